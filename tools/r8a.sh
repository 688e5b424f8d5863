set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r8a_pytest.log
timeout 900 python tools/dist_sweep.py --log2n 20 > gpurun_out/r8a_dist20.log 2>&1
python bench.py > gpurun_out/r8a_bench.log 2>&1
