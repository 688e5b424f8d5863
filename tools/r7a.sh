set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r7a_pytest.log
timeout 1200 python tools/fr_bench.py > gpurun_out/r7a_fr_bench.log 2>&1
