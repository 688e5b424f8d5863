set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r10b_pytest.log
timeout 900 python tools/fr_bench.py --fft 16,20,22,24 --fold 20 --prove 20 > gpurun_out/r10b_fr_bench.log 2>&1
