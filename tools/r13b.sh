set -x
mkdir -p gpurun_out
python -m pytest tests/test_fr_vectors.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r13b_pytest.log
timeout 900 python tools/fr_bench.py --fft 16,20,22,24 --fold "" --prove "" --no-cpu > gpurun_out/r13b_fr_bench.log 2>&1
