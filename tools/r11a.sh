set -x
mkdir -p gpurun_out
python bench.py --group g2 --steps 10 --no-cpu-baseline > gpurun_out/r11a_bench_g2.log 2>&1
python -m pytest tests/test_gpu_msm.py -m gpu -x -q -k "g2" 2>&1 | tail -3 > gpurun_out/r11a_pytest.log
