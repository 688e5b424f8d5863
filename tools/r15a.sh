set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r15a_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r15a_bench.log 2>&1
for l in 16 18 24; do timeout 600 python bench.py --log2n $l --steps 10 --no-cpu-baseline >> gpurun_out/r15a_sizes.log 2>&1; done
python bench.py --group g2 --steps 10 --no-cpu-baseline > gpurun_out/r15a_bench_g2.log 2>&1
