set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r4b_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r4b_bench.log 2>&1
python bench.py --group g2 --steps 10 --no-cpu-baseline > gpurun_out/r4b_bench_g2.log 2>&1
for l in 24; do timeout 600 python bench.py --log2n $l --steps 5 --no-cpu-baseline >> gpurun_out/r4b_bench_sizes.log 2>&1; done
