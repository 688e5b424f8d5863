set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r19a_bench.log 2>&1
python bench.py --group g2 --steps 10 --no-cpu-baseline > gpurun_out/r19a_bench_g2.log 2>&1
for l in 16 18 22 24; do timeout 600 python bench.py --log2n $l --steps 5 --no-cpu-baseline >> gpurun_out/r19a_sizes.log 2>&1; done
