set -x
mkdir -p gpurun_out
rm -f gpurun_out/r9c_sizes.log
for l in 16 18 20 22 24; do timeout 600 python bench.py --log2n $l --steps 10 --no-cpu-baseline >> gpurun_out/r9c_sizes.log 2>&1; done
timeout 900 python bench.py --log2n 26 --steps 3 --no-cpu-baseline >> gpurun_out/r9c_sizes.log 2>&1
for l in 20 22 24; do timeout 900 python bench.py --group g2 --log2n $l --steps 5 --no-cpu-baseline >> gpurun_out/r9c_sizes_g2.log 2>&1; done
