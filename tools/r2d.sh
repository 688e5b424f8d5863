set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/r2d_bench.log 2>&1
for l in 16 18 22 24 26; do python bench.py --log2n $l --steps 5 --no-cpu-baseline >> gpurun_out/r2d_bench_sizes.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2d_launches_2p20.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu20.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2d_launches_2p24.csv python bench.py --log2n 24 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu24.log 2>&1
