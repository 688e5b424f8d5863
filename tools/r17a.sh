set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r17a_pytest.log
python bench.py > gpurun_out/r17a_bench.log 2>&1
for l in 16 18 22 24; do timeout 600 python bench.py --log2n $l --steps 5 --no-cpu-baseline >> gpurun_out/r17a_sizes.log 2>&1; done
python bench.py --group g2 --steps 10 --no-cpu-baseline > gpurun_out/r17a_bench_g2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r17a_launches_2p20.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r17a_ncu20.log 2>&1
