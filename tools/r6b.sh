set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r6b_pytest.log
timeout 900 python tools/pre_sweep.py --log2n 20 --cs 0,16,17,18,19,20 > gpurun_out/r6b_sweep20.log 2>&1
timeout 900 python tools/pre_sweep.py --log2n 24 --cs 0,19,20,22 --reps 4 > gpurun_out/r6b_sweep24.log 2>&1
python bench.py > gpurun_out/r6b_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r6b_bench_ref.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r6b_smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r6b_launches_2p20.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r6b_ncu20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 6 -c 1 -f -o gpurun_out/r6b_acc_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6b_ncufull.log 2>&1
python bench.py --group g2 --steps 10 --no-cpu-baseline > gpurun_out/r6b_bench_g2.log 2>&1
timeout 600 python bench.py --log2n 24 --steps 5 --no-cpu-baseline > gpurun_out/r6b_bench_24.log 2>&1
