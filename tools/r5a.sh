set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r5a_pytest.log
python bench.py > gpurun_out/r5a_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5a_bench_ref.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5a_smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r5a_launches_2p20.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r5a_ncu20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -c 1 -f -o gpurun_out/r5a_acc_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r5a_ncufull.log 2>&1
