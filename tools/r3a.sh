set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r3a_pytest.log
python bench.py > gpurun_out/r3a_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3a_bench_ref.log 2>&1
python bench.py --group g2 --steps 10 --no-cpu-baseline > gpurun_out/r3a_bench_g2.log 2>&1
for l in 16 18 22 24 26; do timeout 600 python bench.py --log2n $l --steps 5 --no-cpu-baseline >> gpurun_out/r3a_bench_sizes.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3a_launches_2p20.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3a_ncu20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -c 1 -f -o gpurun_out/r3a_acc_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r3a_ncufull.log 2>&1
nvidia-smi > gpurun_out/r3a_smi.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3a_smoke.log 2>&1
