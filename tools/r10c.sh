set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r10c_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r10c_smoke.log 2>&1
python bench.py > gpurun_out/r10c_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r10c_bench_ref.log 2>&1
for l in 16 18; do timeout 600 python bench.py --log2n $l --steps 10 --no-cpu-baseline >> gpurun_out/r10c_sizes.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r10c_launches_2p20.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r10c_ncu20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 6 -c 1 -f -o gpurun_out/r10c_acc_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r10c_ncufull.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fr_fft_pass -s 6 -c 3 -f -o gpurun_out/r10c_fft_full python tools/fr_bench.py --fft 20 --fold "" --prove "" --no-cpu --reps 2 > gpurun_out/r10c_ncufft.log 2>&1
