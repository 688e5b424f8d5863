set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r16a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r16a_smoke.log 2>&1
python bench.py > gpurun_out/r16a_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r16a_bench_ref.log 2>&1
python bench.py --group g2 --steps 10 --no-cpu-baseline > gpurun_out/r16a_bench_g2.log 2>&1
for l in 16 24; do timeout 600 python bench.py --log2n $l --steps 5 --no-cpu-baseline >> gpurun_out/r16a_sizes.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r16a_launches_2p20.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r16a_ncu20.log 2>&1
I=integration/_ref
{
  echo "== $a"; ( time B200_GPUS=1 timeout 900 $I/$a ) 2>&1 | grep -E '^\{|real|rror|terminate|what|fault'
done
} > gpurun_out/r16a_integration.log 2>&1
for l in 18 22 26; do timeout 900 python bench.py --log2n $l --steps 3 --no-cpu-baseline >> gpurun_out/r16a_sizes.log 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 6 -c 1 -f -o gpurun_out/r16a_acc_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r16a_ncufull.log 2>&1
