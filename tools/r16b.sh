set -x
mkdir -p gpurun_out
I=integration/_ref
{
for a in "polycommit_b200 20" "cplink_b200 10 5" "groth16matrix_b200 64 0"; do
  echo "== $a"; ( time B200_GPUS=1 timeout 900 $I/$a ) 2>&1 | grep -E '^\{|real|rror|terminate|what|fault'
done
} > gpurun_out/r16b_integration.log 2>&1
for l in 18 20 22 26; do timeout 900 python bench.py --log2n $l --steps 3 --no-cpu-baseline >> gpurun_out/r16b_sizes.log 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 6 -c 1 -f -o gpurun_out/r16b_acc_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r16b_ncufull.log 2>&1
timeout 600 python tools/fr_bench.py --fft 20 --fold 20 --prove 16,20 > gpurun_out/r16b_fr_bench.log 2>&1
