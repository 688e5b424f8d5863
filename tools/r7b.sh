set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r7b_pytest.log
timeout 1200 python tools/fr_bench.py > gpurun_out/r7b_fr_bench.log 2>&1
I=integration/_ref
{
for a in "fft_b200 16" "fft_cpuomp 16" "fft_b200 20 786432" "fft_cpuomp 20 786432" "fft_cpu 18" "fft_b200 18" "fft_b200 22" "fft_cpuomp 22" \
         "polycommit_b200 16" "polycommit_b200 20" "polycommit_cpuomp 20" "cplink_b200 10 5" \
         "groth16matrix_b200 16" "groth16matrix_b200 32" "groth16matrix_b200 64 0" "groth16matrix_cpuomp 64" "groth16matrix_b200 128 0"; do
  echo "== $a"; ( time B200_GPUS=1 timeout 900 $I/$a ) 2>&1 | grep -E '^\{|real|rror|terminate|what|fault'
done
} > gpurun_out/r7b_integration.log 2>&1
