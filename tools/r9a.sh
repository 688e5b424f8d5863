set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r9a_pytest.log
timeout 900 python tools/fr_bench.py --fft 20,22 --fold 20 --prove 20 > gpurun_out/r9a_fr_bench.log 2>&1
timeout 600 python tools/sweep.py --log2n 16,18,20,22 > gpurun_out/r9a_sweep.log 2>&1
I=integration/_ref
{
for a in "fft_b200 20" "fft_b200 22" "polycommit_b200 20" "cplink_b200 10 5" "groth16matrix_b200 64 0"; do
  echo "== $a"; ( time B200_GPUS=1 timeout 900 $I/$a ) 2>&1 | grep -E '^\{|real|rror|terminate|what|fault'
done
} > gpurun_out/r9a_integration.log 2>&1
python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r9a_bench.log 2>&1
