set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r8b_pytest.log
timeout 900 python tools/dist_sweep.py --log2n 20 > gpurun_out/r8b_dist20.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/r8b_bench.log 2>&1
I=integration/_ref
{
for a in "fft_b200 20 786432" "fft_b200 22" "groth16matrix_b200 64 0" "groth16matrix_b200 32"; do
  echo "== $a"; ( time B200_GPUS=1 timeout 900 $I/$a ) 2>&1 | grep -E '^\{|real|rror|terminate|what|fault'
done
} > gpurun_out/r8b_integration.log 2>&1
