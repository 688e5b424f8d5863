set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r3b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3b_smoke.log 2>&1
I=integration/_ref
{
for a in "cplink_b200 10 5" "cplink_cpu 10 3" "cplink_cpuomp 10 3" "cplink_b200 14 3" "cplink_cpuomp 14 1" \
         "polycommit_b200 16" "polycommit_cpuomp 16" "polycommit_b200 20" "polycommit_cpuomp 20" \
         "groth16matrix_b200 16" "groth16matrix_cpuomp 16" "groth16matrix_b200 32" "groth16matrix_b200 64 0" "groth16matrix_cpuomp 64" \
         "groth16matrix_b200 128 0" "groth16matrix_cpuomp 128"; do
  echo "== $a"; ( time timeout 900 $I/$a ) 2>&1 | grep -E '^\{|real|rror|terminate|what' 
done
} > gpurun_out/r3b_integration.log 2>&1
nproc >> gpurun_out/r3b_integration.log
