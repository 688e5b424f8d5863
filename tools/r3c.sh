set -x
mkdir -p gpurun_out
I=integration/_ref
( $I/groth16matrix_b200 16 ) > gpurun_out/r3c_g16.log 2>&1; echo "rc=$?" >> gpurun_out/r3c_g16.log
{
for a in "groth16matrix_b200 32" "groth16matrix_b200 64 0" "groth16matrix_b200 128 0" "cplink_b200 10 5" "polycommit_b200 20"; do
  echo "== $a"; ( time timeout 900 $I/$a ) 2>&1 | grep -E '^\{|real|rror|terminate|what|fault'
done
} > gpurun_out/r3c_integration.log 2>&1
