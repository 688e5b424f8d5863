set -x
mkdir -p gpurun_out
I=integration/_ref
{
for a in "polycommit_b200 16" "polycommit_b200 20" "cplink_b200 10 5"; do
  echo "== $a"; ( time B200_GPUS=1 timeout 900 $I/$a ) 2>&1 | grep -E '^\{|real|rror|terminate|what|fault'
done
} > gpurun_out/r20a_integration.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r20a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r20a_smoke.log 2>&1
