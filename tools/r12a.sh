set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r12a_pytest.log
timeout 600 python tools/wire_bench.py > gpurun_out/r12a_wire_bench.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r12a_bench_torchrun1.log 2>&1
I=integration/_ref
{
for a in "polycommit_b200 16" "polycommit_b200 20" "polycommit_cpuomp 16" "cplink_b200 10 5" "fft_b200 20" "groth16matrix_b200 32" "groth16matrix_b200 64 0"; do
  echo "== $a"; ( time B200_GPUS=1 timeout 900 $I/$a ) 2>&1 | grep -E '^\{|real|rror|terminate|what|fault'
done
} > gpurun_out/r12a_integration.log 2>&1
