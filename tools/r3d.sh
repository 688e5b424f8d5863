set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r3d_gpus.log
python tools/multi_gpu_check.py 2 18 > gpurun_out/r3d_inproc.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3d_bench_n2.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 --log2n 24 > gpurun_out/r3d_bench_n2_2p24.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r3d_bench_ref_n2.log 2>&1
I=integration/_ref
{ for a in "groth16matrix_b200 16" "groth16matrix_b200 32"; do echo "== $a"; ( time timeout 900 $I/$a ) 2>&1 | grep -E '^\{|real|rror|terminate|what|fault'; done; } > gpurun_out/r3d_groth.log 2>&1
