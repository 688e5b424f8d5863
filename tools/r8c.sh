set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r8c_gpus.log
python tools/multi_gpu_check.py 2 20 > gpurun_out/r8c_inproc.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r8c_bench_n2.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r8c_bench_ref_n2.log 2>&1
