set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r11b_gpus.log
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r11b_bench_n$N.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 --log2n 24 > gpurun_out/r11b_bench_n${N}_2p24.log 2>&1
