set -x
mkdir -p gpurun_out
timeout 600 python tools/pre_sweep.py --log2n 16 --cs 0,11,12,13,14,15,16,17,19 > gpurun_out/r9d_sweep16.log 2>&1
timeout 600 python tools/pre_sweep.py --log2n 18 --cs 0,13,14,15,16,17,19,20 > gpurun_out/r9d_sweep18.log 2>&1
timeout 600 python tools/pre_sweep.py --log2n 22 --cs 0,17,19,20,21,22 --reps 5 > gpurun_out/r9d_sweep22.log 2>&1
timeout 600 python tools/pre_sweep.py --log2n 20 --cs 0,16,17,18,19,20,21 > gpurun_out/r9d_sweep20.log 2>&1
