set -x
mkdir -p gpurun_out
timeout 600 python tools/pre_sweep.py --log2n 20 --cs 20 --logS 2,3,4,5 --splits 0,8,16,32 > gpurun_out/r15b_sweep20.log 2>&1
timeout 600 python tools/pre_sweep.py --log2n 16 --cs 17 --logS 0,1,2,3,4 --splits 0,4,8,16 > gpurun_out/r15b_sweep16.log 2>&1
