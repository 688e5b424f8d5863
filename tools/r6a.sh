set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r6a_pytest.log
timeout 900 python tools/pre_sweep.py --log2n 20 --cs 0,16,17,18,19,20,21 > gpurun_out/r6a_sweep20.log 2>&1
timeout 600 python tools/pre_sweep.py --log2n 20 --cs 19,20 --logS 0,1,2,3,4,5 --splits 0,8,32 > gpurun_out/r6a_sweep20_red.log 2>&1
timeout 900 python tools/pre_sweep.py --log2n 24 --cs 0,20,21,22 --reps 4 > gpurun_out/r6a_sweep24.log 2>&1
timeout 600 python tools/pre_sweep.py --log2n 20 --group g2 --cs 0,16,18,19 --reps 4 > gpurun_out/r6a_sweep20_g2.log 2>&1
