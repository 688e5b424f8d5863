set -x
for mb in 0 3 4; do B200_ACC_MINB=$mb python bench.py --group g2 --steps 5 --no-cpu-baseline > gpurun_out/r2c_g2_minb$mb.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c_launches_2p20.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu20.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c_launches_2p24.csv python bench.py --log2n 24 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu24.log 2>&1
