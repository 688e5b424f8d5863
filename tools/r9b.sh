set -x
mkdir -p gpurun_out
nproc > gpurun_out/r9b_probe.log; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/r9b_probe.log
for t in 0 1 2 4 8; do B200_COPY_THREADS=$t python tools/copy_probe.py 22 >> gpurun_out/r9b_probe.log 2>&1; done
