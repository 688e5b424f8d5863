set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.log 2>&1
python bench.py --group g2 --steps 10 > gpurun_out/r2a_bench_g2.log 2>&1
for l in 16 18 22 24 26; do timeout 600 python bench.py --log2n $l --steps 5 --no-cpu-baseline >> gpurun_out/r2a_bench_sizes.log 2>&1; done
for l in 22 24; do timeout 600 python bench.py --group g2 --log2n $l --steps 3 --no-cpu-baseline >> gpurun_out/r2a_bench_sizes_g2.log 2>&1; done
nvidia-smi > gpurun_out/r2a_smi.log
