set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r22a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r22a_smoke.log 2>&1
python bench.py > gpurun_out/r22a_bench.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r22a_bench_ref.log 2>&1
