python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for ls in 2 3 4 5 6; do echo "logS $ls"; B200_LOGS=$ls python bench.py --steps 10 --no-cpu-baseline 2>&1 | python tools/summ.py bench /dev/stdin; done
for ls in 3 4 5 6 7; do echo "logS $ls"; B200_LOGS=$ls python bench.py --log2n 24 --steps 3 --no-cpu-baseline 2>&1 | python tools/summ.py bench /dev/stdin; done
echo auto; python bench.py --steps 10 --no-cpu-baseline 2>&1 | python tools/summ.py bench /dev/stdin
python bench.py --log2n 24 --steps 3 --no-cpu-baseline 2>&1 | python tools/summ.py bench /dev/stdin
