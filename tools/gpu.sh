#!/bin/bash
# tools/gpu.sh — the one parametrised runner for gpurun (replaces the per-round tools/rNN.sh scripts).
#   gpurun -- 'bash tools/gpu.sh <tag> <step> [<step> ...]'
# Every step writes gpurun_out/<tag>_<step>.log (or .csv / .jsonl).  Steps:
#   tests [pytest args]   python -m pytest tests -m gpu -x -q
#   smoke                 __graft_entry__.smoke()
#   bench[:args]          python bench.py <args>   (args separated by '+', a '~' inside an argument is a space:
#                         bench:--log2n+22   tests:tests/test_gpu_msm.py+-k+golden~or~uniform)
#   ref[:args]            python bench.py --impl reference <args>
#   stage:<args>          python tools/stage_times.py <args>
#   py:<script>+<args>    python tools/<script>.py <args>
#   launches[:args]       ncu launch list (gpu__time_duration) of python bench.py --steps 2 --warmup 3 --no-cpu-baseline <args>
#   ncufull:<regex>[+args] ncu --set full of the kernels matching <regex> in the same command; also writes the raw/details CSV pages
#   integ:<binary>+<args> integration/_ref/<binary> <args>
#   sanitize:<tool>+<pytest args>  compute-sanitizer --tool <memcheck|racecheck|synccheck> over python -m pytest -m gpu <args>
#   mgpu:<N>[+args]       torchrun with N ranks on one node: python bench.py --gpus N <args> (needs gpurun --gpus N)
set -u
TAG=$1; shift
mkdir -p gpurun_out
O=gpurun_out/$TAG
for step in "$@"; do
  name=${step%%:*}; rest=""; [[ "$step" == *:* ]] && rest=${step#*:}
  IFS='+' read -ra A <<< "$rest"; A=("${A[@]//\~/ }")
  case $name in
    tests)   timeout 3000 python -m pytest tests -m gpu -x -q "${A[@]}" 2>&1 | tail -15 > ${O}_tests.log ;;
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1 ;;
    bench)   timeout 1200 python bench.py "${A[@]}" >> ${O}_bench.jsonl 2>> ${O}_bench.err ;;
    ref)     timeout 1200 python bench.py --impl reference "${A[@]}" >> ${O}_bench_ref.jsonl 2>> ${O}_bench.err ;;
    stage)   timeout 1200 python tools/stage_times.py "${A[@]}" >> ${O}_stage.jsonl 2>> ${O}_stage.err ;;
    py)      script=${A[0]}
             timeout 1800 python tools/$script.py "${A[@]:1}" >> ${O}_$script.log 2>&1 ;;
    launches) timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches.csv \
               python bench.py --steps 2 --warmup 3 --no-cpu-baseline "${A[@]}" > ${O}_launches_run.log 2>&1 ;;
    ncufull) regex=${A[0]}
             short=$(echo "$regex" | tr -cd 'a-zA-Z0-9_')
             timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$regex -s ${NCU_SKIP:-4} -c ${NCU_COUNT:-4} -f -o ${O}_ncu_$short \
               python bench.py --steps 2 --warmup 3 --no-cpu-baseline "${A[@]:1}" > ${O}_ncu_${short}_run.log 2>&1
             ncu -i ${O}_ncu_$short.ncu-rep --page raw --csv > ${O}_ncu_${short}_raw.csv 2>/dev/null
             ncu -i ${O}_ncu_$short.ncu-rep --page details --csv > ${O}_ncu_${short}_details.csv 2>/dev/null ;;
    integ)   ( time B200_GPUS=1 timeout 1500 integration/_ref/"${A[0]}" "${A[@]:1}" ) 2>&1 | grep -E '^\{|^\[b200\]|leave|real|rror|terminate|what|fault' >> ${O}_integration.log ;;
    sanitize) tool=${A[0]}
             ( timeout 2400 compute-sanitizer --tool $tool --print-limit 20 python -m pytest -m gpu -x -q "${A[@]:1}" 2>&1 | tail -25; echo "exit $?" ) > ${O}_compute_sanitizer_$tool.log ;;
    mgpu)    N=${A[0]}
             timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
               bench.py --gpus $N "${A[@]:1}" >> ${O}_bench_n$N.jsonl 2>> ${O}_bench_n$N.err ;;
    *) echo "unknown step $step" >> ${O}_errors.log ;;
  esac
done
ls -la gpurun_out | tail -30 > ${O}_files.log
