set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r14b_racecheck.log python -m pytest tests -m gpu -x -q -k "fold_vs_oracle or fft_vs_oracle or cplink_shape or golden_cases or hot_buckets" > gpurun_out/r14b_pytest.log 2>&1
echo "exit=$?" >> gpurun_out/r14b_pytest.log
