set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r14a_memcheck.log python -m pytest tests -m gpu -x -q -k "golden or ones_filter or precomputed_key or fold_vs_oracle or fft_vs_oracle or cppoly or wire_vs_oracle or pinned_key_prefixes or knowledge" > gpurun_out/r14a_pytest.log 2>&1
echo "exit=$?" >> gpurun_out/r14a_pytest.log
tail -5 gpurun_out/r14a_memcheck.log >> gpurun_out/r14a_pytest.log
