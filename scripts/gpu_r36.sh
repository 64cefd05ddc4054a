cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "res_ln or bwd_input or linear_fwd" > gpurun_out/r36_pytest.log 2>&1; tail -15 gpurun_out/r36_pytest.log | cut -c1-200
timeout 280 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_tc.py -x -q -k "test_tc_linear_bwd_input and 60001 and relu_mask or test_tc_linear_res_ln and 60001" > gpurun_out/r36_sanitizer.log 2>&1; grep -v "^$" gpurun_out/r36_sanitizer.log | head -60 | cut -c1-220
