cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/r41_pytest.log 2>&1; tail -3 gpurun_out/r41_pytest.log | cut -c1-200
