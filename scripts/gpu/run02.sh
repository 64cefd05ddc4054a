cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -q > gpurun_out/r02_fused.log 2>&1; tail -12 gpurun_out/r02_fused.log | cut -c1-400
timeout 100 python -m pytest "tests/test_gpu_parity.py::test_seq_encoder_dropout_given_mask" -q 2>&1 | tail -3 | cut -c1-300
MVN_FFN_FWD=0 python scripts/bench_fused.py ffn 2>&1 | tail -3
MVN_FFN_FWD=1 python scripts/bench_fused.py ffn 2>&1 | tail -3
