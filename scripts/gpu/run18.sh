cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do
echo NEW; python scripts/bench_fused.py attn 2>&1 | tail -2
echo HEAD; MVN_LIB_PATH=$GRAFT_REPO_ROOT/multimodal-supernovae_b200/libmaven_head.so python scripts/bench_fused.py attn 2>&1 | tail -2
done
