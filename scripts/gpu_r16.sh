set -x
cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"attn_|tc_wgrad" -s 60 -c 12 -o gpurun_out/r16_attn_wgrad python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r16_ncu.log 2>&1
tail -3 gpurun_out/r16_ncu.log | cut -c1-300
ls -la gpurun_out/
