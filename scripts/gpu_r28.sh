cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/host_profile.py c3 2>&1 | head -3 > gpurun_out/r28_host.txt
python scripts/host_profile.py c4 2>&1 | head -3 >> gpurun_out/r28_host.txt
cat gpurun_out/r28_host.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 84 -c 4 -o gpurun_out/r28_attn python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r28_ncu.log 2>&1
tail -3 gpurun_out/r28_ncu.log
ls -la gpurun_out/*.ncu-rep
