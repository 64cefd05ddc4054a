set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd|gemm_kernel|wgrad' -s 200 -c 6 -o gpurun_out/r1_prof python bench.py --steps 1 --warmup 1 --batch 256 --no-cpu-baseline > gpurun_out/r1_ncu.log 2>&1
tail -3 gpurun_out/r1_pytest.log; cat gpurun_out/r1_bench.json | head -c 1500
