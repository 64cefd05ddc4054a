set -x
cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -s 1060 -c 420 --csv --log-file gpurun_out/r10_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r10_ncu1.log 2>&1
tail -2 gpurun_out/r10_ncu1.log | cut -c1-200
wc -l gpurun_out/r10_launches.csv
