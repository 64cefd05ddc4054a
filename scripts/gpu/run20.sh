cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in 1; do
MVN_CONV_FUSED=$f timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r20_launches_c3_$f.csv python bench.py --steps 1 --warmup 3 --workload c3 --no-cpu-baseline --no-sweep --no-graph > /dev/null 2>&1
done
