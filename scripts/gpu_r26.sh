cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/host_profile.py c3 > gpurun_out/r26_host_c3.txt 2>&1
python scripts/host_profile.py c4 > gpurun_out/r26_host_c4.txt 2>&1
head -5 gpurun_out/r26_host_c3.txt; grep enqueue gpurun_out/r26_host_c4.txt
