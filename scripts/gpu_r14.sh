cd $GRAFT_REPO_ROOT
MVN_PDL=1 python scripts/enqueue_time.py 2>&1 | tail -1
MVN_PDL=0 python scripts/enqueue_time.py 2>&1 | tail -1
