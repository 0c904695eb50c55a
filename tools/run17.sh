cd $GRAFT_REPO_ROOT
timeout 600 python tools/stress_ops.py 2>&1 | tail -40
