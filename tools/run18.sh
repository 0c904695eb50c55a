cd $GRAFT_REPO_ROOT
timeout 600 python tools/diag_opsum.py 1080 1920 2>&1 | tail -40
