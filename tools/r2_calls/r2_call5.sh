#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python tools/diag_fuse.py 2160 3840 40 2>&1 | tail -12
timeout 300 python tools/diag_fuse.py 1080 1920 32 2>&1 | tail -12
