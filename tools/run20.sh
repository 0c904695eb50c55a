cd $GRAFT_REPO_ROOT
timeout 120 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -9
timeout 120 python tools/gemm_trace.py 136 240 384 384 0 0 1 2>&1 | tail -9
timeout 120 python tools/gemm_trace.py 136 240 384 1536 1 1 2>&1 | tail -9
timeout 120 python tools/gemm_trace.py 68 120 512 512 0 0 1 2>&1 | tail -9
