cd $GRAFT_REPO_ROOT
timeout 120 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -10
DCVC_B200_GEMM_DBG=3 timeout 120 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -4
DCVC_B200_GEMM_MODE=resident DCVC_B200_GEMM_STAGING=2 DCVC_B200_GEMM_BN=128 timeout 120 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -10
