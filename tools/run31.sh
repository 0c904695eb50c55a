cd $GRAFT_REPO_ROOT
echo "=== streaming 384x384"
DCVC_B200_GEMM_ARES=0 timeout 60 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -12
echo "=== pair 384x384"
timeout 60 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -12 | head -8
echo "=== pair 384x384 no-epilogue (dbg 2)"
DCVC_B200_GEMM_DBG=2 timeout 60 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -12 | head -8
echo "=== pair 384x384 loads only (dbg 3)"
DCVC_B200_GEMM_DBG=3 timeout 60 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -12 | head -8
echo "=== streaming loads only (dbg 3)"
DCVC_B200_GEMM_ARES=0 DCVC_B200_GEMM_DBG=3 timeout 60 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -12 | head -8
echo "=== streaming no epilogue (dbg 2)"
DCVC_B200_GEMM_ARES=0 DCVC_B200_GEMM_DBG=2 timeout 60 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -12 | head -8
