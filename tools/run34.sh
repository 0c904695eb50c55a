cd $GRAFT_REPO_ROOT
echo "=== ares1 384x384 plain"
DCVC_B200_GEMM_ARES=1 DCVC_B200_GEMM_PAIR=0 timeout 60 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -9
echo "=== ares1 384x384 res"
DCVC_B200_GEMM_ARES=1 DCVC_B200_GEMM_PAIR=0 timeout 60 python tools/gemm_trace.py 136 240 384 384 0 0 1 2>&1 | tail -9
echo "=== ares1 384x384 wsilu"
DCVC_B200_GEMM_ARES=1 DCVC_B200_GEMM_PAIR=0 timeout 60 python tools/gemm_trace.py 136 240 384 384 1 2>&1 | tail -9
