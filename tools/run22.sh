cd $GRAFT_REPO_ROOT
export DCVC_B200_GEMM_MODE=resident DCVC_B200_GEMM_STAGING=2
for bn in 128 64; do
export DCVC_B200_GEMM_BN=$bn
echo "--- resident sb=2 bn=$bn"
timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 384 0 0 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 512 0 0 1 2>&1 | tail -1
done
export DCVC_B200_GEMM_BN=128
timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 2048 1 1 2>&1 | tail -1
timeout 120 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -8
timeout 120 python tools/gemm_trace.py 136 240 384 1536 1 1 2>&1 | tail -8
