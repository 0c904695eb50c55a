cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "--- auto"
timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 384 0 0 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 512 0 0 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 2048 1 1 2>&1 | tail -1
echo "--- no cluster"
DCVC_B200_GEMM_CLUSTER=1 timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
DCVC_B200_GEMM_CLUSTER=1 timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
echo "--- variants"
DCVC_B200_GEMM_BN=192 timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
DCVC_B200_GEMM_BN=128 timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
DCVC_B200_GEMM_BN=256 timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
DCVC_B200_GEMM_BN=128 timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
DCVC_B200_GEMM_DBG=3 timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
DCVC_B200_GEMM_DBG=3 timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
DCVC_B200_GEMM_DBG=2 timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
DCVC_B200_GEMM_DBG=2 timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
