cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -11
echo "--- micro"
timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 384 0 0 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 512 0 0 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 2048 1 1 2>&1 | tail -1
echo "--- resident sb2 bn128"
export DCVC_B200_GEMM_MODE=resident DCVC_B200_GEMM_STAGING=2 DCVC_B200_GEMM_BN=128
timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 384 0 0 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 2048 1 1 2>&1 | tail -1
timeout 120 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -11
