cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "--- auto"
timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 384 0 0 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 512 0 0 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 2048 1 1 2>&1 | tail -1
echo "--- stream"
DCVC_B200_GEMM_MODE=stream timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
DCVC_B200_GEMM_MODE=stream timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
echo "--- resident"
DCVC_B200_GEMM_MODE=resident timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
DCVC_B200_GEMM_MODE=resident timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
DCVC_B200_GEMM_MODE=resident DCVC_B200_GEMM_BN=128 timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
DCVC_B200_GEMM_MODE=resident DCVC_B200_GEMM_BN=192 timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
echo "--- bench"
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3
