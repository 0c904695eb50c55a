cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
echo "--- micro"
timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 384 0 0 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 512 0 0 1 2>&1 | tail -1
timeout 120 python tools/gemm_micro.py 68 120 512 2048 1 1 2>&1 | tail -1
echo "--- micro no PDL"
DCVC_B200_PDL=0 timeout 120 python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -1
DCVC_B200_PDL=0 timeout 120 python tools/gemm_micro.py 136 240 384 1536 1 1 2>&1 | tail -1
timeout 120 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -10 | head -3
echo "--- bench"
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1
