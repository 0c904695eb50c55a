cd $GRAFT_REPO_ROOT
echo "--- stream"
DCVC_B200_GEMM_MODE=stream timeout 600 python -m pytest tests/test_hts_gpu.py -x -q -m gpu -k state_consistency 2>&1 | tail -4
echo "--- resident"
DCVC_B200_GEMM_MODE=resident timeout 600 python -m pytest tests/test_hts_gpu.py -x -q -m gpu -k state_consistency 2>&1 | tail -4
echo "--- auto"
timeout 600 python -m pytest tests/test_hts_gpu.py -x -q -m gpu -k state_consistency 2>&1 | tail -4
