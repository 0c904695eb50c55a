cd $GRAFT_REPO_ROOT
echo "--- ares single-CTA op tests"
DCVC_B200_GEMM_PAIR=0 timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "pw_gemm" 2>&1 | tail -5
echo "--- ares pair op tests"
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "pw_gemm" 2>&1 | tail -5
echo "--- micro: streaming / ares1 / pair"
for a in "136 240 384 384" "136 240 384 384 0 0 1" "136 240 384 1536 1 1" "68 120 512 512 1 0 0" "68 120 512 2048 1 1" "136 240 512 256 1" "136 240 256 512 0 0 1"; do
DCVC_B200_GEMM_ARES=0 timeout 60 python tools/gemm_micro.py $a 2>&1 | tail -1
DCVC_B200_GEMM_PAIR=0 timeout 60 python tools/gemm_micro.py $a 2>&1 | tail -1
timeout 60 python tools/gemm_micro.py $a 2>&1 | tail -1
done
