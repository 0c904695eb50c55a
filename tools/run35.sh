cd $GRAFT_REPO_ROOT
echo "--- op tests (streaming)"
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "--- op tests (pair)"
DCVC_B200_GEMM_ARES=1 timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "pw_gemm" 2>&1 | tail -3
echo "--- micro (CUDA graph): streaming / ares1 / pair"
for a in "136 240 384 384" "136 240 384 384 1" "136 240 384 384 0 0 1" "136 240 384 1536 1 1" "68 120 512 512 1 0 0" "68 120 512 512 0 0 1" "136 240 256 512 0 0 1"; do
DCVC_B200_GEMM_ARES=0 timeout 60 python tools/gemm_micro.py $a 2>&1 | tail -1
DCVC_B200_GEMM_ARES=1 DCVC_B200_GEMM_PAIR=0 timeout 60 python tools/gemm_micro.py $a 2>&1 | tail -1
DCVC_B200_GEMM_ARES=1 timeout 60 python tools/gemm_micro.py $a 2>&1 | tail -1
done
echo "=== ares1 384x384 res"
DCVC_B200_GEMM_ARES=1 DCVC_B200_GEMM_PAIR=0 timeout 60 python tools/gemm_trace.py 136 240 384 384 0 0 1 2>&1 | tail -3
echo "=== ares1 384x384 plain"
DCVC_B200_GEMM_ARES=1 DCVC_B200_GEMM_PAIR=0 timeout 60 python tools/gemm_trace.py 136 240 384 384 2>&1 | tail -3
