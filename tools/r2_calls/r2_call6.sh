#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
echo "== dcb_tail tests (v2 kernel)"
timeout 600 python -m pytest tests/test_dcb_tail_gpu.py -q -x 2>&1 | tail -12
echo "== micro"
for V in "0:0" "3:0" "2:0" "1:0" "0:1" "3:1"; do
  IFS=: read DBG KBS <<< "$V"
  DCVC_B200_DT_KBS=$KBS DCVC_B200_GEMM_DBG=$DBG timeout 120 python tools/dcb_tail_micro.py 136 240 384 384 384 fused 2>&1 | tail -1
done
timeout 120 python tools/dcb_tail_micro.py 68 120 512 512 512 2>&1 | tail -2
DCVC_B200_GEMM_DBG=3 timeout 120 python tools/dcb_tail_micro.py 68 120 512 512 512 fused 2>&1 | tail -1
echo "== codec diag"
timeout 300 python tools/diag_fuse.py 1080 1920 32 2>&1 | tail -8
timeout 600 python tools/diag_fuse.py 2160 3840 40 2>&1 | tail -8
