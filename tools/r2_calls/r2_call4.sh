#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
echo "== diag: 4K P8 shape, and 1080p shape with 8 pairs (16 tiles per pair)"
timeout 200 python tools/dcb_tail_diag.py 270 480 384 384 384 3 2>&1 | tail -8
DCVC_B200_DT_MAXPAIRS=8 timeout 200 python tools/dcb_tail_diag.py 136 240 384 384 384 3 2>&1 | tail -8
DCVC_B200_DT_ROT=0 timeout 200 python tools/dcb_tail_diag.py 270 480 384 384 384 2 2>&1 | tail -6
echo "== load pipeline experiments (fused only)"
for V in "0:1:0" "1:1:0" "0:1:3" "1:1:3" "0:8:3" "0:37:3" "1:8:3" "0:8:0"; do
  IFS=: read ROT WREP DBG <<< "$V"
  DCVC_B200_DT_ROT=$ROT DCVC_B200_DT_WREP=$WREP DCVC_B200_GEMM_DBG=$DBG timeout 120 python tools/dcb_tail_micro.py 136 240 384 384 384 fused 2>&1 | tail -1
done
echo "== P16 shape"
for V in "0:1:0" "1:1:0" "1:1:3" "0:8:3"; do
  IFS=: read ROT WREP DBG <<< "$V"
  DCVC_B200_DT_ROT=$ROT DCVC_B200_DT_WREP=$WREP DCVC_B200_GEMM_DBG=$DBG timeout 120 python tools/dcb_tail_micro.py 68 120 512 512 512 fused 2>&1 | tail -1
done
