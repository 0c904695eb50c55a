#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
echo "== dcb_tail tests"
timeout 600 python -m pytest tests/test_dcb_tail_gpu.py -q -x 2>&1 | tail -4
echo "== micro (dbg 16 = old cluster-scope signalling)"
timeout 120 python tools/dcb_tail_micro.py 136 240 384 384 384 2>&1 | tail -2
for DBG in 16 2 8 10; do
  DCVC_B200_GEMM_DBG=$DBG timeout 120 python tools/dcb_tail_micro.py 136 240 384 384 384 fused 2>&1 | tail -1
done
timeout 120 python tools/dcb_tail_micro.py 68 120 512 512 512 2>&1 | tail -2
echo "== trace"
timeout 200 python tools/dcb_tail_trace.py 136 240 384 384 384 2>&1 | tail -48 | head -30
echo "== codec diag 4K"
timeout 600 python tools/diag_fuse.py 2160 3840 40 2>&1 | tail -7 | cut -c1-150
