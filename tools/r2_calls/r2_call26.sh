#!/bin/bash
# after removing the finished experiments' knobs (kernel sources touched): op, fused-kernel and codec tests + a micro number
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_dcb_tail_gpu.py tests/test_ops_gpu.py tests/test_codec_gpu.py tests/test_hts_gpu.py -m gpu -x -q > $O/r2c26_pytest.log 2>&1; echo "rc=$?"; tail -2 $O/r2c26_pytest.log
timeout 60 python tools/dcb_tail_micro.py 136 240 384 384 384 2>&1 | tail -2
