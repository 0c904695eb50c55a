#!/bin/bash
# compute-sanitizer memcheck over the small op tests (one B200; the tcgen05 / TMA kernels run under the sanitizer too)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_ops_gpu.py -x -q -m gpu \
    -k "(pw_gemm and (16-16 or 32-32 or 17-30)) or (dw3x3 and 16-16) or (frame and 16-16) or (warp and 16-24) or (gdn and 16-16) or (conv3x3 and 32-32) or (tconv and 16-16)" \
    > gpurun_out/sanitizer_memcheck.log 2>&1
tail -15 gpurun_out/sanitizer_memcheck.log
