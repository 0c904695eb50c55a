cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# source-level profile of the streaming kernel (plain and +residual) and the pair kernel, M=32640 N=K=384
DCVC_B200_GEMM_ARES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_gemm -s 4 -c 1 -f -o gpurun_out/prof_stream_plain python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -2
DCVC_B200_GEMM_ARES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_gemm -s 4 -c 1 -f -o gpurun_out/prof_stream_res python tools/gemm_micro.py 136 240 384 384 0 0 1 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_gemm -s 4 -c 1 -f -o gpurun_out/prof_pair_plain python tools/gemm_micro.py 136 240 384 384 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
