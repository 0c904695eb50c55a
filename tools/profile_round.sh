#!/bin/bash
# Round evidence (run under gpurun on one B200): launch list of the bench command, DRAM traffic of every pw_gemm
# launch of one 1080p Intra decode, and one --set full capture of the two dominant pw_gemm instantiations.
# Outputs go to gpurun_out/; summaries are copied into profiles/ by hand (see profiles/README.md).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
R=${1:-r1b}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hts > gpurun_out/${R}_bench_under_ncu.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/${R}_launches_bench.csv)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none \
    -k regex:pw_gemm -s 250 -c 142 --csv --log-file gpurun_out/${R}_traffic_pw_gemm.csv python tools/profile_decode.py 1080 1920 1 > /dev/null 2>&1
echo "traffic rows: $(wc -l < gpurun_out/${R}_traffic_pw_gemm.csv)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_gemm_kernel -s 300 -c 6 -f -o gpurun_out/${R}_prof_pw_gemm \
    python tools/profile_decode.py 1080 1920 1 > /dev/null 2>&1
ls -la gpurun_out/ | grep ${R}
