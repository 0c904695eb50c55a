cd $GRAFT_REPO_ROOT
# (1) launch list of the bench command (graph kernel nodes are profiled individually)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1f.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-hts > gpurun_out/launches_r1f.log 2>&1
tail -c 300 gpurun_out/launches_r1f.log
# (2) DRAM traffic + time of every pw_gemm launch of one encode + one decode
DCVC_B200_GRAPHS=0 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none \
    -k regex:pw_gemm -c 400 --csv --log-file gpurun_out/traffic_r1f.csv python tools/profile_decode.py 1080 1920 1 > gpurun_out/traffic_r1f.log 2>&1
tail -c 200 gpurun_out/traffic_r1f.log
# (3) full capture of 6 decoder GEMMs at P8
DCVC_B200_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pw_gemm -s 150 -c 6 -f -o gpurun_out/prof_gemm_r1f \
    python tools/profile_decode.py 1080 1920 1 > gpurun_out/prof_r1f.log 2>&1
tail -c 200 gpurun_out/prof_r1f.log
ls -la gpurun_out/
