#!/bin/bash
# the dcb_tail captures of tools/profile_round.sh again (the first attempt skipped past all of its 20 launches)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out; R=r2; K=dcb_tail
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none \
    -k regex:${K}_kernel -s 7 -c 13 --csv --log-file $O/${R}_traffic_${K}.csv python tools/profile_decode.py 1080 1920 1 > /dev/null 2>&1
echo "traffic rows ${K}: $(wc -l < $O/${R}_traffic_${K}.csv)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 9 -c 3 -f -o $O/${R}_prof_${K} \
    python tools/profile_decode.py 1080 1920 1 > /dev/null 2>&1
timeout 300 ncu -i $O/${R}_prof_${K}.ncu-rep --page raw --csv > $O/${R}_prof_${K}_raw.csv 2>/dev/null
echo "set-full raw rows ${K}: $(wc -l < $O/${R}_prof_${K}_raw.csv)"
ls -la $O/${R}_prof_${K}.ncu-rep
echo "== 2-GPU bench (weak scaling headline + the sharded configs[3] job list)"
