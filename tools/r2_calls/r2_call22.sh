#!/bin/bash
# after the last kernel changes (two MMA issuers + one ring each, vectorised count scan): the touched tests, then the
# round's ncu evidence again for the final dcb_tail (launch list of the bench command, traffic, one --set full capture)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out; R=r2b
echo "== tests"
timeout 600 python -m pytest tests/test_codec_gpu.py tests/test_ops_gpu.py tests/test_ld_gpu.py -m gpu -q -x > $O/${R}_pytest.log 2>&1; echo "rc=$?"; tail -2 $O/${R}_pytest.log
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hts --no-reference-cuda --no-seq8 --no-pipelined > $O/${R}_bench_under_ncu.log 2>&1
echo "launch list rows: $(wc -l < $O/${R}_launches_bench.csv)"
python tools/summarize_ncu.py $O/${R}_launches_bench.csv > $O/${R}_launches_bench.md 2>&1; head -16 $O/${R}_launches_bench.md
K=dcb_tail
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none \
    -k regex:${K}_kernel -s 7 -c 13 --csv --log-file $O/${R}_traffic_${K}.csv python tools/profile_decode.py 1080 1920 1 > /dev/null 2>&1
echo "traffic rows ${K}: $(wc -l < $O/${R}_traffic_${K}.csv)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 9 -c 3 -f -o $O/${R}_prof_${K} \
    python tools/profile_decode.py 1080 1920 1 > /dev/null 2>&1
timeout 300 ncu -i $O/${R}_prof_${K}.ncu-rep --page raw --csv > $O/${R}_prof_${K}_raw.csv 2>/dev/null
echo "set-full raw rows ${K}: $(wc -l < $O/${R}_prof_${K}_raw.csv)"
python tools/summarize_set_full.py $O/${R}_prof_${K}_raw.csv > $O/${R}_${K}_set_full_summary.csv 2>/dev/null; head -5 $O/${R}_${K}_set_full_summary.csv | cut -c1-600
python - <<'PY'
import csv, json, collections
R, K = "r2b", "dcb_tail"
rows = list(csv.DictReader(l for l in open(f"gpurun_out/{R}_traffic_{K}.csv") if not l.startswith("==")))
per = collections.defaultdict(dict)
for r in rows:
    per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "%": 1, "inst": 1}.get(r["Metric Unit"], 1)
n = len(per)
s = lambda m: sum(v.get(m, 0.0) for v in per.values())
out = {"kernel": K + "_kernel", "launches": n, "dram_bytes_per_launch": (s("dram__bytes_read.sum") + s("dram__bytes_write.sum")) / n,
       "lts_bytes_per_launch": s("lts__t_bytes.sum") / n, "avg_launch_us_cold": s("gpu__time_duration.sum") / n / 1e3,
       "tensor_pipe_active_pct_avg": s("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") / n,
       "how": "ncu --clock-control none, cache flushed before every launch, the 13 launches of one 1080p Intra decode (tools/r2_calls/r2_call22.sh)"}
json.dump(out, open(f"gpurun_out/{R}_traffic_{K}.json", "w"), indent=1)
print(out)
PY
