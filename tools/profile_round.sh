#!/bin/bash
# Round evidence (run under gpurun on one B200): launch list of the bench command, DRAM / L2 traffic of every dcb_tail and
# pw_gemm launch of one 1080p Intra decode, one --set full capture of each of the two tcgen05 kernels, the DRAM traffic of a 4K
# HT-S chunk decode (configs[4]) and a memcheck pass over the small op tests.
# Outputs go to gpurun_out/; summaries are copied into profiles/ (see profiles/README.md).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
R=${1:-r2}
O=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hts --no-reference-cuda --no-seq8 > $O/${R}_bench_under_ncu.log 2>&1
echo "launch list rows: $(wc -l < $O/${R}_launches_bench.csv)"
python tools/summarize_ncu.py $O/${R}_launches_bench.csv > $O/${R}_launches_bench.md 2>&1; head -14 $O/${R}_launches_bench.md
for K in dcb_tail pw_gemm; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none \
      -k regex:${K}_kernel -s 60 -c 90 --csv --log-file $O/${R}_traffic_${K}.csv python tools/profile_decode.py 1080 1920 1 > /dev/null 2>&1
  echo "traffic rows ${K}: $(wc -l < $O/${R}_traffic_${K}.csv)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 70 -c 4 -f -o $O/${R}_prof_${K} \
      python tools/profile_decode.py 1080 1920 1 > /dev/null 2>&1
  timeout 300 ncu -i $O/${R}_prof_${K}.ncu-rep --page raw --csv > $O/${R}_prof_${K}_raw.csv 2>/dev/null
  echo "set-full raw rows ${K}: $(wc -l < $O/${R}_prof_${K}_raw.csv)"
done
python - <<'PY'
import csv, json, collections, re, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
for K in ("dcb_tail", "pw_gemm"):
    try:
        rows = list(csv.DictReader(l for l in open(f"gpurun_out/{R}_traffic_{K}.csv") if not l.startswith("==")))
    except OSError:
        continue
    per = collections.defaultdict(dict)
    for r in rows:
        per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "%": 1, "inst": 1}.get(r["Metric Unit"], 1)
    n = len(per)
    if not n:
        continue
    s = lambda m: sum(v.get(m, 0.0) for v in per.values())
    out = {"kernel": K + "_kernel", "launches": n, "dram_bytes_per_launch": (s("dram__bytes_read.sum") + s("dram__bytes_write.sum")) / n,
           "lts_bytes_per_launch": s("lts__t_bytes.sum") / n, "avg_launch_us_cold": s("gpu__time_duration.sum") / n / 1e3,
           "tensor_pipe_active_pct_avg": s("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") / n,
           "how": "ncu --clock-control none, cache flushed before every launch, one 1080p Intra encode + decode (tools/profile_round.sh)"}
    json.dump(out, open(f"gpurun_out/{R}_traffic_{K}.json", "w"), indent=1)
    print(out)
PY
echo "== 4K HT-S chunk decode: DRAM traffic (configs[4])"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv \
    --log-file $O/${R}_traffic_hts4k.csv python tools/profile_hts.py 2160 3840 > $O/${R}_hts4k_under_ncu.log 2>&1
python - <<'PY'
import csv, collections, json, re
try:
    rows = list(csv.DictReader(l for l in open("gpurun_out/r2_traffic_hts4k.csv") if not l.startswith("==")))
    per = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        v = float(r["Metric Value"].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1e-3, "us": 1, "usecond": 1, "nsecond": 1e-3}.get(r["Metric Unit"], 1)
        if r["Metric Name"].startswith("dram"):
            per[name][1] += v
        else:
            per[name][2] += v; per[name][0] += 1
    tot_b = sum(v[1] for v in per.values()); tot_us = sum(v[2] for v in per.values())
    print(f"{sum(v[0] for v in per.values())} launches of one 4K HT-S chunk decode: {tot_b / 1e9:.2f} GB DRAM traffic, {tot_us / 1e3:.2f} ms serialised -> {tot_b / tot_us / 1e3:.0f} GB/s")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][2])[:6]:
        print(f"  {k}: {v[0]} launches, {v[1] / 1e6:.0f} MB, {v[2]:.0f} us, {v[1] / max(v[2], 1e-9) / 1e3:.0f} GB/s")
except Exception as e:
    print("4K summary failed:", e)
PY
echo "== memcheck"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_dcb_tail_gpu.py -q -x -k "16-16 or 17-30 or 32-40 or pitched or 68-120-256" > $O/${R}_sanitizer_memcheck.log 2>&1
tail -4 $O/${R}_sanitizer_memcheck.log
ls -la $O/ | grep ${R}_ | head -30
