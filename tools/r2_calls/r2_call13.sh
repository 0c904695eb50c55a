#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
echo "== micro (acc pre-wait reverted)"
timeout 120 python tools/dcb_tail_micro.py 136 240 384 384 384 2>&1 | tail -2
echo "== pytest -m gpu (everything)"
timeout 2400 python -m pytest tests -m gpu -q -x -s > $O/r2c13_pytest.log 2>&1
echo "rc=$?"; grep -E "parity|block parity|dcb_tail\]" $O/r2c13_pytest.log | head -30; tail -5 $O/r2c13_pytest.log
echo "== bench A/B fused (intra + hts legs only)"
for V in "on:1" "off:0"; do
    IFS=: read NAME FT <<< "$V"
    DCVC_B200_FUSE_TAIL=$FT timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-seq8 --hts-size none > $O/r2c13_bench_$NAME.json 2> $O/r2c13_bench_$NAME.err
    python - <<PY
import json
try:
    d = json.loads(open("$O/r2c13_bench_$NAME.json").read().strip().splitlines()[-1])
    h = d["hts"]; r = d["roofline"]
    print("$NAME: intra dec %.1f e2e %.1f gpu-only %.3f ms enc %.1f launches %d | roofline %s frac %.3f whole %.3f | hts dec %.1f gpu-only %.3f enc %.1f | ld %s %s | htl %s %s" %
          (d["value"], d["e2e"]["value"], d["gpu_only_ms_per_decode"], d["encode_fps"], d["gpu_launches"], r["kernel"], r["frac"], r["whole_decode_frac"],
           h["decode_fps"], h["gpu_only_ms_per_chunk_decode"], h["encode_fps"], d["ld"].get("decode_fps"), d["ld"].get("encode_fps"), d["htl"].get("decode_fps"), d["htl"].get("encode_fps")))
except Exception as e:
    print("$NAME: no result (%s)" % e); print(open("$O/r2c13_bench_$NAME.err").read()[-1500:])
PY
done
echo "== the default bench line (everything on), timed"
T0=$(date +%s); timeout 900 python bench.py > $O/r2c13_bench_default.json 2> $O/r2c13_bench_default.err; RC=$?; echo "elapsed $(( $(date +%s) - T0 )) s"; (exit $RC)
echo "rc=$?"
python - <<PY
import json
d = json.loads(open("$O/r2c13_bench_default.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","encode_fps","gpu_only_ms_per_decode","cpu_baseline","seq8","hts_extra","reference_cuda","parity","speedup_vs_reference_cuda","clocks"):
    print(k, "=", json.dumps(d.get(k))[:700])
PY
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-900
