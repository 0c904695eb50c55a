#!/bin/bash
# dcb_tail without the lock-step option: pre-wait weight depth A/B on one box, host-side decode trace, fused on/off bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
echo "== dcb_tail tests"
timeout 150 python -m pytest tests/test_dcb_tail_gpu.py -q -x > $O/r2c20_tail.log 2>&1; echo "rc=$?"; tail -2 $O/r2c20_tail.log
echo "== micro: weight stages before griddepcontrol.wait (0 = all that fit)"
for R in 1 2 3; do
  for PRE in 0 2 1; do
    echo -n "pre=$PRE: "; DCVC_B200_DT_PRE=$PRE timeout 60 python tools/dcb_tail_micro.py 136 240 384 384 384 fused 2>&1 | grep fused
  done
done
timeout 60 python tools/dcb_tail_micro.py 136 240 384 384 384 perop 2>&1 | grep per-op
timeout 60 python tools/dcb_tail_micro.py 270 480 384 384 384 2>&1 | grep "dbg=0"
echo "== host trace (1080p intra decode)"
DCVC_B200_HOST_TRACE=1 timeout 120 python tools/profile_decode.py 1080 1920 4 2>&1 | grep -E "host trace|done" | cut -c1-600
echo "== 4K fused identity test (C=512 forced)"
timeout 300 python -m pytest tests/test_codec_gpu.py -m gpu -q -x -k "fused_block_tails" > $O/r2c20_pytest.log 2>&1; echo "rc=$?"; tail -2 $O/r2c20_pytest.log
echo "== bench A/B"
for V in "on:1" "off:0" "on2:1"; do
    IFS=: read NAME FT <<< "$V"
    DCVC_B200_FUSE_TAIL=$FT timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-seq8 > $O/r2c20_bench_$NAME.json 2> $O/r2c20_bench_$NAME.err
    python - <<PY
import json
try:
    d = json.loads(open("$O/r2c20_bench_$NAME.json").read().strip().splitlines()[-1])
    h = d["hts"]; r = d["roofline"]; x = d["hts_extra"]
    print("$NAME: intra dec %.1f e2e %.1f gpu-only %.3f ms enc %.1f | roofline %s frac %.3f tensor %.3f whole %.3f | hts dec %.1f gpu-only %.3f enc %.1f | ld %s %s | htl %s %s | 4K %s %s" %
          (d["value"], d["e2e"]["value"], d["gpu_only_ms_per_decode"], d["encode_fps"], r["kernel"], r["frac"], r["tensor_frac"], r["whole_decode_frac"],
           h["decode_fps"], h["gpu_only_ms_per_chunk_decode"], h["encode_fps"], d["ld"].get("decode_fps"), d["ld"].get("encode_fps"), d["htl"].get("decode_fps"), d["htl"].get("encode_fps"), x.get("decode_fps"), x.get("encode_fps")))
except Exception as e:
    print("$NAME: no result (%s)" % e); print(open("$O/r2c20_bench_$NAME.err").read()[-800:])
PY
done
