#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
echo "== micro"
for shape in "136 240 384 384 384" "136 240 384 384 0" "68 120 512 512 512" "34 60 512 512 512"; do
  timeout 120 python tools/dcb_tail_micro.py $shape 2>&1 | tail -2
done
for dbg in 1 2 3; do
  DCVC_B200_GEMM_DBG=$dbg timeout 120 python tools/dcb_tail_micro.py 136 240 384 384 384 fused 2>&1 | tail -1
done
echo "== pytest -m gpu (fused tail on)"
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_reference_surface_gpu.py > $O/r2c3_pytest.log 2>&1
echo "rc=$?"; tail -8 $O/r2c3_pytest.log
echo "== bench fused on / off"
for V in "on:1" "off:0"; do
    IFS=: read NAME FT <<< "$V"
    DCVC_B200_FUSE_TAIL=$FT timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2c3_bench_$NAME.json 2> $O/r2c3_bench_$NAME.err
    python - <<PY
import json
try:
    d = json.loads(open("$O/r2c3_bench_$NAME.json").read().strip().splitlines()[-1])
    h = d["hts"]; r = d["roofline"]
    print("$NAME: intra dec %.1f e2e %.1f gpu-only %.3f ms enc %.1f launches %d | roofline frac %.3f share %.3f whole %.3f | hts dec %.1f gpu-only %.3f enc %.1f | ld %s %s | htl %s %s" %
          (d["value"], d["e2e"]["value"], d["gpu_only_ms_per_decode"], d["encode_fps"], d["gpu_launches"], r["frac"], r["share_of_gpu_time"], r["whole_decode_frac"],
           h["decode_fps"], h["gpu_only_ms_per_chunk_decode"], h["encode_fps"], d["ld"].get("decode_fps"), d["ld"].get("encode_fps"), d["htl"].get("decode_fps"), d["htl"].get("encode_fps")))
except Exception as e:
    print("$NAME: no result (%s)" % e); print(open("$O/r2c3_bench_$NAME.err").read()[-1500:])
PY
done
