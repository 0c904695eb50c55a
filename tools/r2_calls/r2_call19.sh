#!/bin/bash
# dcb_tail epilogue A/B: all 16 warps on every chunk (DCVC_B200_DT_EPI16=1) vs two groups of 8 on alternate chunks
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
for E in 1 0; do
    export DCVC_B200_DT_EPI16=$E
    echo "== EPI16=$E dcb_tail tests"
    timeout 150 python -m pytest tests/test_dcb_tail_gpu.py -q -x > $O/r2c19_tail_$E.log 2>&1; echo "rc=$?"; tail -2 $O/r2c19_tail_$E.log
    echo "== EPI16=$E micro"
    timeout 60 python tools/dcb_tail_micro.py 136 240 384 384 384 2>&1 | grep fused
    timeout 60 python tools/dcb_tail_micro.py 270 480 384 384 384 2>&1 | grep "dbg=0" 
    timeout 60 python tools/dcb_tail_micro.py 68 120 512 512 512 2>&1 | grep "dbg=0"
    timeout 60 python tools/dcb_tail_micro.py 135 240 512 512 512 2>&1 | grep "dbg=0"
done
export DCVC_B200_DT_EPI16=1
echo "== EPI16=1 codec diag 4K"
timeout 200 python tools/diag_fuse.py 2160 3840 40 2>&1 | tail -4 | cut -c1-150
echo "== EPI16=1 codec tests"
timeout 600 python -m pytest tests/test_codec_gpu.py tests/test_hts_gpu.py tests/test_ld_gpu.py -m gpu -q -x > $O/r2c19_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/r2c19_pytest.log
echo "== bench A/B"
for E in 1 0; do
    DCVC_B200_DT_EPI16=$E timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-seq8 > $O/r2c19_bench_$E.json 2> $O/r2c19_bench_$E.err
    python - <<PY
import json
try:
    d = json.loads(open("$O/r2c19_bench_$E.json").read().strip().splitlines()[-1])
    h = d["hts"]; r = d["roofline"]; x = d["hts_extra"]
    print("epi16=$E: intra dec %.1f e2e %.1f gpu-only %.3f ms enc %.1f | roofline %s frac %.3f tensor %.3f whole %.3f | hts dec %.1f gpu-only %.3f enc %.1f | ld %s %s | htl %s %s | 4K %s %s" %
          (d["value"], d["e2e"]["value"], d["gpu_only_ms_per_decode"], d["encode_fps"], r["kernel"], r["frac"], r["tensor_frac"], r["whole_decode_frac"],
           h["decode_fps"], h["gpu_only_ms_per_chunk_decode"], h["encode_fps"], d["ld"].get("decode_fps"), d["ld"].get("encode_fps"), d["htl"].get("decode_fps"), d["htl"].get("encode_fps"), x.get("decode_fps"), x.get("encode_fps")))
except Exception as e:
    print("epi16=$E: no result (%s)" % e); print(open("$O/r2c19_bench_$E.err").read()[-800:])
PY
done
