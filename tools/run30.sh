cd $GRAFT_REPO_ROOT
echo "--- micro (CUDA graph): streaming / ares1 / pair"
for a in "136 240 384 384" "136 240 384 384 0 0 1" "136 240 384 1536 1 1" "68 120 512 512 1 0 0" "68 120 512 2048 1 1" "136 240 512 256 1" "136 240 256 512 0 0 1" "136 240 512 1024 1 1"; do
DCVC_B200_GEMM_ARES=0 timeout 60 python tools/gemm_micro.py $a 2>&1 | tail -1
DCVC_B200_GEMM_PAIR=0 timeout 60 python tools/gemm_micro.py $a 2>&1 | tail -1
timeout 60 python tools/gemm_micro.py $a 2>&1 | tail -1
done
echo "--- bench streaming"
DCVC_B200_GEMM_ARES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['gpu_only_ms_per_decode'], d['encode_fps'], d['roofline']['families_ms_per_step'], d['hts']['decode_fps'], d['hts']['encode_fps'], d['hts']['gpu_only_ms_per_chunk_decode'])"
echo "--- bench pair"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['gpu_only_ms_per_decode'], d['encode_fps'], d['roofline']['families_ms_per_step'], d['hts']['decode_fps'], d['hts']['encode_fps'], d['hts']['gpu_only_ms_per_chunk_decode'])"
