cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "--- tests"
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
echo "--- bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench36.json 2> gpurun_out/bench36.err; tail -c 4000 gpurun_out/bench36.json; tail -3 gpurun_out/bench36.err
