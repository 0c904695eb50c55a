#!/bin/bash
# round-2 validation on one box: the whole GPU suite, smoke(), and the default bench line of both arms
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"
S=$(date +%s)
timeout 1500 python -m pytest tests/ -m gpu -x -q > $O/r2c21_pytest.log 2>&1; echo "rc=$? ($(( $(date +%s) - S )) s)"; tail -4 $O/r2c21_pytest.log
grep "parity vs reference CUDA" $O/r2c21_pytest.log | head -12
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c21_smoke.log 2>&1; echo "rc=$?"; tail -3 $O/r2c21_smoke.log
echo "== bench (default flags)"
S=$(date +%s)
timeout 900 python bench.py > $O/r2c21_bench.json 2> $O/r2c21_bench.err; echo "rc=$? ($(( $(date +%s) - S )) s)"
tail -c 6000 $O/r2c21_bench.json
echo
echo "== bench --impl reference"
S=$(date +%s)
timeout 900 python bench.py --impl reference > $O/r2c21_bench_ref.json 2> $O/r2c21_bench_ref.err; echo "rc=$? ($(( $(date +%s) - S )) s)"
tail -c 1500 $O/r2c21_bench_ref.json
