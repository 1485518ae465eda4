#!/bin/bash
# One bounded GPU session (run under gpurun): A/B the compile-time forward-kernel variants that tools/build_variants.py
# built (ultra_b200/lib/variants/), validate the winner with the full GPU test-suite, take a bench line and an ncu
# capture with it.  Every step has its own timeout and writes into gpurun_out/ as it goes, so a cut-off session still
# leaves its earlier results.  Round-1 use: profiles/r01_ab_session.log (metadata staging variants, BF16x3 epilogue).
#   python tools/build_variants.py && gpurun --timeout 480 -- 'bash tools/gpu_ab_session.sh'
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/session.log; }
DEFAULT_LIB="$PWD/ultra_b200/lib/libultra_rspmm.so"

stamp "ab_bench"
timeout 300 python tools/ab_bench.py --out gpurun_out/ab.jsonl --timeout 90 > gpurun_out/ab.log 2>&1
WIN=$(python - <<'PY'
import json
try:
    print([json.loads(l) for l in open("gpurun_out/ab.jsonl")][-1]["winner_lib"])
except Exception:
    print("")
PY
)
[ -n "$WIN" ] && [ -f "$WIN" ] || WIN="$DEFAULT_LIB"
stamp "winner: $WIN"

stamp "pytest -m gpu on the winner"
ULTRA_RSPMM_LIB=$WIN timeout 240 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_winner.log 2>&1
stamp "pytest rc=$? : $(tail -1 gpurun_out/pytest_winner.log)"

stamp "bench.py with the winner"
ULTRA_RSPMM_LIB=$WIN timeout 240 python bench.py > gpurun_out/bench_winner.json 2> gpurun_out/bench_winner.err
stamp "bench rc=$? : $(head -c 300 gpurun_out/bench_winner.json)"

if [ "$WIN" != "$DEFAULT_LIB" ]; then
    stamp "ncu --set full of the winner's forward kernel"
    ULTRA_RSPMM_LIB=$WIN timeout 150 ncu --set full --clock-control none --import-source on -k regex:seg_reduce_kernel -s 3 -c 1 \
        -f -o gpurun_out/fwd_winner python tools/kbench.py --alphas 1.0 > gpurun_out/ncu_fwd_winner.log 2>&1
    stamp "ncu rc=$?"
fi
stamp "done"
