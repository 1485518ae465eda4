#!/bin/bash
# Round-2 GPU session A (one GPU): full GPU suite incl. the new scale tests, A/B of the queued forward / epilogue
# variants, finer-item timing at C1, a bench line.  Every step has its own timeout and logs into gpurun_out/.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/session.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee -a gpurun_out/session.log
stamp "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --durations=15 > gpurun_out/pytest.log 2>&1
stamp "pytest rc=$? : $(tail -1 gpurun_out/pytest.log)"
stamp "ab_bench"
timeout 420 python tools/ab_bench.py --out gpurun_out/ab.jsonl --timeout 100 > gpurun_out/ab.log 2>&1
stamp "ab rc=$? : $(tail -1 gpurun_out/ab.log | head -c 600)"
for ie in 64 32 16; do
  stamp "kbench wn18rr x4 heads item_edges=$ie"
  timeout 120 python tools/kbench.py --scale wn18rr --heads 4 --item-edges $ie > gpurun_out/kbench_c1_ie$ie.log 2>&1
  timeout 120 python tools/kbench.py --scale fb15k237 --heads 128 --alphas 1.0 --item-edges $ie > gpurun_out/kbench_c2_ie$ie.log 2>&1
  timeout 120 python tools/kbench.py --scale fb15k237 --heads 16 --alphas 1.0 --item-edges $ie > gpurun_out/kbench_c2h16_ie$ie.log 2>&1
done
stamp "bench"
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
stamp "bench rc=$? : $(head -c 400 gpurun_out/bench.json)"
stamp "done"
