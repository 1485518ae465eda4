#!/bin/bash
# Short GPU session: validate the library as built (full GPU suite), take the bench line, capture the forward kernel.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/session2.log; }
stamp "pytest -m gpu (library defaults)"
timeout 150 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_final.log 2>&1
stamp "pytest rc=$? : $(tail -1 gpurun_out/pytest_final.log)"
stamp "bench.py"
timeout 120 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
stamp "bench rc=$? : $(head -c 200 gpurun_out/bench_final2.json)"
stamp "ncu --set full of the forward kernel"
timeout 90 ncu --set full --clock-control none --import-source on -k regex:seg_reduce_kernel -s 3 -c 1 \
    -f -o gpurun_out/r01_fwd_v6 python tools/kbench.py --alphas 1.0 > gpurun_out/ncu_fwd_v6.log 2>&1
stamp "ncu rc=$?"
stamp "done"
