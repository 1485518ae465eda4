#!/bin/bash
# Round-2 GPU session B (one GPU): A/B of the 1024-thread CTA geometry, full bench (strong-scaling records at N=1,
# training layer), the reference arm, ncu launch list of the bench command and --set full captures of the dominant kernels.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/session_b.log; }
stamp "ab_bench (default vs variants)"
timeout 300 python tools/ab_bench.py --out gpurun_out/ab_b.jsonl --timeout 120 > gpurun_out/ab_b.log 2>&1
stamp "ab rc=$? : $(tail -1 gpurun_out/ab_b.log | head -c 500)"
stamp "bench"
timeout 420 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
stamp "bench rc=$? : $(head -c 300 gpurun_out/bench_b.json) $(tail -3 gpurun_out/bench_b.err)"
stamp "reference arm"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_b.json 2> gpurun_out/bench_ref_b.err
stamp "reference rc=$? : $(head -c 300 gpurun_out/bench_ref_b.json)"
stamp "ncu launch list of bench.py"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/ncu_bench_stdout.log 2>&1
stamp "launch list rc=$?"
stamp "ncu --set full: forward"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:seg_reduce_kernel -s 3 -c 1 -f -o gpurun_out/r02_fwd \
    python tools/kbench.py --alphas 1.0 > gpurun_out/ncu_fwd.log 2>&1
stamp "ncu --set full: backward (two passes)"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:seg_reduce_kernel -s 13 -c 2 -f -o gpurun_out/r02_bwd \
    python tools/kbench.py --alphas 1.0 --bwd > gpurun_out/ncu_bwd.log 2>&1
stamp "ncu --set full: update"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:update_tc3 -s 3 -c 1 -f -o gpurun_out/r02_update \
    python tools/update_bench.py > gpurun_out/ncu_upd.log 2>&1
stamp "done"
