#!/bin/bash
# Round-2 final evidence session (one GPU): bench line + reference arm, ncu launch list of the bench command, --set full
# captures of the dominant kernels, sweep with same-run parity, full GPU suite log.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/session_d.log; }
stamp "bench"
timeout 420 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
stamp "bench rc=$? : $(head -c 200 gpurun_out/bench_final.json) $(tail -2 gpurun_out/bench_final.err)"
stamp "reference arm"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
stamp "ncu launch list of bench.py"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/ncu_bench_stdout.log 2>&1
stamp "launch list rc=$?"
stamp "ncu --set full: forward"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:seg_reduce_kernel -s 3 -c 1 -f -o gpurun_out/r02_fwd \
    python tools/kbench.py --alphas 1.0 > gpurun_out/ncu_fwd.log 2>&1
stamp "ncu --set full: backward (two passes)"
timeout 240 ncu --set full --clock-control none -k regex:seg_reduce_kernel -s 13 -c 2 -f -o gpurun_out/r02_bwd \
    python tools/kbench.py --alphas 1.0 --bwd > gpurun_out/ncu_bwd.log 2>&1
stamp "ncu --set full: update kernel"
timeout 200 ncu --set full --clock-control none -k regex:update_ -s 3 -c 1 -f -o gpurun_out/r02_update \
    python tools/update_bench.py > gpurun_out/ncu_upd.log 2>&1
stamp "summaries of the captures (the .ncu-rep files are too large to travel back: 64 MiB limit)"
python tools/ncu_summary.py gpurun_out/r02_fwd.ncu-rep gpurun_out/r02_bwd.ncu-rep gpurun_out/r02_update.ncu-rep > gpurun_out/r02_ncu_summary.md 2>&1
python tools/ncu_roofline_json.py --out=gpurun_out/roofline_traffic.json forward=gpurun_out/r02_fwd.ncu-rep backward_in=gpurun_out/r02_bwd.ncu-rep:0 \
    backward_rel=gpurun_out/r02_bwd.ncu-rep:1 update=gpurun_out/r02_update.ncu-rep > /dev/null 2>&1
ncu -i gpurun_out/r02_fwd.ncu-rep --page source --csv 2>/dev/null | head -400 > gpurun_out/r02_fwd_source_page_head.csv
rm -f gpurun_out/r02_bwd.ncu-rep gpurun_out/r02_update.ncu-rep
stamp "sweep (C1-C5) with same-run parity"
timeout 600 python tools/sweep.py > gpurun_out/r02_sweep_stdout.jsonl 2> gpurun_out/sweep.err
stamp "sweep rc=$? lines=$(wc -l < gpurun_out/r02_sweep_stdout.jsonl)"
stamp "done"
