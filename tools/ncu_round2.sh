#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU):
#   1. launch list of the bench command (device time of every launch; cold-cache and serialised: compare SHARES)
#   2. one `ncu --set full` capture per hot kernel, each after its plain command exited 0:
#        matcher, RANSAC round kernel at the headline threshold (round 0) and at threshold 0.05 (a 128-sample round),
#        root finder (stage B1), single-pair LM, windowed LM (504 windows of BASELINE config 5)
# usage: tools/ncu_round2.sh <tag>   -> gpurun_out/<tag>_*.{csv,ncu-rep}
tag=${1:-r2}
set -x
BENCH="python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline"
$BENCH > gpurun_out/${tag}_bench_plain.json 2> gpurun_out/${tag}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench_steps2_warmup1.csv $BENCH > gpurun_out/${tag}_bench_ncu.log 2>&1
full() { # name regex skip command...
  name=$1; re=$2; skip=$3; shift 3
  "$@" > gpurun_out/${tag}_$name.plain.log 2>&1 &&
  ncu --set full --import-source on --clock-control none -k regex:$re -s $skip -c 1 -f -o gpurun_out/${tag}_$name "$@" > gpurun_out/${tag}_$name.log 2>&1
}
full match match_tile 2 python tools/iters_hist.py
full ess_round_thr1 ess_round_kernel 20 python tools/iters_hist.py
full solve_b1 solve_b1_kernel 20 python tools/iters_hist.py
full lm_pair lm_pair_kernel 2 python tools/iters_hist.py
full ess_round_thr005 ess_round_kernel 24 env THR=0.05 python tools/iters_hist.py
full lm_windows lm_kernel 1 env B5=504 python tools/lm_windows.py
ls -la gpurun_out/${tag}_*.ncu-rep
