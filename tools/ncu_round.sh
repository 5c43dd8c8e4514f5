#!/bin/bash
# One `ncu --set full` capture per hot kernel of the pipeline (warm launch, round-0 launch for the essential rounds).
# usage: tools/ncu_round.sh <tag>   -> gpurun_out/<tag>_<kernel>.ncu-rep
tag=${1:-rX}
run() { # name regex skip
  ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${tag}_$1 python tools/iters_hist.py > gpurun_out/${tag}_$1.log 2>&1
}
run match match_tile 2
run solve_a solve_a_kernel 20
run solve_b1 solve_b1_kernel 20
run solve_b2 solve_b2_kernel 20
run ess_round ess_round_kernel 20
run pose pose_kernel 2
run lm lm_pair_kernel 2
ls -la gpurun_out/${tag}_*.ncu-rep
