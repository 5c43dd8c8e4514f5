#!/bin/bash
# GPU check after a change to the essential-matrix kernels: parity tests of the stage, then stage times of the
# headline call and of the reference's other findEssentialMat shapes on the benchmark sequence.
# usage: tools/check_ess.sh <tag> [ncu]
tag=${1:-ess}
timeout 900 python -m pytest tests/test_gpu_essential.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_edge_pipeline.py tests/test_gpu_cv2_census.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
tail -3 gpurun_out/${tag}_tests.log
{ python tools/iters_hist.py; THR=0.05 python tools/iters_hist.py; LMEDS=1 THR=0.01 python tools/iters_hist.py; THR=0.3 python tools/iters_hist.py; } > gpurun_out/${tag}_stages.log 2>&1
grep product gpurun_out/${tag}_stages.log
if [ "$2" = ncu ]; then
  ncu --set full --import-source on --clock-control none -k regex:ess_round_kernel -s 24 -c 1 -f -o gpurun_out/${tag}_ess_round_thr005 env THR=0.05 python tools/iters_hist.py > gpurun_out/${tag}_ncu.log 2>&1
fi
