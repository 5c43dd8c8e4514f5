#!/usr/bin/env python
"""Copy-only ceiling of the end-to-end input path.

Every rank does nothing but what the e2e step's upload does: one pinned host buffer of the size of the benchmark
sequence (4541 frames x 2000 kp x 40 B = 363 MB) copied host->device with cudaMemcpyAsync, `--steps` times, all
ranks at once (barrier before every step).  Reports per-rank ms and the aggregate GB/s; the e2e number of
`bench.py --gpus N` cannot be better than  pairs_per_rank * N / (this time).

  python tools/h2d_ceiling.py                      # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/h2d_ceiling.py --out gpurun_out/r2_h2d_ceiling_8.json
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=4541 * 2000 * 40)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--pieces", type=int, default=1, help="cut the copy into this many cudaMemcpyAsync calls")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    host = torch.empty(a.bytes, dtype=torch.uint8).pin_memory()
    host.random_(0, 255)
    dev = torch.empty(a.bytes, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    cuts = [a.bytes * i // a.pieces for i in range(a.pieces + 1)]

    def copy():
        with torch.cuda.stream(st):
            for i in range(a.pieces):
                dev[cuts[i]:cuts[i + 1]].copy_(host[cuts[i]:cuts[i + 1]], non_blocking=True)

    for _ in range(3):
        copy()
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        copy()
        e1.record(st)
        st.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = torch.tensor(ms, device="cuda", dtype=torch.float64)
    if world > 1:
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        allt = torch.stack(allt).cpu()
    else:
        allt = t.cpu()[None]
    if rank == 0:
        per_step_max = allt.max(0).values           # a step ends when the slowest rank has its data
        med = float(per_step_max.median())
        out = {"what": "copy-only host->device ceiling, all ranks concurrently, pinned memory, cudaMemcpyAsync",
               "n_gpus": world, "bytes_per_rank": a.bytes, "pieces": a.pieces, "steps": a.steps,
               "ms_per_step_max_over_ranks_median": med,
               "ms_per_rank_median": [float(x) for x in allt.median(1).values],
               "per_rank_GBps_median": [a.bytes / float(x) / 1e6 for x in allt.median(1).values],
               "aggregate_GBps": world * a.bytes / med / 1e6,
               "e2e_pairs_per_s_ceiling": world * 4540 / med * 1e3}
        s = json.dumps(out)
        print(s)
        if a.out:
            open(a.out, "w").write(s + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
