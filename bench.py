#!/usr/bin/env python
"""Benchmark of the per-frame-pair geometric core (BASELINE.json metric):

    KITTI-shape frame-pairs/sec (match + E-RANSAC + pose + LM)

Workload (config 3 of BASELINE.json): a synthetic KITTI seq-00-length run, 4541 frames x 2000
ORB-shaped keypoints (256-bit descriptors) = 4540 consecutive frame pairs per GPU.  One "step"
is one pass of the whole hot path over that sequence:
  cross-check Hamming2 matching (kitti_ba.cpp:602,641) -> findEssentialMat(RANSAC, 0.99, 1.0)
  (kitti.cpp:98-104) -> recoverPose (kitti_E.cpp:120) -> 48-point Rt Levenberg-Marquardt
  (kitti_E.cpp:170-201).
`value`  : inputs already resident in HBM, CUDA events on the library's stream.
`e2e`    : the same through the public API with HOST buffers: every step copies the sequence
           host->device from pinned memory, runs, and copies the per-pair results back.
N > 1    : one process per GPU (torchrun); every rank processes its own 4540-pair sequence
           (weak scaling, no data-path collective); the per-pair poses are all-gathered with
           NCCL inside the e2e region; times are the max over ranks.
`--impl reference`: the reference's CPU path (cv2 = the OpenCV calls the reference makes + the
           plain-C restatement of its LM) on all host cores, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "KITTI-shape frame-pairs/sec (match+E-RANSAC+pose+LM)"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4541)
    ap.add_argument("--kp", type=int, default=2000)
    ap.add_argument("--method", default="ransac", choices=["ransac", "lmeds"])
    ap.add_argument("--thr", type=float, default=1.0)
    ap.add_argument("--match", default="crosscheck", choices=["crosscheck", "ratio"],
                    help="crosscheck = BFMatcher(norm, true).match (the reference, kitti_ba.cpp:602); "
                         "ratio = knnMatch(k=2) + Lowe ratio 0.8 (north_star's matcher mode)")
    ap.add_argument("--norm", default="hamming2", choices=["hamming", "hamming2"])
    ap.add_argument("--cpu-pairs", type=int, default=128, help="bounded CPU-baseline sample (pairs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="pairs", choices=["pairs", "ba_windows"],
                    help="pairs = the headline metric (default); ba_windows = BASELINE.json config 5: the 504 "
                         "kitti_ba windows of a sequence (n_zeta 10, 20 reps x 250) sharded over the GPUs (strong scaling)")
    ap.add_argument("--windows", type=int, default=504)
    return ap.parse_args()


def workload_name(a):
    return (f"kitti_E synthetic seq-00-length run: {a.frames} frames x {a.kp} kp x 256-bit descriptors, "
            f"{a.frames - 1} pairs per GPU; BFMatcher({a.norm.upper()}, "
            f"{'crossCheck' if a.match == 'crosscheck' else 'knnMatch k=2 + ratio 0.8'}) + findEssentialMat("
            f"{a.method.upper()}, 0.99, {a.thr}) + recoverPose + 48-pt LM")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].strip().replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].strip().replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for k, nm in enumerate(names):
                if r[5 + k].strip().lower() == "active":
                    reasons.add(nm)
        out.update(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                   reasons=sorted(reasons), samples=len(rows))
        return out


def bind_near_gpu(local):
    """N > 1: run this rank (and first-touch its pinned staging buffers) on the host cores NVML reports as local
    to its GPU, so that eight ranks pulling 363 MB per step do not all cross the socket interconnect.
    Best effort: any failure leaves the affinity as it was."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def make_inputs(a, rank):
    from epivo_b200 import synth
    seq = synth.make_sequence(a.frames, a.kp, seed=synth.seed_for(3, rank))
    return seq


def run_reference(a):
    """`--impl reference`: rank 0 only; bounded sample per step on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from epivo_b200 import synth
    from oracle import cpu_reference as R
    per_step = max(8, min(a.cpu_pairs, a.frames - 1))
    seq = synth.make_sequence(min(a.frames, per_step + 1), a.kp, seed=synth.seed_for(3, 0))
    method = 8 if a.method == "ransac" else 4
    cores = os.cpu_count() or 1
    pool = R.CpuPool(seq.kps, seq.descs, seq.K, method, 0.99, a.thr, cores=cores,
                     norm=7 if a.norm == "hamming2" else 6, ratio=None if a.match == "crosscheck" else 0.8)
    idx = list(range(seq.n_pairs))
    for _ in range(a.warmup):
        pool.run(idx[:max(cores, 8)])
    t0 = time.perf_counter()
    for _ in range(a.steps):
        pool.run(idx)
    dt = time.perf_counter() - t0
    pool.close()
    value = a.steps * len(idx) / dt
    kind = "reference" if R.HAVE_CV2 else "port"
    sample = (f"{len(idx)} pairs per step of the same synthetic sequence; "
              + ("cv2 %s BFMatcher/findEssentialMat/recoverPose (the OpenCV calls the reference makes) + plain-C "
                 "restatement of its LM (Eigen/Sophus original not buildable)" % R.cv2.__version__ if R.HAVE_CV2
                 else "numpy restatement (cv2 missing)")
              + f"; {cores} worker processes x 1 thread, pair-parallel")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8+f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(a), "pairs_per_step": len(idx)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line when
    NCCL_DEBUG is set on the box), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the
    saved original descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_windows(a):
    """BASELINE.json config 5 (not the headline line): the windows of one sequence, sharded over the ranks in
    contiguous blocks (strong scaling: the total is fixed), one batched Levenberg_Marquardt launch per rank.
    value = windows / max over ranks of the kernel time; e2e = the same through epivo_lm_rt_batch with pinned
    host buffers (H2D of the reprojections + D2H of the refined chains inside the timed region)."""
    import torch
    import torch.distributed as dist
    from epivo_b200 import api, shard, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nz, N = 10, 250
    reps = [(i, i) for i in range(nz)] + [(0, i) for i in range(nz)]
    lo, hi = shard.shard_range(a.windows, world, rank)
    B = hi - lo
    data = [synth.gen_scene_sequence(500 + (lo + b) % 64, N, nz, reps) for b in range(min(B, 64))]

    def pinned(k):
        t = torch.from_numpy(np.stack([data[b % len(data)][k] for b in range(B)])).pin_memory()
        return t, t.numpy()
    keep = [pinned(k) for k in (1, 2, 3)]
    T0, pr, p_r = (x[1] for x in keep)
    ctx = api.Context(local)
    w = [1.0] * len(reps)

    def step():
        return api.Levenberg_Marquardt_batch(nz, 1e-8, reps, w, 1e-2, T0, pr, p_r, huber_delta=1.0, ctx=ctx)
    for _ in range(max(a.warmup, 1)):
        step()
    clocks = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = ctx.launch_count
    t0 = time.perf_counter()
    k_ms = 0.0
    for _ in range(a.steps):
        T, res, its = step()
        k_ms += ctx.last_kernel_ms()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / a.steps
    k_ms /= a.steps
    launches = ctx.launch_count - n0
    if world > 1:
        t = torch.tensor([k_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k_ms, wall_ms = float(t[0].item()), float(t[1].item())
        dist.barrier()
    clk = clocks.stop()
    line = {"metric": "kitti_ba windows/sec (windowed Rt LM, n_zeta 10, 20 reps x 250 correspondences, 30 iterations)",
            "value": a.windows / (k_ms * 1e-3), "unit": "windows/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": k_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"BASELINE config 5: {a.windows} windows (sequence.hpp scenes), contiguous blocks per rank, "
                                   f"{B} on rank {rank}", "parallelism": f"windows sharded, {world} x 1 GPU, no collective"},
            "clocks": clk,
            "e2e": {"value": a.windows / (wall_ms * 1e-3), "unit": "windows/s", "ms_per_step": wall_ms,
                    "h2d_bytes_per_step": int(T0.nbytes + pr.nbytes + p_r.nbytes), "d2h_bytes_per_step": int(T0.nbytes + B * 28)},
            "gpu_launches": int(launches), "mean_iters": float(its.mean())}
    if world == 1 and not a.no_cpu_baseline:
        from oracle import cpu_reference as R
        cores = os.cpu_count() or 1
        n = 2 * cores
        v = R.windows_rate(data, nz, reps, n, cores=cores)
        line["cpu_baseline"] = {"value": v, "unit": "windows/s", "cores": cores, "kind": "port",
                                "sample": f"{n} windows of the same shape; plain-C restatement of the reference LM "
                                          f"(Eigen/Sophus original not buildable), {cores} processes x 1 thread"}
    if rank == 0:
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    quiet_stdout()
    if a.impl == "reference":
        run_reference(a)
        return
    if a.workload == "ba_windows":
        run_windows(a)
        return
    import torch
    import torch.distributed as dist
    from epivo_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    near_cpus = bind_near_gpu(local) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    seq = make_inputs(a, rank)
    F, P, kp = seq.n_frames, seq.n_pairs, a.kp
    # pinned host copies of the inputs (torch owns the pinned allocation; numpy views for the C ABI)
    h_kps = torch.from_numpy(seq.kps).pin_memory()
    h_desc = torch.from_numpy(seq.descs).pin_memory()
    kps_np, desc_np = h_kps.numpy(), h_desc.numpy()
    ctx = api.Context(local)
    pipe = api.SequencePipeline(F, kp, ctx=ctx)
    method = api.RANSAC if a.method == "ransac" else api.LMEDS
    prm = api.default_params(seq.K.astype(np.float32), method=method, threshold=a.thr,
                             match_mode=1 if a.match == "crosscheck" else 2, ratio=0.8,
                             norm=api.NORM_HAMMING2 if a.norm == "hamming2" else api.NORM_HAMMING)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    # results land in pinned host memory too (the library DMAs straight into a pinned caller buffer)
    h_res = torch.zeros(P * api.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    results = h_res.numpy().view(api.RESULT_DTYPE)

    pipe.upload(kps_np, desc_np)
    for _ in range(a.warmup):
        pipe.run(prm, 0, P)
    ctx.sync()

    # ---------------- device-resident timed region: EXACTLY `steps` steps -----------------
    launches0 = ctx.launch_count
    clocks = ClockSampler(local)
    stage_acc = np.zeros(16, dtype=np.float64)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        pipe.run(prm, 0, P)
        stage_acc += pipe.stage_ms()          # syncs the stream (microseconds against a ~100 ms step)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    ms_step = max_over_ranks(ms_total / a.steps)
    value = world * P / (ms_step * 1e-3)

    # ---------------- end to end through the public API with host buffers -----------------
    for _ in range(2):
        pipe.process(prm, kps_np, desc_np, results)
    barrier()
    gathered = None
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(a.steps):
        pipe.process(prm, kps_np, desc_np, results)   # H2D (pipelined under the matcher) + run + D2H + sync
        if world > 1:                         # the only collective: per-pair poses -> every rank (NCCL)
            T = torch.from_numpy(np.ascontiguousarray(results["T"])).cuda(non_blocking=True)
            gathered = torch.empty((world * T.shape[0],) + tuple(T.shape[1:]), dtype=T.dtype, device="cuda")
            dist.all_gather_into_tensor(gathered, T)
    e3.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    ms_e2e = max(e2.elapsed_time(e3), wall * 1e3) / a.steps      # host-side staging counts too
    ms_e2e = max_over_ranks(ms_e2e)
    e2e_value = world * P / (ms_e2e * 1e-3)
    clk = clocks.stop()

    # ---------------- roofline of the dominant kernel (the matcher) -------------------------
    ms_match = stage_acc[7] / a.steps                        # tile kernel alone, summed over chunk launches
    n_chunk_launches = int(round(stage_acc[8] / a.steps))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
    # algorithmic bytes per pair: 32*(nq+nt) descriptor bytes in + 16*nq key/match bytes out
    bytes_per_pair = 32 * (kp + kp) + 16 * kp
    achieved_gbs = bytes_per_pair * P / (ms_match * 1e-3) / 1e9
    popc_peak = ctx.microbench(0)                            # POPC.32 thread-ops/s, measured now on this GPU
    lop_peak = ctx.microbench(4)
    # algorithmic POPC.32 per pair: nq*nt*(256/32) for plain Hamming; the Hamming2 bit-plane form
    # needs nq*nt*4 -- report against the instruction count the kernel actually needs (4)
    popc_per_pair = kp * kp * (4 if a.norm == "hamming2" else 8)
    achieved_popc = popc_per_pair * P / (ms_match * 1e-3)
    # DRAM traffic of one matcher launch over the default workload, from the `ncu --set full` capture in
    # profiles/r1_ncu_full_top_kernels.csv (dram__bytes_read.sum + dram__bytes_write.sum = 327.8 + 59.0 MB).
    # It is BELOW the algorithmic bytes because consecutive pairs share a frame (train set of pair i =
    # query set of pair i+1) and that frame is still in L2.  Only valid for the profiled shape.
    traffic = 386.8e6 if (a.frames == 4541 and a.kp == 2000 and n_chunk_launches == 1 and a.norm == "hamming2"
                          and a.match == "crosscheck") else None
    roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": "match_tile_kernel<8,%s%s>" % (a.norm.upper(), ",TOP2" if a.match == "ratio" else ""), "launches_per_step": n_chunk_launches,
                "algorithmic_bytes_per_launch": bytes_per_pair * P / max(n_chunk_launches, 1),
                "ms_per_step_in_kernel": ms_match,
                "note": "the matcher is bound by the integer pipes (POPC on XU, LOP3 on ALU), not HBM "
                        "(AI ~ 100 popc/byte, so the HBM fraction is < 1 % by construction): see pipe"}
    # POPC instructions the kernel ISSUES per descriptor pair: carry-save compression counts 4 (Hamming2 planes)
    # or 8 (plain Hamming) algorithmic words with 3 or 4 POPC
    issued_per_alg = (3.0 / 4.0) if a.norm == "hamming2" else (4.0 / 8.0)
    issued_popc = achieved_popc * issued_per_alg
    pipe_roof = {"bound": "popc32 (XU pipe) / LOP3 (ALU pipe)", "achieved": issued_popc / 1e9,
                 "peak": popc_peak / 1e9, "unit": "Gpopc/s (issued)", "frac": issued_popc / popc_peak,
                 "algorithmic_Gpopc_s": achieved_popc / 1e9, "algorithmic_frac": achieved_popc / popc_peak,
                 "alu_peak_Gops": lop_peak / 1e9,
                 "work": "nq*nt*%d POPC.32 per pair (%s), one direction + fused column minima" % (
                     (4, "Hamming2 on bit planes") if a.norm == "hamming2" else (8, "plain Hamming")),
                 "note": "frac = issued POPC.32 rate / POPC.32 peak measured in this run (epivo_microbench); the "
                         "algorithmic rate is higher because carry-save adders compress the words before counting; "
                         "ncu (profiles/r1_ncu_full_top_kernels.csv): XU (POPC) pipe 88 %, ALU (LOP3) pipe 89 % of peak"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32-popc+f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "pairs_per_gpu": P, "l2": "inputs (363 MB/GPU) larger than L2",
                       "parallelism": f"pairs sharded, {world} x 1 GPU, no data-path collective",
                       "host_affinity": (f"rank bound to {near_cpus} GPU-local cores (NVML)" if near_cpus else "unchanged")},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(kps_np.nbytes + desc_np.nbytes),
                    "d2h_bytes_per_step": int(results.nbytes), "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "roofline": roofline, "pipe": pipe_roof,
            "stages_ms_per_step": {k: float(stage_acc[i] / a.steps) for i, k in
                                   enumerate(["total", "match", "presolve", "essential", "pose", "lm", "finish",
                                              "match_kernel"])},
            "quality": {"mean_matches": float(results["n_matches"].mean()),
                        "mean_inliers": float(results["n_inliers"].mean()),
                        "mean_good": float(results["n_good"].mean()),
                        "mean_ransac_iters": float(results["ransac_iters"].mean()),
                        "lm_ran_frac": float(results["lm_ran"].mean()),
                        "lm_reverted_frac": float(results["lm_reverted"].mean()),   # kitti_E.cpp:198-200 reverts when r_norm > 1e-9
                        "median_rot_err_rad": float(np.median([np.arccos(np.clip((np.trace(results["R"][i].T @ seq.R[i]) - 1) / 2, -1, 1))
                                                               for i in range(0, P, max(1, P // 256))]))}}

    if world == 1 and rank == 0 and not a.no_cpu_baseline:
        from oracle import cpu_reference as R
        n = max(8, min(a.cpu_pairs, P))
        cores = os.cpu_count() or 1
        pool = R.CpuPool(seq.kps[:n + 1], seq.descs[:n + 1], seq.K, 8 if a.method == "ransac" else 4, 0.99, a.thr,
                         cores=cores, norm=7 if a.norm == "hamming2" else 6, ratio=None if a.match == "crosscheck" else 0.8)
        pool.run(range(min(n, cores)))
        v, res = pool.run(range(n))
        pool.close()
        # SURVEY 8d's other arrangement: one process, OpenCV's own threads over all cores (16 pairs, ~2 s)
        n1 = min(16, n)
        v1 = R.single_process_rate(seq.kps[:n1 + 1], seq.descs[:n1 + 1], seq.K, n1, 8 if a.method == "ransac" else 4, 0.99,
                                   a.thr, threads=cores, norm=7 if a.norm == "hamming2" else 6,
                                   ratio=None if a.match == "crosscheck" else 0.8)
        line["cpu_baseline_single_process"] = {"value": v1, "unit": UNIT, "cores": cores,
                                               "sample": f"first {n1} pairs, one process, cv2.setNumThreads({cores})"}
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference" if R.HAVE_CV2 else "port",
                                "sample": f"first {n} pairs of the same sequence; cv2 "
                                          f"{R.cv2.__version__ if R.HAVE_CV2 else 'missing'} BFMatcher/findEssentialMat/"
                                          f"recoverPose + plain-C restatement of the reference LM; {cores} processes x 1 thread"}
    if rank == 0:
        emit(line)
    pipe.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
