"""Multi-GPU sharding of a sequence and the host-side pose chaining that follows the hot path.

Frame pairs are independent (the reference loop kitti_E.cpp:54-201 carries no state between
iterations until the pose chaining at :225-228), so a sequence is split into contiguous blocks
of pairs, one block per rank (contiguous so that every frame is uploaded to one GPU only, with a
halo of one frame), with NO collective on the data path.  The only exchange is one all-gather of
the fixed-size per-pair pose records at the end (NCCL on GPUs, gloo in the CPU tests), after
which rank 0 does the inherently sequential chaining cT <- cT * dT^-1 (kitti_E.cpp:218-228).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_pairs: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of pair indices for `rank`: blocks of
    ceil(n_pairs / world), the last ranks possibly short or empty."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-n_pairs // world)
    start = min(rank * per, n_pairs)
    return start, min(start + per, n_pairs)


def frames_for(start: int, stop: int) -> tuple[int, int]:
    """Frames a rank needs for pairs [start, stop): [start, stop] inclusive, i.e. a halo of one."""
    return (start, stop + 1) if stop > start else (start, start)


def gather_poses(local_T: np.ndarray, n_pairs: int, world: int, rank: int, dist=None, device=None) -> np.ndarray:
    """All-gather the per-pair 4x4 poses of every rank into sequence order.

    local_T: (stop-start, 4, 4) float64 for this rank's block.  Every rank contributes a block
    padded to ceil(n_pairs/world) records so that a plain all_gather works for ragged tails.
    `dist` is torch.distributed (None when world == 1); `device` the tensor device to stage on
    ("cuda" for NCCL, "cpu" for gloo)."""
    start, stop = shard_range(n_pairs, world, rank)
    assert local_T.shape == (stop - start, 4, 4)
    if world == 1 or dist is None:
        return np.ascontiguousarray(local_T, dtype=np.float64)
    import torch
    per = -(-n_pairs // world)
    buf = np.zeros((per, 4, 4), dtype=np.float64)
    buf[:stop - start] = local_T
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * per, 4, 4), dtype=t.dtype, device=t.device)     # concatenated along dim 0
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy()[:n_pairs]


def chain_poses(T_pairs: np.ndarray, scales: np.ndarray | None = None) -> np.ndarray:
    """Host pose chaining of kitti_E.cpp:218-228.

    T_pairs[i] is the refined point transform of pair i (frame i -> i+1).  The reference
    normalises its translation, rescales it by the ground-truth step length (`scale`,
    kitti_E.cpp:220-222) and accumulates the camera pose cT <- cT * dT^-1 starting from identity;
    all_T[i] is the pose BEFORE pair i is applied (kitti_E.cpp:227), so n+1 poses come back."""
    n = T_pairs.shape[0]
    scales = np.ones(n) if scales is None else np.asarray(scales, dtype=np.float64)
    out = np.empty((n + 1, 4, 4))
    out[0] = np.eye(4)
    if n == 0:
        return out
    T_pairs = np.asarray(T_pairs, dtype=np.float64)
    dT = np.tile(np.eye(4), (n, 1, 1))
    dT[:, :3, :3] = T_pairs[:, :3, :3]
    t = T_pairs[:, :3, 3]
    with np.errstate(all="ignore"):           # |t| = 0 is not guarded: NaN from there on, as in the reference (:221)
        dT[:, :3, 3] = t / np.linalg.norm(t, axis=1, keepdims=True) * scales[:, None]   # and in cloud.cu: scaled_dT
        bad = ~np.isfinite(dT).all(axis=(1, 2))
        dT[bad] = np.eye(4)                   # keep LAPACK away from NaN; the NaN is put back below
        inv = np.linalg.inv(dT)               # the reference inverts the general 4x4 (:228)
        inv[bad] = np.nan
        # cT_i = inv_0 inv_1 ... inv_{i-1}: an inclusive scan of matrix products by doubling (4540 pairs: 13 batched
        # products instead of a 4540-step Python loop; the matrix product is associative, so this differs from the
        # sequential loop by rounding only -- as the device block scan of cloud.cu does)
        acc = inv.copy()
        o = 1
        while o < n:
            acc[o:] = acc[:-o] @ acc[o:]
            o *= 2
    out[1:] = acc
    return out


def write_poses(path: str, poses: np.ndarray) -> None:
    """kitti.T / kitti.GT / euroc.T text format (kitti_E.cpp:271-286); see epivo_b200.io."""
    from . import io
    io.write_poses(path, poses)
