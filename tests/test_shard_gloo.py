"""Multi-GPU host logic on CPU: contiguous pair sharding, the single pose all-gather (gloo,
world_size 2 and 3) and the sequential pose chaining that follows it (kitti_E.cpp:218-228)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from epivo_b200 import shard, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_pose(i):
    """a deterministic pose per pair index, so every rank can fabricate its own block"""
    T = np.eye(4)
    T[:3, :3] = synth.rodrigues(np.array([0.01 * np.sin(i), 0.02 * np.cos(i), 0.005 * i % 0.03]))
    T[:3, 3] = [0.1 * np.cos(i), -0.05, 1.0 + 0.01 * (i % 7)]
    return T


def _worker(rank, world, port, n_pairs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, stop = shard.shard_range(n_pairs, world, rank)
    local = np.stack([_fake_pose(i) for i in range(start, stop)]) if stop > start else np.zeros((0, 4, 4))
    allT = shard.gather_poses(local, n_pairs, world, rank, dist=dist, device="cpu")
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), allT)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_pairs", [(2, 11), (2, 4540), (3, 10)])
def test_gather_is_in_sequence_order(tmp_path, world, n_pairs):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_pairs, str(tmp_path)), nprocs=world, join=True)
    want = np.stack([_fake_pose(i) for i in range(n_pairs)])
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert got.shape == (n_pairs, 4, 4)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("n_pairs,world", [(4540, 1), (4540, 2), (4540, 4), (4540, 8), (5, 8), (0, 2)])
def test_shard_ranges_partition_the_sequence(n_pairs, world):
    covered = []
    for r in range(world):
        a, b = shard.shard_range(n_pairs, world, r)
        assert 0 <= a <= b <= n_pairs
        covered.extend(range(a, b))
        f0, f1 = shard.frames_for(a, b)
        assert f1 - f0 == (b - a + 1 if b > a else 0)       # halo of one frame
    assert covered == list(range(n_pairs))
    sizes = [shard.shard_range(n_pairs, world, r) for r in range(world)]
    assert max(b - a for a, b in sizes) == -(-n_pairs // world) or n_pairs == 0


def test_pose_chaining_matches_reference_loop():
    """kitti_E.cpp:218-228 restated inline: dT = [R | t/|t| * scale], cT = cT * dT^-1."""
    rng = np.random.default_rng(0)
    n = 20
    T = np.stack([_fake_pose(i) for i in range(n)])
    scales = rng.uniform(0.5, 1.5, n)
    got = shard.chain_poses(T, scales)
    cT = np.eye(4)
    for i in range(n):
        assert np.allclose(got[i], cT)
        dT = np.eye(4)
        dT[:3, :3] = T[i][:3, :3]
        dT[:3, 3] = T[i][:3, 3] / np.linalg.norm(T[i][:3, 3]) * scales[i]
        cT = cT @ np.linalg.inv(dT)
    assert np.allclose(got[n], cT)


def test_pose_file_format_roundtrip(tmp_path):
    """the viewers read the pose files with np.fromfile(sep=' ').reshape(-1, 4, 4) (cloud_pango.py:32-34)"""
    T = np.stack([_fake_pose(i) for i in range(5)])
    p = tmp_path / "kitti.T"
    shard.write_poses(str(p), T)
    back = np.fromfile(str(p), sep=" ").reshape(-1, 4, 4)
    assert np.array_equal(back, T)
