"""CPU tests of the multi-process plumbing (gloo, world_size 2): the pair partition and
the gather of per-rank match lists into global pair order."""
import os
import socket

import numpy as np
import pytest

from orthosfm_b200 import distributed as osd
from orthosfm_b200 import synth


def test_partition_covers_every_pair_once_and_balances():
    pairs = synth.all_pairs(17)
    sizes = np.full(17, 8192)
    for world in (1, 2, 3, 8):
        parts = osd.partition_pairs(pairs, sizes, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(pairs)))
        lens = [len(p) for p in parts]
        assert max(lens) - min(lens) <= 1
    # unequal views: cost-balanced within one largest pair
    sizes = np.array([100, 5000, 30000, 7, 12000, 800, 20000, 64, 9000])
    pairs = synth.all_pairs(len(sizes))
    cost = sizes[pairs[:, 0]] * sizes[pairs[:, 1]]
    parts = osd.partition_pairs(pairs, sizes, 4)
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(len(pairs)))
    loads = np.array([cost[p].sum() for p in parts])
    assert loads.max() - loads.min() <= cost.max()


def _fake_list(pair_index: int) -> np.ndarray:
    k = (pair_index * 7) % 5  # some pairs have empty lists
    return np.stack([np.arange(k) * 3 + pair_index, np.arange(k) + 100 * pair_index], axis=1).astype(np.int32)


def _worker(rank: int, world: int, port: int, npairs: int, out_path: str):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pairs = synth.all_pairs(int((1 + (1 + 8 * npairs) ** 0.5) / 2))
    assert len(pairs) == npairs
    owned = osd.partition_pairs(pairs, np.full(64, 10), world)[rank]
    lists = [_fake_list(int(p)) for p in owned]
    offs = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    flat = np.concatenate(lists + [np.zeros((3, 2), np.int32)])  # buffer longer than the payload
    ij, offsets = osd.gather_match_lists(torch.from_numpy(flat), offs, owned, npairs, dst=0)
    if rank == 0:
        np.savez(out_path, ij=ij.numpy(), offsets=offsets)
    else:
        assert ij is None and offsets is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_match_lists_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    npairs = 28
    out_path = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, port, npairs, out_path), nprocs=2, join=True)
    z = np.load(out_path)
    ij, offsets = z["ij"], z["offsets"]
    assert offsets.shape == (npairs + 1,)
    for p in range(npairs):
        assert np.array_equal(ij[offsets[p]:offsets[p + 1]], _fake_list(p)), p
    assert offsets[-1] == ij.shape[0]


def test_gather_single_process_passthrough():
    import torch
    owned = np.array([0, 1, 2])
    lists = [_fake_list(p) for p in owned]
    offs = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    flat = torch.from_numpy(np.concatenate(lists))
    ij, offsets = osd.gather_match_lists(flat, offs, owned, 3)
    assert np.array_equal(offsets, offs) and ij.shape[0] == offs[-1]


def _worker_fixed(rank: int, world: int, port: int, npairs: int, out_path: str):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pairs = synth.all_pairs(int((1 + (1 + 8 * npairs) ** 0.5) / 2))
    all_owned = osd.partition_pairs(pairs, np.full(64, 10), world)
    owned = all_owned[rank]
    fg = osd.FixedGather(all_owned, npairs, capacity=64, device=torch.device("cpu"))
    for step in range(2):          # the buffers are reused step after step
        lists = [_fake_list(int(p) + step) for p in owned]
        offs = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
        flat = np.concatenate(lists + [np.zeros((0, 2), np.int32)])
        fg.out_ij[:len(flat)] = torch.from_numpy(flat)         # what the matcher would have written
        ij, start, count = fg.gather(offs)
        if rank == 0:
            np.savez(out_path + str(step) + ".npz", ij=ij.numpy(), start=start, count=count)
        else:
            assert ij is None
        dist.barrier()
    dist.destroy_process_group()


def test_fixed_gather_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    npairs = 28
    out_path = str(tmp_path / "fixed")
    mp.spawn(_worker_fixed, args=(2, port, npairs, out_path), nprocs=2, join=True)
    for step in range(2):
        z = np.load(out_path + str(step) + ".npz")
        for p in range(npairs):
            want = _fake_list(p + step)
            got = z["ij"][z["start"][p]:z["start"][p] + z["count"][p]]
            assert np.array_equal(got, want), (step, p)


def _worker_list(rank: int, world: int, port: int, npairs: int, out_path: str):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pairs = synth.all_pairs(int((1 + (1 + 8 * npairs) ** 0.5) / 2))
    all_owned = osd.partition_pairs(pairs, np.full(64, 10), world)
    owned = all_owned[rank]
    lg = osd.ListGather(all_owned, npairs, capacity=64, device=torch.device("cpu"))
    for step in range(3):          # the buffers are reused step after step; step 2: every list empty
        lists = [_fake_list(int(p) + step)[:0 if step == 2 else None] for p in owned]
        offs = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
        flat = np.concatenate(lists + [np.zeros((0, 2), np.int32)])
        lg.out_ij[:len(flat)] = torch.from_numpy(flat)         # what the matcher would have written
        lg.out_ij[len(flat):] = -7                             # must never travel
        ij, start, count = lg.gather(offs)
        if rank == 0:
            host = torch.full((lg.used_rows() + 1, 2), -1, dtype=torch.int32)
            hstart = lg.to_host(host, start)
            np.savez(out_path + str(step) + ".npz", ij=ij.numpy(), start=start, count=count,
                     host=host.numpy(), hstart=hstart, used=lg.used_rows())
        else:
            assert ij is None
        dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_list_gather_gloo(tmp_path, world):
    """The exact-size gather bench.py uses for N > 1: only header + used payload travel."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    npairs = 28
    out_path = str(tmp_path / "list")
    mp.spawn(_worker_list, args=(world, port, npairs, out_path), nprocs=world, join=True)
    for step in range(3):
        z = np.load(out_path + str(step) + ".npz")
        total = 0
        for p in range(npairs):
            want = _fake_list(p + step)[:0 if step == 2 else None]
            got = z["ij"][z["start"][p]:z["start"][p] + z["count"][p]]
            assert np.array_equal(got, want), (step, p)
            got_h = z["host"][z["hstart"][p]:z["hstart"][p] + z["count"][p]]
            assert np.array_equal(got_h, want), (step, p)
            total += len(want)
        assert int(z["used"]) == total
        assert not (z["host"][:total] == -7).any()


def test_list_gather_single_process():
    import torch
    all_owned = [np.arange(5)]
    lg = osd.ListGather(all_owned, 5, capacity=32, device=torch.device("cpu"))
    lists = [_fake_list(p) for p in range(5)]
    offs = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    flat = np.concatenate(lists)
    lg.out_ij[:len(flat)] = torch.from_numpy(flat)
    ij, start, count = lg.gather(offs)
    for p in range(5):
        assert np.array_equal(ij[start[p]:start[p] + count[p]].numpy(), lists[p])
