"""Multi-GPU plumbing for the matcher: one process per GPU, image pairs partitioned across
ranks, descriptor pool replicated with one broadcast, match lists gathered to rank 0.

The reference has no multi-GPU or multi-process code at all; its only
parallelism is an OpenMP loop over the flat pair index
(src/mve/sfm/bundler_matching.cc:74-132) whose iterations are independent.  That is the
unit sharded here.  There is no collective on the data path: the broadcast happens once
before matching and the gather once after it.

Works with NCCL (GPU tensors) and gloo (CPU tensors; used by the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def partition_pairs(pairs: np.ndarray, sizes: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Deterministic cost-balanced partition of the pair list.

    Cost of a pair is n1*n2 (the matcher's work is exactly proportional to it).  Pairs are
    visited in the reference's enumeration order and handed to the least-loaded rank
    (ties: lowest rank), so equal-size views degenerate to round-robin.  Returns, per rank,
    the indices into ``pairs`` it owns (ascending)."""
    pairs = np.asarray(pairs, np.int64).reshape(-1, 2)
    sizes = np.asarray(sizes, np.int64)
    cost = sizes[pairs[:, 0]] * sizes[pairs[:, 1]] if pairs.size else np.zeros(0, np.int64)
    owned: List[List[int]] = [[] for _ in range(world_size)]
    if pairs.shape[0] == 0:
        return [np.zeros(0, np.int64) for _ in range(world_size)]
    if np.all(cost == cost[0]):
        for r in range(world_size):
            owned[r] = list(range(r, pairs.shape[0], world_size))
    else:
        load = np.zeros(world_size, np.int64)
        for i in range(pairs.shape[0]):
            r = int(np.argmin(load))
            owned[r].append(i)
            load[r] += cost[i]
    return [np.asarray(o, np.int64) for o in owned]


def broadcast_pool(pool, src: int = 0):
    """Replicates the packed descriptor pool (uint8 tensor [rows, 128]) from ``src`` to
    every rank (NCCL broadcast over NVLink on the GPU box)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(pool, src=src)
    return pool


def gather_match_lists(local_ij, local_offsets: np.ndarray, owned: np.ndarray, npairs: int,
                       dst: int = 0, all_owned: Sequence[np.ndarray] | None = None) -> Tuple[object, np.ndarray]:
    """Gathers every rank's compacted (i, j) lists to ``dst`` and orders them by global
    pair index.

    local_ij      int32 tensor [>= local_offsets[-1], 2] on this rank's device
    local_offsets int64 array, len(owned) + 1
    owned         global pair indices of this rank's lists (ascending)
    all_owned     every rank's ``owned`` (the partition is deterministic, so callers normally
                  pass ``partition_pairs(...)``; if omitted it is exchanged once)
    Returns (ij, offsets) on ``dst`` -- ij int32 [total, 2], offsets int64 [npairs + 1] --
    and (None, None) on the other ranks.

    Two collectives: one all-gather of the (padded) per-pair list lengths, one gather of the
    (padded) payloads; the reorder into pair order is a single indexed copy on the device."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        total = int(local_offsets[-1])
        offsets = np.zeros(npairs + 1, np.int64)
        counts = np.zeros(npairs, np.int64)
        counts[owned] = np.diff(local_offsets)
        offsets[1:] = np.cumsum(counts)
        return local_ij[:total], offsets

    rank, world = dist.get_rank(), dist.get_world_size()
    dev = local_ij.device
    if all_owned is None:
        gathered = [None] * world
        dist.all_gather_object(gathered, np.asarray(owned, np.int64))
        all_owned = gathered
    max_owned = max(len(o) for o in all_owned)

    # 1. per-pair list lengths of every rank
    mine = torch.zeros(max_owned, dtype=torch.int64, device=dev)
    if len(owned):
        mine[:len(owned)] = torch.as_tensor(np.diff(local_offsets), device=dev)
    lens_all = torch.empty(world * max_owned, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(lens_all, mine)
    lens_h = lens_all.cpu().numpy().reshape(world, max_owned)
    counts = np.zeros(npairs, np.int64)
    totals = np.zeros(world, np.int64)
    for r in range(world):
        counts[all_owned[r]] = lens_h[r, :len(all_owned[r])]
        totals[r] = lens_h[r, :len(all_owned[r])].sum()
    offsets = np.zeros(npairs + 1, np.int64)
    offsets[1:] = np.cumsum(counts)

    # 2. payloads, padded to the largest rank total
    pad = int(max(int(totals.max()), 1))
    send = local_ij[:pad] if local_ij.shape[0] >= pad else torch.cat(
        [local_ij, torch.zeros((pad - local_ij.shape[0], 2), dtype=torch.int32, device=dev)])
    send = send.contiguous()
    if rank == dst:
        recv = [torch.empty((pad, 2), dtype=torch.int32, device=dev) for _ in range(world)]
        dist.gather(send, recv, dst=dst)
    else:
        dist.gather(send, None, dst=dst)
        return None, None

    # 3. rank-major lists -> global pair order, one indexed copy per rank
    out = torch.empty((int(offsets[-1]), 2), dtype=torch.int32, device=dev)
    for r in range(world):
        if totals[r] == 0:
            continue
        lens = torch.as_tensor(counts[all_owned[r]], device=dev)
        dst_off = torch.as_tensor(offsets[all_owned[r]], device=dev)
        src_off = torch.cumsum(lens, 0) - lens
        idx = torch.repeat_interleave(dst_off - src_off, lens) + torch.arange(int(totals[r]), device=dev)
        out[idx] = recv[r][:int(totals[r])]
    return out, offsets


class FixedGather:
    """One-collective gather of match lists for a fixed pair partition.

    ``gather_match_lists`` needs the list lengths on every rank before it can size the payload
    exchange (an all-gather, a host round trip, then the gather) and reorders the lists into
    global pair order with a handful of small kernels per rank.  When the same partition is
    matched again and again (the benchmark's step; a pipeline that re-matches after adding
    views), all of that can be laid out once: every rank owns one buffer

        [ max_owned rows of header : per-pair list lengths (int64, viewed as int32 pairs) |
          capacity rows of payload : the rank's compacted (i, j) lists, back to back        ]

    whose payload part is handed to the matcher as its output, so a step is one
    ``dist.gather`` of equal-sized buffers plus one small device-to-host copy of the headers on
    the destination.  The lists stay in rank-major order; ``start[p]`` / ``count[p]`` locate pair
    ``p`` in the flattened result (no reorder: consumers walk the pair list anyway,
    src/mve/sfm/bundler_matching.cc:74-132)."""

    def __init__(self, all_owned: Sequence[np.ndarray], npairs: int, capacity: int, device, dst: int = 0):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dst = dst
        self.all_owned = [np.asarray(o, np.int64) for o in all_owned]
        self.npairs = npairs
        self.max_owned = max(1, max(len(o) for o in self.all_owned))
        self.capacity = int(capacity)
        self.rows = self.max_owned + self.capacity
        self.buf = torch.zeros((self.rows, 2), dtype=torch.int32, device=device)
        self.out_ij = self.buf[self.max_owned:]           # what the matcher writes into
        self.header = self.buf[:self.max_owned].view(torch.int64).view(-1)     # [max_owned]
        pin = device.type == "cuda"
        self.lens_host = torch.zeros(self.max_owned, dtype=torch.int64, pin_memory=pin)
        self.recv = None
        if self.rank == dst:
            self.recv = torch.empty((self.world, self.rows, 2), dtype=torch.int32, device=device)
            self.recv_list = [self.recv[r] for r in range(self.world)]
            self.hdr_host = torch.zeros((self.world, self.max_owned), dtype=torch.int64, pin_memory=pin)

    def gather(self, local_offsets: np.ndarray):
        """local_offsets: this rank's list offsets (len(owned) + 1) as the matcher returned them.
        Returns (flat_ij, start, count) on the destination -- flat_ij int32 [world * rows, 2] --
        and (None, None, None) elsewhere."""
        import torch
        n = len(local_offsets) - 1
        self.lens_host.zero_()
        if n:
            self.lens_host[:n] = torch.from_numpy(np.diff(local_offsets))
        self.header.copy_(self.lens_host, non_blocking=True)
        if self.world == 1:
            lens = self.lens_host.numpy()[None, :]
        else:
            if self.rank == self.dst:
                self.dist.gather(self.buf, self.recv_list, dst=self.dst)
            else:
                self.dist.gather(self.buf, None, dst=self.dst)
                if self.buf.is_cuda:
                    torch.cuda.current_stream().synchronize()    # the payload has left: the next step may overwrite it
                return None, None, None
            self.hdr_host.copy_(self.recv[:, :self.max_owned].reshape(self.world, -1).view(torch.int64),
                                non_blocking=False)
            lens = self.hdr_host.numpy()
        start = np.zeros(self.npairs, np.int64)
        count = np.zeros(self.npairs, np.int64)
        for r in range(self.world):
            o = self.all_owned[r]
            c = lens[r, :len(o)]
            count[o] = c
            start[o] = r * self.rows + self.max_owned + np.cumsum(c) - c
        flat = (self.recv if self.world > 1 else self.buf[None]).view(-1, 2)
        return flat, start, count


class ListGather:
    """Gather of the per-pair match lists to one rank that moves only what was produced.

    Every rank owns one buffer ``[header | payload]``: the matcher writes its compacted (i, j)
    lists straight into the payload part (``out_ij``), the header holds the per-pair list
    lengths.  A step is: one tiny all-gather of the used sizes, then one point-to-point message
    per rank carrying exactly ``header + used payload`` (``FixedGather`` moves the whole
    capacity).  The lists stay rank-major on the destination; ``start[p]`` / ``count[p]`` locate
    pair ``p`` (consumers walk the pair list anyway, src/mve/sfm/bundler_matching.cc:74-132).
    ``gather`` returns after the transfers have completed on every rank, so the buffers may be
    rewritten by the next step at once."""

    def __init__(self, all_owned: Sequence[np.ndarray], npairs: int, capacity: int, device, dst: int = 0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dst = dst
        self.device = device
        self.all_owned = [np.asarray(o, np.int64) for o in all_owned]
        self.npairs = npairs
        self.max_owned = max(1, max(len(o) for o in self.all_owned))
        self.hdr_rows = (self.max_owned + 1) // 2            # int32 lengths, two per (i, j) row
        self.capacity = int(capacity)
        self.rows = self.hdr_rows + self.capacity
        self.cuda = getattr(device, "type", str(device)) == "cuda"
        self.buf = torch.zeros((self.rows, 2), dtype=torch.int32, device=device)
        self.out_ij = self.buf[self.hdr_rows:]               # what the matcher writes into
        self.header = self.buf[:self.hdr_rows].view(-1)      # [2 * hdr_rows] int32
        self.lens_host = torch.zeros(2 * self.hdr_rows, dtype=torch.int32, pin_memory=self.cuda)
        self.sizes_dev = torch.zeros(self.world, dtype=torch.int64, device=device)
        self.recv = None
        self.used = np.zeros(self.world, np.int64)           # payload rows received per rank (dst only)
        if self.rank == dst and self.world > 1:
            self.recv = torch.empty((self.world, self.rows, 2), dtype=torch.int32, device=device)
            self.hdr_host = torch.zeros((self.world, 2 * self.hdr_rows), dtype=torch.int32, pin_memory=self.cuda)

    def gather(self, local_offsets: np.ndarray):
        """local_offsets: this rank's list offsets (len(owned) + 1) as the matcher returned them.
        Returns (flat_ij, start, count) on the destination -- flat_ij int32 [world * rows, 2] --
        and (None, None, None) elsewhere."""
        torch, dist = self.torch, self.dist
        n = len(local_offsets) - 1
        total = int(local_offsets[-1]) if n >= 0 and len(local_offsets) else 0
        if total > self.capacity:
            raise ValueError(f"match lists need {total} rows, capacity {self.capacity}")
        self.lens_host.zero_()
        if n > 0:
            self.lens_host[:n] = torch.from_numpy(np.diff(local_offsets).astype(np.int32))
        self.header.copy_(self.lens_host, non_blocking=True)
        if self.world == 1:
            lens = self.lens_host.numpy()[None, :]
            self.used[0] = total
            flat = self.buf
        else:
            mine = torch.tensor([total], dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(self.sizes_dev, mine)
            used = self.sizes_dev.cpu().numpy()              # synchronises: the header copy above is done too
            send_rows = self.hdr_rows + total
            if self.rank == self.dst:
                ops = [dist.P2POp(dist.irecv, self.recv[r, :self.hdr_rows + int(used[r])], r)
                       for r in range(self.world) if r != self.dst]
                self.recv[self.dst, :send_rows].copy_(self.buf[:send_rows], non_blocking=True)
            else:
                ops = [dist.P2POp(dist.isend, self.buf[:send_rows], self.dst)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            if self.rank != self.dst:
                if self.cuda:
                    torch.cuda.current_stream().synchronize()    # the payload has left: the buffer is free
                return None, None, None
            self.used[:] = used
            self.hdr_host.copy_(self.recv[:, :self.hdr_rows].reshape(self.world, -1), non_blocking=False)
            lens = self.hdr_host.numpy()
            flat = self.recv.view(-1, 2)
        start = np.zeros(self.npairs, np.int64)
        count = np.zeros(self.npairs, np.int64)
        for r in range(self.world):
            o = self.all_owned[r]
            c = lens[r, :len(o)].astype(np.int64)
            count[o] = c
            start[o] = r * self.rows + self.hdr_rows + np.cumsum(c) - c
        return flat, start, count

    def used_rows(self) -> int:
        """Payload rows of the last gather (destination)."""
        return int(self.used.sum())

    def to_host(self, host_ij, start: np.ndarray):
        """Copies the gathered payloads (used parts only) back to back into ``host_ij`` (a pinned
        int32 [>= used_rows, 2] tensor) and returns the pairs' start offsets in it."""
        src = self.recv if self.world > 1 else self.buf[None]
        host_start = start.copy()
        at = 0
        for r in range(self.world):
            u = int(self.used[r])
            if u:
                host_ij[at:at + u].copy_(src[r, self.hdr_rows:self.hdr_rows + u], non_blocking=True)
            o = self.all_owned[r]
            host_start[o] = start[o] - (r * self.rows + self.hdr_rows) + at
            at += u
        return host_start
