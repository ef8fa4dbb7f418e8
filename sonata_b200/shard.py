"""Utterance-level sharding across the GPUs of one box.

Utterances are independent (the reference already treats sentences as independent tasks,
synth/src/lib.rs:316-320), so the path shards with NO data-path collective: the only exchanges are
the ragged scatter of ids from rank 0 and the ragged gather of waveforms back to rank 0, done with
torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).  Weights are replicated.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def lpt_partition(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time greedy: sort by cost descending, give each to the least loaded rank.
    Cost ~ T_x (frames ~ 3 T_x).  Returns, per rank, utterance indices in ascending order."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += costs[i]
    return [sorted(p) for p in parts]


def _dev(group=None) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def scatter_ids(batches: Optional[Sequence[np.ndarray]], group=None) -> List[np.ndarray]:
    """Rank 0 passes the full list of id sequences; every rank returns its own shard (LPT).
    Also returns nothing else: global indices are recovered by gather_waveforms."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _dev(group)
    if rank == 0:
        lens = np.array([len(b) for b in batches], dtype=np.int64)
        parts = lpt_partition(lens.tolist(), world)
        owner = np.zeros(len(batches), dtype=np.int64)
        for r, p in enumerate(parts):
            owner[p] = r
        meta = torch.tensor([len(batches)], dtype=torch.int64, device=dev)
    else:
        meta = torch.zeros(1, dtype=torch.int64, device=dev)
    dist.broadcast(meta, 0, group=group)
    n = int(meta.item())
    table = torch.zeros(2, n, dtype=torch.int64, device=dev)
    if rank == 0:
        table[0] = torch.from_numpy(lens).to(dev)
        table[1] = torch.from_numpy(owner).to(dev)
    dist.broadcast(table, 0, group=group)
    lens_all = table[0].cpu().numpy()
    owner_all = table[1].cpu().numpy()
    mine = np.nonzero(owner_all == rank)[0]
    per_rank_tot = [int(lens_all[owner_all == r].sum()) for r in range(world)]
    cap = max(max(per_rank_tot), 1)
    recv = torch.zeros(cap, dtype=torch.int64, device=dev)
    if rank == 0:
        sl = []
        for r in range(world):
            buf = torch.zeros(cap, dtype=torch.int64)
            idx = np.nonzero(owner_all == r)[0]
            if len(idx):
                buf[:per_rank_tot[r]] = torch.from_numpy(np.concatenate([np.asarray(batches[i], dtype=np.int64) for i in idx]))
            sl.append(buf.to(dev))
        dist.scatter(recv, sl, src=0, group=group)
    else:
        dist.scatter(recv, None, src=0, group=group)
    flat = recv.cpu().numpy()
    out, o = [], 0
    for i in mine:
        out.append(flat[o:o + int(lens_all[i])].copy())
        o += int(lens_all[i])
    scatter_ids.last = {"mine": mine, "owner": owner_all, "n": n}
    return out


def gather_waveforms(local_wave: torch.Tensor, local_lens: Sequence[int], group=None, to_host: bool = True):
    """`local_wave`: this rank's utterances concatenated (device tensor for NCCL, 1-D float32);
    `local_lens`: samples per local utterance.  Rank 0 returns the waveforms in GLOBAL utterance order."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = _dev(group)
    info = scatter_ids.last
    n, owner, mine = info["n"], info["owner"], info["mine"]
    lens = torch.zeros(n, dtype=torch.int64, device=dev)
    if len(mine):
        lens[torch.from_numpy(mine).to(dev)] = torch.tensor(list(local_lens), dtype=torch.int64, device=dev)
    dist.all_reduce(lens, group=group)
    lens_all = lens.cpu().numpy()
    tot = [int(lens_all[owner == r].sum()) for r in range(world)]
    cap = max(max(tot), 1)
    send = torch.zeros(cap, dtype=torch.float32, device=dev)
    send[:tot[rank]] = local_wave[:tot[rank]].to(dev)
    if rank == 0:
        gl = [torch.empty(cap, dtype=torch.float32, device=dev) for _ in range(world)]
        dist.gather(send, gl, dst=0, group=group)
        if not to_host:          # device-resident result: per-rank buffers + the tables to slice them
            return gl, owner, lens_all
        out: List[Optional[np.ndarray]] = [None] * n
        for r in range(world):
            flat = gl[r].cpu().numpy()
            o = 0
            for i in np.nonzero(owner == r)[0]:
                out[i] = flat[o:o + int(lens_all[i])].copy()
                o += int(lens_all[i])
        return out  # type: ignore[return-value]
    dist.gather(send, None, dst=0, group=group)
    return None


class HostGather:
    """Gather into ONE host buffer without funnelling every GPU's audio through rank 0's PCIe link.

    A POSIX shared-memory segment (grown on demand, reused across calls) is mapped by every rank; each rank copies its
    own waveforms device->host into its slice -- N links in parallel -- and rank 0 reads all of them from the same
    pages.  The collectives carry only the length table and the segment's name.  With `pin=True` on a CUDA build the
    mapping is page-locked (`cudaHostRegister`) so the copies are DMA transfers.  (NCCL gather + one copy reached
    80k audio-s/s at N = 8 where per-rank host buffers reach 174k: profiles/notes_r01.md.)
    """

    def __init__(self, group=None, pin: bool = True):
        self.group, self.pin = group, pin
        self.shm = None
        self.cap = 0
        self._registered = None

    def _ensure(self, nbytes: int):
        from multiprocessing import shared_memory
        rank = dist.get_rank(self.group)
        need = torch.tensor([int(nbytes > self.cap)], dtype=torch.int64, device=_dev(self.group))
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
        if int(need.item()) == 0:
            return
        self.close(unlink=rank == 0)
        cap = max(int(nbytes * 1.25), 1 << 20)
        name = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=cap)
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=0, group=self.group)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
        self.cap = cap
        if self.pin and torch.cuda.is_available() and dist.get_backend(self.group) == "nccl":
            buf = np.frombuffer(self.shm.buf, dtype=np.uint8)
            if int(torch.cuda.cudart().cudaHostRegister(buf.ctypes.data, cap, 0)) == 0:
                self._registered = buf.ctypes.data     # on failure the copies still work, through pageable memory
        dist.barrier(group=self.group)

    def gather(self, local_wave: torch.Tensor, local_lens: Sequence[int]):
        """Same contract as `gather_waveforms(..., to_host=True)`: rank 0 returns the waveforms in GLOBAL utterance
        order (views into the shared segment, valid until the next call), other ranks return None."""
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        dev = _dev(self.group)
        info = scatter_ids.last
        n, owner, mine = info["n"], info["owner"], info["mine"]
        lens = torch.zeros(n, dtype=torch.int64, device=dev)
        if len(mine):
            lens[torch.from_numpy(mine).to(dev)] = torch.tensor(list(local_lens), dtype=torch.int64, device=dev)
        dist.all_reduce(lens, group=self.group)
        lens_all = lens.cpu().numpy()
        offs = np.zeros(n + 1, dtype=np.int64)
        order = [i for r in range(world) for i in np.nonzero(owner == r)[0]]      # rank-major layout of the segment
        pos = 0
        for i in order:
            offs[i] = pos
            pos += int(lens_all[i])
        self._ensure(pos * 4)
        seg = np.frombuffer(self.shm.buf, dtype=np.float32, count=self.cap // 4)
        my_tot = int(lens_all[mine].sum()) if len(mine) else 0
        if my_tot:
            start = int(offs[mine[0]])                                            # a rank's utterances are contiguous
            torch.from_numpy(seg[start:start + my_tot]).copy_(local_wave[:my_tot])
        if dev.type == "cuda":
            torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if rank != 0:
            return None
        return [seg[int(offs[i]):int(offs[i]) + int(lens_all[i])] for i in range(n)]

    def close(self, unlink: bool = False):
        if self.shm is None:
            return
        if self._registered is not None:
            torch.cuda.cudart().cudaHostUnregister(self._registered)
            self._registered = None
        try:
            self.shm.close()
        except BufferError:
            pass                                   # views handed out by gather() are still alive
        if unlink:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass
        self.shm, self.cap = None, 0


def sharded_synthesize(batches: Optional[Sequence[np.ndarray]], synth: Callable[[List[np.ndarray]], List[np.ndarray]],
                       group=None) -> Optional[List[np.ndarray]]:
    """scatter -> local synthesis (`synth`: list of id arrays -> list of float32 waveforms) -> gather."""
    mine = scatter_ids(batches, group)
    waves = synth(mine) if mine else []
    dev = _dev(group)
    if waves:
        local = torch.from_numpy(np.concatenate(waves)).to(dev)
    else:
        local = torch.zeros(0, dtype=torch.float32, device=dev)
    return gather_waveforms(local, [len(w) for w in waves], group)
