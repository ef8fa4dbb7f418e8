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
