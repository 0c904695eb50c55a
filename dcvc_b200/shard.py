"""Multi-GPU plumbing of the hot path: the path shards by job — (sequence x rate point) or
independent GOP segments — with NO data-path collective.  One process per GPU; rank 0's checkpoint is
broadcast once (NCCL over NVLink on the GPU box, gloo in the CPU tests); jobs are dealt round-robin in the
deterministic order of the reference's job list (test_video.py:527-564)."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist


def shard_jobs(jobs, rank: int, world: int):
    """jobs j with j % world == rank, in order."""
    return [j for i, j in enumerate(jobs) if i % world == rank]


def broadcast_state_dict(sd, spec, src: int = 0, device="cpu"):
    """Broadcast a checkpoint as ONE flat fp32 blob (a single collective).  `sd` is only read on `src`;
    every rank passes the same `spec` (name -> shape)."""
    numel = sum(int(np.prod(s)) for s in spec.values())
    blob = torch.empty(numel, dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        blob.copy_(torch.cat([sd[k].reshape(-1).float() for k in spec]))
    dist.broadcast(blob, src)
    host = blob.cpu()
    out, off = OrderedDict(), 0
    for k, s in spec.items():
        n = int(np.prod(s))
        out[k] = host[off:off + n].view(s).clone()
        off += n
    return out


def gather_results(local_results, dst: int = 0):
    """Host-side merge of per-job result dicts (the reference merges JSON on the host, test_video.py:566-586)."""
    world = dist.get_world_size()
    gathered = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(local_results, gathered, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = []
    for part in gathered:
        merged.extend(part)
    return merged
