"""Multi-GPU plumbing of the hot path: the path shards by job — (sequence x rate point) or
independent GOP segments — with NO data-path collective.  One process per GPU; rank 0's checkpoint is
broadcast once (NCCL over NVLink on the GPU box, gloo in the CPU tests); jobs are dealt in rotated blocks
(shard_jobs) from the deterministic order of the reference's job list (test_video.py:527-564)."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist


def shard_jobs(jobs, rank: int, world: int):
    """This rank's share of a job list, in list order: every block of `world` consecutive jobs goes to `world` different
    ranks, and the blocks are dealt with a rotation — job i belongs to rank (i + i // world) % world.  The reference hands
    jobs to whichever worker is free (test_video.py:527-564, a process pool); a static deal has to balance by itself: its
    list is (sequence, rate point) with the rate point innermost, and a plain i % world gives a rank the same rate point of
    every sequence whenever the number of rate points divides (or is divided by) `world` — the ranks that drew the highest
    rate then carry 2-3x the entropy-coding work (measured at 8 ranks: 0.63 of the ideal aggregate encode FPS)."""
    return [j for i, j in enumerate(jobs) if (i + i // world) % world == rank]


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


# ---------------------------------------------------------------------------------------------- host placement
def _parse_cpulist(text: str):
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the kernel's cpulist format)."""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_cpus(pci_bus_id: str, sysfs_root: str = "/sys"):
    """CPUs of the NUMA node the GPU at `pci_bus_id` ('0000:1b:00.0') hangs off, or None when the platform does not
    say (numa_node -1: single node or a virtualised topology)."""
    import os
    dev = os.path.join(sysfs_root, "bus", "pci", "devices", pci_bus_id.lower())
    try:
        node = int(open(os.path.join(dev, "numa_node")).read().strip())
        if node < 0:
            return None
        return _parse_cpulist(open(os.path.join(sysfs_root, "devices", "system", "node", f"node{node}", "cpulist")).read()) or None
    except (OSError, ValueError):
        return None


def pin_to_gpu_numa_node(device_index: int, sysfs_root: str = "/sys"):
    """One process per GPU on a two-socket box: keep this process — and every thread it starts afterwards: the rANS
    fork-join pool is created with the codec handle, pinned host buffers are placed by first touch — on the socket the
    GPU is attached to, so the symbol hand-off (D2H indexes -> host rANS -> H2D symbols, five times per picture) never
    crosses the inter-socket link.  Call it before the first proxy is created.  Returns the CPU set applied, or None
    when nothing was changed (no topology information, no permission, or the set would be empty)."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{getattr(p, 'pci_domain_id'):04x}:{getattr(p, 'pci_bus_id'):02x}:{getattr(p, 'pci_device_id'):02x}.0"
    except Exception:  # noqa: BLE001 — older torch without the pci_* properties, or no device
        return None
    cpus = gpu_numa_cpus(bus, sysfs_root)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except OSError:
        return None


def core_slices(cpus, local_world: int, sysfs_root: str = "/sys"):
    """Deal the hardware threads in `cpus` to `local_world` ranks as contiguous runs of whole cores (both SMT siblings
    of a core go to the same rank; cores ordered by package, then core id): [set, ...] of length local_world.  Without
    topology files every CPU counts as its own core, in numeric order."""
    import os
    cores = {}
    for c in sorted(cpus):
        top = os.path.join(sysfs_root, "devices", "system", "cpu", f"cpu{c}", "topology")
        try:
            key = (int(open(os.path.join(top, "physical_package_id")).read()), int(open(os.path.join(top, "core_id")).read()))
        except (OSError, ValueError):
            key = (0, c)
        cores.setdefault(key, []).append(c)
    order = [cores[k] for k in sorted(cores)]
    n = len(order)
    out = []
    for r in range(local_world):
        lo, hi = r * n // local_world, (r + 1) * n // local_world
        out.append({c for core in order[lo:hi] for c in core})
    return out


def _device_bus(device_index: int):
    try:
        p = torch.cuda.get_device_properties(device_index)
        return f"{getattr(p, 'pci_domain_id'):04x}:{getattr(p, 'pci_bus_id'):02x}:{getattr(p, 'pci_device_id'):02x}.0"
    except Exception:  # noqa: BLE001 — older torch without the pci_* properties, or no device
        return None


def pin_rank(local_rank: int, local_world: int, sysfs_root: str = "/sys", node_cpus=None):
    """CPU placement of one rank of a one-process-per-GPU launch (device index = local rank), before its first proxy
    (and with it the rANS pool) is created: the GPU's NUMA node when the platform names it, and inside that set — or
    inside the whole allowed set when it does not — this rank's own run of whole cores, dealt among the local ranks that
    share the set, so that the ranks' spinning rANS workers neither share a core nor migrate.  `node_cpus(d)` -> CPU set
    of local device d or None (default: sysfs).  Opt-in (DCVC_B200_PIN=1): measured on a 2-GPU B200 box (two sockets, one
    GPU per NUMA node, round 2 call 24) the Intra headline of two ranks was 497.7 / 497.8 FPS unpinned and 507.0 / 473.8 /
    468.1 pinned, against 2 x 254.8 for one rank — the run-to-run spread is larger than the effect, so the default leaves
    the affinity alone.  Returns the CPU set applied or None."""
    import os
    if os.environ.get("DCVC_B200_PIN", "0") != "1":
        return None
    if node_cpus is None:
        def node_cpus(d):
            bus = _device_bus(d)
            return gpu_numa_cpus(bus, sysfs_root) if bus else None
    try:
        allowed = set(os.sched_getaffinity(0))
        sets = [frozenset((node_cpus(d) or allowed) & allowed) for d in range(max(1, local_world))]
        mine_node = sets[local_rank % len(sets)]
        if not mine_node:
            return None
        peers = [d for d in range(len(sets)) if sets[d] == mine_node]
        mine = core_slices(mine_node, len(peers), sysfs_root)[peers.index(local_rank % len(sets))]
        if len(mine) < 2 or mine == allowed:
            return None
        os.sched_setaffinity(0, mine)
        return mine
    except OSError:
        return None
