"""world_size-2 gloo test of the multi-GPU host logic (no GPU): weight broadcast + job sharding + gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dcvc_b200.shard import broadcast_state_dict, gather_results, shard_jobs
    from dcvc_b200.spec import dmci_spec, synth_state_dict
    spec = {k: v for k, v in dmci_spec().items() if k.startswith("hyper_enc.") or k.startswith("q_scale")}
    sd = synth_state_dict(spec, 0) if rank == 0 else None
    got = broadcast_state_dict(sd, spec, 0)
    ref = synth_state_dict(spec, 0)
    ok = all(torch.equal(got[k], ref[k]) for k in spec)
    jobs = [(s, q_) for s in range(5) for q_ in range(4)]       # 5 sequences x 4 rate points (runtime_avg.json)
    mine = shard_jobs(jobs, rank, world)
    merged = gather_results([(j, rank) for j in mine], 0)
    if rank == 0:
        q.put((ok, sorted(j for j, _ in merged) == sorted(jobs), len(mine)))
    else:
        q.put((ok, True, len(mine)))
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] for r in res), "broadcast checkpoint differs"
    assert all(r[1] for r in res), "job shards do not cover the job list exactly once"
    assert sorted(r[2] for r in res) == [10, 10]


def test_gpu_numa_cpu_set_from_a_sysfs_tree(tmp_path):
    """host placement helper (bench.py: pin_rank): the CPUs of the GPU's NUMA node come from
    sysfs; a platform that does not say (numa_node -1, missing files) changes nothing"""
    from dcvc_b200.shard import _parse_cpulist, gpu_numa_cpus
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    dev = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
    dev.mkdir(parents=True)
    node = tmp_path / "devices" / "system" / "node" / "node1"
    node.mkdir(parents=True)
    (node / "cpulist").write_text("32-63,96-127\n")
    (dev / "numa_node").write_text("1\n")
    cpus = gpu_numa_cpus("0000:1B:00.0", str(tmp_path))
    assert cpus == set(range(32, 64)) | set(range(96, 128))
    (dev / "numa_node").write_text("-1\n")
    assert gpu_numa_cpus("0000:1b:00.0", str(tmp_path)) is None
    assert gpu_numa_cpus("0000:ff:00.0", str(tmp_path)) is None


def test_ranks_get_disjoint_runs_of_whole_cores(tmp_path, monkeypatch):
    """pin_rank: every local rank gets its own run of whole cores (SMT siblings stay together) inside the CPU set of its
    GPU's NUMA node — or inside the whole allowed set when the platform names no node; ranks that share a node split it"""
    import os
    from dcvc_b200 import shard
    # a two-package box: cpu c and c + 8 are siblings of core c % 4 of package c // 4 % 2
    for c in range(16):
        top = tmp_path / "devices" / "system" / "cpu" / f"cpu{c}" / "topology"
        top.mkdir(parents=True)
        (top / "physical_package_id").write_text(f"{(c % 8) // 4}\n")
        (top / "core_id").write_text(f"{c % 4}\n")
    sl = shard.core_slices(set(range(16)), 4, str(tmp_path))
    assert sl == [{0, 8, 1, 9}, {2, 10, 3, 11}, {4, 12, 5, 13}, {6, 14, 7, 15}]
    assert shard.core_slices({0, 1, 2, 3, 4}, 2, str(tmp_path / "nothing")) == [{0, 1}, {2, 3, 4}]

    applied = {}
    monkeypatch.setenv("DCVC_B200_PIN", "1")       # opt-in (measured: within the run-to-run spread on a 2-GPU box)
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(16)))
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: applied.__setitem__("cpus", set(cpus)))
    node = {0: {0, 1, 2, 3, 8, 9, 10, 11}, 1: {0, 1, 2, 3, 8, 9, 10, 11}, 2: {4, 5, 6, 7, 12, 13, 14, 15}, 3: {4, 5, 6, 7, 12, 13, 14, 15}}
    got = [shard.pin_rank(r, 4, str(tmp_path), node_cpus=lambda d: node[d]) for r in range(4)]
    assert got == [{0, 8, 1, 9}, {2, 10, 3, 11}, {4, 12, 5, 13}, {6, 14, 7, 15}] and applied["cpus"] == got[3]
    # no NUMA information: the local ranks split everything that is allowed
    got = [shard.pin_rank(r, 2, str(tmp_path), node_cpus=lambda d: None) for r in range(2)]
    assert got == [{0, 1, 2, 3, 8, 9, 10, 11}, {4, 5, 6, 7, 12, 13, 14, 15}]
    # one rank and no NUMA information, or the switch: nothing changes
    assert shard.pin_rank(0, 1, str(tmp_path), node_cpus=lambda d: None) is None
    monkeypatch.delenv("DCVC_B200_PIN")
    assert shard.pin_rank(0, 2, str(tmp_path), node_cpus=lambda d: None) is None


def test_job_deal_is_a_partition_and_mixes_the_rate_points():
    """shard_jobs: every job exactly once, equal shares, and with the reference's (sequence, rate) job order every rank
    sees every rate point when there are at least as many ranks as rate points (8 ranks x 32 jobs: one of each)"""
    from dcvc_b200.shard import shard_jobs
    jobs = [(s, r) for s in range(8) for r in range(4)]
    for world in (1, 2, 3, 4, 8):
        shares = [shard_jobs(jobs, k, world) for k in range(world)]
        assert sorted(j for sh in shares for j in sh) == jobs
        assert max(len(sh) for sh in shares) - min(len(sh) for sh in shares) <= 1
        if world in (4, 8):
            for sh in shares:
                assert sorted({r for _, r in sh}) == [0, 1, 2, 3], (world, sh)
        if world == 2:   # two rate points per rank would be {0, 2} / {1, 3} with a plain modulo deal
            assert [sum(r for _, r in sh) for sh in shares] == [24, 24]
