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
    """host placement helper (opt-in in bench.py: DCVC_B200_NUMA_PIN=1): the CPUs of the GPU's NUMA node come from
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
