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
