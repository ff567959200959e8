"""CPU tests of the N>1 path: particles / poses are sharded contiguously, no data-path collective, results gathered.
world_size = 2 over gloo; each rank computes its shard with the CPU emulation of the product code (the GPU run uses the same
shard_range / gather_records helpers in bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from rmcl_b200.shard import shard_range
    for n in (0, 1, 7, 100, 100001):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import pyemul
    from oracle import pyoracle as po
    from rmcl_b200 import synth
    from rmcl_b200.shard import gather_records, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V, F = synth.cube(12)
        sc = pyemul.Scene(V, F)
        Tsb = synth.scenario_tsb()
        m = synth.c1_sensor()
        o, d = po.model_rays(m)
        pts = po.Scene(V, F).simulate(synth.make_transform(), Tsb, o, d, 80.0)["points"]
        beams = synth.pf_beams(pts, 16)
        P, A = synth.pf_particles(101, footprint=(16.0, 16.0), z=0.0, margin=0.0)
        P["t"][:, :2] -= 8.0
        b, e = shard_range(len(P), rank, world)
        prm = po.PFParams.defaults()
        local = sc.pf_update(P[b:e], A[b:e], Tsb, beams, prm)
        full = gather_records(local, dist, dst=0)
        # the one exchange step of the cycle: {sum, max} of likelihood.mean, all-reduced over the shards
        import torch
        ls, lm = po.pf_likelihood_stats(local)
        ts, tm = torch.tensor([ls], dtype=torch.float64), torch.tensor([lm], dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        # resampling: the stage WITH an exchange (opponents come from all particles): all-gather, then resample the own champions.
        # The product passes its CUDA kernel as `resample_fn`; here the CPU emulation of the same device code stands in.
        from rmcl_b200.shard import gladiator_resample_sharded
        n_eq = 100                                                                  # equal shard sizes
        Pq, Aq = P[:n_eq].copy(), sc.pf_update(P[:n_eq], A[:n_eq], Tsb, beams, prm)
        cfg = po.GladiatorConfig.defaults()
        lb, le = shard_range(n_eq, rank, world)

        def resample_fn(P_all, A_all, first, n_local):
            Pa = P_all.numpy().view(Pq.dtype).reshape(-1)
            Aa = A_all.numpy().view(Aq.dtype).reshape(-1)
            Pn, An, _, _ = pyemul.gladiator(Pa, Aa, first, n_local, cfg, 1234, 5)
            return Pn, An

        Pl = torch.from_numpy(Pq[lb:le].view(np.float32).reshape(-1, 8).copy())
        Al = torch.from_numpy(Aq[lb:le].view(np.float32).reshape(-1, 9).copy())
        Pn, An = gladiator_resample_sharded(resample_fn, Pl, Al, dist)
        Pfull, Afull = gather_records(Pn, dist, dst=0), gather_records(An, dist, dst=0)
        if rank == 0:
            ref = sc.pf_update(P, A, Tsb, beams, prm)
            rs, rm = po.pf_likelihood_stats(ref)
            stats_ok = float(tm[0]) == rm and abs(float(ts[0]) - rs) <= 1e-5 * abs(rs) + 1e-6
            Pr, Ar, _, _ = pyemul.gladiator(Pq, Aq, 0, n_eq, cfg, 1234, 5)
            res_ok = Pfull.tobytes() == Pr.tobytes() and Afull.tobytes() == Ar.tobytes() and not np.array_equal(Ar["likelihood"]["mean"], Aq["likelihood"]["mean"])
            q.put(("ok", full.tobytes() == ref.tobytes() and stats_ok and res_ok, len(full)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_particle_sharding_invariance_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    tag, same, n = q.get(timeout=5)
    assert tag == "ok" and same and n == 101
