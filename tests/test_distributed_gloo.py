"""CPU tests of the N>1 path (world_size 2, gloo): shard geometry + the record exchange of
control_toolkit_b200.distributed.ShardPlan, with numpy twins of the per-shard record / combine arithmetic of the
kernels (K1 record, K2 combine; CEM candidate merge).  The sharded result must equal the unsharded oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mppi_record(S, z, neg_inv_lbd):
    """numpy twin of the K1 block/shard softmin record: [rho, a, b_z[n_ind]] (ctk_kernels_mppi.cuh)."""
    rho = S.min()
    e = np.exp((S - rho) * neg_inv_lbd)
    return np.concatenate([[rho, e.sum()], (e[:, None] * z).sum(0)]).astype(np.float32)


def _mppi_combine(records, neg_inv_lbd):
    """numpy twin of K2 mppi_combine_kernel: exact rescaling of per-shard records to the global minimum."""
    rho = records[:, 0].min()
    sc = np.exp((records[:, 0] - rho) * neg_inv_lbd)
    a = (sc * records[:, 1]).sum()
    bz = (sc[:, None] * records[:, 2:]).sum(0)
    return rho, a, bz


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from control_toolkit_b200.distributed import ShardPlan
    from helpers import load_golden, make_oracle, replay
    from oracle.mppi import interpolation_matrix
    plan = ShardPlan.from_env()
    assert (plan.rank, plan.world_size) == (rank, world)

    # ---- MPPI: every rank evaluates ONLY its shard of the global population (global noise rows off..off+cnt) ----
    z, meta = load_golden("mppi_c1_n2000")
    o = make_oracle(meta)
    o.step(z["states"][0], replay(meta))  # full-population oracle (reference result)
    S_all, du_all = o.last["J"], o.last["delta_u"][..., 0]
    N, n_ind = o.N, o.n_ind
    zz = replay(meta)
    zz.as_torch = False
    z_all = zz.standard_draws("normal", (N, n_ind, 1))[..., 0]
    off, cnt = plan.local_offset(N), plan.local_count(N)
    nil = np.float32(-1.0 / o.LBD)
    rec = _mppi_record(S_all[off:off + cnt], z_all[off:off + cnt], nil)
    gathered = plan.all_gather(torch.from_numpy(rec)).numpy().reshape(world, -1)
    rho, a, bz = _mppi_combine(gathered, nil)
    W = interpolation_matrix(o.H, o.period)
    b = (bz @ W) * float(o.SQRTRHODTINV) / a
    u_nom_prev = np.zeros(o.H, np.float32)  # reset state, shifted
    u_nom = np.clip(u_nom_prev + b, -1, 1)
    err = np.abs(u_nom - o.u_nom.numpy()[0, :, 0]).max() / np.abs(o.u_nom.numpy()).max()

    # ---- the two forms of the fused exchange (ctk_kernels_mppi.cuh mppi_tick_finish): every shard's grid emits one record per BLOCK;
    #      two hops = blocks -> shard record -> world's shard records, one hop = all world x G block records at once.  Both must
    #      give the unsharded result (the rescaling to the common minimum is exact up to rounding, in any grouping) ----
    G = 5  # blocks of this shard's grid, ragged shares
    cuts = np.linspace(0, cnt, G + 1).astype(int)
    cuts[1:-1] += np.array([3, -2, 1, 0])[: G - 1]
    blocks = np.stack([_mppi_record(S_all[off + a:off + b], z_all[off + a:off + b], nil) for a, b in zip(cuts[:-1], cuts[1:])])
    rho_s, a_s, bz_s = _mppi_combine(blocks, nil)  # first hop, on this shard
    shard_rec = np.concatenate([[rho_s, a_s], bz_s]).astype(np.float32)
    two = _mppi_combine(plan.all_gather(torch.from_numpy(shard_rec)).numpy().reshape(world, -1), nil)
    one = _mppi_combine(plan.all_gather(torch.from_numpy(blocks.ravel().copy())).numpy().reshape(world * G, -1), nil)
    ref = _mppi_record(S_all, z_all, nil)
    for got in (two, one):
        assert got[0] == ref[0]  # the global minimum is exact
        assert abs(got[1] - ref[1]) <= 2e-6 * abs(ref[1])
        assert np.abs(got[2] - ref[2:]).max() <= 2e-6 * max(np.abs(ref[2:]).max(), 1e-6) + 2e-6 * abs(ref[1])

    # ---- CEM: per-shard top-k candidates (cost, GLOBAL id) -> all-gather -> merge == global stable argsort[:k] ----
    z2, meta2 = load_golden("cem_c2_n4096_k64")
    J = z2["J_0"]
    k = meta2["cfg"]["cem_best_k"]
    off2, cnt2 = plan.local_offset(J.size), plan.local_count(J.size)
    loc = np.argsort(J[off2:off2 + cnt2], kind="stable")[:k]
    cand = np.stack([J[off2:off2 + cnt2][loc], (loc + off2).astype(np.float32)], 1).astype(np.float32).ravel()
    allc = plan.all_gather(torch.from_numpy(cand)).numpy().reshape(-1, 2)
    order = np.lexsort((allc[:, 1], allc[:, 0]))[:k]
    merged = allc[order, 1].astype(np.int64)
    ok_cem = np.array_equal(merged, np.argsort(J, kind="stable")[:k])
    q.put((rank, float(err), bool(ok_cem), off, cnt))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_exchange_matches_unsharded_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert [r[3] for r in res] == [0, 1000] and [r[4] for r in res] == [1000, 1000]
    for rank, err, ok_cem, _, _ in res:
        assert err < 1e-6, (rank, err)  # sharded softmin combine == unsharded, up to summation-order rounding
        assert ok_cem
