"""world_size-2 and -3 tests of the multi-GPU host logic on CPU (gloo).  What is exercised is the
PRODUCT's exchange plan, through the C ABI (w3d_comm_plan / w3d_comm_owned in
libw3d_engine.so -- host-only routines, the same ones the engine uses at create): blocks are
partitioned into contiguous slabs, every rank computes the internal forces of its own blocks (the
CPU restatement restricted to the rank's blocks stands in for the rank's GPU kernels, which need a
device), packs its partial sums on the dofs the plan says it shares with each neighbour, exchanges
them pairwise (gloo send/recv in the role of ncclSend/ncclRecv) and adds what it receives --
exactly the sequence w3d_comm.cuh runs on the device.  Checks: both sides of a pair hold the same
dof list; after the exchange every rank holds the complete force on every dof its blocks touch;
the owned dofs of all ranks tile 1..nodof once, so the concatenation is the full vector; the
5-double scalar reduction equals the one-domain scalars."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle.oracle_py import Oracle
    from warp3d_b200 import build_structured_hex8, new_state, slab_domains, smooth_displacement
    from warp3d_b200.engine import comm_owned, comm_plan
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    m = build_structured_hex8(6, 4, 5, coord_jitter=0.1, seed=2, block_size=16)
    m.elblks[:, 2] = slab_domains(m, world)
    u = np.zeros(m.nodof); du = smooth_displacement(m, strain=3.0e-3)
    # the one-domain answer (every rank computes it for the check)
    ofull, sfull = Oracle(m), new_state(m)
    f_full, sc_full = ofull.eps_sig_ifv(sfull, 1, 1, u, du)
    # this rank's share: a sub-model holding only the owned blocks (same global dof numbering)
    own = np.nonzero(m.elblks[:, 2] == rank)[0]
    sub = build_structured_hex8(6, 4, 5, coord_jitter=0.1, seed=2, block_size=16)
    off = sub.block_offsets(24)
    sub.edest = np.concatenate([sub.edest[off[b]:off[b + 1]] for b in own])
    sub.cdest = np.concatenate([sub.cdest[off[b]:off[b + 1]] for b in own])
    sub.elblks = sub.elblks[own].copy(); sub.hist_size = sub.hist_size[own].copy(); sub.nelblk = len(own)
    o, s = Oracle(sub), new_state(sub)
    f_loc, sc = o.eps_sig_ifv(s, 1, 1, u, du)
    touched = np.zeros(m.nodof, dtype=bool)
    touched[sub.edest.astype(np.int64) - 1] = True

    # ---- the product's plan (C ABI, host only) and the exchange it prescribes
    plan = comm_plan(m, rank, world)                    # per rank: 1-based ascending dofs
    assert len(plan[rank]) == 0
    f = f_loc.copy()
    reqs, recvs = [], {}
    for r in range(world):
        if r == rank or len(plan[r]) == 0:
            continue
        send = torch.from_numpy(f_loc[plan[r] - 1].copy())          # k_iface_pack
        recvs[r] = torch.zeros(len(plan[r]), dtype=torch.float64)
        reqs.append(dist.isend(send, dst=r)); reqs.append(dist.irecv(recvs[r], src=r))
    for rq in reqs:
        rq.wait()
    for r, buf in recvs.items():
        np.add.at(f, plan[r] - 1, buf.numpy())                       # k_iface_add
    err_touched = float(np.abs(f[touched] - f_full[touched]).max() / np.abs(f_full).max())
    # both sides of every pair hold the same list
    lists = [None] * world
    dist.all_gather_object(lists, [p.tolist() for p in plan])
    sym = all(lists[a][b] == lists[b][a] for a in range(world) for b in range(world))
    # owned dofs tile the vector; concatenation of the owned parts = the full vector
    owned = comm_owned(m, rank, world)
    part = torch.from_numpy(np.where(owned, f, 0.0))
    cover = torch.from_numpy(owned.astype(np.float64))
    dist.all_reduce(part); dist.all_reduce(cover)
    err_full = float(np.abs(part.numpy() - f_full).max() / np.abs(f_full).max())
    # scalars + cut flag: one 5-double reduction
    red = torch.tensor([sc["internal_energy"], sc["plastic_work"], sc["sum_ifv"], float(sc["num_term_ifv"]),
                        1.0 if sc["material_cut_step"] else 0.0], dtype=torch.float64)
    dist.all_reduce(red)
    serr = max(abs(red[k].item() - sc_full[n]) / abs(sc_full[n])
               for k, n in enumerate(("internal_energy", "plastic_work", "sum_ifv")))
    q.put(dict(rank=rank, err_touched=err_touched, err_full=err_full, sym=sym,
               cover_ok=bool((cover.numpy() == 1.0).all()), serr=serr, nterm=int(red[3].item()),
               nterm_full=sc_full["num_term_ifv"], cut=red[4].item(), nown=int(len(own)),
               nshared=[len(p) for p in plan]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_exchange_plan_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r["err_touched"] < 1e-13, r
        assert r["err_full"] < 1e-13, r
        assert r["sym"] and r["cover_ok"], r
        assert r["serr"] < 1e-12 and r["nterm"] == r["nterm_full"] and r["cut"] == 0.0 and r["nown"] > 0, r
    if world == 3:      # the middle slab has two neighbours, the outer ones one each
        by_rank = {r["rank"]: r["nshared"] for r in res}
        assert by_rank[0][2] == 0 and by_rank[2][0] == 0 and by_rank[1][0] > 0 and by_rank[1][2] > 0
