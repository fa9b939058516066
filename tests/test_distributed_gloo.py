"""world_size-2 test of the multi-GPU host logic on CPU (gloo): blocks are partitioned into
contiguous slabs, each rank computes the internal forces of its own blocks (here with the oracle
restricted to the rank's blocks, standing in for the rank's GPU engine) and the interface-force
sum is one all-reduce -- the result must equal the single-rank vector."""
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
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    m = build_structured_hex8(6, 4, 4, coord_jitter=0.1, seed=2, block_size=16)
    dom = slab_domains(m, world)
    u = np.zeros(m.nodof); du = smooth_displacement(m, strain=3.0e-3)
    # full model on every rank for the check
    ofull, sfull = Oracle(m), new_state(m)
    f_full, sc_full = ofull.eps_sig_ifv(sfull, 1, 1, u, du)
    # this rank's share: a sub-model holding only the owned blocks (same global dof numbering)
    own = np.nonzero(dom == rank)[0]
    sub = build_structured_hex8(6, 4, 4, coord_jitter=0.1, seed=2, block_size=16)
    off = sub.block_offsets(24)
    sub.edest = np.concatenate([sub.edest[off[b]:off[b + 1]] for b in own])
    sub.cdest = np.concatenate([sub.cdest[off[b]:off[b + 1]] for b in own])
    sub.elblks = sub.elblks[own].copy(); sub.hist_size = sub.hist_size[own].copy(); sub.nelblk = len(own)
    o, s = Oracle(sub), new_state(sub)
    f_loc, sc = o.eps_sig_ifv(s, 1, 1, u, du)
    t = torch.from_numpy(f_loc.copy())
    dist.all_reduce(t)                      # the path's only exchange step
    sc_t = torch.tensor([sc["internal_energy"], sc["plastic_work"], sc["sum_ifv"], float(sc["num_term_ifv"])],
                        dtype=torch.float64)
    dist.all_reduce(sc_t)
    err = float(np.abs(t.numpy() - f_full).max() / np.abs(f_full).max())
    serr = abs(sc_t[1].item() - sc_full["plastic_work"]) / abs(sc_full["plastic_work"])
    # interface dofs = dofs touched by both ranks' blocks
    q.put((rank, err, serr, int(sc_t[3].item()), sc_full["num_term_ifv"], int(len(own))))
    dist.destroy_process_group()


def test_two_rank_partition_and_ifv_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, serr, nterm, nterm_full, nown in res:
        assert err < 1e-13, (rank, err)
        assert serr < 1e-12
        assert nterm == nterm_full and nown > 0
