"""Multi-GPU check of the exchange step behind the C ABI (run under torchrun on N >= 2 GPUs of one
box; not a pytest file: the round-end GPU test box has one GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank owns a slab of blocks; after w3d_eps_sig_ifv (collective: NCCL neighbour exchange or
all-reduce inside the engine) the internal forces are compared with the one-domain CPU restatement
at 1e-10: on the dofs the rank's blocks touch (neighbour mode), on all dofs (full mode), as the
concatenation of the owned ranges, and the reduced scalars.  Also with u, du given on rank 0 only
(ncclBroadcast).  Prints one line per rank and exits non-zero on a mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from oracle.oracle_py import Oracle
    from warp3d_b200 import build_structured_hex8, new_state, smooth_displacement
    from warp3d_b200.distributed import DomainEngine
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = build_structured_hex8(12, 8, 4 * world, coord_jitter=0.1, seed=7, block_size=32)
    u, du = np.zeros(m.nodof), smooth_displacement(m, strain=3.0e-3)
    orc, so = Oracle(m, nthreads=4), new_state(m)
    orc.eps_sig_ifv(so, 1, 0, u, du)
    f_ref, sc_ref = orc.eps_sig_ifv(so, 1, 1, u, du)
    scale = np.abs(f_ref).max()
    # one-domain residual and norms from the CPU restatement (upres + cvtest residual loop)
    cst = np.zeros(m.nodof, dtype=np.int32)
    xs = m.c.reshape(-1, 3)[:, 0]
    cst.reshape(-1, 3)[xs <= xs.min() + 1e-9] = 1
    rng = np.random.default_rng(3)
    pbar = scale * 0.3 * rng.normal(size=m.nodof); load = scale * 0.1 * rng.normal(size=m.nodof)
    from oracle import oracle_py as op
    res_ref, nrm_ref = op.upres(cst, pbar, load, f_ref, du, 1.0, 0.25)
    nrm_ref.update(op.cvtest_residual(res_ref))
    ok = True
    for exchange, bcast in (("neighbours", False), ("full", False), ("neighbours", True)):
        de = DomainEngine(m, rank, world, device=local, exchange=exchange, bcast_inputs=bcast)
        touched = np.zeros(m.nodof, dtype=bool)
        off = m.block_offsets(24)
        for b in de.engine.owned_blocks:
            touched[m.edest[off[b]:off[b + 1]].astype(np.int64) - 1] = True
        uu, dd = (u, du) if (rank == 0 or not bcast) else (None, None)
        de.drive_eps_sig_internal_forces(1, 0, uu, dd)
        f, sc = de.drive_eps_sig_internal_forces(1, 1, uu, dd)
        sel = touched if exchange == "neighbours" else np.ones(m.nodof, dtype=bool)
        e1 = float(np.abs(f[sel] - f_ref[sel]).max() / scale)
        own = de.owned_mask()
        part = torch.from_numpy(np.where(own, f, 0.0)).cuda()
        dist.all_reduce(part)
        e2 = float(np.abs(part.cpu().numpy() - f_ref).max() / scale)
        first, n = de.owned_range()
        fo, _ = de.drive_eps_sig_internal_forces(1, 1, uu, dd, owned_only=True)
        e3 = float(np.abs(fo - f_ref[first - 1:first - 1 + n])[own[first - 1:first - 1 + n]].max() / scale)
        es = max(abs(sc[k] - sc_ref[k]) / abs(sc_ref[k]) for k in ("internal_energy", "plastic_work", "sum_ifv"))
        # structure residual and its norms (w3d_residual is collective with several domains): the
        # norms equal the one-domain ones on every rank, the ranks' res vectors add up to the
        # structure residual (neighbour mode: each rank evaluates the dofs it owns)
        de.engine.residual_set_loads(cst, pbar=pbar, load=load)
        res, nrm = de.engine.residual(1, 1.0, 0.25)
        rsum = torch.from_numpy(res.copy()).cuda()
        if exchange == "neighbours":
            dist.all_reduce(rsum)
        er = float(np.abs(rsum.cpu().numpy() - res_ref).max() / np.abs(res_ref).max())
        en = max(abs(nrm[k] - nrm_ref[k]) / abs(nrm_ref[k]) for k in ("norm_res", "mgload", "sum_loads", "res_max"))
        good_res = (er < 1e-10 and en < 1e-12 and nrm["num_term_loads"] == nrm_ref["num_term_loads"]
                    and nrm["res_max_dof"] == nrm_ref["res_max_dof"])
        good = (max(e1, e2, e3) < 1e-10 and es < 1e-10 and sc["num_term_ifv"] == sc_ref["num_term_ifv"]
                and not sc["material_cut_step"] and good_res)
        ok = ok and good
        print(f"rank {rank}/{world} exchange={exchange} bcast={bcast}: touched {e1:.1e} owned-concat {e2:.1e} "
              f"owned-range {e3:.1e} scalars {es:.1e} num_term {sc['num_term_ifv']} residual {er:.1e} norms {en:.1e} "
              f"max dof {nrm['res_max_dof']} -> {'OK' if good else 'MISMATCH'}",
              flush=True)
        de.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
