#!/usr/bin/env python
"""Replay the block records an ifort build of the UNMODIFIED reference wrote through
fortran/w3d_dump_hook.f (stream file w3d_dump.bin) and report how far the CPU restatement -- and,
on a GPU box, the CUDA engine -- are from the Fortran's own numbers at block level.  This is the
function-level pin (north_star: 1e-10) that cannot be produced in the development image (no
Fortran compiler); the tool and its record format are exercised there by tests/test_dump_replay.py
with records written from the restatement itself.

usage: compare_fortran_dump.py w3d_dump.bin --e 200e9 --nu 0.3 --yld 400e6 --tan-e 2e9 [--beta 1]
                               [--engine] [--eps-bbar 0] [--no-qbar]
       compare_fortran_dump.py w3d_dump.bin --mises --e 200e9 --nu 0.3 --yld 400e6 --n-power 10
                               [--m-power 0 --ref-eps 0] [--dt 1e6] ...      (blocks of model 3)

Record (little endian): 13 x int32 header (magic 87033001, blk, step, iter, span, ngp, hist_size,
totdof, mxvl, mat_type, geonl, bbar, felem) then float64 arrays in Fortran order:
ce_0(mxvl,totdof) ue due | elem_hist(span,hist,ngp) urcs_blk_n(mxvl,9,ngp) strain_n(mxvl,6,ngp) |
ddtse(mxvl,6,ngp) urcs_blk_n1(mxvl,9,ngp) elem_hist1(span,hist,ngp) rot_blk_n1(mxvl,9,ngp).
Every record is replayed on a stand-alone model of `span` disconnected elements (an element's
results depend only on its own nodes)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MAGIC = 87033001


def read_records(path):
    with open(path, "rb") as f:
        while True:
            hdr = np.fromfile(f, dtype="<i4", count=13)
            if hdr.size < 13:
                return
            if hdr[0] != MAGIC:
                raise ValueError("bad record magic")
            _, blk, step, it, span, ngp, hist, totdof, mxvl, mat, geonl, bbar, felem = (int(x) for x in hdr)

            def arr(*shape):          # Fortran-order array -> numpy with reversed axes
                n = int(np.prod(shape))
                a = np.fromfile(f, dtype="<f8", count=n)
                if a.size != n:
                    raise ValueError("truncated record")
                return a.reshape(shape[::-1])

            rec = dict(blk=blk, step=step, iter=it, span=span, ngp=ngp, hist=hist, totdof=totdof, mat=mat,
                       geonl=bool(geonl), bbar=bool(bbar), felem=felem)
            for name in ("ce_0", "ue", "due"):
                rec[name] = arr(mxvl, totdof)[:, :span].T                      # (span, totdof)
            rec["hist_n"] = arr(span, hist, ngp).transpose(2, 0, 1)             # (span, ngp, hist)
            rec["urcs_n"] = arr(mxvl, 9, ngp)[:, :, :span].transpose(2, 0, 1)   # (span, ngp, 9)
            rec["eps_n"] = arr(mxvl, 6, ngp)[:, :, :span].transpose(2, 0, 1)
            rec["ddtse"] = arr(mxvl, 6, ngp)[:, :, :span].transpose(2, 0, 1)
            rec["urcs_n1"] = arr(mxvl, 9, ngp)[:, :, :span].transpose(2, 0, 1)
            rec["hist_n1"] = arr(span, hist, ngp).transpose(2, 0, 1)
            rec["rot_n1"] = arr(mxvl, 9, ngp)[:, :, :span].transpose(2, 0, 1)
            yield rec


def write_record(f, rec, mxvl=128):
    """the hook's record from Python (tests): the same bytes the Fortran would write."""
    span, ngp, hist, totdof = rec["span"], rec["ngp"], rec["hist"], rec["totdof"]
    np.array([MAGIC, rec["blk"], rec["step"], rec["iter"], span, ngp, hist, totdof, mxvl, rec["mat"],
              int(rec["geonl"]), int(rec["bbar"]), rec["felem"]], dtype="<i4").tofile(f)

    def elem_table(a):                  # (span, totdof) -> Fortran (mxvl, totdof)
        out = np.zeros((totdof, mxvl))
        out[:, :span] = a.T
        out.astype("<f8").tofile(f)

    def hist_block(a):                  # (span, ngp, hist) -> Fortran (span, hist, ngp)
        np.ascontiguousarray(a.transpose(1, 2, 0)).astype("<f8").tofile(f)

    def gp_block(a):                    # (span, ngp, k) -> Fortran (mxvl, k, ngp)
        out = np.zeros((ngp, a.shape[2], mxvl))
        out[:, :, :span] = a.transpose(1, 2, 0)
        out.astype("<f8").tofile(f)

    for name in ("ce_0", "ue", "due"):
        elem_table(rec[name])
    hist_block(rec["hist_n"]); gp_block(rec["urcs_n"]); gp_block(rec["eps_n"])
    gp_block(rec["ddtse"]); gp_block(rec["urcs_n1"]); hist_block(rec["hist_n1"]); gp_block(rec["rot_n1"])


def replay(rec, material, eps_bbar=0.0, qbar=True, use_engine=False, dt=1.0):
    """Returns dict name -> (max abs difference, scale) for eps_n1, urcs_n1, hist_n1, rot_n1."""
    from oracle.oracle_py import Oracle
    from warp3d_b200 import build_hex8_model, new_state
    span = rec["span"]
    ce = rec["ce_0"]                                            # (span, 24): x(8), y(8), z(8)
    coords = np.stack([ce[:, 0:8], ce[:, 8:16], ce[:, 16:24]], axis=2).reshape(-1, 3)
    incid = (np.arange(span * 8).reshape(span, 8) + 1).astype(np.int32)
    m = build_hex8_model(coords, incid, materials=[material], geonl=rec["geonl"], bbar=rec["bbar"],
                         block_size=128, eps_bbar=eps_bbar, qbar_flag=qbar, dt=dt)

    def nodal(a):                                               # (span, 24) -> (nodof,)
        return np.stack([a[:, 0:8], a[:, 8:16], a[:, 16:24]], axis=2).reshape(-1)

    u, du = nodal(rec["ue"]), nodal(rec["due"])
    st = new_state(m)
    st.history[:] = rec["hist_n"].reshape(-1)
    st.urcs_n[:] = rec["urcs_n"].reshape(-1)
    st.eps_n[:] = rec["eps_n"].reshape(-1)
    if use_engine:
        from warp3d_b200.engine import Engine
        be = Engine(m)
        be.push_state(st, names=("history", "urcs_n", "eps_n"))
        be.drive_eps_sig_internal_forces(rec["step"], rec["iter"], u, du)
        out = be.pull_state(new_state(m))
        be.close()
    else:
        Oracle(m).eps_sig_ifv(st, rec["step"], rec["iter"], u, du)
        out = st
    res = {}
    for name, got in (("eps_n1", out.eps_n1), ("urcs_n1", out.urcs_n1), ("hist_n1", out.history1),
                      ("rot_n1", out.rot_n1)):
        want = (rec["eps_n"] + rec["ddtse"]) if name == "eps_n1" else rec[name]
        if name == "hist_n1" and rec["iter"] == 0 and rec["mat"] != 10:
            continue                                            # history is not saved at iteration 0
        if name == "rot_n1" and (rec["iter"] == 0 or not rec["geonl"]):
            continue
        w = want.reshape(-1)
        res[name] = (float(np.abs(got - w).max()), float(np.abs(w).max()))
    return res


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("dump")
    ap.add_argument("--e", type=float, required=True); ap.add_argument("--nu", type=float, required=True)
    ap.add_argument("--yld", type=float, required=True); ap.add_argument("--tan-e", type=float, required=True)
    ap.add_argument("--beta", type=float, default=1.0); ap.add_argument("--eps-bbar", type=float, default=0.0)
    ap.add_argument("--no-qbar", action="store_true"); ap.add_argument("--engine", action="store_true")
    ap.add_argument("--mises", action="store_true", help="the blocks use the `mises` card (material model 3)")
    ap.add_argument("--n-power", type=float, default=0.0); ap.add_argument("--m-power", type=float, default=0.0)
    ap.add_argument("--ref-eps", type=float, default=0.0); ap.add_argument("--dt", type=float, default=1.0)
    a = ap.parse_args()
    from warp3d_b200 import Material
    if a.mises:
        mat = Material(mat_type=3, e=a.e, nu=a.nu, yld_pt=a.yld, tan_e=a.tan_e, beta=1.0,
                       extra=dict(n_power=a.n_power, m_power=a.m_power, ref_eps=a.ref_eps))
    else:
        mat = Material(mat_type=1, e=a.e, nu=a.nu, yld_pt=a.yld, tan_e=a.tan_e, beta=a.beta)
    worst = {}
    n = 0
    for rec in read_records(a.dump):
        if rec["mat"] != mat.mat_type:
            print(f"block {rec['blk']}: material model {rec['mat']} -- give the model's properties through the API")
            continue
        r = replay(rec, mat, a.eps_bbar, not a.no_qbar, a.engine, dt=a.dt)
        n += 1
        for k, (d, s) in r.items():
            rel = d / s if s > 0 else d
            worst[k] = max(worst.get(k, 0.0), rel)
        print(f"block {rec['blk']} step {rec['step']} iter {rec['iter']}: " +
              "  ".join(f"{k} {d / s if s > 0 else d:.2e}" for k, (d, s) in r.items()))
    print(f"{n} records; worst relative difference: " + "  ".join(f"{k} {v:.2e}" for k, v in worst.items()))
    sys.exit(0 if all(v <= 1e-10 for v in worst.values()) else 1)


if __name__ == "__main__":
    main()
