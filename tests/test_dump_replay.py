"""tools/compare_fortran_dump.py and the record format of fortran/w3d_dump_hook.f: block records
written from the restatement's own state (standing in for the ifort build's dump, which cannot be
produced here) are read back and replayed on stand-alone models of disconnected elements; the
replay must return the recorded results.  This is what makes the function-level pin a one-command
job once a Fortran binary exists."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import compare_fortran_dump as cfd  # noqa: E402

from common import load_path  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402
from warp3d_b200 import Material, build_structured_hex8, new_state  # noqa: E402


def _records(m, mat, path, hist=11, mat_type=1):
    """run the load path on the restatement and cut one record per (block, iteration)"""
    o, st = Oracle(m), new_state(m)
    off24 = m.block_offsets(24)
    recs = []
    for step, iters in path:
        for itr, u, du in iters:
            before = st.copy()
            o.eps_sig_ifv(st, step, itr, u, du)
            for blk in range(m.nelblk):
                span = m.span(blk)
                ed = m.edest[off24[blk]:off24[blk + 1]].reshape(span, 24).astype(np.int64) - 1

                def cut(a, k):
                    off = m.block_offsets(8 * k)
                    return a[off[blk]:off[blk + 1]].reshape(span, 8, k).copy()

                recs.append(dict(blk=blk + 1, step=step, iter=itr, span=span, ngp=8, hist=hist, totdof=24, mat=mat_type,
                                 geonl=True, bbar=True, felem=m.felem(blk), ce_0=m.c[ed], ue=u[ed], due=du[ed],
                                 hist_n=cut(before.history, hist), urcs_n=cut(before.urcs_n, 9),
                                 eps_n=cut(before.eps_n, 6), ddtse=cut(st.eps_n1, 6) - cut(before.eps_n, 6),
                                 urcs_n1=cut(st.urcs_n1, 9), hist_n1=cut(st.history1, hist), rot_n1=cut(st.rot_n1, 9)))
        o.update(st); o.adaptive_save(st)
    return recs


def _write(tmp_path, recs):
    p = str(tmp_path / "w3d_dump.bin")
    with open(p, "wb") as f:
        for r in recs:
            cfd.write_record(f, r)
    return p


def test_dump_records_round_trip_and_replay(tmp_path):
    mat = Material(e=200.0e9, nu=0.3, yld_pt=400.0e6, tan_e=2.0e9, beta=1.0)
    m = build_structured_hex8(5, 3, 3, materials=[mat], coord_jitter=0.1, seed=8, block_size=20)
    recs = _records(m, mat, load_path(m, nsteps=2, strain=2.5e-3))
    back = list(cfd.read_records(_write(tmp_path, recs)))
    assert len(back) == len(recs) == 2 * 3 * m.nelblk
    for a, b in zip(recs, back):
        for k in ("blk", "step", "iter", "span", "felem"):
            assert a[k] == b[k]
        for k in ("ce_0", "ue", "due", "hist_n", "urcs_n", "eps_n", "ddtse", "urcs_n1", "hist_n1", "rot_n1"):
            assert np.array_equal(np.asarray(a[k], dtype=float), b[k]), k
    nplastic = 0
    for rec in back:
        res = cfd.replay(rec, mat)
        for name, (d, s) in res.items():
            assert d <= 1e-12 * max(s, 1e-300), (rec["blk"], rec["step"], rec["iter"], name, d, s)
        nplastic += int((rec["hist_n1"][:, :, 3].view(np.int64) & 0xFFFFFFFF == 1).sum())
    assert nplastic > 0


def test_dump_replay_of_mises_blocks(tmp_path):
    """the same for blocks of material model 3 (`mises ... n_power`: what the seb and test_1 decks ship)"""
    mat = Material(mat_type=3, e=200.0e9, nu=0.3, yld_pt=400.0e6, tan_e=0.0, extra={"n_power": 10.0})
    m = build_structured_hex8(4, 3, 2, materials=[mat], coord_jitter=0.1, seed=9, block_size=7, dt=1.0e6)
    recs = _records(m, mat, load_path(m, nsteps=2, strain=2.5e-3), hist=15, mat_type=3)
    nplastic = 0
    for rec in cfd.read_records(_write(tmp_path, recs)):
        assert rec["mat"] == 3 and rec["hist"] == 15
        for name, (d, s) in cfd.replay(rec, mat, dt=1.0e6).items():
            assert d <= 1e-12 * max(s, 1e-300), (rec["blk"], rec["step"], rec["iter"], name, d, s)
        nplastic += int((rec["hist_n1"][:, :, 5].view(np.int64) & 0xFFFFFFFF == 1).sum())
    assert nplastic > 0


@pytest.mark.gpu
def test_dump_replay_through_the_engine(tmp_path):
    mat = Material(e=200.0e9, nu=0.3, yld_pt=400.0e6, tan_e=2.0e9, beta=1.0)
    m = build_structured_hex8(5, 3, 3, materials=[mat], coord_jitter=0.1, seed=8, block_size=20)
    recs = _records(m, mat, load_path(m, nsteps=2, strain=2.5e-3))
    for rec in cfd.read_records(_write(tmp_path, recs)):
        for name, (d, s) in cfd.replay(rec, mat, use_engine=True).items():
            assert d <= 1e-10 * max(s, 1e-300), (rec["blk"], rec["step"], rec["iter"], name, d, s)
