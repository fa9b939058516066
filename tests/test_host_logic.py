"""Host-side integer tables must be bit-exact with the reference's rules."""
import numpy as np

from warp3d_b200 import build_structured_hex8, slab_domains
from warp3d_b200.model import HEX8_NODE_SIGNS, simple_blocking


def _blocking_sequential(noelem, keys, auto_size):
    """literal restatement of inelbk_simple_blocking (src/inelbk.f:282-345, 497-548)"""
    blocks, felem, size = [], 1, 1
    for element in range(2, noelem + 1):
        size += 1
        newblk = size > auto_size or not np.array_equal(keys[element - 1], keys[felem - 1])
        if not newblk:
            continue
        blocks.append((size - 1, felem))
        felem, size = element, 1
    blocks.append((size, felem))
    return blocks


def test_auto_blocking_matches_the_sequential_rule():
    rng = np.random.default_rng(3)
    for n, bs in ((1, 128), (127, 128), (128, 128), (129, 128), (1000, 64), (333, 7)):
        keys = np.cumsum(rng.random(n) < 0.02).reshape(-1, 1) % 3
        assert simple_blocking(n, keys, bs) == _blocking_sequential(n, keys, bs)


def test_block_table_and_dest_maps():
    m = build_structured_hex8(5, 4, 3, block_size=16)
    assert m.noelem == 60 and m.nonode == 6 * 5 * 4 and m.nodof == 3 * m.nonode
    assert list(m.elblks[:, 0]) == [16, 16, 16, 12] and list(m.elblks[:, 1]) == [1, 17, 33, 49]
    ed = m.edest.reshape(-1, 24)
    # all u, all v, all w ordering: edest(j)=dstmap(node_j), edest(8+j)=+1, edest(16+j)=+2
    assert np.array_equal(ed[:, :8], 3 * (m.incid - 1) + 1)
    assert np.array_equal(ed[:, 8:16], ed[:, :8] + 1) and np.array_equal(ed[:, 16:], ed[:, :8] + 2)
    # node ordering gives a positive Jacobian for the unit cube with WARP3D's parent coordinates
    x = m.c.reshape(-1, 3)[m.incid[0] - 1]
    assert np.allclose((x - x.mean(0)) * 2, HEX8_NODE_SIGNS)


def test_props_table_is_the_reference_word_layout():
    m = build_structured_hex8(2, 2, 2)
    p = m.props
    assert p.dtype == np.int32 and p.shape[1] == 43
    assert p[0, 0] == 2 and p[0, 1] == 8 and p[0, 3] == 3 and p[0, 5] == 8 and p[0, 24] == 1
    assert p[0, 6:8].view(np.float32)[0] == np.float32(200.0e9)      # E is real*4
    e, et = np.float32(200.0e9), np.float32(2.0e9)
    assert p[0, 14:15].view(np.float32)[0] == np.float32((e * et) / (e - et))   # H' in real*4 arithmetic
    assert p[0, 17] & 1 and p[0, 18] & 1


def test_slab_domains_are_contiguous_and_balanced():
    m = build_structured_hex8(20, 10, 10)           # 2000 elements, 16 blocks
    for nproc in (2, 4, 8):
        d = slab_domains(m, nproc)
        assert np.all(np.diff(d) >= 0) and d[0] == 0 and d[-1] == nproc - 1
        counts = np.bincount(d, weights=m.elblks[:, 0], minlength=nproc)
        assert counts.max() - counts.min() <= 2 * 128


def test_every_bench_workload_builds_and_states_its_algorithmic_bytes():
    """bench.py's workloads on a small box: the tables build, the load increment has the model's
    size, and the description carries what the roofline needs."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    for name, (dims, kind, cpu_dims) in bench.WORKLOADS.items():
        m, du, nprep = bench.make_workload(name, (6, 4, 4))
        assert m.noelem == (17318 if kind == "seb" else 96), name     # the deck has its own mesh
        assert du.shape == (m.nodof,) and nprep in (1, 2), name
        wd = bench.workload_description(name)
        assert wd["bytes_stress"] >= 776.0 and wd["bytes_tanstf"] >= 708.0 and wd["material"], name
        if kind != "seb":
            assert np.prod(dims) >= np.prod(cpu_dims), name      # the CPU sample is a sub-box
    beta = bench.make_workload("hex8_2m_mm01_beta05", (2, 2, 2))[0].matprp[0, 2]
    assert beta == np.float32(0.5)
    e, p = (abs(bench.make_workload(w, (2, 2, 2))[1]).max() for w in ("hex8_2m_mm01_elastic", "hex8_2m_mm01"))
    assert e < 0.5 * p                                           # the elastic variant stays under yield


def test_mises_material_words_follow_elprp():
    """`properties mises e .. nu .. yld_pt .. n_power ..` (inmat.f:816-823, defaults :307-327) ->
    element property words of elprp.f:866-897: 15 = h (0 for tan_e 0), 20 m_power, 21 n_power,
    22 = 1 / ref_eps (or 0), 23 yield stress, 25 model 3, 26 f0 = 0, 27-29 q1-q3 = 0, 30 no
    nucleation; history size 15 (mm03_set_sizes, mm03.f:1623-1640)."""
    import numpy as np
    from warp3d_b200 import Material, build_structured_hex8
    mat = Material(mat_type=3, e=200.0e9, nu=0.3, yld_pt=400.0e6, tan_e=0.0,
                   extra=dict(n_power=10.0, m_power=12.0, ref_eps=0.5))
    m = build_structured_hex8(3, 2, 2, materials=[mat], block_size=5)
    f32 = m.props.view(np.float32)
    assert (m.hist_size == 15).all()
    assert (m.props[:, 24] == 3).all()
    assert (f32[:, 14] == 0.0).all() and (f32[:, 22] == np.float32(400.0e6)).all()
    assert (f32[:, 19] == np.float32(12.0)).all() and (f32[:, 20] == np.float32(10.0)).all()
    assert (f32[:, 21] == np.float32(2.0)).all()                 # eta = 1 / D
    assert (f32[:, 25:29] == 0.0).all() and (m.props[:, 29] == 0).all()
    # a bilinear material leaves those words untouched (zero)
    m1 = build_structured_hex8(3, 2, 2, block_size=5)
    assert (m1.props[:, 19:22] == 0).all() and (m1.hist_size == 11).all()
