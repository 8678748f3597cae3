"""Structural KATs the reference's own tests pin for this path:
test/mat_seq.cpp:45-51 (packed offsets), test/indexing.cpp:4-86 (composed indexers, strides 0,6,0,6,0,15),
plus the time shifting of the doubly advanced constraint (Appendix A)."""
import numpy as np


def test_regular_indexer(pkg):
    L = pkg.layout if hasattr(pkg, "layout") else __import__("ddp_pinocchio_b200.layout", fromlist=["x"])
    idx = L.vec_regular_indexer(0, 12, 3)
    assert len(idx) == 12 and all(idx.rows(t) == 3 and idx.cols(t) == 1 for t in range(12))
    assert idx.offsets()[-1] == 36


def test_compose_indexers_strides(pkg):
    # test/indexing.cpp:42-86
    from ddp_pinocchio_b200 import layout as L
    idx, idx2, idx3 = L.vec_regular_indexer(0, 12, 3), L.vec_regular_indexer(0, 12, 2), L.vec_regular_indexer(0, 12, 3)
    prod = L.outer_prod(L.row_concat(L.periodic_row_filter(idx, 3, 2), idx2), L.periodic_row_filter(idx3, 2, 1))
    assert prod.max_rows() == 5 and prod.max_cols() == 3
    assert [(prod.rows(t), prod.cols(t)) for t in range(6)] == [(2, 0), (2, 3), (5, 0), (2, 3), (2, 0), (5, 3)]
    assert [prod.stride(t) for t in range(6)] == [0, 6, 0, 6, 0, 15]


def test_mat_seq_packing(pkg):
    # test/mat_seq.cpp:9-52: ragged matrices stored back to back, column-major
    from ddp_pinocchio_b200 import layout as L
    reg, reg2, reg3 = L.vec_regular_indexer(0, 12, 3), L.vec_regular_indexer(0, 12, 2), L.vec_regular_indexer(0, 12, 3)
    idx = L.outer_prod(L.row_concat(L.periodic_row_filter(reg, 3, 2), reg2), L.periodic_row_filter(reg3, 2, 1))
    seq = L.MatSeq(idx)
    assert np.isnan(seq.data).all()  # mat_seq.hpp:36
    assert seq[0].shape == (2, 0) and seq[1].shape == (2, 3)
    m = seq[1]
    m[0, 0], m[1, 0], m[0, 1] = 1.0, 2.0, 3.0
    assert seq[5].shape == (5, 3)
    m = seq[5]
    m[0, 0], m[1, 0], m[0, 1] = 4.0, 5.0, 6.0
    p = seq.data
    assert (p[0], p[1], p[2]) == (1.0, 2.0, 3.0)
    assert (p[12 + 5 * 0 + 0], p[12 + 5 * 0 + 1], p[12 + 5 * 1 + 0]) == (4.0, 5.0, 6.0)


def test_advanced_constraint_rows(pkg):
    # base indexer on [2, T+2) filtered to [T, T+1) (test/pendulum_ddp.cpp:37-47) -> rows only at knot T-2
    from ddp_pinocchio_b200 import layout as L, synthetic as S
    T = 200
    base = L.range_row_filter(L.vec_regular_indexer(2, T + 2, 1), T, T + 1)
    ne = L.eq_rows_per_knot(T, base)
    assert ne.sum() == 1 and ne[T - 2] == 1
    assert np.array_equal(ne, S.workload("c1").ne_per_knot())
    # every-knot variant of test/pinocchio_ddp.cpp:37
    assert (L.eq_rows_per_knot(10, L.vec_regular_indexer(2, 12, 6)) == 6).all()


def test_tensor_index(pkg):
    from ddp_pinocchio_b200 import layout as L
    O, Ld, R = 4, 3, 5
    seen = {L.tensor_index(k, i, j, O, Ld) for k in range(O) for i in range(Ld) for j in range(R)}
    assert seen == set(range(O * Ld * R)) and L.tensor_index(1, 2, 3, O, Ld) == 1 + 2 * 4 + 3 * 12


def test_host_packed_offsets_match_batch_sizes(pkg):
    """The per-problem packed sizes the C ABI expects equal the reference's mat_seq sizes."""
    from ddp_pinocchio_b200 import layout as L, synthetic as S, capi
    w = S.workload("c4", B=2, T=9)
    bb = capi.BatchBase(); bb._init_dims(w.model, w.B, w.T, w.ne_per_knot(), w.eq_kind)
    nd, nu = w.ndx, w.nu
    eq = L.shift_time(L.shift_time(L.range_row_filter(L.vec_regular_indexer(2, w.T + 2, 3), w.T, w.T + 1), 1), 1)
    x = L.vec_regular_indexer(0, w.T, nd); u = L.vec_regular_indexer(0, w.T, nu)
    sizes = bb.deriv_sizes()
    assert sizes["eq_x"] == L.outer_prod(eq, x).offsets()[-1] and sizes["fu"] == L.outer_prod(x, u).offsets()[-1]
    assert sizes["eq_val"] == eq.offsets()[-1] == 3
