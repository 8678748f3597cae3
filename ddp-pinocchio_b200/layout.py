"""Packed per-problem sequence layouts at the host boundary (the reference's indexers and mat_seq_t).

Restates, as plain index arithmetic, what the reference expresses with iterator/proxy templates:
  regular_indexer_t        /root/reference/include/ddp/indexer.hpp:249-289
  outer_product_indexer_t  indexer.hpp:197-247        row_concat_indexer_t  indexer.hpp:152-195
  shift_time_idx_t         indexer.hpp:291-318        range_row_filter_t    indexer.hpp:328-393
  periodic_row_filter_t    indexer.hpp:395-448
  mat_seq_t                detail/mat_seq.hpp:10-128 (one flat array, knot t at offset sum_{t'<t} stride(t'),
                           matrices column-major, NaN-initialised)
  tensor index             detail/tensor.hpp:141-147 (k + i*O + j*O*L)
Every host array of the C ABI (include/ddpb.h) uses these offsets inside a problem's block."""
import numpy as np


class Indexer:
    """rows(t) / cols(t) for t in [begin, end); stride(t) = rows * cols."""

    def __init__(self, begin, end, rows_fn, cols_fn=lambda t: 1, max_rows=None, max_cols=1):
        self.begin, self.end, self._r, self._c = begin, end, rows_fn, cols_fn
        self._max_rows, self._max_cols = max_rows, max_cols

    def rows(self, t):
        return self._r(t)

    def cols(self, t):
        return self._c(t)

    def stride(self, t):
        return self.rows(t) * self.cols(t)

    def max_rows(self):
        return self._max_rows if self._max_rows is not None else max(self.rows(t) for t in range(self.begin, self.end))

    def max_cols(self):
        return self._max_cols

    def offsets(self):
        """offset of every knot + total size (mat_seq_t storage size)."""
        s = np.array([self.stride(t) for t in range(self.begin, self.end)], dtype=np.int64)
        return np.concatenate([[0], np.cumsum(s)])

    def __len__(self):
        return self.end - self.begin


def vec_regular_indexer(begin, end, dim):
    return Indexer(begin, end, lambda t: dim, max_rows=dim)


def periodic_row_filter(idx, period, first_offset):
    """rows only where (t - begin) % period == first_offset, else 0 (indexer.hpp:395-448)."""
    return Indexer(idx.begin, idx.end, lambda t: idx.rows(t) if (t - idx.begin) % period == first_offset else 0, idx.cols,
                   max_rows=idx.max_rows(), max_cols=idx.max_cols())


def range_row_filter(idx, begin, end):
    """rows only for t in [begin, end) (indexer.hpp:328-393); test/pendulum_ddp.cpp:37-47 uses (T, T+1)."""
    return Indexer(idx.begin, idx.end, lambda t: idx.rows(t) if begin <= t < end else 0, idx.cols, max_rows=idx.max_rows(),
                   max_cols=idx.max_cols())


def row_concat(a, b):
    assert (a.begin, a.end) == (b.begin, b.end)
    return Indexer(a.begin, a.end, lambda t: a.rows(t) + b.rows(t), a.cols, max_rows=a.max_rows() + b.max_rows(), max_cols=a.max_cols())


def outer_prod(l, r):
    assert (l.begin, l.end) == (r.begin, r.end)
    return Indexer(l.begin, l.end, l.rows, r.rows, max_rows=l.max_rows(), max_cols=r.max_rows())


def shift_time(idx, dt):
    """indexer over [begin - dt, end - dt) whose knot t has the rows of knot t + dt (indexer.hpp:291-318)."""
    return Indexer(idx.begin - dt, idx.end - dt, lambda t: idx.rows(t + dt), lambda t: idx.cols(t + dt), max_rows=idx.max_rows(),
                   max_cols=idx.max_cols())


class MatSeq:
    """mat_seq_t: flat NaN-initialised storage + per-knot column-major views."""

    def __init__(self, idx):
        self.idx = idx
        self.off = idx.offsets()
        self.data = np.full(int(self.off[-1]), np.nan)

    def __getitem__(self, t):
        i = t - self.idx.begin
        r, c = self.idx.rows(t), self.idx.cols(t)
        return self.data[self.off[i]:self.off[i + 1]].reshape((c, r)).T  # column-major view


def tensor_index(k, i, j, O, L):
    """tensor_view_t::operator() (detail/tensor.hpp:141-147)."""
    return k + i * O + j * O * L


def eq_rows_per_knot(T, base_indexer):
    """Rows per knot of a constraint advanced twice in time (constraint_advance_time_t applied twice,
    problem.hpp:553): knot t carries the rows the base indexer has at t + 2.  This is `ne_per_knot` of
    ddpb_batch_create."""
    s = shift_time(shift_time(base_indexer, 1), 1)
    return np.array([s.rows(t) for t in range(0, T)], dtype=np.int32)
