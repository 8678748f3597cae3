// layout.hpp -- the reference's packed per-problem sequence layouts for C++ callers of the C ABI.
//
// Every host array of include/ddpb.h is, inside one problem's block, laid out as the reference lays out its sequences:
// one flat array per sequence, knot t at offset sum_{t' < t} stride(t'), stride = rows * cols, matrices column-major,
// third-order tensors T(k, i, j) at k + i O + j O L.  The reference expresses this with compile-time indexer / iterator /
// proxy templates; this header restates the same arithmetic at run time, under the reference's names, so that code written
// against them (and the reference's own layout tests, tests/cpp/layout_kats.cpp) carries over:
//   regular_indexer_t, vec_/mat_regular_indexer   /root/reference/include/ddp/indexer.hpp:249-289, 485-500
//   row_concat                                    indexer.hpp:152-195
//   outer_prod                                    indexer.hpp:197-247, 450-478
//   shift_time_idx                                indexer.hpp:291-318
//   range_row_filter / periodic_row_filter        indexer.hpp:328-393 / 395-448
//   iteration (current_index, rows, cols, offset) indexer.hpp:44-150
//   mat_seq_t                                     detail/mat_seq.hpp:10-128 (NaN-initialised, :36)
//   tensor_view_t                                 detail/tensor.hpp:141-147
// Host code only; depends on the standard library.
#pragma once
#include <cassert>
#include <functional>
#include <limits>
#include <utility>
#include <vector>

namespace ddpb {

using index_t = long;

namespace indexing {

// rows(t) x cols(t) for t in [index_begin, index_end)
class indexer_t {
 public:
  using dim_fn = std::function<index_t(index_t)>;
  indexer_t(index_t begin, index_t end, dim_fn rows, dim_fn cols, index_t max_rows, index_t max_cols)
      : m_begin(begin), m_end(end), m_rows(std::move(rows)), m_cols(std::move(cols)), m_max_rows(max_rows), m_max_cols(max_cols) {
    assert(begin <= end);
    for (index_t t = begin; t < end; ++t) m_required += stride(t);
  }
  index_t index_begin() const { return m_begin; }
  index_t index_end() const { return m_end; }
  index_t rows(index_t t) const { return m_rows(t); }
  index_t cols(index_t t) const { return m_cols(t); }
  index_t max_rows() const { return m_max_rows; }
  index_t max_cols() const { return m_max_cols; }
  index_t stride(index_t t) const { return rows(t) * cols(t); }
  index_t required_memory() const { return m_required; }  // doubles of a sequence over this indexer
  indexer_t clone() const { return *this; }

  // *it of the reference: the knot's index, dimensions and position in the flat storage
  struct proxy_t {
    const indexer_t* idx;
    index_t t, offset;
    index_t current_index() const { return t; }
    index_t rows() const { return idx->rows(t); }
    index_t cols() const { return idx->cols(t); }
    index_t max_rows() const { return idx->max_rows(); }
    index_t max_cols() const { return idx->max_cols(); }
    index_t memory_offset() const { return offset; }
  };
  class iterator {
   public:
    iterator(const indexer_t* idx, index_t t, index_t offset) : p{idx, t, offset} {}
    proxy_t operator*() const { return p; }
    iterator& operator++() { p.offset += p.idx->stride(p.t); ++p.t; return *this; }
    iterator& operator--() { --p.t; p.offset -= p.idx->stride(p.t); return *this; }
    iterator operator++(int) { iterator c = *this; ++*this; return c; }
    iterator operator--(int) { iterator c = *this; --*this; return c; }
    iterator& operator+=(index_t n) { for (; n > 0; --n) ++*this; for (; n < 0; ++n) --*this; return *this; }
    iterator& operator-=(index_t n) { return *this += -n; }
    friend iterator operator+(iterator a, index_t n) { return a += n; }
    friend iterator operator-(iterator a, index_t n) { return a -= n; }
    friend bool operator==(const iterator& a, const iterator& b) { return a.p.t == b.p.t; }
    friend bool operator!=(const iterator& a, const iterator& b) { return a.p.t != b.p.t; }
    friend bool operator<(const iterator& a, const iterator& b) { return a.p.t < b.p.t; }
    friend bool operator>(const iterator& a, const iterator& b) { return a.p.t > b.p.t; }
    friend bool operator<=(const iterator& a, const iterator& b) { return a.p.t <= b.p.t; }
    friend bool operator>=(const iterator& a, const iterator& b) { return a.p.t >= b.p.t; }

   private:
    proxy_t p;
  };
  iterator begin() const { return iterator(this, m_begin, 0); }
  iterator end() const { return iterator(this, m_end, m_required); }

 private:
  index_t m_begin, m_end;
  dim_fn m_rows, m_cols;
  index_t m_max_rows, m_max_cols;
  index_t m_required = 0;
};
inline indexer_t::iterator begin(const indexer_t& i) { return i.begin(); }
inline indexer_t::iterator end(const indexer_t& i) { return i.end(); }

inline indexer_t mat_regular_indexer(index_t begin, index_t end, index_t rows, index_t cols) {
  return indexer_t(begin, end, [rows](index_t) { return rows; }, [cols](index_t) { return cols; }, rows, cols);
}
inline indexer_t vec_regular_indexer(index_t begin, index_t end, index_t rows) { return mat_regular_indexer(begin, end, rows, 1); }
// rows of both at every knot (vectors over the same range)
inline indexer_t row_concat(const indexer_t& a, const indexer_t& b) {
  assert(a.index_begin() == b.index_begin() && a.index_end() == b.index_end());
  return indexer_t(a.index_begin(), a.index_end(), [a, b](index_t t) { return a.rows(t) + b.rows(t); }, [a](index_t t) { return a.cols(t); },
                   a.max_rows() + b.max_rows(), a.max_cols());
}
// matrices rows(l) x rows(r) from two vector indexers
inline indexer_t outer_prod(const indexer_t& l, const indexer_t& r) {
  assert(l.index_begin() == r.index_begin() && l.index_end() == r.index_end());
  return indexer_t(l.index_begin(), l.index_end(), [l](index_t t) { return l.rows(t); }, [r](index_t t) { return r.rows(t); }, l.max_rows(),
                   r.max_rows());
}
// knot t has what `idx` has at t + dt; the range moves to [begin - dt, end - dt)
inline indexer_t shift_time_idx(const indexer_t& idx, index_t dt) {
  return indexer_t(idx.index_begin() - dt, idx.index_end() - dt, [idx, dt](index_t t) { return idx.rows(t + dt); },
                   [idx, dt](index_t t) { return idx.cols(t + dt); }, idx.max_rows(), idx.max_cols());
}
// rows only for t in [begin, end)
inline indexer_t range_row_filter(const indexer_t& idx, index_t begin, index_t end) {
  return indexer_t(idx.index_begin(), idx.index_end(), [idx, begin, end](index_t t) { return begin <= t && t < end ? idx.rows(t) : 0; },
                   [idx](index_t t) { return idx.cols(t); }, idx.max_rows(), idx.max_cols());
}
// rows only where (t - index_begin) % period == first_offset
inline indexer_t periodic_row_filter(const indexer_t& idx, index_t period, index_t first_offset) {
  assert(period > 0 && first_offset < period);
  const index_t b = idx.index_begin();
  return indexer_t(b, idx.index_end(), [idx, b, period, first_offset](index_t t) { return (t - b) % period == first_offset ? idx.rows(t) : 0; },
                   [idx](index_t t) { return idx.cols(t); }, idx.max_rows(), idx.max_cols());
}

}  // namespace indexing

// column-major view of one knot's matrix
template <typename T> struct mat_view_t {
  T* ptr;
  index_t m_rows, m_cols;
  index_t rows() const { return m_rows; }
  index_t cols() const { return m_cols; }
  T& operator()(index_t r, index_t c) const { assert(r >= 0 && r < m_rows && c >= 0 && c < m_cols); return ptr[r + c * m_rows]; }
  T* data() const { return ptr; }
};
// T(k, i, j), k < O (output), i < L (left), j < R (right), at k + i O + j O L
template <typename T> struct tensor_view_t {
  T* ptr;
  index_t O, L, R;
  T& operator()(index_t k, index_t i, index_t j) const { assert(k < O && i < L && j < R); return ptr[k + i * O + j * O * L]; }
};

// one flat array, the knots' matrices back to back
template <typename T> class mat_seq_t {
 public:
  explicit mat_seq_t(indexing::indexer_t idx)
      : m_idx(std::move(idx)), m_data(static_cast<std::size_t>(m_idx.required_memory()), std::numeric_limits<T>::quiet_NaN()) {
    index_t o = 0;
    for (index_t t = m_idx.index_begin(); t < m_idx.index_end(); ++t) { m_off.push_back(o); o += m_idx.stride(t); }
    m_off.push_back(o);
  }
  // over EXISTING storage laid out the same way (one problem's block of a C-ABI host array): no copy
  mat_seq_t(indexing::indexer_t idx, T* storage) : mat_seq_t(std::move(idx)) { m_ext = storage; m_data.clear(); }
  T* data() { return m_ext ? m_ext : m_data.data(); }
  const indexing::indexer_t& indexer() const { return m_idx; }
  index_t offset(index_t t) const { return m_off[static_cast<std::size_t>(t - m_idx.index_begin())]; }
  mat_view_t<T> operator[](index_t t) { return mat_view_t<T>{data() + offset(t), m_idx.rows(t), m_idx.cols(t)}; }

  class iterator {
   public:
    iterator(mat_seq_t* s, index_t t) : s(s), t(t) {}
    // **it of the reference is the matrix; here *it is
    mat_view_t<T> operator*() const { return (*s)[t]; }
    iterator& operator++() { ++t; return *this; }
    iterator& operator--() { --t; return *this; }
    friend bool operator==(const iterator& a, const iterator& b) { return a.t == b.t; }
    friend bool operator!=(const iterator& a, const iterator& b) { return a.t != b.t; }

   private:
    mat_seq_t* s;
    index_t t;
  };
  iterator begin() { return iterator(this, m_idx.index_begin()); }
  iterator end() { return iterator(this, m_idx.index_end()); }

 private:
  indexing::indexer_t m_idx;
  std::vector<T> m_data;
  std::vector<index_t> m_off;
  T* m_ext = nullptr;
};

// ne_per_knot of ddpb_batch_create from the BASE indexer of a constraint that is advanced twice in time
// (constraint_advance_time_t o constraint_advance_time_t, problem.hpp:553: knot t carries the rows the base has at t + 2)
inline std::vector<int> eq_rows_per_knot(index_t horizon, const indexing::indexer_t& base) {
  const indexing::indexer_t s = indexing::shift_time_idx(indexing::shift_time_idx(base, 1), 1);
  std::vector<int> ne;
  for (index_t t = 0; t < horizon; ++t) ne.push_back(int(s.rows(t)));
  return ne;
}

}  // namespace ddpb
