// The structural known-answer tests of the reference's own unit tests, on ddpb/layout.hpp:
//   /root/reference/test/indexing.cpp:4-40   regular indexer: iterator algebra, current_index / rows / cols
//   /root/reference/test/indexing.cpp:42-86  composed indexers: per-knot rows / cols, max_rows / max_cols, strides 0,6,0,6,0,15
//   /root/reference/test/mat_seq.cpp:9-52    mat_seq packing: ragged matrices back to back, column-major, offsets 0 and 12
// plus the time shift of the doubly advanced constraint (test/pendulum_ddp.cpp:37-47) and the tensor index order
// (detail/tensor.hpp:146).  Host code only.
#include <cmath>
#include <cstdio>

#include "ddpb/ddpb.hpp"    // (both headers together: they share ddpb::index_t)
#include "ddpb/layout.hpp"

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

int main() {
  using namespace ddpb;
  {  // regular indexer over [3, 12) of 3 x 4 matrices: iterator algebra and what a dereferenced iterator reports
    const auto idx = indexing::mat_regular_indexer(3, 12, 3, 4);
    CHECK(idx.index_begin() == 3 && idx.index_end() == 12 && idx.required_memory() == 9 * 12);
    auto it = begin(idx);
    ++it; --it; it++; it--;              // back at the first knot
    CHECK((*it).current_index() == 3 && (*it).memory_offset() == 0);
    it += 4; it -= 2;                    // third knot
    const auto here = *it;
    CHECK(here.current_index() == 5 && here.rows() == 3 && here.cols() == 4 && here.memory_offset() == 2 * 12);
    // post-increment / post-decrement return the old position, the prefix forms the new one
    const auto before_inc = it++;
    CHECK(before_inc == it - 1);
    const auto before_dec = it--;
    CHECK(before_dec == it + 1);
    CHECK(++it == before_dec);
    CHECK(--it == before_inc);
    // ordering
    const auto next = it + 1, prev = it - 1;
    CHECK(it == it && it != next && next > it && prev < it && next >= it && prev <= it && it >= it && it <= it);
    // walking the whole range visits every knot once and ends at end()
    index_t count = 0;
    for (auto w = begin(idx); w != end(idx); ++w) ++count;
    CHECK(count == 9);
  }
  {  // composed indexers
    auto idx = indexing::vec_regular_indexer(0, 12, 3);
    auto idx2 = indexing::vec_regular_indexer(0, 12, 2);
    auto idx3 = indexing::vec_regular_indexer(0, 12, 3);
    auto filtered = indexing::periodic_row_filter(idx, 3, 2);
    auto filtered2 = indexing::periodic_row_filter(idx3, 2, 1);
    auto concatenated = indexing::row_concat(filtered, idx2);
    auto prod = indexing::outer_prod(concatenated, filtered2);
    auto it = begin(prod);
    CHECK((*it).max_rows() == 5);
    CHECK((*it).max_cols() == 3);
    const index_t want[6][2] = {{2, 0}, {2, 3}, {5, 0}, {2, 3}, {2, 0}, {5, 3}};
    for (int k = 0; k < 6; ++k, ++it) { CHECK((*it).rows() == want[k][0]); CHECK((*it).cols() == want[k][1]); }
    const index_t strides[6] = {0, 6, 0, 6, 0, 15};
    for (int k = 0; k < 6; ++k) CHECK(prod.stride(k) == strides[k]);
  }
  {  // matrix sequence
    auto reg = indexing::vec_regular_indexer(0, 12, 3);
    auto reg2 = indexing::vec_regular_indexer(0, 12, 2);
    auto reg3 = indexing::vec_regular_indexer(0, 12, 3);
    auto idx = indexing::outer_prod(indexing::row_concat(indexing::periodic_row_filter(reg, 3, 2), reg2), indexing::periodic_row_filter(reg3, 2, 1));
    mat_seq_t<double> seq(idx);
    double* ptr = seq.data();
    CHECK(std::isnan(ptr[0]));  // detail/mat_seq.hpp:36
    auto it = seq.begin();
    CHECK((*it).rows() == 2);
    CHECK((*it).cols() == 0);
    ++it;
    CHECK((*it).rows() == 2);
    CHECK((*it).cols() == 3);
    (*it)(0, 0) = 1.0;
    (*it)(1, 0) = 2.0;
    (*it)(0, 1) = 3.0;
    ++it, ++it, ++it, ++it;
    CHECK((*it).rows() == 5);
    CHECK((*it).cols() == 3);
    (*it)(0, 0) = 4.0;
    (*it)(1, 0) = 5.0;
    (*it)(0, 1) = 6.0;
    CHECK(ptr[0 + 2 * 0 + 0] == 1.0);
    CHECK(ptr[0 + 2 * 0 + 1] == 2.0);
    CHECK(ptr[0 + 2 * 1 + 0] == 3.0);
    CHECK(ptr[12 + 5 * 0 + 0] == 4.0);
    CHECK(ptr[12 + 5 * 0 + 1] == 5.0);
    CHECK(ptr[12 + 5 * 1 + 0] == 6.0);
    // a view over existing storage (one problem's block of a C-ABI host array) addresses the same doubles
    mat_seq_t<double> view(idx, ptr);
    CHECK(view[5](0, 1) == 6.0 && view.offset(5) == 12);
  }
  {  // the reference's terminal constraint: base indexer on [2, T+2) filtered to [T, T+1) -> one row at knot T-2
    const index_t T = 200;
    auto base = indexing::range_row_filter(indexing::vec_regular_indexer(2, T + 2, 1), T, T + 1);
    auto ne = eq_rows_per_knot(T, base);
    int sum = 0;
    for (int r : ne) sum += r;
    CHECK(sum == 1 && ne[T - 2] == 1);
    auto every = eq_rows_per_knot(10, indexing::vec_regular_indexer(2, 12, 6));  // test/pinocchio_ddp.cpp:37
    for (int r : every) CHECK(r == 6);
  }
  {  // tensor index order
    double buf[4 * 3 * 5];
    for (int i = 0; i < 60; ++i) buf[i] = i;
    tensor_view_t<double> t{buf, 4, 3, 5};
    CHECK(t(1, 2, 3) == 1 + 2 * 4 + 3 * 12);
  }
  std::printf("%d failures\n", failures);
  return failures != 0;
}
