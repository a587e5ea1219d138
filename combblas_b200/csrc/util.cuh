#pragma once
#include "common.cuh"

namespace cbgpu {

struct BinResult {
  int64_t count[256];  // tasks per bucket
  int64_t weight[256]; // summed weight per bucket (when a weight array is given)
  int64_t weight2[256];
  int64_t offset[256]; // start of the bucket's tasks in `order` (buckets are laid out in DESCENDING id)
  int64_t listed;      // number of tasks listed (bucket 0 is skipped)
};

// Groups the tasks for launching. `class_of_bucket[b]` (0..14, 15 = not launched) names the kernel class of every
// bucket; order[] receives the task ids sorted by (class, row window, larger bucket first), so that one launch walks the
// row windows one after the other (the window-major slice of A stays in L2) with its biggest tasks first.
// class_begin/class_count describe the slice of order[] of every class. Synchronises.
struct ClassRanges {
  int64_t begin[16];
  int64_t count[16];
};
int bin_tasks(cbgpu_ctx_impl *ctx, const uint8_t *bucket, const int64_t *weight, const int64_t *weight2, int64_t n,
              const uint32_t *task_win, const uint8_t class_of_bucket[256], int32_t *order, BinResult *res,
              ClassRanges *classes);
// window-major copy of a column-major (colptr, rows, vals) matrix; see Source::T2
int build_window_major(cbgpu_ctx_impl *ctx, const int64_t *colptr, const int32_t *rows, const void *vals, int vbytes,
                       int64_t ncols, int64_t nnz, int nwin, int wlog2, int64_t **T2, int32_t **Wir, void **Wval);
int build_window_table(cbgpu_ctx_impl *ctx, const int64_t *colptr, const int32_t *rows, int64_t ncols, int nwin,
                       int wlog2, int64_t **T);

// range of buckets [lo, hi] -> contiguous slice of `order` (because of the descending layout)
inline void bucket_range(const BinResult &r, int lo, int hi, int64_t *begin, int64_t *count) {
  int64_t c = 0;
  for (int b = lo; b <= hi; ++b) c += r.count[b];
  *begin = r.offset[hi];
  *count = c;
}

} // namespace cbgpu
