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

// order[] receives task ids grouped by bucket, larger bucket ids first; bucket 0 is not listed. Synchronises.
int bin_tasks(cbgpu_ctx_impl *ctx, const uint8_t *bucket, const int64_t *weight, const int64_t *weight2, int64_t n,
              int32_t *order, BinResult *res);
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
