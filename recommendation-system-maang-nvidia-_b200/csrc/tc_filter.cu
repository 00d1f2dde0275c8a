// tc_filter.cu -- tensor-core (tcgen05) pre-filter + exact fp32 re-score for batched retrieval.
// Placeholder until the tcgen05 path lands: reports "not applicable", so RS_AUTO uses the exact scan
// and RS_TENSOR_BF16 fails loudly.
#include "flat_index.cuh"

namespace rs {

bool tensor_path_supported(const rs_index*, int64_t, int) { return false; }

void search_tensor_device(rs_index*, const float*, int64_t, int, int64_t, float*, int64_t, int64_t*, int64_t,
                          cudaStream_t) {
  RS_CHECK(false, "tensor-core search path is not built");
}

}  // namespace rs
