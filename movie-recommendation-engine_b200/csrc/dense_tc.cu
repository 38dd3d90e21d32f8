// dense_tc.cu -- tcgen05 (kind::tf32) path of pb200_gather_dense.  Placeholder until the
// tensor-core kernel lands: reports UNSUPPORTED so callers fail loudly instead of silently
// running something else.
#include "pool.cuh"

namespace pb200 {
struct DenseParams;
int gather_dense_tf32(const DenseParams&, cudaStream_t) {
    set_error("gather_dense: PB200_PREC_TF32 (tcgen05) path not built in this version");
    return PB200_ERR_UNSUPPORTED;
}
}  // namespace pb200
