// dense.cuh -- parameters shared by the CUDA-core (dense.cu) and tcgen05 (dense_tc.cu) paths
// of pb200_gather_dense.
#pragma once
#include "pool.cuh"

namespace pb200 {

struct DenseParams {
    const float* __restrict__ a1; int k1;
    const float* __restrict__ a2; int k2;
    const float* __restrict__ pool_x;
    ListArgs lists;
    const float* __restrict__ w;      // [N, K], nn.Linear layout
    const float* __restrict__ bias;   // [N] or null
    const float* __restrict__ ln_gamma;
    const float* __restrict__ ln_beta;
    int64_t n; int n_out; int flags;
    float* __restrict__ out;
};

int gather_dense_fp32(const DenseParams& p, cudaStream_t stream);
int gather_dense_tf32(const DenseParams& p, cudaStream_t stream);
bool gather_dense_tf32_supported(const DenseParams& p);

}  // namespace pb200
