// Micro-benchmark behind the visit-table design of the walk kernel (csrc/walk_topt.cu): cost of
// match.any, shared-memory atomics and plain LDS/STS with 32 spread addresses per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/ubench_smem tools/ubench_smem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int kOp>
__global__ void bench(uint32_t* out, long long* cyc, int iters, int distinct) {
    __shared__ uint32_t tab[8][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < 512; i += 32) tab[warp][i] = 0;
    __syncwarp();
    uint32_t v = (lane % distinct) * 2654435761u + blockIdx.x, acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const uint32_t h = (v >> 23) & 511u;
        if (kOp == 0) acc += __match_any_sync(0xFFFFFFFFu, v);
        if (kOp == 1) acc += atomicCAS(&tab[warp][h], 0u, v);
        if (kOp == 2) acc += atomicAdd(&tab[warp][h], 1u);
        if (kOp == 3) atomicMin(&tab[warp][h], v);
        if (kOp == 4) { acc += tab[warp][h]; tab[warp][h] = v; }
        if (kOp == 5) acc += __ballot_sync(0xFFFFFFFFu, v & 1);
        if (kOp == 6) acc += __reduce_max_sync(0xFFFFFFFFu, v);
        v = v * 1664525u + 1013904223u + acc;
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + v;
}

template <int kOp>
void run(const char* name, int distinct) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&cyc, 148 * 8 * 8 * 8);
    const int iters = 4096;
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int blocks = cfg ? 148 * 6 : 1, threads = cfg ? 256 : 32;
        bench<kOp><<<blocks, threads>>>(out, cyc, iters, distinct);   // warm
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        bench<kOp><<<blocks, threads>>>(out, cyc, iters, distinct);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
        if (!cfg) printf("%-28s distinct=%2d  1 warp: %7.1f cyc/iter", name, distinct, (double)c0 / iters);
        else printf("   48 warps/SM: %7.2f cyc/warp-instr/SM (%.3f ms)\n", ms * 1e-3 * 1.9e9 / (double)(iters * 48), ms);
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int d : {32, 8, 1}) run<0>("match.any", d);
    for (int d : {32, 8, 1}) run<1>("atomicCAS smem", d);
    for (int d : {32, 1}) run<2>("atomicAdd smem", d);
    for (int d : {32, 1}) run<3>("atomicMin smem", d);
    for (int d : {32, 1}) run<4>("LDS+STS", d);
    run<5>("ballot", 32);
    run<6>("redux.max", 32);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
