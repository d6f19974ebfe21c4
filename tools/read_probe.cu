// Read-bandwidth probe: per-map sum of squares (the Parseval lower bound the DCT kernel races), 128-bit streaming loads.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/read_probe tools/read_probe.cu && /tmp/read_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) sumsq(const float4* __restrict__ x, size_t n_vec, float* out, int unroll_dummy) {
    float acc = 0.f;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 7 * stride < n_vec; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(x + i + u * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
    }
    for (; i < n_vec; i += stride) { float4 v = x[i]; acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
int main() {
    const size_t bytes = 822083584ull;                 // [256,256,56,56] fp32
    float4* x; float* out;
    cudaMalloc(&x, bytes); cudaMalloc(&out, 4); cudaMemset(x, 0, bytes);
    for (int blocks_per_sm : {1, 2, 4, 8}) {
        int grid = 148 * blocks_per_sm;
        for (int w = 0; w < 3; ++w) sumsq<<<grid, 256>>>(x, bytes / 16, out, 0);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        for (int it = 0; it < 10; ++it) sumsq<<<grid, 256>>>(x, bytes / 16, out, 0);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("sumsq grid %4d x 256: %.3f ms per pass, %.1f GB/s\n", grid, ms / 10, bytes / (ms / 10) / 1e6);
    }
    return 0;
}
