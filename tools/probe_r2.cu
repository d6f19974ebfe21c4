// Round-2 hardware probes (B200, sm_100a) behind the design of score_stack.cuh / score_kron.cuh:
//   P1  register <-> (lane, column) mapping of tcgen05.ld.16x256b (the fragment shape the stacked epilogues rely on)
//   P2  tcgen05.ld throughput: 32x32b.x32 and 16x256b.x8, 4 / 8 / 16 warps
//   P4  no-swizzle K-major shared-memory operand: which descriptor field is the K-direction stride (LBO vs SBO)
//   P5  tcgen05.mma cycles, A from TMEM, N = 112 / 128 / 224 / 256, operand in the no-swizzle layout
//   P6  2-D tensor-map TMA: landing layout, out-of-bounds fill, and a pure-read bandwidth run (ring of S stages per CTA)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dct_pruning_b200/csrc -o tools/probe_r2_bin tools/probe_r2.cu
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "umma.cuh"
using namespace dctp::umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31]) : "r"(taddr) : "memory");
}

// ------------------------------------------------------------------ P1: fragment layout of 16x256b
__global__ void __launch_bounds__(128) p1_layout(uint32_t* out) {
    __shared__ uint32_t slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc<64>(&slot);
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = slot;
    uint32_t v[16];
    for (int part = 0; part < 2; ++part) {
        for (int i = 0; i < 16; ++i) v[i] = tid * 256 + part * 16 + i;           // lane tid, column part*16+i
        tmem_st16(tmem + ((warp * 32u) << 16) + part * 16, v);
    }
    tmem_st_wait();
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    for (int half = 0; half < 2; ++half) {
        uint32_t r[8];
        tmem_ld_16x256b_x2(tmem + ((warp * 32u + half * 16u) << 16), r);
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) out[((warp * 2 + half) * 32 + (tid & 31)) * 8 + i] = r[i];
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

// ------------------------------------------------------------------ P7: fragment layout of tcgen05.st.16x128b
__global__ void __launch_bounds__(128) p7_st_layout(uint32_t* out) {
    __shared__ uint32_t slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc<64>(&slot);
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = slot;
    for (int half = 0; half < 2; ++half) {
        uint32_t v[4];
        for (int i = 0; i < 4; ++i) v[i] = 0x10000u * (half * 32 + lane) + i;      // (thread id within the instruction, register)
        asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(tmem + ((warp * 32u + half * 16u) << 16)),
                     "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
    }
    tmem_st_wait();
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    uint32_t r[16];
    tmem_ld16(tmem + ((warp * 32u) << 16), r);
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out[tid * 8 + i] = r[i];
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

// ------------------------------------------------------------------ P2: tcgen05.ld throughput
template <int SHAPE>   // 0: 32x32b.x32 (4 KB per warp instruction), 1: 16x256b.x8 (2 KB)
__global__ void __launch_bounds__(512) p2_ldtm(long long* out, int reps, uint32_t* sink) {
    __shared__ uint32_t slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = slot;
    const uint32_t lane_base = ((warp & 3u) * 32u) << 16;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        uint32_t a[32], b[32];
        if (SHAPE == 0) {
            tmem_ld_32x32b_x32(tmem + lane_base + ((r * 64) & 255) + (warp >> 2) * 0, a);
            tmem_ld_32x32b_x32(tmem + lane_base + ((r * 64 + 32) & 255), b);
        } else {
            tmem_ld_16x256b_x8(tmem + lane_base + ((r * 64) & 255), a);
            tmem_ld_16x256b_x8(tmem + lane_base + (16u << 16) + ((r * 64) & 255), b);
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += a[i] ^ b[i];
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[tid] = acc;
    if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------ P4/P5: no-swizzle K-major B operand, A from TMEM
// B element (n, k) at  (k/8)*KSTRIDE + (n/8)*128 + (n%8)*16 + (k%8)*2   (each 16-byte row of a core matrix is one k-chunk of one n)
template <int N, bool SW = false>
__global__ void __launch_bounds__(128) p4_mma(float* d_out, long long* cyc, int swap_fields, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    constexpr uint32_t KSTRIDE = N * 16 + 16;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 8 * (256 * 16 + 16) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    __syncthreads();
    for (int idx = tid; idx < N * 64; idx += 128) {
        const int n = idx / 64, k = idx % 64;
        const float val = float((n * 5 + k) % 7 - 3);
        const uint32_t off = SW ? (n / 8) * 1024 + (n % 8) * 128 + (((k / 8) ^ (n % 8)) * 16) + (k % 8) * 2
                                : (k / 8) * KSTRIDE + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<uint16_t*>(smem + off) = static_cast<uint16_t>(__float_as_uint(val) >> 16);
    }
    if (warp == 0) tmem_alloc<512>(&slot);
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    fence_async_smem(); tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = slot;
    {   // A[m][k] = (m*7 + k*3) % 5 - 2, packed pairs: column c of lane m = (A[m][2c] | A[m][2c+1] << 16)
        uint32_t v[16];
        for (int part = 0; part < 2; ++part) {
            for (int i = 0; i < 16; ++i) {
                const int c = part * 16 + i, m = tid;
                const float a0 = float((m * 7 + (2 * c) * 3) % 5 - 2), a1 = float((m * 7 + (2 * c + 1) * 3) % 5 - 2);
                v[i] = (__float_as_uint(a0) >> 16) | (__float_as_uint(a1) & 0xFFFF0000u);
            }
            tmem_st16(tmem + ((warp * 32u) << 16) + part * 16, v);
        }
        tmem_st_wait();
    }
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint64_t desc = SW ? make_smem_desc(smem_u32(smem), 16, 1024, SWIZZLE_128B) : swap_fields ? make_smem_desc(smem_u32(smem), 128, KSTRIDE, SWIZZLE_NONE)     // LBO = 128 (n direction), SBO = k stride
                                      : make_smem_desc(smem_u32(smem), KSTRIDE, 128, SWIZZLE_NONE);    // LBO = k stride, SBO = 128 (n direction)
    if (warp == 0) {
        if (elect_one()) {
            uint32_t phase = 0;
            const long long t0 = clock64();
            const uint32_t dlo = static_cast<uint32_t>(desc);
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        mma_bf16_ts(tmem + 256, tmem + 8 * ks, desc_with_lo(desc, dlo + (SW ? 2 * ks : ((2 * ks * KSTRIDE) >> 4))), idesc, (pass | ks) != 0 ? 1u : 0u);
                mma_commit(&bar);
                mbar_wait(&bar, phase); phase ^= 1;
            }
            const long long t1 = clock64();
            if (blockIdx.x == 0) cyc[0] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    // D = last rep only (its first MMA overwrote): two passes over the same operands -> 2 * A * B^T
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + 256 + ((warp * 32u) << 16) + c0, r);
        tmem_ld_wait();
        if (blockIdx.x == 0)
            for (int i = 0; i < 16; ++i) d_out[tid * 256 + c0 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------ P6: TMA 2-D tensor map
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// one tile into shared memory, copied out verbatim
__global__ void __launch_bounds__(128) p6_layout(const __grid_constant__ CUtensorMap map, float* out, int row0, int box_rows) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const uint32_t tid = threadIdx.x;
    for (int i = tid; i < box_rows * 32; i += 128) reinterpret_cast<float*>(smem)[i] = -7.f;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar, box_rows * 128);
        tma_load_2d(smem, &map, 0, row0, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = tid; i < box_rows * 32; i += 128) out[i] = reinterpret_cast<float*>(smem)[i];
}
// pure read: every CTA walks tiles blockIdx.x, +gridDim.x, ... through a ring of STAGES buffers
template <int STAGES>
__global__ void __launch_bounds__(64) p6_read(const __grid_constant__ CUtensorMap map, int box_rows, int num_tiles, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t tile_bytes = box_rows * 128;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init_fence();
    }
    __syncthreads();
    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(empty + s, ((it / STAGES) - 1) & 1);
                mbar_arrive_expect_tx(full + s, tile_bytes);
                tma_load_2d(smem + s * tile_bytes, &map, 0, tile * box_rows, full + s);
            }
        }
    } else {
        float acc = 0.f;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int s = it % STAGES;
            mbar_wait(full + s, (it / STAGES) & 1);
            acc += reinterpret_cast<float*>(smem + s * tile_bytes)[(tid & 31) * 4];
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(empty + s);
        }
        if (acc == 123.456f) sink[tid] = acc;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    CK(cudaSetDevice(0));
    // ---------------- P1
    {
        uint32_t* d; CK(cudaMalloc(&d, 4 * 2 * 32 * 8 * 4));
        p1_layout<<<1, 128>>>(d);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> h(4 * 2 * 32 * 8);
        CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
        printf("P1 16x256b.x2: thread t register i -> (lane, column)   [warp 0, lane base 0]\n");
        for (int t = 0; t < 32; ++t) {
            printf("  t%02d:", t);
            for (int i = 0; i < 8; ++i) printf(" (%u,%u)", h[(0 * 32 + t) * 8 + i] >> 8, h[(0 * 32 + t) * 8 + i] & 255);
            printf("\n");
        }
        printf("  [warp 0, lane base 16] t0: ");
        for (int i = 0; i < 8; ++i) printf(" (%u,%u)", h[(1 * 32 + 0) * 8 + i] >> 8, h[(1 * 32 + 0) * 8 + i] & 255);
        printf("\n  [warp 2, lane base 64] t5: ");
        for (int i = 0; i < 8; ++i) printf(" (%u,%u)", h[(4 * 32 + 5) * 8 + i] >> 8, h[(4 * 32 + 5) * 8 + i] & 255);
        printf("\n");
        // check the expected mapping: r[4g + 0,1] = (t/4, 8g + 2(t%4) + 0,1), r[4g + 2,3] = (t/4 + 8, same)
        int bad = 0;
        for (int w = 0; w < 4; ++w) for (int half = 0; half < 2; ++half) for (int t = 0; t < 32; ++t) for (int i = 0; i < 8; ++i) {
            const uint32_t got = h[((w * 2 + half) * 32 + t) * 8 + i];
            const uint32_t lane = w * 32 + half * 16 + t / 4 + ((i & 2) ? 8 : 0), col = (i / 4) * 8 + 2 * (t % 4) + (i & 1);
            if (got != lane * 256 + col) ++bad;
        }
        printf("P1 expected-mapping mismatches: %d\n", bad);
    }
    // ---------------- P7
    {
        uint32_t* d; CK(cudaMalloc(&d, 128 * 8 * 4));
        p7_st_layout<<<1, 128>>>(d);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> h(128 * 8);
        CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        // expected: thread t register 2g + s of the instruction at lane base L wrote lane L + t/4 + 8s, column 4g + t%4
        for (int lane = 0; lane < 128; ++lane) for (int col = 0; col < 8; ++col) {
            const int half = (lane % 32) / 16, l16 = lane % 16, t = (l16 % 8) * 4 + col % 4, reg = 2 * (col / 4) + l16 / 8;
            if (h[lane * 8 + col] != 0x10000u * (half * 32 + t) + reg) ++bad;
        }
        printf("P7 st.16x128b.x2 expected-mapping mismatches: %d\n", bad);
        if (bad) for (int lane = 0; lane < 32; ++lane) {
            printf("  lane %2d:", lane);
            for (int col = 0; col < 8; ++col) printf(" (t%u,r%u)", (h[lane * 8 + col] >> 16) & 31, h[lane * 8 + col] & 0xffff);
            printf("\n");
        }
    }
    // ---------------- P2
    {
        long long* d; CK(cudaMalloc(&d, 8));
        uint32_t* sink; CK(cudaMalloc(&sink, 4096));
        const int reps = 2000;
        for (int nt : {128, 256, 512}) {
            long long h;
            p2_ldtm<0><<<148, nt>>>(d, reps, sink); p2_ldtm<0><<<148, nt>>>(d, reps, sink);
            CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
            printf("P2 32x32b.x32  %2d warps: %.1f cycles per pair of loads per warp -> %.1f B/cycle/SM\n", nt / 32, double(h) / reps,
                   double(nt / 32) * 8192.0 * reps / double(h));
            p2_ldtm<1><<<148, nt>>>(d, reps, sink); p2_ldtm<1><<<148, nt>>>(d, reps, sink);
            CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
            printf("P2 16x256b.x8  %2d warps: %.1f cycles per pair of loads per warp -> %.1f B/cycle/SM\n", nt / 32, double(h) / reps,
                   double(nt / 32) * 4096.0 * reps / double(h));
        }
    }
    // ---------------- P4 / P5
    {
        float* d; CK(cudaMalloc(&d, 128 * 256 * 4));
        long long* cyc; CK(cudaMalloc(&cyc, 8));
        std::vector<float> h(128 * 256);
        auto check = [&](int N, int swap) {
            CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
                float want = 0;
                for (int k = 0; k < 64; ++k) want += float((m * 7 + k * 3) % 5 - 2) * float((n * 5 + k) % 7 - 3);
                if (h[m * 256 + n] != 2 * want) ++bad;
            }
            long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
            printf("P4 N=%3d %s: %d mismatches of %d   | P5 %.1f cycles per MMA (8 per commit+wait)\n", N,
                   swap ? "LBO=128(n)  SBO=kstride" : "LBO=kstride SBO=128(n) ", bad, 128 * N, double(c) / (64 * 8));
        };
        const size_t sm = 8 * (256 * 16 + 16) + 1024;
        CK(cudaFuncSetAttribute(p4_mma<112>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaFuncSetAttribute(p4_mma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaFuncSetAttribute(p4_mma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaFuncSetAttribute(p4_mma<224>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaFuncSetAttribute(p4_mma<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        p4_mma<112><<<148, 128, sm>>>(d, cyc, 0, 64); CK(cudaDeviceSynchronize()); check(112, 0);
        p4_mma<64><<<148, 128, sm>>>(d, cyc, 0, 64); CK(cudaDeviceSynchronize()); check(64, 0);
        CK(cudaFuncSetAttribute(p4_mma<112, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaFuncSetAttribute(p4_mma<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        printf("(next two: SWIZZLE_128B layout)\n");
        p4_mma<112, true><<<148, 128, sm>>>(d, cyc, 0, 64); CK(cudaDeviceSynchronize()); check(112, 0);
        p4_mma<64, true><<<148, 128, sm>>>(d, cyc, 0, 64); CK(cudaDeviceSynchronize()); check(64, 0);
        p4_mma<128><<<148, 128, sm>>>(d, cyc, 0, 64); CK(cudaDeviceSynchronize()); check(128, 0);
        p4_mma<224><<<148, 128, sm>>>(d, cyc, 0, 64); CK(cudaDeviceSynchronize()); check(224, 0);
        p4_mma<256><<<148, 128, sm>>>(d, cyc, 0, 64); CK(cudaDeviceSynchronize()); check(256, 0);
    }
    // ---------------- P6
    {
        EncodeFn encode = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode), cudaEnableDefault, &qres));
        if (!encode) { printf("P6: no cuTensorMapEncodeTiled\n"); return 0; }
        const size_t total = 256ull * 256 * 56 * 56;              // [256,256,56,56] fp32 = 822 MB
        float* x; CK(cudaMalloc(&x, total * 4 + 4096));
        {
            std::vector<float> h(1 << 20);
            for (size_t i = 0; i < h.size(); ++i) h[i] = float(i);
            CK(cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        }
        const int box_rows = 196;                                  // 196 x 128 B = 25088 B = two 56x56 maps
        for (int rows_avail : {100, 1 << 30}) {
            CUtensorMap map;
            cuuint64_t gdim[2] = {32, rows_avail < (1 << 30) ? (cuuint64_t)rows_avail : total / 32};
            cuuint64_t gstr[1] = {128};
            cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, estr[2] = {1, 1};
            CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            printf("P6 encode rows=%llu -> %d\n", (unsigned long long)gdim[1], (int)r);
            if (r != CUDA_SUCCESS) continue;
            if (rows_avail < (1 << 30)) {
                float* out; CK(cudaMalloc(&out, box_rows * 128));
                CK(cudaFuncSetAttribute(p6_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
                p6_layout<<<1, 128, 32 * 1024>>>(map, out, 0, box_rows);
                CK(cudaDeviceSynchronize());
                std::vector<float> h(box_rows * 32);
                CK(cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost));
                int bad = 0, badz = 0;
                for (int i = 0; i < box_rows * 32; ++i) {
                    if (i < 100 * 32) { if (h[i] != float(i)) ++bad; }
                    else if (h[i] != 0.f) ++badz;
                }
                printf("P6 layout: %d mismatches in the 100 in-bounds rows, %d non-zero of the %d out-of-bounds rows' elements (first oob value %g)\n",
                       bad, badz, (box_rows - 100) * 32, h[100 * 32]);
            } else {
                const int num_tiles = (int)(total / 32 / box_rows);
                float* sink; CK(cudaMalloc(&sink, 4096));
                cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                auto run = [&](auto kern, int stages, int ctas_per_sm) {
                    const size_t sm = (size_t)stages * box_rows * 128;
                    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                    kern<<<148 * ctas_per_sm, 64, sm>>>(map, box_rows, num_tiles, sink);
                    CK(cudaEventRecord(e0));
                    for (int i = 0; i < 5; ++i) kern<<<148 * ctas_per_sm, 64, sm>>>(map, box_rows, num_tiles, sink);
                    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    printf("P6 TMA read, %d stages x 25 KB, %d CTA/SM: %.3f ms per pass = %.0f GB/s\n", stages, ctas_per_sm, ms / 5,
                           double(num_tiles) * box_rows * 128 / (ms / 5 * 1e6));
                };
                run(p6_read<2>, 2, 1); run(p6_read<3>, 3, 1); run(p6_read<4>, 4, 1); run(p6_read<6>, 6, 1); run(p6_read<8>, 8, 1);
                run(p6_read<2>, 2, 2); run(p6_read<4>, 4, 2);
            }
        }
    }
    return 0;
}
