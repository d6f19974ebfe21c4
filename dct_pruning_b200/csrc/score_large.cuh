// DCT-score hook kernel for LARGE square maps (128 < N <= 320, N % 16 == 0: U^2-Netp's 144 / 160 / 288 / 320 stages).
//
// Same contract as score_umma.cuh (/root/reference/utils/common.py:262-277 and, for the side inputs,
// :296-309: per-(image,channel) orthonormal 2-D DCT-II energy, summed per channel).  A map no longer fits on
// chip (320 x 320 fp32 = 400 KB; the bf16 hi/lo basis is another 400 KB), so the separable contraction is tiled:
//
//   work item   (map, v-chunk): 128 output columns v of Y = X C^T / 128 rows of Z^T
//   for each h-tile (128 rows of the map):
//     stage 1   D1[h, v] = sum_w X[h,w] C[v,w]      K-loop over 64-wide w blocks; A = X slab (fp32 -> bf16 hi/lo,
//               K-major), B = C[v-chunk, w-block] slab copied from L2; accumulates in TMEM columns [0,128)
//     epi   1   D1 row h -> bf16 hi/lo -> A2[k = h][m = v]  MN-major in shared memory (the transpose is free)
//     stage 2   D2[v, u] += sum_{h in tile} A2[v,h] C[u,h]   for u-chunks of <= 160, B = C[u-chunk, h-block] slab;
//               D2 (N columns) stays in TMEM columns [128, 128+N) across all h-tiles: the K-split over h
//   epi   2   D2 -> sum of squares per lane -> block reduction -> one fp64 atomicAdd per (map, v-chunk)
//
// Every product is three bf16 MMAs (hi*hi + lo*hi + hi*lo, fp32 accumulate), like the small-map kernels.
// Phases run back to back (one mbarrier, strictly alternating phase); the kernel is tensor-bound by design
// (2*N^3 flops per stage and map), the map is re-read once per v-chunk out of L2, the basis slabs come from L2.
// The operands of the next step (map slab, basis slab) are loaded into registers while the tensor core works on
// the current one; there is one resident CTA of 16 warps per SM (141 KB of shared memory, 512 TMEM columns).
#pragma once
#include "score_umma.cuh"

namespace dctp {

struct LargeScoreArgs {
    const float* x_dense;           // first scored element; all scored maps back to back (stride_h == N), 16-B aligned
    int n_maps, c_count;
    int N, NP;                      // map side; basis leading dimension (N rounded up to 64, zero padded)
    int NVC;                        // v-chunks per map = ceil(N / 128)
    int NU, NUC;                    // u-chunk width (N if N <= 160 else N/2 rounded up to 16) and count
    int n_items;                    // n_maps * NVC
    const uint16_t* c_hi;           // [NP][NP] bf16 bits of C_N (row = output index, col = contraction index), zero padded
    const uint16_t* c_lo;
    double* accum;
    float* energy_out;              // optional [n_maps], accumulated with float atomics over the v-chunks (caller zeroes)
    float* dump;                    // optional [n_maps][N][N] coefficients Z[u][v]
    int* status;
};

struct LargeSmem {
    static constexpr uint32_t A1_HALF = 128 * 128;                 // X slab: 128 rows x 64 k (hi or lo)
    static constexpr uint32_t B_HALF = 160 * 128;                  // basis slab: up to 160 rows x 64 k
    static constexpr uint32_t A2_HALF = 128 * 128 * 2;             // A2: 128 k-rows x 128 m, MN-major
    static constexpr uint32_t OFF_A1_HI = 0, OFF_A1_LO = A1_HALF;
    static constexpr uint32_t OFF_B_HI = 2 * A1_HALF, OFF_B_LO = OFF_B_HI + B_HALF;
    static constexpr uint32_t OFF_A2_HI = ((OFF_B_LO + B_HALF + 1023) / 1024) * 1024, OFF_A2_LO = OFF_A2_HI + A2_HALF;
    static constexpr uint32_t OFF_CTRL = OFF_A2_LO + A2_HALF;
    static constexpr uint32_t TOTAL = OFF_CTRL + 64 + 512 * 4;
};

constexpr int LARGE_NT = 512;          // 16 warps: 4 per scheduler hide the latency of the load / convert / store phases;
                                       // warp w owns TMEM lane quarter w % 4 and every 4th 16-column block (w / 4) of an epilogue

__global__ void __launch_bounds__(LARGE_NT, 1) score_large_kernel(const LargeScoreArgs a) {
    constexpr int NT = LARGE_NT;
    constexpr int XV = 2048 / NT, BV = (1280 + NT - 1) / NT;       // prefetch registers: float4 of the map slab, uint4 of a basis slab
    using S = LargeSmem;
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a1_hi = smem + S::OFF_A1_HI;
    uint8_t* a1_lo = smem + S::OFF_A1_LO;
    uint8_t* b_hi = smem + S::OFF_B_HI;
    uint8_t* b_lo = smem + S::OFF_B_LO;
    uint8_t* a2_hi = smem + S::OFF_A2_HI;
    uint8_t* a2_lo = smem + S::OFF_A2_LO;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::OFF_CTRL);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::OFF_CTRL + 8);
    float* red = reinterpret_cast<float*>(smem + S::OFF_CTRL + 64);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (tid == 0) atomicExch(a.status, DCTP_DEV_SMEM_ALIGN);
        return;
    }
    for (uint32_t off = tid * 16; off < S::OFF_CTRL; off += NT * 16) *reinterpret_cast<uint4*>(smem + off) = make_uint4(0, 0, 0, 0);
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_idx = (warp & 3) * 32 + (tid & 31);        // this thread's TMEM lane
    const uint32_t col_blk = warp >> 2;                            // which 16-column blocks of an epilogue it handles
    const uint32_t tmem_lane = tmem + (((warp & 3) * 32u) << 16);
    const uint32_t d1_col = 0, d2_col = 128;

    const uint64_t desc_k = make_smem_desc(0, 16, 1024, SWIZZLE_128B);
    const uint64_t desc_mn = make_smem_desc(0, 16384, 1024, SWIZZLE_128B);        // A2: 64-wide M blocks 16 KB apart
    const uint32_t k_lo = static_cast<uint32_t>(desc_k), mn_lo = static_cast<uint32_t>(desc_mn);
    const uint32_t lo_a1_hi = smem_u32(a1_hi) >> 4, lo_a1_lo = smem_u32(a1_lo) >> 4;
    const uint32_t lo_b_hi = smem_u32(b_hi) >> 4, lo_b_lo = smem_u32(b_lo) >> 4;
    const uint32_t lo_a2_hi = smem_u32(a2_hi) >> 4, lo_a2_lo = smem_u32(a2_lo) >> 4;

    const int N = a.N, NN = N * N;
    uint32_t phase = 0;
    bool alive = true;
    auto issue_mmas = [&](bool stage2, int ksteps, int kofs, uint32_t dcol, uint32_t idesc, bool acc0) {
        if (warp == 0) {                                           // one elected thread issues 3 passes x ksteps MMAs
            if (elect_one()) {
                tc_fence_after_sync();
                if (!stage2)
                    detail::issue_ss3_n<64, false>(ksteps, tmem + dcol, k_lo + lo_a1_hi, k_lo + lo_a1_lo, k_lo + lo_b_hi, k_lo + lo_b_lo,
                                                   desc_k, desc_k, idesc, acc0);
                else
                    detail::issue_ss3_n<64, true>(ksteps, tmem + dcol, mn_lo + lo_a2_hi + kofs * 128, mn_lo + lo_a2_lo + kofs * 128,
                                                  k_lo + lo_b_hi, k_lo + lo_b_lo, desc_mn, desc_k, idesc, acc0);
                mma_commit(bar);
            }
            __syncwarp();
        }
    };
    auto wait_mmas = [&]() {
        if (alive && !mbar_wait(bar, phase)) alive = false;
        phase ^= 1;
        tc_fence_after_sync();
    };

    // prefetch registers: the operands of the NEXT step are loaded while the tensor core works on the current one
    float4 xr[XV];
    uint4 bh[BV], bl[BV];
    uint32_t x_vpr = 1, x_total = 0, b_total = 0;
    auto load_x = [&](const float* xm, int h0, int w0) {
        const int MH = min(128, N - h0), kvalid = min(64, N - w0);
        x_vpr = kvalid / 4;                                        // float4 vectors per row (4 / 8 / 12 / 16)
        x_total = (uint32_t)MH * x_vpr;
#pragma unroll
        for (int j = 0; j < XV; ++j) {
            const uint32_t i = tid + j * NT;
            if (i < x_total) {
                const uint32_t r = x_vpr == 16 ? i >> 4 : i / x_vpr, q = i - r * x_vpr;
                xr[j] = detail::ldg_stream(reinterpret_cast<const float4*>(xm + (size_t)(h0 + r) * N + w0) + q);
            }
        }
    };
    auto store_x = [&]() {
#pragma unroll
        for (int j = 0; j < XV; ++j) {
            const uint32_t i = tid + j * NT;
            if (i < x_total) {
                const uint32_t r = x_vpr == 16 ? i >> 4 : i / x_vpr, q = i - r * x_vpr;
                detail::Scatter<1>::st(a1_hi, a1_lo, static_cast<uint16_t>(detail::kmajor_off(r, q * 4, 128)), xr[j]);
            }
        }
    };
    auto load_b = [&](int row0, int col0, int rows) {              // [rows x 64] block of C (hi and lo) at (row0, col0)
        b_total = (uint32_t)rows * 8;
        const uint16_t* sh = a.c_hi + (size_t)row0 * a.NP + col0;
        const uint16_t* sl = a.c_lo + (size_t)row0 * a.NP + col0;
#pragma unroll
        for (int j = 0; j < BV; ++j) {
            const uint32_t i = tid + j * NT;
            if (i < b_total) {
                const size_t o = (size_t)(i >> 3) * a.NP + (i & 7) * 8;
                bh[j] = *reinterpret_cast<const uint4*>(sh + o);
                bl[j] = *reinterpret_cast<const uint4*>(sl + o);
            }
        }
    };
    auto store_b = [&]() {
#pragma unroll
        for (int j = 0; j < BV; ++j) {
            const uint32_t i = tid + j * NT;
            if (i < b_total) {
                const uint32_t r = i >> 3, ch = i & 7;
                const uint32_t off = (r >> 3) * 1024u + (r & 7) * 128u + ((ch ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(b_hi + off) = bh[j];
                *reinterpret_cast<uint4*>(b_lo + off) = bl[j];
            }
        }
    };

    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int map = item / a.NVC, vc = item - map * a.NVC;
        const int v0 = vc * 128, MV = min(128, N - v0), MV16 = (MV + 15) & ~15;
        const float* xm = a.x_dense + (size_t)map * NN;
        const uint32_t idesc1 = make_idesc_bf16(128, MV16, false, false);
        const uint32_t idesc2 = make_idesc_bf16(128, a.NU, true, false);

        load_x(xm, 0, 0);
        load_b(v0, 0, MV16);
        for (int h0 = 0, ht = 0; h0 < N; h0 += 128, ++ht) {
            const int MH = min(128, N - h0);
            // ---- stage 1: D1[h, v] = sum_w X[h,w] C[v,w], K-loop over 64-wide w blocks
            for (int w0 = 0; w0 < N; w0 += 64) {
                const int kvalid = min(64, N - w0);                // multiple of 16
                store_x();                                         // X slab rows h0.., columns w0.. (fp32 -> bf16 hi/lo, K-major)
                store_b();                                         // basis slab rows v0..v0+MV16, columns w0..w0+64
                fence_async_smem();
                tc_fence_before_sync();
                __syncthreads();
                issue_mmas(false, kvalid / 16, 0, d1_col, idesc1, w0 != 0);
                if (w0 + 64 < N) {                                 // next step's operands: in flight while the tensor core works
                    load_x(xm, h0, w0 + 64);
                    load_b(v0, w0 + 64, MV16);
                } else {
                    load_b(0, h0, a.NU);                           // first stage-2 slab of this h-tile
                }
                wait_mmas();
            }
            // ---- epilogue 1: D1 row h (lane) -> bf16 hi/lo -> A2[k = h][m = v]  (MN-major)
#pragma unroll 1
            for (int c0 = col_blk * 16; c0 < MV16; c0 += 64) {      // this warp's 16-column blocks
                uint32_t r[16];
                tmem_ld16(tmem_lane + d1_col + c0, r);
                tmem_ld_wait();
                if ((int)lane_idx < MH) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t h4[4], l4[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p)
                            split2(__uint_as_float(r[8 * half + 2 * p]), __uint_as_float(r[8 * half + 2 * p + 1]), h4[p], l4[p]);
                        const uint32_t off = detail::mnmajor_off(c0 + 8 * half, lane_idx, 16384);
                        *reinterpret_cast<uint4*>(a2_hi + off) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
                        *reinterpret_cast<uint4*>(a2_lo + off) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
                    }
                }
            }
            // ---- stage 2: D2[v, u] += sum_{h in tile} A2[v,h] C[u,h], per u-chunk and 64-wide h block
            const int nkb = (MH + 63) / 64;
            for (int uc = 0; uc < a.NUC; ++uc) {
                const int u0 = uc * a.NU;
                for (int kb = 0; kb < nkb; ++kb) {
                    const int kvalid = min(64, MH - kb * 64);
                    store_b();                                     // basis slab rows u0..u0+NU, columns h0 + 64*kb ..
                    fence_async_smem();
                    tc_fence_before_sync();
                    __syncthreads();
                    issue_mmas(true, kvalid / 16, kb * 4, d2_col + u0, idesc2, !(ht == 0 && kb == 0));
                    if (kb + 1 < nkb) load_b(u0, h0 + 64 * (kb + 1), a.NU);
                    else if (uc + 1 < a.NUC) load_b(u0 + a.NU, h0, a.NU);
                    else if (h0 + 128 < N) {                       // next h-tile's first stage-1 operands
                        load_x(xm, h0 + 128, 0);
                        load_b(v0, 0, MV16);
                    }
                    wait_mmas();
                }
            }
            __syncthreads();
        }

        // ---- epilogue 2: lane v < MV holds Z[:, v]; energy of the v-chunk = sum over lanes and all N columns
        float e = 0.f;
#pragma unroll 1
        for (int c0 = col_blk * 16; c0 < a.NU * a.NUC; c0 += 64) {
            uint32_t r[16];
            tmem_ld16(tmem_lane + d2_col + c0, r);
            tmem_ld_wait();
            if ((int)lane_idx < MV) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int u = c0 + i;
                    const float z = u < N ? __uint_as_float(r[i]) : 0.f;
                    e = fmaf(z, z, e);
                    if (a.dump != nullptr && u < N) a.dump[(size_t)map * NN + (size_t)u * N + v0 + lane_idx] = z;
                }
            }
        }
        tc_fence_before_sync();
        red[tid] = (int)lane_idx < MV ? e : 0.f;
        __syncthreads();
        if (tid < 32) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NT / 32; ++j) s += red[tid + 32 * j];          // fixed order
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (tid == 0) {
                atomicAdd(a.accum + (map % a.c_count), (double)s);
                if (a.energy_out) atomicAdd(a.energy_out + map, s);
            }
        }
        __syncthreads();
    }
    if (!alive && tid == 0) atomicExch(a.status, DCTP_DEV_MMA_TIMEOUT);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace dctp
