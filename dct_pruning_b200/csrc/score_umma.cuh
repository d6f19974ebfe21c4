// Fused DCT-score hook kernel on 5th-gen tensor cores (tcgen05 + TMEM), bf16x3 split precision.
//
// Replaces, for one hooked activation, the per-slice Python loop of
// /root/reference/utils/common.py:262-277 (dct_2d per (image, channel) -> sum of squared
// coefficients -> per-channel batch sum).  One launch reads every scored map from HBM exactly
// once, keeps the cosine basis resident in shared memory, runs C_H * X * C_W^T on tensor cores
// and reduces the coefficients to energies straight out of TMEM: no coefficient tensor is ever
// written to HBM (unless the debug `dump` pointer asks for it).
//
// Tile = 128 TMEM lanes.
//   TWO_STAGE (9 <= N <= 128):  G = 128 / Ms maps per tile (Ms = N rounded up to 8), lane = g*Ms + h.
//     stage 1  D1[(g,h), v] = sum_w  X_g[h,w]   * C[v,w]      A1 = split(X) K-major   (K = w)
//     epi   1  D1 -> bf16 hi/lo -> A2[(g,v), h]                A2 MN-major            (K = h)
//     stage 2  D2[(g,v), u] = sum_h  A2[(g,v),h] * C[u,h]  ==  Z_g[u,v]
//     epi   2  energy_g = sum_{u,v} D2^2  (fixed-order fp32 tree) -> fp64 atomicAdd per channel
//   single stage (N <= 8): lane = map, K = flattened map (N*N <= 64), basis = C (x) C (Kronecker),
//     D1[g, (u,v)] = Z_g[u,v], energy straight from the lane's row.
// Every product runs as three bf16 MMAs: hi*hi + lo*hi + hi*lo (fp32 accumulate in TMEM).
// Operands live in shared memory in the canonical SWIZZLE_128B layouts (8-row x 128-byte atoms).
#pragma once
#include "umma.cuh"

namespace dctp {

struct FastDiv {            // q = n / d for n < 2^32 / d
    uint32_t mul, d;
    __host__ void set(uint32_t div) { d = div; mul = static_cast<uint32_t>(0x100000000ull / div) + 1u; }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : __umulhi(n, mul); }
};

struct UmmaScoreArgs {
    const float* x;                 // activation base pointer (fp32, maps contiguous: stride_h = W, stride_w = 1)
    long long stride_b, stride_c;   // in elements
    int c_begin, c_count;           // scored channel window (DenseNet: last 12)
    int n_maps;                     // B * c_count
    int N;                          // map side (H == W)
    int NN;                         // N * N
    int row_len;                    // TWO_STAGE: N ; single stage: N*N
    int Ms;                         // TMEM lanes per map (TWO_STAGE: roundup8(N); single: 1)
    int G;                          // maps per tile = 128 / Ms
    int num_tiles;
    FastDiv div_vpm, div_row, div_ms;
    const uint16_t* basis_hi;       // [KP x KP] bf16 bits, row = output index, col = contraction index, zero padded
    const uint16_t* basis_lo;
    double* accum;                  // [c_count] per-channel energy sums (fp64)
    float* energy_out;              // optional [n_maps] per-(image,channel) energies
    float* dump;                    // optional [n_maps x NN] DCT coefficients (debug / parity of the transform itself)
};

namespace detail {

// byte offset of bf16 element (row, k) in a K-major SWIZZLE_128B operand with `rows` rows
__device__ __forceinline__ uint32_t kmajor_off(uint32_t row, uint32_t k, uint32_t rows) {
    uint32_t kb = k >> 6, kk = k & 63;
    return kb * (rows * 128u) + (row >> 3) * 1024u + (row & 7) * 128u + ((((kk >> 3) ^ row) & 7) << 4) + ((kk & 7) << 1);
}
// byte offset of bf16 element (m, k) in an MN-major SWIZZLE_128B operand; lbo = stride between 64-wide M blocks
__device__ __forceinline__ uint32_t mnmajor_off(uint32_t m, uint32_t k, uint32_t lbo) {
    return (m >> 6) * lbo + (k >> 3) * 1024u + (k & 7) * 128u + (((((m & 63) >> 3) ^ k) & 7) << 4) + ((m & 7) << 1);
}

template <int VEC> struct Ld;
template <> struct Ld<4> {
    __device__ static __forceinline__ void ld(const float* p, float (&v)[4]) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
    }
    __device__ static __forceinline__ void st(uint8_t* hi, uint8_t* lo, uint32_t off, const float (&v)[4]) {
        uint32_t h0, l0, h1, l1;
        umma::split2(v[0], v[1], h0, l0);
        umma::split2(v[2], v[3], h1, l1);
        *reinterpret_cast<uint2*>(hi + off) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(lo + off) = make_uint2(l0, l1);
    }
};
template <> struct Ld<2> {
    __device__ static __forceinline__ void ld(const float* p, float (&v)[2]) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
    }
    __device__ static __forceinline__ void st(uint8_t* hi, uint8_t* lo, uint32_t off, const float (&v)[2]) {
        uint32_t h0, l0;
        umma::split2(v[0], v[1], h0, l0);
        *reinterpret_cast<uint32_t*>(hi + off) = h0;
        *reinterpret_cast<uint32_t*>(lo + off) = l0;
    }
};
template <> struct Ld<1> {
    __device__ static __forceinline__ void ld(const float* p, float (&v)[1]) {
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[0]) : "l"(p));
    }
    __device__ static __forceinline__ void st(uint8_t* hi, uint8_t* lo, uint32_t off, const float (&v)[1]) {
        uint32_t h0, l0;
        umma::split2(v[0], 0.f, h0, l0);
        *reinterpret_cast<uint16_t*>(hi + off) = static_cast<uint16_t>(h0 & 0xFFFFu);
        *reinterpret_cast<uint16_t*>(lo + off) = static_cast<uint16_t>(l0 & 0xFFFFu);
    }
};

constexpr uint32_t tmem_cols_for(int need) {
    return need <= 32 ? 32u : need <= 64 ? 64u : need <= 128 ? 128u : need <= 256 ? 256u : 512u;
}

}  // namespace detail

template <int KP, bool TWO_STAGE>
struct UmmaScoreSmem {
    static constexpr uint32_t KB = (KP + 63) / 64;                 // 64-element K blocks
    static constexpr uint32_t A_BYTES = KB * 128u * 128u;          // one 128-row K-major operand (hi or lo)
    static constexpr uint32_t A2_BYTES = 128u * KP * 2u;           // MN-major operand: 2 M blocks x KP/8 atoms
    static constexpr uint32_t B_BYTES = KB * KP * 128u;
    static constexpr uint32_t OFF_A1_HI = 0;
    static constexpr uint32_t OFF_A1_LO = OFF_A1_HI + A_BYTES;
    static constexpr uint32_t OFF_A2_HI = OFF_A1_LO + A_BYTES;
    static constexpr uint32_t OFF_A2_LO = OFF_A2_HI + (TWO_STAGE ? A2_BYTES : 0);
    static constexpr uint32_t OFF_B_HI = OFF_A2_LO + (TWO_STAGE ? A2_BYTES : 0);
    static constexpr uint32_t OFF_B_LO = OFF_B_HI + B_BYTES;
    static constexpr uint32_t OFF_MISC = OFF_B_LO + B_BYTES;       // barrier, tmem slot, reduction scratch
    static constexpr uint32_t MISC_BYTES = 16 + 128 * 4 + 128 * 8;
    static constexpr uint32_t TOTAL = OFF_MISC + MISC_BYTES + 1024;  // + slack for manual 1024-B alignment
    static constexpr uint32_t TMEM_COLS = detail::tmem_cols_for(TWO_STAGE ? 2 * KP : KP);
};

template <int KP, bool TWO_STAGE, int VEC>
__global__ void __launch_bounds__(128) score_umma_kernel(const UmmaScoreArgs a) {
    using S = UmmaScoreSmem<KP, TWO_STAGE>;
    using namespace umma;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a1_hi = smem + S::OFF_A1_HI;
    uint8_t* a1_lo = smem + S::OFF_A1_LO;
    uint8_t* a2_hi = smem + S::OFF_A2_HI;
    uint8_t* a2_lo = smem + S::OFF_A2_LO;
    uint8_t* b_hi = smem + S::OFF_B_HI;
    uint8_t* b_lo = smem + S::OFF_B_LO;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::OFF_MISC);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::OFF_MISC + 8);
    float* red = reinterpret_cast<float*>(smem + S::OFF_MISC + 16);
    const float** mptr = reinterpret_cast<const float**>(smem + S::OFF_MISC + 16 + 128 * 4);   // per-tile map base pointers

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: zero operands (pad rows/cols must stay finite zeros), stage the basis, TMEM, barrier
    for (uint32_t off = tid * 16; off < S::OFF_B_HI; off += 128 * 16)
        *reinterpret_cast<uint4*>(smem + off) = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < KP * (KP / 8); i += 128) {
        uint32_t n = i / (KP / 8), c8 = i % (KP / 8);
        uint4 vh = *reinterpret_cast<const uint4*>(a.basis_hi + n * KP + c8 * 8);
        uint4 vl = *reinterpret_cast<const uint4*>(a.basis_lo + n * KP + c8 * 8);
        uint32_t off = detail::kmajor_off(n, c8 * 8, KP);
        *reinterpret_cast<uint4*>(b_hi + off) = vh;
        *reinterpret_cast<uint4*>(b_lo + off) = vl;
    }
    if (warp == 0) tmem_alloc<S::TMEM_COLS>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_lane = tmem + ((warp * 32u) << 16);

    constexpr uint32_t IDESC1 = make_idesc_bf16(128, KP, false, false);
    constexpr uint32_t IDESC2 = make_idesc_bf16(128, KP, true, false);
    constexpr uint32_t A2_LBO = (KP / 8) * 1024u;
    const uint64_t d_a1_hi = make_smem_desc(smem_u32(a1_hi), 16, 1024, SWIZZLE_128B);
    const uint64_t d_a1_lo = make_smem_desc(smem_u32(a1_lo), 16, 1024, SWIZZLE_128B);
    const uint64_t d_b_hi = make_smem_desc(smem_u32(b_hi), 16, 1024, SWIZZLE_128B);
    const uint64_t d_b_lo = make_smem_desc(smem_u32(b_lo), 16, 1024, SWIZZLE_128B);
    const uint64_t d_a2_hi = make_smem_desc(smem_u32(a2_hi), A2_LBO, 1024, SWIZZLE_128B);
    const uint64_t d_a2_lo = make_smem_desc(smem_u32(a2_lo), A2_LBO, 1024, SWIZZLE_128B);

    // this thread's TMEM lane as (map-in-tile, row-in-map)
    const uint32_t my_g = a.div_ms.div(tid);
    const uint32_t my_r = tid - my_g * a.Ms;
    const uint32_t vpm = a.NN / VEC;                 // vectors per map
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int map0 = tile * a.G;
        const int maps_here = min(a.G, a.n_maps - map0);

        // ---- stage 0: HBM -> registers -> bf16 hi/lo -> A1 (each scored map is read exactly once)
        if ((int)tid < maps_here) {
            int m = map0 + (int)tid;
            int b = m / a.c_count, c = m - b * a.c_count;
            mptr[tid] = a.x + b * a.stride_b + (long long)(a.c_begin + c) * a.stride_c;
        }
        __syncthreads();
        {
            const uint32_t total = maps_here * vpm;
            constexpr int U = 4;
            for (uint32_t base = 0; base < total; base += 128 * U) {
                float v[U][VEC];
                uint32_t off[U];
                bool ok[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t idx = base + u * 128 + tid;
                    ok[u] = idx < total;
                    if (ok[u]) {
                        uint32_t g = a.div_vpm.div(idx);
                        uint32_t e = (idx - g * vpm) * VEC;
                        uint32_t h = a.div_row.div(e);
                        uint32_t w = e - h * a.row_len;
                        detail::Ld<VEC>::ld(mptr[g] + e, v[u]);
                        off[u] = detail::kmajor_off(g * a.Ms + h, w, 128);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (ok[u]) detail::Ld<VEC>::st(a1_hi, a1_lo, off[u], v[u]);
            }
        }
        fence_async_smem();
        __syncthreads();

        // ---- stage 1 MMA: D1 = A1 * B^T  (hi*hi + lo*hi + hi*lo)
        if (tid == 0) {
            tc_fence_after_sync();
            uint32_t acc = 0;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint64_t da = pass == 1 ? d_a1_lo : d_a1_hi;
                const uint64_t db = pass == 2 ? d_b_lo : d_b_hi;
#pragma unroll
                for (uint32_t ks = 0; ks < KP / 16; ++ks) {
                    uint32_t a_off = (ks >> 2) * (128u * 128u) + (ks & 3) * 32u;
                    uint32_t b_off = (ks >> 2) * (KP * 128u) + (ks & 3) * 32u;
                    mma_bf16_ss(tmem, desc_advance(da, a_off), desc_advance(db, b_off), IDESC1, acc);
                    acc = 1;
                }
            }
            mma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after_sync();

        const bool lane_valid = (int)my_g < maps_here && my_r < (uint32_t)(TWO_STAGE ? a.N : 1);
        float energy = 0.f;

        if constexpr (TWO_STAGE) {
            // ---- epilogue 1: D1 row (g,h) -> bf16 hi/lo -> A2[(g, v), h]   (transpose-free: MN-major operand)
#pragma unroll 1
            for (uint32_t c0 = 0; c0 < KP; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_lane + c0, r);
                tmem_ld_wait();
                if (lane_valid) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint32_t v0 = c0 + 8 * j;
                        if (v0 < (uint32_t)a.Ms) {
                            uint32_t h4[4], l4[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                split2(__uint_as_float(r[8 * j + 2 * q]), __uint_as_float(r[8 * j + 2 * q + 1]), h4[q], l4[q]);
                            uint32_t off = detail::mnmajor_off(my_g * a.Ms + v0, my_r, A2_LBO);
                            *reinterpret_cast<uint4*>(a2_hi + off) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
                            *reinterpret_cast<uint4*>(a2_lo + off) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
                        }
                    }
                }
            }
            tc_fence_before_sync();
            fence_async_smem();
            __syncthreads();

            // ---- stage 2 MMA: D2 = A2 * B^T, contraction over the map's rows
            if (tid == 0) {
                tc_fence_after_sync();
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint64_t da = pass == 1 ? d_a2_lo : d_a2_hi;
                    const uint64_t db = pass == 2 ? d_b_lo : d_b_hi;
#pragma unroll
                    for (uint32_t ks = 0; ks < KP / 16; ++ks) {
                        uint32_t b_off = (ks >> 2) * (KP * 128u) + (ks & 3) * 32u;
                        mma_bf16_ss(tmem + KP, desc_advance(da, ks * 2048u), desc_advance(db, b_off), IDESC2, acc);
                        acc = 1;
                    }
                }
                mma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after_sync();
        }

        // ---- final epilogue: coefficients -> energy, never leaving the SM
        {
            const uint32_t dcol = TWO_STAGE ? KP : 0;
            const int m = map0 + (int)my_g;
#pragma unroll 1
            for (uint32_t c0 = 0; c0 < KP; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_lane + dcol + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float z = __uint_as_float(r[i]);
                    energy = fmaf(z, z, energy);
                }
                if (a.dump != nullptr && lane_valid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        uint32_t col = c0 + i;
                        if (TWO_STAGE) {       // lane = (g, v), column = u  ->  Z[u][v]
                            if (col < (uint32_t)a.N) a.dump[(long long)m * a.NN + col * a.N + my_r] = __uint_as_float(r[i]);
                        } else {               // lane = map, column = u*N + v
                            if (col < (uint32_t)a.NN) a.dump[(long long)m * a.NN + col] = __uint_as_float(r[i]);
                        }
                    }
                }
            }
            tc_fence_before_sync();
            if constexpr (TWO_STAGE) {
                red[tid] = lane_valid ? energy : 0.f;
                __syncthreads();
                for (int g = warp; g < maps_here; g += 4) {
                    float s = 0.f;
                    for (int i = lane; i < a.N; i += 32) s += red[g * a.Ms + i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == 0) {
                        int mm = map0 + g;
                        int c = mm % a.c_count;
                        atomicAdd(a.accum + c, (double)s);
                        if (a.energy_out) a.energy_out[mm] = s;
                    }
                }
                __syncthreads();
            } else {
                if (lane_valid) {
                    int c = m % a.c_count;
                    atomicAdd(a.accum + c, (double)energy);
                    if (a.energy_out) a.energy_out[m] = energy;
                }
                __syncthreads();
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<S::TMEM_COLS>(tmem);
}

}  // namespace dctp
