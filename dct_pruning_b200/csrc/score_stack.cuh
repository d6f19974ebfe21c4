// Fused DCT-score hook kernel, "stacked basis" formulation (dense tensors, even map side 10..64).
//
// Same contract as score_umma.cuh / score_tmem.cuh (/root/reference/utils/common.py:262-277: per-(image, channel)
// orthonormal 2-D DCT-II energy, summed per channel).  What changes is how the first contraction is laid on the
// tensor core.  score_tmem.cuh gets its transposed intermediate from a block-diagonal TMEM operand I_G (x) C_N and
// pays a factor G of wasted MMA work plus three passes; here the TMEM operand is the basis itself with its bf16
// hi and lo parts STACKED along M (the 128 accumulator lanes), and the maps of a tile are concatenated along the
// MMA's N dimension:
//
//   lanes     lane = 32q + 16p + r :  p = 0 rows of C_hi, p = 1 rows of C_lo, basis row v = v(q, r)
//             J = 64 / KP map "sets" share the 128 lanes (KP = N rounded up to 16):
//               KP 48/64: J = 1, v = 16q + r     KP 32: J = 2, set = q / 2, v = 16 (q % 2) + r     KP 16: J = 4, set = q, v = r
//   stage 1   D1[(q,p,r), (g,h)] = sum_{(set',w)} A[(q,p,r), (set',w)] * Bx[(g,h), (set',w)]         2 passes: Bx_hi, Bx_lo
//             A  = the stacked basis, block structured over the sets, resident in TMEM for the whole kernel
//             Bx[(g,h), (set,w)] = X_{g,set}[h,w]: the raw rows of the maps, K-major in shared memory (N-side operand)
//             -> lanes p = 0 hold C_hi (X_hi + X_lo)^T, lanes p = 1 hold C_lo (X_hi + X_lo)^T: all four split-precision
//                terms in two passes, no block-diagonal waste (the maps sit side by side along N, not along K)
//   epi   1   tcgen05.ld 16x256b fragments of both lane halves -> y = za + zb (fp32) -> bf16 hi/lo pairs ->
//             tcgen05.st 16x128b: A2[(q,s,r), h] for map g = 2t + s of the warp's set; never leaves TMEM
//   stage 2   per A2 tile t:  D2[(q,s,r), u] = sum_h A2[(q,s,r), h] * C[u,h]       3 passes (hi*hi + lo*hi + hi*lo)
//             B = C_N resident in shared memory; every lane of a tile is a live row (two maps per warp)
//   epi   2   tcgen05.ld 32x32b -> sum of squares per lane -> 16-lane shuffle tree -> fixed-order sum over the
//             warps of a map -> one fp64 atomicAdd per map
//
// Tensor-core time per 25 KB of input: 448 + 384 cycles at 56x56 (score_tmem.cuh: 672 + 384), 512 + 192 at 28x28,
// 512 + 96 at 14x14, against 1113 cycles of HBM time at the measured 6.55 TB/s.
//
// The kernel is warp specialised (one CTA per SM, 18 warps), every hand-over is an mbarrier:
//   warp 16     TMA producer: one cp.async.bulk.tensor.2d per tile (tensor map over the dense fp32 stream viewed as
//               [rows, 32 floats]; a tile is a box of tile_rows rows; rows past the end arrive as zeros) into a ring of 3
//   warps 0-3   converters: fp32 tile -> bf16 hi/lo -> Bx (two buffers), offsets from a host-built table
//   warp 17     MMA issuer (one thread): stage 1 of tile i, then stage 2 of tile i-1
//   warps 4-11  epilogue 1 (warp = lane quarter q x row slot s)
//   warps 12-15 epilogue 2
// TMEM (512 columns): A 32 | D1 2 x 128 | A2 2 x 64 | D2 NB2 x 64 (NB2 = 1: 480 columns in use).
#pragma once
#include <cuda.h>
#include "score_umma.cuh"

namespace dctp {

struct StackArgs {
    const float* x_dense;           // first scored element; all scored maps back to back, 16-B aligned
    long long total_elems;          // n_maps * NN
    int n_maps, c_count;
    int N, NN, Np;                  // map side, N*N, rows a map takes in Bx (= D1 columns per map): N rounded up to 8
    int G, MT;                      // maps per set and per tile (MT = G * J)
    int ncols;                      // G * Np = MMA N of stage 1
    int tile_elems, tile_rows, tile_vec, num_tiles;
    int tail_tile;                  // tile converted straight from global memory (the stream does not end on a 128-byte row), or -1
    uint32_t idesc1, idesc2;
    const uint32_t* a_img;          // [128][32] packed bf16 pairs: the stacked basis as it sits in TMEM
    const uint8_t* c2_hi;           // stage-2 basis as a shared-memory operand image (hi, lo): C2_HALF bytes each
    const uint8_t* c2_lo;
    const uint16_t* table;          // [tile_vec] (VEC 4) / [2 * tile_vec] (VEC 2) byte offsets of a float4's pieces in Bx
    uint32_t table_bytes;
    double* accum;
    float* energy_out;
    float* dump;
    int* status;
};

struct StackSmem {
    static constexpr uint32_t NSTG = 3, STG_STRIDE = 32768;              // TMA ring: fp32 tiles of at most 32 KB
    static constexpr uint32_t LBO1 = 128 * 16 + 16;                      // Bx: k-chunk c of row n at c * LBO1 + n * 16 (+16: bank spread)
    static constexpr uint32_t BX_HALF = 8 * LBO1;                        // 16512
    static constexpr uint32_t LBO2 = 64 * 16 + 16;                       // stage-2 basis, same layout
    static constexpr uint32_t C2_HALF = 8 * LBO2;                        // 8320
    static constexpr uint32_t OFF_BX = NSTG * STG_STRIDE;
    static constexpr uint32_t OFF_C2 = OFF_BX + 4 * BX_HALF;
    static constexpr uint32_t OFF_TABLE = OFF_C2 + 2 * C2_HALF;
    static constexpr uint32_t TABLE_MAX = 8192;
    static constexpr uint32_t OFF_RED = OFF_TABLE + TABLE_MAX;
    static constexpr uint32_t OFF_BARS = OFF_RED + 256;
    static constexpr uint32_t OFF_SLOT = OFF_BARS + 256;
    static constexpr uint32_t TOTAL = OFF_SLOT + 128;
};

constexpr int STACK_NT = 576;

// KP: contraction length per map (N rounded up to 16); VEC: granularity of a row in the fp32 stream (4: N % 4 == 0, 2: N even)
template <int KP, int VEC>
__global__ void __launch_bounds__(STACK_NT, 1) score_stack_kernel(const __grid_constant__ CUtensorMap tmap, const StackArgs a) {
    using S = StackSmem;
    using namespace umma;
    constexpr int J = KP <= 32 ? 64 / KP : 1;
    constexpr int K1S = J * KP / 16, K2S = KP / 16;
    constexpr int G = KP == 16 ? 8 : KP == 32 ? 4 : 2, T2 = G / 2;
    constexpr uint32_t TM_A = 0, TM_D1 = 32, TM_A2 = 288, TM_D2 = 416;   // TMEM columns (D1, A2 double buffered)
    constexpr uint32_t STEP1 = (2 * S::LBO1) >> 4, STEP2 = (2 * S::LBO2) >> 4;   // one k-step (two 8-element chunks) in descriptor units

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* stg = smem;
    uint8_t* bx = smem + S::OFF_BX;                                       // [buffer][hi | lo]
    uint8_t* c2 = smem + S::OFF_C2;
    const uint8_t* tab = smem + S::OFF_TABLE;
    float* red = reinterpret_cast<float*>(smem + S::OFF_RED);             // [parity][q][8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BARS);
    uint64_t* stg_full = bars;                                            // TMA landed (tx bytes)
    uint64_t* stg_free = bars + 3;                                        // 4 converter warps
    uint64_t* bx_full = bars + 6;                                         // 4 converter warps
    uint64_t* bx_free = bars + 8;                                         // stage-1 MMAs of the buffer have completed
    uint64_t* d1_full = bars + 10;                                        // ... and D1 is complete
    uint64_t* d1_free = bars + 12;                                        // 8 epilogue-1 warps have read it
    uint64_t* a2_full = bars + 14;                                        // 8 epilogue-1 warps have written A2
    uint64_t* a2_free = bars + 16;                                        // stage-2 MMAs have read it
    uint64_t* d2_full = bars + 18;
    uint64_t* d2_free = bars + 19;                                        // 4 epilogue-2 warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::OFF_SLOT);

    // ---- prologue (independent of the activation: overlaps the preceding kernel under a dependent launch)
    for (uint32_t off = tid * 16; off < 4 * S::BX_HALF; off += STACK_NT * 16) *reinterpret_cast<uint4*>(bx + off) = make_uint4(0, 0, 0, 0);
    for (uint32_t off = tid * 16; off < S::C2_HALF; off += STACK_NT * 16) {
        *reinterpret_cast<uint4*>(c2 + off) = *reinterpret_cast<const uint4*>(a.c2_hi + off);
        *reinterpret_cast<uint4*>(c2 + S::C2_HALF + off) = *reinterpret_cast<const uint4*>(a.c2_lo + off);
    }
    for (uint32_t off = tid * 16; off < a.table_bytes; off += STACK_NT * 16)
        *reinterpret_cast<uint4*>(smem + S::OFF_TABLE + off) = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(a.table) + off);
    if (warp == 17) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        for (int s = 0; s < 3; ++s) { mbar_init(stg_full + s, 1); mbar_init(stg_free + s, 4); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bx_full + b, 4); mbar_init(bx_free + b, 1);
            mbar_init(d1_full + b, 1); mbar_init(d1_free + b, 8);
            mbar_init(a2_full + b, 8); mbar_init(a2_free + b, 1);
        }
        mbar_init(d2_full, 1); mbar_init(d2_free, 4);
        mbar_init_fence();
    }
    if (warp == 16 && lane == 0) tma_prefetch_desc(&tmap);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    if (warp < 4) {                                                       // the stacked basis -> TMEM columns [0,32)
        uint32_t v[16];
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const uint4* src = reinterpret_cast<const uint4*>(a.a_img + tid * 32 + part * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 q4 = src[i];
                v[4 * i] = q4.x; v[4 * i + 1] = q4.y; v[4 * i + 2] = q4.z; v[4 * i + 3] = q4.w;
            }
            tmem_st16(tmem + ((warp * 32u) << 16) + TM_A + part * 16, v);
        }
        tmem_st_wait();
    } else if (warp < 8) {                                                // A2 (both buffers): columns a tile never writes stay zero
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
        for (int part = 0; part < 8; ++part) tmem_st16(tmem + (((warp & 3u) * 32u) << 16) + TM_A2 + part * 16, z);
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    launch_dependents();                                                  // this CTA holds its TMEM columns (see score_umma.cuh)
    grid_dependency_wait();                                               // the activation is complete

    const int first = blockIdx.x, stride = gridDim.x;
    const uint32_t tile_bytes = static_cast<uint32_t>(a.tile_rows) * 128u;
    bool dead = false;
#define STACK_WAIT(bar, par)                         \
    if (!mbar_wait((bar), (par))) {                  \
        dead = true;                                 \
        break;                                       \
    }

    if (warp == 16) {
        // ================================================================ TMA producer
        if (elect_one()) {
            uint32_t it = 0;
            for (int tile = first; tile < a.num_tiles; tile += stride) {
                if (tile == a.tail_tile) continue;
                const uint32_t s = it % S::NSTG;
                if (it >= S::NSTG) STACK_WAIT(stg_free + s, ((it / S::NSTG) - 1u) & 1u);
                mbar_arrive_expect_tx(stg_full + s, tile_bytes);
                tma_load_2d(stg + s * S::STG_STRIDE, &tmap, 0, tile * a.tile_rows, stg_full + s);
                ++it;
            }
        }
        __syncwarp();
    } else if (warp == 17) {
        // ================================================================ MMA issuer
        if (elect_one()) {
            const uint64_t desc = make_smem_desc(0, S::LBO1, 128, SWIZZLE_NONE);
            const uint64_t desc2 = make_smem_desc(0, S::LBO2, 128, SWIZZLE_NONE);
            const uint32_t lo_c2_hi = static_cast<uint32_t>(desc2) + (smem_u32(c2) >> 4);
            const uint32_t lo_c2_lo = lo_c2_hi + (S::C2_HALF >> 4);
            auto stage2 = [&](uint32_t m) -> bool {                       // tile number m of this CTA
                const uint32_t b = m & 1u;
                if (!mbar_wait(a2_full + b, (m >> 1) & 1u)) return false;
                if (m >= 1 && !mbar_wait(d2_free, (m - 1u) & 1u)) return false;
                tc_fence_after_sync();
#pragma unroll
                for (int t = 0; t < T2; ++t) {
                    const uint32_t d = tmem + TM_D2 + t * KP;
                    const uint32_t ahi = tmem + TM_A2 + b * 64 + t * KP, alo = ahi + KP / 2;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                        for (int ks = 0; ks < K2S; ++ks)
                            mma_bf16_ts(d, (pass == 1 ? alo : ahi) + 8 * ks,
                                        desc_with_lo(desc2, (pass == 2 ? lo_c2_lo : lo_c2_hi) + ks * STEP2), a.idesc2, (pass | ks) != 0);
                }
                mma_commit(d2_full);
                mma_commit(a2_free + b);
                return true;
            };
            uint32_t n = 0;
            for (int tile = first; tile < a.num_tiles; tile += stride, ++n) {
                const uint32_t b = n & 1u;
                STACK_WAIT(bx_full + b, (n >> 1) & 1u);
                if (n >= 2) STACK_WAIT(d1_free + b, ((n >> 1) - 1u) & 1u);
                tc_fence_after_sync();
                const uint32_t d1 = tmem + TM_D1 + b * 128;
                const uint32_t lo_hi = static_cast<uint32_t>(desc) + (smem_u32(bx + b * 2 * S::BX_HALF) >> 4);
                const uint32_t lo_lo = lo_hi + (S::BX_HALF >> 4);
#pragma unroll
                for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                    for (int ks = 0; ks < K1S; ++ks)
                        mma_bf16_ts(d1, tmem + TM_A + 8 * ks, desc_with_lo(desc, (pass ? lo_lo : lo_hi) + ks * STEP1), a.idesc1,
                                    (pass | ks) != 0);
                mma_commit(d1_full + b);
                mma_commit(bx_free + b);
                if (n >= 1 && !stage2(n - 1)) { dead = true; break; }
            }
            if (!dead && n >= 1 && !stage2(n - 1)) dead = true;
        }
        __syncwarp();
    } else if (warp < 4) {
        // ================================================================ converters: fp32 tile -> bf16 hi/lo -> Bx
        uint32_t it = 0, n = 0;
        for (int tile = first; tile < a.num_tiles; tile += stride, ++n) {
            const uint32_t b = n & 1u;
            if (n >= 2) STACK_WAIT(bx_free + b, ((n >> 1) - 1u) & 1u);
            uint8_t* hi = bx + b * 2 * S::BX_HALF;
            uint8_t* lo = hi + S::BX_HALF;
            auto emit = [&](int f, const float4 v) {
                uint32_t h0, l0, h1, l1;
                split2(v.x, v.y, h0, l0);
                split2(v.z, v.w, h1, l1);
                if constexpr (VEC == 4) {
                    const uint32_t off = reinterpret_cast<const uint16_t*>(tab)[f];
                    *reinterpret_cast<uint2*>(hi + off) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(lo + off) = make_uint2(l0, l1);
                } else {
                    const uint32_t o2 = reinterpret_cast<const uint32_t*>(tab)[f];
                    const uint32_t o0 = o2 & 0xFFFFu, o1 = o2 >> 16;
                    *reinterpret_cast<uint32_t*>(hi + o0) = h0;
                    *reinterpret_cast<uint32_t*>(lo + o0) = l0;
                    *reinterpret_cast<uint32_t*>(hi + o1) = h1;
                    *reinterpret_cast<uint32_t*>(lo + o1) = l1;
                }
            };
            if (tile != a.tail_tile) {
                const uint32_t s = it % S::NSTG;
                STACK_WAIT(stg_full + s, (it / S::NSTG) & 1u);
                const float4* src = reinterpret_cast<const float4*>(stg + s * S::STG_STRIDE);
#pragma unroll 4
                for (int f = tid; f < a.tile_vec; f += 128) emit(f, src[f]);
                __syncwarp();
                if (lane == 0) mbar_arrive(stg_free + s);
                ++it;
            } else {                                                      // the stream's last, partial 128-byte row is not in the tensor map
                const long long e0 = static_cast<long long>(tile) * a.tile_elems;
                for (int f = tid; f < a.tile_vec; f += 128) {
                    const long long e = e0 + 4ll * f;
                    float4 v;
                    v.x = e + 0 < a.total_elems ? a.x_dense[e + 0] : 0.f;
                    v.y = e + 1 < a.total_elems ? a.x_dense[e + 1] : 0.f;
                    v.z = e + 2 < a.total_elems ? a.x_dense[e + 2] : 0.f;
                    v.w = e + 3 < a.total_elems ? a.x_dense[e + 3] : 0.f;
                    emit(f, v);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bx_full + b);
        }
    } else if (warp < 12) {
        // ================================================================ epilogue 1: D1 -> y = za + zb -> bf16 hi/lo -> A2
        const uint32_t q = warp & 3u, s = (warp - 4u) >> 2;
        const uint32_t lane_q = (q * 32u) << 16, lane_s = (q * 32u + s * 16u) << 16;
        const int np8 = a.Np >> 3;
        uint32_t n = 0;
        for (int tile = first; tile < a.num_tiles; tile += stride, ++n) {
            const uint32_t b = n & 1u;
            STACK_WAIT(d1_full + b, (n >> 1) & 1u);
            if (n >= 2) STACK_WAIT(a2_free + b, ((n >> 1) - 1u) & 1u);
            tc_fence_after_sync();
            const uint32_t d1 = tmem + TM_D1 + b * 128;
#pragma unroll
            for (int t = 0; t < T2; ++t) {
                const uint32_t col0 = (2 * t + s) * a.Np;                 // this warp's map of A2 tile t
                const uint32_t src_a = d1 + lane_q + col0, src_b = src_a + (16u << 16);
                const uint32_t dst_hi = tmem + TM_A2 + b * 64 + t * KP + lane_s, dst_lo = dst_hi + KP / 2;
                int c8 = 0;
                for (; c8 + 2 <= np8; c8 += 2) {                          // 16 columns of h = 8 packed A2 columns
                    uint32_t za[8], zb[8], h[4], l[4];
                    tmem_ld_frag16(src_a + c8 * 8, za);
                    tmem_ld_frag16(src_b + c8 * 8, zb);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        split2_packed(add2_packed(pack2(za[2 * i], za[2 * i + 1]), pack2(zb[2 * i], zb[2 * i + 1])), h[i], l[i]);
                    tmem_st_frag8(dst_hi + c8 * 4, h);
                    tmem_st_frag8(dst_lo + c8 * 4, l);
                }
                if (c8 < np8) {                                           // 8 more columns
                    uint32_t za[4], zb[4], h[2], l[2];
                    tmem_ld_frag8(src_a + c8 * 8, za);
                    tmem_ld_frag8(src_b + c8 * 8, zb);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        split2_packed(add2_packed(pack2(za[2 * i], za[2 * i + 1]), pack2(zb[2 * i], zb[2 * i + 1])), h[i], l[i]);
                    tmem_st_frag4(dst_hi + c8 * 4, h[0], h[1]);
                    tmem_st_frag4(dst_lo + c8 * 4, l[0], l[1]);
                }
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(d1_free + b);
                mbar_arrive(a2_full + b);
            }
        }
    } else if (warp < 16) {
        // ================================================================ epilogue 2: D2 -> energies
        const uint32_t q = warp & 3u, et = tid - 12 * 32;                 // thread within the role
        const uint32_t lane_q = (q * 32u) << 16;
        const uint32_t s = lane >> 4, r = lane & 15u;
        const uint32_t my_set = J == 4 ? q : J == 2 ? (q >> 1) : 0u;
        const uint32_t my_v = J == 4 ? r : J == 2 ? 16u * (q & 1u) + r : 16u * q + r;
        uint32_t n = 0;
        for (int tile = first; tile < a.num_tiles; tile += stride, ++n) {
            STACK_WAIT(d2_full, n & 1u);
            tc_fence_after_sync();
            float* red_w = red + (n & 1u) * 32;
#pragma unroll
            for (int t = 0; t < T2; ++t) {
                float e0 = 0.f, e1 = 0.f;
#pragma unroll
                for (int c = 0; c < KP; c += 16) {
                    uint32_t z[16];
                    tmem_ld16(tmem + TM_D2 + lane_q + t * KP + c, z);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        e0 = fmaf(__uint_as_float(z[i]), __uint_as_float(z[i]), e0);
                        e1 = fmaf(__uint_as_float(z[8 + i]), __uint_as_float(z[8 + i]), e1);
                    }
                    if (a.dump != nullptr) {
                        const long long m = static_cast<long long>(tile) * a.MT + (2 * t + s) * J + my_set;
                        if (m < a.n_maps && my_v < (uint32_t)a.N)
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c + i < a.N) a.dump[m * a.NN + (c + i) * a.N + my_v] = __uint_as_float(z[i]);
                    }
                }
                float e = e0 + e1;
                e += __shfl_xor_sync(0xffffffffu, e, 8);
                e += __shfl_xor_sync(0xffffffffu, e, 4);
                e += __shfl_xor_sync(0xffffffffu, e, 2);
                e += __shfl_xor_sync(0xffffffffu, e, 1);
                if (r == 0) red_w[q * 8 + 2 * t + s] = e;
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_free);
            named_bar_sync(1, 128);
            if (et < (uint32_t)a.MT) {
                const uint32_t g = et / J, set = et % J;
                float e;
                if (J == 4) e = red_w[set * 8 + g];
                else if (J == 2) e = red_w[(2 * set) * 8 + g] + red_w[(2 * set + 1) * 8 + g];
                else e = (red_w[g] + red_w[8 + g]) + (red_w[16 + g] + red_w[24 + g]);
                const long long m = static_cast<long long>(tile) * a.MT + et;
                if (m < a.n_maps) {
                    atomicAdd(a.accum + (m % a.c_count), static_cast<double>(e));
                    if (a.energy_out) a.energy_out[m] = e;
                }
            }
        }
    }
#undef STACK_WAIT
    if (dead) {                                                           // a hand-over never came: flag it and poison the result
        atomicExch(a.status, DCTP_DEV_MMA_TIMEOUT);
        for (int c = lane; c < a.c_count; c += 32) a.accum[c] = __longlong_as_double(0x7FF8000000000000ll);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 17) tmem_dealloc<512>(tmem);
}

}  // namespace dctp
