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
// The kernel is tensor-bound by design (2*N^3 flops per stage and map); the map is re-read once per v-chunk out
// of L2 and the basis slabs come from L2.  Warp-specialised pipeline, operand slabs double buffered in shared memory:
//   basis producer (one elected thread) basis slabs by 1-D bulk copy into a ring of three buffers (the basis image in
//                                       global memory is already in the swizzled operand layout: a slab is one blob)
//   MMA issuer (one elected thread)     issues the MMAs of a step once its slabs are full and does nothing else: the
//                                       tensor pipe's queue is shallow, every cycle this thread spends elsewhere is idle
//                                       tensor time; tcgen05.commit frees the buffers
//   16 converter / epilogue warps       map slabs fp32 -> bf16 hi/lo from registers loaded two steps earlier; the two
//                                       epilogues.  No block-wide barrier inside the pipeline, only mbarriers.
// One resident CTA of 18 warps per SM (226 KB of shared memory, 512 TMEM columns).
#pragma once
#include "score_umma.cuh"

namespace dctp {

struct LargeScoreArgs {
    const float* x_dense;           // first scored element; all scored maps back to back (stride_h == N), 16-B aligned
    int n_maps, c_count;
    int N, NPR;                     // map side; rows per column block of the basis image (>= N and >= NU * NUC, zero padded)
    int NVC;                        // v-chunks per map = ceil(N / 128)
    int NU, NUC;                    // u-chunk width (multiple of 16, <= 128) and count: NU * NUC >= N
    int n_items;                    // n_maps * NVC
    const uint8_t* c_hi;            // basis image, bf16: [column blocks of 64][NPR rows][128 B], 16-B chunks XOR-swizzled by
    const uint8_t* c_lo;            //   (row & 7): any (8-aligned row range, column block) is one contiguous operand slab
    double* accum;
    float* energy_out;              // optional [n_maps], accumulated with float atomics over the v-chunks (caller zeroes)
    float* dump;                    // optional [n_maps][N][N] coefficients Z[u][v]
    int* status;
    long long* trace;               // bring-up aid: cycles the control thread of CTA 0 spent in each kind of wait
    int exp_flags;                  // bring-up aid (DCTP_L_EXP): 1 = skip the map slab stores, 2 = skip the basis bulk copies
};

struct LargeSmem {
    static constexpr uint32_t A1_HALF = 128 * 128;                 // X slab: 128 rows x 64 k (hi or lo)
    static constexpr uint32_t A1_BUF = 2 * A1_HALF;                // hi | lo
    static constexpr uint32_t B_HALF = 128 * 128;                  // basis slab: up to 128 rows x 64 k
    static constexpr uint32_t B_BUF = 2 * B_HALF;                  // hi | lo
    static constexpr uint32_t NXB = 2, NBB = 3;                    // X slab buffers (converters are one step ahead), basis slab
                                                                   // buffers (a bulk copy takes about one step: two in flight)
    static constexpr uint32_t A2_HALF = 128 * 128 * 2;             // A2: 128 k-rows x 128 m, MN-major
    static constexpr uint32_t OFF_A1 = 0;
    static constexpr uint32_t OFF_B = NXB * A1_BUF;
    static constexpr uint32_t OFF_A2_HI = OFF_B + NBB * B_BUF, OFF_A2_LO = OFF_A2_HI + A2_HALF;
    static constexpr uint32_t OFF_CTRL = OFF_A2_LO + A2_HALF;      // 13 mbarriers, TMEM slot
    static constexpr uint32_t OFF_RED = OFF_CTRL + 128;
    static constexpr uint32_t TOTAL = OFF_RED + 512 * 4;
    static_assert(OFF_B % 1024 == 0 && B_BUF % 1024 == 0 && OFF_A2_HI % 1024 == 0, "swizzle atoms are 1024-byte aligned");
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

constexpr int LARGE_CONV = 512;        // 16 converter / epilogue warps: warp w owns TMEM lane quarter w % 4 and every 4th
                                       // 16-column block (w / 4) of an epilogue
constexpr int LARGE_NT = LARGE_CONV + 64;   // + the MMA issuer warp and the basis producer warp (one elected thread each)

namespace detail {
// The steps of the whole kernel in issue order: per item (map, v-chunk), per h-tile: one stage-1 step per 64-wide
// w block, then one stage-2 step per (u-chunk, 64-wide h block).  The issuer and the basis producer each walk one.
struct LargeStep {
    int item, v0, MV16;             // work item
    int h0, ht, MH;                 // h-tile
    int w0;                         // stage 1: w block
    int uc, kb;                     // stage 2: u-chunk, h block inside the tile
    bool stage2;
    __device__ __forceinline__ void set_item(const LargeScoreArgs& a, int it) {
        item = it;
        const int map = it / a.NVC;
        v0 = (it - map * a.NVC) * 128;
        MV16 = (min(128, a.N - v0) + 15) & ~15;
        h0 = 0; ht = 0; MH = min(128, a.N);
        w0 = 0; uc = 0; kb = 0; stage2 = false;
    }
    __device__ __forceinline__ bool valid(const LargeScoreArgs& a) const { return item < a.n_items; }
    __device__ __forceinline__ int nkb() const { return (MH + 63) >> 6; }
    __device__ __forceinline__ bool last_stage1(const LargeScoreArgs& a) const { return !stage2 && w0 + 64 >= a.N; }
    __device__ __forceinline__ bool first_stage2() const { return stage2 && uc == 0 && kb == 0; }
    __device__ __forceinline__ bool last_of_item(const LargeScoreArgs& a) const {
        return stage2 && h0 + 128 >= a.N && uc + 1 == a.NUC && kb + 1 == nkb();
    }
    __device__ __forceinline__ int ksteps(const LargeScoreArgs& a) const {
        return (stage2 ? min(64, MH - kb * 64) : min(64, a.N - w0)) >> 4;
    }
    // basis slab of this step: byte offset into the image and size (per hi / lo half)
    __device__ __forceinline__ uint32_t slab_off(const LargeScoreArgs& a) const {
        const int cb = stage2 ? ((h0 >> 6) + kb) : (w0 >> 6), row0 = stage2 ? uc * a.NU : v0;
        return (uint32_t)(cb * a.NPR + row0) * 128u;
    }
    __device__ __forceinline__ uint32_t slab_bytes(const LargeScoreArgs& a) const { return (uint32_t)(stage2 ? a.NU : MV16) * 128u; }
    __device__ __forceinline__ void advance(const LargeScoreArgs& a, int grid) {
        if (!stage2) {
            w0 += 64;
            if (w0 >= a.N) { stage2 = true; uc = 0; kb = 0; }
            return;
        }
        if (++kb < nkb()) return;
        kb = 0;
        if (++uc < a.NUC) return;
        h0 += 128; ++ht;
        if (h0 < a.N) { MH = min(128, a.N - h0); stage2 = false; w0 = 0; uc = 0; return; }
        set_item(a, item + grid);
    }
};
}  // namespace detail

__global__ void __launch_bounds__(LARGE_NT, 1) score_large_kernel(const LargeScoreArgs a) {
    constexpr int NC = LARGE_CONV;
    constexpr int XV = 2048 / NC;                                  // prefetch registers per map slab: float4 per thread
    using S = LargeSmem;
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a1_base = smem + S::OFF_A1;                           // buffer p at + p * A1_BUF: hi, then lo
    uint8_t* b_base = smem + S::OFF_B;                             // buffer p at + p * B_BUF: hi, then lo
    uint8_t* a2_hi = smem + S::OFF_A2_HI;
    uint8_t* a2_lo = smem + S::OFF_A2_LO;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_CTRL);
    uint64_t* x_full = bars;            // [2] map slab stored (16 warp arrivals)
    uint64_t* x_empty = bars + 2;       // [2] MMAs that read it are done (tcgen05.commit)
    uint64_t* b_full = bars + 4;        // [3] basis slab landed (bulk-copy bytes)
    uint64_t* b_empty = bars + 7;       // [3] MMAs that read it are done
    uint64_t* acc_ready = bars + 10;    // D1 of an h-tile complete / D2 of an item complete
    uint64_t* a2_full = bars + 11;      // epilogue 1 stored A2 (16 warp arrivals)
    uint64_t* d2_free = bars + 12;      // epilogue 2 has read D2 (16 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::OFF_CTRL + 112);
    float* red = reinterpret_cast<float*>(smem + S::OFF_RED);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (tid == 0) atomicExch(a.status, DCTP_DEV_SMEM_ALIGN);
        return;
    }
    launch_dependents();
    for (uint32_t off = tid * 16; off < S::OFF_CTRL; off += LARGE_NT * 16) *reinterpret_cast<uint4*>(smem + off) = make_uint4(0, 0, 0, 0);
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        for (uint32_t i = 0; i < S::NXB; ++i) {
            mbar_init(x_full + i, NC / 32);
            mbar_init(x_empty + i, 1);
        }
        for (uint32_t i = 0; i < S::NBB; ++i) {
            mbar_init(b_full + i, 1);
            mbar_init(b_empty + i, 1);
        }
        mbar_init(acc_ready, 1);
        mbar_init(a2_full, NC / 32);
        mbar_init(d2_free, NC / 32);
        mbar_init_fence();
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    grid_dependency_wait();                                        // the activation (written by the preceding kernel) is complete
    const uint32_t d1_col = 0, d2_col = 128;
    const int N = a.N, NN = N * N;

    // a timed-out wait (pipeline bug) is recorded once; afterwards this thread stops waiting so the kernel still ends
    bool dead = false;
    auto wait = [&](uint64_t* bar, uint32_t parity) {
        if (!dead && !mbar_wait(bar, parity)) {
            dead = true;
            atomicExch(a.status, DCTP_DEV_MMA_TIMEOUT);
        }
    };

    if (warp == NC / 32) {
        // =========================================================== MMA issuer: one elected thread
        if (elect_one()) {
            const uint64_t desc_k = make_smem_desc(0, 16, 1024, SWIZZLE_128B);
            const uint64_t desc_mn = make_smem_desc(0, 16384, 1024, SWIZZLE_128B);    // A2: 64-wide M blocks 16 KB apart
            const uint32_t k_lo = static_cast<uint32_t>(desc_k), mn_lo = static_cast<uint32_t>(desc_mn);
            const uint32_t lo_a1 = smem_u32(a1_base) >> 4, lo_b = smem_u32(b_base) >> 4;
            const uint32_t lo_a2_hi = smem_u32(a2_hi) >> 4, lo_a2_lo = smem_u32(a2_lo) >> 4;
            const uint32_t idesc2 = make_idesc_bf16(128, a.NU, true, false);
            const int grid = (int)gridDim.x;
            detail::LargeStep cur;
            cur.set_item(a, blockIdx.x);
            uint32_t s = 0, bp = 0, bpar = 0;                       // steps issued; basis buffer of this step and its parity
            uint32_t xs = 0, tiles = 0, items = 0;                  // stage-1 steps, h-tiles, items issued so far
            long long tr[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            const bool tracing = a.trace != nullptr && blockIdx.x == 0;
            auto twait = [&](int slot, uint64_t* bar, uint32_t parity) {
                if (tracing) {
                    const long long t0 = clock64();
                    wait(bar, parity);
                    tr[slot] += clock64() - t0;
                } else {
                    wait(bar, parity);
                }
            };
            const long long t_begin = clock64();
            bool have = false;                                      // the operands of `cur` have been waited for already
            while (cur.valid(a)) {
                const uint32_t xp = xs & 1u;
                if (!have) {
                    if (!cur.stage2) twait(1, x_full + xp, (xs >> 1) & 1u);
                    twait(4, b_full + bp, bpar);
                }
                if (cur.first_stage2()) {
                    twait(2, a2_full, tiles & 1u);
                    ++tiles;
                    if (cur.ht == 0 && items >= 1) twait(3, d2_free, (items - 1) & 1u);   // the first MMA of an item overwrites D2
                }
                tc_fence_after_sync();
                const uint32_t bh = k_lo + lo_b + bp * (S::B_BUF >> 4), bl = bh + (S::B_HALF >> 4);
                const int ks = cur.ksteps(a);
                const bool st2 = cur.stage2;
                uint32_t ah, al, d, idesc;
                bool acc0;
                if (!st2) {
                    ah = k_lo + lo_a1 + xp * (S::A1_BUF >> 4); al = ah + (S::A1_HALF >> 4);
                    d = tmem + d1_col; idesc = make_idesc_bf16(128, cur.MV16, false, false); acc0 = cur.w0 != 0;
                } else {
                    ah = mn_lo + lo_a2_hi + cur.kb * 4 * 128; al = mn_lo + lo_a2_lo + cur.kb * 4 * 128;
                    d = tmem + d2_col + cur.uc * a.NU; idesc = idesc2; acc0 = !(cur.ht == 0 && cur.kb == 0);
                }
                // hi*hi and lo*hi go out first; the bookkeeping and the look-ahead waits below then run while the tensor
                // pipe still has those queued
                if (!st2) {
                    detail::issue_ss_pass_n<false>(ks, d, ah, bh, desc_k, desc_k, idesc, acc0);
                    detail::issue_ss_pass_n<false>(ks, d, al, bh, desc_k, desc_k, idesc, true);
                } else {
                    detail::issue_ss_pass_n<true>(ks, d, ah, bh, desc_mn, desc_k, idesc, acc0);
                    detail::issue_ss_pass_n<true>(ks, d, al, bh, desc_mn, desc_k, idesc, true);
                }
                const bool commit_acc = cur.last_stage1(a) || cur.last_of_item(a);
                if (cur.last_of_item(a)) ++items;
                cur.advance(a, grid);
                uint32_t nbp = bp + 1, nbpar = bpar;
                if (nbp == S::NBB) { nbp = 0; nbpar ^= 1u; }
                const uint32_t nxs = xs + (st2 ? 0u : 1u);
                have = false;
                if (cur.valid(a)) {                                // next step's operands (never waits on this step's commits)
                    if (!cur.stage2) twait(1, x_full + (nxs & 1u), (nxs >> 1) & 1u);
                    twait(5, b_full + nbp, nbpar);
                    have = true;
                }
                if (!st2) detail::issue_ss_pass_n<false>(ks, d, ah, bl, desc_k, desc_k, idesc, true);
                else detail::issue_ss_pass_n<true>(ks, d, ah, bl, desc_mn, desc_k, idesc, true);
                if (!st2) mma_commit(x_empty + xp);
                mma_commit(b_empty + bp);
                if (commit_acc) mma_commit(acc_ready);
                xs = nxs; bp = nbp; bpar = nbpar;
                ++s;
            }
            if (tracing) {
                tr[6] = clock64() - t_begin;
                tr[7] = s;
                for (int i = 0; i < 12; ++i) a.trace[i] = tr[i];
            }
        }
        __syncwarp();
    } else if (warp == NC / 32 + 1) {
        // =========================================================== basis producer: one elected thread
        if (elect_one()) {
            const int grid = (int)gridDim.x;
            detail::LargeStep st;
            st.set_item(a, blockIdx.x);
            uint32_t bp = 0, round = 0;                             // buffer of this step, how often the ring wrapped
            while (st.valid(a)) {
                if (round >= 1) wait(b_empty + bp, (round - 1) & 1u);               // the MMAs that last read this buffer are done
                const uint32_t bytes = st.slab_bytes(a), off = st.slab_off(a);
                if (a.exp_flags & 2) {
                    mbar_arrive(b_full + bp);
                } else {
                    mbar_arrive_expect_tx(b_full + bp, 2 * bytes);
                    bulk_g2s(b_base + bp * S::B_BUF, a.c_hi + off, bytes, b_full + bp);
                    bulk_g2s(b_base + bp * S::B_BUF + S::B_HALF, a.c_lo + off, bytes, b_full + bp);
                }
                st.advance(a, grid);
                if (++bp == S::NBB) { bp = 0; ++round; }
            }
        }
        __syncwarp();
    } else {
        // =========================================================== converter / epilogue warps
        const uint32_t lane = tid & 31;
        const uint32_t lane_idx = (warp & 3) * 32 + lane;          // this thread's TMEM lane
        const uint32_t col_blk = warp >> 2;                        // which 16-column blocks of an epilogue it handles
        const uint32_t tmem_lane = tmem + (((warp & 3) * 32u) << 16);
        uint32_t xs = 0, acc_cnt = 0;                              // stage-1 steps stored, accumulator hand-overs consumed

        float4 xr0[XV], xr1[XV];                                   // two map slabs in flight (two steps ahead: HBM latency)
        uint32_t xm0 = 0, xm1 = 0;                                 // per set: float4 per row (low 8 bits) | vectors in the slab << 8
        auto load_x = [&](float4 (&xr)[XV], uint32_t& meta, const float* xm, int h0, int w0) {
            const int MH = min(128, N - h0), kvalid = min(64, N - w0);
            const uint32_t vpr = kvalid / 4;                       // float4 vectors per row (4 / 8 / 12 / 16)
            const uint32_t total = (uint32_t)MH * vpr;
            meta = vpr | (total << 8);
            if (vpr == 16) {                                       // full-width block: row = i >> 4, fixed per thread
                const float4* src = reinterpret_cast<const float4*>(xm + (size_t)(h0 + (tid >> 4)) * N + w0) + (tid & 15);
#pragma unroll
                for (int j = 0; j < XV; ++j)
                    if (tid + j * NC < total) xr[j] = detail::ldg_stream(src + (size_t)j * (NC / 16) * (N / 4));
            } else {
#pragma unroll
                for (int j = 0; j < XV; ++j) {
                    const uint32_t i = tid + j * NC;
                    if (i < total) {
                        const uint32_t r = i / vpr, q = i - r * vpr;
                        xr[j] = detail::ldg_stream(reinterpret_cast<const float4*>(xm + (size_t)(h0 + r) * N + w0) + q);
                    }
                }
            }
        };
        const uint32_t xoff16 = detail::kmajor_off(tid >> 4, (tid & 15) * 4, 128);    // full-width block: + j * 4096 B per 32 rows
        auto store_x = [&](const float4 (&xr)[XV], uint32_t meta, uint32_t p) {
            const uint32_t vpr = meta & 255u, total = meta >> 8;
            uint8_t* hi = a1_base + p * S::A1_BUF;
            uint8_t* lo = hi + S::A1_HALF;
            if (vpr == 16) {
#pragma unroll
                for (int j = 0; j < XV; ++j)
                    if (tid + j * NC < total) detail::Scatter<1>::st(hi + j * (NC / 16) * 128, lo + j * (NC / 16) * 128, static_cast<uint16_t>(xoff16), xr[j]);
            } else {
#pragma unroll
                for (int j = 0; j < XV; ++j) {
                    const uint32_t i = tid + j * NC;
                    if (i < total) {
                        const uint32_t r = i / vpr, q = i - r * vpr;
                        detail::Scatter<1>::st(hi, lo, static_cast<uint16_t>(detail::kmajor_off(r, q * 4, 128)), xr[j]);
                    }
                }
            }
        };
        // one stage-1 step of this thread: wait for the buffer, convert + store the slab, hand it over, refill the registers
        auto x_step = [&](float4 (&xr)[XV], uint32_t& meta, const float* xm, int h0, int w_next) {
            const uint32_t xp = xs & 1u, k = xs >> 1;
            if (k >= 1) wait(x_empty + xp, (k - 1) & 1u);
            if (!(a.exp_flags & 1)) store_x(xr, meta, xp);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(x_full + xp);
            if (w_next < N) load_x(xr, meta, xm, h0, w_next);
            ++xs;
        };

        int item = blockIdx.x;
        if (item < a.n_items) {                                    // operands of the very first steps
            const float* xm = a.x_dense + (size_t)(item / a.NVC) * NN;
            load_x(xr0, xm0, xm, 0, 0);
            load_x(xr1, xm1, xm, 0, 64);
        }
        bool pre_stored = false;                                   // the tile's first two slabs were stored ahead of time
        for (; item < a.n_items; item += gridDim.x) {
            const int map = item / a.NVC, vc = item - map * a.NVC;
            const int v0 = vc * 128, MV = min(128, N - v0), MV16 = (MV + 15) & ~15;
            const float* xm = a.x_dense + (size_t)map * NN;
            const int nitem = item + (int)gridDim.x;               // what this CTA works on next
            const float* nxm = a.x_dense + (size_t)(nitem / a.NVC) * NN;

            for (int h0 = 0; h0 < N; h0 += 128) {
                const int MH = min(128, N - h0);
                // ---- stage 1 operands: X slab rows h0.., one step per 64-wide w block, register sets alternate
                for (int w0 = pre_stored ? 128 : 0; w0 < N; w0 += 128) {
                    x_step(xr0, xm0, xm, h0, w0 + 128);
                    if (w0 + 64 < N) x_step(xr1, xm1, xm, h0, w0 + 192);
                }
                // the tile after this one (next h-tile, or the first of the next item): its first two slabs travel now
                const bool more = h0 + 128 < N || nitem < a.n_items;
                const float* txm = h0 + 128 < N ? xm : nxm;
                const int th0 = h0 + 128 < N ? h0 + 128 : 0;
                if (more) {
                    load_x(xr0, xm0, txm, th0, 0);
                    load_x(xr1, xm1, txm, th0, 64);
                }
                // ---- epilogue 1: D1 row h (lane) -> bf16 hi/lo -> A2[k = h][m = v]  (MN-major)
                wait(acc_ready, acc_cnt & 1u);                     // all stage-1 MMAs of the tile (and everything before) are complete
                ++acc_cnt;
                tc_fence_after_sync();
#pragma unroll 1
                for (int c0 = col_blk * 16; c0 < MV16; c0 += 64) {  // this warp's 16-column blocks
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
                fence_async_smem();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(a2_full);
                // ---- the X buffers are idle during stage 2: store the next tile's first two slabs now, so its stage 1
                //      starts the moment this tile's stage 2 is issued
                pre_stored = more;
                if (more) {
                    x_step(xr0, xm0, txm, th0, 128);
                    x_step(xr1, xm1, txm, th0, 192);
                }
            }

            // ---- epilogue 2: lane v < MV holds Z[:, v]; energy of the v-chunk = sum over lanes and all N columns
            wait(acc_ready, acc_cnt & 1u);
            ++acc_cnt;
            tc_fence_after_sync();
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
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_free);
            red[tid] = (int)lane_idx < MV ? e : 0.f;
            named_bar_sync(1, NC);
            if (tid < 32) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < NC / 32; ++j) s += red[tid + 32 * j];          // fixed order
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (tid == 0) {
                    atomicAdd(a.accum + (map % a.c_count), (double)s);
                    if (a.energy_out) atomicAdd(a.energy_out + map, s);
                }
            }
            named_bar_sync(1, NC);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace dctp
