// Segmented top-k channel selection + score finalisation.
//
// One segment = one score file of the reference; the selection replaces
//   select_index = np.argsort(imp)[C-k:]; select_index.sort()
// at /root/reference/utils/load_models.py:39-41 and its sibling sites (:102-104, :265-267, :313-315,
// :352-354, :407-409, :469-471, :521-523, :629-631 ... :746-748).
//
// Tie rule (documented deviation): the reference's default argsort is unstable, so which
// of several channels equal to the cut value survive is implementation-defined there.  Here a channel is
// kept iff fewer than k channels beat it, where j beats i when  s_j > s_i  or  (s_j == s_i and j > i):
// exactly np.argsort(imp, kind='stable')[C-k:].  NaN sorts last (largest) as in numpy; -0.0 == +0.0.
//
// One CTA per segment: exact rank by counting (C <= a few thousand, O(C^2) compares out of shared memory),
// then an order-preserving compaction so the kept ids come out ascending.  HBM traffic is 4C bytes in,
// 8k bytes out per segment: the kernel is latency-bound, not bandwidth-bound, and runs once per layer.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dctp {

constexpr int TOPK_THREADS = 256;
constexpr int TOPK_SMEM_KEYS = 8192;      // segments up to this many channels rank out of shared memory

// monotone map float -> uint32 (total order: -inf < ... < -0 == +0 < ... < +inf < NaN)
__device__ __forceinline__ uint32_t topk_key(float f) {
    if (f != f) return 0xFFFFFFFFu;
    if (f == 0.f) f = 0.f;                // folds -0.0 onto +0.0
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(TOPK_THREADS) topk_segmented_kernel(const float* __restrict__ scores,
                                                                      const int* __restrict__ seg_offsets,   // [n_seg+1]
                                                                      const int* __restrict__ seg_k,         // [n_seg]
                                                                      long long* __restrict__ out_idx,
                                                                      const int* __restrict__ out_offsets) { // [n_seg+1]
    __shared__ uint32_t keys[TOPK_SMEM_KEYS];
    __shared__ int warp_sums[TOPK_THREADS / 32];
    __shared__ int carry;
    const int seg = blockIdx.x, tid = threadIdx.x;
    const int lo = seg_offsets[seg], C = seg_offsets[seg + 1] - lo;
    int k = seg_k[seg];
    k = k < 0 ? 0 : (k > C ? C : k);
    const float* s = scores + lo;
    long long* out = out_idx + out_offsets[seg];
    const bool in_smem = C <= TOPK_SMEM_KEYS;
    if (in_smem)
        for (int i = tid; i < C; i += TOPK_THREADS) keys[i] = topk_key(s[i]);
    if (tid == 0) carry = 0;
    __syncthreads();

    for (int base = 0; base < C; base += TOPK_THREADS) {
        const int i = base + tid;
        int keep = 0;
        if (i < C) {
            const uint32_t ki = in_smem ? keys[i] : topk_key(s[i]);
            int beat = 0;                                 // channels that outrank i
            if (in_smem) {
                for (int j = 0; j < C; ++j) {
                    uint32_t kj = keys[j];
                    beat += (kj > ki) || (kj == ki && j > i);
                }
            } else {
                for (int j = 0; j < C; ++j) {
                    uint32_t kj = topk_key(s[j]);
                    beat += (kj > ki) || (kj == ki && j > i);
                }
            }
            keep = beat < k;
        }
        // order-preserving compaction: exclusive scan of `keep` over the block
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        const int lane = tid & 31, w = tid >> 5;
        const int before = __popc(ballot & ((1u << lane) - 1));
        if (lane == 0) warp_sums[w] = __popc(ballot);
        __syncthreads();
        int woff = 0, total = 0;
        for (int q = 0; q < TOPK_THREADS / 32; ++q) {
            int c = warp_sums[q];
            if (q < w) woff += c;
            total += c;
        }
        const int start = carry;
        if (keep) out[start + woff + before] = i;
        __syncthreads();
        if (tid == 0) carry = start + total;
        __syncthreads();
    }
}

// out[i] = float(accum[i] / n_images): the reference's running mean over images
// (/root/reference/utils/common.py:275-277) collapses to sum / N.
__global__ void finalize_kernel(const double* __restrict__ accum, double n_images, float* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<float>(accum[i] / n_images);
}

}  // namespace dctp
