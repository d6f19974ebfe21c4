// Generic fp32 CUDA-core DCT-score kernels: any H x W (non-square, odd, > 128), strided rows.
// They cover the shapes the tensor-core kernel does not take and serve as an on-device
// cross-check of it (same C-ABI entry, path = DCTP_PATH_SIMT).  Same contract as
// /root/reference/utils/common.py:262-277: energy of the orthonormal 2-D DCT-II per (image, channel),
// summed per channel.
//
//   small  (H, W <= 64): G = 64 / H maps stacked per CTA pass; X, Y = X*C_W^T and both bases in smem.
//   large  (anything)  : one CTA per (map, 64-column panel v0..v0+63):
//                          Y[:, panel] = X * C_W[panel, :]^T  (X streamed in 64x64 blocks, Y panel kept in smem)
//                          Z[:, panel] = C_H * Y[:, panel]    -> sum of squares
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dctp {

struct SimtScoreArgs {
    const float* x;
    long long stride_b, stride_c, stride_h;   // elements; stride_w == 1
    int c_begin, c_count, n_maps, H, W;
    const float* basis_h_t;    // [H x H]  basis_h_t[h*H + u] = C_H[u][h]
    const float* basis_w_t;    // [W x W]  basis_w_t[w*W + v] = C_W[v][w]
    double* accum;
    float* energy_out;         // optional [n_maps]; the large kernel accumulates into it (caller zeroes)
    float* dump;               // optional [n_maps x H x W] coefficients Z[u][v]
};

__device__ __forceinline__ const float* simt_map_ptr(const SimtScoreArgs& a, int m) {
    int b = m / a.c_count, c = m - b * a.c_count;
    return a.x + b * a.stride_b + (long long)(a.c_begin + c) * a.stride_c;
}

// deterministic block-wide sum of one float per thread (fixed tree), result valid in thread 0
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0)
        for (int i = 0; i < THREADS / 32; ++i) s += scratch[i];
    __syncthreads();
    return s;
}

constexpr int SIMT_T = 64;          // tile edge
constexpr int SIMT_LD = SIMT_T + 4; // padded leading dimension (keeps float4 rows 16-B aligned)
constexpr int SIMT_SMALL_SMEM = (4 * SIMT_T * SIMT_LD + SIMT_T * 16) * 4;

__global__ void __launch_bounds__(256) score_simt_small_kernel(const SimtScoreArgs a) {
    extern __shared__ __align__(16) float smem_f[];       // SIMT_SMALL_SMEM bytes
    float (*Xs)[SIMT_LD] = reinterpret_cast<float (*)[SIMT_LD]>(smem_f);
    float (*Ys)[SIMT_LD] = reinterpret_cast<float (*)[SIMT_LD]>(smem_f + SIMT_T * SIMT_LD);
    float (*Cw)[SIMT_LD] = reinterpret_cast<float (*)[SIMT_LD]>(smem_f + 2 * SIMT_T * SIMT_LD);   // Cw[w][v]
    float (*Ch)[SIMT_LD] = reinterpret_cast<float (*)[SIMT_LD]>(smem_f + 3 * SIMT_T * SIMT_LD);   // Ch[h][u]
    float* part = smem_f + 4 * SIMT_T * SIMT_LD;
    const int tid = threadIdx.x, H = a.H, W = a.W;
    const int G = SIMT_T / H, Wq = (W + 3) / 4;

    for (int i = tid; i < SIMT_T * SIMT_T; i += 256) {
        int r = i / SIMT_T, c = i % SIMT_T;
        Cw[r][c] = (r < W && c < W) ? a.basis_w_t[r * W + c] : 0.f;
        Ch[r][c] = (r < H && c < H) ? a.basis_h_t[r * H + c] : 0.f;
    }
    const int num_tiles = (a.n_maps + G - 1) / G;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int map0 = tile * G, maps_here = min(G, a.n_maps - map0), rows = maps_here * H;
        __syncthreads();
        for (int i = tid; i < rows * W; i += 256) {
            int r = i / W, w = i - r * W, g = r / H, h = r - g * H;
            Xs[r][w] = simt_map_ptr(a, map0 + g)[(long long)h * a.stride_h + w];
        }
        __syncthreads();
        // Y[r][v] = sum_w X[r][w] * C_W[v][w]
        for (int o = tid; o < rows * Wq; o += 256) {
            int r = o / Wq, v0 = (o - r * Wq) * 4;
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
            for (int w = 0; w < W; ++w) {
                float xv = Xs[r][w];
                float4 cv = *reinterpret_cast<const float4*>(&Cw[w][v0]);
                acc0 = fmaf(xv, cv.x, acc0); acc1 = fmaf(xv, cv.y, acc1);
                acc2 = fmaf(xv, cv.z, acc2); acc3 = fmaf(xv, cv.w, acc3);
            }
            *reinterpret_cast<float4*>(&Ys[r][v0]) = make_float4(acc0, acc1, acc2, acc3);
        }
        __syncthreads();
        // Z[(g,u)][v] = sum_h C_H[u][h] * Y[(g,h)][v]
        for (int o = tid; o < SIMT_T * 16; o += 256) part[o] = 0.f;
        __syncthreads();
        for (int o = tid; o < rows * Wq; o += 256) {
            int r = o / Wq, q = o - r * Wq, v0 = q * 4, g = r / H, u = r - g * H;
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
            for (int h = 0; h < H; ++h) {
                float cv = Ch[h][u];
                float4 yv = *reinterpret_cast<const float4*>(&Ys[g * H + h][v0]);
                acc0 = fmaf(cv, yv.x, acc0); acc1 = fmaf(cv, yv.y, acc1);
                acc2 = fmaf(cv, yv.z, acc2); acc3 = fmaf(cv, yv.w, acc3);
            }
            float z[4] = {acc0, acc1, acc2, acc3};
            float e = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (v0 + j < W) {
                    e = fmaf(z[j], z[j], e);
                    if (a.dump) a.dump[((long long)(map0 + g) * H + u) * W + v0 + j] = z[j];
                }
            part[r * 16 + q] = e;
        }
        __syncthreads();
        // fixed-order per-map reduction: warp `g % 8` sums map g's H*16 partials
        for (int g = tid >> 5; g < maps_here; g += 8) {
            float s = 0.f;
            for (int i = (tid & 31); i < H * 16; i += 32) s += part[g * H * 16 + i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((tid & 31) == 0) {
                int m = map0 + g;
                atomicAdd(a.accum + (m % a.c_count), (double)s);
                if (a.energy_out) a.energy_out[m] = s;
            }
        }
    }
}

// grid = n_maps * panels (map-major); dynamic smem = Ybuf[Hpad][SIMT_LD] + 2 staging tiles
__global__ void __launch_bounds__(256) score_simt_large_kernel(const SimtScoreArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    const int H = a.H, W = a.W;
    const int Hpad = (H + SIMT_T - 1) / SIMT_T * SIMT_T;
    float (*Yb)[SIMT_LD] = reinterpret_cast<float (*)[SIMT_LD]>(smem_f);                       // [Hpad][LD]
    float (*Ts)[SIMT_LD] = reinterpret_cast<float (*)[SIMT_LD]>(smem_f + Hpad * SIMT_LD);        // X block / C_H block
    float (*Bs)[SIMT_LD] = reinterpret_cast<float (*)[SIMT_LD]>(smem_f + (Hpad + SIMT_T) * SIMT_LD);
    __shared__ float scratch[8];
    const int panels = (W + SIMT_T - 1) / SIMT_T;
    const int tid = threadIdx.x, m = blockIdx.x / panels, v0 = (blockIdx.x % panels) * SIMT_T;
    const float* xm = simt_map_ptr(a, m);
    const int row = tid >> 2, cq = (tid & 3) * 16;      // thread owns 1 row x 16 columns of a 64x64 block

    // ---- stage 1: Y[h][v0+j] for all h
    for (int rb = 0; rb < Hpad; rb += SIMT_T) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        for (int wc = 0; wc < W; wc += SIMT_T) {
            __syncthreads();
            for (int i = tid; i < SIMT_T * SIMT_T; i += 256) {
                int r = i >> 6, c = i & 63;
                Ts[r][c] = (rb + r < H && wc + c < W) ? xm[(long long)(rb + r) * a.stride_h + wc + c] : 0.f;
                Bs[r][c] = (wc + r < W && v0 + c < W) ? a.basis_w_t[(long long)(wc + r) * W + v0 + c] : 0.f;   // [w][v]
            }
            __syncthreads();
#pragma unroll 4
            for (int w = 0; w < SIMT_T; ++w) {
                float xv = Ts[row][w];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float4 cv = *reinterpret_cast<const float4*>(&Bs[w][cq + 4 * j4]);
                    acc[4 * j4 + 0] = fmaf(xv, cv.x, acc[4 * j4 + 0]);
                    acc[4 * j4 + 1] = fmaf(xv, cv.y, acc[4 * j4 + 1]);
                    acc[4 * j4 + 2] = fmaf(xv, cv.z, acc[4 * j4 + 2]);
                    acc[4 * j4 + 3] = fmaf(xv, cv.w, acc[4 * j4 + 3]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) Yb[rb + row][cq + j] = acc[j];
    }
    // ---- stage 2: Z[u][v0+j] = sum_h C_H[u][h] * Y[h][v0+j]
    float e = 0.f;
    for (int ub = 0; ub < Hpad; ub += SIMT_T) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        for (int hc = 0; hc < Hpad; hc += SIMT_T) {
            __syncthreads();
            for (int i = tid; i < SIMT_T * SIMT_T; i += 256) {
                int r = i >> 6, c = i & 63;                     // Ts[h][u] = C_H[ub+u][hc+h]
                Ts[r][c] = (hc + r < H && ub + c < H) ? a.basis_h_t[(long long)(hc + r) * H + ub + c] : 0.f;
            }
            __syncthreads();
#pragma unroll 4
            for (int h = 0; h < SIMT_T; ++h) {
                float cv = Ts[h][row];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float4 yv = *reinterpret_cast<const float4*>(&Yb[hc + h][cq + 4 * j4]);
                    acc[4 * j4 + 0] = fmaf(cv, yv.x, acc[4 * j4 + 0]);
                    acc[4 * j4 + 1] = fmaf(cv, yv.y, acc[4 * j4 + 1]);
                    acc[4 * j4 + 2] = fmaf(cv, yv.z, acc[4 * j4 + 2]);
                    acc[4 * j4 + 3] = fmaf(cv, yv.w, acc[4 * j4 + 3]);
                }
            }
        }
        const int u = ub + row;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int v = v0 + cq + j;
            if (u < H && v < W) {
                e = fmaf(acc[j], acc[j], e);
                if (a.dump) a.dump[((long long)m * H + u) * W + v] = acc[j];
            }
        }
    }
    float s = block_sum<256>(e, scratch);
    if (tid == 0) {
        atomicAdd(a.accum + (m % a.c_count), (double)s);
        if (a.energy_out) atomicAdd(a.energy_out + m, s);
    }
}

}  // namespace dctp
