// X ingest: everything the reference does to X on the host before the loop (base.py:243-267, 519-528,
// 200-201), as device passes over the uploaded image:
//   retile  raw X (any strides, f32/f64) -> tile-major Xt, + non-finite / negative flags, per-channel and
//           per-pixel "has a non-zero" marks (remove_zeros_lines), block sums (mean(X) for normalize)
//   fixup   all-zero rows / columns <- eps, then * scale           (only when needed)
//   const   sum X log max(X, ls) - sum X  (const_KL_, base.py:200-201), per tile, in fp64
#pragma once
#include "common.cuh"
#include "ingest_decl.h"

namespace espm {

// grid = (n_tiles, n_pad/32), 256 threads.
template <typename TS, typename TX>
__global__ void __launch_bounds__(256) retile_kernel(const TS* __restrict__ src, long long stride_c,
                                                     long long stride_p, long long j0, int n, int n_pad, int p_loc,
                                                     double scale, TX* __restrict__ Xt, IngestOut io) {
    __shared__ TX sm[32][TILE_PX + 1];
    __shared__ double sm_sum[8];
    __shared__ uint32_t sm_rownz, sm_colnz[4], sm_flags;
    const int tile = blockIdx.x, cb = blockIdx.y * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) sm_rownz = sm_flags = 0u;
    if (threadIdx.x < 4) sm_colnz[threadIdx.x] = 0u;
    __syncthreads();
    double sum = 0.0;
    uint32_t flags = 0u;
    auto look = [&](TS raw, bool valid) -> bool {   // returns "non-zero"
        if (!valid) return false;
        const double v = (double)raw;
        if (v != v) flags |= ESPM_X_NAN;
        else if (v > 1.7976931348623157e308 || v < -1.7976931348623157e308) flags |= ESPM_X_INF;
        else sum += v;
        if (v < 0.0) flags |= ESPM_X_NEGATIVE;
        return raw != TS(0);
    };
    if (stride_p == 1 || stride_c != 1) {
        // pixels contiguous (or generic): read rows of 128 pixels
        uint32_t colbits = 0u;   // bit i: pixel lane + 32*i of this tile has a non-zero in my rows
        for (int ci = warp; ci < 32; ci += 8) {
            const int c = cb + ci;
            bool row_any = false;
#pragma unroll
            for (int i = 0; i < TILE_PX / 32; ++i) {
                const int q = lane + 32 * i;
                const long long j = (long long)tile * TILE_PX + q;
                const bool valid = c < n && j < p_loc;
                const TS raw = valid ? src[(long long)c * stride_c + (j0 + j) * stride_p] : TS(0);
                const bool nz = look(raw, valid);
                row_any |= nz;
                if (nz) colbits |= 1u << i;
                sm[ci][q] = valid ? (TX)((double)raw * scale) : TX(0);
            }
            if (__any_sync(0xffffffffu, row_any) && lane == 0) atomicOr(&sm_rownz, 1u << ci);
        }
#pragma unroll
        for (int i = 0; i < TILE_PX / 32; ++i) {
            const uint32_t b = __ballot_sync(0xffffffffu, (colbits >> i) & 1u);
            if (lane == 0 && b) atomicOr(&sm_colnz[i], b);
        }
    } else {
        // channels contiguous (hyperspy layout): read 32 channels of one pixel per warp access
        uint32_t rowbits = 0u;
        for (int q = warp; q < TILE_PX; q += 8) {
            const long long j = (long long)tile * TILE_PX + q;
            const int c = cb + lane;
            const bool valid = c < n && j < p_loc;
            const TS raw = valid ? src[(long long)c * stride_c + (j0 + j) * stride_p] : TS(0);
            const bool nz = look(raw, valid);
            if (nz) rowbits = 1u;
            if (__any_sync(0xffffffffu, nz) && lane == 0) atomicOr(&sm_colnz[q >> 5], 1u << (q & 31));
            sm[lane][q] = valid ? (TX)((double)raw * scale) : TX(0);
        }
        const uint32_t b = __ballot_sync(0xffffffffu, rowbits);
        if (lane == 0 && b) atomicOr(&sm_rownz, b);
    }
    sum = warp_sum(sum);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0) {
        sm_sum[warp] = sum;
        if (flags) atomicOr(&sm_flags, flags);
    }
    __syncthreads();
    TX* dst = Xt + ((size_t)tile * n_pad + cb) * TILE_PX;
    for (int i = threadIdx.x; i < 32 * TILE_PX; i += 256) dst[i] = sm[i / TILE_PX][i % TILE_PX];
    if (io.flags) {
        // marks are idempotent stores of 1 => deterministic whatever the block order
        if (threadIdx.x < 32 && ((sm_rownz >> threadIdx.x) & 1u)) io.row_nz[cb + threadIdx.x] = 1;
        if (threadIdx.x < TILE_PX && ((sm_colnz[threadIdx.x >> 5] >> (threadIdx.x & 31)) & 1u))
            io.col_nz[(size_t)tile * TILE_PX + threadIdx.x] = 1;
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += sm_sum[w];
            io.sum_part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
            if (sm_flags) atomicOr(io.flags, sm_flags);
        }
    }
}

// Pre-scan of the raw X: NaN / inf / negative / fractional flags, the largest entry, and the non-zero marks of channels
// and pixels.  A CTA covers 256 "fast" indices (the contiguous axis: pixels, or channels in the hyperspy layout) times
// 64 "slow" ones; grid = (ceil(fast / 256), ceil(slow / 64)).  Marks are idempotent stores of 1, the maximum is an
// atomicMax on the bit pattern of a non-negative float: deterministic whatever the block order.
template <typename TS>
__global__ void __launch_bounds__(256) prescan_kernel(const TS* __restrict__ src, long long stride_c, long long stride_p,
                                                      long long j0, int n, long long p_loc, int px_fast,
                                                      uint32_t* __restrict__ out4, int32_t* __restrict__ row_nz,
                                                      int32_t* __restrict__ col_nz) {
    const long long n_fast = px_fast ? p_loc : n, n_slow = px_fast ? n : p_loc;
    const long long f = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long s0 = (long long)blockIdx.y * 64;
    uint32_t flags = 0u;
    float vmax = 0.f;
    bool fast_any = false;
    for (int i = 0; i < 64; ++i) {
        const long long sl = s0 + i;
        bool nz = false;
        if (f < n_fast && sl < n_slow) {
            const long long c = px_fast ? sl : f, j = px_fast ? f : sl;
            const TS raw = src[c * stride_c + (j0 + j) * stride_p];
            const double v = (double)raw;
            if (v != v) flags |= ESPM_X_NAN;
            else if (v > 1.7976931348623157e308 || v < -1.7976931348623157e308) flags |= ESPM_X_INF;
            else {
                if (v < 0.0) flags |= ESPM_X_NEGATIVE;
                if (v != floor(v)) flags |= ESPM_X_FRACTION;
                vmax = fmaxf(vmax, v > 3.0e38 ? 3.0e38f : __double2float_ru(v));
            }
            nz = raw != TS(0);
        }
        fast_any |= nz;
        // the slow index is common to the warp: one mark per (warp, slow index) when any lane saw a non-zero
        if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0 && sl < n_slow) (px_fast ? row_nz : col_nz)[sl] = 1;
    }
    if (fast_any && f < n_fast) (px_fast ? col_nz : row_nz)[f] = 1;
    flags = __reduce_or_sync(0xffffffffu, flags);
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0) {
        if (flags) atomicOr(&out4[0], flags);
        atomicMax(&out4[1], __float_as_uint(vmax));
    }
}

// Xt <- (zero row or zero column ? eps : Xt) * scale, on the real entries only.  grid = (n_tiles, n_pad/32).
template <typename TX>
__global__ void __launch_bounds__(256) xt_fixup_kernel(TX* __restrict__ Xt, const int32_t* __restrict__ row_zero,
                                                       const int32_t* __restrict__ col_zero, int n, int n_pad,
                                                       int p_loc, double eps, double scale) {
    const int tile = blockIdx.x, cb = blockIdx.y * 32;
    TX* base = Xt + ((size_t)tile * n_pad + cb) * TILE_PX;
    for (int i = threadIdx.x; i < 32 * TILE_PX; i += 256) {
        const int c = cb + i / TILE_PX, q = i % TILE_PX;
        const long long j = (long long)tile * TILE_PX + q;
        if (c < n && j < p_loc) {
            double v = (double)base[i];
            if ((row_zero && row_zero[c]) || (col_zero && col_zero[j])) v = eps;
            base[i] = (TX)(v * scale);
        }
    }
}

// part[tile] = sum_{c<n, j in tile} v*log(max(v, ls)) - v   (fp64 accumulation, fixed order)
template <typename TX>
__global__ void __launch_bounds__(256) xt_const_kernel(const TX* __restrict__ Xt, int n, int n_pad, int p_loc,
                                                       double ls, double* __restrict__ part) {
    __shared__ double sm[8];
    const int tile = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TX* base = Xt + (size_t)tile * n_pad * TILE_PX;
    const int q = threadIdx.x & (TILE_PX - 1);
    const bool px_ok = (long long)tile * TILE_PX + q < p_loc;
    double acc = 0.0;
    if (px_ok) {
        for (int c = threadIdx.x / TILE_PX; c < n; c += 256 / TILE_PX) {
            const double v = (double)base[(size_t)c * TILE_PX + q];
            if (v != 0.0) acc += v * log(fmax(v, ls)) - v;   // 0*log(ls) - 0 == 0
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        part[tile] = t;
    }
}

// out[0] = sum(in[0..n)) with a fixed summation tree (single CTA of 1024 threads).
__global__ void __launch_bounds__(1024) reduce_sum_kernel(const double* __restrict__ in, long long n,
                                                          double* __restrict__ out) {
    __shared__ double sm[32];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) acc += in[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += sm[w];
        out[0] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// Sums of X for the Bregman branches: one CTA per pixel tile of Xt ([n_pad][128]).
//   colsum[j] = sum_c X[c][j]  (updates.py:121);  rowsum_part[tile][c] = sum_{j in tile} X[c][j]  (updates.py:43)
// ------------------------------------------------------------------------------------------------
template <typename TX, typename TC>
__global__ void __launch_bounds__(128) x_sums_kernel(const TX* Xt, int n_pad, TC* colsum, double* rowsum_part) {
    const TX* tile = Xt + (size_t)blockIdx.x * n_pad * TILE_PX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double cs = 0.0;
    __shared__ double sm[4];
    for (int c = 0; c < n_pad; ++c) {
        const double v = (double)tile[(size_t)c * TILE_PX + threadIdx.x];
        cs += v;
        const double r = warp_sum(v);
        if (lane == 0) sm[warp] = r;
        __syncthreads();
        if (threadIdx.x == 0) rowsum_part[(size_t)blockIdx.x * n_pad + c] = (sm[0] + sm[1]) + (sm[2] + sm[3]);
        __syncthreads();
    }
    colsum[(size_t)blockIdx.x * TILE_PX + threadIdx.x] = (TC)cs;
}


}  // namespace espm
