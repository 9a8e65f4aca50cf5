// The two kernels that stream X from HBM: the H pass and the W pass.
//
// Both are warp-specialised persistent kernels: one producer warp issues TMA bulk copies
// (cp.async.bulk, SASS UBLKCP) of 16 KiB X chunks + the matching GW rows into a ring of shared-memory
// stages guarded by mbarriers; eight consumer warps read the stages with 128-bit LDS, rebuild
// y = GW.H in registers (never materialised), form x/y and contract it on the fly.
//
//   h_pass : numraw[k][j] = sum_c GW[c][k] * X[c][j] / y[c][j]      (updates.py:127-128)
//            + partial sums of max(X,ls)*log(y)                      (measures.py:497-503)
//   w_pass : S[c][k]      = sum_j X[c][j] / y'[c][j] * H'[k][j]      (updates.py:53-59, G^T(R H^T))
#pragma once
#include "common.cuh"

namespace espm {

struct XPassArgs {
    const void* Xt;
    const void* GW;
    const void* GWc;
    const void* H;       // H_cur for the H pass, H_next for the W pass (points at local pixel 0)
    void* numraw;        // [nsplit][KP][P_pad]
    double* xlogy_part;  // [grid]
    void* s_part;        // [w_nr][n_pad][KP]
    int n_pad, k, n_tiles, ldh, p_pad;
    int nstages_tile;    // n_pad / CS
    int nsplit;          // H pass
    int w_nb, w_nr;      // W pass
    int depth;           // pipeline stages
    int sacc_rows;       // W pass: channel rows reserved per half in the smem accumulator
    int clamp_y, dual;   // SAFE variants only
    double log_shift;
};

template <typename TX, typename TC, int KP, bool SAFE>
struct XPassSmem {
    using G = PassGeom<TX, TC>;
    static constexpr int X_BYTES = STAGE_BYTES;
    static constexpr int GW_BYTES = G::CS * KP * (int)sizeof(TC);
    static constexpr int GW_BYTES_AL = (GW_BYTES + 127) / 128 * 128;
    static constexpr int STAGE_STRIDE = X_BYTES + GW_BYTES_AL * (SAFE ? 2 : 1);
    static constexpr int BAR_BYTES = 256;  // up to 16 full + 16 empty barriers
    static constexpr int RED_BYTES = G::NSLOT * KP * TILE_PX * (int)sizeof(TC);
    static constexpr int MISC_BYTES = 128;
    static __host__ __device__ int h_bytes(int depth) { return BAR_BYTES + MISC_BYTES + depth * STAGE_STRIDE + RED_BYTES; }
    static __host__ __device__ int w_bytes(int depth, int sacc_rows) {
        return BAR_BYTES + MISC_BYTES + depth * STAGE_STRIDE + G::HALVES * sacc_rows * KP * (int)sizeof(TC);
    }
};

template <typename T, int N>
__device__ __forceinline__ void lds_vec(T (&dst)[N], const void* src) {
    constexpr int BYTES = N * (int)sizeof(T);
    if constexpr (BYTES == 16) {
        uint4 v = *reinterpret_cast<const uint4*>(src);
        memcpy(dst, &v, 16);
    } else if constexpr (BYTES == 8) {
        uint2 v = *reinterpret_cast<const uint2*>(src);
        memcpy(dst, &v, 8);
    } else {
        const T* s = reinterpret_cast<const T*>(src);
#pragma unroll
        for (int i = 0; i < N; ++i) dst[i] = s[i];
    }
}

template <typename TC, int KP>
__device__ __forceinline__ void lds_gw(TC (&gw)[KP], const void* row) {
    constexpr int BYTES = KP * (int)sizeof(TC);
    if constexpr (BYTES % 16 == 0) {
        const uint4* s = reinterpret_cast<const uint4*>(row);
        uint4 tmp[BYTES / 16];
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) tmp[i] = s[i];
        memcpy(gw, tmp, BYTES);
    } else if constexpr (BYTES % 8 == 0) {
        const uint2* s = reinterpret_cast<const uint2*>(row);
        uint2 tmp[BYTES / 8];
#pragma unroll
        for (int i = 0; i < BYTES / 8; ++i) tmp[i] = s[i];
        memcpy(gw, tmp, BYTES);
    } else {
        const TC* s = reinterpret_cast<const TC*>(row);
#pragma unroll
        for (int i = 0; i < KP; ++i) gw[i] = s[i];
    }
}

// Loads the PPL consecutive H values of KP rows for this lane's pixels (H pad pixels hold 1, pad
// rows are treated as 0).
template <typename TC, int KP, int PPL>
__device__ __forceinline__ void load_h(TC (&h)[KP][PPL], const TC* H, int ldh, int k, int px0) {
#pragma unroll
    for (int kk = 0; kk < KP; ++kk) {
        if (kk < k) {
            lds_vec<TC, PPL>(h[kk], H + (size_t)kk * ldh + px0);
        } else {
#pragma unroll
            for (int q = 0; q < PPL; ++q) h[kk][q] = TC(0);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Producer: one elected lane streams (tile, stage) chunks into the ring.
// ------------------------------------------------------------------------------------------------
template <typename TX, typename TC, int KP, bool SAFE>
struct Ring {
    using S = XPassSmem<TX, TC, KP, SAFE>;
    using G = PassGeom<TX, TC>;
    uint64_t* full;
    uint64_t* empty;
    unsigned char* stages;
    int depth;
    __device__ __forceinline__ Ring(unsigned char* smem, int depth_) : depth(depth_) {
        full = reinterpret_cast<uint64_t*>(smem);
        empty = full + 16;
        stages = smem + S::BAR_BYTES + S::MISC_BYTES;
    }
    __device__ __forceinline__ void init() {
        for (int i = 0; i < depth; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], N_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __device__ __forceinline__ unsigned char* x_ptr(int slot) { return stages + (size_t)slot * S::STAGE_STRIDE; }
    __device__ __forceinline__ unsigned char* gw_ptr(int slot) { return x_ptr(slot) + S::X_BYTES; }
    __device__ __forceinline__ unsigned char* gwc_ptr(int slot) { return gw_ptr(slot) + S::GW_BYTES_AL; }

    __device__ __forceinline__ void produce(const XPassArgs& a, int tile, int st, uint32_t cnt, uint64_t pol_x,
                                            uint64_t pol_gw) {
        const int slot = cnt % depth;
        const uint32_t phase = (cnt / depth) & 1u;
        mbar_wait(&empty[slot], phase ^ 1u);
        const bool dual = SAFE && a.dual;
        mbar_expect_tx(&full[slot], S::X_BYTES + S::GW_BYTES * (dual ? 2 : 1));
        const TX* xsrc = reinterpret_cast<const TX*>(a.Xt) + ((size_t)tile * a.n_pad + (size_t)st * G::CS) * TILE_PX;
        tma_bulk_g2s(x_ptr(slot), xsrc, S::X_BYTES, &full[slot], pol_x);
        const TC* gsrc = reinterpret_cast<const TC*>(a.GW) + (size_t)st * G::CS * KP;
        tma_bulk_g2s(gw_ptr(slot), gsrc, S::GW_BYTES, &full[slot], pol_gw);
        if (dual) {
            const TC* csrc = reinterpret_cast<const TC*>(a.GWc) + (size_t)st * G::CS * KP;
            tma_bulk_g2s(gwc_ptr(slot), csrc, S::GW_BYTES, &full[slot], pol_gw);
        }
    }
    __device__ __forceinline__ void consumer_wait(uint32_t cnt) {
        mbar_wait(&full[cnt % depth], (cnt / depth) & 1u);
    }
    __device__ __forceinline__ void consumer_release(uint32_t cnt, int lane) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[cnt % depth]);
    }
};

// ------------------------------------------------------------------------------------------------
// H pass
// ------------------------------------------------------------------------------------------------
template <typename TX, typename TC, int KP, bool SAFE>
__global__ void __launch_bounds__(XPASS_THREADS, (KP * sizeof(TC) <= 32) ? 2 : 1)
h_pass_kernel(const XPassArgs a) {
    using G = PassGeom<TX, TC>;
    using S = XPassSmem<TX, TC, KP, SAFE>;
    constexpr int PPL = G::PPL;
    extern __shared__ __align__(128) unsigned char smem[];
    Ring<TX, TC, KP, SAFE> ring(smem, a.depth);
    double* misc = reinterpret_cast<double*>(smem + S::BAR_BYTES);
    TC* red = reinterpret_cast<TC*>(smem + S::BAR_BYTES + S::MISC_BYTES + (size_t)a.depth * S::STAGE_STRIDE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) ring.init();
    __syncthreads();

    const int n_items = a.n_tiles * a.nsplit;
    const int NS = a.nstages_tile;

    if (warp == N_CONSUMER_WARPS) {
        // ---------------- producer warp ----------------
        if (lane == 0) {
            const uint64_t pol_x = l2_policy_evict_first();
            const uint64_t pol_gw = l2_policy_evict_last();
            uint32_t cnt = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int tile = it / a.nsplit, split = it - tile * a.nsplit;
                const int s0 = (int)((long long)split * NS / a.nsplit), s1 = (int)((long long)(split + 1) * NS / a.nsplit);
                for (int st = s0; st < s1; ++st, ++cnt) ring.produce(a, tile, st, cnt, pol_x, pol_gw);
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int half = warp % G::HALVES, slot = warp / G::HALVES;
    const int lane_px = half * (32 * PPL) + lane * PPL;  // pixel offset inside the tile
    const TC ls = (TC)a.log_shift;
    const TC* Hc = reinterpret_cast<const TC*>(a.H);
    double xl_total = 0.0;   // sum x*log2(y) over x>0
    double zl_total = 0.0;   // sum log2(y) over x==0 (fp64 only; weighted by log_shift at the end)
    uint32_t cnt = 0;

    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int tile = it / a.nsplit, split = it - tile * a.nsplit;
        const int s0 = (int)((long long)split * NS / a.nsplit), s1 = (int)((long long)(split + 1) * NS / a.nsplit);
        TC h[KP][PPL], hc[SAFE ? KP : 1][PPL], num[KP][PPL];
        load_h<TC, KP, PPL>(h, Hc, a.ldh, a.k, tile * TILE_PX + lane_px);
#pragma unroll
        for (int kk = 0; kk < KP; ++kk)
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                num[kk][q] = TC(0);
                if constexpr (SAFE) hc[kk][q] = (kk < a.k) ? Num<TC>::vmax(h[kk][q], ls) : TC(0);
            }

        for (int st = s0; st < s1; ++st, ++cnt) {
            ring.consumer_wait(cnt);
            const unsigned char* xs = ring.x_ptr(cnt % a.depth);
            const unsigned char* gs = ring.gw_ptr(cnt % a.depth);
            const unsigned char* gcs = ring.gwc_ptr(cnt % a.depth);
            TC xl = TC(0);
            float zl = 0.f;
#pragma unroll
            for (int ci = 0; ci < G::CPW; ++ci) {
                const int c = slot + ci * G::NSLOT;
                TX xv[PPL];
                TC gw[KP];
                lds_vec<TX, PPL>(xv, xs + ((size_t)c * TILE_PX + lane_px) * sizeof(TX));
                lds_gw<TC, KP>(gw, gs + (size_t)c * KP * sizeof(TC));
                TC y[PPL], r[PPL];
#pragma unroll
                for (int q = 0; q < PPL; ++q) {
                    y[q] = gw[0] * h[0][q];
#pragma unroll
                    for (int kk = 1; kk < KP; ++kk) y[q] = fma(gw[kk], h[kk][q], y[q]);
                }
                if constexpr (SAFE) {
                    if (a.clamp_y) {
#pragma unroll
                        for (int q = 0; q < PPL; ++q) y[q] = Num<TC>::vmax(y[q], ls);
                    }
                }
#pragma unroll
                for (int q = 0; q < PPL; ++q) r[q] = Num<TC>::ratio((TC)xv[q], y[q]);
#pragma unroll
                for (int kk = 0; kk < KP; ++kk)
#pragma unroll
                    for (int q = 0; q < PPL; ++q) num[kk][q] = fma(gw[kk], r[q], num[kk][q]);
                // ---- loss of the current iterate: sum max(x,ls)*log(Y), Y from clamped GW, H ----
                TC yl[PPL];
                if constexpr (SAFE) {
                    if (a.dual) {
                        TC gwc[KP];
                        lds_gw<TC, KP>(gwc, gcs + (size_t)c * KP * sizeof(TC));
#pragma unroll
                        for (int q = 0; q < PPL; ++q) {
                            yl[q] = gwc[0] * hc[0][q];
#pragma unroll
                            for (int kk = 1; kk < KP; ++kk) yl[q] = fma(gwc[kk], hc[kk][q], yl[q]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < PPL; ++q) yl[q] = y[q];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < PPL; ++q) yl[q] = y[q];
                }
                if constexpr (sizeof(TC) == 4 && !SAFE) {
                    // fp32 fast path: 0*log2(y) == 0 because y > 0 is guaranteed here; the reference's
                    // ls*log(Y) terms of zero entries are below fp32 rounding of the sum.
#pragma unroll
                    for (int q = 0; q < PPL; ++q) xl = fma((TC)xv[q], Num<TC>::log2_fast(yl[q]), xl);
                } else {
#pragma unroll
                    for (int q = 0; q < PPL; ++q) {
                        const TC x = (TC)xv[q];
                        if (x > TC(0)) {
                            xl = fma(Num<TC>::vmax(x, ls), Num<TC>::log2_fast(yl[q]), xl);
                        } else {
                            zl += __log2f((float)yl[q]);
                        }
                    }
                }
            }
            ring.consumer_release(cnt, lane);
            xl_total += (double)xl;
            zl_total += (double)zl;
        }

        // ---- cross-warp reduction of the ratio sums of this item (fixed order => deterministic) ----
#pragma unroll
        for (int kk = 0; kk < KP; ++kk)
#pragma unroll
            for (int q = 0; q < PPL; ++q) red[((size_t)slot * KP + kk) * TILE_PX + lane_px + q] = num[kk][q];
        named_bar_sync(1, N_CONSUMER_THREADS);
        TC* out = reinterpret_cast<TC*>(a.numraw) + (size_t)split * KP * a.p_pad + (size_t)tile * TILE_PX;
        for (int idx = threadIdx.x; idx < KP * TILE_PX; idx += N_CONSUMER_THREADS) {
            const int kk = idx / TILE_PX, q = idx - kk * TILE_PX;
            TC s = red[(size_t)kk * TILE_PX + q];
#pragma unroll
            for (int sl = 1; sl < G::NSLOT; ++sl) s += red[((size_t)sl * KP + kk) * TILE_PX + q];
            out[(size_t)kk * a.p_pad + q] = s;
        }
        named_bar_sync(1, N_CONSUMER_THREADS);
    }

    // ---- loss partial of this CTA ----
    // pad pixels / pad channels contribute x == 0 only; the zero-entry term is weighted by log_shift.
    double v = xl_total + a.log_shift * zl_total;
    v = warp_sum(v);
    if (lane == 0) misc[warp] = v;
    named_bar_sync(1, N_CONSUMER_THREADS);
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < N_CONSUMER_WARPS; ++w) s += misc[w];
        a.xlogy_part[blockIdx.x] = s * 0.6931471805599453094;  // log2 -> ln
    }
}

// ------------------------------------------------------------------------------------------------
// W pass
// ------------------------------------------------------------------------------------------------
template <typename TX, typename TC, int KP, bool SAFE>
__global__ void __launch_bounds__(XPASS_THREADS, (KP * sizeof(TC) <= 32) ? 2 : 1)
w_pass_kernel(const XPassArgs a) {
    using G = PassGeom<TX, TC>;
    using S = XPassSmem<TX, TC, KP, SAFE>;
    constexpr int PPL = G::PPL;
    extern __shared__ __align__(128) unsigned char smem[];
    Ring<TX, TC, KP, SAFE> ring(smem, a.depth);
    TC* sacc = reinterpret_cast<TC*>(smem + S::BAR_BYTES + S::MISC_BYTES + (size_t)a.depth * S::STAGE_STRIDE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NS = a.nstages_tile;
    const int b = blockIdx.x % a.w_nb, r = blockIdx.x / a.w_nb;
    const int sb0 = (int)((long long)b * NS / a.w_nb), sb1 = (int)((long long)(b + 1) * NS / a.w_nb);
    const int t0 = (int)((long long)r * a.n_tiles / a.w_nr), t1 = (int)((long long)(r + 1) * a.n_tiles / a.w_nr);
    const int rows = (sb1 - sb0) * G::CS;

    if (threadIdx.x == 0) ring.init();
    for (int i = threadIdx.x; i < G::HALVES * a.sacc_rows * KP; i += blockDim.x) sacc[i] = TC(0);
    __syncthreads();

    if (warp == N_CONSUMER_WARPS) {
        if (lane == 0) {
            const uint64_t pol_x = l2_policy_evict_first();
            const uint64_t pol_gw = l2_policy_evict_last();
            uint32_t cnt = 0;
            for (int tile = t0; tile < t1; ++tile)
                for (int st = sb0; st < sb1; ++st, ++cnt) ring.produce(a, tile, st, cnt, pol_x, pol_gw);
        }
        return;
    }

    const int half = warp % G::HALVES, slot = warp / G::HALVES;
    const int lane_px = half * (32 * PPL) + lane * PPL;
    const TC ls = (TC)a.log_shift;
    const TC* Hn = reinterpret_cast<const TC*>(a.H);
    TC* my_acc = sacc + (size_t)half * a.sacc_rows * KP;
    uint32_t cnt = 0;

    for (int tile = t0; tile < t1; ++tile) {
        TC h[KP][PPL];
        load_h<TC, KP, PPL>(h, Hn, a.ldh, a.k, tile * TILE_PX + lane_px);
        for (int st = sb0; st < sb1; ++st, ++cnt) {
            ring.consumer_wait(cnt);
            const unsigned char* xs = ring.x_ptr(cnt % a.depth);
            const unsigned char* gs = ring.gw_ptr(cnt % a.depth);
#pragma unroll
            for (int ci = 0; ci < G::CPW; ++ci) {
                const int c = slot + ci * G::NSLOT;
                TX xv[PPL];
                TC gw[KP];
                lds_vec<TX, PPL>(xv, xs + ((size_t)c * TILE_PX + lane_px) * sizeof(TX));
                lds_gw<TC, KP>(gw, gs + (size_t)c * KP * sizeof(TC));
                TC t[KP];
#pragma unroll
                for (int kk = 0; kk < KP; ++kk) t[kk] = TC(0);
#pragma unroll
                for (int q = 0; q < PPL; ++q) {
                    TC y = gw[0] * h[0][q];
#pragma unroll
                    for (int kk = 1; kk < KP; ++kk) y = fma(gw[kk], h[kk][q], y);
                    if constexpr (SAFE) {
                        if (a.clamp_y) y = Num<TC>::vmax(y, ls);
                    }
                    const TC rq = Num<TC>::ratio((TC)xv[q], y);
#pragma unroll
                    for (int kk = 0; kk < KP; ++kk) t[kk] = fma(rq, h[kk][q], t[kk]);
                }
                // reduce over the 32 lanes (pixels) and accumulate into this warp's private rows
                TC mine = TC(0);
#pragma unroll
                for (int kk = 0; kk < KP; ++kk) {
                    const TC s = warp_sum(t[kk]);
                    if (lane == kk) mine = s;
                }
                if (lane < KP) {
                    TC* dst = my_acc + ((size_t)(st - sb0) * G::CS + c) * KP + lane;
                    *dst += mine;
                }
            }
            ring.consumer_release(cnt, lane);
        }
    }
    named_bar_sync(1, N_CONSUMER_THREADS);
    TC* out = reinterpret_cast<TC*>(a.s_part) + ((size_t)r * a.n_pad + (size_t)sb0 * G::CS) * KP;
    for (int idx = threadIdx.x; idx < rows * KP; idx += N_CONSUMER_THREADS) {
        TC s = sacc[idx];
        if constexpr (G::HALVES == 2) s += sacc[(size_t)a.sacc_rows * KP + idx];
        out[idx] = s;
    }
}

}  // namespace espm
