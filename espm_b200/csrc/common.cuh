// Shared device helpers for the espm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include <cstdlib>
#include <utility>

#include "../../include/espm_b200.h"

#ifndef __CUDA_ARCH__
#define ESPM_DEVICE_ARCH 0
#else
#define ESPM_DEVICE_ARCH __CUDA_ARCH__
#endif

namespace espm {

constexpr int TILE_PX = ESPM_TILE_PX;        // pixels per tile
constexpr int STAGE_BYTES = ESPM_STAGE_BYTES; // X bytes per pipeline stage
constexpr int N_CONSUMER_WARPS = 8;
constexpr int N_CONSUMER_THREADS = N_CONSUMER_WARPS * 32;
constexpr int XPASS_THREADS = N_CONSUMER_THREADS + 32;  // + one TMA producer warp

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
#define ESPM_CUDA_CHECK(expr)                                                               \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            espm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                      \
            return ESPM_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

// ---------------------------------------------------------------- mbarrier / bulk-copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared (contiguous bytes, 16 B aligned, size % 16 == 0); completion is
// signalled on `bar` through complete_tx.  SASS: UBLKCP.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar, uint64_t l2_policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(l2_policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- programmatic dependent launch
// The per-iteration kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel's CTAs
// may be scheduled while the previous kernel of the stream drains.  pdl_wait() (griddepcontrol.wait) is the FIRST
// statement of every such kernel -- nothing of the predecessor is read or overwritten before it returns, so the only
// effect is that launch latency and CTA start-up overlap the predecessor's tail.  pdl_trigger() lets the next kernel
// of the stream do the same with us.  Both are no-ops for kernels launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    static const bool enabled = [] {   // ESPM_B200_PDL=0 restores plain stream-ordered launches (A/B measurements)
        const char* e = getenv("ESPM_B200_PDL");
        return !(e && e[0] == '0');
    }();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = enabled ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------- arithmetic traits
template <typename T>
struct Num;
template <>
struct Num<float> {
    // x / y for the streaming passes: MUFU.RCP + FMUL (<= 2 ulp; the fp32 mode's bar is 1e-5)
    static __device__ __forceinline__ float ratio(float x, float y) { return __fdividef(x, y); }
    static __device__ __forceinline__ float log2_fast(float y) { return __log2f(y); }
    static __device__ __forceinline__ float vmax(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float vmin(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float vabs(float a) { return fabsf(a); }
    static __device__ __forceinline__ float vsqrt(float a) { return sqrtf(a); }
    static __device__ __forceinline__ float vlog(float a) { return logf(a); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};
template <>
struct Num<double> {
    // fp64 reciprocal: MUFU.RCP64H seed (rcp.approx.ftz.f64, ~20 good bits, full exponent range) + two
    // Newton steps in fp64 => <= 1 ulp.  y == 0 gives NaN/inf like the division would (callers flag
    // non-finite results); 7 instructions instead of the ~25 of an IEEE fp64 division.
    static __device__ __forceinline__ double rcp(double y) {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
        double e = fma(-y, r, 1.0);
        r = fma(r, e, r);
        e = fma(-y, r, 1.0);
        r = fma(r, e, r);
        return r;
    }
    static __device__ __forceinline__ double ratio(double x, double y) { return x * rcp(y); }
    static __device__ __forceinline__ double log2_fast(double y) { return log2(y); }
    static __device__ __forceinline__ double vmax(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double vmin(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double vabs(double a) { return fabs(a); }
    static __device__ __forceinline__ double vsqrt(double a) { return sqrt(a); }
    static __device__ __forceinline__ double vlog(double a) { return log(a); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = v > u ? v : u;
    }
    return v;
}

// Geometry of the X-pass kernels for a (storage, compute) type pair.
template <typename TX, typename TC>
struct PassGeom {
    static constexpr int PPL = (sizeof(TC) == 8) ? 2 : 4;          // pixels per lane
    static constexpr int HALVES = 4 / PPL;                         // warps covering one 128-px row
    static constexpr int NSLOT = N_CONSUMER_WARPS / HALVES;        // channel rows processed at once
    static constexpr int CS = STAGE_BYTES / (TILE_PX * (int)sizeof(TX));  // channels per stage
    static constexpr int CPW = CS / NSLOT;                         // channels per warp per stage
};

// host-side helpers implemented in api.cu
int plan_stage_channels(int x_dtype);

}  // namespace espm
