// Explicit instantiation helpers for the X-pass kernels: one translation unit per (storage, compute)
// type pair so that the 48 template instances compile in parallel.
#pragma once
#include "xpass.cuh"

namespace espm {

enum XPassKind { XPASS_H = 0, XPASS_W = 1 };

struct XPassLaunch {
    int kind;   // XPassKind
    int kp;
    int safe;
    int grid;
    int smem;
};

template <typename TX, typename TC, int KP, bool SAFE>
static int xpass_do(const XPassLaunch& l, const XPassArgs* a, int* occ_out, cudaStream_t stream) {
    auto kern = (l.kind == XPASS_H) ? h_pass_kernel<TX, TC, KP, SAFE> : w_pass_kernel<TX, TC, KP, SAFE>;
    ESPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, l.smem));
    if (occ_out) {
        ESPM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_out, kern, XPASS_THREADS, l.smem));
        return ESPM_OK;
    }
    kern<<<l.grid, XPASS_THREADS, l.smem, stream>>>(*a);
    ESPM_CUDA_CHECK(cudaGetLastError());
    return ESPM_OK;
}

template <typename TX, typename TC, bool SAFE>
static int xpass_by_kp(const XPassLaunch& l, const XPassArgs* a, int* occ_out, cudaStream_t stream) {
    switch (l.kp) {
        case 2: return xpass_do<TX, TC, 2, SAFE>(l, a, occ_out, stream);
        case 3: return xpass_do<TX, TC, 3, SAFE>(l, a, occ_out, stream);
        case 4: return xpass_do<TX, TC, 4, SAFE>(l, a, occ_out, stream);
        case 5: return xpass_do<TX, TC, 5, SAFE>(l, a, occ_out, stream);
        case 6: return xpass_do<TX, TC, 6, SAFE>(l, a, occ_out, stream);
        case 8: return xpass_do<TX, TC, 8, SAFE>(l, a, occ_out, stream);
        case 12: return xpass_do<TX, TC, 12, SAFE>(l, a, occ_out, stream);
        case 16: return xpass_do<TX, TC, 16, SAFE>(l, a, occ_out, stream);
        default:
            set_error("unsupported padded component count kp=%d", l.kp);
            return ESPM_ERR_BAD_ARG;
    }
}

template <typename TX, typename TC>
static int xpass_entry(const XPassLaunch& l, const XPassArgs* a, int* occ_out, cudaStream_t stream) {
    return l.safe ? xpass_by_kp<TX, TC, true>(l, a, occ_out, stream) : xpass_by_kp<TX, TC, false>(l, a, occ_out, stream);
}

// sizes needed by the planner (host)
template <typename TX, typename TC>
static void xpass_sizes(int kp, int safe, int* stage_stride, int* red_bytes, int* fixed_bytes, int* cs, int* halves) {
    using G = PassGeom<TX, TC>;
    const int gw_bytes = G::CS * kp * (int)sizeof(TC);
    const int gw_al = (gw_bytes + 127) / 128 * 128;
    *stage_stride = STAGE_BYTES + gw_al * (safe ? 2 : 1);
    *red_bytes = G::NSLOT * kp * TILE_PX * (int)sizeof(TC);
    *fixed_bytes = 256 + 128;
    *cs = G::CS;
    *halves = G::HALVES;
}

// defined in xpass_f32f32.cu / xpass_f32f64.cu / xpass_f64f64.cu
int xpass_f32f32(const XPassLaunch& l, const XPassArgs* a, int* occ_out, cudaStream_t stream);
int xpass_f32f64(const XPassLaunch& l, const XPassArgs* a, int* occ_out, cudaStream_t stream);
int xpass_f64f64(const XPassLaunch& l, const XPassArgs* a, int* occ_out, cudaStream_t stream);

}  // namespace espm
