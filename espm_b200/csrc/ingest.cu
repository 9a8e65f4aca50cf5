// X ingest kernels (see ingest.cuh) and their launchers.
#include "ingest.cuh"

namespace espm {

template <typename TS, typename TX>
static int retile_launch_t(const espm_state* st, const void* src, long long stride_c, long long stride_p,
                           long long j0, double scale, const IngestOut& io, cudaStream_t s) {
    dim3 grid(st->n_tiles, st->n_pad / 32);
    retile_kernel<TS, TX><<<grid, 256, 0, s>>>((const TS*)src, stride_c, stride_p, j0, st->n, st->n_pad, st->p_loc,
                                              scale, (TX*)st->Xt, io);
    ESPM_CUDA_CHECK(cudaGetLastError());
    return ESPM_OK;
}

int retile_launch(const espm_state* st, const void* src, int src_dtype, long long stride_c, long long stride_p,
                  long long j0, double scale, const IngestOut& io, cudaStream_t s) {
    if (src_dtype == ESPM_F32 && st->x_dtype == ESPM_F32)
        return retile_launch_t<float, float>(st, src, stride_c, stride_p, j0, scale, io, s);
    if (src_dtype == ESPM_F64 && st->x_dtype == ESPM_F64)
        return retile_launch_t<double, double>(st, src, stride_c, stride_p, j0, scale, io, s);
    if (src_dtype == ESPM_F64 && st->x_dtype == ESPM_F32)
        return retile_launch_t<double, float>(st, src, stride_c, stride_p, j0, scale, io, s);
    if (src_dtype == ESPM_F32 && st->x_dtype == ESPM_F64)
        return retile_launch_t<float, double>(st, src, stride_c, stride_p, j0, scale, io, s);
    // compact count storage: the caller has established (espm_x_prescan) that every entry is an integer in range
    if (st->x_dtype == ESPM_U8 || st->x_dtype == ESPM_U16) {
        if (scale != 1.0) {
            set_error("retile: compact storage cannot be scaled (scale=%g)", scale);
            return ESPM_ERR_BAD_ARG;
        }
        if (src_dtype == ESPM_F32 && st->x_dtype == ESPM_U8)
            return retile_launch_t<float, uint8_t>(st, src, stride_c, stride_p, j0, scale, io, s);
        if (src_dtype == ESPM_F32 && st->x_dtype == ESPM_U16)
            return retile_launch_t<float, uint16_t>(st, src, stride_c, stride_p, j0, scale, io, s);
        if (src_dtype == ESPM_F64 && st->x_dtype == ESPM_U8)
            return retile_launch_t<double, uint8_t>(st, src, stride_c, stride_p, j0, scale, io, s);
        if (src_dtype == ESPM_F64 && st->x_dtype == ESPM_U16)
            return retile_launch_t<double, uint16_t>(st, src, stride_c, stride_p, j0, scale, io, s);
    }
    set_error("retile: unsupported dtype combination %d -> %d", src_dtype, st->x_dtype);
    return ESPM_ERR_BAD_ARG;
}

int prescan_launch(const void* src, int src_dtype, int n, long long p_loc, long long stride_c, long long stride_p,
                   long long j0, uint32_t* out4, int32_t* row_nz, int32_t* col_nz, cudaStream_t s) {
    const int px_fast = (stride_p == 1 || stride_c != 1) ? 1 : 0;     // which axis is contiguous (see retile_kernel)
    const long long n_fast = px_fast ? p_loc : n, n_slow = px_fast ? n : p_loc;
    dim3 grid((unsigned)((n_fast + 255) / 256), (unsigned)((n_slow + 63) / 64));
    if (src_dtype == ESPM_F32)
        prescan_kernel<float><<<grid, 256, 0, s>>>((const float*)src, stride_c, stride_p, j0, n, p_loc, px_fast, out4,
                                                   row_nz, col_nz);
    else
        prescan_kernel<double><<<grid, 256, 0, s>>>((const double*)src, stride_c, stride_p, j0, n, p_loc, px_fast, out4,
                                                    row_nz, col_nz);
    ESPM_CUDA_CHECK(cudaGetLastError());
    return ESPM_OK;
}

int xt_fixup_launch(const espm_state* st, const int32_t* row_zero, const int32_t* col_zero, double eps, double scale,
                    cudaStream_t s) {
    if (st->x_dtype != ESPM_F32 && st->x_dtype != ESPM_F64) {
        set_error("xt_fixup: compact storage holds unpatched integer counts only");
        return ESPM_ERR_UNSUPPORTED;
    }
    dim3 grid(st->n_tiles, st->n_pad / 32);
    if (st->x_dtype == ESPM_F32)
        xt_fixup_kernel<float><<<grid, 256, 0, s>>>((float*)st->Xt, row_zero, col_zero, st->n, st->n_pad, st->p_loc, eps,
                                                   scale);
    else
        xt_fixup_kernel<double><<<grid, 256, 0, s>>>((double*)st->Xt, row_zero, col_zero, st->n, st->n_pad, st->p_loc,
                                                    eps, scale);
    ESPM_CUDA_CHECK(cudaGetLastError());
    return ESPM_OK;
}

int xt_const_launch(const espm_state* st, double* part, cudaStream_t s) {
    if (st->x_dtype == ESPM_F32)
        xt_const_kernel<float><<<st->n_tiles, 256, 0, s>>>((const float*)st->Xt, st->n, st->n_pad, st->p_loc,
                                                          st->log_shift, part);
    else if (st->x_dtype == ESPM_U8)
        xt_const_kernel<uint8_t><<<st->n_tiles, 256, 0, s>>>((const uint8_t*)st->Xt, st->n, st->n_pad, st->p_loc,
                                                            st->log_shift, part);
    else if (st->x_dtype == ESPM_U16)
        xt_const_kernel<uint16_t><<<st->n_tiles, 256, 0, s>>>((const uint16_t*)st->Xt, st->n, st->n_pad, st->p_loc,
                                                             st->log_shift, part);
    else
        xt_const_kernel<double><<<st->n_tiles, 256, 0, s>>>((const double*)st->Xt, st->n, st->n_pad, st->p_loc,
                                                           st->log_shift, part);
    ESPM_CUDA_CHECK(cudaGetLastError());
    return ESPM_OK;
}

int x_sums_launch(const espm_state* st, void* colsum, double* rowsum_part, cudaStream_t s) {
    if (st->x_dtype == ESPM_U8)
        x_sums_kernel<uint8_t, float><<<st->n_tiles, 128, 0, s>>>((const uint8_t*)st->Xt, st->n_pad, (float*)colsum, rowsum_part);
    else if (st->x_dtype == ESPM_U16)
        x_sums_kernel<uint16_t, float><<<st->n_tiles, 128, 0, s>>>((const uint16_t*)st->Xt, st->n_pad, (float*)colsum, rowsum_part);
    else if (st->x_dtype == ESPM_F32 && st->c_dtype == ESPM_F32)
        x_sums_kernel<float, float><<<st->n_tiles, 128, 0, s>>>((const float*)st->Xt, st->n_pad, (float*)colsum, rowsum_part);
    else if (st->x_dtype == ESPM_F32)
        x_sums_kernel<float, double><<<st->n_tiles, 128, 0, s>>>((const float*)st->Xt, st->n_pad, (double*)colsum, rowsum_part);
    else
        x_sums_kernel<double, double><<<st->n_tiles, 128, 0, s>>>((const double*)st->Xt, st->n_pad, (double*)colsum, rowsum_part);
    ESPM_CUDA_CHECK(cudaGetLastError());
    return ESPM_OK;
}

// out[i] = log2_tab(y[i]): the table-driven log2 of the fp64 H pass, exposed for the accuracy test
__global__ void __launch_bounds__(256) log2_table_kernel(const double* y, long long n, double* out) {
    __shared__ double2 tab[LOG2TAB_N * LOG2TAB_R];
    log2tab_fill(tab, threadIdx.x, blockDim.x);
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = log2_tab(y[i], log2tab_lane(tab, threadIdx.x & 31));
}

int log2_table_launch(const double* y, long long n, double* out, cudaStream_t s) {
    const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    if (grid > 0) log2_table_kernel<<<grid, 256, 0, s>>>(y, n, out);
    ESPM_CUDA_CHECK(cudaGetLastError());
    return ESPM_OK;
}

int reduce_sum_launch(const double* in, long long n, double* out, cudaStream_t s) {
    reduce_sum_kernel<<<1, 1024, 0, s>>>(in, n, out);
    ESPM_CUDA_CHECK(cudaGetLastError());
    return ESPM_OK;
}

}  // namespace espm
