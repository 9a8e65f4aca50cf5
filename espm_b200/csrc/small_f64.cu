#include "small_inst.cuh"
namespace espm {
int small_launch_f64(int op, const espm_state* st, cudaStream_t s) { return small_launch_t<double>(op, st, s); }
int dicho_launch_f64(const DichoArgs& d, cudaStream_t s) { return dicho_launch_t<double>(d, s); }
int colsum_launch_f64(const espm_state* st, void* out, cudaStream_t s) { return colsum_launch_t<double>(st, out, s); }
int retile_launch(const espm_state* st, const void* src, int src_dtype, long long stride_c, long long stride_p,
                  long long j0, double scale, cudaStream_t s) {
    if (src_dtype == ESPM_F32 && st->x_dtype == ESPM_F32)
        return retile_launch_t<float, float>(st, src, stride_c, stride_p, j0, scale, s);
    if (src_dtype == ESPM_F64 && st->x_dtype == ESPM_F64)
        return retile_launch_t<double, double>(st, src, stride_c, stride_p, j0, scale, s);
    if (src_dtype == ESPM_F64 && st->x_dtype == ESPM_F32)
        return retile_launch_t<double, float>(st, src, stride_c, stride_p, j0, scale, s);
    if (src_dtype == ESPM_F32 && st->x_dtype == ESPM_F64)
        return retile_launch_t<float, double>(st, src, stride_c, stride_p, j0, scale, s);
    set_error("retile: unsupported dtype combination %d -> %d", src_dtype, st->x_dtype);
    return ESPM_ERR_BAD_ARG;
}
}  // namespace espm
