#include "small_inst.cuh"
namespace espm {
int small_launch_f64(int op, const espm_state* st, cudaStream_t s) { return small_launch_t<double>(op, st, s); }
int dicho_launch_f64(const DichoArgs& d, cudaStream_t s) { return dicho_launch_t<double>(d, s); }
int colsum_launch_f64(const espm_state* st, void* out, cudaStream_t s) { return colsum_launch_t<double>(st, out, s); }
}  // namespace espm
