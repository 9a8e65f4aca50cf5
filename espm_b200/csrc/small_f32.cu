#include "small_inst.cuh"
namespace espm {
int small_launch_f32(int op, const espm_state* st, cudaStream_t s) { return small_launch_t<float>(op, st, s); }
int dicho_launch_f32(const DichoArgs& d, cudaStream_t s) { return dicho_launch_t<float>(d, s); }
int colsum_launch_f32(const espm_state* st, void* out, cudaStream_t s) { return colsum_launch_t<float>(st, out, s); }
}  // namespace espm
