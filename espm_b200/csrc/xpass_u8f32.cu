#include "xpass_inst.cuh"
namespace espm {
int xpass_u8f32(const XPassLaunch& l, const XPassArgs* a, int* occ_out, cudaStream_t stream) {
    return xpass_entry<uint8_t, float>(l, a, occ_out, stream);
}
}  // namespace espm
