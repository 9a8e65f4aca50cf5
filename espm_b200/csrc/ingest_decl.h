// Host-side declarations of the X-ingest launchers (ingest.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/espm_b200.h"

namespace espm {
struct IngestOut {
    int32_t* row_nz;
    int32_t* col_nz;
    uint32_t* flags;
    double* sum_part;
};
int retile_launch(const espm_state* st, const void* src, int src_dtype, long long stride_c, long long stride_p,
                  long long j0, double scale, const IngestOut& io, cudaStream_t s);
int prescan_launch(const void* src, int src_dtype, int n, long long p_loc, long long stride_c, long long stride_p,
                   long long j0, uint32_t* out4, int32_t* row_nz, int32_t* col_nz, cudaStream_t s);
int xt_fixup_launch(const espm_state* st, const int32_t* row_zero, const int32_t* col_zero, double eps, double scale,
                    cudaStream_t s);
int xt_const_launch(const espm_state* st, double* part, cudaStream_t s);
int x_sums_launch(const espm_state* st, void* colsum, double* rowsum_part, cudaStream_t s);
int log2_table_launch(const double* y, long long n, double* out, cudaStream_t s);
int reduce_sum_launch(const double* in, long long n, double* out, cudaStream_t s);
}  // namespace espm
