// Probe: accuracy of MUFU.RCP64H seeds and of Newton / Halley refinements (decides the fp64 reciprocal of common.cuh).
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k(const double* y, int n, double* seed_err, double* newton2_err, double* halley_err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = y[i], r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
    double exact = 1.0 / v;
    seed_err[i] = fabs(r - exact) / exact;
    double e = fma(-v, r, 1.0);
    double r2 = fma(r, e, r);
    e = fma(-v, r2, 1.0);
    r2 = fma(r2, e, r2);
    newton2_err[i] = fabs(r2 - exact) / exact;
    e = fma(-v, r, 1.0);
    double t = fma(e, e, e);
    double r3 = fma(r, t, r);
    halley_err[i] = fabs(r3 - exact) / exact;
}
int main() {
    const int n = 1 << 24;
    double* h = (double*)malloc(n * sizeof(double));
    srand(1);
    for (int i = 0; i < n; ++i) {
        double u = (rand() + 0.5) / ((double)RAND_MAX + 1), w = (rand() + 0.5) / ((double)RAND_MAX + 1);
        h[i] = (1.0 + u + w * 1e-9) * exp2((double)(rand() % 600 - 300));
    }
    double *d, *a, *b, *c;
    cudaMalloc(&d, n * 8); cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8); cudaMalloc(&c, n * 8);
    cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(d, n, a, b, c);
    double *ha = (double*)malloc(n * 8), *hb = (double*)malloc(n * 8), *hc = (double*)malloc(n * 8);
    cudaMemcpy(ha, a, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hb, b, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, c, n * 8, cudaMemcpyDeviceToHost);
    double ma = 0, mb = 0, mc = 0;
    for (int i = 0; i < n; ++i) { ma = fmax(ma, ha[i]); mb = fmax(mb, hb[i]); mc = fmax(mc, hc[i]); }
    printf("max rel err: seed %.3e (2^%.1f)  newton2 %.3e  halley %.3e  (%s)\n", ma, log2(ma), mb, mc, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
