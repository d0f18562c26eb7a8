// Development probe: is the FP64 tensor-core path (DMMA) on B200 a pipe separate from the FP64 vector pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe tools/dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int MODE>   // 0: DFMA only, 1: DMMA m8n8k4 only, 2: both in every warp, 3: DMMA m16n8k8 only, 4: DFMA + m16n8k8
__global__ void __launch_bounds__(256) probe(double* out, int iters, double seed) {
    double f[8], c[8][2], c4[4][4];
    const double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed;
    double a4[4] = {a, a + 1, a + 2, a + 3}, b2[2] = {b, b + 1};
#pragma unroll
    for (int k = 0; k < 8; ++k) { f[k] = k + seed; c[k][0] = c[k][1] = k; }
#pragma unroll
    for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) c4[k][j] = k + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0 || MODE == 2 || MODE == 4) f[k] = fma(f[k], a, b);
            if (MODE == 1 || MODE == 2) dmma884(c[k][0], c[k][1], a, b);
        }
        if (MODE == 3 || MODE == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) dmma1688(c4[k], a4, b2);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k] + c[k][0] + c[k][1];
#pragma unroll
    for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) s += c4[k][j];
    if (s == 12345.678) out[0] = s;
}

template <int MODE>
void run(const char* name, double dfma_per_iter, double dmma_flop_per_iter_per_warp) {
    double* out; cudaMalloc(&out, 8);
    const int iters = 20000, grid = 148 * 8, block = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<grid, block>>>(out, 100, 0.5);
    cudaEventRecord(e0);
    probe<MODE><<<grid, block>>>(out, iters, 0.5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double threads = (double)grid * block, warps = threads / 32;
    const double vec = 2.0 * dfma_per_iter * iters * threads / (ms * 1e-3) / 1e12;
    const double ten = dmma_flop_per_iter_per_warp * iters * warps / (ms * 1e-3) / 1e12;
    printf("%-28s %8.3f ms   vector %6.2f TF   tensor %6.2f TF   sum %6.2f TF\n", name, ms, vec, ten, vec + ten);
    cudaFree(out);
}

int main() {
    run<0>("DFMA only", 8, 0);
    run<1>("DMMA m8n8k4 only", 0, 8 * 2.0 * 8 * 8 * 4);
    run<2>("DFMA + DMMA m8n8k4", 8, 8 * 2.0 * 8 * 8 * 4);
    run<3>("DMMA m16n8k8 only", 0, 4 * 2.0 * 16 * 8 * 8);
    run<4>("DFMA + DMMA m16n8k8", 8, 4 * 2.0 * 16 * 8 * 8);
    return 0;
}
