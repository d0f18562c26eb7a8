// Development probe: which DFMA operand patterns / occupancies reach which fraction of the nominal FP64 rate
// (148 SMs x 64 lanes x 2 flop x clock) on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/dfma_probe tools/dfma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int V, int CH>
__global__ void __launch_bounds__(128) probe(double* out, int iters, double a, double b) {
    double r[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) r[k] = (threadIdx.x + k) * 1e-3;
    double a2 = a * 1.0000001, b2 = b * 0.999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 64 / CH; ++u) {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                if (V == 0) r[k] = fma(r[k], a, b);
                if (V == 1) r[k] = fma(r[k], r[k], r[k]);
                if (V == 2) r[k] = fma(r[k], a, r[k]);
                if (V == 3) r[k] = fma(a, b, r[k]);
                if (V == 4) r[k] = r[k] + a;
                if (V == 5) r[k] = r[k] * a;
                if (V == 6) r[k] = fma(r[k], (k & 1) ? a : a2, (k & 2) ? b : b2);
                if (V == 7) r[k] = fma(r[k], r[(k + 1) % CH], r[(k + 2) % CH]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += r[k];
    if (s == 1.2345) out[0] = s;
}

template <int V, int CH>
void run(const char* name, int ctas_per_sm) {
    double* out; cudaMalloc(&out, 8);
    const int iters = 8192, grid = 148 * ctas_per_sm, block = 128;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<V, CH><<<grid, block>>>(out, 64, 0.999999, 1e-9);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<V, CH><<<grid, block>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)grid * block * iters * 64.0;
    const double rate = ops / (best * 1e-3);                       // lane-instructions per second
    printf("%-34s chains=%2d warps/SMSP=%d  %7.3f ms  %6.2f TF(2 flop)  %5.1f %% of 148x64x1.965GHz\n", name, CH,
           ctas_per_sm, best, 2 * rate / 1e12, 100.0 * rate / (148.0 * 64 * 1.965e9));
    cudaFree(out);
}

int main() {
    run<0, 8>("fma(r,a,b)", 1); run<0, 8>("fma(r,a,b)", 2); run<0, 8>("fma(r,a,b)", 4); run<0, 8>("fma(r,a,b)", 8);
    run<0, 16>("fma(r,a,b)", 2); run<0, 4>("fma(r,a,b)", 8);
    run<1, 8>("fma(r,r,r)", 4); run<2, 8>("fma(r,a,r)", 4); run<3, 8>("fma(a,b,r)", 4);
    run<4, 8>("r+a (DADD)", 4); run<5, 8>("r*a (DMUL)", 4);
    run<6, 8>("fma(r,a|a2,b|b2)", 4); run<7, 8>("fma(r_k,r_k+1,r_k+2)", 4); run<7, 16>("fma(r_k,r_k+1,r_k+2)", 2);
    return 0;
}
