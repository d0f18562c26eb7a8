// Development probe 2: accumulate-type DFMAs  acc_k = fma(s, d_k, acc_k)  with different operand sharing.
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void __launch_bounds__(128) probe(double* out, int iters, double a, double b) {
    double acc[12], d[12], s[4];
#pragma unroll
    for (int k = 0; k < 12; ++k) { acc[k] = (threadIdx.x + k) * 1e-3; d[k] = a + k * b + threadIdx.x * 1e-9; }
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = b * (k + 1) + threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                if (V == 0) acc[k] = fma(s[0], d[k], acc[k]);            // one s shared by all
                if (V == 1) acc[k] = fma(s[k / 3], d[k], acc[k]);        // groups of 3 share s
                if (V == 2) acc[k] = fma(s[k & 1], d[k >> 1], acc[k]);   // pairs share d
                if (V == 3) acc[k] = fma(s[k & 3], d[(k * 5) % 12], acc[k]);   // nothing shared between neighbours
                if (V == 4) acc[k] = fma(s[k / 6], d[k], acc[k]);        // groups of 6 share s
            }
        }
    }
    double t = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) t += acc[k];
    if (t == 1.2345) out[0] = t;
}

template <int V>
void run(const char* name, int ctas_per_sm) {
    double* out; cudaMalloc(&out, 8);
    const int iters = 8192, grid = 148 * ctas_per_sm, block = 128;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<V><<<grid, block>>>(out, 64, 0.999999, 1e-9);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<V><<<grid, block>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double rate = (double)grid * block * iters * 48.0 / (best * 1e-3);
    printf("%-44s warps/SMSP=%d  %7.3f ms  %5.1f %% of 148x64x1.965GHz  (%.2f cycles/DFMA)\n", name, ctas_per_sm, best,
           100.0 * rate / (148.0 * 64 * 1.965e9), 2.0 / (rate / (148.0 * 64 * 1.965e9)));
    cudaFree(out);
}

int main() {
    for (int w : {1, 2, 4}) {
        run<0>("acc_k = fma(s, d_k, acc_k)  s shared by all", w);
        run<4>("groups of 6 share s", w);
        run<1>("groups of 3 share s", w);
        run<2>("pairs share d", w);
        run<3>("nothing shared", w);
    }
    return 0;
}
