// Development aid: dependent-issue latency of DFMA / DADD / DMUL / SHFL / MUFU.RSQ64H on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency tools/fp64_latency.cu && /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b, int iters) {
    double r = a;
    const long long c0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) r = fma(r, a, b);
            if (OP == 1) r = r + b;
            if (OP == 2) r = r * a;
            if (OP == 3) r = __shfl_sync(0xffffffffu, r, (threadIdx.x + 1) & 31);
            if (OP == 4) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r)); r = y; }
            if (OP == 5) r = 1.0 / r;
            if (OP == 6) r = sqrt(r);
            if (OP == 7) r = (double)(float)r;
        }
    }
    const long long c1 = clock64();
    out[threadIdx.x] = r;
    if (threadIdx.x == 0) *cyc = c1 - c0;
}

int main() {
    double* d; long long* c; cudaMalloc(&d, 8 * 32); cudaMalloc(&c, 8);
    const char* names[] = {"DFMA", "DADD", "DMUL", "SHFL(64b=2x32)", "MUFU.RSQ64H", "ddiv (1/x)", "dsqrt", "F2F f64->f32->f64"};
    const int iters = 2000;
    auto run = [&](int op) {
        switch (op) {
            case 0: chain<0><<<1, 32>>>(d, c, 0.999999, 1e-9, iters); break;
            case 1: chain<1><<<1, 32>>>(d, c, 0.999999, 1e-9, iters); break;
            case 2: chain<2><<<1, 32>>>(d, c, 0.999999, 1e-9, iters); break;
            case 3: chain<3><<<1, 32>>>(d, c, 0.999999, 1e-9, iters); break;
            case 4: chain<4><<<1, 32>>>(d, c, 1.7, 1e-9, iters); break;
            case 5: chain<5><<<1, 32>>>(d, c, 1.7, 1e-9, iters); break;
            case 6: chain<6><<<1, 32>>>(d, c, 1.7, 1e-9, iters); break;
            case 7: chain<7><<<1, 32>>>(d, c, 1.7, 1e-9, iters); break;
        }
    };
    for (int op = 0; op < 8; ++op) {
        run(op); run(op);
        cudaDeviceSynchronize();
        long long cyc; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
        printf("%-20s %7.1f cycles per dependent op\n", names[op], (double)cyc / (iters * 16.0));
    }
    return 0;
}
