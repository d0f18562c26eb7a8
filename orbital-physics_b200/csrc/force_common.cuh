// force_common.cuh -- device helpers shared by the one-sided (force.cu) and pair-symmetric
// (force_sym.cu) fast force kernels.
#pragma once
#include "kernels.h"

namespace orb {

constexpr int kFastThreads = 128;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kTile = 256;     // source bodies per TMA tile (8 KiB)
#ifndef ORB_STAGES
#define ORB_STAGES 4
#endif
constexpr int kStages = ORB_STAGES;

// m_j * (r^2)^(-3/2) to ~1 ulp from the 20-bit MUFU seed:
//   y0 = rsqrt(r2)(1+d), e = 1 - r2*y0^2,  (r2)^(-3/2) = y0^3 (1-e)^(-3/2)
//   (1-e)^(-3/2) = 1 + 3/2 e + 15/8 e^2 + O(e^3),  |e| <~ 2^-19  => O(e^3) < 1e-17
__device__ __forceinline__ double inv_r3_mass(double r2, double mj, int& y0_hi) {
    const double y0 = rsqrt_seed(r2);
    y0_hi = __double2hiint(y0);
    const double u = y0 * y0;                 // exact: y0 has <= 21 significant bits
    const double e = fma(-r2, u, 1.0);
    const double c = mj * y0;
    const double w = c * u;
    const double p = fma(1.875, e, 1.5);
    const double q = e * p;
    return fma(w, q, w);
}

// y0^3 (1-e)^(-3/2) without the mass factor: 6 FP64 instructions
__device__ __forceinline__ double inv_r3_plain(double r2, int& y0_hi) {
    const double y0 = rsqrt_seed(r2);
    y0_hi = __double2hiint(y0);
    const double u = y0 * y0;
    const double e = fma(-r2, u, 1.0);
    const double w = y0 * u;
    const double p = fma(1.875, e, 1.5);
    const double q = e * p;
    return fma(w, q, w);
}

// UNI: all masses are equal -- the mass factor is applied once, by the reduction kernel
template <int TI, bool DETECT, bool CHECKED, bool UNI = false>
__device__ __forceinline__ void tile_loop(const double2* __restrict__ tile, int cnt, long long j0, double eps2,
                                          const double (&xi)[TI], const double (&yi)[TI], const double (&zi)[TI],
                                          const long long (&idx)[TI], double (&ax)[TI], double (&ay)[TI],
                                          double (&az)[TI], int (&maxhi)[TI]) {
#pragma unroll 2
    for (int j = 0; j < cnt; ++j) {
        const double2 a = tile[2 * j];
        const double2 b = tile[2 * j + 1];
#pragma unroll
        for (int k = 0; k < TI; ++k) {
            const double dx = a.x - xi[k];
            const double dy = a.y - yi[k];
            const double dz = b.x - zi[k];
            const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
            int hi;
            double s = UNI ? inv_r3_plain(r2, hi) : inv_r3_mass(r2, b.y, hi);
            if (CHECKED) {
                const bool self = (j0 + j) == idx[k];
                s = self ? 0.0 : s;
                hi = self ? 0 : hi;
            }
            if (DETECT) maxhi[k] = max(maxhi[k], hi);
            ax[k] = fma(s, dx, ax[k]);
            ay[k] = fma(s, dy, ay[k]);
            az[k] = fma(s, dz, az[k]);
        }
    }
}

// Rare path: a seed exceeded the conservative threshold -- test the tile exactly.
static __device__ __noinline__ void rescan_tile(const double2* tile, int cnt, long long j0, long long i, double xi,
                                         double yi, double zi, double Ri, const double* __restrict__ radius,
                                         Ctl* ctl, long long* pairs) {
    for (int j = 0; j < cnt; ++j) {
        const long long jg = j0 + j;
        if (jg <= i) continue;                      // each unordered pair once (i < j)
        const double2 a = tile[2 * j];
        const double2 b = tile[2 * j + 1];
        // handle_collisions forms ri - rj (physics.py:517); squares are sign-independent
        if (overlap_exact(xi - a.x, yi - a.y, zi - b.x, Ri, radius[jg])) record_overlap(ctl, pairs, i, jg);
    }
}


}  // namespace orb
