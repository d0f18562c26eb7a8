// ensemble.cu -- batched ensemble of independent small systems, one CTA per system (sm_100a).
//
// Equivalent to `nsys` separate reference SimulationEngine instances
// (core/engine.py:19-46,65-97; force: core/physics.py:125-159) stepped in
// lockstep, without collision handling. State is SoA [nsys][nbody] fp64:
// x y z vx vy vz ax ay az (read+write) and m (read): 152 B per body-step when a
// launch covers one step (HBM-bound mode); with `nsteps` fused in one launch the
// state stays in registers and the kernel is FP64-bound.
//
//   FAITHFUL: lane i owns body i and accumulates j ascending with the reference's
//             rounding sequence (bit-exact with the reference engine).
//   FAST:     ens_step_fast_kernel -- every unordered pair once (the reference's own half-matrix trick,
//             physics.py:136-155), 20 FP64 instructions per pair.  A lane owns TWO bodies (I and I + nbp/2), so
//             a system takes nbp/2 lanes and a warp carries 64/nbp systems.  Lanes of a system form a ring: at
//             offset s lane I meets the two bodies of lane I+s (positions from shared memory), evaluates the
//             2x2 pairs, keeps its own accelerations and adds the reactions to a travelling accumulator that
//             rotates one lane per offset (12 SHFL per 4 pairs) and is sent home after the last offset.
#include "kernels.h"
#include "ensemble.h"
#include "force_common.cuh"

namespace orb {


constexpr int kEnsMaxWarps = 8;

// reference rounding of the integrator with the velocity dtype fixed at compile time
template <bool F32>
__device__ __forceinline__ double ens_kick(double v, double h, double a) {
    const double r = __dadd_rn(v, __dmul_rn(h, a));
    return F32 ? (double)__double2float_rn(r) : r;
}
template <bool F32>
__device__ __forceinline__ double ens_drift(double r, double v, double dt, float dt32) {
    if (F32) return __dadd_rn(r, (double)__fmul_rn(__double2float_rn(v), dt32));
    return __dadd_rn(r, __dmul_rn(v, dt));
}

// Faithful mode.  One warp per system; `blockDim.x / 32` systems per CTA (1 = one CTA per system); lane i owns
// body i and adds its terms in ascending j with the reference's rounding sequence.
// NBP = bodies rounded up to a power of two (compile time: loops fully unrolled, no index arithmetic).
template <int NBP, bool F32>
__global__ void __launch_bounds__(32 * kEnsMaxWarps) ens_step_kernel(const EnsArgs g) {
    __shared__ double4 sp_all[kEnsMaxWarps][NBP];   // {x, y, z, G*m}
    const int warp = threadIdx.x >> 5;
    double4* sp = sp_all[warp];
    const long long sys = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (sys >= g.nsys) return;                      // whole warp exits together
    const int i = threadIdx.x & 31;
    const int nb = g.nb;
    const bool body = i < nb;
    const long long o = sys * nb + i;

    double x = 0, y = 0, z = 0, m = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0;
    if (body) {
        x = g.x[o]; y = g.y[o]; z = g.z[o]; m = g.m[o];
        vx = g.vx[o]; vy = g.vy[o]; vz = g.vz[o];
        ax = g.ax[o]; ay = g.ay[o]; az = g.az[o];
    }
    const double gm = __dmul_rn(g.G, m);            // G * m (physics.py:151-152)
    const double h = g.h, dt = g.dt, eps2 = g.eps2;
    const float dt32 = g.dt32;

    for (long long s = 0; s < g.nsteps; ++s) {
        vx = ens_kick<F32>(vx, h, ax);                            // engine.py:69-70
        vy = ens_kick<F32>(vy, h, ay);
        vz = ens_kick<F32>(vz, h, az);
        x = ens_drift<F32>(x, vx, dt, dt32);                      // engine.py:73-75
        y = ens_drift<F32>(y, vy, dt, dt32);
        z = ens_drift<F32>(z, vz, dt, dt32);
        if (i < NBP) sp[i] = make_double4(x, y, z, gm);
        __syncwarp();
        double bx = 0.0, by = 0.0, bz = 0.0;                      // physics.py:132
        if (body) {
#pragma unroll
            for (int j = 0; j < NBP; ++j) {
                if (j == i || j >= nb) continue;
                const double4 q = sp[j];
                pair_faithful(__dsub_rn(q.x, x), __dsub_rn(q.y, y), __dsub_rn(q.z, z), eps2, q.w, bx, by, bz);
            }
        }
        ax = bx; ay = by; az = bz;
        vx = ens_kick<F32>(vx, h, ax);                            // engine.py:81-82
        vy = ens_kick<F32>(vy, h, ay);
        vz = ens_kick<F32>(vz, h, az);
        __syncwarp();
    }
    if (body) {
        g.x[o] = x; g.y[o] = y; g.z[o] = z;
        g.vx[o] = vx; g.vy[o] = vy; g.vz[o] = vz;
        g.ax[o] = ax; g.ay[o] = ay; g.az[o] = az;
    }
}

// ---------------------------------------------------------------------------------------------
// Fast mode.  NBP = bodies rounded up to a power of two; LPS = NBP/2 lanes per system.
// Padded bodies (index >= nb) and the bodies of systems past the end get zero mass and a far-away dummy
// position (distinct per body), so every pair is finite and contributes exactly 0 -- the inner loop needs
// no predicates.
template <int NBP, bool F32>
__global__ void __launch_bounds__(32 * kEnsMaxWarps) ens_step_fast_kernel(const EnsArgs g) {
    constexpr int LPS = NBP / 2;                 // lanes per system
    constexpr int SPW = 32 / LPS;                // systems per warp
    constexpr int NS = LPS / 2;                  // ring offsets 1..NS (offset NS pairs antipodes: lower half only)
    constexpr int LM = LPS - 1;
    constexpr int STRIDE = 3 * LPS;              // shared-memory stride per system: bank-conflict free for all NBP
    __shared__ double2 sxy_all[kEnsMaxWarps][SPW * STRIDE];
    __shared__ double sz_all[kEnsMaxWarps][SPW * STRIDE];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int I = lane & LM;
    const int sw = lane / LPS;
    const long long sys0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * SPW;
    if (sys0 >= g.nsys) return;                  // whole warp exits together
    const long long sys = sys0 + sw;
    const int nb = g.nb;
    double2* sxy = sxy_all[warp] + sw * STRIDE;
    double* sz = sz_all[warp] + sw * STRIDE;

    bool has[2];
    long long o[2];
    double x[2], y[2], z[2], m[2], vx[2], vy[2], vz[2], ax[2], ay[2], az[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int b = I + k * LPS;
        has[k] = sys < g.nsys && b < nb;
        o[k] = sys * nb + b;
        x[k] = 1e150 * (double)(b + 1); y[k] = 0.0; z[k] = 0.0; m[k] = 0.0;
        vx[k] = vy[k] = vz[k] = ax[k] = ay[k] = az[k] = 0.0;
        if (has[k]) {
            x[k] = g.x[o[k]]; y[k] = g.y[o[k]]; z[k] = g.z[o[k]]; m[k] = g.m[o[k]];
            vx[k] = g.vx[o[k]]; vy[k] = g.vy[o[k]]; vz[k] = g.vz[o[k]];
            ax[k] = g.ax[o[k]]; ay[k] = g.ay[o[k]]; az[k] = g.az[o[k]];
        }
    }
    // partner masses do not change: fetch them once (through the same shared-memory slots)
    double mj[NS > 0 ? NS : 1][2];
    double mi_half[2] = {m[0], m[1]};            // own masses as seen by the antipodal offset
    if (NS > 0) {
        sz[I] = m[0]; sz[I + LPS] = m[1];
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const int J = (I + t + 1) & LM;
            const bool vt = (t + 1 < NS) || (I < NS);
            mj[t][0] = vt ? sz[J] : 0.0;
            mj[t][1] = vt ? sz[J + LPS] : 0.0;
            if (t + 1 == NS && !vt) mi_half[0] = mi_half[1] = 0.0;
        }
        __syncwarp();
    }
    const double h = g.h, dt = g.dt, eps2 = g.eps2, G = g.G;
    const float dt32 = g.dt32;
    const int group = lane & ~LM;

    for (long long s = 0; s < g.nsteps; ++s) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (has[k]) {
                vx[k] = ens_kick<F32>(vx[k], h, ax[k]);                  // engine.py:69-70
                vy[k] = ens_kick<F32>(vy[k], h, ay[k]);
                vz[k] = ens_kick<F32>(vz[k], h, az[k]);
                x[k] = ens_drift<F32>(x[k], vx[k], dt, dt32);            // engine.py:73-75
                y[k] = ens_drift<F32>(y[k], vy[k], dt, dt32);
                z[k] = ens_drift<F32>(z[k], vz[k], dt, dt32);
            }
            if (NS > 0) {
                sxy[I + k * LPS] = make_double2(x[k], y[k]);
                sz[I + k * LPS] = z[k];
            }
        }
        if (NS > 0) __syncwarp();
        double a0x, a0y, a0z, a1x, a1y, a1z;
        {   // the lane's own pair (I, I + LPS)
            const double dx = x[1] - x[0], dy = y[1] - y[0], dz = z[1] - z[0];
            const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
            int hi;
            const double s0 = inv_r3_plain(r2, hi);
            const double si = s0 * m[1], sj = s0 * m[0];
            a0x = si * dx; a0y = si * dy; a0z = si * dz;                 // physics.py:151
            a1x = -sj * dx; a1y = -sj * dy; a1z = -sj * dz;              // physics.py:152
        }
        double c0x = 0.0, c0y = 0.0, c0z = 0.0, c1x = 0.0, c1y = 0.0, c1z = 0.0;   // travelling: reactions on lane I+s
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const int J = (I + t + 1) & LM;
            const double mi0 = (t + 1 == NS) ? mi_half[0] : m[0];
            const double mi1 = (t + 1 == NS) ? mi_half[1] : m[1];
#pragma unroll
            for (int kj = 0; kj < 2; ++kj) {
                const double2 pxy = sxy[J + kj * LPS];
                const double pz = sz[J + kj * LPS];
                const double mjj = mj[t][kj];
                double& cx = kj == 0 ? c0x : c1x;
                double& cy = kj == 0 ? c0y : c1y;
                double& cz = kj == 0 ? c0z : c1z;
                int hi;
                {
                    const double dx = pxy.x - x[0], dy = pxy.y - y[0], dz = pz - z[0];
                    const double s0 = inv_r3_plain(fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2))), hi);
                    const double si = s0 * mjj, sj = s0 * mi0;
                    a0x = fma(si, dx, a0x); a0y = fma(si, dy, a0y); a0z = fma(si, dz, a0z);
                    cx = fma(-sj, dx, cx); cy = fma(-sj, dy, cy); cz = fma(-sj, dz, cz);
                }
                {
                    const double dx = pxy.x - x[1], dy = pxy.y - y[1], dz = pz - z[1];
                    const double s0 = inv_r3_plain(fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2))), hi);
                    const double si = s0 * mjj, sj = s0 * mi1;
                    a1x = fma(si, dx, a1x); a1y = fma(si, dy, a1y); a1z = fma(si, dz, a1z);
                    cx = fma(-sj, dx, cx); cy = fma(-sj, dy, cy); cz = fma(-sj, dz, cz);
                }
            }
            // the next offset meets lane I+s+1, whose accumulators sit one lane up; after the last offset they
            // go home: they belong to lane I + NS
            const int src = group | ((t + 1 < NS ? I + 1 : I - NS) & LM);
            c0x = __shfl_sync(0xffffffffu, c0x, src); c0y = __shfl_sync(0xffffffffu, c0y, src);
            c0z = __shfl_sync(0xffffffffu, c0z, src); c1x = __shfl_sync(0xffffffffu, c1x, src);
            c1y = __shfl_sync(0xffffffffu, c1y, src); c1z = __shfl_sync(0xffffffffu, c1z, src);
        }
        ax[0] = G * (a0x + c0x); ay[0] = G * (a0y + c0y); az[0] = G * (a0z + c0z);
        ax[1] = G * (a1x + c1x); ay[1] = G * (a1y + c1y); az[1] = G * (a1z + c1z);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (has[k]) {
                vx[k] = ens_kick<F32>(vx[k], h, ax[k]);                  // engine.py:81-82
                vy[k] = ens_kick<F32>(vy[k], h, ay[k]);
                vz[k] = ens_kick<F32>(vz[k], h, az[k]);
            }
        }
        if (NS > 0) __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (has[k]) {
            g.x[o[k]] = x[k]; g.y[o[k]] = y[k]; g.z[o[k]] = z[k];
            g.vx[o[k]] = vx[k]; g.vy[o[k]] = vy[k]; g.vz[o[k]] = vz[k];
            g.ax[o[k]] = ax[k]; g.ay[o[k]] = ay[k]; g.az[o[k]] = az[k];
        }
    }
}

// Initial accelerations (engine.py:41) -- same code path with zero steps would skip the force;
// this kernel evaluates the force once on the resident positions.
template <bool FAITHFUL>
__global__ void __launch_bounds__(32) ens_accel_kernel(const EnsArgs g) {
    __shared__ double4 sp[32];
    const long long sys = blockIdx.x;
    const int i = threadIdx.x;
    const bool body = i < g.nb;
    const long long o = sys * g.nb + i;
    double x = 0, y = 0, z = 0;
    if (body) {
        x = g.x[o]; y = g.y[o]; z = g.z[o];
        sp[i] = make_double4(x, y, z, FAITHFUL ? __dmul_rn(g.G, g.m[o]) : g.m[o]);
    }
    __syncwarp();
    if (!body) return;
    double bx = 0.0, by = 0.0, bz = 0.0;
    for (int j = 0; j < g.nb; ++j) {
        if (j == i) continue;
        const double4 q = sp[j];
        if (FAITHFUL) {
            pair_faithful(__dsub_rn(q.x, x), __dsub_rn(q.y, y), __dsub_rn(q.z, z), g.eps2, q.w, bx, by, bz);
        } else {
            const double dx = q.x - x, dy = q.y - y, dz = q.z - z;
            const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, g.eps2)));
            const double y0 = rsqrt_seed(r2);
            const double u = y0 * y0;
            const double e = fma(-r2, u, 1.0);
            const double w = (q.w * y0) * u;
            const double sc = fma(w, e * fma(1.875, e, 1.5), w);
            bx = fma(sc, dx, bx); by = fma(sc, dy, by); bz = fma(sc, dz, bz);
        }
    }
    if (!FAITHFUL) { bx *= g.G; by *= g.G; bz *= g.G; }
    g.ax[o] = bx; g.ay[o] = by; g.az[o] = bz;
}

// E = K + U per system (fp64, straightforward order)
__global__ void __launch_bounds__(32) ens_energy_kernel(const EnsArgs g, double* E) {
    __shared__ double4 sp[32];
    const long long sys = blockIdx.x;
    const int i = threadIdx.x;
    const bool body = i < g.nb;
    const long long o = sys * g.nb + i;
    double x = 0, y = 0, z = 0, m = 0, e = 0;
    if (body) {
        x = g.x[o]; y = g.y[o]; z = g.z[o]; m = g.m[o];
        sp[i] = make_double4(x, y, z, m);
        const double vx = g.vx[o], vy = g.vy[o], vz = g.vz[o];
        e = 0.5 * m * (vx * vx + vy * vy + vz * vz);
    }
    __syncwarp();
    if (body) {
        double u = 0.0;
        for (int j = i + 1; j < g.nb; ++j) {
            const double4 q = sp[j];
            const double dx = q.x - x, dy = q.y - y, dz = q.z - z;
            u += q.w / sqrt(fma(dx, dx, fma(dy, dy, fma(dz, dz, g.eps2))));
        }
        e -= g.G * m * u;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    if (i == 0) E[sys] = e;
}

template <bool FAITHFUL, int NBP>
static void launch_ens_step_t(const EnsArgs& a, int w, cudaStream_t st) {
    const int block = 32 * w;
    if (FAITHFUL) {
        const unsigned grid = (unsigned)((a.nsys + w - 1) / w);                 // one warp per system
        if (a.vel_f32)
            ens_step_kernel<NBP, true><<<grid, block, 0, st>>>(a);
        else
            ens_step_kernel<NBP, false><<<grid, block, 0, st>>>(a);
    } else {
        const long long per_cta = (long long)w * (64 / NBP);                    // 64/NBP systems per warp
        const unsigned grid = (unsigned)((a.nsys + per_cta - 1) / per_cta);
        if (a.vel_f32)
            ens_step_fast_kernel<NBP, true><<<grid, block, 0, st>>>(a);
        else
            ens_step_fast_kernel<NBP, false><<<grid, block, 0, st>>>(a);
    }
}

cudaError_t launch_ens_step(const EnsArgs& a, bool faithful, cudaStream_t st) {
    int w = a.warps_per_cta;
    if (w < 1) w = 1;
    if (w > kEnsMaxWarps) w = kEnsMaxWarps;
#define ORB_ENS_CASE(P)                                                       \
    case P:                                                                   \
        if (faithful) launch_ens_step_t<true, P>(a, w, st);                   \
        else launch_ens_step_t<false, P>(a, w, st);                           \
        break;
    switch (a.nbp) {
        ORB_ENS_CASE(2)
        ORB_ENS_CASE(4)
        ORB_ENS_CASE(8)
        ORB_ENS_CASE(16)
        ORB_ENS_CASE(32)
        default: return cudaErrorInvalidValue;
    }
#undef ORB_ENS_CASE
    return cudaGetLastError();
}

cudaError_t launch_ens_accel(const EnsArgs& a, bool faithful, cudaStream_t st) {
    if (faithful)
        ens_accel_kernel<true><<<(unsigned)a.nsys, 32, 0, st>>>(a);
    else
        ens_accel_kernel<false><<<(unsigned)a.nsys, 32, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_ens_energy(const EnsArgs& a, double* E, cudaStream_t st) {
    ens_energy_kernel<<<(unsigned)a.nsys, 32, 0, st>>>(a, E);
    return cudaGetLastError();
}

}  // namespace orb
