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
//   FAST:     all 32 lanes busy: 32/nbp lanes share a body, each taking a strided
//             subset of the sources (rsqrt seed + fp64 polynomial), xor-shuffle combine.
#include "kernels.h"
#include "ensemble.h"

namespace orb {


template <bool FAITHFUL>
__global__ void __launch_bounds__(32) ens_step_kernel(const EnsArgs g) {
    __shared__ double4 sp[32];                      // {x,y,z,m or G*m}
    const long long sys = blockIdx.x;
    const int lane = threadIdx.x;
    const int i = FAITHFUL ? lane : (lane & (g.nbp - 1));
    const int part = FAITHFUL ? 0 : lane / g.nbp;
    const int nparts = FAITHFUL ? 1 : 32 / g.nbp;
    const bool owner = (i < g.nb) && (part == 0);   // lane that holds / stores body i
    const bool body = i < g.nb;
    const long long o = sys * g.nb + i;
    const bool f32 = g.vel_f32 != 0;

    double x = 0, y = 0, z = 0, m = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0;
    if (body) {                                     // replicas load the same lines (L1 hit)
        x = g.x[o]; y = g.y[o]; z = g.z[o]; m = g.m[o];
        vx = g.vx[o]; vy = g.vy[o]; vz = g.vz[o];
        ax = g.ax[o]; ay = g.ay[o]; az = g.az[o];
    }
    const double mw = FAITHFUL ? __dmul_rn(g.G, m) : m;

    for (long long s = 0; s < g.nsteps; ++s) {
        if (body) {
            vx = kick_faithful(vx, g.h, ax, f32);                 // engine.py:69-70
            vy = kick_faithful(vy, g.h, ay, f32);
            vz = kick_faithful(vz, g.h, az, f32);
            x = drift_faithful(x, vx, g.dt, g.dt32, f32);         // engine.py:73-75
            y = drift_faithful(y, vy, g.dt, g.dt32, f32);
            z = drift_faithful(z, vz, g.dt, g.dt32, f32);
            if (part == 0) sp[i] = make_double4(x, y, z, mw);
        }
        __syncwarp();
        double bx = 0.0, by = 0.0, bz = 0.0;
        if (FAITHFUL) {
            if (body) {
#pragma unroll 4
                for (int j = 0; j < g.nb; ++j) {
                    if (j == i) continue;
                    const double4 q = sp[j];
                    pair_faithful(__dsub_rn(q.x, x), __dsub_rn(q.y, y), __dsub_rn(q.z, z), g.eps2, q.w, bx, by, bz);
                }
            }
        } else {
            if (body) {
                for (int j = part; j < g.nb; j += nparts) {
                    const double4 q = sp[j];
                    const double dx = q.x - x, dy = q.y - y, dz = q.z - z;
                    const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, g.eps2)));
                    const double y0 = rsqrt_seed(r2);
                    const double u = y0 * y0;
                    const double e = fma(-r2, u, 1.0);
                    const double w = (q.w * y0) * u;
                    double sc = fma(w, e * fma(1.875, e, 1.5), w);
                    sc = (j == i) ? 0.0 : sc;
                    bx = fma(sc, dx, bx); by = fma(sc, dy, by); bz = fma(sc, dz, bz);
                }
            }
            for (int off = g.nbp; off < 32; off <<= 1) {          // combine the parts
                bx += __shfl_xor_sync(0xffffffffu, bx, off);
                by += __shfl_xor_sync(0xffffffffu, by, off);
                bz += __shfl_xor_sync(0xffffffffu, bz, off);
            }
            bx *= g.G; by *= g.G; bz *= g.G;
        }
        ax = bx; ay = by; az = bz;
        if (body) {
            vx = kick_faithful(vx, g.h, ax, f32);                 // engine.py:81-82
            vy = kick_faithful(vy, g.h, ay, f32);
            vz = kick_faithful(vz, g.h, az, f32);
        }
        __syncwarp();
    }
    if (owner) {
        g.x[o] = x; g.y[o] = y; g.z[o] = z;
        g.vx[o] = vx; g.vy[o] = vy; g.vz[o] = vz;
        g.ax[o] = ax; g.ay[o] = ay; g.az[o] = az;
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

cudaError_t launch_ens_step(const EnsArgs& a, bool faithful, cudaStream_t st) {
    if (faithful)
        ens_step_kernel<true><<<(unsigned)a.nsys, 32, 0, st>>>(a);
    else
        ens_step_kernel<false><<<(unsigned)a.nsys, 32, 0, st>>>(a);
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
