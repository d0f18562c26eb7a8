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

// One warp per system; `blockDim.x / 32` systems per CTA (1 = one CTA per system).
// NBP = bodies rounded up to a power of two (compile time: loops fully unrolled, no index arithmetic).
template <bool FAITHFUL, int NBP, bool F32>
__global__ void __launch_bounds__(32 * kEnsMaxWarps) ens_step_kernel(const EnsArgs g) {
    __shared__ double4 sp_all[kEnsMaxWarps][NBP];   // {x,y,z,m or G*m}
    const int warp = threadIdx.x >> 5;
    double4* sp = sp_all[warp];
    const int sys = blockIdx.x * (blockDim.x >> 5) + warp;
    if (sys >= (int)g.nsys) return;                 // whole warp exits together
    const int lane = threadIdx.x & 31;
    // lanes sharing one body in fast mode: 32/NBP, but never more than there are sources
    constexpr int NPARTS = FAITHFUL ? 1 : ((32 / NBP) < NBP ? (32 / NBP) : NBP);
    constexpr int NSRC = NBP / NPARTS;                      // sources per lane
    const int i = FAITHFUL ? lane : (lane & (NBP - 1));
    const int part = FAITHFUL ? 0 : lane / NBP;
    const int nb = g.nb;
    const bool body = i < nb;
    const bool owner = body && (part == 0);         // lane that stores body i
    const int o = sys * nb + i;

    double x = 0, y = 0, z = 0, m = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0;
    if (body) {                                     // replicas load the same lines (L1 hit)
        x = g.x[o]; y = g.y[o]; z = g.z[o]; m = g.m[o];
        vx = g.vx[o]; vy = g.vy[o]; vz = g.vz[o];
        ax = g.ax[o]; ay = g.ay[o]; az = g.az[o];
    }
    const double mw = FAITHFUL ? __dmul_rn(g.G, m) : m;
    const double h = g.h, dt = g.dt, eps2 = g.eps2;
    const float dt32 = g.dt32;

    for (long long s = 0; s < g.nsteps; ++s) {
        vx = ens_kick<F32>(vx, h, ax);                            // engine.py:69-70
        vy = ens_kick<F32>(vy, h, ay);
        vz = ens_kick<F32>(vz, h, az);
        x = ens_drift<F32>(x, vx, dt, dt32);                      // engine.py:73-75
        y = ens_drift<F32>(y, vy, dt, dt32);
        z = ens_drift<F32>(z, vz, dt, dt32);
        if (lane < NBP) sp[lane] = make_double4(x, y, z, body ? mw : 0.0);   // padded bodies: zero mass
        __syncwarp();
        double bx = 0.0, by = 0.0, bz = 0.0;
        if (FAITHFUL) {
            if (body) {
#pragma unroll
                for (int j = 0; j < NBP; ++j) {
                    if (j == i || j >= nb) continue;
                    const double4 q = sp[j];
                    pair_faithful(__dsub_rn(q.x, x), __dsub_rn(q.y, y), __dsub_rn(q.z, z), eps2, q.w, bx, by, bz);
                }
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < NSRC; ++jj) {
                if (part >= NPARTS) break;                        // tiny systems: surplus lanes contribute zero
                const int j = jj * NPARTS + part;
                const double4 q = sp[j];
                const double dx = q.x - x, dy = q.y - y, dz = q.z - z;
                const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
                const double y0 = rsqrt_seed(r2);
                const double u = y0 * y0;
                const double e = fma(-r2, u, 1.0);
                const double w = (q.w * y0) * u;
                double sc = fma(w, e * fma(1.875, e, 1.5), w);
                sc = (j == i) ? 0.0 : sc;                         // self pair (also kills the eps = 0 NaN)
                bx = fma(sc, dx, bx); by = fma(sc, dy, by); bz = fma(sc, dz, bz);
            }
#pragma unroll
            for (int off = NBP; off < 32; off <<= 1) {            // combine the lanes sharing a body
                bx += __shfl_xor_sync(0xffffffffu, bx, off);
                by += __shfl_xor_sync(0xffffffffu, by, off);
                bz += __shfl_xor_sync(0xffffffffu, bz, off);
            }
            bx *= g.G; by *= g.G; bz *= g.G;
        }
        ax = bx; ay = by; az = bz;
        vx = ens_kick<F32>(vx, h, ax);                            // engine.py:81-82
        vy = ens_kick<F32>(vy, h, ay);
        vz = ens_kick<F32>(vz, h, az);
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

template <bool FAITHFUL, int NBP>
static void launch_ens_step_t(const EnsArgs& a, unsigned grid, int block, cudaStream_t st) {
    if (a.vel_f32)
        ens_step_kernel<FAITHFUL, NBP, true><<<grid, block, 0, st>>>(a);
    else
        ens_step_kernel<FAITHFUL, NBP, false><<<grid, block, 0, st>>>(a);
}

cudaError_t launch_ens_step(const EnsArgs& a, bool faithful, cudaStream_t st) {
    int w = a.warps_per_cta;
    if (w < 1) w = 1;
    if (w > kEnsMaxWarps) w = kEnsMaxWarps;
    const unsigned grid = (unsigned)((a.nsys + w - 1) / w);
    const int block = 32 * w;
#define ORB_ENS_CASE(P)                                                       \
    case P:                                                                   \
        if (faithful) launch_ens_step_t<true, P>(a, grid, block, st);         \
        else launch_ens_step_t<false, P>(a, grid, block, st);                 \
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
