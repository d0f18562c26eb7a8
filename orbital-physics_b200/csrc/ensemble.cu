// ensemble.cu -- batched ensemble of independent small systems (sm_100a): one warp per system (bit-exact) or
// nbody/2 lanes per system, 64/nbody systems per warp (fast).
//
// Equivalent to `nsys` separate reference SimulationEngine instances
// (core/engine.py:19-46,65-97; force: core/physics.py:125-159) stepped in
// lockstep, including -- when radii are given -- the contact sweep each engine runs after its step
// (engine.py:85 -> physics.py:510-535) and per-body velocity dtypes.  State is SoA [nsys][nbody] fp64.
// One step per launch (HBM-bound): x, u (half-kicked velocity), m in and x, u out = 104 B per body-step
// (see the first / last note below); with `nsteps` fused in one launch the state stays in registers and the
// kernel is FP64-bound.
//
//   FAITHFUL: lane i owns body i and accumulates j ascending with the reference's
//             rounding sequence (bit-exact with the reference engine).
//   FAST:     ens_step_fast_kernel -- every unordered pair once (the reference's own half-matrix trick,
//             physics.py:136-155), 20 FP64 instructions per pair.  A lane owns TWO bodies (I and I + nbp/2), so
//             a system takes nbp/2 lanes and a warp carries 64/nbp systems.  Lanes of a system form a ring: at
//             offset s lane I meets the two bodies of lane I+s (positions from shared memory), evaluates the
//             2x2 pairs, keeps its own accelerations and adds the reactions to a travelling accumulator that
//             rotates one lane per offset (12 SHFL per 4 pairs) and is sent home after the last offset.
#include <algorithm>
#include <cstdlib>

#include "kernels.h"
#include "contacts.cuh"
#include "ensemble.h"
#include "force_common.cuh"

namespace orb {


constexpr int kEnsMaxWarps = 8;

// Programmatic dependent launch (sm_90+): a step launch lets the next one start filling the SMs at once
// (launch_dependents) and waits for its predecessor's memory only right before it reads the state (wait).  In the
// one-step-per-launch mode of a small batch a launch is a few microseconds of work, about as much as the gap
// between two dependent launches.  Both are no-ops for launches without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
constexpr int kEnsDetectWarps = 4;      // warps per CTA of the variants that carry the contact sweep's shared memory

// Velocity dtype of the bodies: VM 0 = every velocity is a float64 array, 1 = every velocity is
// float32 (what Object(...) stores, physics.py:184), 2 = per-body flags (mixed systems, physics.py:448-449).
template <int VM>
__device__ __forceinline__ double ens_kick(double v, double h, double a, bool f32) {
    const double r = __dadd_rn(v, __dmul_rn(h, a));
    return (VM == 1 || (VM == 2 && f32)) ? (double)__double2float_rn(r) : r;
}
template <int VM>
__device__ __forceinline__ double ens_drift(double r, double v, double dt, float dt32, bool f32) {
    if (VM == 1 || (VM == 2 && f32)) return __dadd_rn(r, (double)__fmul_rn(__double2float_rn(v), dt32));
    return __dadd_rn(r, __dmul_rn(v, dt));
}

// handle_collisions(restitution) of ONE small system held in shared memory (physics.py:510-535 -> 391-422):
// the reference's own sequential loop, run by a single lane.  Returns the touching pairs processed.
__device__ inline int ens_sweep(ContactBody* b, int nb, double restitution) {
    int hits = 0;
    for (int i = 0; i < nb; ++i)
        for (int j = i + 1; j < nb; ++j) {
            ContactBody& A = b[i];
            ContactBody& B = b[j];
            if (overlap_exact(__dsub_rn(A.x, B.x), __dsub_rn(A.y, B.y), __dsub_rn(A.z, B.z), A.R, B.R)) {
                collide_bodies(A, B, restitution);
                ++hits;
            }
        }
    return hits;
}

// What one launch does with the state (`first` / `last` in EnsArgs):
//   first: the velocity plane holds v_n and the acceleration planes a_n (the synchronised state every
//          orb_ens_step call starts from and ends with) -> u = kick(v_n, a_n) is formed here;
//          otherwise the velocity plane already holds u (the half-kicked velocity) and a is not read.
//   per step: x += dt u; a = F(x); v = kick(u, a); [contacts]; u = kick(v, a) unless this is the very last step.
//   last:  v_{n+k} and a_{n+k} are written; otherwise only x and u.
// A one-step-per-launch sequence therefore moves x, u, m in and x, u out per body-step (104 B) instead
// of 152 B, with exactly the reference's rounding sequence: two separately rounded half-kicks per step.

// Faithful mode.  One warp per system; `blockDim.x / 32` systems per CTA; lane i owns body i and adds its terms in
// ascending j with the reference's rounding sequence.
// NBP = bodies rounded up to a power of two (compile time: loops fully unrolled, no index arithmetic).
template <int NBP, int VM, bool DETECT>
__global__ void __launch_bounds__(32 * kEnsMaxWarps) ens_step_kernel(const EnsArgs g) {
    __shared__ double4 sp_all[kEnsMaxWarps][NBP];   // {x, y, z, G*m}
    __shared__ ContactBody cb_all[DETECT ? kEnsDetectWarps : 1][DETECT ? NBP : 1];
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5;
    double4* sp = sp_all[warp];
    const long long sys = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (sys >= g.nsys) return;                      // whole warp exits together
    pdl_wait();
    const int i = threadIdx.x & 31;
    const int nb = g.nb;
    const bool body = i < nb;
    const long long o = sys * nb + i;

    double x = 0, y = 0, z = 0, m = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0, R = 0;
    bool f32 = VM == 1;
    if (body) {
        x = g.x[o]; y = g.y[o]; z = g.z[o]; m = g.m[o];
        vx = g.vx[o]; vy = g.vy[o]; vz = g.vz[o];
        if (g.first) { ax = g.ax[o]; ay = g.ay[o]; az = g.az[o]; }
        if (VM == 2) f32 = g.vf32[o] != 0;
        if (DETECT) R = g.radius[o];
    }
    const double gm = __dmul_rn(g.G, m);            // G * m (physics.py:151-152)
    const double h = g.h, dt = g.dt, eps2 = g.eps2;
    const float dt32 = g.dt32;
    if (g.first) {
        vx = ens_kick<VM>(vx, h, ax, f32);                        // engine.py:69-70
        vy = ens_kick<VM>(vy, h, ay, f32);
        vz = ens_kick<VM>(vz, h, az, f32);
    }

    for (long long s = 0; s < g.nsteps; ++s) {
        x = ens_drift<VM>(x, vx, dt, dt32, f32);                  // engine.py:73-75
        y = ens_drift<VM>(y, vy, dt, dt32, f32);
        z = ens_drift<VM>(z, vz, dt, dt32, f32);
        if (i < NBP) sp[i] = make_double4(x, y, z, gm);
        __syncwarp();
        double bx = 0.0, by = 0.0, bz = 0.0;                      // physics.py:132
        if (body) {
#pragma unroll
            for (int j = 0; j < NBP; ++j) {
                if (j == i || j >= nb) continue;
                const double4 q = sp[j];
                const double dx = __dsub_rn(q.x, x), dy = __dsub_rn(q.y, y), dz = __dsub_rn(q.z, z);
                pair_faithful(dx, dy, dz, eps2, q.w, bx, by, bz);
            }
        }
        ax = bx; ay = by; az = bz;
        vx = ens_kick<VM>(vx, h, ax, f32);                        // engine.py:81-82
        vy = ens_kick<VM>(vy, h, ay, f32);
        vz = ens_kick<VM>(vz, h, az, f32);
        __syncwarp();
        if (DETECT) {
            // engine.py:85: every lane publishes its body, lane 0 tests and resolves in the reference's order
            ContactBody* cb = cb_all[warp];
            if (body) cb[i] = ContactBody{x, y, z, vx, vy, vz, m, R, f32};
            __syncwarp();
            bool any = false;
            if (body) {
                for (int j = i + 1; j < nb; ++j)
                    any |= overlap_exact(__dsub_rn(x, cb[j].x), __dsub_rn(y, cb[j].y), __dsub_rn(z, cb[j].z), R, cb[j].R);
            }
            if (__any_sync(0xffffffffu, any)) {
                if (i == 0) {
                    const int hits = ens_sweep(cb, nb, g.restitution);
                    if (g.contacts) atomicAdd(g.contacts, (unsigned long long)hits);
                }
                __syncwarp();
                if (body) {
                    const ContactBody c = cb[i];
                    x = c.x; y = c.y; z = c.z; vx = c.vx; vy = c.vy; vz = c.vz;
                }
            }
            __syncwarp();
        }
        if (s + 1 < g.nsteps || !g.last) {
            vx = ens_kick<VM>(vx, h, ax, f32);                    // the next step's first half-kick (engine.py:69-70)
            vy = ens_kick<VM>(vy, h, ay, f32);
            vz = ens_kick<VM>(vz, h, az, f32);
        }
    }
    if (body) {
        g.x[o] = x; g.y[o] = y; g.z[o] = z;
        g.vx[o] = vx; g.vy[o] = vy; g.vz[o] = vz;
        if (g.last) { g.ax[o] = ax; g.ay[o] = ay; g.az[o] = az; }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast mode.  NBP = bodies rounded up to a power of two; LPS = NBP/2 lanes per system.
// Padded bodies (index >= nb) and the bodies of systems past the end get zero mass and a far-away dummy
// position (distinct per body), so every pair is finite and contributes exactly 0 -- the inner loop needs
// no predicates.  DETECT: a conservative per-pair threshold on r^2 (one compare per pair) flags systems that may
// have a touching pair; those run the exact sequential sweep.
// One warp advances the SPW systems [sys0, sys0 + SPW) by `nsteps` steps.  sxy_w / sz_w / cb_w: this warp's shared
// memory.  State loads bypass L1 (__ldcg): the time-sliced kernel below hands a system from one SM to another.
template <int NBP, int VM, bool DETECT>
__device__ __forceinline__ void ens_fast_body(const EnsArgs& g, long long sys0, long long nsteps, bool first,
                                              bool last, double2* sxy_w, double* sz_w, ContactBody* cb_w) {
    constexpr int LPS = NBP / 2;                 // lanes per system
    constexpr int NS = LPS / 2;                  // ring offsets 1..NS (offset NS pairs antipodes: lower half only)
    constexpr int LM = LPS - 1;
    constexpr int STRIDE = 3 * LPS;              // shared-memory stride per system: bank-conflict free for all NBP
    const int lane = threadIdx.x & 31;
    const int I = lane & LM;
    const int sw = lane / LPS;
    const long long sys = sys0 + sw;
    const int nb = g.nb;
    double2* sxy = sxy_w + sw * STRIDE;
    double* sz = sz_w + sw * STRIDE;

    bool has[2], f32[2];
    long long o[2];
    double x[2], y[2], z[2], m[2], vx[2], vy[2], vz[2], ax[2], ay[2], az[2], R[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int b = I + k * LPS;
        has[k] = sys < g.nsys && b < nb;
        o[k] = sys * nb + b;
        x[k] = 1e150 * (double)(b + 1); y[k] = 0.0; z[k] = 0.0; m[k] = 0.0; R[k] = 0.0;
        vx[k] = vy[k] = vz[k] = ax[k] = ay[k] = az[k] = 0.0;
        f32[k] = VM == 1;
        if (has[k]) {
            x[k] = __ldcg(g.x + o[k]); y[k] = __ldcg(g.y + o[k]); z[k] = __ldcg(g.z + o[k]); m[k] = g.m[o[k]];
            vx[k] = __ldcg(g.vx + o[k]); vy[k] = __ldcg(g.vy + o[k]); vz[k] = __ldcg(g.vz + o[k]);
            if (first) { ax[k] = __ldcg(g.ax + o[k]); ay[k] = __ldcg(g.ay + o[k]); az[k] = __ldcg(g.az + o[k]); }
            if (VM == 2) f32[k] = g.vf32[o[k]] != 0;
            if (DETECT) R[k] = g.radius[o[k]];
        }
    }
    // partner masses (and radii) do not change: fetch them once (through the same shared-memory slots)
    double mj[NS > 0 ? NS : 1][2];
    double thr[DETECT && NS > 0 ? NS : 1][2][2];     // [offset][partner body][own body]: conservative r^2 bound
    double mi_half[2] = {m[0], m[1]};            // own masses as seen by the antipodal offset
    const double eps2 = g.eps2;
    constexpr double kSlack = 1.0 + 9.5367431640625e-07;      // 1 + 2^-20: covers the rounding of either r^2 form
    double thr_own = 0.0;
    if (DETECT) { const double rs = R[0] + R[1]; thr_own = fma(rs, rs, eps2) * kSlack; }
    if (NS > 0) {
        sz[I] = m[0]; sz[I + LPS] = m[1];
        if (DETECT) { sxy[I] = make_double2(R[0], 0.0); sxy[I + LPS] = make_double2(R[1], 0.0); }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const int J = (I + t + 1) & LM;
            const bool vt = (t + 1 < NS) || (I < NS);
            mj[t][0] = vt ? sz[J] : 0.0;
            mj[t][1] = vt ? sz[J + LPS] : 0.0;
            if (t + 1 == NS && !vt) mi_half[0] = mi_half[1] = 0.0;
            if (DETECT) {
#pragma unroll
                for (int kj = 0; kj < 2; ++kj)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const double rs = R[k] + sxy[J + kj * LPS].x;
                        thr[t][kj][k] = vt ? fma(rs, rs, eps2) * kSlack : -1.0;   // the upper half skips the antipodes
                    }
            }
        }
        __syncwarp();
    }
    const double h = g.h, dt = g.dt, G = g.G;
    const float dt32 = g.dt32;
    const int group = lane & ~LM;
    if (first) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (has[k]) {
                vx[k] = ens_kick<VM>(vx[k], h, ax[k], f32[k]);           // engine.py:69-70
                vy[k] = ens_kick<VM>(vy[k], h, ay[k], f32[k]);
                vz[k] = ens_kick<VM>(vz[k], h, az[k], f32[k]);
            }
        }
    }

    for (long long s = 0; s < nsteps; ++s) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (has[k]) {
                x[k] = ens_drift<VM>(x[k], vx[k], dt, dt32, f32[k]);     // engine.py:73-75
                y[k] = ens_drift<VM>(y[k], vy[k], dt, dt32, f32[k]);
                z[k] = ens_drift<VM>(z[k], vz[k], dt, dt32, f32[k]);
            }
            if (NS > 0) {
                sxy[I + k * LPS] = make_double2(x[k], y[k]);
                sz[I + k * LPS] = z[k];
            }
        }
        if (NS > 0) __syncwarp();
        bool flag = false;
        double a0x, a0y, a0z, a1x, a1y, a1z;
        {   // the lane's own pair (I, I + LPS)
            const double dx = x[1] - x[0], dy = y[1] - y[0], dz = z[1] - z[0];
            const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
            if (DETECT) flag |= r2 <= thr_own;
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
                    const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
                    if (DETECT) flag |= r2 <= thr[t][kj][0];
                    const double s0 = inv_r3_plain(r2, hi);
                    const double si = s0 * mjj, sj = s0 * mi0;
                    a0x = fma(si, dx, a0x); a0y = fma(si, dy, a0y); a0z = fma(si, dz, a0z);
                    cx = fma(-sj, dx, cx); cy = fma(-sj, dy, cy); cz = fma(-sj, dz, cz);
                }
                {
                    const double dx = pxy.x - x[1], dy = pxy.y - y[1], dz = pz - z[1];
                    const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
                    if (DETECT) flag |= r2 <= thr[t][kj][1];
                    const double s0 = inv_r3_plain(r2, hi);
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
                vx[k] = ens_kick<VM>(vx[k], h, ax[k], f32[k]);           // engine.py:81-82
                vy[k] = ens_kick<VM>(vy[k], h, ay[k], f32[k]);
                vz[k] = ens_kick<VM>(vz[k], h, az[k], f32[k]);
            }
        }
        if (NS > 0) __syncwarp();
        if (DETECT) {
            // engine.py:85 for the systems whose prefilter fired: publish the bodies, one lane per system runs the
            // reference's sequential sweep with the exact test, everybody reloads
            const unsigned fired = __ballot_sync(0xffffffffu, flag);
            if (fired) {
                const bool mine = ((fired >> group) & ((1u << LPS) - 1u)) != 0u && sys < g.nsys;
                ContactBody* cb = cb_w + sw * NBP;
                if (mine) {
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        cb[I + k * LPS] = ContactBody{x[k], y[k], z[k], vx[k], vy[k], vz[k], m[k], R[k], f32[k]};
                }
                __syncwarp();
                if (mine && I == 0) {
                    const int hits = ens_sweep(cb, nb, g.restitution);
                    if (hits && g.contacts) atomicAdd(g.contacts, (unsigned long long)hits);
                }
                __syncwarp();
                if (mine) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const ContactBody c = cb[I + k * LPS];
                        x[k] = c.x; y[k] = c.y; z[k] = c.z; vx[k] = c.vx; vy[k] = c.vy; vz[k] = c.vz;
                    }
                }
                __syncwarp();
            }
        }
        if (s + 1 < nsteps || !last) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (has[k]) {
                    vx[k] = ens_kick<VM>(vx[k], h, ax[k], f32[k]);       // the next step's first half-kick
                    vy[k] = ens_kick<VM>(vy[k], h, ay[k], f32[k]);
                    vz[k] = ens_kick<VM>(vz[k], h, az[k], f32[k]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (has[k]) {
            g.x[o[k]] = x[k]; g.y[o[k]] = y[k]; g.z[o[k]] = z[k];
            g.vx[o[k]] = vx[k]; g.vy[o[k]] = vy[k]; g.vz[o[k]] = vz[k];
            if (last) { g.ax[o[k]] = ax[k]; g.ay[o[k]] = ay[k]; g.az[o[k]] = az[k]; }
        }
    }
}

// 4 warps per CTA, 4 CTAs per SM (128 registers): with the resident-CTA hint ptxas schedules ~1 % better than with
// the bare 256-thread bound (profiles/r2b_ens_narrow.txt)
constexpr int kEnsFastWarps = 4;
template <int NBP, int VM, bool DETECT>
__global__ void __launch_bounds__(32 * kEnsFastWarps, 4) ens_step_fast_kernel(const EnsArgs g) {
    constexpr int SPW = 64 / NBP;                // systems per warp
    constexpr int SLOTS = SPW * 3 * (NBP / 2);
    __shared__ double2 sxy_all[kEnsFastWarps][SLOTS];
    __shared__ double sz_all[kEnsFastWarps][SLOTS];
    __shared__ ContactBody cb_all[DETECT ? kEnsDetectWarps : 1][DETECT ? SPW * NBP : 1];
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5;
    const long long sys0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * SPW;
    if (sys0 >= g.nsys) return;                  // whole warp exits together
    pdl_wait();
    ens_fast_body<NBP, VM, DETECT>(g, sys0, g.nsteps, g.first != 0, g.last != 0, sxy_all[warp], sz_all[warp],
                                   cb_all[DETECT ? warp : 0]);
}

// ---------------------------------------------------------------------------------------------
// Fast mode, NARROW layout for small batches: ONE body per lane, NBP lanes per system, 32 / NBP systems per warp.
// The two-body layout above needs 64 / NBP x fewer warps per system, which is what a big batch wants (fewer shuffles
// and shared-memory loads per pair); a batch of 8,192 16-body systems (BASELINE configs[3] split over 8 GPUs) then
// has only 3.5 warps per SM sub-partition and cannot hide its FP64 latency (r2a: 87 % of the full-batch rate even
// with the time-sliced kernel).  Here the same batch is twice as many warps with half the work each: every unordered
// pair still once -- ring offsets 1 .. NBP/2, the last one (antipodes) by the lower half only -- own accelerations
// in the lane, reactions in a travelling accumulator that moves one lane per offset and goes home after the last
// (3 SHFL and 2 LDS per pair instead of 1.5 and 1).  Same launch contract (first / last) and rounding of the
// integrator as ens_fast_body; no contact handling (systems with radii take the two-body variant).
template <int NBP, int VM>
__device__ __forceinline__ void ens_fast_body1(const EnsArgs& g, long long sys0, long long nsteps, bool first,
                                               bool last, double2* sxy_w, double* sz_w) {
    constexpr int NS = NBP / 2;                  // ring offsets 1..NS
    constexpr int LM = NBP - 1;
    const int lane = threadIdx.x & 31;
    const int I = lane & LM;
    const int sw = lane / NBP;
    const long long sys = sys0 + sw;
    const int nb = g.nb;
    double2* sxy = sxy_w + sw * NBP;
    double* sz = sz_w + sw * NBP;
    const bool has = sys < g.nsys && I < nb;
    const long long o = sys * nb + I;
    double x = 1e150 * (double)(I + 1), y = 0.0, z = 0.0, m = 0.0;       // padded slot: far away, massless
    double vx = 0.0, vy = 0.0, vz = 0.0, ax = 0.0, ay = 0.0, az = 0.0;
    bool f32 = VM == 1;
    if (has) {
        x = __ldcg(g.x + o); y = __ldcg(g.y + o); z = __ldcg(g.z + o); m = g.m[o];
        vx = __ldcg(g.vx + o); vy = __ldcg(g.vy + o); vz = __ldcg(g.vz + o);
        if (first) { ax = __ldcg(g.ax + o); ay = __ldcg(g.ay + o); az = __ldcg(g.az + o); }
        if (VM == 2) f32 = g.vf32[o] != 0;
    }
    // partner masses do not change: fetch them once (through the z slots)
    double mj[NS];
    sz[I] = m;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const bool vt = (t + 1 < NS) || (I < NS);                        // the upper half skips the antipodes
        mj[t] = vt ? sz[(I + t + 1) & LM] : 0.0;
    }
    const double mi_last = (I < NS) ? m : 0.0;                           // own mass as the antipodal offset sees it
    __syncwarp();
    const double h = g.h, dt = g.dt, G = g.G, eps2 = g.eps2;
    const float dt32 = g.dt32;
    const int group = lane & ~LM;
    if (first && has) {
        vx = ens_kick<VM>(vx, h, ax, f32);                               // engine.py:69-70
        vy = ens_kick<VM>(vy, h, ay, f32);
        vz = ens_kick<VM>(vz, h, az, f32);
    }
    for (long long s = 0; s < nsteps; ++s) {
        if (has) {
            x = ens_drift<VM>(x, vx, dt, dt32, f32);                     // engine.py:73-75
            y = ens_drift<VM>(y, vy, dt, dt32, f32);
            z = ens_drift<VM>(z, vz, dt, dt32, f32);
        }
        sxy[I] = make_double2(x, y);
        sz[I] = z;
        __syncwarp();
        double a0x = 0.0, a0y = 0.0, a0z = 0.0;                          // own acceleration (physics.py:151)
        double cx = 0.0, cy = 0.0, cz = 0.0;                             // travelling: reactions on body I + offset (:152)
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const int J = (I + t + 1) & LM;
            const double2 pxy = sxy[J];
            const double pz = sz[J];
            const double mi = (t + 1 == NS) ? mi_last : m;
            const double dx = pxy.x - x, dy = pxy.y - y, dz = pz - z;
            const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
            int hi;
            const double s0 = inv_r3_plain(r2, hi);
            const double si = s0 * mj[t], sj = s0 * mi;
            a0x = fma(si, dx, a0x); a0y = fma(si, dy, a0y); a0z = fma(si, dz, a0z);
            cx = fma(-sj, dx, cx); cy = fma(-sj, dy, cy); cz = fma(-sj, dz, cz);
            // the next offset meets body I + t + 2, whose accumulator sits one lane up; after the last offset the
            // accumulator of body I + NS goes home
            const int src = group | ((t + 1 < NS ? I + 1 : I - NS) & LM);
            cx = __shfl_sync(0xffffffffu, cx, src);
            cy = __shfl_sync(0xffffffffu, cy, src);
            cz = __shfl_sync(0xffffffffu, cz, src);
        }
        ax = G * (a0x + cx); ay = G * (a0y + cy); az = G * (a0z + cz);
        if (has) {
            vx = ens_kick<VM>(vx, h, ax, f32);                           // engine.py:81-82
            vy = ens_kick<VM>(vy, h, ay, f32);
            vz = ens_kick<VM>(vz, h, az, f32);
            if (s + 1 < nsteps || !last) {
                vx = ens_kick<VM>(vx, h, ax, f32);                       // the next step's first half-kick
                vy = ens_kick<VM>(vy, h, ay, f32);
                vz = ens_kick<VM>(vz, h, az, f32);
            }
        }
        __syncwarp();
    }
    if (has) {
        g.x[o] = x; g.y[o] = y; g.z[o] = z;
        g.vx[o] = vx; g.vy[o] = vy; g.vz[o] = vz;
        if (last) { g.ax[o] = ax; g.ay[o] = ay; g.az[o] = az; }
    }
}

// ORB_ENS_NARROW_MINB CTAs of 4 warps per SM (register cap 65536 / (128 * MINB))
constexpr int kEnsNarrowWarps = 4;
#ifndef ORB_ENS_NARROW_MINB
#define ORB_ENS_NARROW_MINB 5
#endif
template <int NBP, int VM>
__global__ void __launch_bounds__(32 * kEnsNarrowWarps, ORB_ENS_NARROW_MINB) ens_step_fast1_kernel(const EnsArgs g) {
    constexpr int SPW = 32 / NBP;                // systems per warp
    __shared__ double2 sxy_all[kEnsNarrowWarps][32];
    __shared__ double sz_all[kEnsNarrowWarps][32];
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5;
    const long long sys0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * SPW;
    if (sys0 >= g.nsys) return;                  // whole warp exits together
    pdl_wait();
    ens_fast_body1<NBP, VM>(g, sys0, g.nsteps, g.first != 0, g.last != 0, sxy_all[warp], sz_all[warp]);
}

// Time-sliced variant for small batches in fused mode.  With only a few warps of work per SM sub-partition
// (8,192 16-body systems = 3.46 warps per scheduler on 148 SMs) a static assignment leaves the schedulers that got
// 3 warps idle a quarter of the time.  Here a fixed crew of warps (one CTA of 4 per SM slot) pulls (group of SPW
// systems, slice of `slice` steps) items from a queue; slice s of a group is handed out after slice s-1 and waits
// for it (per-group progress counter), the state travels through global memory (L2) in its synchronised
// (x, v, a) form, so the arithmetic -- and the result, bit for bit -- is that of the single pass.
template <int NBP, int VM, bool DETECT>
__global__ void __launch_bounds__(32 * kEnsDetectWarps) ens_fast_sliced_kernel(const EnsArgs g, int slice,
                                                                               unsigned long long* queue,
                                                                               int* progress) {
    constexpr int SPW = 64 / NBP;
    constexpr int SLOTS = SPW * 3 * (NBP / 2);
    __shared__ double2 sxy_all[kEnsDetectWarps][SLOTS];
    __shared__ double sz_all[kEnsDetectWarps][SLOTS];
    __shared__ ContactBody cb_all[DETECT ? kEnsDetectWarps : 1][DETECT ? SPW * NBP : 1];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long ngroups = (g.nsys + SPW - 1) / SPW;
    const long long nseg = (g.nsteps + slice - 1) / slice;
    const unsigned long long total = (unsigned long long)(ngroups * nseg);
    for (;;) {
        unsigned long long id = 0;
        if (lane == 0) id = atomicAdd(queue, 1ull);
        id = __shfl_sync(0xffffffffu, id, 0);
        if (id >= total) break;
        const long long seg = (long long)(id / (unsigned long long)ngroups);
        const long long grp = (long long)(id % (unsigned long long)ngroups);
        if (seg > 0) {
            if (lane == 0) {
                volatile int* p = progress + grp;
                while (*p < (int)seg) __nanosleep(64);
                __threadfence();
            }
            __syncwarp();
        }
        const long long steps = min((long long)slice, g.nsteps - seg * slice);
        ens_fast_body<NBP, VM, DETECT>(g, grp * SPW, steps, true, true, sxy_all[warp], sz_all[warp],
                                       cb_all[DETECT ? warp : 0]);
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicExch(progress + grp, (int)seg + 1);
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

template <bool FAITHFUL, int NBP, int VM, bool DETECT>
static void launch_ens_variant(const EnsArgs& a, int w, cudaStream_t st) {
    if (DETECT && w > kEnsDetectWarps) w = kEnsDetectWarps;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(32 * w);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = a.pdl ? 1 : 0;
    if (FAITHFUL) {
        cfg.gridDim = dim3((unsigned)((a.nsys + w - 1) / w));                   // one warp per system
        cudaLaunchKernelEx(&cfg, ens_step_kernel<NBP, VM, DETECT>, a);
    } else if (!DETECT && a.narrow) {
        w = std::min(w, kEnsNarrowWarps);
        cfg.blockDim = dim3(32 * w);
        const long long per_cta = (long long)w * (32 / NBP);                    // one body per lane
        cfg.gridDim = dim3((unsigned)((a.nsys + per_cta - 1) / per_cta));
        cudaLaunchKernelEx(&cfg, ens_step_fast1_kernel<NBP, VM>, a);
    } else {
        w = std::min(w, kEnsFastWarps);
        cfg.blockDim = dim3(32 * w);
        const long long per_cta = (long long)w * (64 / NBP);                    // 64/NBP systems per warp
        cfg.gridDim = dim3((unsigned)((a.nsys + per_cta - 1) / per_cta));
        cudaLaunchKernelEx(&cfg, ens_step_fast_kernel<NBP, VM, DETECT>, a);
    }
}

template <bool FAITHFUL, int NBP>
static void launch_ens_step_t(const EnsArgs& a, int w, cudaStream_t st) {
    if (a.radius)                     // contacts: always with per-body dtype flags (one variant carries the sweep)
        launch_ens_variant<FAITHFUL, NBP, 2, true>(a, w, st);
    else if (a.vf32)
        launch_ens_variant<FAITHFUL, NBP, 2, false>(a, w, st);
    else if (a.vel_f32)
        launch_ens_variant<FAITHFUL, NBP, 1, false>(a, w, st);
    else
        launch_ens_variant<FAITHFUL, NBP, 0, false>(a, w, st);
}

cudaError_t launch_ens_step(const EnsArgs& a, bool faithful, cudaStream_t st) {
    int w = a.warps_per_cta;
    if (w < 1) w = 1;
    if (w > kEnsMaxWarps) w = kEnsMaxWarps;
    if (a.radius && !a.vf32) return cudaErrorInvalidValue;       // the contact variant reads the dtype flags
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

// Fused fast mode for a small batch: the time-sliced kernel (see ens_fast_sliced_kernel).
template <int NBP>
static cudaError_t launch_sliced_t(const EnsArgs& a, int slice, unsigned long long* queue, int* progress, int sm_count,
                                   cudaStream_t st) {
    const int block = 32 * kEnsDetectWarps;
    int per_sm = 0;
    const char* env = getenv("ORBITAL_B200_ENS_CREW");       // CTAs per SM of the crew (0: all that fit)
    const int crew = env ? atoi(env) : 3;                     // measured: 3 CTAs (12 warps) per SM beat 4 and 2
#define ORB_SLICED(VMv, DETv)                                                                                   \
    do {                                                                                                        \
        auto kern = ens_fast_sliced_kernel<NBP, VMv, DETv>;                                                     \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, 0) != cudaSuccess || per_sm < 1) \
            per_sm = 1;                                                                                         \
        if (crew > 0 && crew < per_sm) per_sm = crew;                                                           \
        kern<<<sm_count * per_sm, block, 0, st>>>(a, slice, queue, progress);                                   \
    } while (0)
    if (a.radius) ORB_SLICED(2, true);
    else if (a.vf32) ORB_SLICED(2, false);
    else if (a.vel_f32) ORB_SLICED(1, false);
    else ORB_SLICED(0, false);
#undef ORB_SLICED
    return cudaGetLastError();
}

cudaError_t launch_ens_step_sliced(const EnsArgs& a, int slice, unsigned long long* queue, int* progress, int sm_count,
                                   cudaStream_t st) {
    if (a.radius && !a.vf32) return cudaErrorInvalidValue;
    switch (a.nbp) {
        case 2: return launch_sliced_t<2>(a, slice, queue, progress, sm_count, st);
        case 4: return launch_sliced_t<4>(a, slice, queue, progress, sm_count, st);
        case 8: return launch_sliced_t<8>(a, slice, queue, progress, sm_count, st);
        case 16: return launch_sliced_t<16>(a, slice, queue, progress, sm_count, st);
        case 32: return launch_sliced_t<32>(a, slice, queue, progress, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
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
