// integrate.cu -- leapfrog update kernels, the fused single-CTA multi-step kernel,
// history ring writer and diagnostics reductions (sm_100a).
//
// Replaces SimulationEngine.step phases 1,2,4,6 (reference core/engine.py:69-75,
// 81-82,88-92) and total_energy / angular_momentum (core/engine.py:104-121).
// The arithmetic is the reference's exact rounding sequence in
// both engine modes: these passes are O(N) and HBM-trivial next to the force pass.
#include <algorithm>
#include <cstdio>

#include "contacts.cuh"
#include <cstdlib>

#include "kernels.h"

namespace orb {

// ---------------------------------------------------------------------------
// engine.py:69-75  half-kick + drift of the local targets; writes the packed
// source array the force pass (and the multi-GPU all-gather) reads.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kick_drift_kernel(double4* pos4, double* vel, const double* __restrict__ acc,
                                                         const uint8_t* __restrict__ vf32, long long n,
                                                         long long lo, long long hi, double h, double dt,
                                                         float dt32, const Ctl* ctl) {
    if (ctl->halted) return;
    const long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const bool f32 = vf32[i] != 0;
    const double vx = kick_faithful(vel[i], h, acc[i], f32);
    const double vy = kick_faithful(vel[i + n], h, acc[i + n], f32);
    const double vz = kick_faithful(vel[i + 2 * n], h, acc[i + 2 * n], f32);
    vel[i] = vx; vel[i + n] = vy; vel[i + 2 * n] = vz;
    double4 p = pos4[i];
    p.x = drift_faithful(p.x, vx, dt, dt32, f32);
    p.y = drift_faithful(p.y, vy, dt, dt32, f32);
    p.z = drift_faithful(p.z, vz, dt, dt32, f32);
    pos4[i] = p;
}

// engine.py:81-82 second half-kick, then engine.py:88-92 history append (skipped
// in a halting step: the host appends after resolving the contacts).
__global__ void __launch_bounds__(256) kick_hist_kernel(const double4* __restrict__ pos4, double* vel,
                                                        const double* __restrict__ acc,
                                                        const uint8_t* __restrict__ vf32, long long n, long long lo,
                                                        long long hi, double h, double* hist, long long hist_cap,
                                                        const Ctl* ctl) {
    if (ctl->halted) return;
    const long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const bool f32 = vf32[i] != 0;
    vel[i] = kick_faithful(vel[i], h, acc[i], f32);
    vel[i + n] = kick_faithful(vel[i + n], h, acc[i + n], f32);
    vel[i + 2 * n] = kick_faithful(vel[i + 2 * n], h, acc[i + 2 * n], f32);
    if (hist_cap > 0 && ctl->overlap_count == 0) {
        const double4 p = pos4[i];
        double* row = hist + ((ctl->hist_count % hist_cap) * n + i) * 3;
        row[0] = p.x; row[1] = p.y; row[2] = p.z;
    }
}

__global__ void advance_kernel(Ctl* ctl, long long hist_cap) {
    if (ctl->halted) return;
    ctl->steps_done += 1;
    if (ctl->overlap_count > 0)
        ctl->halted = 1;
    else if (hist_cap > 0)
        ctl->hist_count += 1;
}

// ---- device-side contacts (physics.py:510-535 after the second half-kick, engine.py:85) -------------------
// All four kernels are launched every step of an engine that can have contacts and return at once when the
// force pass flagged nothing.

// U belongs to the force build: stash it before the push-out moves bodies (only where `last_potential` is
// evaluated in reference order: faithful mode, n <= 4096).
__global__ void __launch_bounds__(256) contacts_stash_u_kernel(const double4* __restrict__ pos4, int n, double eps2,
                                                               double G, Ctl* ctl) {
    __shared__ double terms[256];
    if (ctl->halted) return;
    if (ctl->overlap_count == 0) {
        if (threadIdx.x == 0) ctl->u_valid = 0;
        return;
    }
    const double U = potential_ordered_block(pos4, n, eps2, G, terms);
    if (threadIdx.x == 0) { ctl->u_stash = U; ctl->u_valid = 1; }
}

__global__ void __launch_bounds__(256) contacts_resolve_kernel(double4* pos4, double* vel, long long n,
                                                               const double* __restrict__ radius,
                                                               const uint8_t* __restrict__ vf32, double restitution,
                                                               Ctl* ctl, long long* pairs) {
    __shared__ unsigned long long scratch[4];
    if (ctl->halted || ctl->overlap_count == 0) return;
    const int hits = resolve_contacts_block(pos4, vel, n, radius, vf32, restitution, ctl, pairs, scratch);
    if (threadIdx.x == 0) ctl->contacts_total += hits;
}

// engine.py:88-92 for a step with contacts (kick_hist_kernel skipped it): positions after the sweep
__global__ void __launch_bounds__(256) contacts_hist_kernel(const double4* __restrict__ pos4, long long n,
                                                            double* hist, long long hist_cap, const Ctl* ctl) {
    if (ctl->halted || ctl->overlap_count == 0 || hist_cap <= 0) return;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 p = pos4[i];
    double* row = hist + ((ctl->hist_count % hist_cap) * n + i) * 3;
    row[0] = p.x; row[1] = p.y; row[2] = p.z;
}

__global__ void advance_contacts_kernel(Ctl* ctl, long long hist_cap) {
    if (ctl->halted) return;
    ctl->steps_done += 1;
    if (hist_cap > 0) ctl->hist_count += 1;
    ctl->overlap_count = 0;                      // the next force pass starts a fresh list
    ctl->overlap_overflow = 0;
}

cudaError_t launch_contacts(const DeviceState& s, const StepParams& p, bool ordered_potential, cudaStream_t st,
                            int* launches) {
    if (ordered_potential) {
        contacts_stash_u_kernel<<<1, 256, 0, st>>>(s.pos4, (int)s.n, p.eps2, p.G, s.ctl);
        if (launches) ++*launches;
    }
    contacts_resolve_kernel<<<1, 256, 0, st>>>(s.pos4, s.vel, s.n, s.radius, s.vf32, p.restitution, s.ctl, s.pairs);
    if (s.hist_cap > 0) {
        contacts_hist_kernel<<<(unsigned)((s.n + 255) / 256), 256, 0, st>>>(s.pos4, s.n, s.hist, s.hist_cap, s.ctl);
        if (launches) ++*launches;
    }
    advance_contacts_kernel<<<1, 1, 0, st>>>(s.ctl, s.hist_cap);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) hist_append_kernel(const double4* __restrict__ pos4, long long n,
                                                          double* hist, long long hist_cap, const Ctl* ctl) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || hist_cap <= 0) return;
    const double4 p = pos4[i];
    double* row = hist + ((ctl->hist_count % hist_cap) * n + i) * 3;
    row[0] = p.x; row[1] = p.y; row[2] = p.z;
}

__global__ void hist_bump_kernel(Ctl* ctl) { ctl->hist_count += 1; }

static inline int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

cudaError_t launch_kick_drift(const DeviceState& s, const StepParams& p, cudaStream_t st) {
    const long long nt = s.tgt_hi - s.tgt_lo;
    kick_drift_kernel<<<grid_for(nt, 256), 256, 0, st>>>(s.pos4, s.vel, s.acc, s.vf32, s.n, s.tgt_lo, s.tgt_hi, p.h,
                                                         p.dt, p.dt32, s.ctl);
    return cudaGetLastError();
}

cudaError_t launch_kick_hist(const DeviceState& s, const StepParams& p, cudaStream_t st) {
    const long long nt = s.tgt_hi - s.tgt_lo;
    // a sharded handle appends the whole snapshot in launch_step_end, after the contacts of all ranks are resolved
    const bool sharded = !(s.tgt_lo == 0 && s.tgt_hi == s.n);
    kick_hist_kernel<<<grid_for(nt, 256), 256, 0, st>>>(s.pos4, s.vel, s.acc, s.vf32, s.n, s.tgt_lo, s.tgt_hi, p.h,
                                                        s.hist, sharded ? 0 : s.hist_cap, s.ctl);
    return cudaGetLastError();
}

// End of a sharded step (engine.py:85-92 once every rank holds the same positions, velocities and pair list):
// replicated contact sweep, history append of ALL n bodies, step bookkeeping.
cudaError_t launch_step_end(const DeviceState& s, const StepParams& p, bool resolve, bool ordered_potential,
                            cudaStream_t st, int* launches) {
    if (resolve) {
        if (ordered_potential) {
            contacts_stash_u_kernel<<<1, 256, 0, st>>>(s.pos4, (int)s.n, p.eps2, p.G, s.ctl);
            if (launches) ++*launches;
        }
        contacts_resolve_kernel<<<1, 256, 0, st>>>(s.pos4, s.vel, s.n, s.radius, s.vf32, p.restitution, s.ctl, s.pairs);
        if (launches) ++*launches;
    }
    if (s.hist_cap > 0) {
        hist_append_kernel<<<grid_for(s.n, 256), 256, 0, st>>>(s.pos4, s.n, s.hist, s.hist_cap, s.ctl);
        if (launches) ++*launches;
    }
    advance_contacts_kernel<<<1, 1, 0, st>>>(s.ctl, s.hist_cap);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// the merged pair list of all ranks was uploaded: make it this handle's list
__global__ void ctl_set_overlaps_kernel(Ctl* ctl, int count, int overflow) {
    ctl->overlap_count = count;
    ctl->overlap_overflow = overflow;
}

cudaError_t launch_set_overlaps(const DeviceState& s, int count, int overflow, cudaStream_t st) {
    ctl_set_overlaps_kernel<<<1, 1, 0, st>>>(s.ctl, count, overflow);
    return cudaGetLastError();
}

cudaError_t launch_advance(const DeviceState& s, cudaStream_t st) {
    advance_kernel<<<1, 1, 0, st>>>(s.ctl, s.hist_cap);
    return cudaGetLastError();
}

cudaError_t launch_hist_append(const DeviceState& s, cudaStream_t st) {
    if (s.hist_cap <= 0) return cudaSuccess;
    hist_append_kernel<<<grid_for(s.n, 256), 256, 0, st>>>(s.pos4, s.n, s.hist, s.hist_cap, s.ctl);
    hist_bump_kernel<<<1, 1, 0, st>>>(s.ctl);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Fused small-system kernel: ONE CTA integrates the whole system for nsteps
// steps (engine.py:65-97 per step) with all state in registers / shared memory.
// Thread i owns body i for the kicks and the drift.  The force pass is spread over
// the whole CTA: a warp takes a target, its lanes evaluate the pair terms of 32
// sources at once (IEEE sqrt + two divides each: the long-latency part), and the
// terms are then added in ascending source order -- the reference's accumulation
// order, so the result is bit-exact -- via warp shuffles.  Two block barriers per
// step; the solar-system configs run 10,000 steps in a single launch.
// ---------------------------------------------------------------------------
template <bool DETECT>
__global__ void __launch_bounds__(512, 1) tiny_steps_kernel(double4* pos4, double* vel, double* acc,
                                                          const double* __restrict__ radius,
                                                          const uint8_t* __restrict__ vf32, int n, long long nsteps,
                                                          double h, double dt, float dt32, double eps2, double G,
                                                          double* hist, long long hist_cap, double restitution, int device_contacts, int stash_u, Ctl* ctl,
                                                          long long* pairs) {
    if (ctl->halted) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double4* sp = reinterpret_cast<double4*>(smem_raw);     // {x,y,z,G*m}
    double* sr = reinterpret_cast<double*>(sp + n);         // radius
    double* sa = sr + n;                                    // 3 x n accelerations of the current force build
    double* cterms = sa + 3 * n;                            // contact path: 512 doubles + 4 words of scratch
    unsigned long long* cscratch = reinterpret_cast<unsigned long long*>(cterms + 512);
    bool u_set = false;
    if (threadIdx.x == 0) ctl->u_valid = 0;
    const int i = threadIdx.x;
    const int lane = i & 31;
    const int warp = i >> 5;
    const int nwarps = blockDim.x >> 5;
    const bool active = i < n;
    double x = 0, y = 0, z = 0, gm = 0, vx = 0, vy = 0, vz = 0, ax = 0, ay = 0, az = 0;
    bool f32 = false;
    if (active) {
        const double4 p = pos4[i];
        x = p.x; y = p.y; z = p.z;
        gm = __dmul_rn(G, p.w);                              // G * mj (physics.py:151)
        vx = vel[i]; vy = vel[i + n]; vz = vel[i + 2 * n];
        ax = acc[i]; ay = acc[i + n]; az = acc[i + 2 * n];
        f32 = vf32[i] != 0;
        sr[i] = radius[i];
    }
    const long long hist0 = ctl->hist_count;
    long long slot = hist_cap > 0 ? hist0 % hist_cap : 0;    // ring cursor, advanced without a 64-bit modulo per step
    long long done = 0;
    for (long long s = 0; s < nsteps; ++s) {
        if (active) {
            vx = kick_faithful(vx, h, ax, f32);                      // engine.py:69-70
            vy = kick_faithful(vy, h, ay, f32);
            vz = kick_faithful(vz, h, az, f32);
            x = drift_faithful(x, vx, dt, dt32, f32);                // engine.py:73-75
            y = drift_faithful(y, vy, dt, dt32, f32);
            z = drift_faithful(z, vz, dt, dt32, f32);
            sp[i] = make_double4(x, y, z, gm);
        }
        __syncthreads();
        int hit = 0;
        for (int t = warp; t < n; t += nwarps) {                     // warp-uniform: one target per warp pass
            const double4 me = sp[t];
            const double Rt = DETECT ? sr[t] : 0.0;
            double bx = 0.0, by = 0.0, bz = 0.0;                     // physics.py:132 (replicated in every lane)
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                double tx = 0.0, ty = 0.0, tz = 0.0;
                if (j < n && j != t) {
                    const double4 q = sp[j];
                    const double dx = __dsub_rn(q.x, me.x), dy = __dsub_rn(q.y, me.y), dz = __dsub_rn(q.z, me.z);
                    pair_term_faithful(dx, dy, dz, eps2, q.w, tx, ty, tz);
                    if (DETECT && j > t) {
                        if (overlap_exact(-dx, -dy, -dz, Rt, sr[j])) {
                            record_overlap(ctl, pairs, t, j);
                            hit = 1;
                        }
                    }
                }
                const int cnt = min(32, n - j0);
#pragma unroll 4
                for (int l = 0; l < cnt; ++l) {                      // ascending j: the reference's order
                    const double px = __shfl_sync(0xffffffffu, tx, l);
                    const double py = __shfl_sync(0xffffffffu, ty, l);
                    const double pz = __shfl_sync(0xffffffffu, tz, l);
                    if (j0 + l != t) {                               // uniform: the self pair is skipped, not added
                        bx = __dadd_rn(bx, px);
                        by = __dadd_rn(by, py);
                        bz = __dadd_rn(bz, pz);
                    }
                }
            }
            if (lane == 0) {
                sa[t] = bx; sa[n + t] = by; sa[2 * n + t] = bz;
            }
        }
        const int hits = DETECT ? __syncthreads_or(hit) : (__syncthreads(), 0);
        if (active) {
            ax = sa[i]; ay = sa[n + i]; az = sa[2 * n + i];
            vx = kick_faithful(vx, h, ax, f32);                      // engine.py:81-82
            vy = kick_faithful(vy, h, ay, f32);
            vz = kick_faithful(vz, h, az, f32);
        }
        ++done;
        if (hits) {
            if (!device_contacts) break;                             // the host resolves and resumes
            // engine.py:85: contacts after the second half-kick, resolved here in the reference's order
            if (active) {
                pos4[i] = make_double4(x, y, z, pos4[i].w);
                vel[i] = vx; vel[i + n] = vy; vel[i + 2 * n] = vz;
            }
            __threadfence_block();
            __syncthreads();
            if (stash_u) {                                           // U of this force build (pre push-out)
                const double U = potential_ordered_block(pos4, n, eps2, G, cterms);
                if (i == 0) { ctl->u_stash = U; ctl->u_valid = 1; }
                u_set = true;
            }
            const int nh = resolve_contacts_block(pos4, vel, n, radius, vf32, restitution, ctl, pairs, cscratch);
            if (active) {
                const double4 p = pos4[i];
                x = p.x; y = p.y; z = p.z;
                vx = vel[i]; vy = vel[i + n]; vz = vel[i + 2 * n];
            }
            if (i == 0) { ctl->contacts_total += nh; ctl->overlap_count = 0; }
            __syncthreads();
        } else if (u_set) {
            if (i == 0) ctl->u_valid = 0;
            u_set = false;
        }
        if (active && hist_cap > 0) {                                // engine.py:88-92
            double* row = hist + (slot * n + i) * 3;
            row[0] = x; row[1] = y; row[2] = z;
        }
        if (++slot >= hist_cap) slot = 0;
    }
    if (active) {
        pos4[i] = make_double4(x, y, z, pos4[i].w);
        vel[i] = vx; vel[i + n] = vy; vel[i + 2 * n] = vz;
        acc[i] = ax; acc[i + n] = ay; acc[i + 2 * n] = az;
    }
    if (i == 0) {
        ctl->steps_done += done;
        if (ctl->overlap_count > 0) {
            ctl->halted = 1;
            if (hist_cap > 0) ctl->hist_count = hist0 + (done - 1);
        } else if (hist_cap > 0) {
            ctl->hist_count = hist0 + done;
        }
    }
}

// ---------------------------------------------------------------------------
// Micro-system kernel (n <= kMicroMax, e.g. the solar-system configs): same contract as
// tiny_steps_kernel, but every unordered pair is evaluated ONCE per step -- the reference's own
// half-matrix loop (physics.py:136-155) -- and both of its terms are parked in a shared-memory
// matrix T[target][source]; one lane per target then adds its row in ascending source order.
//   pair (i<j): d = rj - ri, inv_r3 shared;   T[i][j] = (G mj inv_r3) d
//                                            T[j][i] = -((G mi inv_r3) d)   == (G mi inv_r3)(ri - rj) bit for bit
// (IEEE negation is exact and rounding is sign-symmetric).  The expensive IEEE sqrt + two divides are
// paid n(n-1)/2 times instead of n(n-1), spread over all lanes of the CTA.
// ---------------------------------------------------------------------------
constexpr int kMicroMax = 64;

// One CTA for the whole system saves every launch but uses one SM.  Measured (profiles/r1_sweep_tiny.txt, us per
// step): micro_steps_kernel 0.76 (n=15) ... 3.3 (n=64) against ~10 for the multi-CTA kernel sequence (two-pass force
// under a CUDA graph), but tiny_steps_kernel 25 (n=65) ... 1256 (n=512) against 10.7 ... 19.4: the fused path is the
// default only up to kMicroMax.  ORBITAL_B200_TINY_MAX (<= kTinyMax) moves the limit (cross-checks).
int tiny_limit() {
    const char* env = getenv("ORBITAL_B200_TINY_MAX");
    const int v = env ? atoi(env) : kMicroMax;
    return v < 0 ? 0 : (v > kTinyMax ? kTinyMax : v);
}

// ordered sum of 8 * NB row entries starting from +0.0 (ascending j, physics.py:154)
template <int NB>
__device__ __forceinline__ double row_sum_blocks(const double* row) {
    double v[8 * NB];
#pragma unroll
    for (int j = 0; j < 8 * NB; ++j) v[j] = row[j];
    double b = 0.0;
#pragma unroll
    for (int j = 0; j < 8 * NB; ++j) b = __dadd_rn(b, v[j]);
    return b;
}

template <bool DETECT>
__global__ void __launch_bounds__(256, 1) micro_steps_kernel(double4* pos4, double* vel, double* acc,
                                                             const double* __restrict__ radius,
                                                             const uint8_t* __restrict__ vf32, int n,
                                                             long long nsteps, double h, double dt, float dt32,
                                                             double eps2, double G, double* hist, long long hist_cap,
                                                             double restitution, int device_contacts, int stash_u, Ctl* ctl, long long* pairs) {
    if (ctl->halted) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int npad = (n + 7) & ~7;                          // row sums run in whole blocks of 8 (padding: +0.0)
    const int stride = npad | 1;                            // odd row stride: conflict-free column walks
    double4* sp = reinterpret_cast<double4*>(smem_raw);     // {x,y,z,G*m}
    double* sr = reinterpret_cast<double*>(sp + n);         // radius
    double* T = sr + n;                                     // 3 x n x stride
    double* cterms = T + 3 * n * stride;                    // contact path: 256 doubles + 4 words of scratch
    unsigned long long* cscratch = reinterpret_cast<unsigned long long*>(cterms + 256);
    unsigned short* plist = reinterpret_cast<unsigned short*>(cscratch + 4);   // (i << 8) | j, i < j
    bool u_set = false;
    __shared__ volatile int hitflag[2];
    if (threadIdx.x == 0) { ctl->u_valid = 0; hitflag[0] = 0; hitflag[1] = 0; }
    const int tid = threadIdx.x;
    const int npairs = n * (n - 1) / 2;
    for (int p = tid; p < npairs; p += blockDim.x) {        // lexicographic pair list: p -> (i, j), i < j
        // row i starts at offset i(2n - i - 1)/2: invert with a float sqrt, then correct by at most one
        int i = (int)((2.0f * n - 1.0f - sqrtf((2.0f * n - 1.0f) * (2.0f * n - 1.0f) - 8.0f * p)) * 0.5f);
        i = max(0, min(i, n - 2));
        while (i > 0 && i * (2 * n - i - 1) / 2 > p) --i;
        while ((i + 1) * (2 * n - i - 2) / 2 <= p) ++i;
        const int j = p - i * (2 * n - i - 1) / 2 + i + 1;
        plist[p] = (unsigned short)((i << 8) | j);
    }
    // Integrator and row sums: one thread per (body, COMPONENT), the three components of a body in three different
    // warps.  The step is one dependent chain (kick -> drift -> pairs -> ordered row sum -> kick); with a thread per
    // body, one warp issued ~190 instructions of it per step, now each of three warps issues a third (ncu, N = 15:
    // 373 instructions per step on warp 0, issue slots 12 % busy, everything stalled on the previous result).
    const int S = n <= 32 ? 32 : 64;                        // threads per component plane
    const int comp = tid / S, body = tid - comp * S;
    const bool active = comp < 3 && body < n;
    const int plane = n * stride;
    double q = 0, v = 0, a = 0, gm = 0;                     // this thread's component of position / velocity / acc
    bool f32 = false;
    if (active) {
        const double4 p = pos4[body];
        q = comp == 0 ? p.x : (comp == 1 ? p.y : p.z);
        gm = __dmul_rn(G, p.w);                              // G * m (physics.py:151-152)
        v = vel[body + comp * n];
        a = acc[body + comp * n];
        f32 = vf32[body] != 0;
        if (comp == 0) sr[body] = radius[body];
        T[comp * plane + body * stride + body] = 0.0;        // diagonal of the three term planes: +0.0
        for (int j = n; j < npad; ++j) T[comp * plane + body * stride + j] = 0.0;   // and the padding of the row
    }
    double* spq = reinterpret_cast<double*>(sp + body) + comp;       // this thread's slot of sp[body]
    __syncthreads();
    const int first_pair = tid < npairs ? plist[tid] : 0;   // most systems: at most one pair per lane
    const int fi = first_pair >> 8, fj = first_pair & 0xff;
    double* const fTi = T + fi * stride + fj;
    double* const fTj = T + fj * stride + fi;
    const double fRi = DETECT ? sr[fi] : 0.0, fRj = DETECT ? sr[fj] : 0.0;     // radii do not change inside a launch
    const long long hist0 = ctl->hist_count;
    long long slot = hist_cap > 0 ? hist0 % hist_cap : 0;    // ring cursor, advanced without a 64-bit modulo per step
    const long long hstep = 3LL * n;                          // ... and this thread's slot in it, advanced by addition
    double* hp = hist_cap > 0 ? hist + (slot * n + body) * 3 + comp : nullptr;
    long long done = 0;
    if (active && comp == 0) sp[body].w = gm;
    // -DORB_MICRO_PROFILE: cycles per phase and warp, printed at the end of a launch (development aid; build a second
    // library with `make OUT=... BUILD=... EXTRA=-DORB_MICRO_PROFILE` and point ORBITAL_B200_LIB at it)
#ifdef ORB_MICRO_PROFILE
    long long tA = 0, tB = 0, tC = 0, tD = 0, t0 = clock64(), t1;
#define ORB_MP(acc) { t1 = clock64(); acc += t1 - t0; t0 = t1; }
#else
#define ORB_MP(acc)
#endif
    for (long long s = 0; s < nsteps; ++s) {
        if (active) {
            // kick_faithful + drift_faithful (common.cuh) with the float32 velocity kept from the kick: the drift's
            // float32(v) is that very value, so one conversion less sits on the step's dependent chain
            const double r = __dadd_rn(v, __dmul_rn(h, a));          // engine.py:69-70
            if (f32) {
                const float vf = __double2float_rn(r);
                v = (double)vf;
                q = __dadd_rn(q, (double)__fmul_rn(vf, dt32));       // engine.py:73-75, product in float32
            } else {
                v = r;
                q = __dadd_rn(q, __dmul_rn(v, dt));
            }
            *spq = q;
        }
        __syncthreads();
        ORB_MP(tA)
        // physics.py:136-155, one pair per lane: the lane's first pair with everything that does not change between
        // steps (slots, radii, term addresses) held in registers; further pairs (n > 23 at 256 threads) by index
        auto pair = [&](const double4* ppi, const double4* ppj, double Ri, double Rj, double* Ti, double* Tj, int i,
                        int j) {
            const double4 pi = *ppi, pj = *ppj;
            const double dx = __dsub_rn(pj.x, pi.x), dy = __dsub_rn(pj.y, pi.y), dz = __dsub_rn(pj.z, pi.z);  // :145
            // physics.py:517-518 (ri - rj): tested up front so that it overlaps the sqrt / divide chain below
            const bool touching = DETECT && overlap_exact(-dx, -dy, -dz, Ri, Rj);
            const double r2 = __dadd_rn(dot3_numpy(dx, dy, dz), eps2);                                        // :146
            const double inv_r = __ddiv_rn(1.0, __dsqrt_rn(r2));                                              // :147
            const double inv_r3 = __ddiv_rn(inv_r, r2);                                                       // :148
            const double si = __dmul_rn(pj.w, inv_r3);               // (G mj) inv_r3          :151
            const double sj = __dmul_rn(pi.w, inv_r3);               // (G mi) inv_r3          :152 (sign applied below)
            Ti[0] = __dmul_rn(si, dx); Ti[plane] = __dmul_rn(si, dy); Ti[2 * plane] = __dmul_rn(si, dz);
            Tj[0] = -__dmul_rn(sj, dx); Tj[plane] = -__dmul_rn(sj, dy); Tj[2 * plane] = -__dmul_rn(sj, dz);
            if (touching) {
                record_overlap(ctl, pairs, i, j);
                hitflag[s & 1] = 1;
            }
        };
        if (tid < npairs) pair(sp + fi, sp + fj, fRi, fRj, fTi, fTj, fi, fj);
        for (int p = tid + blockDim.x; p < npairs; p += blockDim.x) {
            const int code = plist[p];
            const int i = code >> 8, j = code & 0xff;
            pair(sp + i, sp + j, DETECT ? sr[i] : 0.0, DETECT ? sr[j] : 0.0, T + i * stride + j, T + j * stride + i, i, j);
        }
        __syncthreads();
        ORB_MP(tB)
        // two flags, alternating by step: the one of the next step is cleared while nobody can be setting it
        const int hits = DETECT ? hitflag[s & 1] : 0;
        if (DETECT && tid == 0) hitflag[(s + 1) & 1] = 0;
        if (active) {
            const double* row = T + comp * plane + body * stride;
            double b = 0.0;                                          // physics.py:132
            // ascending j.  The diagonal entry holds +0.0: x + 0.0 == x bit for bit because a running sum that
            // starts at +0.0 can never be -0.0, so the self pair is "skipped" without a branch and the
            // shared-memory loads pipeline freely.  The row is padded with +0.0 to a multiple of 8 for the same
            // reason: whole unrolled blocks, no remainder loop (r2b: clock64 per phase at N = 15 showed this phase at
            // 610 of the step's 1,900 cycles -- seven branchy remainder iterations around 14 dependent additions).
            // Up to 32 bodies the whole row is one straight-line block: all loads issue before the first addition, so
            // the chain pays one shared-memory latency, not one per block of 8 (N = 15: 490 -> 2xx cycles).
            // (uniform compare-and-branch chain: a `switch` becomes a jump table -- constant load + indirect branch)
            if (npad == 16) b = row_sum_blocks<2>(row);
            else if (npad == 8) b = row_sum_blocks<1>(row);
            else if (npad == 24) b = row_sum_blocks<3>(row);
            else if (npad == 32) b = row_sum_blocks<4>(row);
            else {
                for (int j0 = 0; j0 < npad; j0 += 8) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) b = __dadd_rn(b, row[j0 + j]);
                }
            }
            a = b;
            v = kick_faithful(v, h, a, f32);                         // engine.py:81-82
        }
        ++done;
        ORB_MP(tC)
        if (hits) {
            if (!device_contacts) break;                             // the host resolves and resumes
            // engine.py:85: contacts after the second half-kick, resolved here in the reference's order
            if (active) {
                reinterpret_cast<double*>(pos4 + body)[comp] = q;
                vel[body + comp * n] = v;
            }
            __threadfence_block();
            __syncthreads();
            if (stash_u) {                                           // U of this force build (pre push-out)
                const double U = potential_ordered_block(pos4, n, eps2, G, cterms);
                if (tid == 0) { ctl->u_stash = U; ctl->u_valid = 1; }
                u_set = true;
            }
            const int nh = resolve_contacts_block(pos4, vel, n, radius, vf32, restitution, ctl, pairs, cscratch);
            if (active) {
                q = reinterpret_cast<const double*>(pos4 + body)[comp];
                v = vel[body + comp * n];
            }
            if (tid == 0) { ctl->contacts_total += nh; ctl->overlap_count = 0; ctl->overlap_overflow = 0; }
            __syncthreads();
        } else if (u_set) {
            if (tid == 0) ctl->u_valid = 0;
            u_set = false;
        }
        if (active && hist_cap > 0) *hp = q;                                  // engine.py:88-92
        hp += hstep;
        if (++slot >= hist_cap) { slot = 0; hp -= hist_cap * hstep; }
        ORB_MP(tD)
    }
#ifdef ORB_MICRO_PROFILE
    if ((tid & 31) == 0 && nsteps >= 100)
        printf("micro profile warp %d: cycles per step  kick+drift+bar %.0f  pairs+bar %.0f  rowsum+kick %.0f  tail %.0f\n",
               tid >> 5, (double)tA / nsteps, (double)tB / nsteps, (double)tC / nsteps, (double)tD / nsteps);
#endif
    if (active) {
        reinterpret_cast<double*>(pos4 + body)[comp] = q;
        vel[body + comp * n] = v;
        acc[body + comp * n] = a;
    }
    if (tid == 0) {
        ctl->steps_done += done;
        if (ctl->overlap_count > 0) {
            ctl->halted = 1;
            if (hist_cap > 0) ctl->hist_count = hist0 + (done - 1);
        } else if (hist_cap > 0) {
            ctl->hist_count = hist0 + done;
        }
    }
}

static int micro_block(int n) {
    const int npairs = n * (n - 1) / 2;
    const int want = std::max(3 * (n <= 32 ? 32 : 64), npairs);     // a thread per (body, component); a pair per lane
    return std::max(96, std::min(256, ((want + 31) / 32) * 32));
}

static size_t micro_smem(int n) {
    const int stride = ((n + 7) & ~7) | 1;
    return (size_t)n * (sizeof(double4) + sizeof(double)) + (size_t)3 * n * stride * sizeof(double) +
           (256 + 4) * sizeof(double) + (size_t)(n * (n - 1) / 2 + 8) * sizeof(unsigned short) + 32;
}

// one thread per body for the integrator, one warp per target (up to 16 warps) for the force pass
int tiny_block(int n) {
    if (n <= kMicroMax) return micro_block(n);
    const int threads_bodies = ((n + 31) / 32) * 32;
    const int threads_force = 32 * std::min(16, n);
    return std::max(32, std::min(512, std::max(threads_bodies, threads_force)));
}

cudaError_t launch_tiny_steps(const DeviceState& s, const StepParams& p, long long nsteps, bool detect,
                              cudaStream_t st) {
    const int n = (int)s.n;
    const int block = tiny_block(n);
    const int stash = 1;          // the fused kernels only run in faithful mode: `last_potential` is reference-ordered
    if (n <= kMicroMax) {
        const size_t msmem = micro_smem(n);
        static DeviceOnce attr_set;                             // the attribute is per device
        if (attr_set.first()) {
            const int cap = (int)micro_smem(kMicroMax);
            cudaFuncSetAttribute(micro_steps_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
            cudaFuncSetAttribute(micro_steps_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        }
        if (detect)
            micro_steps_kernel<true><<<1, block, msmem, st>>>(s.pos4, s.vel, s.acc, s.radius, s.vf32, n, nsteps, p.h,
                                                              p.dt, p.dt32, p.eps2, p.G, s.hist, s.hist_cap, p.restitution,
                                                              p.device_contacts, stash, s.ctl, s.pairs);
        else
            micro_steps_kernel<false><<<1, block, msmem, st>>>(s.pos4, s.vel, s.acc, s.radius, s.vf32, n, nsteps, p.h,
                                                               p.dt, p.dt32, p.eps2, p.G, s.hist, s.hist_cap, p.restitution,
                                                               p.device_contacts, stash, s.ctl, s.pairs);
        return cudaGetLastError();
    }
    const size_t smem = (size_t)n * (sizeof(double4) + 4 * sizeof(double)) + (512 + 4) * sizeof(double) + 16;
    if (detect)
        tiny_steps_kernel<true><<<1, block, smem, st>>>(s.pos4, s.vel, s.acc, s.radius, s.vf32, n, nsteps, p.h, p.dt,
                                                        p.dt32, p.eps2, p.G, s.hist, s.hist_cap, p.restitution, p.device_contacts, stash,
                                                        s.ctl, s.pairs);
    else
        tiny_steps_kernel<false><<<1, block, smem, st>>>(s.pos4, s.vel, s.acc, s.radius, s.vf32, n, nsteps, p.h, p.dt,
                                                         p.dt32, p.eps2, p.G, s.hist, s.hist_cap, p.restitution, p.device_contacts, stash,
                                                        s.ctl, s.pairs);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// SoA <-> packed {x,y,z,m}
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_kernel(double4* pos4, const double* __restrict__ x,
                                                   const double* __restrict__ y, const double* __restrict__ z,
                                                   const double* __restrict__ m, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) pos4[i] = make_double4(x[i], y[i], z[i], m[i]);
}

__global__ void __launch_bounds__(256) unpack_kernel(const double4* __restrict__ pos4, double* x, double* y,
                                                     double* z, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) {
        const double4 p = pos4[i];
        x[i] = p.x; y[i] = p.y; z[i] = p.z;
    }
}

cudaError_t launch_pack(double4* pos4, const double* x, const double* y, const double* z, const double* m,
                        long long n, cudaStream_t st) {
    pack_kernel<<<grid_for(n, 256), 256, 0, st>>>(pos4, x, y, z, m, n);
    return cudaGetLastError();
}

cudaError_t launch_unpack(const double4* pos4, double* x, double* y, double* z, long long n, cudaStream_t st) {
    unpack_kernel<<<grid_for(n, 256), 256, 0, st>>>(pos4, x, y, z, n);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Potential energy  U = sum_{i<j} ((-G m_i) m_j) / sqrt(r2)   (physics.py:158)
// ---------------------------------------------------------------------------
// Reference order: one running sum over pairs in lexicographic (i, j>i) order.
// The terms of a row chunk are produced by the whole CTA, then added by thread 0
// in order, so the result is bit-identical to the Python loop.
__global__ void __launch_bounds__(256) potential_faithful_kernel(const double4* __restrict__ pos4, int n,
                                                                 double eps2, double G, double* out) {
    __shared__ double terms[256];
    const double U = potential_ordered_block(pos4, n, eps2, G, terms);
    if (threadIdx.x == 0) *out = U;
}

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;   // valid in warp 0
}

// Tree-order potential for large n: thread i sums its row (j > i), block + grid reduction.
__global__ void __launch_bounds__(256) potential_tree_kernel(const double4* __restrict__ pos4, long long n,
                                                             double eps2, double* partial) {
    __shared__ double4 tile[256];
    __shared__ double sh[8];
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long li = min(i, n - 1);
    const double4 me = pos4[li];
    double u = 0.0;
    const long long jstart = (blockIdx.x * (long long)blockDim.x / 256) * 256;
    for (long long j0 = jstart; j0 < n; j0 += 256) {
        const long long jl = j0 + threadIdx.x;
        if (jl < n) tile[threadIdx.x] = pos4[jl];
        __syncthreads();
        const int cnt = (int)min(256LL, n - j0);
        for (int j = 0; j < cnt; ++j) {
            if (j0 + j > i && i < n) {
                const double4 q = tile[j];
                const double dx = q.x - me.x, dy = q.y - me.y, dz = q.z - me.z;
                const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
                u += q.w * (1.0 / sqrt(r2));
            }
        }
        __syncthreads();
    }
    u *= me.w;
    const double tot = block_sum_256(u, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// Large-n potential: each unordered pair once (j > i), seed + polynomial reciprocal square root
// (inv_r = y0 (1 + e/2 + 3/8 e^2), truncation 5/16 e^3 < 1e-17), TI targets per thread, 256-body source tiles in
// shared memory.  A CTA is (block of 128*TI targets, slab of the tiles at or after it); tiles that overlap the
// target block mask j <= i, the others run unmasked.  12 FP64 instructions per pair; U = -G sum_i m_i u_i.
template <int TI>
__global__ void __launch_bounds__(128) potential_fast_kernel(const double4* __restrict__ pos4, long long n,
                                                             double eps2, int slabs, double* partial) {
    __shared__ double4 tile[256];
    __shared__ double sh[4];
    const int b = blockIdx.x / slabs, sl = blockIdx.x - b * slabs;
    const long long i_lo = (long long)b * 128 * TI;
    const long long i_hi = i_lo + 128 * TI;
    const int n_tiles = (int)((n + 255) / 256);
    const int t_first = (int)(i_lo / 256);
    const int per = (n_tiles - t_first + slabs - 1) / slabs;
    const int t0 = t_first + sl * per;
    const int t1 = min(n_tiles, t0 + per);
    double xi[TI], yi[TI], zi[TI], u[TI];
    long long idx[TI];
#pragma unroll
    for (int k = 0; k < TI; ++k) {
        idx[k] = i_lo + (long long)k * 128 + threadIdx.x;
        const double4 me = pos4[min(idx[k], n - 1)];
        xi[k] = me.x; yi[k] = me.y; zi[k] = me.z;
        u[k] = 0.0;
    }
    for (int t = t0; t < t1; ++t) {
        const long long j0 = (long long)t * 256;
        const int cnt = (int)min(256LL, n - j0);
        for (int e = threadIdx.x; e < cnt; e += 128) tile[e] = pos4[j0 + e];
        __syncthreads();
        const bool masked = j0 < i_hi;
#pragma unroll 2
        for (int j = 0; j < cnt; ++j) {
            const double4 q = tile[j];
#pragma unroll
            for (int k = 0; k < TI; ++k) {
                const double dx = q.x - xi[k], dy = q.y - yi[k], dz = q.z - zi[k];
                const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, eps2)));
                const double y0 = rsqrt_seed(r2);
                const double e = fma(-r2, y0 * y0, 1.0);
                double inv = fma(y0, e * fma(0.375, e, 0.5), y0);
                if (masked) inv = (j0 + j > idx[k]) ? inv : 0.0;
                u[k] = fma(q.w, inv, u[k]);
            }
        }
        __syncthreads();
    }
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < TI; ++k) tot += idx[k] < n ? u[k] * pos4[idx[k]].w : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) partial[blockIdx.x] = (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

__global__ void __launch_bounds__(256) final_sum_kernel(const double* __restrict__ partial, int count, int ncomp,
                                                        double scale0, double* out) {
    __shared__ double sh[8];
    for (int c = 0; c < ncomp; ++c) {
        double v = 0.0;
        for (int k = threadIdx.x; k < count; k += blockDim.x) v += partial[(long long)c * count + k];
        const double tot = block_sum_256(v, sh);
        if (threadIdx.x == 0) out[c] = (c == 0 ? scale0 : 1.0) * tot;
    }
}

cudaError_t launch_potential(const DeviceState& s, const StepParams& p, bool faithful_order, double* d_out,
                             cudaStream_t st, int* launches) {
    if (faithful_order) {
        potential_faithful_kernel<<<1, 256, 0, st>>>(s.pos4, (int)s.n, p.eps2, p.G, d_out);
        if (launches) ++*launches;
        return cudaGetLastError();
    }
    int grid = grid_for(s.n, 256);
    if (s.n > 4096) {
        // reduce_buf holds 4 * (ceil(n/256) + 1) partials: 8 slabs per block of 512 targets always fit
        constexpr int kTI = 4, kSlabs = 8;
        grid = grid_for(s.n, 128 * kTI) * kSlabs;
        potential_fast_kernel<kTI><<<grid, 128, 0, st>>>(s.pos4, s.n, p.eps2, kSlabs, s.reduce_buf);
    } else {
        potential_tree_kernel<<<grid, 256, 0, st>>>(s.pos4, s.n, p.eps2, s.reduce_buf);
    }
    final_sum_kernel<<<1, 256, 0, st>>>(s.reduce_buf, grid, 1, -p.G, d_out);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

// Potential energy of ONE body against all others, unsoftened, in the reference's loop order -- the potential term
// of Object.lagrangian (physics.py:275-279):  pe = sum_{j != i, ascending} ((-G m_i) m_j) / ||r_i - r_j||  with
// np.linalg.norm = sqrt(x.dot(x)).  The CTA forms 256 terms at a time, thread 0 adds them in order: the same bits as
// the Python loop (a coincident body gives -inf, as there).
__global__ void __launch_bounds__(256) body_potential_kernel(const double4* __restrict__ pos4, long long n, long long i,
                                                             double G, double* out) {
    __shared__ double terms[256];
    const double4 pi = pos4[i];
    const double gmi = __dmul_rn(-G, pi.w);
    double pe = 0.0;
    for (long long j0 = 0; j0 < n; j0 += 256) {
        const long long j = j0 + threadIdx.x;
        if (j < n && j != i) {
            const double4 pj = pos4[j];
            const double dx = __dsub_rn(pi.x, pj.x), dy = __dsub_rn(pi.y, pj.y), dz = __dsub_rn(pi.z, pj.z);
            terms[threadIdx.x] = __ddiv_rn(__dmul_rn(gmi, pj.w), __dsqrt_rn(dot3_numpy(dx, dy, dz)));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int cnt = (int)min(256LL, n - j0);
            for (int k = 0; k < cnt; ++k)
                if (j0 + k != i) pe = __dadd_rn(pe, terms[k]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = pe;
}

cudaError_t launch_body_potential(const DeviceState& s, long long i, double G, double* d_out, cudaStream_t st) {
    body_potential_kernel<<<1, 256, 0, st>>>(s.pos4, s.n, i, G, d_out);
    return cudaGetLastError();
}

// Sum of the ranks' partial accelerations for this rank's slab, read from peer memory (see orb_peer_reduce).
// ld.global.cv: the addresses were read in the previous step too and nothing may be served from a stale line; all
// peers' loads are issued before the first addition (memory-level parallelism over NVLink).
struct PeerAcc { const double* p[16]; };
__global__ void __launch_bounds__(256) peer_reduce_kernel(const PeerAcc peers, int world, double* acc, long long n,
                                                          long long lo, long long hi) {
    const long long per = hi - lo;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= 3 * per) return;
    const long long c = t / per;
    const long long idx = c * n + lo + (t - c * per);
    double v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r)             // compile-time index: the pointers stay in the parameter bank
        v[r] = r < world ? __ldcv(peers.p[r] + idx) : 0.0;
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 16; ++r)
        if (r < world) s += v[r];            // rank order
    acc[idx] = s;
}

cudaError_t launch_peer_reduce(const double* const* peer_acc, int world, double* acc, long long n, long long lo,
                               long long hi, cudaStream_t st) {
    PeerAcc pa{};
    for (int r = 0; r < world && r < 16; ++r) pa.p[r] = peer_acc[r];
    const long long tot = 3 * (hi - lo);
    peer_reduce_kernel<<<grid_for(tot, 256), 256, 0, st>>>(pa, world, acc, n, lo, hi);
    return cudaGetLastError();
}

// K = sum 1/2 m v.v ; L = sum r x (m v)     (engine.py:104-121, fp64)
__global__ void __launch_bounds__(256) energy_angmom_kernel(const double4* __restrict__ pos4,
                                                            const double* __restrict__ vel, long long n,
                                                            long long lo, long long hi, double* partial,
                                                            int nblocks) {
    __shared__ double sh[8];
    const long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double k = 0, lx = 0, ly = 0, lz = 0;
    if (i < hi) {
        const double4 p = pos4[i];
        const double vx = vel[i], vy = vel[i + n], vz = vel[i + 2 * n];
        k = 0.5 * p.w * (vx * vx + vy * vy + vz * vz);
        const double px = p.w * vx, py = p.w * vy, pz = p.w * vz;
        lx = p.y * pz - p.z * py;
        ly = p.z * px - p.x * pz;
        lz = p.x * py - p.y * px;
    }
    double t;
    t = block_sum_256(k, sh);  if (threadIdx.x == 0) partial[0 * nblocks + blockIdx.x] = t;
    t = block_sum_256(lx, sh); if (threadIdx.x == 0) partial[1 * nblocks + blockIdx.x] = t;
    t = block_sum_256(ly, sh); if (threadIdx.x == 0) partial[2 * nblocks + blockIdx.x] = t;
    t = block_sum_256(lz, sh); if (threadIdx.x == 0) partial[3 * nblocks + blockIdx.x] = t;
}

cudaError_t launch_energy_angmom(const DeviceState& s, double* d_out4, cudaStream_t st, int* launches) {
    const long long nt = s.tgt_hi - s.tgt_lo;
    const int grid = grid_for(nt, 256);
    energy_angmom_kernel<<<grid, 256, 0, st>>>(s.pos4, s.vel, s.n, s.tgt_lo, s.tgt_hi, s.reduce_buf, grid);
    final_sum_kernel<<<1, 256, 0, st>>>(s.reduce_buf, grid, 4, 1.0, d_out4);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

}  // namespace orb
