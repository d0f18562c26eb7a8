// common.cuh -- shared device helpers for liborbital_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef __CUDA_ARCH__
#define ORB_HOST_ONLY 1
#endif

namespace orb {

// ---------------------------------------------------------------------------
// Device-side control block shared by every kernel of one engine handle.
// Lives in device memory; a copy is read back into pinned host memory at the
// end of orb_step.
// ---------------------------------------------------------------------------
struct Ctl {
    long long steps_done;     // completed steps in the current orb_step call
    long long hist_count;     // snapshots appended to the history ring so far
    int halted;               // set once a step saw an overlap: later kernels no-op
    int overlap_count;        // overlapping pairs (i<j) seen in the halting step
    int overlap_overflow;     // pairs dropped because the list was full
    int u_valid;              // u_stash holds U of the last force build (stashed before contacts moved bodies)
    double u_stash;
    long long contacts_total; // touching pairs resolved on the device since the last orb_step began
    unsigned int rows_done;   // CTAs of faithful_rows_kernel<true> that finished their tail (last one advances)
    int pairs_cap;            // capacity of the pair list (set once at allocation, never reset)
    int full_sweeps;          // contact sweeps that fell back to the list-free lexicographic scan (list overflow)
    int pad_;
};

// Pair-list capacity per step: 2^20 pairs (16 MiB) by default; ORBITAL_B200_OVERLAP_CAP overrides (tests force the
// overflow path with a tiny list).  When a step flags more pairs than fit, the device-side sweep switches to an
// exact list-free scan (contacts.cuh) and the host-resolved mode sweeps all pairs (core/engine.py).
constexpr int kOverlapCapDefault = 1 << 20;

// ---------------------------------------------------------------------------
// Bit-faithful arithmetic.  Every operation is a single
// explicitly rounded intrinsic so ptxas cannot contract mul+add into FMA.
// ---------------------------------------------------------------------------

// core/physics.py:145-146: `float(rij @ rij)` -- OpenBLAS ddot tail: fma(z,z,fma(y,y,x*x))
__device__ __forceinline__ double dot3_numpy(double dx, double dy, double dz) {
    return __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
}

// One pair term of core/physics.py:145-151. Gm = G*m_j (rounded once, :151 left-to-right).
__device__ __forceinline__ void pair_faithful(double dx, double dy, double dz, double eps2, double Gm,
                                              double& ax, double& ay, double& az) {
    const double r2 = __dadd_rn(dot3_numpy(dx, dy, dz), eps2);     // :146
    const double inv_r = __ddiv_rn(1.0, __dsqrt_rn(r2));           // :147
    const double inv_r3 = __ddiv_rn(inv_r, r2);                    // :148
    const double s = __dmul_rn(Gm, inv_r3);                        // :151
    ax = __dadd_rn(ax, __dmul_rn(s, dx));                          // :154  (mul, then add: no FMA)
    ay = __dadd_rn(ay, __dmul_rn(s, dy));
    az = __dadd_rn(az, __dmul_rn(s, dz));
}

// The same pair, but returning the three products s*d without accumulating: lets many lanes evaluate
// terms in parallel while one lane adds them in the reference's ascending-j order.
__device__ __forceinline__ void pair_term_faithful(double dx, double dy, double dz, double eps2, double Gm,
                                                   double& tx, double& ty, double& tz) {
    const double r2 = __dadd_rn(dot3_numpy(dx, dy, dz), eps2);     // :146
    const double inv_r = __ddiv_rn(1.0, __dsqrt_rn(r2));           // :147
    const double inv_r3 = __ddiv_rn(inv_r, r2);                    // :148
    const double s = __dmul_rn(Gm, inv_r3);                        // :151
    tx = __dmul_rn(s, dx);                                         // :151 (a_i = s * rij)
    ty = __dmul_rn(s, dy);
    tz = __dmul_rn(s, dz);
}

// core/engine.py:70,82: `obj.velocity += 0.5 * dt * acc`  (h = 0.5*dt formed on the host)
__device__ __forceinline__ double kick_faithful(double v, double h, double a, bool f32) {
    const double r = __dadd_rn(v, __dmul_rn(h, a));
    return f32 ? (double)__double2float_rn(r) : r;                 // f32 array: cast on store
}

// core/engine.py:74: `obj.position() + obj.velocity * dt`; f32 velocity -> product in float32
__device__ __forceinline__ double drift_faithful(double r, double v, double dt, float dt32, bool f32) {
    if (f32) {
        const float p = __fmul_rn(__double2float_rn(v), dt32);
        return __dadd_rn(r, (double)p);
    }
    return __dadd_rn(r, __dmul_rn(v, dt));
}

// core/physics.py:517-518: `np.linalg.norm(ri - rj) <= Ri + Rj`
__device__ __forceinline__ bool overlap_exact(double dx, double dy, double dz, double Ri, double Rj) {
    const double t = dot3_numpy(dx, dy, dz);
    const double rs = __dadd_rn(Ri, Rj);
    // cheap conservative prefilter, then the exact correctly-rounded sqrt compare
    if (t > rs * rs * 1.0000000000001) return false;
    return __dsqrt_rn(t) <= rs;
}

__device__ __forceinline__ void record_overlap(Ctl* ctl, long long* pairs, long long i, long long j) {
    const int slot = atomicAdd(&ctl->overlap_count, 1);
    if (slot < ctl->pairs_cap) {
        pairs[2 * slot] = i;
        pairs[2 * slot + 1] = j;
    } else {
        atomicAdd(&ctl->overlap_overflow, 1);
    }
}

// ---------------------------------------------------------------------------
// mbarrier + 1-D bulk TMA (cp.async.bulk, SASS: UBLKCP) wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared bulk copy, completion signalled on `bar` (bytes: multiple of 16, 16-B aligned)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Approximate reciprocal square root seed of a double (SASS: MUFU.RSQ64H on the
// high word; low word of the result is zero; relative error ~2^-20).
__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

}  // namespace orb
