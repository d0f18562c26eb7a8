// contacts.cuh -- device-side contact resolution with the reference's sequential semantics.
//
// Reference: ObjectCollection.handle_collisions (core/physics.py:510-535, merge_on_capture=False, the only
// branch the engine reaches) and collide_spheres (core/physics.py:391-422).  The sweep visits pairs i<j in
// lexicographic order and mutates bodies in place, so a contact can create or remove later contacts.
//
// The force pass has already flagged every pair that overlaps at the drifted positions.  resolve_contacts_block
// (one CTA) then replays the sweep exactly: repeatedly pick the smallest not-yet-visited candidate pair
// (parallel min), let ONE thread re-test it against the current positions and apply collide_spheres with the
// reference's rounding sequence (every operation an explicitly rounded intrinsic; float32-velocity bodies round
// as NumPy does), and, if a body was pushed out, have the whole CTA scan its later pairs for new overlaps and
// queue them.  A pair that was never flagged and whose bodies never moved cannot be touching, so the visited
// set equals the reference's -- the same argument as core/physics.py::ObjectCollection.resolve_contacts (host).
// If the pair list ever overflows (Ctl::pairs_cap), the sweep continues list-free from the current pair
// (full_sweep_block): slower, but still exactly the reference's loop.
#pragma once
#include "common.cuh"

namespace orb {

// One body as collide_spheres sees it.
struct ContactBody {
    double x, y, z, vx, vy, vz, m, R;
    bool f32;            // velocity is a float32 array in the reference (physics.py:184)
};

// collide_spheres(obj1 = A, obj2 = B, restitution) -- core/physics.py:391-422 with the reference's rounding
// sequence.  Returns true when both bodies were pushed out of overlap (their positions changed).
__device__ inline bool collide_bodies(ContactBody& A, ContactBody& B, double restitution) {
    double nx = __dsub_rn(A.x, B.x), ny = __dsub_rn(A.y, B.y), nz = __dsub_rn(A.z, B.z);          // :394
    const double dist = __dsqrt_rn(dot3_numpy(nx, ny, nz));                                        // :395
    if (dist == 0.0) return false;                                                                 // :396
    nx = __ddiv_rn(nx, dist); ny = __ddiv_rn(ny, dist); nz = __ddiv_rn(nz, dist);                  // :398
    const double m1 = A.m, m2 = B.m;
    double wx, wy, wz;                                       // :401 velocity difference keeps the common dtype
    if (A.f32 && B.f32) {
        wx = (double)__fsub_rn((float)A.vx, (float)B.vx);
        wy = (double)__fsub_rn((float)A.vy, (float)B.vy);
        wz = (double)__fsub_rn((float)A.vz, (float)B.vz);
    } else {
        wx = __dsub_rn(A.vx, B.vx); wy = __dsub_rn(A.vy, B.vy); wz = __dsub_rn(A.vz, B.vz);
    }
    const double v_rel = __fma_rn(wz, nz, __fma_rn(wy, ny, __dmul_rn(wx, nx)));   // np.dot -> ddot
    if (v_rel >= 0.0) return false;                                                // :402 separating
    const double m1_inv = __ddiv_rn(1.0, m1), m2_inv = __ddiv_rn(1.0, m2);         // :408-409
    double e = restitution;                                                        // :410 np.clip
    e = e < 0.0 ? 0.0 : (e > 1.0 ? 1.0 : e);
    const double inv_sum = __dadd_rn(m1_inv, m2_inv);
    const double j = __ddiv_rn(__dmul_rn(-__dadd_rn(1.0, e), v_rel), inv_sum);     // :412
    const double ix = __dmul_rn(j, nx), iy = __dmul_rn(j, ny), iz = __dmul_rn(j, nz);   // :413
    double t;                                                                      // :414-415 in place, dtype kept
    t = __dadd_rn(A.vx, __ddiv_rn(ix, m1)); A.vx = A.f32 ? (double)__double2float_rn(t) : t;
    t = __dadd_rn(A.vy, __ddiv_rn(iy, m1)); A.vy = A.f32 ? (double)__double2float_rn(t) : t;
    t = __dadd_rn(A.vz, __ddiv_rn(iz, m1)); A.vz = A.f32 ? (double)__double2float_rn(t) : t;
    t = __dsub_rn(B.vx, __ddiv_rn(ix, m2)); B.vx = B.f32 ? (double)__double2float_rn(t) : t;
    t = __dsub_rn(B.vy, __ddiv_rn(iy, m2)); B.vy = B.f32 ? (double)__double2float_rn(t) : t;
    t = __dsub_rn(B.vz, __ddiv_rn(iz, m2)); B.vz = B.f32 ? (double)__double2float_rn(t) : t;
    const double overlap = __dsub_rn(__dadd_rn(A.R, B.R), dist);                   // :418
    if (!(overlap > 0.0)) return false;
    const double corr = __ddiv_rn(overlap, inv_sum);                               // :420
    const double c1 = __ddiv_rn(corr, m1), c2 = __ddiv_rn(corr, m2);
    A.x = __dadd_rn(A.x, __dmul_rn(nx, c1)); A.y = __dadd_rn(A.y, __dmul_rn(ny, c1));   // :421
    A.z = __dadd_rn(A.z, __dmul_rn(nz, c1));
    B.x = __dsub_rn(B.x, __dmul_rn(nx, c2)); B.y = __dsub_rn(B.y, __dmul_rn(ny, c2));   // :422
    B.z = __dsub_rn(B.z, __dmul_rn(nz, c2));
    return true;
}

// collide_spheres(obj a, obj b) on the resident state. moved: both bodies were pushed out of overlap.
__device__ inline void collide_pair_dev(long long a, long long b, double4* pos4, double* vel, long long n,
                                        const double* radius, const uint8_t* vf32, double restitution, bool* moved) {
    const double4 pa = pos4[a], pb = pos4[b];
    ContactBody A = {pa.x, pa.y, pa.z, vel[a], vel[a + n], vel[a + 2 * n], pa.w, radius[a], vf32[a] != 0};
    ContactBody B = {pb.x, pb.y, pb.z, vel[b], vel[b + n], vel[b + 2 * n], pb.w, radius[b], vf32[b] != 0};
    *moved = collide_bodies(A, B, restitution);
    vel[a] = A.vx; vel[a + n] = A.vy; vel[a + 2 * n] = A.vz;
    vel[b] = B.vx; vel[b + n] = B.vy; vel[b + 2 * n] = B.vz;
    if (*moved) {
        pos4[a] = make_double4(A.x, A.y, A.z, pa.w);
        pos4[b] = make_double4(B.x, B.y, B.z, pb.w);
    }
}

// List-free continuation of the sweep (the pair list overflowed, so it can no longer be trusted to hold every
// candidate): the reference's own loop (physics.py:513-518) from the pair after `cursor` on -- for each row i the CTA
// finds the smallest j >= jstart that touches i AT THE CURRENT POSITIONS, one thread collides it, and the scan
// of the row resumes behind it.  Pairs that do not touch are no-ops in the reference, so visiting only the touching
// ones in lexicographic order is the same sweep.  O(n^2 / blockDim) per call: a fallback, not a fast path.
__device__ inline int full_sweep_block(double4* pos4, double* vel, long long n, const double* radius,
                                       const uint8_t* vf32, double restitution, unsigned long long cursor,
                                       unsigned long long* s_best) {
    const unsigned long long kNone = ~0ull;
    const int tid = threadIdx.x, nth = blockDim.x;
    int hits = 0;
    long long i = cursor ? (long long)((cursor - 1) / n) : 0;
    long long jstart = cursor ? (long long)((cursor - 1) % n) + 1 : 1;
    for (; i < n - 1; ++i, jstart = i + 1) {
        for (;;) {
            if (tid == 0) *s_best = kNone;
            __syncthreads();
            const double4 pi = pos4[i];
            const double Ri = radius[i];
            for (long long j = jstart + tid; j < n; j += nth) {
                const double4 pj = pos4[j];
                if (overlap_exact(__dsub_rn(pi.x, pj.x), __dsub_rn(pi.y, pj.y), __dsub_rn(pi.z, pj.z), Ri, radius[j])) {
                    atomicMin(s_best, (unsigned long long)j);       // ascending per thread: the first hit is its smallest
                    break;
                }
            }
            __syncthreads();
            const unsigned long long j = *s_best;
            __syncthreads();                                        // everyone has read s_best before it is reset
            if (j == kNone) break;
            if (tid == 0) {
                bool moved;
                collide_pair_dev(i, (long long)j, pos4, vel, n, radius, vf32, restitution, &moved);
                __threadfence_block();
            }
            ++hits;
            jstart = (long long)j + 1;
            __syncthreads();
        }
    }
    return hits;
}

// Replay of the lexicographic sweep over the flagged pairs, executed by every thread of one CTA.
// `scratch` : 4 x 8 bytes of shared memory.  Returns (in every thread) the number of touching pairs processed.
__device__ inline int resolve_contacts_block(double4* pos4, double* vel, long long n, const double* radius,
                                             const uint8_t* vf32, double restitution, Ctl* ctl, long long* pairs,
                                             unsigned long long* scratch) {
    const unsigned long long kNone = ~0ull;
    unsigned long long* s_best = scratch;          // smallest candidate key above the cursor
    unsigned long long* s_cursor = scratch + 1;    // key of the last visited pair (+1), 0 = none yet
    unsigned long long* s_moved = scratch + 2;     // the two bodies of the last contact if they were pushed out
    unsigned long long* s_hits = scratch + 3;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int cap = ctl->pairs_cap;
    if (tid == 0) { *s_cursor = 0; *s_hits = 0; }
    __syncthreads();
    for (;;) {
        if (*(volatile int*)&ctl->overlap_overflow > 0) {
            // a pair was dropped (by the force pass, or by the re-queue below): finish without the list
            const unsigned long long cursor = *s_cursor;
            __syncthreads();
            const int more = full_sweep_block(pos4, vel, n, radius, vf32, restitution, cursor, s_best);
            if (tid == 0) { *s_hits += more; ctl->full_sweeps += 1; }
            __syncthreads();
            break;
        }
        if (tid == 0) { *s_best = kNone; *s_moved = kNone; }
        __syncthreads();
        const int count = min(*(volatile int*)&ctl->overlap_count, cap);
        const unsigned long long cursor = *s_cursor;
        unsigned long long best = kNone;
        for (int c = tid; c < count; c += nth) {
            const unsigned long long key = (unsigned long long)pairs[2 * c] * n + pairs[2 * c + 1] + 1;
            if (key > cursor && key < best) best = key;
        }
        if (best != kNone) atomicMin(s_best, best);
        __syncthreads();
        const unsigned long long key = *s_best;
        if (key == kNone) break;
        const long long i = (long long)((key - 1) / n), j = (long long)((key - 1) % n);
        if (tid == 0) {
            *s_cursor = key;
            const double4 pi = pos4[i], pj = pos4[j];
            // physics.py:517-518 against the CURRENT positions (an earlier contact may have moved either body)
            if (overlap_exact(__dsub_rn(pi.x, pj.x), __dsub_rn(pi.y, pj.y), __dsub_rn(pi.z, pj.z), radius[i], radius[j])) {
                bool moved;
                collide_pair_dev(i, j, pos4, vel, n, radius, vf32, restitution, &moved);
                *s_hits += 1;
                if (moved) *s_moved = key;
            }
            __threadfence_block();
        }
        __syncthreads();
        if (*s_moved != kNone) {
            // both bodies were displaced: every LATER pair of either one may now overlap -> queue the touching ones
            for (int which = 0; which < 2; ++which) {
                const long long k = which ? j : i;
                const double4 pk = pos4[k];
                const double Rk = radius[k];
                for (long long c = tid; c < n; c += nth) {
                    if (c == k) continue;
                    const long long a = min(k, c), b = max(k, c);
                    const unsigned long long ck = (unsigned long long)a * n + b + 1;
                    if (ck <= key) continue;
                    const double4 pc = pos4[c];
                    // |ra - rb| is symmetric in (a, b): squares of exact negations
                    if (overlap_exact(__dsub_rn(pk.x, pc.x), __dsub_rn(pk.y, pc.y), __dsub_rn(pk.z, pc.z), Rk, radius[c]))
                        record_overlap(ctl, pairs, a, b);
                }
            }
            __threadfence_block();
        }
        __syncthreads();
    }
    const int hits = (int)*s_hits;
    __syncthreads();
    return hits;
}

// U = sum_{i<j} ((-G mi) mj) / sqrt(r2) in lexicographic pair order (physics.py:158), by one CTA of <= 256 threads.
// `terms`: blockDim doubles of shared memory. Result valid in thread 0.
__device__ inline double potential_ordered_block(const double4* pos4, int n, double eps2, double G, double* terms) {
    double U = 0.0;
    const int nth = blockDim.x;
    for (int i = 0; i < n - 1; ++i) {
        const double4 pi = pos4[i];
        const double gmi = __dmul_rn(-G, pi.w);
        for (int j0 = i + 1; j0 < n; j0 += nth) {
            const int j = j0 + threadIdx.x;
            if (j < n) {
                const double4 pj = pos4[j];
                const double dx = __dsub_rn(pj.x, pi.x), dy = __dsub_rn(pj.y, pi.y), dz = __dsub_rn(pj.z, pi.z);
                const double r2 = __dadd_rn(dot3_numpy(dx, dy, dz), eps2);
                const double inv_r = __ddiv_rn(1.0, __dsqrt_rn(r2));
                terms[threadIdx.x] = __dmul_rn(__dmul_rn(gmi, pj.w), inv_r);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const int cnt = min(nth, n - j0);
                for (int k = 0; k < cnt; ++k) U = __dadd_rn(U, terms[k]);
            }
            __syncthreads();
        }
    }
    return U;
}

}  // namespace orb
