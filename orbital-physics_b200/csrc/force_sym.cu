// force_sym.cu -- pair-symmetric fast force kernel (sm_100a).
//
// Same result as force_fast_kernel (reference core/physics.py:125-159) but each unordered pair
// (i<j) is evaluated ONCE and applied to both bodies, which is how the reference's half-matrix
// loop works (physics.py:136-155: a_i += G m_j s d, a_j -= G m_i s d).  FP64-pipe cost per pair:
//   3 DADD (d) + 3 DFMA (r^2) + 6 (y0^3 (1 + 3/2 e + 15/8 e^2) from the MUFU.RSQ64H seed)
//   + [1 DMUL + 3 DFMA] for body i + [1 DMUL + 3 DFMA] for body j = 20 per pair
//   = 10 FP64 instructions per ordered interaction (the one-sided kernel needs 16).
//
// Decomposition.  Bodies are cut into I-blocks of 128*TI bodies and tiles of 256 bodies.  A work
// item (CTA) is (I-block, chunk of consecutive tiles at or after the I-block): the thread keeps TI
// bodies of the I-block (position, mass, accumulator) in registers; tiles stream through the same
// 4-stage bulk-TMA ring as the one-sided kernel.
//   * tiles that overlap the I-block ("diagonal") are processed one-sided with the self-pair masked;
//   * tiles after it are processed symmetrically with a warp-level systolic rotation: each lane
//     meets one tile body j per step (read from the shared-memory tile) and carries j's accumulator;
//     after every TI pairs the accumulator rotates to the neighbouring lane (6 SHFL) and the lane moves
//     on to the next body, so after 32 steps every lane's TI bodies have met all 32
//     bodies of the round and every j accumulator is back in its home lane.  The four warps'
//     j-accumulators are summed in fixed order through shared memory and written to the partial
//     plane P_j[I-block][j].
// A reduction kernel then forms  a[x] = G (sum_chunks P_i[chunk][x] + sum_{I before x} P_j[I][x])
// in fixed order, so the result is run-to-run deterministic.  P_j is processed in panels of
// I-blocks to bound its footprint.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "force_common.cuh"
#include "force_sym.h"

namespace orb {

struct SymArgs {
    const double4* pos4;
    const double* radius;
    const SymItem* items;
    double* Pi;                // [n_chunks][3][n]
    double* Pj;                // [panel_blocks][3][n]
    long long n;
    int n_tiles;
    int tile;                  // bodies per source tile: 64, 128 or 256 (<= kTile, the stage size of the ring)
    double eps2;
    double rmax1, rmax2;
    long long rmax1_idx;
    Ctl* ctl;
    long long* pairs;
};

__device__ __forceinline__ double rot1(double v, int src_lane) {
    return __shfl_sync(0xffffffffu, v, src_lane);
}

#ifndef SYM_UNROLL
#define SYM_UNROLL 4
#endif
#ifndef SYM_MINB_HI
#define SYM_MINB_HI 2   // resident CTAs/SM requested for TI >= 5
#endif
#ifndef SYM_SMEMROT
#define SYM_SMEMROT 1     // 1: rotate the accumulators through shared memory (2 STS + 2 LDS) instead of 6 SHFL
#endif
#ifndef SYM_SPLITB
#define SYM_SPLITB 0      // 1: two j-accumulator chains (even/odd k), merged before the rotation
#endif
#define ORB_STR2(x) #x
#define ORB_STR(x) ORB_STR2(x)

constexpr int kSymSlabBytes = kFastWarps * 3 * kTile * 8;
constexpr int kSymXchgBytes = kFastWarps * 2 * 32 * 24;      // per warp: 2 buffers x 32 lanes x {xy: 16 B, z: 8 B}
constexpr int kSymSmem = kStages * kTile * 32 + 128 + kSymSlabBytes + kSymXchgBytes;

// UNI: every body has the same mass and n is a whole number of I-blocks and tiles (no padded slots, no
// shadow threads): the two per-pair mass multiplies disappear (18 FP64 instructions per pair) and the
// reduction kernel scales by G*m.
// T256: the source tile is the full 256-body stage (what every system above ~6k bodies uses) -- a compile-time tile
// keeps the index arithmetic of the tile loop out of the FP64 schedule (measured 1.3 % on the N = 262,144 pass).
template <int TI, bool DETECT, bool UNI, bool T256 = false>
__global__ void __launch_bounds__(kFastThreads, (TI >= 5 ? SYM_MINB_HI : 3))
force_sym_kernel(const SymArgs g) {
    if (g.ctl->halted) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tiles = reinterpret_cast<double2*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * kTile * 32);
    uint64_t* empty = full + kStages;
    double* slab = reinterpret_cast<double*>(smem_raw + kStages * kTile * 32 + 128);   // [warp][3][kTile]
    unsigned char* xchg_raw = smem_raw + kStages * kTile * 32 + 128 + kSymSlabBytes;

    const SymItem it = g.items[blockIdx.x];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int next_lane = (lane + 1) & 31;
    constexpr long long B = (long long)kFastThreads * TI;
    const long long i_lo = (long long)it.I * B;
    const long long i_hi = min(i_lo + B, g.n);
    const int tile_n = T256 ? kTile : g.tile;
    const int diag_end = (int)min((long long)g.n_tiles, (i_hi + tile_n - 1) / tile_n);
    const int ntiles = it.t1 - it.t0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kFastWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t) {
        const int s = t % kStages;
        const long long j0 = (long long)(it.t0 + t) * tile_n;
        const int cnt = (int)min((long long)tile_n, g.n - j0);
        const uint32_t bytes = (uint32_t)cnt * 32u;
        mbar_expect_tx(&full[s], bytes);
        tma_load_1d(tiles + (size_t)s * kTile * 2, g.pos4 + j0, bytes, &full[s]);
    };
    if (tid == 0) {
        const int pre = min(kStages, ntiles);
        for (int t = 0; t < pre; ++t) issue(t);
    }

    double xi[TI], yi[TI], zi[TI], mi[TI], ax[TI], ay[TI], az[TI];
    long long idx[TI];
    int maxhi[TI], thr[TI];
    double Ri[TI];
#pragma unroll
    for (int k = 0; k < TI; ++k) {
        idx[k] = i_lo + (long long)k * kFastThreads + tid;
        const long long ld = min(idx[k], g.n - 1);          // tail threads shadow the last body with zero mass
        const double4 p = g.pos4[ld];
        xi[k] = p.x; yi[k] = p.y; zi[k] = p.z;
        mi[k] = idx[k] < g.n ? p.w : 0.0;
        ax[k] = ay[k] = az[k] = 0.0;
        maxhi[k] = 0;
        thr[k] = 0x7fffffff;
        Ri[k] = 0.0;
        if (DETECT) {
            Ri[k] = g.radius[ld];
            const double partner = (ld == g.rmax1_idx) ? g.rmax2 : g.rmax1;
            const double rs = Ri[k] + partner;
            const double bound = (1.0 / sqrt(fma(rs, rs, g.eps2))) * (1.0 - 1.52587890625e-05);
            thr[k] = __double2hiint(bound) - 1;
        }
    }

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages;
        mbar_wait(&full[s], (uint32_t)((t / kStages) & 1));
        const double2* tile = tiles + (size_t)s * kTile * 2;
        const int tile_index = it.t0 + t;
        const long long j0 = (long long)tile_index * tile_n;
        const int cnt = (int)min((long long)tile_n, g.n - j0);

        if (tile_index < diag_end) {
            // ---- diagonal tile: one-sided, self-pair masked (both directions are evaluated by their owners)
            tile_loop<TI, DETECT, true, UNI>(tile, cnt, j0, g.eps2, xi, yi, zi, idx, ax, ay, az, maxhi);
        } else {
            // ---- symmetric tile: systolic rotation, 32 bodies per round
            const int rounds = (cnt + 31) >> 5;
            for (int r = 0; r < rounds; ++r) {
                const int slot = (r << 5) + lane;
                const int r32 = r << 5;
                // body seen at step st sits in tile slot r32 + ((lane + st) & 31); padded slots (>= cnt)
                // read a real position with zero mass
                auto fetch = [&](int st, double& px, double& py, double& pz, double& pm) {
                    const int sl = r32 + ((lane + st) & 31);
                    const bool ok = UNI || sl < cnt;          // UNI: whole tiles only
                    const int src = ok ? sl : cnt - 1;
                    const double2 pa = tile[2 * src];
                    const double2 pb = tile[2 * src + 1];
                    px = pa.x; py = pa.y; pz = pb.x;
                    pm = ok ? pb.y : 0.0;
                };
                double jx, jy, jz, jm;
                fetch(0, jx, jy, jz, jm);
                double bx = 0.0, by = 0.0, bz = 0.0;
#if SYM_SPLITB
                double cx = 0.0, cy = 0.0, cz = 0.0;
#endif
_Pragma(ORB_STR(unroll SYM_UNROLL))
                for (int st = 0; st < 32; ++st) {
                    // next step's body comes straight from the shared-memory tile (2 LDS.128, in flight while
                    // this step computes); only the three accumulators have to travel by shuffle
                    double nx, ny, nz, nm;
                    fetch(st + 1, nx, ny, nz, nm);
#pragma unroll
                    for (int k = 0; k < TI; ++k) {
                        const double dx = jx - xi[k];
                        const double dy = jy - yi[k];
                        const double dz = jz - zi[k];
                        const double r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, g.eps2)));
                        int hi;
                        const double s0 = inv_r3_plain(r2, hi);
                        if (DETECT) maxhi[k] = max(maxhi[k], hi);
                        const double si = UNI ? s0 : s0 * jm;      // physics.py:151  a_i += (G m_j / r^3) d
                        const double sj = UNI ? s0 : s0 * mi[k];   // physics.py:152  a_j -= (G m_i / r^3) d
                        ax[k] = fma(si, dx, ax[k]);
                        ay[k] = fma(si, dy, ay[k]);
                        az[k] = fma(si, dz, az[k]);
#if SYM_SPLITB
                        if (k & 1) {
                            cx = fma(-sj, dx, cx); cy = fma(-sj, dy, cy); cz = fma(-sj, dz, cz);
                        } else
#endif
                        {
                            bx = fma(-sj, dx, bx);
                            by = fma(-sj, dy, by);
                            bz = fma(-sj, dz, bz);
                        }
                    }
                    jx = nx; jy = ny; jz = nz; jm = nm;
#if SYM_SPLITB
                    bx += cx; by += cy; bz += cz;
                    cx = cy = cz = 0.0;
#endif
#if SYM_SMEMROT
                    {
                        double2* exy = reinterpret_cast<double2*>(xchg_raw + (size_t)warp * 2 * 32 * 24) + (st & 1) * 32;
                        double* ez = reinterpret_cast<double*>(xchg_raw + (size_t)warp * 2 * 32 * 24 + 2 * 32 * 16) + (st & 1) * 32;
                        exy[lane] = make_double2(bx, by);
                        ez[lane] = bz;
                        __syncwarp();
                        const double2 t2 = exy[next_lane];
                        bx = t2.x; by = t2.y; bz = ez[next_lane];
                    }
#else
                    bx = rot1(bx, next_lane); by = rot1(by, next_lane); bz = rot1(bz, next_lane);
#endif
                }
                // 32 rotations: every accumulator is back in its home lane
                double* mine = slab + (size_t)warp * 3 * kTile;
                mine[slot] = bx;
                mine[kTile + slot] = by;
                mine[2 * kTile + slot] = bz;
            }
        }
        if (DETECT) {
#pragma unroll
            for (int k = 0; k < TI; ++k) {
                if (maxhi[k] >= thr[k] && idx[k] < g.n)
                    rescan_tile(tile, cnt, j0, idx[k], xi[k], yi[k], zi[k], Ri[k], g.radius, g.ctl, g.pairs);
                maxhi[k] = 0;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (tile_index >= diag_end) {
            __syncthreads();
            // fixed-order sum of the four warps' contributions to the tile bodies -> P_j[I][j]
            double* out = g.Pj + (long long)it.slot * 3 * g.n + j0;
            for (int e = tid; e < 3 * tile_n; e += kFastThreads) {
                const int c = e / tile_n;
                const int slot = e - c * tile_n;
                if (slot < cnt) {
                    double v = slab[c * kTile + slot];
#pragma unroll
                    for (int w = 1; w < kFastWarps; ++w) v += slab[(w * 3 + c) * kTile + slot];
                    out[(long long)c * g.n + slot] = v;
                }
            }
            __syncthreads();
        }
        if (tid == 0 && t >= 1 && (t - 1 + kStages) < ntiles) {
            const int tp = t - 1;
            mbar_wait(&empty[tp % kStages], (uint32_t)((tp / kStages) & 1));
            issue(tp + kStages);
        }
    }

    double* pi = g.Pi + (long long)it.chunk * 3 * g.n;
#pragma unroll
    for (int k = 0; k < TI; ++k) {
        if (idx[k] < g.n) {
            pi[idx[k]] = ax[k];
            pi[g.n + idx[k]] = ay[k];
            pi[2 * g.n + idx[k]] = az[k];
        }
    }
}

// Every element x is summed by kRedSplit threads (a quarter of the planes each, combined in fixed order): at
// N = 4,096 one thread per element meant 16 CTAs walking ~75 dependent-latency loads each -- 11 us, more than half
// of the force kernel it follows.
constexpr int kRedSplit = 4;
constexpr int kRedX = 256 / kRedSplit;           // elements per CTA

// share `sub` of  sum_{chunks of x's I-block} P_i[chunk][x] + sum_{owned I-blocks of this panel before x} P_j[slot][x]
__device__ __forceinline__ void sym_reduce_row(const double* __restrict__ Pi, const double* __restrict__ Pj, long long n,
                                               long long B, int tile, int chunk_tiles, int n_chunks, int rank, int world,
                                               int ka, int kb, int sub, long long x, double (&out)[3]) {
    const long long X = x / B;
    const long long T = x / tile;
    double s[3] = {0.0, 0.0, 0.0};
    // ownership (multi-GPU): every group of `world` consecutive I-blocks gives one block to each rank, in ascending
    // rank order in even groups and descending order in odd ones ("snake"), which balances the triangle: block I
    // costs ~(N - I B) B pairs, so plain cyclic ownership hands rank 0 a whole group-width more work per group than
    // the last rank (5.5 % of the pass at N = 262,144 on 8 ranks).
    const long long gx = X / world;
    const int posx = (int)(X - gx * world);
    if (((gx & 1) ? world - 1 - posx : posx) == rank) {
        const long long kx = gx;
        if (kx >= ka && kx < kb) {
            const int c_first = (int)(((X * B) / tile) / chunk_tiles);
            for (int c = c_first + sub; c < n_chunks; c += kRedSplit) {
#pragma unroll
                for (int k = 0; k < 3; ++k) s[k] += Pi[((long long)c * 3 + k) * n + x];
            }
        }
    }
    // owned I-blocks that treated x's tile symmetrically: I_k < nI, I_k = world k + pos(k): all full groups below nI,
    // plus the partial group if this rank's block in it comes before nI
    const long long nI = (T * tile) / B;
    const long long gI = nI / world;
    const int rem = (int)(nI - gI * world);
    long long kend = gI + ((((gI & 1) ? world - 1 - rank : rank) < rem) ? 1 : 0);
    if (kend > kb) kend = kb;
    for (long long k = ka + sub; k < kend; k += kRedSplit) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s[c] += Pj[((k - ka) * 3 + c) * n + x];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = s[k];
}

// The rest of the leapfrog step, riding along in the reduction (unsharded engines, one panel): a step is then
// kick_drift -> force_sym -> reduce, and inside a multi-step graph even the kick_drift of the NEXT step is done here,
// which leaves two launches per step.  What matters below ~16k bodies, where a launch costs about as much as the
// arithmetic (a dependent kernel node is ~3 us; the whole N = 4,096 force pass is ~12 us of FP64 work).
struct SymTail {
    int kick;              // second half-kick of this step (engine.py:81-82)
    int close;             // history append + step bookkeeping (engine.py:88-92, advance_kernel)
    int next;              // first half-kick + drift of the next step (engine.py:69-75), unless this step halts
    double4* pos4;
    double* vel;
    const uint8_t* vf32;
    double* hist;
    long long hist_cap;
    double h, dt;
    float dt32;
    Ctl* ctl_rw;
};

// a[x] (+)= G * ( sum_{chunks of x's I-block} P_i[chunk][x] + sum_{owned I-blocks of this panel before x} P_j[slot][x] )
// I-blocks are owned group-wise in snake order: I_k = world k + (k odd ? world-1-rank : rank); the panel holds k in [ka, kb).
__global__ void __launch_bounds__(256) reduce_sym_kernel(const double* __restrict__ Pi, const double* __restrict__ Pj,
                                                         double* acc, long long n, long long B, int tile,
                                                         int chunk_tiles, int n_chunks, int rank, int world, int ka,
                                                         int kb, double G, int accumulate, const Ctl* ctl,
                                                         const SymTail tail) {
    if (ctl->halted) return;
    __shared__ double part[kRedSplit - 1][3][kRedX];
    const int xl = threadIdx.x % kRedX, sub = threadIdx.x / kRedX;
    const long long x = blockIdx.x * (long long)kRedX + xl;
    double a3[3] = {0.0, 0.0, 0.0};
    if (x < n) sym_reduce_row(Pi, Pj, n, B, tile, chunk_tiles, n_chunks, rank, world, ka, kb, sub, x, a3);
    if (sub > 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) part[sub - 1][k][xl] = a3[k];
    }
    __syncthreads();
    if (sub == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double v = a3[k];
#pragma unroll
            for (int q = 0; q < kRedSplit - 1; ++q) v += part[q][k][xl];     // fixed order: deterministic
            a3[k] = G * v;
        }
    }
    const bool mine = sub == 0 && x < n;
    if (tail.close) {
        // every thread has read the ring cursor / overlap count before the last CTA to arrive moves them
        const long long hist_count = ctl->hist_count;
        const int overlaps = ctl->overlap_count;
        if (mine) {
            acc[x] = a3[0]; acc[x + n] = a3[1]; acc[x + 2 * n] = a3[2];
            const bool f32 = tail.vf32[x] != 0;
            double vx = kick_faithful(tail.vel[x], tail.h, a3[0], f32);
            double vy = kick_faithful(tail.vel[x + n], tail.h, a3[1], f32);
            double vz = kick_faithful(tail.vel[x + 2 * n], tail.h, a3[2], f32);
            double4 p = tail.pos4[x];
            if (tail.hist_cap > 0 && overlaps == 0) {        // a halting step appends on the host
                double* row = tail.hist + ((hist_count % tail.hist_cap) * n + x) * 3;
                row[0] = p.x; row[1] = p.y; row[2] = p.z;
            }
            if (tail.next && overlaps == 0) {                // the next step's first half (no contact in between)
                vx = kick_faithful(vx, tail.h, a3[0], f32);
                vy = kick_faithful(vy, tail.h, a3[1], f32);
                vz = kick_faithful(vz, tail.h, a3[2], f32);
                p.x = drift_faithful(p.x, vx, tail.dt, tail.dt32, f32);
                p.y = drift_faithful(p.y, vy, tail.dt, tail.dt32, f32);
                p.z = drift_faithful(p.z, vz, tail.dt, tail.dt32, f32);
                tail.pos4[x] = p;
            }
            tail.vel[x] = vx; tail.vel[x + n] = vy; tail.vel[x + 2 * n] = vz;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            Ctl* c = tail.ctl_rw;
            if (atomicAdd(&c->rows_done, 1u) == gridDim.x - 1) {
                c->rows_done = 0;
                c->steps_done += 1;
                if (overlaps > 0)
                    c->halted = 1;
                else if (tail.hist_cap > 0)
                    c->hist_count = hist_count + 1;
            }
        }
        return;
    }
    if (!mine) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double v = accumulate ? acc[x + k * n] + a3[k] : a3[k];
        acc[x + k * n] = v;
        a3[k] = v;
    }
    if (tail.kick) {                                         // device-resolved contacts follow: kick only
        const bool f32 = tail.vf32[x] != 0;
        tail.vel[x] = kick_faithful(tail.vel[x], tail.h, a3[0], f32);
        tail.vel[x + n] = kick_faithful(tail.vel[x + n], tail.h, a3[1], f32);
        tail.vel[x + 2 * n] = kick_faithful(tail.vel[x + 2 * n], tail.h, a3[2], f32);
        if (tail.hist_cap > 0 && ctl->overlap_count == 0) {  // as kick_hist_kernel: a step with contacts appends after the sweep
            const double4 p = tail.pos4[x];
            double* row = tail.hist + ((ctl->hist_count % tail.hist_cap) * n + x) * 3;
            row[0] = p.x; row[1] = p.y; row[2] = p.z;
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <int TI, bool DETECT>
static int sym_occupancy() {
    int nb = 0;
    auto kern = force_sym_kernel<TI, DETECT, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSymSmem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kFastThreads, kSymSmem) != cudaSuccess) nb = 0;
    return nb;
}

template <int TI, bool DETECT, bool UNI, bool T256>
static cudaError_t launch_sym_tt(const SymArgs& a, int grid, cudaStream_t st) {
    auto kern = force_sym_kernel<TI, DETECT, UNI, T256>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSymSmem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kFastThreads, kSymSmem, st>>>(a);
    return cudaGetLastError();
}

template <int TI, bool DETECT, bool UNI>
static cudaError_t launch_sym_t(const SymArgs& a, int grid, cudaStream_t st) {
    if (TI >= 6 && a.tile == kTile) return launch_sym_tt<TI, DETECT, UNI, (TI >= 6)>(a, grid, st);
    return launch_sym_tt<TI, DETECT, UNI, false>(a, grid, st);
}

const char* sym_kernel_name(int ti, bool detect, bool uniform) {
    static thread_local char buf[64];
    snprintf(buf, sizeof buf, "force_sym_kernel<%d,%s,%s>", ti, detect ? "true" : "false", uniform ? "true" : "false");
    return buf;
}

void free_sym(SymPlan& p) {
    for (auto& pan : p.panels) cudaFree(pan.d_items);
    p.panels.clear();
    cudaFree(p.Pi);
    cudaFree(p.Pj);
    p.Pi = p.Pj = nullptr;
    p.valid = false;
}

// Build the item lists (one per panel of I-blocks) and allocate the partial planes.
cudaError_t plan_sym(SymPlan& p, long long n, int sm_count, int rank, int world) {
    free_sym(p);
    const char* env_ti = getenv("ORBITAL_B200_SYM_TI");
    // I-blocks (128*TI bodies) must be tile aligned (a multiple or a divisor of the source tile), otherwise a
    // tile straddling two I-blocks would be treated one-sided by the lower block and its other bodies would miss
    // those pairs: TI in {1, 2, 4, 6, 8} with tiles of 64 / 128 / 256 bodies.  Geometry from the whole-step sweep on
    // B200 (profiles/r2_sweep_step_uniform.txt, us per step for the best / runner-up): N=1,024 13.4 (TI 1, tile 64),
    // 2,048 15.4 (TI 2, tile 64), 4,096 24.2 (TI 4, tile 128; TI 2 / tile 64: 35.4), 8,192 58.5 (TI 8, tile 256;
    // tile 64: 66.8), 16,384 197, 32,768 694, 65,536 2,609 (TI 8, tile 256).  Small systems want many small items
    // (one CTA per SM leaves a warp alone on its scheduler); from ~8k bodies up the larger tile wins because every
    // chunk adds a P_i plane to the reduction.
    int ti = 8;
    if (n <= 6144) ti = 4;
    if (n <= 3072) ti = 2;
    if (n <= 1536) ti = 1;
    if (env_ti) {
        const int v = atoi(env_ti);
        if (v == 1 || v == 2 || v == 4 || v == 6 || v == 8) ti = v;
    }
    p.ti = ti;
    p.n = n;
    p.rank = rank;
    p.world = world;
    p.B = (long long)kFastThreads * ti;
    p.nb_I = (int)((n + p.B - 1) / p.B);
    // Source-tile size: see the sweep above.
    p.tile = n <= 3072 ? 64 : (n <= 6144 ? 128 : kTile);
    {
        const char* env_tile = getenv("ORBITAL_B200_SYM_TILE");
        if (env_tile) {
            const int v = atoi(env_tile);
            if (v == 64 || v == 128 || v == 256) p.tile = v;
        }
    }
    p.n_tiles = (int)((n + p.tile - 1) / p.tile);
    int occ = 0;
    switch (ti) {
        case 1: occ = sym_occupancy<1, false>(); break;
        case 2: occ = sym_occupancy<2, false>(); break;
        case 4: occ = sym_occupancy<4, false>(); break;
        case 6: occ = sym_occupancy<6, false>(); break;
        default: occ = sym_occupancy<8, false>(); break;
    }
    if (occ <= 0) occ = 2;
    p.ctas_per_sm = occ;
    const long long slots = (long long)sm_count * occ;
    // I-blocks owned by this rank: one per group of `world` blocks, snake order (see reduce_sym_kernel)
    auto owned_block = [&](int k) { return world * k + ((k & 1) ? world - 1 - rank : rank); };
    int my_blocks = 0;
    while (owned_block(my_blocks) < p.nb_I) ++my_blocks;                     // owned_block is increasing in k
    // panels: bound the P_j footprint (one 3 x n plane per owned I-block of the panel)
    long long budget = 16LL << 30;
    const char* env_b = getenv("ORBITAL_B200_SYM_PJ_BYTES");
    if (env_b) budget = atoll(env_b);
    const long long per_block = 3 * n * 8;
    p.panel_blocks = (int)std::max<long long>(1, std::min<long long>(std::max(1, my_blocks), budget / per_block));
    // chunking: ~32 items per resident slot in every panel (an item is one I-block x one chunk of tiles;
    // about half of the (block, chunk) grid lies in the upper triangle)
    long long want_chunks = (64 * slots + p.panel_blocks - 1) / p.panel_blocks;
    const char* env_c = getenv("ORBITAL_B200_SYM_CHUNKS");
    if (env_c) want_chunks = atoi(env_c);
    want_chunks = std::max<long long>(1, std::min<long long>(std::min<long long>(want_chunks, 512), p.n_tiles));
    p.chunk_tiles = (int)((p.n_tiles + want_chunks - 1) / want_chunks);
    p.n_chunks = (p.n_tiles + p.chunk_tiles - 1) / p.chunk_tiles;
    cudaError_t e;
    if ((e = cudaMalloc(&p.Pi, sizeof(double) * 3 * n * p.n_chunks)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&p.Pj, sizeof(double) * 3 * n * p.panel_blocks)) != cudaSuccess) return e;
    for (int ka = 0; ka < my_blocks; ka += p.panel_blocks) {
        SymPanel pan;
        pan.ka = ka;
        pan.kb = std::min(my_blocks, ka + p.panel_blocks);
        std::vector<SymItem> items;
        for (int k = pan.ka; k < pan.kb; ++k) {
            const int I = owned_block(k);
            const int first_tile = (int)(((long long)I * p.B) / p.tile);
            for (int c = first_tile / p.chunk_tiles; c < p.n_chunks; ++c) {
                SymItem it;
                it.I = I;
                it.t0 = std::max(c * p.chunk_tiles, first_tile);
                it.t1 = std::min((c + 1) * p.chunk_tiles, p.n_tiles);
                it.chunk = c;
                it.slot = k - pan.ka;
                if (it.t1 > it.t0) items.push_back(it);
            }
        }
        std::stable_sort(items.begin(), items.end(),
                         [](const SymItem& a, const SymItem& b) { return (a.t1 - a.t0) > (b.t1 - b.t0); });
        pan.n_items = (int)items.size();
        if (pan.n_items == 0) continue;
        if ((e = cudaMalloc(&pan.d_items, sizeof(SymItem) * items.size())) != cudaSuccess) return e;
        if ((e = cudaMemcpy(pan.d_items, items.data(), sizeof(SymItem) * items.size(), cudaMemcpyHostToDevice)) !=
            cudaSuccess)
            return e;
        p.panels.push_back(pan);
    }
    p.valid = true;
    return cudaSuccess;
}

// The mass-free variant needs one common mass and no partial I-block or tile.
bool sym_uniform(const SymPlan& p, const StepParams& sp) {
    const char* v = getenv("ORBITAL_B200_SYM_UNI");      // "0": keep the per-pair mass multiplies (cross-check)
    const bool off = v && v[0] == '0';
    return !off && sp.uniform_mass != 0.0 && p.n % p.tile == 0 && p.n % p.B == 0;
}

bool sym_tail_applicable(const SymPlan& p) { return p.world == 1 && p.panels.size() == 1; }

cudaError_t launch_force_sym(const DeviceState& s, const StepParams& sp, const SymPlan& p, bool detect,
                             cudaStream_t st, int* launches, int tail_mode) {
    SymTail tail = {};
    if (tail_mode != kSymTailNone) {
        if (!sym_tail_applicable(p)) return cudaErrorInvalidValue;
        tail.kick = tail_mode == kSymTailKick;
        tail.close = tail_mode == kSymTailClose || tail_mode == kSymTailCloseNext;
        tail.next = tail_mode == kSymTailCloseNext;
        tail.pos4 = s.pos4; tail.vel = s.vel; tail.vf32 = s.vf32; tail.hist = s.hist; tail.hist_cap = s.hist_cap;
        tail.h = sp.h; tail.dt = sp.dt; tail.dt32 = sp.dt32; tail.ctl_rw = s.ctl;
    }
    SymArgs a;
    a.pos4 = s.pos4;
    a.radius = s.radius;
    a.Pi = p.Pi;
    a.Pj = p.Pj;
    a.n = s.n;
    a.n_tiles = p.n_tiles;
    a.tile = p.tile;
    a.eps2 = sp.eps2;
    a.rmax1 = sp.rmax1;
    a.rmax2 = sp.rmax2;
    a.rmax1_idx = sp.rmax1_idx;
    a.ctl = s.ctl;
    a.pairs = s.pairs;
    const bool uni = sym_uniform(p, sp);
    const double scale = uni ? sp.G * sp.uniform_mass : sp.G;
    bool first = true;
    for (const SymPanel& pan : p.panels) {
        a.items = pan.d_items;
        cudaError_t e;
#define ORB_SYM_CASE(T)                                                                                         \
    case T:                                                                                                     \
        if (uni)                                                                                                \
            e = detect ? launch_sym_t<T, true, true>(a, pan.n_items, st)                                        \
                       : launch_sym_t<T, false, true>(a, pan.n_items, st);                                      \
        else                                                                                                    \
            e = detect ? launch_sym_t<T, true, false>(a, pan.n_items, st)                                       \
                       : launch_sym_t<T, false, false>(a, pan.n_items, st);                                     \
        break;
        switch (p.ti) {
            ORB_SYM_CASE(1)
            ORB_SYM_CASE(2)
            ORB_SYM_CASE(4)
            ORB_SYM_CASE(6)
            ORB_SYM_CASE(8)
            default: return cudaErrorInvalidValue;
        }
#undef ORB_SYM_CASE
        if (e != cudaSuccess) return e;
        const int grid = (int)((s.n + kRedX - 1) / kRedX);
        reduce_sym_kernel<<<grid, 256, 0, st>>>(p.Pi, p.Pj, s.acc, s.n, p.B, p.tile, p.chunk_tiles, p.n_chunks, p.rank, p.world,
                                                pan.ka, pan.kb, scale, first ? 0 : 1, s.ctl, tail);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (launches) *launches += 2;
        first = false;
    }
    if (first) return cudaMemsetAsync(s.acc, 0, sizeof(double) * 3 * s.n, st);
    return cudaSuccess;
}

}  // namespace orb
