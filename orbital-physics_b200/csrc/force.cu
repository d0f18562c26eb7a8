// force.cu -- all-pairs softened Newtonian acceleration kernels (sm_100a).
//
// Replaces pairwise_accelerations (reference core/physics.py:125-159) and the
// overlap *detection* half of ObjectCollection.handle_collisions
// (core/physics.py:513-518).
//
//   force_fast_kernel<TI,DETECT>   roofline kernel. Source bodies {x,y,z,m} are
//       streamed through a 4-stage shared-memory ring by 1-D bulk TMA
//       (cp.async.bulk + mbarrier, SASS UBLKCP); every thread keeps TI targets
//       and their accumulators in registers and reads each source with two
//       broadcast LDS.128. Per interaction: 3 DADD + 3 DFMA (r^2) + MUFU.RSQ64H
//       seed + 7 DMUL/DFMA (m * r^-3 to full fp64 accuracy) + 3 DFMA = 16 FP64
//       pipe instructions. FP64-pipe bound; HBM traffic is 56 B/body/pass.
//       Work is a 2-D grid of (target block x source slab) items sized to fill
//       148 SMs in whole waves; slab partials are combined in fixed order.
//   force_faithful_kernel<TW,DETECT>  bit-exact: the reference's rounding sequence with IEEE sqrt/div
//       (SURVEY.md A.1); lanes evaluate pair terms in parallel, one lane per target adds them in ascending j.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "force_common.cuh"
#include "kernels.h"

namespace orb {

// ===========================================================================
// Fast kernel (one-sided: every target sums over all sources)
// ===========================================================================
struct FastArgs {
    const double4* pos4;
    const double* radius;
    double* out;               // acc (3 x out_n SoA) when slabs == 1, else scratch (slabs x 3 x n_tgt)
    long long out_n;           // SoA stride of `out`
    long long out_off;         // index offset of target tgt_lo inside `out`
    long long slab_stride;     // elements between slab partial planes (0 when slabs == 1)
    long long n_src, tgt_lo, n_tgt;
    int slabs, tiles_per_slab, n_tiles;
    double eps2, G;
    double rmax1, rmax2;
    long long rmax1_idx;
    Ctl* ctl;
    long long* pairs;
};

template <int TI, bool DETECT>
__global__ void __launch_bounds__(kFastThreads, (TI >= 6 ? 2 : (TI >= 3 ? 3 : 4)))
force_fast_kernel(const FastArgs g) {
    if (g.ctl->halted) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tiles = reinterpret_cast<double2*>(smem_raw);                       // kStages x kTile x 2 double2
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * kTile * 32);
    uint64_t* empty = full + kStages;

    const int tid = threadIdx.x;
    const int slab = blockIdx.x % g.slabs;
    const long long tblock = blockIdx.x / g.slabs;
    const long long cta_lo = g.tgt_lo + tblock * (long long)(kFastThreads * TI);
    const long long tgt_hi = g.tgt_lo + g.n_tgt;
    const long long cta_hi = min(cta_lo + (long long)(kFastThreads * TI), tgt_hi);

    const int tile_lo = slab * g.tiles_per_slab;
    const int tile_hi = min(tile_lo + g.tiles_per_slab, g.n_tiles);
    const int ntiles = tile_hi - tile_lo;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kFastWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t) {   // tile index relative to tile_lo
        const int s = t % kStages;
        const long long j0 = (long long)(tile_lo + t) * kTile;
        const int cnt = (int)min((long long)kTile, g.n_src - j0);
        const uint32_t bytes = (uint32_t)cnt * 32u;
        mbar_expect_tx(&full[s], bytes);
        tma_load_1d(tiles + (size_t)s * kTile * 2, g.pos4 + j0, bytes, &full[s]);
    };
    if (tid == 0) {
        const int pre = min(kStages, ntiles);
        for (int t = 0; t < pre; ++t) issue(t);
    }

    double xi[TI], yi[TI], zi[TI], ax[TI], ay[TI], az[TI];
    long long idx[TI];
    int maxhi[TI], thr[TI];
    double Ri[TI];
#pragma unroll
    for (int k = 0; k < TI; ++k) {
        idx[k] = cta_lo + (long long)k * kFastThreads + tid;
        const long long ld = min(idx[k], tgt_hi - 1);      // tail threads shadow the last target, never store
        const double4 p = g.pos4[ld];
        xi[k] = p.x; yi[k] = p.y; zi[k] = p.z;
        ax[k] = ay[k] = az[k] = 0.0;
        maxhi[k] = 0;
        thr[k] = 0x7fffffff;
        Ri[k] = 0.0;
        if (DETECT) {
            Ri[k] = g.radius[ld];
            const double partner = (ld == g.rmax1_idx) ? g.rmax2 : g.rmax1;
            const double rs = Ri[k] + partner;
            // overlap => r2 <= rs^2(1+tiny) + eps2 => seed >= rsqrt(rs^2+eps2) (1 - 2^-19)
            const double bound = (1.0 / sqrt(fma(rs, rs, g.eps2))) * (1.0 - 1.52587890625e-05);
            thr[k] = __double2hiint(bound) - 1;
        }
    }

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages;
        mbar_wait(&full[s], (uint32_t)((t / kStages) & 1));
        const double2* tile = tiles + (size_t)s * kTile * 2;
        const long long j0 = (long long)(tile_lo + t) * kTile;
        const int cnt = (int)min((long long)kTile, g.n_src - j0);
        const bool diag = (j0 < cta_hi) && (j0 + cnt > cta_lo);        // CTA-uniform
        if (diag)
            tile_loop<TI, DETECT, true>(tile, cnt, j0, g.eps2, xi, yi, zi, idx, ax, ay, az, maxhi);
        else
            tile_loop<TI, DETECT, false>(tile, cnt, j0, g.eps2, xi, yi, zi, idx, ax, ay, az, maxhi);
        if (DETECT) {
#pragma unroll
            for (int k = 0; k < TI; ++k) {
                if (maxhi[k] >= thr[k] && idx[k] < tgt_hi)
                    rescan_tile(tile, cnt, j0, idx[k], xi[k], yi[k], zi[k], Ri[k], g.radius, g.ctl, g.pairs);
                maxhi[k] = 0;
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        // refill the stage freed one tile ago: the producer never waits on the current tile's stragglers
        if (tid == 0 && t >= 1 && (t - 1 + kStages) < ntiles) {
            const int tp = t - 1;
            mbar_wait(&empty[tp % kStages], (uint32_t)((tp / kStages) & 1));
            issue(tp + kStages);
        }
    }

    // epilogue
    const double scale = (g.slabs == 1) ? g.G : 1.0;
    double* ox = g.out + (long long)slab * g.slab_stride;
#pragma unroll
    for (int k = 0; k < TI; ++k) {
        if (idx[k] < tgt_hi) {
            const long long o = g.out_off + (idx[k] - g.tgt_lo);
            ox[o] = scale * ax[k];
            ox[o + g.out_n] = scale * ay[k];
            ox[o + 2 * g.out_n] = scale * az[k];
        }
    }
}

// acc[i] = G * (((p_0 + p_1) + p_2) + ...)  -- fixed slab order: deterministic
__global__ void __launch_bounds__(256) reduce_slabs_kernel(const double* __restrict__ scratch, double* acc,
                                                           long long n_tgt, long long acc_n, long long acc_off,
                                                           int slabs, double G, const Ctl* ctl) {
    if (ctl->halted) return;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_tgt) return;
    const long long plane = 3 * n_tgt;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double s = scratch[c * n_tgt + i];
        for (int k = 1; k < slabs; ++k) s += scratch[k * plane + c * n_tgt + i];
        acc[acc_off + i + c * acc_n] = G * s;
    }
}

template <int TI, bool DETECT>
static cudaError_t launch_fast_t(const FastArgs& a, const FastPlan& plan, cudaStream_t st) {
    auto kern = force_fast_kernel<TI, DETECT>;
    static DeviceOnce attr_set;                                 // the attribute is per device
    if (attr_set.first()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<plan.grid, plan.block, plan.smem, st>>>(a);
    return cudaGetLastError();
}

template <int TI, bool DETECT>
static int occupancy_t(int smem) {
    int nb = 0;
    auto kern = force_fast_kernel<TI, DETECT>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kFastThreads, smem) != cudaSuccess) nb = 0;
    return nb;
}

static int fast_smem_bytes() { return kStages * kTile * 32 + 2 * kStages * 8 + 64; }

static int occupancy_for(int ti) {
    const int smem = fast_smem_bytes();
    switch (ti) {
        case 1: return occupancy_t<1, false>(smem);
        case 2: return occupancy_t<2, false>(smem);
        case 4: return occupancy_t<4, false>(smem);
        case 6: return occupancy_t<6, false>(smem);
        default: return occupancy_t<8, false>(smem);
    }
}

const char* fast_kernel_name(int ti, bool detect) {
    static thread_local char buf[64];
    snprintf(buf, sizeof buf, "force_fast_kernel<%d,%s>", ti, detect ? "true" : "false");
    return buf;
}

// Choose targets-per-thread and the number of source slabs so that the grid is
// (close to) a whole number of full waves of the 148 x ctas_per_sm resident slots.
FastPlan plan_fast(long long n_tgt, long long n_src, int sm_count) {
    FastPlan p;
    const char* env_ti = getenv("ORBITAL_B200_TI");
    const long long per_sm = (n_tgt + sm_count - 1) / sm_count;
    // measured on B200 at N=262144 (profiles/r1_sweep.txt): TI=8 61.7 %, 6 60.8 %, 4 59.7 %, 2 57.4 %, 1 55.7 % of
    // the DFMA peak -- more targets per thread amortise the LDS/MUFU/issue overhead of each source.
    int ti = per_sm >= 1024 ? 8 : (per_sm >= 256 ? 4 : (per_sm >= 64 ? 2 : 1));
    if (env_ti) {
        const int v = atoi(env_ti);
        if (v == 1 || v == 2 || v == 4 || v == 6 || v == 8) ti = v;
    }
    p.ti = ti;
    p.block = kFastThreads;
    p.smem = fast_smem_bytes();
    int occ = occupancy_for(ti);
    if (occ <= 0) occ = 2;
    p.ctas_per_sm = occ;
    const long long slots = (long long)sm_count * occ;
    const long long tblocks = (n_tgt + (long long)kFastThreads * ti - 1) / ((long long)kFastThreads * ti);
    const int n_tiles = (int)((n_src + kTile - 1) / kTile);
    // candidate slab counts: best wave efficiency with at least ~6 waves when the problem is big enough
    int best_s = 1;
    double best_eff = -1.0;
    const int max_s = (int)std::min<long long>(n_tiles, 64);
    const char* env_s = getenv("ORBITAL_B200_SLABS");
    for (int s = 1; s <= max_s; ++s) {
        const int tps = (n_tiles + s - 1) / s;
        const int s_eff = (n_tiles + tps - 1) / tps;      // slabs actually non-empty
        if (s_eff != s) continue;
        const long long items = tblocks * s;
        const long long waves = (items + slots - 1) / slots;
        // every item costs tps tiles (last slab may be shorter: ignore); time ~ waves * tps
        const double eff = (double)n_tiles * tblocks / ((double)waves * slots * tps);
        // mild penalty per extra slab (partial-sum traffic + reduce launch)
        const double score = eff - 0.002 * (s - 1) - (s > 1 ? 0.01 : 0.0);
        if (score > best_eff) { best_eff = score; best_s = s; }
    }
    if (env_s) {
        const int v = atoi(env_s);
        if (v >= 1 && v <= n_tiles) best_s = v;
    }
    p.tiles_per_slab = (n_tiles + best_s - 1) / best_s;
    p.slabs = (n_tiles + p.tiles_per_slab - 1) / p.tiles_per_slab;
    p.grid = (int)(tblocks * p.slabs);
    return p;
}

cudaError_t launch_force_fast(const DeviceState& s, const StepParams& p, const FastPlan& plan, bool detect,
                              cudaStream_t st, int* launches) {
    FastArgs a;
    const long long n_tgt = s.tgt_hi - s.tgt_lo;
    a.pos4 = s.pos4;
    a.radius = s.radius;
    a.n_src = s.n;
    a.tgt_lo = s.tgt_lo;
    a.n_tgt = n_tgt;
    a.slabs = plan.slabs;
    a.tiles_per_slab = plan.tiles_per_slab;
    a.n_tiles = (int)((s.n + kTile - 1) / kTile);
    a.eps2 = p.eps2;
    a.G = p.G;
    a.rmax1 = p.rmax1;
    a.rmax2 = p.rmax2;
    a.rmax1_idx = p.rmax1_idx;
    a.ctl = s.ctl;
    a.pairs = s.pairs;
    if (plan.slabs == 1) {
        a.out = s.acc; a.out_n = s.n; a.out_off = s.tgt_lo; a.slab_stride = 0;
    } else {
        a.out = s.scratch; a.out_n = n_tgt; a.out_off = 0; a.slab_stride = 3 * n_tgt;
    }
    cudaError_t e;
#define ORB_FAST_CASE(T)                                                                        \
    case T:                                                                                     \
        e = detect ? launch_fast_t<T, true>(a, plan, st) : launch_fast_t<T, false>(a, plan, st); \
        break;
    switch (plan.ti) {
        ORB_FAST_CASE(1)
        ORB_FAST_CASE(2)
        ORB_FAST_CASE(4)
        ORB_FAST_CASE(6)
        ORB_FAST_CASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef ORB_FAST_CASE
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    if (plan.slabs > 1) {
        const int blk = 256;
        const int grd = (int)((n_tgt + blk - 1) / blk);
        reduce_slabs_kernel<<<grd, blk, 0, st>>>(s.scratch, s.acc, n_tgt, s.n, s.tgt_lo, plan.slabs, p.G, s.ctl);
        e = cudaGetLastError();
        if (launches) ++*launches;
    }
    return e;
}

// ===========================================================================
// Faithful kernel: one thread per target, ascending j, exact rounding sequence
// ===========================================================================
// Bit-exact force kernel.  A warp owns TW targets; for every chunk of 32 sources lane l evaluates the TW
// terms (target t, source l) with the reference's rounding sequence and parks them in shared memory;
// lanes t < TW then add their target's 32 terms in ascending source order.  The expensive part (IEEE
// sqrt + two divides per pair) is spread over all lanes; only the three adds per pair stay sequential,
// which is what bit-exactness requires (physics.py:154 accumulates in loop order).
constexpr int kFaithWarps = 4;

template <int TW, bool DETECT>
__global__ void __launch_bounds__(32 * kFaithWarps)
force_faithful_kernel(const double4* __restrict__ pos4, const double* __restrict__ radius, double* acc, long long n,
                      long long tgt_lo, long long tgt_hi, double eps2, double G, Ctl* ctl, long long* pairs) {
    if (ctl->halted) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // per warp: terms[3][TW][33] (+1 padding: conflict-free row reads) and its targets {x,y,z,radius}
    typedef double TermBlock[3][TW][33];
    TermBlock& terms = reinterpret_cast<TermBlock*>(smem_raw)[warp];
    double4* tgt = reinterpret_cast<double4*>(smem_raw + kFaithWarps * sizeof(TermBlock)) + warp * TW;
    const long long t_base = tgt_lo + ((long long)blockIdx.x * kFaithWarps + warp) * TW;
    if (t_base >= tgt_hi) return;                           // warp-uniform
    const int nt = (int)min((long long)TW, tgt_hi - t_base);
    if (lane < TW) {
        const long long ti = min(t_base + lane, tgt_hi - 1);
        const double4 p = pos4[ti];
        tgt[lane] = make_double4(p.x, p.y, p.z, DETECT ? radius[ti] : 0.0);
    }
    __syncwarp();
    double bx = 0.0, by = 0.0, bz = 0.0;                    // lane t < nt: accumulator of target t (physics.py:132)
    for (long long j0 = 0; j0 < n; j0 += 32) {
        const long long j = j0 + lane;
        const bool jvalid = j < n;
        const double4 q = pos4[jvalid ? j : n - 1];
        const double Gm = __dmul_rn(G, q.w);                // G * mj (physics.py:151)
        const double Rj = (DETECT && jvalid) ? radius[j] : 0.0;
#pragma unroll
        for (int t = 0; t < TW; ++t) {
            double tx = 0.0, ty = 0.0, tz = 0.0;            // masked pairs park +0.0: x + 0.0 == x (x is never -0.0)
            const long long it = t_base + t;
            if (jvalid && j != it && t < nt) {
                const double4 me = tgt[t];
                const double dx = __dsub_rn(q.x, me.x), dy = __dsub_rn(q.y, me.y), dz = __dsub_rn(q.z, me.z);
                pair_term_faithful(dx, dy, dz, eps2, Gm, tx, ty, tz);
                if (DETECT && j > it) {
                    if (overlap_exact(-dx, -dy, -dz, me.w, Rj)) record_overlap(ctl, pairs, it, j);
                }
            }
            terms[0][t][lane] = tx;
            terms[1][t][lane] = ty;
            terms[2][t][lane] = tz;
        }
        __syncwarp();
        if (lane < nt) {
#pragma unroll 8
            for (int l = 0; l < 32; ++l) {                  // ascending j: the reference's accumulation order
                bx = __dadd_rn(bx, terms[0][lane][l]);
                by = __dadd_rn(by, terms[1][lane][l]);
                bz = __dadd_rn(bz, terms[2][lane][l]);
            }
        }
        __syncwarp();
    }
    if (lane < nt) {
        const long long i = t_base + lane;
        acc[i] = bx;
        acc[i + n] = by;
        acc[i + 2 * n] = bz;
    }
}

// Two-pass bit-exact force for mid-size systems.
// Pass 1 (faithful_pairs_kernel): inv_r3 of every unordered pair, once (physics.py:145-148: IEEE sqrt and two
// divides, ~100 FP64 instructions) into a symmetric matrix; the overlap test of handle_collisions
// (physics.py:517-518) rides along (i < j).  Layout: column blocks of 32 bodies, each holding n_rows rows of 32
// doubles, so a 32 x 32 tile is one contiguous 8 KiB slab -- written whole by pass 1 (directly and transposed
// through shared memory) and streamed by pass 2 with 1-D bulk TMA.
__device__ __forceinline__ long long pair_slot(long long row, long long col, long long n_rows) {
    return ((col >> 5) * n_rows + row) * 32 + (col & 31);
}

template <bool DETECT>
__global__ void __launch_bounds__(256) faithful_pairs_kernel(const double4* __restrict__ pos4,
                                                             const double* __restrict__ radius, double* invr3,
                                                             long long n, long long n_rows, double eps2, Ctl* ctl,
                                                             long long* pairs) {
    if (ctl->halted) return;
    // triangular grid: block t -> tile (bi <= bj), t = bj (bj + 1) / 2 + bi
    int bj = (int)((sqrtf(8.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
    while ((long long)(bj + 1) * (bj + 2) / 2 <= (long long)blockIdx.x) ++bj;
    while ((long long)bj * (bj + 1) / 2 > (long long)blockIdx.x) --bj;
    const int bi = (int)((long long)blockIdx.x - (long long)bj * (bj + 1) / 2);
    __shared__ double tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const long long i0 = (long long)bi * 32, j0 = (long long)bj * 32;
    const long long j = j0 + tx;
    const double4 pj = pos4[min(j, n - 1)];
    const double Rj = DETECT ? radius[min(j, n - 1)] : 0.0;
    if (bi < bj && j0 + 32 <= n) {
        // Whole off-diagonal tile (all but ~2 n / 32 of the n^2 / 2048 tiles): every pair is valid, so no per-element
        // predicates, and both copies of the tile are contiguous 8 KiB slabs addressed with 32-bit offsets.  The kernel
        // is issue-bound (ncu: issue slots 71 % busy, FP64 pipe 36 %): 112 warp instructions per 32 pairs, only 31 of
        // them FP64 / MUFU -- index arithmetic and predication are what there is to save.
        __shared__ double4 srow[32];
        __shared__ double srad[32];
        if (ty == 0) {
            srow[tx] = pos4[i0 + tx];
            if (DETECT) srad[tx] = radius[i0 + tx];
        }
        __syncthreads();
        double* dst = invr3 + ((long long)bj * n_rows + i0) * 32 + tx;        // rows i0.., column block bj
        double* dstT = invr3 + ((long long)bi * n_rows + j0) * 32 + tx;       // rows j0.., column block bi (mirror)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = ty + 8 * k;
            const double4 pi = srow[r];                                        // broadcast
            const double dx = __dsub_rn(pj.x, pi.x), dy = __dsub_rn(pj.y, pi.y), dz = __dsub_rn(pj.z, pi.z);   // :145
            const double r2 = __dadd_rn(dot3_numpy(dx, dy, dz), eps2);                                         // :146
            const double inv_r = __ddiv_rn(1.0, __dsqrt_rn(r2));                                               // :147
            const double v = __ddiv_rn(inv_r, r2);                                                             // :148
            if (DETECT) {
                if (overlap_exact(-dx, -dy, -dz, srad[r], Rj)) record_overlap(ctl, pairs, i0 + r, j);
            }
            dst[r * 32] = v;
            tile[r][tx] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = ty + 8 * k;
            dstT[r * 32] = tile[tx][r];
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = ty + 8 * k;
        const long long i = i0 + r;
        double v = 0.0;
        if (i < j && j < n) {
            const double4 pi = pos4[i];
            const double dx = __dsub_rn(pj.x, pi.x), dy = __dsub_rn(pj.y, pi.y), dz = __dsub_rn(pj.z, pi.z);   // :145
            const double r2 = __dadd_rn(dot3_numpy(dx, dy, dz), eps2);                                         // :146
            const double inv_r = __ddiv_rn(1.0, __dsqrt_rn(r2));                                               // :147
            v = __ddiv_rn(inv_r, r2);                                                                          // :148
            if (DETECT) {
                if (overlap_exact(-dx, -dy, -dz, radius[i], Rj)) record_overlap(ctl, pairs, i, j);
            }
            invr3[pair_slot(i, j, n_rows)] = v;
        }
        tile[r][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = ty + 8 * k;                 // row of the transposed tile: body j0 + r, column body i0 + tx
        const long long jj = j0 + r, ii = i0 + tx;
        if (ii < jj && jj < n) invr3[pair_slot(jj, ii, n_rows)] = tile[tx][r];
        else if (ii == jj && jj < n) invr3[pair_slot(jj, ii, n_rows)] = 0.0;   // self term: +0.0 in pass 2
    }
}

// Pass 2 (faithful_rows_kernel): one warp per column block, lane = target.  For every source j in ascending order
// (the reference's accumulation order, physics.py:154) finish the term with the target-centric tail
// ((G m_j) inv_r3) (r_j - r_i) -- bit-identical for both bodies of a pair because rounding is sign-symmetric --
// and add it.  The block's slabs and their source bodies arrive through a kRowStages-deep bulk-TMA ring (the warp
// is alone on its scheduler: nobody else hides the latency).  Rows past the end are zero (memset at allocation)
// and such sources get a zero position and mass: the term is (+-)0 and x + (+-0.0) == x for a running sum that
// started at +0.0; the diagonal entry is 0 for the same reason.  No predicates in the loop.
// (A variant with one component of one target per lane -- 4 FP64 instructions per source instead of 10, four warps
// per block -- was slower: 0.154 vs 0.128 ms at n = 4096, 1.75 vs 0.91 ms at n = 16384.)
#ifndef ORB_ROW_STAGES
#define ORB_ROW_STAGES 4
#endif
constexpr int kRowStages = ORB_ROW_STAGES;       // x 9 KiB of shared memory (8 or 16 stages: no faster at n = 4096,
                                                 // slower at n = 16384 where they cut the resident CTAs per SM)
constexpr int kRowSmem = kRowStages * (8192 + 1024) + 1024 + 64;

// FUSE_TAIL: the rest of the leapfrog step rides along -- second half-kick (engine.py:81-82), history append
// (engine.py:88-92) and, by the last CTA to finish, the step bookkeeping of advance_kernel -- which takes the step
// from five launches to three (what counts below a few thousand bodies, where a step is launch latency).
struct RowsTail {
    double* vel;
    const uint8_t* vf32;
    double* hist;
    long long hist_cap;
    double h;
};

template <bool FUSE_TAIL>
__global__ void __launch_bounds__(32) faithful_rows_kernel(const double4* __restrict__ pos4,
                                                           const double* __restrict__ invr3, double* acc, long long n,
                                                           long long n_rows, double G, Ctl* ctl, const RowsTail tail) {
    if (ctl->halted) return;
    extern __shared__ __align__(128) unsigned char rows_smem[];
    double (*mt)[32 * 32] = reinterpret_cast<double (*)[32 * 32]>(rows_smem);                  // slabs of the matrix
    double4 (*praw)[32] = reinterpret_cast<double4 (*)[32]>(rows_smem + kRowStages * 8192);    // their source bodies
    double4* sp = reinterpret_cast<double4*>(rows_smem + kRowStages * (8192 + 1024));          // ... as {x,y,z,G*m}
    uint64_t* full = reinterpret_cast<uint64_t*>(rows_smem + kRowStages * (8192 + 1024) + 1024);
    const int lane = threadIdx.x;
    const long long i = blockIdx.x * 32LL + lane;
    const double4 me = pos4[min(i, n - 1)];
    const double* blk = invr3 + (long long)blockIdx.x * n_rows * 32;
    const int ntile = (int)(n_rows / 32);
    auto issue = [&](int t) {
        const int st = t % kRowStages;
        const uint32_t pbytes = (uint32_t)min(32LL, n - 32LL * t) * 32u;
        mbar_expect_tx(&full[st], 8192u + pbytes);
        tma_load_1d(mt[st], blk + (long long)t * 1024, 8192u, &full[st]);
        tma_load_1d(praw[st], pos4 + 32LL * t, pbytes, &full[st]);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kRowStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        for (int t = 0; t < min(kRowStages, ntile); ++t) issue(t);
    }
    __syncwarp();
    double bx = 0.0, by = 0.0, bz = 0.0;                        // physics.py:132
    for (int t = 0; t < ntile; ++t) {
        const int stage = t % kRowStages;
        mbar_wait(&full[stage], (uint32_t)((t / kRowStages) & 1));
        {
            const double4 q = (32LL * t + lane < n) ? praw[stage][lane] : make_double4(0.0, 0.0, 0.0, 0.0);
            sp[lane] = make_double4(q.x, q.y, q.z, __dmul_rn(G, q.w));         // G * m_j            :151
        }
        __syncwarp();
        const double* m = mt[stage] + lane;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const double4 q = sp[k];                                            // broadcast
            const double s = __dmul_rn(q.w, m[k * 32]);                         // (G m_j) inv_r3     :151
            bx = __dadd_rn(bx, __dmul_rn(s, __dsub_rn(q.x, me.x)));             // a += s * rij       :154
            by = __dadd_rn(by, __dmul_rn(s, __dsub_rn(q.y, me.y)));
            bz = __dadd_rn(bz, __dmul_rn(s, __dsub_rn(q.z, me.z)));
        }
        __syncwarp();
        if (lane == 0 && t + kRowStages < ntile) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // our reads, then the async refill
            issue(t + kRowStages);
        }
    }
    if (i < n) {
        acc[i] = bx;
        acc[i + n] = by;
        acc[i + 2 * n] = bz;
    }
    if (FUSE_TAIL) {
        if (i < n) {
            const bool f32 = tail.vf32[i] != 0;
            tail.vel[i] = kick_faithful(tail.vel[i], tail.h, bx, f32);
            tail.vel[i + n] = kick_faithful(tail.vel[i + n], tail.h, by, f32);
            tail.vel[i + 2 * n] = kick_faithful(tail.vel[i + 2 * n], tail.h, bz, f32);
            if (tail.hist_cap > 0 && ctl->overlap_count == 0) {            // a halting step appends on the host
                double* row = tail.hist + ((ctl->hist_count % tail.hist_cap) * n + i) * 3;
                row[0] = me.x; row[1] = me.y; row[2] = me.z;
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            if (atomicAdd(&ctl->rows_done, 1u) == gridDim.x - 1) {         // every CTA has read hist_count by now
                ctl->rows_done = 0;
                ctl->steps_done += 1;
                if (ctl->overlap_count > 0)
                    ctl->halted = 1;
                else if (tail.hist_cap > 0)
                    ctl->hist_count += 1;
            }
        }
    }
}

// Pass 2, producer / consumer form (default).  The ordered sum is a chain of N dependent additions per target and
// component -- 8 cycles each, 17 us at N = 4,096, and nothing can shorten it -- so the kernel is built to make that
// chain the ONLY thing its warp does: a CTA owns 32 targets (lane = target); four producer warps finish the terms
// ((G m_j) inv_r3) (r_j - r_i) of 8 sources each per 32-source tile (7 FP64 per term, unordered, embarrassingly
// parallel) into a shared-memory ring, and one consumer warp adds them in ascending source order: 3 LDS + 3 DADD
// per source.  The matrix slabs and their source bodies arrive through a bulk-TMA ring as before.
// Measured (profiles/r2_sweep_faithful_step*.txt, whole step): N = 1,024 27.8 us (one warp per 32 targets doing
// everything: 30.8), 2,048 49.8 (55.3), 4,096 111 (119); from 8,192 up the one-warp kernel wins (298 vs 408 us) because
// its small CTAs pack six to an SM while this one's 108 KiB of rings allow one -- so it serves up to 148 blocks.
// What bounds BOTH at N = 4,096 is the 134 MB pair matrix coming back from DRAM at ~1.7 TB/s (ncu: L2 hit rate 4 %;
// neither a 12-deep ring, nor an odd slab stride against channel aliasing, nor plain register-prefetched loads
// instead of TMA moved it): the ordered chain itself is 17 us.
#ifndef R2_PROD
#define R2_PROD 4
#endif
#ifndef R2_TSTAGES
#define R2_TSTAGES 3
#endif
#ifndef R2_MSTAGES
#define R2_MSTAGES 4
#endif
constexpr int kR2Prod = R2_PROD;                 // producer warps
constexpr int kR2TStages = R2_TSTAGES;                    // term ring: 3 x [3][32 sources][32 targets] doubles (24 KiB each)
constexpr int kR2MStages = R2_MSTAGES;                    // matrix-slab ring: 4 x (8 KiB slab + 1 KiB of source bodies)
constexpr int kR2TermBytes = 3 * 32 * 32 * 8;
constexpr int kR2Smem = kR2TStages * kR2TermBytes + kR2MStages * (8192 + 1024) + 256;

template <bool FUSE_TAIL>
__global__ void __launch_bounds__(32 * (1 + kR2Prod)) faithful_rows2_kernel(const double4* __restrict__ pos4,
                                                                           const double* __restrict__ invr3, double* acc,
                                                                           long long n, long long n_rows, double G,
                                                                           Ctl* ctl, const RowsTail tail) {
    if (ctl->halted) return;
    extern __shared__ __align__(128) unsigned char rows_smem[];
    double (*terms)[3][32][32] = reinterpret_cast<double (*)[3][32][32]>(rows_smem);
    double (*mt)[32 * 32] = reinterpret_cast<double (*)[32 * 32]>(rows_smem + kR2TStages * kR2TermBytes);
    double4 (*praw)[32] = reinterpret_cast<double4 (*)[32]>(rows_smem + kR2TStages * kR2TermBytes + kR2MStages * 8192);
    uint64_t* bars = reinterpret_cast<uint64_t*>(rows_smem + kR2TStages * kR2TermBytes + kR2MStages * (8192 + 1024));
    uint64_t* mfull = bars;                       // TMA landed                      (1 arrival + tx bytes)
    uint64_t* mempty = bars + kR2MStages;         // producers are done with a slab  (kR2Prod arrivals)
    uint64_t* tfull = bars + 2 * kR2MStages;      // producers filled a term stage   (kR2Prod arrivals)
    uint64_t* tempty = tfull + kR2TStages;        // consumer drained a term stage   (1 arrival)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;            // 0: consumer, 1..kR2Prod: producers
    const long long i = blockIdx.x * 32LL + lane;
    const double4 me = pos4[min(i, n - 1)];
    const double* blk = invr3 + (long long)blockIdx.x * n_rows * 32;
    const int ntile = (int)(n_rows / 32);
    auto issue = [&](int t) {
        const int st = t % kR2MStages;
        const uint32_t pbytes = (uint32_t)min(32LL, n - 32LL * t) * 32u;
        mbar_expect_tx(&mfull[st], 8192u + pbytes);
        tma_load_1d(mt[st], blk + (long long)t * 1024, 8192u, &mfull[st]);
        tma_load_1d(praw[st], pos4 + 32LL * t, pbytes, &mfull[st]);
    };
    if (threadIdx.x == 32) {                      // producer 0, lane 0 owns the TMA ring
        for (int s = 0; s < kR2MStages; ++s) { mbar_init(&mfull[s], 1); mbar_init(&mempty[s], kR2Prod); }
        for (int s = 0; s < kR2TStages; ++s) { mbar_init(&tfull[s], kR2Prod); mbar_init(&tempty[s], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 32)
        for (int t = 0; t < min(kR2MStages, ntile); ++t) issue(t);

    if (warp > 0) {
        // ---------------- producers: terms of sources [8 p, 8 p + 8) of every tile, lane = target
        const int p = warp - 1;
        for (int t = 0; t < ntile; ++t) {
            const int ms = t % kR2MStages, ts = t % kR2TStages;
            if (p == 0 && lane == 0 && t >= 1 && t - 1 + kR2MStages < ntile) {
                // refill the slab freed one tile ago: never waits on this tile's stragglers
                mbar_wait(&mempty[(t - 1) % kR2MStages], (uint32_t)(((t - 1) / kR2MStages) & 1));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(t - 1 + kR2MStages);
            }
            mbar_wait(&mfull[ms], (uint32_t)((t / kR2MStages) & 1));
            if (t >= kR2TStages) mbar_wait(&tempty[ts], (uint32_t)(((t / kR2TStages) + 1) & 1));
            const double* m = mt[ms] + lane;
            constexpr int kPer = 32 / kR2Prod;
            // all loads first: the term stores below live in the same shared array, so the compiler cannot hoist a
            // load over them by itself (ncu: every source stalled on its own LDS)
            double4 q[kPer];
            double inv[kPer];
#pragma unroll
            for (int kk = 0; kk < kPer; ++kk) {
                const int k = p * kPer + kk;
                q[kk] = praw[ms][k];                                            // broadcast
                inv[kk] = m[k * 32];
            }
            const int valid = (int)min(32LL, n - 32LL * t);                     // sources past the end: zero bodies
#pragma unroll
            for (int kk = 0; kk < kPer; ++kk) {
                const int k = p * kPer + kk;
                if (k >= valid) q[kk] = make_double4(0.0, 0.0, 0.0, 0.0);       // (their inv_r3 rows are zero too)
                const double s = __dmul_rn(__dmul_rn(G, q[kk].w), inv[kk]);     // (G m_j) inv_r3     :151
                terms[ts][0][k][lane] = __dmul_rn(s, __dsub_rn(q[kk].x, me.x));  // s * rij            :154
                terms[ts][1][k][lane] = __dmul_rn(s, __dsub_rn(q[kk].y, me.y));
                terms[ts][2][k][lane] = __dmul_rn(s, __dsub_rn(q[kk].z, me.z));
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&tfull[ts]);
                mbar_arrive(&mempty[ms]);
            }
        }
        return;
    }
    // ---------------- consumer: the ordered sums (physics.py:154 accumulates in loop order)
    double bx = 0.0, by = 0.0, bz = 0.0;                        // physics.py:132
    for (int t = 0; t < ntile; ++t) {
        const int ts = t % kR2TStages;
        mbar_wait(&tfull[ts], (uint32_t)((t / kR2TStages) & 1));
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            bx = __dadd_rn(bx, terms[ts][0][k][lane]);
            by = __dadd_rn(by, terms[ts][1][k][lane]);
            bz = __dadd_rn(bz, terms[ts][2][k][lane]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ts]);
    }
    if (i < n) {
        acc[i] = bx;
        acc[i + n] = by;
        acc[i + 2 * n] = bz;
    }
    if (FUSE_TAIL) {
        if (i < n) {
            const bool f32 = tail.vf32[i] != 0;
            tail.vel[i] = kick_faithful(tail.vel[i], tail.h, bx, f32);
            tail.vel[i + n] = kick_faithful(tail.vel[i + n], tail.h, by, f32);
            tail.vel[i + 2 * n] = kick_faithful(tail.vel[i + 2 * n], tail.h, bz, f32);
            if (tail.hist_cap > 0 && ctl->overlap_count == 0) {            // a halting step appends on the host
                double* row = tail.hist + ((ctl->hist_count % tail.hist_cap) * n + i) * 3;
                row[0] = me.x; row[1] = me.y; row[2] = me.z;
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            if (atomicAdd(&ctl->rows_done, 1u) == gridDim.x - 1) {         // every CTA has read hist_count by now
                ctl->rows_done = 0;
                ctl->steps_done += 1;
                if (ctl->overlap_count > 0)
                    ctl->halted = 1;
                else if (tail.hist_cap > 0)
                    ctl->hist_count += 1;
            }
        }
    }
}

// Pass 2, four lanes per target (ORBITAL_B200_ROWS=4; kept as a measured alternative and cross-check).  The ordered sum is a chain of N dependent additions per target and
// component (8 cycles each: 17 us at N = 4,096); one warp doing everything for 32 targets has to issue 11 FP64 + 3 LDS
// per source and is issue-bound at ~36 cycles per source, and handing finished terms from producer warps to a consumer
// warp through shared memory (faithful_rows2_kernel) costs more in ring traffic and handshakes than it saves.  Here a
// warp owns 8 targets; the 4 lanes of a target finish the terms of 4 consecutive sources at once (8 FP64, one warp
// instruction each) and every lane then adds the 4 terms in source order, fetching them with shuffles -- the
// accumulator is replicated in the 4 lanes, bit for bit.  Per source: 2 + 3 FP64 and 3 SHFL warp instructions, no
// shared-memory stores, no cross-warp synchronisation beyond the TMA ring of matrix slabs the 4 warps of a CTA share.
// Measured: 186 us per step at N = 4,096 against 108 -- the 12 shuffles per 4 sources cost more than the ring they
// replace (SHFL issues at a fraction of the FP64 rate on this part), so this is not the default.
constexpr int kR4Stages = 4;
constexpr int kR4Smem = kR4Stages * (8192 + 1024) + 128;

template <bool FUSE_TAIL>
__global__ void __launch_bounds__(128) faithful_rows4_kernel(const double4* __restrict__ pos4,
                                                             const double* __restrict__ invr3, double* acc, long long n,
                                                             long long n_rows, double G, Ctl* ctl, const RowsTail tail) {
    if (ctl->halted) return;
    extern __shared__ __align__(128) unsigned char rows_smem[];
    double (*mt)[32 * 32] = reinterpret_cast<double (*)[32 * 32]>(rows_smem);                  // slabs of the matrix
    double4 (*praw)[32] = reinterpret_cast<double4 (*)[32]>(rows_smem + kR4Stages * 8192);     // their source bodies
    uint64_t* full = reinterpret_cast<uint64_t*>(rows_smem + kR4Stages * (8192 + 1024));
    uint64_t* empty = full + kR4Stages;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane & 3;                     // which of 4 consecutive sources this lane finishes
    const int col = warp * 8 + (lane >> 2);       // target column inside the 32-wide block
    const int grp = lane & ~3;
    const long long i = blockIdx.x * 32LL + col;
    const double4 me = pos4[min(i, n - 1)];
    const double* blk = invr3 + (long long)blockIdx.x * n_rows * 32;
    const int ntile = (int)((n + 31) / 32);       // rows past the last body are never needed
    auto issue = [&](int t) {
        const int st = t % kR4Stages;
        const uint32_t pbytes = (uint32_t)min(32LL, n - 32LL * t) * 32u;
        mbar_expect_tx(&full[st], 8192u + pbytes);
        tma_load_1d(mt[st], blk + (long long)t * 1024, 8192u, &full[st]);
        tma_load_1d(praw[st], pos4 + 32LL * t, pbytes, &full[st]);
    };
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kR4Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int t = 0; t < min(kR4Stages, ntile); ++t) issue(t);

    double bx = 0.0, by = 0.0, bz = 0.0;                        // physics.py:132 (replicated in the 4 lanes)
    for (int t = 0; t < ntile; ++t) {
        const int st = t % kR4Stages;
        if (threadIdx.x == 0 && t >= 1 && t - 1 + kR4Stages < ntile) {
            // refill the slab freed one tile ago: never waits on this tile's stragglers
            mbar_wait(&empty[(t - 1) % kR4Stages], (uint32_t)(((t - 1) / kR4Stages) & 1));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(t - 1 + kR4Stages);
        }
        mbar_wait(&full[st], (uint32_t)((t / kR4Stages) & 1));
        const int valid = (int)min(32LL, n - 32LL * t);
        const double* m = mt[st] + col;
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 4) {
            const int k = k0 + sub;
            double4 q = praw[st][k];
            if (k >= valid) q = make_double4(0.0, 0.0, 0.0, 0.0);               // past the end: a zero body, inv_r3 = 0
            const double sc = __dmul_rn(__dmul_rn(G, q.w), m[k * 32]);          // (G m_j) inv_r3     :151
            const double tx = __dmul_rn(sc, __dsub_rn(q.x, me.x));               // s * rij            :154
            const double ty = __dmul_rn(sc, __dsub_rn(q.y, me.y));
            const double tz = __dmul_rn(sc, __dsub_rn(q.z, me.z));
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                       // ascending source order
                bx = __dadd_rn(bx, __shfl_sync(0xffffffffu, tx, grp | j));
                by = __dadd_rn(by, __shfl_sync(0xffffffffu, ty, grp | j));
                bz = __dadd_rn(bz, __shfl_sync(0xffffffffu, tz, grp | j));
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
    const bool writer = sub == 0 && i < n;
    if (writer) {
        acc[i] = bx;
        acc[i + n] = by;
        acc[i + 2 * n] = bz;
    }
    if (FUSE_TAIL) {
        if (writer) {
            const bool f32 = tail.vf32[i] != 0;
            tail.vel[i] = kick_faithful(tail.vel[i], tail.h, bx, f32);
            tail.vel[i + n] = kick_faithful(tail.vel[i + n], tail.h, by, f32);
            tail.vel[i + 2 * n] = kick_faithful(tail.vel[i + 2 * n], tail.h, bz, f32);
            if (tail.hist_cap > 0 && ctl->overlap_count == 0) {            // a halting step appends on the host
                double* row = tail.hist + ((ctl->hist_count % tail.hist_cap) * n + i) * 3;
                row[0] = me.x; row[1] = me.y; row[2] = me.z;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&ctl->rows_done, 1u) == gridDim.x - 1) {         // every CTA has read hist_count by now
                ctl->rows_done = 0;
                ctl->steps_done += 1;
                if (ctl->overlap_count > 0)
                    ctl->halted = 1;
                else if (tail.hist_cap > 0)
                    ctl->hist_count += 1;
            }
        }
    }
}

bool faithful_pairs_applicable(long long n, bool sharded) {
    const char* env = getenv("ORBITAL_B200_FAITHFUL_PAIRS");     // "0": always the one-pass kernel (cross-check)
    if (env && env[0] == '0') return false;
    return !sharded && n > tiny_limit() && n <= 32768;               // 8 n^2 bytes of scratch: 8 GiB at n = 32768
}

// pass-2 kernel of the two-pass path: 2 = producer / consumer warps (default up to 148 column blocks), 1 = one warp
// per 32 targets (default above), 4 = four lanes per target (ORBITAL_B200_ROWS selects; read per call)
static int faithful_rows_variant(long long n) {
    const char* rows_env = getenv("ORBITAL_B200_ROWS");
    const int v = rows_env ? atoi(rows_env) : 2;
    if (v == 4 || v == 1) return v;
    return (n + 31) / 32 <= 148 ? 2 : 1;
}

const char* faithful_two_pass_name(long long n) {
    switch (faithful_rows_variant(n)) {
        case 4: return "faithful_pairs_kernel+faithful_rows4_kernel";
        case 2: return "faithful_pairs_kernel+faithful_rows2_kernel";
        default: return "faithful_pairs_kernel+faithful_rows_kernel";
    }
}

long long faithful_pairs_ld(long long n) { return (n + 31) / 32 * 32; }
long long faithful_pairs_elems(long long n) { return faithful_pairs_ld(n) * faithful_pairs_ld(n); }

// targets per warp: few, so that many warps are in flight and the fully unrolled term loop stays small
static int faithful_tw(long long n_tgt) {
    const char* env = getenv("ORBITAL_B200_FAITHFUL_TW");
    if (env) {
        const int v = atoi(env);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) return v;
    }
    // measured on B200 (faithful force pass, profiles/r1_sweep_faithful.txt), ms for TW = 1 / 2 / 4 / 8:
    //   n=1024: 0.042 0.045 0.063 0.090   n=2048: 0.087 0.084 0.111 0.167   n=4096: 0.256 0.184 0.249 0.316
    //   n=8192: 0.937 0.562 0.527 0.710   n=16384: 3.65 2.10 1.80 1.54 (FP64-throughput bound)
    if (n_tgt <= 1536) return 1;
    if (n_tgt <= 6144) return 2;
    return n_tgt <= 12288 ? 4 : 8;
}

void faithful_geometry(long long n_tgt, int* grid, int* block) {
    const int tw = faithful_tw(n_tgt);
    *block = 32 * kFaithWarps;
    *grid = (int)((n_tgt + (long long)tw * kFaithWarps - 1) / ((long long)tw * kFaithWarps));
}

template <int TW>
static void launch_faithful_t(const DeviceState& s, const StepParams& p, bool detect, int grid, cudaStream_t st,
                              bool fuse_tail) {
    const int smem = kFaithWarps * (3 * TW * 33 * 8 + TW * 32);
    static DeviceOnce attr_set;                                 // the attribute is per device
    if (attr_set.first()) {
        cudaFuncSetAttribute(force_faithful_kernel<TW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(force_faithful_kernel<TW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    if (s.invr3) {
        // two passes: pair matrix (each pair's sqrt/div once, overlap test included), then ordered row sums
        const int nb = (int)((s.n + 31) / 32);
        const unsigned tri = (unsigned)((long long)nb * (nb + 1) / 2);     // tiles on or above the diagonal
        if (detect)
            faithful_pairs_kernel<true><<<tri, dim3(32, 8), 0, st>>>(s.pos4, s.radius, s.invr3, s.n, s.invr3_ld,
                                                                              p.eps2, s.ctl, s.pairs);
        else
            faithful_pairs_kernel<false><<<tri, dim3(32, 8), 0, st>>>(s.pos4, s.radius, s.invr3, s.n,
                                                                               s.invr3_ld, p.eps2, s.ctl, s.pairs);
        static DeviceOnce rows_attr;
        if (rows_attr.first()) {
            cudaFuncSetAttribute(faithful_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmem);
            cudaFuncSetAttribute(faithful_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmem);
        }
        const RowsTail tail = {s.vel, s.vf32, s.hist, s.hist_cap, p.h};
        // pass-2 kernel: 2 = producer / consumer warps (default up to 148 column blocks), 1 = one warp per 32 targets
        // (default above), 4 = four lanes per target.  Measured whole steps, us (profiles/r2_sweep_faithful_rows*.txt):
        //   N      rows2   rows1   rows4          N       rows2*  rows1   rows4      (* falls back to rows1 above 148 blocks)
        //   1,024   27.1    29.9    46.1          8,192    297.9   297.6   524.8
        //   4,096  107.5   119.0   186.5         16,384    900.6   899.7  1984.7
        const int rows_variant = faithful_rows_variant(s.n);
        if (rows_variant == 4) {
            static DeviceOnce rows4_attr;
            if (rows4_attr.first()) {
                cudaFuncSetAttribute(faithful_rows4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kR4Smem);
                cudaFuncSetAttribute(faithful_rows4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kR4Smem);
            }
            if (fuse_tail)
                faithful_rows4_kernel<true><<<nb, 128, kR4Smem, st>>>(s.pos4, s.invr3, s.acc, s.n, s.invr3_ld, p.G, s.ctl,
                                                                      tail);
            else
                faithful_rows4_kernel<false><<<nb, 128, kR4Smem, st>>>(s.pos4, s.invr3, s.acc, s.n, s.invr3_ld, p.G, s.ctl,
                                                                       tail);
            return;
        }
        if (rows_variant == 2) {
            static DeviceOnce rows2_attr;
            if (rows2_attr.first()) {
                cudaFuncSetAttribute(faithful_rows2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kR2Smem);
                cudaFuncSetAttribute(faithful_rows2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kR2Smem);
            }
            if (fuse_tail)
                faithful_rows2_kernel<true><<<nb, 32 * (1 + kR2Prod), kR2Smem, st>>>(s.pos4, s.invr3, s.acc, s.n,
                                                                                   s.invr3_ld, p.G, s.ctl, tail);
            else
                faithful_rows2_kernel<false><<<nb, 32 * (1 + kR2Prod), kR2Smem, st>>>(s.pos4, s.invr3, s.acc, s.n,
                                                                                    s.invr3_ld, p.G, s.ctl, tail);
            return;
        }
        if (fuse_tail)
            faithful_rows_kernel<true><<<nb, 32, kRowSmem, st>>>(s.pos4, s.invr3, s.acc, s.n, s.invr3_ld, p.G, s.ctl,
                                                                tail);
        else
            faithful_rows_kernel<false><<<nb, 32, kRowSmem, st>>>(s.pos4, s.invr3, s.acc, s.n, s.invr3_ld, p.G, s.ctl,
                                                                 tail);
        return;
    }
    if (detect)
        force_faithful_kernel<TW, true><<<grid, 32 * kFaithWarps, smem, st>>>(s.pos4, s.radius, s.acc, s.n, s.tgt_lo,
                                                                            s.tgt_hi, p.eps2, p.G, s.ctl, s.pairs);
    else
        force_faithful_kernel<TW, false><<<grid, 32 * kFaithWarps, smem, st>>>(s.pos4, s.radius, s.acc, s.n, s.tgt_lo,
                                                                             s.tgt_hi, p.eps2, p.G, s.ctl, s.pairs);
}

cudaError_t launch_force_faithful(const DeviceState& s, const StepParams& p, bool detect, cudaStream_t st,
                                  int* launches, bool fuse_tail) {
    if (fuse_tail && !s.invr3) return cudaErrorInvalidValue;      // only the two-pass path carries the step tail
    int grid, block;
    faithful_geometry(s.tgt_hi - s.tgt_lo, &grid, &block);
    switch (faithful_tw(s.tgt_hi - s.tgt_lo)) {
        case 1: launch_faithful_t<1>(s, p, detect, grid, st, fuse_tail); break;
        case 2: launch_faithful_t<2>(s, p, detect, grid, st, fuse_tail); break;
        case 4: launch_faithful_t<4>(s, p, detect, grid, st, fuse_tail); break;
        case 8: launch_faithful_t<8>(s, p, detect, grid, st, fuse_tail); break;
        case 16: launch_faithful_t<16>(s, p, detect, grid, st, fuse_tail); break;
        default: launch_faithful_t<32>(s, p, detect, grid, st, fuse_tail); break;
    }
    if (launches) *launches += s.invr3 ? 2 : 1;
    return cudaGetLastError();
}

}  // namespace orb
