// kernels.h -- host-visible launch interface between the translation units of liborbital_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "common.cuh"

namespace orb {

// Resident state of one engine handle (device pointers).
struct DeviceState {
    long long n = 0;          // all bodies (sources)
    long long tgt_lo = 0;     // locally integrated targets [tgt_lo, tgt_hi)
    long long tgt_hi = 0;
    double4* pos4 = nullptr;  // n x {x,y,z,m}
    double* vel = nullptr;    // 3 x n (SoA: vx | vy | vz)
    double* acc = nullptr;    // 3 x n
    double* radius = nullptr; // n
    uint8_t* vf32 = nullptr;  // n : velocity stored as float32 in the reference
    Ctl* ctl = nullptr;
    long long* pairs = nullptr;     // 2 x pairs_cap (Ctl::pairs_cap)
    int pairs_cap = 0;
    double* hist = nullptr;         // hist_cap x n x 3
    long long hist_cap = 0;
    double* scratch = nullptr;      // fast kernel partial sums: slabs x 3 x n_tgt
    long long scratch_elems = 0;
    double* reduce_buf = nullptr;   // diagnostics partials
    double* invr3 = nullptr;        // faithful two-pass path: n x invr3_ld matrix of 1/r^3 (both triangles)
    long long invr3_ld = 0;
};

struct StepParams {
    double dt, h, eps2, G;
    float dt32;
    double rmax1, rmax2;      // largest / second-largest radius (fast-mode overlap prefilter)
    long long rmax1_idx;
    int detect;               // any radius > 0 (or coincident check wanted)
    double restitution;       // collide_spheres restitution (core/engine.py:85)
    int device_contacts;      // 1: contacts are resolved on the device, the step never halts
    double uniform_mass;      // the common mass when every body has the same (non-zero) mass, else 0
};

// "Has this been done on the current device yet?" -- cudaFuncSetAttribute is per device, and one process may
// drive engines on several GPUs.
struct DeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        return (mask.fetch_or(bit) & bit) == 0;
    }
};

// Geometry chosen for the fast force kernel.
struct FastPlan {
    int ti = 1;               // targets per thread
    int slabs = 1;            // source slabs (2-D decomposition)
    int tiles_per_slab = 0;
    int grid = 0;
    int block = 128;
    int smem = 0;
    int ctas_per_sm = 0;
};

// ---- force.cu ----
FastPlan plan_fast(long long n_tgt, long long n_src, int sm_count);
cudaError_t launch_force_fast(const DeviceState& s, const StepParams& p, const FastPlan& plan, bool detect,
                              cudaStream_t st, int* launches);
// fuse_tail (two-pass path only, s.invr3 != nullptr): second half-kick, history append and step bookkeeping ride
// along in the force pass (replaces launch_kick_hist + launch_advance)
cudaError_t launch_force_faithful(const DeviceState& s, const StepParams& p, bool detect, cudaStream_t st,
                                  int* launches, bool fuse_tail = false);
void faithful_geometry(long long n_tgt, int* grid, int* block);
// two-pass bit-exact force (pair matrix of 1/r^3, then ordered row sums): sizes it is used for, leading dimension
bool faithful_pairs_applicable(long long n, bool sharded);
long long faithful_pairs_ld(long long n);
const char* faithful_two_pass_name(long long n);     // which pass-2 kernel the two-pass path runs at this size
long long faithful_pairs_elems(long long n);
const char* fast_kernel_name(int ti, bool detect);

// ---- integrate.cu ----
cudaError_t launch_kick_drift(const DeviceState& s, const StepParams& p, cudaStream_t st);
cudaError_t launch_kick_hist(const DeviceState& s, const StepParams& p, cudaStream_t st);
cudaError_t launch_advance(const DeviceState& s, cudaStream_t st);
// device-side contact handling (physics.py:391-422,510-535): stash U, resolve in reference order, append history
cudaError_t launch_contacts(const DeviceState& s, const StepParams& p, bool ordered_potential, cudaStream_t st,
                            int* launches);
cudaError_t launch_hist_append(const DeviceState& s, cudaStream_t st);
// sharded step tail: replicated contact sweep (resolve), full-snapshot history append, bookkeeping
cudaError_t launch_step_end(const DeviceState& s, const StepParams& p, bool resolve, bool ordered_potential,
                            cudaStream_t st, int* launches);
cudaError_t launch_set_overlaps(const DeviceState& s, int count, int overflow, cudaStream_t st);
// single-CTA fused multi-step kernel (faithful arithmetic), n <= kTinyMax
constexpr int kTinyMax = 512;        // capacity of the fused kernels
int tiny_limit();                    // sizes that actually take them (<= kTinyMax; ORBITAL_B200_TINY_MAX overrides)
int tiny_block(int n);
cudaError_t launch_tiny_steps(const DeviceState& s, const StepParams& p, long long nsteps, bool detect,
                              cudaStream_t st);
cudaError_t launch_pack(double4* pos4, const double* x, const double* y, const double* z, const double* m,
                        long long n, cudaStream_t st);
cudaError_t launch_unpack(const double4* pos4, double* x, double* y, double* z, long long n, cudaStream_t st);
cudaError_t launch_potential(const DeviceState& s, const StepParams& p, bool faithful_order, double* d_out,
                             cudaStream_t st, int* launches);
// acc[slab] = sum over ranks of peer_acc[r][slab], rank order (orb_peer_reduce)
cudaError_t launch_peer_reduce(const double* const* peer_acc, int world, double* acc, long long n, long long lo,
                               long long hi, cudaStream_t st);
// potential term of Object.lagrangian (physics.py:275-279) for body i, reference order, unsoftened
cudaError_t launch_body_potential(const DeviceState& s, long long i, double G, double* d_out, cudaStream_t st);
cudaError_t launch_energy_angmom(const DeviceState& s, double* d_out4, cudaStream_t st, int* launches);

// ---- peak.cu ----
cudaError_t run_fp64_peak(int device, double seconds, double* tflops_best, double* tflops_mean, double* mhz);

}  // namespace orb
