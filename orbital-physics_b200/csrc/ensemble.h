// ensemble.h -- launch interface of ensemble.cu
#pragma once
#include <cuda_runtime.h>
namespace orb {
struct EnsArgs {
    double *x, *y, *z, *vx, *vy, *vz, *ax, *ay, *az;
    const double* m;
    long long nsys;
    int nb, nbp;
    long long nsteps;
    double h, dt, eps2, G;
    float dt32;
    int vel_f32;           // every velocity is float32 (ignored when vf32 is set)
    int warps_per_cta;     // warps per CTA
    const double* radius;           // [nsys][nb] or nullptr: no contact handling
    const unsigned char* vf32;      // [nsys][nb] per-body "velocity is float32" flags, or nullptr
    double restitution;             // collide_spheres' coefficient (core/engine.py:85)
    unsigned long long* contacts;   // device counter: touching pairs resolved (or nullptr)
    int pdl;               // host only: launch with programmatic stream serialization (overlaps the predecessor's tail)
    int first, last;       // see ensemble.cu: does the launch start from / end with the synchronised (x, v, a) state
    int narrow;            // fast mode without contacts: one body per lane (ens_fast_body1) instead of two
};
cudaError_t launch_ens_step(const EnsArgs& a, bool faithful, cudaStream_t st);
// fast mode, all a.nsteps in one launch, balanced over the SMs in slices of `slice` steps (small batches)
cudaError_t launch_ens_step_sliced(const EnsArgs& a, int slice, unsigned long long* queue, int* progress, int sm_count,
                                   cudaStream_t st);
cudaError_t launch_ens_accel(const EnsArgs& a, bool faithful, cudaStream_t st);
cudaError_t launch_ens_energy(const EnsArgs& a, double* E, cudaStream_t st);
// kepler.cu: Keplerian elements -> Cartesian state (core/physics.py:43-71, core/body.py:184-249)
// trig: ORB_TRIG_* (include/orbital_b200.h) -- which sin / cos stands in for the reference's math.sin / math.cos
cudaError_t launch_kepler_states(const double* d_el9, double* d_out7, long long count, double tol, int max_iter,
                                 int trig, cudaStream_t st);
cudaError_t launch_ens_elements(const EnsArgs& a, const double* d_el8, double tol, int max_iter, int trig,
                                cudaStream_t st);
}  // namespace orb
