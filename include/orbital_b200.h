/*
 * orbital_b200.h -- C ABI of liborbital_b200.so (sm_100a).
 *
 * The reference (trevormcguire/orbital-physics) is pure Python and has no FFI:
 * its only seam for the hot path is the Python import surface of
 * core.engine / core.physics.  This header is the boundary a
 * drop-in replacement binds instead: plain pointers and sizes, no torch or
 * NumPy types.  orbital-physics_b200/core/_native.py is the ctypes binding;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Each entry point names the reference code it replaces (file:line relative to
 * the reference repository root).
 *
 * Conventions
 *   - every function returns 0 (ORB_OK) or an ORB_ERR_* code; the message for
 *     the calling thread's last failure is orb_last_error().
 *   - host arrays are structure-of-arrays fp64, one entry per body.
 *   - a handle is internally serialised by a mutex: one writer thread calling
 *     orb_step and any number of reader threads calling orb_download_* is the
 *     reference app's threading contract (app/app.py:104-115).
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 *     with ORB_ERR_NO_DEVICE.
 */
#ifndef ORBITAL_B200_H
#define ORBITAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORB_ABI_VERSION 1

enum orb_status {
    ORB_OK = 0,
    ORB_ERR_INVALID = 1,     /* bad argument / state                         */
    ORB_ERR_CUDA = 2,        /* a CUDA runtime call failed                   */
    ORB_ERR_NO_DEVICE = 3,   /* no usable CUDA device (no CPU fallback)      */
    ORB_ERR_OOM = 4
};

enum orb_mode {
    /* Bit-faithful: the reference's exact rounding sequence,
     * ascending-j accumulation, IEEE sqrt/div. Results equal the reference bit for bit. */
    ORB_MODE_FAITHFUL = 0,
    /* Roofline kernel: TMA-staged source tiles, rsqrt seed + fp64 polynomial
     * refinement, register accumulators. Relative acceleration error <= 1e-12. */
    ORB_MODE_FAST = 1
};

typedef struct orb_engine orb_engine;       /* one N-body system            */
typedef struct orb_ensemble orb_ensemble;   /* many independent small ones  */

/* ---- library / device ------------------------------------------------- */
int orb_abi_version(void);
const char* orb_last_error(void);
int orb_device_count(int* count);
int orb_device_info(int device, char* name, int name_len, int* sm_count,
                    int* cc_major, int* cc_minor, int64_t* total_mem_bytes);
/* Pinned host memory for the upload/download staging of large systems. */
int orb_host_alloc(void** ptr, int64_t bytes);
int orb_host_free(void* ptr);
/* Measured FP64 DFMA-chain throughput of `device` (TFLOP/s, 2 flop per DFMA):
 * the roofline denominator for the force kernel. Runs back-to-back launches for
 * about `seconds`; best = fastest launch (burst), mean = average over the second
 * half of the run (sustained, under the power cap); clock from clock64(). */
int orb_fp64_peak(int device, double seconds, double* tflops_best, double* tflops_mean,
                  double* sm_clock_mhz);

/* ---- engine lifecycle --------------------------------------------------
 * Replaces the state held by SimulationEngine / ObjectCollection / Object
 * (core/engine.py:19-46, core/physics.py:161-191,452-508).
 * orb_create:         all n bodies are integrated on `device`.
 * orb_create_sharded: bodies [tgt_lo, tgt_hi) are integrated here, all n act as
 *                     sources (multi-GPU target partition). Equal
 *                     slabs imply (rank, world) = (tgt_lo / slab, n / slab).
 * orb_create_ranked:  the same with an explicit rank / world, for slabs of
 *                     ceil(n / world) bodies whose last one is shorter. The packed
 *                     position buffer (orb_pos4_ptr) then holds world * ceil(n / world)
 *                     entries so that equal-size in-place all-gathers fit; entries
 *                     >= n are padding no kernel reads.                          */
int orb_create(orb_engine** out, int64_t n, int device, int mode);
int orb_create_sharded(orb_engine** out, int64_t n, int64_t tgt_lo, int64_t tgt_hi,
                       int device, int mode);
int orb_create_ranked(orb_engine** out, int64_t n, int64_t tgt_lo, int64_t tgt_hi,
                      int rank, int world, int device, int mode);
int orb_destroy(orb_engine* e);

/* SimulationEngine(dt=, softening=) + STANDARD.G (core/engine.py:31-32, core/constants.py:49-58). */
int orb_set_params(orb_engine* e, double dt, double eps, double G);
int orb_set_mode(orb_engine* e, int mode);
/* Contact handling (core/engine.py:85 -> core/physics.py:510-535,391-422). restitution: collide_spheres'
 * coefficient. resolve_on_device != 0: overlapping pairs are resolved on the device with the reference's
 * sequential in-place semantics (lexicographic pair order, re-tested against current positions, float32
 * velocity rounding), orb_step never halts and *n_overlaps reports the touching pairs resolved.
 * resolve_on_device == 0 (default): orb_step halts and the caller resolves (see orb_step). */
int orb_set_contacts(orb_engine* e, double restitution, int resolve_on_device);
/* Device ring of the last `capacity` position snapshots (engine.history, core/engine.py:34,88-92).
 * 0 disables recording. Resets the ring. */
int orb_set_history(orb_engine* e, int64_t capacity);
/* Run on a caller-owned CUDA stream (cudaStream_t), e.g. torch's current stream. NULL = the handle's
 * own stream; pass cudaStreamLegacy (0x1) for the legacy default stream. */
int orb_set_stream(orb_engine* e, void* cuda_stream);

/* ---- state transfer ----------------------------------------------------
 * Upload replaces positions, velocities, masses, radii (Object attributes,
 * core/physics.py:181-184). vel_is_f32[i] != 0 marks a body whose velocity is a
 * float32 array in the reference (core/physics.py:184): its velocity is rounded
 * to float32 after every update and the drift product is formed in float32
 * (SURVEY.md A.2). NULL = all fp64. Accelerations are NOT touched (the
 * reference keeps a stale self.acc after external mutation, core/engine.py:85). */
int orb_upload(orb_engine* e, const double* x, const double* y, const double* z,
               const double* vx, const double* vy, const double* vz,
               const double* m, const double* radius, const uint8_t* vel_is_f32);
int orb_download_state(orb_engine* e, double* x, double* y, double* z,
                       double* vx, double* vy, double* vz);
int orb_download_acc(orb_engine* e, double* ax, double* ay, double* az);   /* engine.acc */
int orb_upload_acc(orb_engine* e, const double* ax, const double* ay, const double* az);

/* ---- the hot path ------------------------------------------------------ */
/* pairwise_accelerations (core/physics.py:125-159) on the resident positions;
 * fills the resident acceleration arrays (engine.py:41). Asynchronous. */
int orb_accel(orb_engine* e);
/* SimulationEngine.step x nsteps (core/engine.py:65-97): half-kick, drift,
 * force, half-kick, overlap detection, history append -- one launch sequence
 * per step under a CUDA graph (one fused single-CTA launch for small n).
 * Collision *detection* (core/physics.py:517-518) runs on the device; if any
 * pair overlaps in a step the device halts after that step (before the history
 * append), *steps_done < nsteps and *n_overlaps > 0: the caller resolves the
 * contacts with the reference's sequential semantics (core/physics.py:391-422),
 * re-uploads, calls orb_history_append and resumes. Synchronous on return. */
int orb_step(orb_engine* e, int64_t nsteps, int64_t* steps_done, int64_t* n_overlaps);
/* Overlapping pairs (i<j) of the halted step (or, on a sharded handle, of the last orb_step_force), unsorted;
 * *count may exceed the pairs returned (then the list overflowed and the caller must sweep all pairs). */
int orb_overlap_pairs(orb_engine* e, int64_t* pairs_ij, int64_t cap, int64_t* count);
/* Split step for sharded engines (one handle per rank; core/engine.py:65-97 per step):
 *   orb_step_begin   half-kick + drift of the local targets (engine.py:69-75)
 *   [caller]         all-gather the orb_pos4_ptr() slabs across ranks
 *   orb_step_force   force pass (engine.py:78) with the overlap test of engine.py:85 ->
 *                    physics.py:517-518 fused in: fills this rank's pair list
 *   [caller]         all-reduce orb_acc_ptr() if orb_acc_needs_allreduce()
 *   orb_step_kick    second half-kick of the local targets (engine.py:81-82)
 *   [caller]         only when some rank reports pairs (orb_overlap_count): all-gather the
 *                    velocities and the pair lists, orb_set_overlap_pairs(merged list)
 *   orb_step_end     contact sweep (physics.py:510-535, 391-422) replicated on every rank over
 *                    the identical full state -- sequential, lexicographic, bit-identical
 *                    everywhere, each rank keeps its slab -- then the history append of all n
 *                    bodies (engine.py:88-92) and the step bookkeeping.
 * orb_step_finish = orb_step_force + orb_step_kick for engines whose accelerations are complete.
 * On an UNSHARDED handle orb_step_begin / orb_accel / orb_step_kick is a whole step (orb_step_kick then
 * also appends the history point and advances), which lets a caller bracket the force pass with its own
 * CUDA events. All asynchronous except orb_overlap_count / orb_set_overlap_pairs. */
int orb_step_begin(orb_engine* e);
int orb_step_force(orb_engine* e);
int orb_step_finish(orb_engine* e);
int orb_step_kick(orb_engine* e);
int orb_step_end(orb_engine* e);
/* Pairs this handle's last force pass flagged (clamped to the list capacity); *overflowed != 0 if some were
 * dropped. Synchronous. */
int orb_overlap_count(orb_engine* e, int64_t* count, int* overflowed);
/* Replace the handle's pair list (e.g. with the union over ranks). overflowed != 0: the list is incomplete, the
 * sweep of orb_step_end scans all pairs instead (exact, slower). */
int orb_set_overlap_pairs(orb_engine* e, const int64_t* pairs_ij, int64_t count, int overflowed);
/* Touching pairs resolved on the device since the last orb_step began / in this handle's life (sharded), and how
 * many sweeps had to abandon the pair list (more than ORBITAL_B200_OVERLAP_CAP pairs, default 2^20) and scan all
 * pairs. */
int orb_contact_stats(orb_engine* e, int64_t* contacts_total, int64_t* full_sweeps);
/* *flag != 0: on this (sharded, fast-mode) engine orb_accel leaves a PARTIAL acceleration of all n bodies
 * (the pair-symmetric kernel evaluates every world-th pair block per rank); the caller must
 * all-reduce (sum) the 3 x n buffer at orb_acc_ptr across ranks before orb_step_kick. orb_step_finish is
 * not available then. */
int orb_acc_needs_allreduce(orb_engine* e, int* flag);
int orb_synchronize(orb_engine* e);

/* ---- device-resident views (for torch.distributed / CUDA-event timing) -- */
/* Packed sources: n x {x,y,z,m} fp64 (32 bytes per body). */
int orb_pos4_ptr(orb_engine* e, void** device_ptr, int64_t* n_bodies);
int orb_vel_ptr(orb_engine* e, void** device_ptr);   /* 3 x n fp64, SoA */
int orb_acc_ptr(orb_engine* e, void** device_ptr);   /* 3 x n fp64, SoA */
/* Peer-memory reduction of the partial accelerations (one process per GPU on one NVLink / NVSwitch node):
 * instead of all-reducing orb_acc_ptr with NCCL, every rank exports its buffer (CUDA IPC handle + offset,
 * orb_peer_export), opens every other rank's (orb_peer_open) and, after each force pass, sums the columns
 * of ITS OWN slab straight out of peer memory in fixed rank order (orb_peer_reduce, asynchronous) -- the
 * reduce-scatter the second half-kick needs, in one kernel over NVLink. The caller synchronises the ranks:
 * all force passes are complete before any orb_peer_reduce (e.g. a 1-element all-reduce on the same
 * stream), and no rank starts its next force pass before all have reduced (the position all-gather of the
 * next step does that). Replaces the all-reduce of the reference-side multi-GPU plan. */
int orb_peer_export(orb_engine* e, void* handle64, int64_t* offset);
int orb_peer_open(orb_engine* e, int rank, const void* handle64, int64_t offset);
int orb_peer_reduce(orb_engine* e);
/* Unmap the other ranks' buffers (synchronises this handle's stream first). An exporting rank must not
 * destroy its engine before every importer has closed: close on all ranks, barrier, then orb_destroy. */
int orb_peer_close(orb_engine* e);
/* Name and launch geometry of the force kernel the current mode/size selects. */
int orb_force_kernel_info(orb_engine* e, char* name, int name_len, int* grid, int* block,
                          int* smem_bytes, int* launches_per_step);
/* Total kernel launches issued by this handle so far (bench.py gpu_launches). */
int orb_launch_count(orb_engine* e, int64_t* launches);

/* ---- diagnostics ------------------------------------------------------- */
/* U = -sum_{i<j} G m_i m_j / sqrt(r^2+eps^2)  (core/physics.py:158; last_potential).
 * Faithful mode and n <= 4096: lexicographic sequential order (bit-exact). */
int orb_potential(orb_engine* e, double* U);
/* Potential term of Object.lagrangian (core/physics.py:275-279) for one body:
 * pe = sum_{j != body, ascending j} ((-G m_body) m_j) / ||r_body - r_j||, unsoftened, added in the
 * reference's loop order (bit-identical to the Python loop; a coincident body gives -inf as there).
 * G is the caller's (Object.unit_profile.G). The positions are the resident ones (all n bodies; on a
 * sharded engine every rank holds them after a step). Synchronous. */
int orb_body_potential(orb_engine* e, int64_t body, double G, double* pe);
/* K = sum 1/2 m v.v, L = sum r x (m v)  (core/engine.py:104-121), fp64 tree reduction. */
int orb_energy_angmom(orb_engine* e, double* K, double* L3);

/* ---- history ring (engine.history) -------------------------------------- */
int orb_history_count(orb_engine* e, int64_t* total_appended);
/* Append the resident positions now (ctor seed, engine.py:34; post-collision append). */
int orb_history_append(orb_engine* e);
/* The most recent min(last_k, stored) snapshots, oldest first: out[k][body][3]. */
int orb_history_download(orb_engine* e, int64_t last_k, double* out, int64_t* got);

/* ---- batched ensemble: nsys independent systems of nbody bodies ---------
 * BASELINE config C3. Bit-exact mode: one warp per system; fast mode: nbody/2 lanes
 * per system, 64/nbody systems per warp (small batches: one body per lane, nbody lanes
 * per system -- twice the warps). Arrays are [nsys][nbody] fp64.
 * Equivalent to nsys separate SimulationEngine instances stepped in lockstep
 * (core/engine.py:19-46,65-97) without collision handling. */
int orb_ens_create(orb_ensemble** out, int64_t nsys, int nbody, int device, int mode, int vel_f32);
int orb_ens_destroy(orb_ensemble* s);
int orb_ens_set_params(orb_ensemble* s, double dt, double eps, double G);
int orb_ens_set_stream(orb_ensemble* s, void* cuda_stream);
/* Optional per-body attributes, [nsys][nbody] each, NULL = none (call before orb_ens_upload):
 *   radius      -> every step ends with the reference's contact sweep per system (core/engine.py:85 ->
 *                  core/physics.py:510-535, 391-422: sequential, lexicographic, in place);
 *   vel_is_f32  -> per-body "velocity is a float32 array" flags (core/physics.py:184 vs :448-449), overriding
 *                  the vel_f32 argument of orb_ens_create. */
int orb_ens_set_bodies(orb_ensemble* s, const double* radius, const uint8_t* vel_is_f32);
int orb_ens_set_contacts(orb_ensemble* s, double restitution);
/* Touching pairs resolved since creation (all systems). Synchronous. */
int orb_ens_contact_count(orb_ensemble* s, int64_t* contacts);
int orb_ens_upload(orb_ensemble* s, const double* x, const double* y, const double* z,
                   const double* vx, const double* vy, const double* vz, const double* m);
/* Generate the ensemble's initial condition on the device from orbital elements
 * (replaces nsys*(nbody-1) host calls of Body.get_state, core/body.py:184-249, and
 * solve_kepler, core/physics.py:43-71). Body 0 of every system is the central
 * mass at rest at the origin; bodies 1.. are parent-relative with
 * n = sqrt(G m_0 / a^3) (core/body.py:159-169) and b = a sqrt(1-e^2) (:120-124).
 * Element arrays are [nsys][nbody-1] (radians, metres); m is [nsys][nbody].
 * Same operation order as the reference; sin/cos as orb_set_trig_mode selects
 * (default ORB_TRIG_LIBM: bit-identical to the reference on an x86-64 glibc host). */
int orb_ens_upload_elements(orb_ensemble* s, const double* M, const double* e, const double* a,
                            const double* inc, const double* Omega, const double* omega,
                            const double* m);
/* fused != 0: all nsteps inside one launch, state in registers/shared memory
 * (FP64-bound); fused == 0: one launch per step, state round-trips HBM (HBM-bound).
 * Between the launches of one call only x and the half-kicked velocity travel (read x,u,m +
 * write x,u = 104 B per body-step); the first launch also reads and the last also
 * writes the accelerations the reference keeps between steps (engine.py:41,78), so the state is
 * synchronised (x, v, a) whenever the call returns. Same rounding sequence in both forms.
 * Asynchronous. */
int orb_ens_step(orb_ensemble* s, int64_t nsteps, int fused);
int orb_ens_download(orb_ensemble* s, double* x, double* y, double* z,
                     double* vx, double* vy, double* vz);
int orb_ens_download_acc(orb_ensemble* s, double* ax, double* ay, double* az);   /* each system's engine.acc */
int orb_ens_energy(orb_ensemble* s, double* E_per_system);
int orb_ens_synchronize(orb_ensemble* s);
int orb_ens_launch_count(orb_ensemble* s, int64_t* launches);

/* ---- initial-condition pipeline ----------------------------
 * Batched Keplerian elements -> parent-relative Cartesian state: count bodies,
 * host arrays in (M, inc, Omega, omega in radians; a, b in metres; n = mean
 * motion in rad/s), host arrays out (r3, v3: [3][count]; E optional: eccentric
 * anomaly). Replaces a Python loop over Body.get_state (core/body.py:184-249)
 * / solve_kepler(M, e, tol, max_iter) (core/physics.py:43-71).
 *
 * The reference's math.sin / math.cos (core/body.py:218-223, core/physics.py:63)
 * are the host libm's; orb_set_trig_mode picks what stands in for them on the
 * device, process-wide (start value: env ORBITAL_B200_TRIG = libm | cr | fast):
 *   ORB_TRIG_LIBM  glibc 2.39 sin/cos restated operation by operation
 *                  (csrc/sincos_libm.h): states bit-identical to the reference
 *                  on an x86-64 glibc host with FMA. Default.
 *   ORB_TRIG_CR    correctly rounded sin/cos (csrc/sincos_cr.h): host independent.
 *   ORB_TRIG_FAST  CUDA sincos (<= 2 ulp): tolerance path.
 * The two integer powers of the reference's IC code, e ** 2 (core/body.py:216)
 * and a ** 3 (core/body.py:166), are CPython float pow = the host libm's pow(),
 * which is not correctly rounded either; the library evaluates them with the
 * same host pow() (one call per body, before the launch). */
#define ORB_TRIG_LIBM 0
#define ORB_TRIG_CR 1
#define ORB_TRIG_FAST 2
int orb_set_trig_mode(int mode);
int orb_get_trig_mode(void);
int orb_kepler_states(int device, int64_t count, const double* M, const double* e, const double* a,
                      const double* b, const double* n, const double* inc, const double* Omega,
                      const double* omega, double tol, int max_iter,
                      double* r3, double* v3, double* E);

#ifdef __cplusplus
}
#endif
#endif /* ORBITAL_B200_H */
