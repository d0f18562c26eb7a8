// api.cu -- C ABI of liborbital_b200.so (include/orbital_b200.h): handle
// lifetime, host<->device staging, step orchestration (CUDA graph / fused
// single-CTA kernel), history ring, diagnostics.  No CPU compute path exists:
// every entry point that computes needs a CUDA device.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/orbital_b200.h"
#include "ensemble.h"
#include "force_sym.h"
#include "kernels.h"

// which sin / cos the initial-condition kernels use (kepler.cu); process-wide, ORBITAL_B200_TRIG sets the start value
static std::atomic<int> g_trig{-1};
static int trig_mode() {
    int t = g_trig.load(std::memory_order_relaxed);
    if (t < 0) {
        const char* env = getenv("ORBITAL_B200_TRIG");
        t = ORB_TRIG_LIBM;
        if (env && !strcmp(env, "cr")) t = ORB_TRIG_CR;
        if (env && !strcmp(env, "fast")) t = ORB_TRIG_FAST;
        g_trig.store(t, std::memory_order_relaxed);
    }
    return t;
}

// x ** y as CPython evaluates it: the host libm's pow().  The reference's IC code has two integer powers,
// e ** 2 (core/body.py:216) and a ** 3 (core/body.py:166); glibc's pow is < 1 ulp but not correctly rounded
// (0.08 % of squares differ from e * e), and unlike sin / cos its table-driven algorithm is not restated for
// the device, so these two scalars per body are formed here, on the host, and travel with the elements.
static void host_pow_plane(const double* x, double y, double* out, int64_t n) {
    const int nt = n >= (1 << 16) ? (int)std::min<int64_t>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
    auto work = [=](int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; ++i) out[i] = pow(x[i], y); };
    if (nt == 1) return work(0, n);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, n * t / nt, n * (t + 1) / nt);
    for (auto& t : th) t.join();
}

using namespace orb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    const int code = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? ORB_ERR_NO_DEVICE
                     : (e == cudaErrorMemoryAllocation ? ORB_ERR_OOM : ORB_ERR_CUDA);
    cudaGetLastError();   // clear sticky-free errors
    return fail(code, buf);
}

#define CU(call)                                              \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);   \
    } while (0)

__global__ void ctl_reset_kernel(Ctl* ctl, int reset_hist, int clear_u) {
    ctl->steps_done = 0;
    ctl->halted = 0;
    ctl->overlap_count = 0;
    ctl->overlap_overflow = 0;
    ctl->contacts_total = 0;
    if (clear_u) ctl->u_valid = 0;
    if (reset_hist) ctl->hist_count = 0;
}

}  // namespace

struct orb_engine {
    std::mutex mu;
    int device = 0;
    int mode = ORB_MODE_FAITHFUL;
    int sm_count = 148;
    bool sharded = false;
    int rank = 0, world = 1;         // pair-block ownership of the pair-symmetric kernel (world 0: not available)
    long long pos4_cap = 0;          // entries allocated for pos4 (>= n: padded so equal-size all-gathers fit)
    DeviceState s;
    StepParams p{};
    FastPlan plan;
    bool plan_valid = false;
    SymPlan sym;                     // pair-symmetric kernel (unsharded fast mode)
    bool use_sym = true;
    bool pairs_unavailable = false;  // the two-pass faithful force could not allocate its pair matrix
    bool detect = false;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    double* d_stage = nullptr;       // 4 x n staging for SoA <-> packed conversion
    double* d_diag = nullptr;        // 8 doubles
    double* h_diag = nullptr;        // pinned
    Ctl* h_ctl = nullptr;            // pinned
    cudaGraphExec_t graph1 = nullptr;   // one step
    cudaGraphExec_t graphK = nullptr;   // kGraphSteps steps
    int kernels_per_step = 0;
    long long launches = 0;
    bool have_state = false;
    // the device-side step / contact counters run on between orb_step calls: the host keeps the values it has
    // already reported and only resets the control block after a halted step (one launch less per call)
    bool ctl_needs_reset = false;
    long long base_steps = 0, base_contacts = 0;
    // peer-memory reduction of the partial accelerations (orb_peer_*): the other ranks' acc buffers, mapped through
    // CUDA IPC (one process per GPU, NVLink / NVSwitch); [rank] is this engine's own buffer
    static const int kMaxPeers = 16;
    void* peer_base[kMaxPeers] = {};          // what cudaIpcOpenMemHandle returned (closed at destroy)
    const double* peer_acc[kMaxPeers] = {};
    int peers_open = 0;
};

namespace {

static const int kGraphSteps = [] {            // steps per replayed graph (ORBITAL_B200_GRAPH_STEPS)
    const char* env = getenv("ORBITAL_B200_GRAPH_STEPS");
    const int v = env ? atoi(env) : 16;
    return v >= 2 && v <= 1024 ? v : 16;
}();

void drop_graphs(orb_engine* e) {
    if (e->graph1) { cudaGraphExecDestroy(e->graph1); e->graph1 = nullptr; }
    if (e->graphK) { cudaGraphExecDestroy(e->graphK); e->graphK = nullptr; }
}

bool use_tiny(const orb_engine* e) {
    return e->mode == ORB_MODE_FAITHFUL && !e->sharded && e->s.n <= tiny_limit();
}

bool sym_applicable(const orb_engine* e) {
    return e->mode == ORB_MODE_FAST && e->use_sym && (!e->sharded || e->world > 0);
}

// pair-symmetric force on a sharded engine: every rank evaluates its share of the I-blocks (one per group of `world`, snake order) and
// ends up with a PARTIAL acceleration of all n bodies; the caller all-reduces (sum) acc across ranks.
bool acc_is_partial(const orb_engine* e) { return e->sharded && sym_applicable(e); }

int ensure_plan(orb_engine* e) {
    if (e->mode == ORB_MODE_FAITHFUL) {
        // two-pass bit-exact force: n x ld scratch matrix of 1/r^3 (allocated outside graph capture)
        const bool want = faithful_pairs_applicable(e->s.n, e->sharded) && !use_tiny(e) && !e->pairs_unavailable;
        if (want && !e->s.invr3) {
            e->s.invr3_ld = faithful_pairs_ld(e->s.n);
            const size_t bytes = sizeof(double) * (size_t)faithful_pairs_elems(e->s.n);
            if (cudaMalloc(&e->s.invr3, bytes) != cudaSuccess) {
                // no room for the 8 n^2-byte pair matrix: the one-pass kernel computes the same bits
                cudaGetLastError();
                e->s.invr3 = nullptr;
                e->pairs_unavailable = true;
                return ORB_OK;
            }
            CU(cudaMemsetAsync(e->s.invr3, 0, bytes, e->stream));      // padding rows / columns stay zero
        } else if (!want && e->s.invr3) {
            cudaFree(e->s.invr3);
            e->s.invr3 = nullptr;
        }
        return ORB_OK;
    }
    if (sym_applicable(e)) {
        if (!e->sym.valid) {
            CU(plan_sym(e->sym, e->s.n, e->sm_count, e->sharded ? e->rank : 0, e->sharded ? e->world : 1));
        }
        return ORB_OK;
    }
    if (e->mode == ORB_MODE_FAST && !e->plan_valid) {
        e->plan = plan_fast(e->s.tgt_hi - e->s.tgt_lo, e->s.n, e->sm_count);
        const long long need = e->plan.slabs > 1 ? (long long)e->plan.slabs * 3 * (e->s.tgt_hi - e->s.tgt_lo) : 0;
        if (need > e->s.scratch_elems) {
            if (e->s.scratch) cudaFree(e->s.scratch);
            e->s.scratch = nullptr;
            e->s.scratch_elems = 0;
            CU(cudaMalloc(&e->s.scratch, sizeof(double) * need));
            e->s.scratch_elems = need;
        }
        e->plan_valid = true;
    }
    return ORB_OK;
}

// one force evaluation on the resident positions (physics.py:125-159)
int enqueue_force(orb_engine* e, bool detect, int* launches) {
    if (e->mode == ORB_MODE_FAST) {
        int rc = ensure_plan(e);
        if (rc) return rc;
        if (sym_applicable(e))
            CU(launch_force_sym(e->s, e->p, e->sym, detect, e->stream, launches));
        else
            CU(launch_force_fast(e->s, e->p, e->plan, detect, e->stream, launches));
    } else {
        int rc = ensure_plan(e);
        if (rc) return rc;
        CU(launch_force_faithful(e->s, e->p, detect, e->stream, launches));
    }
    return ORB_OK;
}

// fast mode, pair-symmetric kernel, one panel: the reduction launch can carry the rest of the step
bool sym_tail_ok(orb_engine* e) {
    if (!(e->mode == ORB_MODE_FAST && sym_applicable(e) && !e->sharded)) return false;
    if (ensure_plan(e) != ORB_OK) return false;
    return sym_tail_applicable(e->sym);
}

// one full leapfrog step as separate kernels (engine.py:65-97).  begin: launch the first half-kick + drift (false
// when the previous step's reduction already did it); fuse_next: let this step's reduction do it for the next one.
int enqueue_step(orb_engine* e, int* launches, bool begin = true, bool fuse_next = false) {
    if (begin) {
        CU(launch_kick_drift(e->s, e->p, e->stream));
        ++*launches;
    }
    const bool contacts_here = e->p.device_contacts && e->detect;
    if (sym_tail_ok(e)) {
        // two launches (+ the contact sweep's): force_sym, then reduce + half-kick + history + bookkeeping
        const int mode = contacts_here ? kSymTailKick : (fuse_next ? kSymTailCloseNext : kSymTailClose);
        CU(launch_force_sym(e->s, e->p, e->sym, e->detect, e->stream, launches, mode));
        if (contacts_here) CU(launch_contacts(e->s, e->p, false, e->stream, launches));
        return ORB_OK;
    }
    if (e->mode == ORB_MODE_FAITHFUL && !contacts_here) {
        int rc = ensure_plan(e);
        if (rc) return rc;
        if (e->s.invr3) {
            // two-pass bit-exact force with the rest of the step fused into its second pass: three launches
            CU(launch_force_faithful(e->s, e->p, e->detect, e->stream, launches, true));
            return ORB_OK;
        }
    }
    int rc = enqueue_force(e, e->detect, launches);
    if (rc) return rc;
    CU(launch_kick_hist(e->s, e->p, e->stream));
    ++*launches;
    if (contacts_here) {
        // engine.py:85: contacts after the second half-kick -- resolved on the device, the step never halts
        const bool ordered_u = e->mode == ORB_MODE_FAITHFUL && e->s.n <= 4096;
        CU(launch_contacts(e->s, e->p, ordered_u, e->stream, launches));
    } else {
        CU(launch_advance(e->s, e->stream));
        ++*launches;
    }
    return ORB_OK;
}

int build_graph(orb_engine* e, int steps, cudaGraphExec_t* out) {
    int rc = ensure_plan(e);   // allocations must happen outside capture
    if (rc) return rc;
    cudaGraph_t graph = nullptr;
    // capture on the library's own stream: the bound stream may be the legacy default stream, which cannot
    // capture; the instantiated graph is launched into the bound stream
    cudaStream_t bound = e->stream;
    e->stream = e->own_stream;
    cudaError_t cb = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
    if (cb != cudaSuccess) { e->stream = bound; return cuda_fail(cb, "cudaStreamBeginCapture"); }
    int launches = 0;
    // inside a multi-step graph the reduction of step k also starts step k+1 (no contact sweep in between)
    const bool chain = steps > 1 && sym_tail_ok(e) && !(e->p.device_contacts && e->detect);
    for (int k = 0; k < steps && rc == ORB_OK; ++k)
        rc = enqueue_step(e, &launches, k == 0 || !chain, chain && k + 1 < steps);
    cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    e->stream = bound;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
    ce = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaGraphInstantiate");
    e->kernels_per_step = launches / steps;
    return ORB_OK;
}

int alloc_engine(orb_engine* e) {
    const long long n = e->s.n;
    CU(cudaMalloc(&e->s.pos4, sizeof(double4) * std::max(n, e->pos4_cap)));
    CU(cudaMemset(e->s.pos4, 0, sizeof(double4) * std::max(n, e->pos4_cap)));
    CU(cudaMalloc(&e->s.vel, sizeof(double) * 3 * n));
    CU(cudaMalloc(&e->s.acc, sizeof(double) * 3 * n));
    CU(cudaMalloc(&e->s.radius, sizeof(double) * n));
    CU(cudaMalloc(&e->s.vf32, n));
    CU(cudaMalloc(&e->s.ctl, sizeof(Ctl)));
    {
        const char* env = getenv("ORBITAL_B200_OVERLAP_CAP");
        const long long v = env ? atoll(env) : (long long)kOverlapCapDefault;
        e->s.pairs_cap = (int)std::max<long long>(1, std::min<long long>(v, 1LL << 28));
    }
    CU(cudaMalloc(&e->s.pairs, sizeof(long long) * 2 * e->s.pairs_cap));
    CU(cudaMalloc(&e->s.reduce_buf, sizeof(double) * 4 * ((n + 255) / 256 + 1)));
    CU(cudaMalloc(&e->d_stage, sizeof(double) * 4 * n));
    CU(cudaMalloc(&e->d_diag, sizeof(double) * 8));
    CU(cudaMallocHost(&e->h_diag, sizeof(double) * 8));
    CU(cudaMallocHost(&e->h_ctl, sizeof(Ctl)));
    CU(cudaMemset(e->s.vel, 0, sizeof(double) * 3 * n));
    CU(cudaMemset(e->s.acc, 0, sizeof(double) * 3 * n));
    {
        Ctl init;
        memset(&init, 0, sizeof init);
        init.pairs_cap = e->s.pairs_cap;
        CU(cudaMemcpy(e->s.ctl, &init, sizeof(Ctl), cudaMemcpyHostToDevice));
    }
    CU(cudaMemset(e->s.vf32, 0, n));
    CU(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    e->stream = e->own_stream;
    return ORB_OK;
}

void free_engine(orb_engine* e) {
    drop_graphs(e);
    free_sym(e->sym);
    cudaFree(e->s.pos4); cudaFree(e->s.vel); cudaFree(e->s.acc); cudaFree(e->s.radius); cudaFree(e->s.vf32);
    cudaFree(e->s.ctl); cudaFree(e->s.pairs); cudaFree(e->s.hist); cudaFree(e->s.scratch);
    cudaFree(e->s.reduce_buf); cudaFree(e->d_stage); cudaFree(e->d_diag); cudaFree(e->s.invr3);
    if (e->h_diag) cudaFreeHost(e->h_diag);
    if (e->h_ctl) cudaFreeHost(e->h_ctl);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
}

int select_device(int device) {
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(ORB_ERR_NO_DEVICE,
                    "no CUDA device available: liborbital_b200 has no CPU fallback (cudaGetDeviceCount: " +
                        std::string(ce == cudaSuccess ? "0 devices" : cudaGetErrorString(ce)) + ")");
    }
    if (device < 0 || device >= count) return fail(ORB_ERR_INVALID, "device index out of range");
    CU(cudaSetDevice(device));
    return ORB_OK;
}

#define LOCK(e)                                                  \
    if (!(e)) return fail(ORB_ERR_INVALID, "null handle");        \
    std::lock_guard<std::mutex> _lk((e)->mu);                     \
    CU(cudaSetDevice((e)->device))

}  // namespace

extern "C" {

int orb_abi_version(void) { return ORB_ABI_VERSION; }

const char* orb_last_error(void) { return g_err.c_str(); }

int orb_device_count(int* count) {
    if (!count) return fail(ORB_ERR_INVALID, "null argument");
    int c = 0;
    cudaError_t ce = cudaGetDeviceCount(&c);
    if (ce != cudaSuccess) { cudaGetLastError(); c = 0; }
    *count = c;
    return ORB_OK;
}

int orb_device_info(int device, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor,
                    int64_t* total_mem_bytes) {
    int rc = select_device(device);
    if (rc) return rc;
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, device));
    if (name && name_len > 0) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem_bytes) *total_mem_bytes = (int64_t)prop.totalGlobalMem;
    return ORB_OK;
}

int orb_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail(ORB_ERR_INVALID, "bad argument");
    CU(cudaMallocHost(ptr, (size_t)bytes));
    return ORB_OK;
}

int orb_host_free(void* ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return ORB_OK;
}

int orb_fp64_peak(int device, double seconds, double* tflops_best, double* tflops_mean, double* sm_clock_mhz) {
    int rc = select_device(device);
    if (rc) return rc;
    double a = 0, b = 0, c = 0;
    CU(run_fp64_peak(device, seconds, &a, &b, &c));
    if (tflops_best) *tflops_best = a;
    if (tflops_mean) *tflops_mean = b;
    if (sm_clock_mhz) *sm_clock_mhz = c;
    return ORB_OK;
}

}  // extern "C"

namespace {
int create_engine(orb_engine** out, int64_t n, int64_t tgt_lo, int64_t tgt_hi, int rank, int world, int device,
                  int mode) {
    if (!out) return fail(ORB_ERR_INVALID, "null out pointer");
    *out = nullptr;
    if (n <= 0) return fail(ORB_ERR_INVALID, "n must be positive");
    if (tgt_lo < 0 || tgt_hi > n || tgt_lo >= tgt_hi) return fail(ORB_ERR_INVALID, "bad target range");
    if (mode != ORB_MODE_FAITHFUL && mode != ORB_MODE_FAST) return fail(ORB_ERR_INVALID, "bad mode");
    if (world < 0 || (world > 0 && (rank < 0 || rank >= world))) return fail(ORB_ERR_INVALID, "bad rank / world");
    int rc = select_device(device);
    if (rc) return rc;
    orb_engine* e = new orb_engine();
    e->device = device;
    e->mode = mode;
    e->s.n = n;
    e->s.tgt_lo = tgt_lo;
    e->s.tgt_hi = tgt_hi;
    e->sharded = !(tgt_lo == 0 && tgt_hi == n);
    e->rank = e->sharded ? rank : 0;
    e->world = e->sharded ? world : 1;
    // equal-size in-place all-gathers of the packed positions need world * ceil(n / world) entries
    e->pos4_cap = world > 0 ? (long long)world * ((n + world - 1) / world) : n;
    {
        const char* env = getenv("ORBITAL_B200_SYM");     // "0": one-sided fast kernel even when unsharded
        e->use_sym = !(env && env[0] == '0');
    }
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) e->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        delete e;
        return fail(ORB_ERR_NO_DEVICE, "liborbital_b200 is built for sm_100a only; device is older");
    }
    e->p.dt = 1.0; e->p.h = 0.5; e->p.dt32 = 1.0f; e->p.eps2 = 0.0; e->p.G = 6.67430e-11;
    e->p.rmax1 = e->p.rmax2 = 0.0; e->p.rmax1_idx = -1; e->p.detect = 0;
    e->p.restitution = 1.0; e->p.device_contacts = 0; e->p.uniform_mass = 0.0;
    rc = alloc_engine(e);
    if (rc) { free_engine(e); delete e; return rc; }
    *out = e;
    return ORB_OK;
}
}  // namespace

extern "C" {

int orb_create_sharded(orb_engine** out, int64_t n, int64_t tgt_lo, int64_t tgt_hi, int device, int mode) {
    // equal slabs imply (rank, world); anything else cannot use the pair-symmetric kernel (world 0)
    int rank = 0, world = 0;
    const int64_t per = tgt_hi - tgt_lo;
    if (per > 0 && n % per == 0 && tgt_lo % per == 0) {
        world = (int)(n / per);
        rank = (int)(tgt_lo / per);
    }
    return create_engine(out, n, tgt_lo, tgt_hi, rank, world, device, mode);
}

int orb_create_ranked(orb_engine** out, int64_t n, int64_t tgt_lo, int64_t tgt_hi, int rank, int world, int device,
                      int mode) {
    if (world < 1) return fail(ORB_ERR_INVALID, "world must be >= 1");
    return create_engine(out, n, tgt_lo, tgt_hi, rank, world, device, mode);
}

int orb_create(orb_engine** out, int64_t n, int device, int mode) {
    return create_engine(out, n, 0, n, 0, 1, device, mode);
}

int orb_destroy(orb_engine* e) {
    if (!e) return ORB_OK;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        cudaSetDevice(e->device);
        cudaStreamSynchronize(e->stream);
        for (int r = 0; r < orb_engine::kMaxPeers; ++r)
            if (e->peer_base[r]) cudaIpcCloseMemHandle(e->peer_base[r]);
        free_engine(e);
    }
    delete e;
    return ORB_OK;
}

int orb_set_params(orb_engine* e, double dt, double eps, double G) {
    LOCK(e);
    e->p.dt = dt;
    e->p.h = 0.5 * dt;            // `0.5 * dt * acc` parses as (0.5*dt)*acc  (engine.py:70)
    e->p.dt32 = (float)dt;        // NEP-50 weak scalar -> float32(dt)          (engine.py:74)
    e->p.eps2 = eps * eps;        // physics.py:134
    e->p.G = G;
    drop_graphs(e);
    return ORB_OK;
}

int orb_set_contacts(orb_engine* e, double restitution, int resolve_on_device) {
    LOCK(e);
    e->p.restitution = restitution;
    e->p.device_contacts = resolve_on_device ? 1 : 0;
    drop_graphs(e);
    return ORB_OK;
}

int orb_set_mode(orb_engine* e, int mode) {
    LOCK(e);
    if (mode != ORB_MODE_FAITHFUL && mode != ORB_MODE_FAST) return fail(ORB_ERR_INVALID, "bad mode");
    e->mode = mode;
    e->plan_valid = false;
    drop_graphs(e);
    return ORB_OK;
}

int orb_set_history(orb_engine* e, int64_t capacity) {
    LOCK(e);
    if (capacity < 0) return fail(ORB_ERR_INVALID, "negative capacity");
    CU(cudaStreamSynchronize(e->stream));
    drop_graphs(e);
    if (e->s.hist) { cudaFree(e->s.hist); e->s.hist = nullptr; }
    e->s.hist_cap = 0;
    if (capacity > 0) {
        CU(cudaMalloc(&e->s.hist, sizeof(double) * 3 * e->s.n * capacity));
        e->s.hist_cap = capacity;
    }
    ctl_reset_kernel<<<1, 1, 0, e->stream>>>(e->s.ctl, 1, 0);
    CU(cudaGetLastError());
    e->base_steps = e->base_contacts = 0;
    e->ctl_needs_reset = false;
    return ORB_OK;
}

int orb_set_stream(orb_engine* e, void* cuda_stream) {
    LOCK(e);
    CU(cudaStreamSynchronize(e->stream));
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return ORB_OK;
}

int orb_upload(orb_engine* e, const double* x, const double* y, const double* z, const double* vx,
               const double* vy, const double* vz, const double* m, const double* radius,
               const uint8_t* vel_is_f32) {
    LOCK(e);
    if (!x || !y || !z || !vx || !vy || !vz || !m || !radius) return fail(ORB_ERR_INVALID, "null array");
    const long long n = e->s.n;
    const size_t nb = sizeof(double) * n;
    cudaStream_t st = e->stream;
    CU(cudaMemcpyAsync(e->d_stage, x, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_stage + n, y, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_stage + 2 * n, z, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_stage + 3 * n, m, nb, cudaMemcpyHostToDevice, st));
    CU(launch_pack(e->s.pos4, e->d_stage, e->d_stage + n, e->d_stage + 2 * n, e->d_stage + 3 * n, n, st));
    ++e->launches;
    CU(cudaMemcpyAsync(e->s.vel, vx, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->s.vel + n, vy, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->s.vel + 2 * n, vz, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->s.radius, radius, nb, cudaMemcpyHostToDevice, st));
    if (vel_is_f32)
        CU(cudaMemcpyAsync(e->s.vf32, vel_is_f32, n, cudaMemcpyHostToDevice, st));
    else
        CU(cudaMemsetAsync(e->s.vf32, 0, n, st));
    // overlap-prefilter bounds (largest two radii) for the fast kernel
    double r1 = 0.0, r2 = 0.0;
    long long i1 = -1;
    for (long long i = 0; i < n; ++i) {
        const double r = radius[i];
        if (r > r1) { r2 = r1; r1 = r; i1 = i; }
        else if (r > r2) { r2 = r; }
    }
    // one common mass lets the pair-symmetric kernel drop its per-pair mass multiplies
    double um = m[0];
    for (long long i = 1; i < n && um != 0.0; ++i)
        if (m[i] != um) um = 0.0;
    if (!(um == um) || um - um != 0.0) um = 0.0;      // NaN / inf masses take the general path
    if (um != e->p.uniform_mass) drop_graphs(e);
    e->p.uniform_mass = um;
    const bool detect = r1 > 0.0;      // all radii zero: a contact can only be dist == 0, a no-op (physics.py:396)
    if (detect != e->detect || r1 != e->p.rmax1 || r2 != e->p.rmax2 || i1 != e->p.rmax1_idx) drop_graphs(e);
    e->detect = detect;
    e->p.detect = detect;
    e->p.rmax1 = r1; e->p.rmax2 = r2; e->p.rmax1_idx = i1;
    CU(cudaStreamSynchronize(st));     // host buffers may be reused by the caller
    e->have_state = true;
    return ORB_OK;
}

int orb_download_state(orb_engine* e, double* x, double* y, double* z, double* vx, double* vy, double* vz) {
    LOCK(e);
    const long long n = e->s.n;
    const size_t nb = sizeof(double) * n;
    cudaStream_t st = e->stream;
    if (x || y || z) {
        CU(launch_unpack(e->s.pos4, e->d_stage, e->d_stage + n, e->d_stage + 2 * n, n, st));
        ++e->launches;
        if (x) CU(cudaMemcpyAsync(x, e->d_stage, nb, cudaMemcpyDeviceToHost, st));
        if (y) CU(cudaMemcpyAsync(y, e->d_stage + n, nb, cudaMemcpyDeviceToHost, st));
        if (z) CU(cudaMemcpyAsync(z, e->d_stage + 2 * n, nb, cudaMemcpyDeviceToHost, st));
    }
    if (vx) CU(cudaMemcpyAsync(vx, e->s.vel, nb, cudaMemcpyDeviceToHost, st));
    if (vy) CU(cudaMemcpyAsync(vy, e->s.vel + n, nb, cudaMemcpyDeviceToHost, st));
    if (vz) CU(cudaMemcpyAsync(vz, e->s.vel + 2 * n, nb, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orb_download_acc(orb_engine* e, double* ax, double* ay, double* az) {
    LOCK(e);
    const long long n = e->s.n;
    const size_t nb = sizeof(double) * n;
    if (ax) CU(cudaMemcpyAsync(ax, e->s.acc, nb, cudaMemcpyDeviceToHost, e->stream));
    if (ay) CU(cudaMemcpyAsync(ay, e->s.acc + n, nb, cudaMemcpyDeviceToHost, e->stream));
    if (az) CU(cudaMemcpyAsync(az, e->s.acc + 2 * n, nb, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return ORB_OK;
}

int orb_upload_acc(orb_engine* e, const double* ax, const double* ay, const double* az) {
    LOCK(e);
    if (!ax || !ay || !az) return fail(ORB_ERR_INVALID, "null array");
    const long long n = e->s.n;
    const size_t nb = sizeof(double) * n;
    CU(cudaMemcpyAsync(e->s.acc, ax, nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->s.acc + n, ay, nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->s.acc + 2 * n, az, nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return ORB_OK;
}

int orb_accel(orb_engine* e) {
    LOCK(e);
    if (!e->have_state) return fail(ORB_ERR_INVALID, "orb_accel before orb_upload");
    ctl_reset_kernel<<<1, 1, 0, e->stream>>>(e->s.ctl, 0, 1);     // a new force build: any stashed U is stale
    CU(cudaGetLastError());
    e->base_steps = e->base_contacts = 0;
    e->ctl_needs_reset = false;
    int launches = 1;
    int rc = enqueue_force(e, false, &launches);
    e->launches += launches;
    return rc;
}

int orb_step(orb_engine* e, int64_t nsteps, int64_t* steps_done, int64_t* n_overlaps) {
    LOCK(e);
    if (!e->have_state) return fail(ORB_ERR_INVALID, "orb_step before orb_upload");
    if (nsteps < 0) return fail(ORB_ERR_INVALID, "negative nsteps");
    if (e->sharded) return fail(ORB_ERR_INVALID, "sharded engines step with orb_step_begin/orb_step_finish");
    cudaStream_t st = e->stream;
    if (e->ctl_needs_reset) {                 // the previous call ended in a halted step
        ctl_reset_kernel<<<1, 1, 0, st>>>(e->s.ctl, 0, 0);
        CU(cudaGetLastError());
        ++e->launches;
        e->base_steps = e->base_contacts = 0;
        e->ctl_needs_reset = false;
    }
    if (nsteps > 0) {
        if (use_tiny(e)) {
            CU(launch_tiny_steps(e->s, e->p, nsteps, e->detect, st));
            ++e->launches;
        } else if (st == cudaStreamLegacy || st == nullptr) {
            // (stream capture is not available on the legacy default stream)
            for (int64_t k = 0; k < nsteps; ++k) {
                int launches = 0;
                int rc = enqueue_step(e, &launches);
                e->launches += launches;
                if (rc) return rc;
            }
        } else {
            int64_t left = nsteps;
            if (left >= kGraphSteps) {
                if (!e->graphK) { int rc = build_graph(e, kGraphSteps, &e->graphK); if (rc) return rc; }
                while (left >= kGraphSteps) {
                    CU(cudaGraphLaunch(e->graphK, st));
                    e->launches += (long long)e->kernels_per_step * kGraphSteps;
                    left -= kGraphSteps;
                }
            }
            if (left > 0) {
                if (!e->graph1) { int rc = build_graph(e, 1, &e->graph1); if (rc) return rc; }
                for (; left > 0; --left) {
                    CU(cudaGraphLaunch(e->graph1, st));
                    e->launches += e->kernels_per_step;
                }
            }
        }
    }
    CU(cudaMemcpyAsync(e->h_ctl, e->s.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (steps_done) *steps_done = e->h_ctl->steps_done - e->base_steps;
    if (n_overlaps)
        *n_overlaps = e->p.device_contacts ? (int64_t)(e->h_ctl->contacts_total - e->base_contacts)
                                           : (int64_t)e->h_ctl->overlap_count;
    e->base_steps = e->h_ctl->steps_done;
    e->base_contacts = e->h_ctl->contacts_total;
    if (e->h_ctl->halted) e->ctl_needs_reset = true;
    return ORB_OK;
}

int orb_overlap_pairs(orb_engine* e, int64_t* pairs_ij, int64_t cap, int64_t* count) {
    LOCK(e);
    if (!count) return fail(ORB_ERR_INVALID, "null count");
    CU(cudaMemcpyAsync(e->h_ctl, e->s.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    const int64_t total = e->h_ctl->overlap_count;
    *count = total;
    const int64_t stored = std::min<int64_t>(std::min<int64_t>(total, e->s.pairs_cap), cap);
    if (pairs_ij && stored > 0) {
        CU(cudaMemcpyAsync(pairs_ij, e->s.pairs, sizeof(long long) * 2 * stored, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    // total > the list's capacity: the caller sees count > stored pairs and must sweep all pairs
    return ORB_OK;
}

int orb_step_begin(orb_engine* e) {
    LOCK(e);
    if (!e->have_state) return fail(ORB_ERR_INVALID, "step before upload");
    CU(launch_kick_drift(e->s, e->p, e->stream));
    ++e->launches;
    return ORB_OK;
}

int orb_acc_needs_allreduce(orb_engine* e, int* flag) {
    LOCK(e);
    if (flag) *flag = acc_is_partial(e) ? 1 : 0;
    return ORB_OK;
}

int orb_step_finish(orb_engine* e) {
    LOCK(e);
    if (acc_is_partial(e))
        return fail(ORB_ERR_INVALID, "this sharded engine produces partial accelerations: use orb_accel, all-reduce "
                                     "orb_acc_ptr across ranks, then orb_step_kick");
    int launches = 0;
    int rc = enqueue_force(e, e->detect, &launches);     // engine.py:78 + the overlap test of :85 (physics.py:517)
    if (rc) return rc;
    CU(launch_kick_hist(e->s, e->p, e->stream));
    e->launches += launches + 1;
    return ORB_OK;
}

int orb_step_kick(orb_engine* e) {
    LOCK(e);
    CU(launch_kick_hist(e->s, e->p, e->stream));
    ++e->launches;
    if (!e->sharded) {              // a sharded step is closed by orb_step_end, once all ranks' contacts are known
        CU(launch_advance(e->s, e->stream));
        ++e->launches;
    }
    return ORB_OK;
}

int orb_step_force(orb_engine* e) {
    LOCK(e);
    if (!e->have_state) return fail(ORB_ERR_INVALID, "orb_step_force before orb_upload");
    int launches = 0;
    int rc = enqueue_force(e, e->detect, &launches);
    e->launches += launches;
    return rc;
}

int orb_overlap_count(orb_engine* e, int64_t* count, int* overflowed) {
    LOCK(e);
    CU(cudaMemcpyAsync(e->h_ctl, e->s.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (count) *count = std::min<int64_t>(e->h_ctl->overlap_count, e->s.pairs_cap);
    if (overflowed) *overflowed = (e->h_ctl->overlap_overflow > 0 || e->h_ctl->overlap_count > e->s.pairs_cap) ? 1 : 0;
    return ORB_OK;
}

int orb_set_overlap_pairs(orb_engine* e, const int64_t* pairs_ij, int64_t count, int overflowed) {
    LOCK(e);
    if (count < 0 || (count > 0 && !pairs_ij)) return fail(ORB_ERR_INVALID, "bad pair list");
    if (count > e->s.pairs_cap) { count = e->s.pairs_cap; overflowed = 1; }
    if (count > 0)
        CU(cudaMemcpyAsync(e->s.pairs, pairs_ij, sizeof(long long) * 2 * count, cudaMemcpyHostToDevice, e->stream));
    CU(launch_set_overlaps(e->s, (int)count, overflowed ? 1 : 0, e->stream));
    ++e->launches;
    CU(cudaStreamSynchronize(e->stream));      // the host list may be reused by the caller
    return ORB_OK;
}

int orb_step_end(orb_engine* e) {
    LOCK(e);
    if (!e->sharded) return fail(ORB_ERR_INVALID, "orb_step_end closes a sharded step; orb_step_kick closes an unsharded one");
    const bool resolve = e->detect && e->p.device_contacts;
    const bool ordered_u = e->mode == ORB_MODE_FAITHFUL && e->s.n <= 4096;
    int launches = 0;
    CU(launch_step_end(e->s, e->p, resolve, ordered_u, e->stream, &launches));
    e->launches += launches;
    return ORB_OK;
}

int orb_contact_stats(orb_engine* e, int64_t* contacts_total, int64_t* full_sweeps) {
    LOCK(e);
    CU(cudaMemcpyAsync(e->h_ctl, e->s.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (contacts_total) *contacts_total = e->h_ctl->contacts_total;
    if (full_sweeps) *full_sweeps = e->h_ctl->full_sweeps;
    return ORB_OK;
}

int orb_synchronize(orb_engine* e) {
    LOCK(e);
    CU(cudaStreamSynchronize(e->stream));
    return ORB_OK;
}

int orb_pos4_ptr(orb_engine* e, void** device_ptr, int64_t* n_bodies) {
    LOCK(e);
    if (device_ptr) *device_ptr = e->s.pos4;
    if (n_bodies) *n_bodies = e->s.n;
    return ORB_OK;
}

int orb_vel_ptr(orb_engine* e, void** device_ptr) {
    LOCK(e);
    if (device_ptr) *device_ptr = e->s.vel;
    return ORB_OK;
}

// ---- peer-memory reduction of the partial accelerations (one process per GPU) -------------------------------
// The pair-symmetric kernel leaves every rank with a PARTIAL acceleration of all n bodies; a rank needs the total
// only for its own slab (the second half-kick).  Instead of an NCCL reduce-scatter / all-reduce (latency-bound:
// 0.33 ms for 6 MB at 8 GPUs) each rank maps the other ranks' acc buffers (CUDA IPC) and sums its slab's columns
// straight out of peer memory over NVLink, in fixed rank order (deterministic).  The caller provides the two
// synchronisation points: every rank has finished its force pass before any rank reduces (a 1-element all-reduce),
// and nobody starts the next force pass before all ranks have reduced (the next step's position all-gather).
static CUresult (*p_cuMemGetAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr) = nullptr;

int orb_peer_export(orb_engine* e, void* handle64, int64_t* offset) {
    LOCK(e);
    if (!handle64 || !offset) return fail(ORB_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!p_cuMemGetAddressRange) {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        CU(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
        if (!fn) return fail(ORB_ERR_CUDA, "cuMemGetAddressRange unavailable");
        p_cuMemGetAddressRange = reinterpret_cast<decltype(p_cuMemGetAddressRange)>(fn);
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (p_cuMemGetAddressRange(&base, &size, (CUdeviceptr)e->s.acc) != CUDA_SUCCESS)
        return fail(ORB_ERR_CUDA, "cuMemGetAddressRange failed");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, (void*)base));
    memcpy(handle64, &h, 64);
    *offset = (int64_t)((CUdeviceptr)e->s.acc - base);
    return ORB_OK;
}

int orb_peer_open(orb_engine* e, int rank, const void* handle64, int64_t offset) {
    LOCK(e);
    if (!e->sharded || e->world < 2) return fail(ORB_ERR_INVALID, "not a multi-rank engine");
    if (rank < 0 || rank >= e->world || rank >= orb_engine::kMaxPeers || rank == e->rank || !handle64)
        return fail(ORB_ERR_INVALID, "bad peer rank");
    if (e->peer_base[rank]) return fail(ORB_ERR_INVALID, "peer already open");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* base = nullptr;
    CU(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    e->peer_base[rank] = base;
    e->peer_acc[rank] = reinterpret_cast<const double*>(static_cast<char*>(base) + offset);
    ++e->peers_open;
    return ORB_OK;
}

int orb_peer_close(orb_engine* e) {
    LOCK(e);
    CU(cudaStreamSynchronize(e->stream));
    for (int r = 0; r < orb_engine::kMaxPeers; ++r) {
        if (e->peer_base[r]) cudaIpcCloseMemHandle(e->peer_base[r]);
        e->peer_base[r] = nullptr;
        e->peer_acc[r] = nullptr;
    }
    e->peers_open = 0;
    return ORB_OK;
}

int orb_peer_reduce(orb_engine* e) {
    LOCK(e);
    if (!acc_is_partial(e)) return fail(ORB_ERR_INVALID, "this engine's accelerations are not partial sums");
    if (e->peers_open != e->world - 1) return fail(ORB_ERR_INVALID, "orb_peer_open every other rank first");
    e->peer_acc[e->rank] = e->s.acc;
    CU(launch_peer_reduce(e->peer_acc, e->world, e->s.acc, e->s.n, e->s.tgt_lo, e->s.tgt_hi, e->stream));
    ++e->launches;
    return ORB_OK;
}

int orb_acc_ptr(orb_engine* e, void** device_ptr) {
    LOCK(e);
    if (device_ptr) *device_ptr = e->s.acc;
    return ORB_OK;
}

int orb_force_kernel_info(orb_engine* e, char* name, int name_len, int* grid, int* block, int* smem_bytes,
                          int* launches_per_step) {
    LOCK(e);
    const char* nm;
    int g = 0, b = 0, sm = 0, lps = 4;
    if (use_tiny(e)) {
        nm = e->s.n <= 64 ? "micro_steps_kernel" : "tiny_steps_kernel";
        g = 1; b = tiny_block((int)e->s.n);
        sm = (int)(e->s.n * 64 + 16); lps = 1;
    } else if (e->mode == ORB_MODE_FAST) {
        int rc = ensure_plan(e);
        if (rc) return rc;
        if (sym_applicable(e)) {
            nm = sym_kernel_name(e->sym.ti, e->detect, sym_uniform(e->sym, e->p));
            g = 0;
            for (const auto& pan : e->sym.panels) g += pan.n_items;
            b = 128; sm = 0;
            lps = 3 + 2 * (int)e->sym.panels.size();
        } else {
            nm = fast_kernel_name(e->plan.ti, e->detect);
            g = e->plan.grid; b = e->plan.block; sm = e->plan.smem;
            lps = 4 + (e->plan.slabs > 1 ? 1 : 0);
        }
    } else {
        const bool two_pass = faithful_pairs_applicable(e->s.n, e->sharded) && !e->pairs_unavailable;
        nm = two_pass ? faithful_two_pass_name(e->s.n) : "force_faithful_kernel";
        faithful_geometry(e->s.tgt_hi - e->s.tgt_lo, &g, &b);
        sm = b * 40;
        if (two_pass) lps = (e->p.device_contacts && e->detect) ? 5 : 3;     // step tail fused into pass 2
    }
    if (name && name_len > 0) { strncpy(name, nm, name_len - 1); name[name_len - 1] = 0; }
    if (grid) *grid = g;
    if (block) *block = b;
    if (smem_bytes) *smem_bytes = sm;
    if (launches_per_step) *launches_per_step = lps;
    return ORB_OK;
}

int orb_launch_count(orb_engine* e, int64_t* launches) {
    LOCK(e);
    if (launches) *launches = e->launches;
    return ORB_OK;
}

int orb_potential(orb_engine* e, double* U) {
    LOCK(e);
    if (!U) return fail(ORB_ERR_INVALID, "null U");
    // a step with device-resolved contacts stashed U of its force build before the push-out moved bodies
    CU(cudaMemcpyAsync(e->h_ctl, e->s.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (e->h_ctl->u_valid) { *U = e->h_ctl->u_stash; return ORB_OK; }
    const bool ordered = (e->mode == ORB_MODE_FAITHFUL) && e->s.n <= 4096;
    int launches = 0;
    CU(launch_potential(e->s, e->p, ordered, e->d_diag, e->stream, &launches));
    e->launches += launches;
    CU(cudaMemcpyAsync(e->h_diag, e->d_diag, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    *U = e->h_diag[0];
    return ORB_OK;
}

int orb_body_potential(orb_engine* e, int64_t body, double G, double* pe) {
    LOCK(e);
    if (!pe) return fail(ORB_ERR_INVALID, "null pe");
    if (body < 0 || body >= e->s.n) return fail(ORB_ERR_INVALID, "body index out of range");
    CU(launch_body_potential(e->s, body, G, e->d_diag, e->stream));
    e->launches += 1;
    CU(cudaMemcpyAsync(e->h_diag, e->d_diag, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    *pe = e->h_diag[0];
    return ORB_OK;
}

int orb_energy_angmom(orb_engine* e, double* K, double* L3) {
    LOCK(e);
    int launches = 0;
    CU(launch_energy_angmom(e->s, e->d_diag, e->stream, &launches));
    e->launches += launches;
    CU(cudaMemcpyAsync(e->h_diag, e->d_diag, sizeof(double) * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (K) *K = e->h_diag[0];
    if (L3) { L3[0] = e->h_diag[1]; L3[1] = e->h_diag[2]; L3[2] = e->h_diag[3]; }
    return ORB_OK;
}

int orb_history_count(orb_engine* e, int64_t* total_appended) {
    LOCK(e);
    CU(cudaMemcpyAsync(e->h_ctl, e->s.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (total_appended) *total_appended = e->h_ctl->hist_count;
    return ORB_OK;
}

int orb_history_append(orb_engine* e) {
    LOCK(e);
    if (e->s.hist_cap <= 0) return ORB_OK;
    CU(launch_hist_append(e->s, e->stream));
    e->launches += 2;
    return ORB_OK;
}

int orb_history_download(orb_engine* e, int64_t last_k, double* out, int64_t* got) {
    LOCK(e);
    if (!got) return fail(ORB_ERR_INVALID, "null got");
    *got = 0;
    if (e->s.hist_cap <= 0 || last_k <= 0) return ORB_OK;
    CU(cudaMemcpyAsync(e->h_ctl, e->s.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    const long long total = e->h_ctl->hist_count;
    const long long stored = std::min<long long>(total, e->s.hist_cap);
    const long long k = std::min<long long>(last_k, stored);
    if (k <= 0 || !out) { *got = k; return ORB_OK; }
    const long long row = 3 * e->s.n;
    const long long first = total - k;                  // logical index of the oldest wanted snapshot
    long long done = 0;
    while (done < k) {
        const long long slot = (first + done) % e->s.hist_cap;
        const long long run = std::min<long long>(k - done, e->s.hist_cap - slot);
        CU(cudaMemcpyAsync(out + done * row, e->s.hist + slot * row, sizeof(double) * row * run,
                           cudaMemcpyDeviceToHost, e->stream));
        done += run;
    }
    CU(cudaStreamSynchronize(e->stream));
    *got = k;
    return ORB_OK;
}

}  // extern "C"

// ===========================================================================
// Ensemble
// ===========================================================================
constexpr int kEnsNarrowBelow = 6;      // warps per SM sub-partition below which the fast kernel goes one body per lane

struct orb_ensemble {
    std::mutex mu;
    int device = 0;
    int mode = ORB_MODE_FAST;
    EnsArgs a{};
    double* base = nullptr;     // 10 planes of nsys*nb doubles
    double* d_radius = nullptr; // optional plane: contact handling (orb_ens_set_bodies)
    uint8_t* d_vf32 = nullptr;  // optional per-body velocity-dtype flags
    unsigned long long* d_contacts = nullptr;
    double* d_E = nullptr;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    long long launches = 0;
    bool have_state = false;
    cudaGraphExec_t step_graph = nullptr;     // kEnsGraphSteps un-fused steps (one launch each) as one graph
    bool use_pdl = false;                     // ORBITAL_B200_ENS_PDL=1: programmatic dependent launch between the steps
                                              // (measured: no gain at any batch size, a loss at 4,096 systems)
    cudaStream_t br_stream[4] = {nullptr, nullptr, nullptr, nullptr};   // capture-only: branches of the step graph
    cudaEvent_t br_event[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork_event = nullptr;
    int sm_count = 148;
    int slice = 32;                           // steps per item of the time-sliced fused kernel (ORBITAL_B200_ENS_SLICE, 0 = off)
    unsigned long long* d_queue = nullptr;    // work queue head + per-group progress of the time-sliced kernel
    int* d_progress = nullptr;
    long long progress_len = 0;
};

namespace {
static const int kEnsGraphSteps = [] {         // un-fused ensemble steps per replayed graph (ORBITAL_B200_ENS_GRAPH_STEPS)
    const char* env = getenv("ORBITAL_B200_ENS_GRAPH_STEPS");
    const int v = env ? atoi(env) : 16;
    return v >= 2 && v <= 1024 ? v : 16;
}();
constexpr int kEnsBranches = 4;          // independent sub-batches of a small ensemble inside one step graph

void ens_drop_graph(orb_ensemble* s) {
    if (s->step_graph) { cudaGraphExecDestroy(s->step_graph); s->step_graph = nullptr; }
}

// small per-GPU batches make the one-step-per-launch mode launch-bound: replay 16 launches as one graph.
// The graph starts from and ends with the synchronised (x, v, a) state (first / last, ensemble.cu); the 14
// launches in between exchange only x and the half-kicked velocity.
// the systems [lo, hi) of an ensemble as an ensemble of their own (they are contiguous in every plane)
EnsArgs ens_subrange(const EnsArgs& a, long long lo, long long hi) {
    EnsArgs b = a;
    const long long off = lo * a.nb;
    b.x += off; b.y += off; b.z += off; b.vx += off; b.vy += off; b.vz += off; b.ax += off; b.ay += off; b.az += off;
    b.m += off;
    if (b.radius) b.radius += off;
    if (b.vf32) b.vf32 += off;
    b.nsys = hi - lo;
    return b;
}

int ens_build_graph(orb_ensemble* s) {
    EnsArgs a = s->a;
    a.nsteps = 1;
    cudaGraph_t graph = nullptr;
    // A small batch is one wave of warps that all load, then all compute, then all store: a launch is a
    // dependent latency chain (~4 us at 8,192 systems) with nothing to overlap it.  Cut the batch into
    // independent branches of the graph -- systems never interact -- so that one branch computes while another
    // waits on memory.
    // Measured (profiles/r2_ens_branch_sweep*.txt, 16-body systems, us per step for 1 / 3 branches): 16,384 systems
    // 6.4 / 5.2, 32,768: 11.3 / 9.7, 65,536: 21.0 / 19.3; no gain from 131,072 up (several waves overlap by
    // themselves) nor at <= 8,192, where a step is bound by the ~3 us a dependent kernel node costs.
    const long long warps = (a.nsys * (long long)a.nbp + 63) / 64;
    int branches = 1;
    if (s->mode == ORB_MODE_FAST && warps <= 16384) branches = 3;
    {
        const char* env = getenv("ORBITAL_B200_ENS_BRANCHES");
        if (env) branches = std::max(1, std::min(kEnsBranches, atoi(env)));
    }
    branches = (int)std::min<long long>(branches, a.nsys);
    // capture on the library's own stream (the caller's may be the legacy default stream, which cannot capture);
    // the instantiated graph is launched into whatever stream the handle is bound to
    CU(cudaStreamBeginCapture(s->own_stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t ce = cudaSuccess;
    if (branches > 1) {
        for (int b = 0; b < kEnsBranches && ce == cudaSuccess; ++b) {
            if (!s->br_stream[b]) ce = cudaStreamCreateWithFlags(&s->br_stream[b], cudaStreamNonBlocking);
            if (ce == cudaSuccess && !s->br_event[b]) ce = cudaEventCreateWithFlags(&s->br_event[b], cudaEventDisableTiming);
        }
        if (ce == cudaSuccess && !s->fork_event) ce = cudaEventCreateWithFlags(&s->fork_event, cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventRecord(s->fork_event, s->own_stream);
        for (int b = 0; b < branches && ce == cudaSuccess; ++b) {
            const long long lo = a.nsys * b / branches, hi = a.nsys * (b + 1) / branches;
            EnsArgs sub = ens_subrange(a, lo, hi);
            ce = cudaStreamWaitEvent(s->br_stream[b], s->fork_event, 0);
            for (int k = 0; k < kEnsGraphSteps && ce == cudaSuccess; ++k) {
                sub.first = k == 0;
                sub.last = k == kEnsGraphSteps - 1;
                sub.pdl = (k > 0 && s->use_pdl) ? 1 : 0;
                ce = launch_ens_step(sub, false, s->br_stream[b]);
            }
            if (ce == cudaSuccess) ce = cudaEventRecord(s->br_event[b], s->br_stream[b]);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s->own_stream, s->br_event[b], 0);
        }
    } else {
        for (int k = 0; k < kEnsGraphSteps && ce == cudaSuccess; ++k) {
            a.first = k == 0;
            a.last = k == kEnsGraphSteps - 1;
            a.pdl = (k > 0 && s->use_pdl) ? 1 : 0;     // programmatic edge to the previous step of the graph
            ce = launch_ens_step(a, s->mode == ORB_MODE_FAITHFUL, s->own_stream);
        }
    }
    cudaError_t ce2 = cudaStreamEndCapture(s->own_stream, &graph);
    if (ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return cuda_fail(ce, "ensemble graph capture"); }
    if (ce2 != cudaSuccess) return cuda_fail(ce2, "cudaStreamEndCapture");
    ce = cudaGraphInstantiate(&s->step_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaGraphInstantiate");
    return ORB_OK;
}
}  // namespace

extern "C" {

int orb_ens_create(orb_ensemble** out, int64_t nsys, int nbody, int device, int mode, int vel_f32) {
    if (!out) return fail(ORB_ERR_INVALID, "null out pointer");
    *out = nullptr;
    if (nsys <= 0 || nbody < 2 || nbody > 32) return fail(ORB_ERR_INVALID, "need nsys > 0 and 2 <= nbody <= 32");
    if (mode != ORB_MODE_FAITHFUL && mode != ORB_MODE_FAST) return fail(ORB_ERR_INVALID, "bad mode");
    int rc = select_device(device);
    if (rc) return rc;
    orb_ensemble* s = new orb_ensemble();
    s->device = device;
    s->mode = mode;
    const long long tot = nsys * (long long)nbody;
    cudaError_t ce = cudaMalloc(&s->base, sizeof(double) * 10 * tot);
    if (ce == cudaSuccess) ce = cudaMalloc(&s->d_E, sizeof(double) * nsys);
    if (ce == cudaSuccess) ce = cudaMalloc(&s->d_contacts, sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMemset(s->d_contacts, 0, sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) {
        cudaFree(s->base); cudaFree(s->d_E); cudaFree(s->d_contacts);
        delete s;
        return cuda_fail(ce, "ensemble alloc");
    }
    s->stream = s->own_stream;
    double* b = s->base;
    s->a.x = b; s->a.y = b + tot; s->a.z = b + 2 * tot; s->a.vx = b + 3 * tot; s->a.vy = b + 4 * tot;
    s->a.vz = b + 5 * tot; s->a.ax = b + 6 * tot; s->a.ay = b + 7 * tot; s->a.az = b + 8 * tot; s->a.m = b + 9 * tot;
    s->a.nsys = nsys;
    s->a.nb = nbody;
    int nbp = 1;
    while (nbp < nbody) nbp <<= 1;
    s->a.nbp = nbp;
    s->a.vel_f32 = vel_f32 ? 1 : 0;
    s->a.radius = nullptr; s->a.vf32 = nullptr; s->a.restitution = 1.0; s->a.contacts = s->d_contacts;
    s->a.first = s->a.last = 1;
    s->a.pdl = 0;
    {
        const char* env = getenv("ORBITAL_B200_ENS_PDL");
        s->use_pdl = env && env[0] == '1';
        env = getenv("ORBITAL_B200_ENS_SLICE");
        if (env) s->slice = std::max(0, atoi(env));
        cudaDeviceProp prop{};
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) s->sm_count = prop.multiProcessorCount;
        s->progress_len = (nsys * (long long)nbp + 63) / 64;        // groups of 64 / nbp systems (one warp each)
        cudaError_t c2 = cudaMalloc(&s->d_queue, sizeof(unsigned long long));
        if (c2 == cudaSuccess) c2 = cudaMalloc(&s->d_progress, sizeof(int) * s->progress_len);
        if (c2 != cudaSuccess) { cudaGetLastError(); s->slice = 0; }
    }
    {
        // one warp per system; several systems share a CTA (measured: 1 -> 3.5 TB/s, >= 2 -> 4.1 TB/s)
        const char* env = getenv("ORBITAL_B200_ENS_WARPS");
        s->a.warps_per_cta = env ? atoi(env) : 4;
        // fast mode: one body per lane when the two-body layout would leave fewer than kEnsNarrowBelow warps per SM
        // sub-partition (ORBITAL_B200_ENS_NARROW = 0 | 1 forces it off / on)
        env = getenv("ORBITAL_B200_ENS_NARROW");
        const long long warps2 = (nsys * (long long)nbp + 63) / 64;
        s->a.narrow = env ? (env[0] == '1') : (warps2 < (long long)kEnsNarrowBelow * 4 * std::max(1, s->sm_count));
    }
    s->a.dt = 1.0; s->a.h = 0.5; s->a.dt32 = 1.0f; s->a.eps2 = 0.0; s->a.G = 6.67430e-11;
    *out = s;
    return ORB_OK;
}

int orb_ens_destroy(orb_ensemble* s) {
    if (!s) return ORB_OK;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->stream);
        ens_drop_graph(s);
        cudaFree(s->base); cudaFree(s->d_E); cudaFree(s->d_radius); cudaFree(s->d_vf32); cudaFree(s->d_contacts);
        cudaFree(s->d_queue); cudaFree(s->d_progress);
        for (int b = 0; b < 4; ++b) {
            if (s->br_stream[b]) cudaStreamDestroy(s->br_stream[b]);
            if (s->br_event[b]) cudaEventDestroy(s->br_event[b]);
        }
        if (s->fork_event) cudaEventDestroy(s->fork_event);
        if (s->own_stream) cudaStreamDestroy(s->own_stream);
    }
    delete s;
    return ORB_OK;
}

int orb_ens_set_params(orb_ensemble* s, double dt, double eps, double G) {
    LOCK(s);
    s->a.dt = dt; s->a.h = 0.5 * dt; s->a.dt32 = (float)dt; s->a.eps2 = eps * eps; s->a.G = G;
    ens_drop_graph(s);
    return ORB_OK;
}

int orb_ens_set_bodies(orb_ensemble* s, const double* radius, const uint8_t* vel_is_f32) {
    LOCK(s);
    const long long tot = s->a.nsys * (long long)s->a.nb;
    CU(cudaStreamSynchronize(s->stream));
    ens_drop_graph(s);
    bool any_radius = false;
    if (radius)
        for (long long i = 0; i < tot && !any_radius; ++i) any_radius = radius[i] > 0.0;
    // all radii zero: a contact can only be dist == 0, a no-op (physics.py:396) -- no sweep needed
    if (any_radius) {
        if (!s->d_radius) CU(cudaMalloc(&s->d_radius, sizeof(double) * tot));
        CU(cudaMemcpy(s->d_radius, radius, sizeof(double) * tot, cudaMemcpyHostToDevice));
        s->a.radius = s->d_radius;
    } else {
        s->a.radius = nullptr;
    }
    if (vel_is_f32 || any_radius) {               // the contact variant always reads per-body flags
        if (!s->d_vf32) CU(cudaMalloc(&s->d_vf32, (size_t)tot));
        if (vel_is_f32)
            CU(cudaMemcpy(s->d_vf32, vel_is_f32, (size_t)tot, cudaMemcpyHostToDevice));
        else
            CU(cudaMemset(s->d_vf32, s->a.vel_f32 ? 1 : 0, (size_t)tot));
        s->a.vf32 = s->d_vf32;
    } else {
        s->a.vf32 = nullptr;
    }
    return ORB_OK;
}

int orb_ens_set_contacts(orb_ensemble* s, double restitution) {
    LOCK(s);
    s->a.restitution = restitution;
    ens_drop_graph(s);
    return ORB_OK;
}

int orb_ens_contact_count(orb_ensemble* s, int64_t* contacts) {
    LOCK(s);
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, s->d_contacts, sizeof v, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (contacts) *contacts = (int64_t)v;
    return ORB_OK;
}

int orb_ens_download_acc(orb_ensemble* s, double* ax, double* ay, double* az) {
    LOCK(s);
    const size_t nb = sizeof(double) * s->a.nsys * s->a.nb;
    if (ax) CU(cudaMemcpyAsync(ax, s->a.ax, nb, cudaMemcpyDeviceToHost, s->stream));
    if (ay) CU(cudaMemcpyAsync(ay, s->a.ay, nb, cudaMemcpyDeviceToHost, s->stream));
    if (az) CU(cudaMemcpyAsync(az, s->a.az, nb, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return ORB_OK;
}

int orb_ens_set_stream(orb_ensemble* s, void* cuda_stream) {
    LOCK(s);
    CU(cudaStreamSynchronize(s->stream));
    s->stream = cuda_stream ? (cudaStream_t)cuda_stream : s->own_stream;
    return ORB_OK;
}

int orb_ens_upload(orb_ensemble* s, const double* x, const double* y, const double* z, const double* vx,
                   const double* vy, const double* vz, const double* m) {
    LOCK(s);
    if (!x || !y || !z || !vx || !vy || !vz || !m) return fail(ORB_ERR_INVALID, "null array");
    const size_t nb = sizeof(double) * s->a.nsys * s->a.nb;
    cudaStream_t st = s->stream;
    CU(cudaMemcpyAsync(s->a.x, x, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->a.y, y, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->a.z, z, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->a.vx, vx, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->a.vy, vy, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->a.vz, vz, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(const_cast<double*>(s->a.m), m, nb, cudaMemcpyHostToDevice, st));
    EnsArgs a = s->a;
    a.nsteps = 0;
    CU(launch_ens_accel(a, s->mode == ORB_MODE_FAITHFUL, st));     // engine.py:41
    ++s->launches;
    CU(cudaStreamSynchronize(st));
    s->have_state = true;
    return ORB_OK;
}

int orb_ens_upload_elements(orb_ensemble* s, const double* M, const double* e, const double* a, const double* inc,
                            const double* Omega, const double* omega, const double* m) {
    LOCK(s);
    if (!M || !e || !a || !inc || !Omega || !omega || !m) return fail(ORB_ERR_INVALID, "null array");
    const long long pl = s->a.nsys * (long long)(s->a.nb - 1);
    const size_t pb = sizeof(double) * pl;
    cudaStream_t st = s->stream;
    double* d_el = nullptr;
    CU(cudaMalloc(&d_el, 8 * pb));
    std::vector<double> e2(pl), a3(pl);
    host_pow_plane(e, 2.0, e2.data(), pl);
    host_pow_plane(a, 3.0, a3.data(), pl);
    const double* src[8] = {M, e, a, inc, Omega, omega, e2.data(), a3.data()};
    cudaError_t ce = cudaSuccess;
    for (int k = 0; k < 8 && ce == cudaSuccess; ++k)
        ce = cudaMemcpyAsync(d_el + k * pl, src[k], pb, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(const_cast<double*>(s->a.m), m, sizeof(double) * s->a.nsys * s->a.nb,
                             cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = launch_ens_elements(s->a, d_el, 1e-12, 50, trig_mode(), st);       // physics.py:43 defaults
    EnsArgs a0 = s->a;
    a0.nsteps = 0;
    if (ce == cudaSuccess) ce = launch_ens_accel(a0, s->mode == ORB_MODE_FAITHFUL, st);   // engine.py:41
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(d_el);
    if (ce != cudaSuccess) return cuda_fail(ce, "orb_ens_upload_elements");
    s->launches += 2;
    s->have_state = true;
    return ORB_OK;
}

int orb_set_trig_mode(int mode) {
    if (mode != ORB_TRIG_LIBM && mode != ORB_TRIG_CR && mode != ORB_TRIG_FAST)
        return fail(ORB_ERR_INVALID, "trig mode must be ORB_TRIG_LIBM, ORB_TRIG_CR or ORB_TRIG_FAST");
    g_trig.store(mode, std::memory_order_relaxed);
    return ORB_OK;
}

int orb_get_trig_mode(void) { return trig_mode(); }

int orb_kepler_states(int device, int64_t count, const double* M, const double* e, const double* a, const double* b,
                      const double* n, const double* inc, const double* Omega, const double* omega, double tol,
                      int max_iter, double* r3, double* v3, double* E) {
    if (count < 0) return fail(ORB_ERR_INVALID, "negative count");
    if (!M || !e || !a || !b || !n || !inc || !Omega || !omega || !r3 || !v3)
        return fail(ORB_ERR_INVALID, "null array");
    if (count == 0) return ORB_OK;
    int rc = select_device(device);
    if (rc) return rc;
    const size_t pb = sizeof(double) * count;
    double *d_el = nullptr, *d_out = nullptr;
    CU(cudaMalloc(&d_el, 9 * pb));
    cudaError_t ce = cudaMalloc(&d_out, 7 * pb);
    std::vector<double> e2(count);
    host_pow_plane(e, 2.0, e2.data(), count);
    const double* src[9] = {M, e, a, b, n, inc, Omega, omega, e2.data()};
    for (int k = 0; k < 9 && ce == cudaSuccess; ++k)
        ce = cudaMemcpy(d_el + k * count, src[k], pb, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = launch_kepler_states(d_el, d_out, count, tol, max_iter, trig_mode(), nullptr);
    if (ce == cudaSuccess) ce = cudaMemcpy(r3, d_out, 3 * pb, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaMemcpy(v3, d_out + 3 * count, 3 * pb, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && E) ce = cudaMemcpy(E, d_out + 6 * count, pb, cudaMemcpyDeviceToHost);
    cudaFree(d_el);
    cudaFree(d_out);
    if (ce != cudaSuccess) return cuda_fail(ce, "orb_kepler_states");
    return ORB_OK;
}

int orb_ens_step(orb_ensemble* s, int64_t nsteps, int fused) {
    LOCK(s);
    if (!s->have_state) return fail(ORB_ERR_INVALID, "step before upload");
    if (nsteps < 0) return fail(ORB_ERR_INVALID, "negative nsteps");
    EnsArgs a = s->a;
    const bool faithful = s->mode == ORB_MODE_FAITHFUL;
    if (fused) {
        a.nsteps = nsteps;
        a.first = a.last = 1;
        // few warps of work per SM sub-partition: balance them dynamically in time slices (bit-identical result)
        const bool narrow = a.narrow && !a.radius;           // one body per lane: twice the warps, no slicing needed
        const bool sliced = !faithful && !narrow && s->slice > 0 && nsteps >= 2 * s->slice &&
                            s->progress_len <= (long long)s->sm_count * 4 * 4;    // < 4 warps per SM sub-partition
        if (sliced) {
            CU(cudaMemsetAsync(s->d_queue, 0, sizeof(unsigned long long), s->stream));
            CU(cudaMemsetAsync(s->d_progress, 0, sizeof(int) * s->progress_len, s->stream));
            CU(launch_ens_step_sliced(a, s->slice, s->d_queue, s->d_progress, s->sm_count, s->stream));
            ++s->launches;
        } else if (nsteps > 0) {
            CU(launch_ens_step(a, faithful, s->stream));
            ++s->launches;
        }
    } else {
        a.nsteps = 1;
        int64_t k = 0;
        if (nsteps >= kEnsGraphSteps) {
            if (!s->step_graph) { int rc = ens_build_graph(s); if (rc) return rc; }
            for (; k + kEnsGraphSteps <= nsteps; k += kEnsGraphSteps) {
                CU(cudaGraphLaunch(s->step_graph, s->stream));
                s->launches += kEnsGraphSteps;
            }
        }
        const int64_t k0 = k;
        for (; k < nsteps; ++k) {
            a.first = k == k0;                   // the call (and every graph) starts / ends synchronised
            a.last = k == nsteps - 1;
            a.pdl = (k > k0 && s->use_pdl) ? 1 : 0;
            CU(launch_ens_step(a, faithful, s->stream));
            ++s->launches;
        }
    }
    return ORB_OK;
}

int orb_ens_download(orb_ensemble* s, double* x, double* y, double* z, double* vx, double* vy, double* vz) {
    LOCK(s);
    const size_t nb = sizeof(double) * s->a.nsys * s->a.nb;
    cudaStream_t st = s->stream;
    if (x) CU(cudaMemcpyAsync(x, s->a.x, nb, cudaMemcpyDeviceToHost, st));
    if (y) CU(cudaMemcpyAsync(y, s->a.y, nb, cudaMemcpyDeviceToHost, st));
    if (z) CU(cudaMemcpyAsync(z, s->a.z, nb, cudaMemcpyDeviceToHost, st));
    if (vx) CU(cudaMemcpyAsync(vx, s->a.vx, nb, cudaMemcpyDeviceToHost, st));
    if (vy) CU(cudaMemcpyAsync(vy, s->a.vy, nb, cudaMemcpyDeviceToHost, st));
    if (vz) CU(cudaMemcpyAsync(vz, s->a.vz, nb, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orb_ens_energy(orb_ensemble* s, double* E_per_system) {
    LOCK(s);
    if (!E_per_system) return fail(ORB_ERR_INVALID, "null array");
    CU(launch_ens_energy(s->a, s->d_E, s->stream));
    ++s->launches;
    CU(cudaMemcpyAsync(E_per_system, s->d_E, sizeof(double) * s->a.nsys, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return ORB_OK;
}

int orb_ens_synchronize(orb_ensemble* s) {
    LOCK(s);
    CU(cudaStreamSynchronize(s->stream));
    return ORB_OK;
}

int orb_ens_launch_count(orb_ensemble* s, int64_t* launches) {
    LOCK(s);
    if (launches) *launches = s->launches;
    return ORB_OK;
}

}  // extern "C"
