// force_sym.h -- host interface of the pair-symmetric fast force kernel (force_sym.cu)
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "kernels.h"

namespace orb {

struct SymItem {
    int I;        // I-block index
    int t0, t1;   // tile range [t0, t1)
    int chunk;    // which P_i plane this item writes
    int slot;     // which P_j plane of the current panel this item writes
};

struct SymPanel {
    int ka = 0, kb = 0;       // owned I-blocks I = rank + world*k, k in [ka, kb)
    int n_items = 0;
    SymItem* d_items = nullptr;
};

struct SymPlan {
    bool valid = false;
    int ti = 4;
    long long n = 0;
    int rank = 0, world = 1;  // snake-order ownership of I-blocks (multi-GPU); 0/1 on one GPU
    long long B = 0;          // bodies per I-block (128 * ti)
    int nb_I = 0, n_tiles = 0;
    int tile = 256;           // bodies per source tile (64 / 128 / 256)
    int chunk_tiles = 0, n_chunks = 0;
    int panel_blocks = 0;
    int ctas_per_sm = 0;
    double* Pi = nullptr;     // [n_chunks][3][n]
    double* Pj = nullptr;     // [panel_blocks][3][n]
    std::vector<SymPanel> panels;
};

cudaError_t plan_sym(SymPlan& p, long long n, int sm_count, int rank, int world);
void free_sym(SymPlan& p);
// What the reduction launch does besides summing the partial planes (unsharded engines with one panel only):
enum SymTailMode {
    kSymTailNone = 0,       // accelerations only
    kSymTailKick,           // + second half-kick (device-resolved contacts follow in their own launches)
    kSymTailClose,          // + second half-kick, history append, step bookkeeping
    kSymTailCloseNext       // + the first half-kick and drift of the next step (multi-step graphs)
};
bool sym_tail_applicable(const SymPlan& p);
cudaError_t launch_force_sym(const DeviceState& s, const StepParams& sp, const SymPlan& p, bool detect,
                             cudaStream_t st, int* launches, int tail_mode = kSymTailNone);
bool sym_uniform(const SymPlan& p, const StepParams& sp);
const char* sym_kernel_name(int ti, bool detect, bool uniform);

}  // namespace orb
