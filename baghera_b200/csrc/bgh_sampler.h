// Internal declarations of the batched NUTS kernels (bgh_sampler.cu) used by bgh_api.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "baghera_b200.h"
#include "bgh_peer.h"

namespace bgh {

constexpr int kNutsMaxDepth = BGH_NUTS_MAX_DEPTH;
constexpr int kNutsMaxPartials = 1 + 2 * kNutsMaxDepth;

// vector slots of the state tensor [n_slots][D][Cp] (public numbering: include/baghera_b200.h)
enum : int {
    kSlotQP = BGH_NUTS_SLOT_Q, kSlotGP = BGH_NUTS_SLOT_GRAD, kSlotVar = BGH_NUTS_SLOT_VAR,
    kSlotQL = 3, kSlotPL, kSlotGL, kSlotQR, kSlotPR, kSlotGR, kSlotQW, kSlotPW, kSlotGW, kSlotQS, kSlotGS,
    kSlotPSumTree, kSlotPSumRun,
    kSlotFgMean = BGH_NUTS_SLOT_FG_MEAN, kSlotFgRaw = BGH_NUTS_SLOT_FG_RAW,
    kSlotBgMean = BGH_NUTS_SLOT_BG_MEAN, kSlotBgRaw = BGH_NUTS_SLOT_BG_RAW,
    kSlotZ,             // async machine: standard normals of the chain's NEXT momentum refresh
    kSlotCount = BGH_NUTS_N_SLOTS
};
static_assert(kSlotPSumRun + 1 == kSlotFgMean, "slot numbering");
static_assert(kSlotZ + 1 == kSlotCount, "slot numbering");

// per-chain scalars sc[k][Cp]
enum : int {
    kScEps = BGH_NUTS_SC_EPS, kScLogpProp = BGH_NUTS_SC_LOGP, kScEnergyProp, kScE0, kScLogSizeTree, kScLogSizeSub,
    kScLogpSub, kScEnergySub, kScAcceptSum, kScNProp, kScMaxDE,
    kScDaLogStep, kScDaLogBar, kScDaHbar, kScDaCount, kScDaMu,
    kScCount = BGH_NUTS_N_SC
};
static_assert(kScDaMu + 1 == kScCount, "scalar numbering");

// per-chain flags fl[k][Cp]
enum : int {
    kFlDir = 0, kFlActive, kFlDone, kFlSubValid, kFlTake, kFlTakeTop, kFlMerge, kFlDepth, kFlDiverging, kFlTurning,
    kFlBadStart,
    kFlCount = BGH_NUTS_N_FL
};
static_assert(kFlBadStart + 1 == kFlCount, "flag numbering");

struct NutsParams {
    double* state;      // [kSlotCount][D][Cp]
    double* ckpt;       // [2][kNutsMaxDepth][D][Cp]   {p, running p_sum} at the even leaves
    double* sc;         // [kScCount][Cp]
    int* fl;            // [kFlCount][Cp]
    double* partials;   // [n_vec_blocks][kNutsMaxPartials][Cp]
    double* sums;       // [kNutsMaxPartials][Cp]
    const double* logp_w;   // [Cp] logp at the working position (written by the likelihood path)
    int* n_active;      // device scalar
    int D, Cp, C;
    int n_vec_blocks;
    // gene-block sharded runs: local rows [0, head_rows) are the shared parameters, replicated on
    // every rank; local row d >= head_rows is global row d + coord_offset.  A replicated row enters the
    // per-chain sums (kinetic energy, U-turn dots) on ONE rank only: the head rows on rank 0 (own_head), the
    // first gene row on the lower rank when the gene continues from it (own_first = 0).  Single rank:
    // head_rows = 0, coord_offset = 0, own_* = 1, peer.world <= 1.
    int coord_offset, head_rows, own_head, own_first;
    const int* gcoord;  // optional [D]: global coordinate of every local row (bgh_set_row_coords), or NULL
    int chain_offset;   // index of this context's chain 0 among all chains of the analysis (Philox key)
    PeerParams peer;    // sums are all-reduced over peer memory inside the scalar kernels when peer.world > 1
    int* status;        // device status word (peer time-out), may be NULL on a single rank
    unsigned long long seed;
    long long draw;
    double emax, target_accept, da_gamma, da_kappa, da_t0;
};

cudaError_t nuts_launch_begin(const NutsParams& p, cudaStream_t s);
cudaError_t nuts_launch_dir(const NutsParams& p, cudaStream_t s);
cudaError_t nuts_launch_leap1(const NutsParams& p, int first, cudaStream_t s);
cudaError_t nuts_launch_leap2(const NutsParams& p, int leaf, int store_level, int lvl_min, int lvl_max, cudaStream_t s);
cudaError_t nuts_launch_end_doubling(const NutsParams& p, int max_depth, cudaStream_t s);
cudaError_t nuts_launch_adapt_mass(const NutsParams& p, double fg_w, double bg_w, int switch_window, cudaStream_t s);
cudaError_t nuts_launch_finish(const NutsParams& p, int tune, double* stats, cudaStream_t s);
cudaError_t nuts_launch_stop_tuning(const NutsParams& p, cudaStream_t s);
cudaError_t nuts_launch_init(const NutsParams& p, double step0, cudaStream_t s);
constexpr int kNutsLaunchesPerLeaf = 3;

// ---- asynchronous per-chain state machine (bgh_nuts_run): control flags afl[k][Cp], scalars asc[k][Cp]
constexpr int kAsyncPartials = 2 * kNutsMaxDepth + 4;   // kinetic, 2*depth sub-block dots, 2 whole-tree dots, kinetic0
constexpr int kAqTree0 = 1 + 2 * kNutsMaxDepth;          // whole-tree dots live at [kAqTree0], [kAqTree0+1]
constexpr int kAqKin0 = kAqTree0 + 2;
enum : int {
    kAfPhase = 0, kAfDepth, kAfLeaf, kAfDir, kAfDirPrev, kAfTake, kAfTakeTop, kAfMerge, kAfNew, kAfFirst, kAfRecSlot,
    kAfAdapt, kAfSwitch, kAfStoreLevel, kAfLvlMin, kAfNLvl, kAfLastLeaf, kAfDraw, kAfDiverging, kAfTurning, kAfNAdapt,
    kAfBadStart, kAfSubValid,
    kAfCount = BGH_NUTS_ASYNC_N_FL
};
static_assert(kAfSubValid + 1 <= kAfCount, "async flag numbering");
enum : int { kPhaseRun = 0, kPhaseFlush = 1, kPhaseFinished = 2 };
enum : int {   // extra scalars of the async machine, appended to the lock-step ones
    kScH = kScCount, kScFgW, kScBgW, kScFgUse, kScBgUse,
    kAscCount = BGH_NUTS_ASYNC_N_SC
};
static_assert(kScBgUse + 1 <= kAscCount, "async scalar numbering");
// device timestamps (%globaltimer >> 10, i.e. units of 1.024 us): start of the run (set by the init kernel) and the end of
// every chain's last tuning transition — seconds_tune of sampler.py is MEASURED per chain, not apportioned by leapfrogs
constexpr int kScRunStartUs = BGH_NUTS_ASYNC_N_SC - 1;     // double: exact
constexpr int kAfTuneEndUs = BGH_NUTS_ASYNC_N_FL - 1;      // int32: low 32 bits (differences are valid for 73 minutes)

struct AsyncParams {
    NutsParams n;        // n.sc has kAscCount rows, n.partials kAsyncPartials quantities per block
    int* afl;            // [kAfCount][Cp]
    void* trace;         // [draws][D][C] fp64 or fp32
    double* stats;       // [tune+draws][BGH_NUTS_N_STATS][C]
    int* n_finished;     // device scalar
    int trace_f32;
    int tune, draws;
    int early_draws, early_depth, max_depth, window;
};
cudaError_t nuts_async_launch_init(const AsyncParams& p, double step0, double initial_weight, cudaStream_t s, bool fill_z = true);   // leap1, leap2, leaf scalar (plus the likelihood + finalize)

}  // namespace bgh
