// Internal declarations of the fused persistent NUTS kernel (bgh_tick2.cu) used by bgh_api.cu.
#pragma once

#include <cuda_runtime.h>

#include "bgh_fused.h"
#include "bgh_internal.h"
#include "bgh_peer.h"
#include "bgh_sampler.h"

namespace bgh {

// per-chain quantities every CTA hands to the chain's scalar CTA each tick:
//   [0, kT2Lik)            likelihood sums  {sum r^2/l, sum r/l, S1, S2}
//   [kT2Lik, kT2Nq)        the sampler's sums, numbered as the kAq* quantities of bgh_sampler.h
//                          (kinetic, 2 x 10 sub-block U-turn dots, 2 whole-tree dots, kinetic of fresh momenta)
constexpr int kT2Lik = 4;
constexpr int kT2Nq = kT2Lik + kAsyncPartials;

// per-chain command word written by the scalar phase for the vector work of the NEXT tick
enum : int {
    kCbRun = 1 << 0,         // phase == kPhaseRun
    kCbPlain = 1 << 1,       // running chain without rare-event work this tick
    kCbTake = 1 << 2,        // the last leaf was drawn as the proposal of its subtree
    kCbMerge = 1 << 3,       // the last leaf completed a valid doubling
    kCbTakeTop = 1 << 4,     // ... whose proposal becomes the tree's proposal
    kCbNew = 1 << 5,         // a transition starts this tick
    kCbAdapt = 1 << 7,
    kCbSwitch = 1 << 8,
    kCbFlush = 1 << 9,       // phase == kPhaseFlush
    kCbLast = 1 << 10,       // last leaf of the doubling: whole-tree U-turn dots
    kCbW = 1 << 12,          // index of the chain's working buffer (X0 / X1)
    kCbFlip = 1 << 13,       // the index was flipped for this doubling (the previous leaf lives in the other buffer)
    kCbStoreShift = 16,      // 4 bits: store_level + 1 (0 = no checkpoint store)
    kCbLvlMinShift = 20,     // 4 bits: first checkpoint level checked
    kCbNLvlShift = 24,       // 4 bits: checkpoint levels checked
};
// extra per-chain flags of the fused kernel (rows of afl[kAfCount][Cp])
enum : int { kAfW = kAfSubValid + 1, kAfFlip };
static_assert(kAfFlip + 1 <= kAfCount, "tick2 flag numbering");

// extra per-chain scalars of the fused kernel (rows of sc[kAscCount][Cp])
enum : int { kScKin0LateLocal = kScBgUse + 1, kScKin0LateRep };
static_assert(kScKin0LateRep < kScRunStartUs && kAfSubValid + 2 < kAfTuneEndUs, "the fused kernel's extra slots stay below the timestamp slots");
static_assert(kScKin0LateRep + 1 <= kAscCount, "tick2 scalar numbering");

// LL-protocol mailbox of the per-chain exchange: cell = {lo32(data), tag, hi32(data), tag}
struct T2Peer {
    uint4* box[kMaxPeers];   // box[r]: rank r's mailbox [2 parities][world][Cp][n_pay]   (box[rank] is local)
    int rank, world;
    int n_pay;               // doubles per chain and rank: kT2Nq + replicated genes
    unsigned epoch_base;     // tag of tick t of this run = epoch_base + t + 1 (never 0)
    int send_only;           // tests (bgh_debug_peer_mode): deposit the payload in the peers' mailboxes, do not wait for theirs
};

struct Tick2Params {
    LikParams lik;           // q / grad = buffer X0 of the state (gene-major [D][Cp]); woff = the chains' buffer offsets
    FinParams fin;           // closing formulas: q / grad = buffer X0 (the scalar phase adds the chain's offset), logp = logp_w
    AsyncParams a;
    unsigned long long* bar; // grid barrier word, zeroed by the host before every launch
    int* status;
    int chain_blocks;
    int n_ticks;
    int* ticks_done;
    int first_launch;        // 1: the late rows' first half step has not been taken yet
    int tick_base;           // global index of this launch's first tick
    int row0;                // head rows in front of the per-gene rows (3 normal, 2 gamma, 2 = D genome-wide)
    int n_unowned;           // sharded: genes [0, n_unowned) are replicated and counted by rank 0 only
    int* cmd_bits;           // [Cp]
    double* cmd_h;           // [Cp]
    int* cmd_woff;           // [Cp] element offset of the chain's working buffer (== lik.woff)
    double* part;            // [n_cta][kAsyncPartials][Cp] sampler sums of the rows each CTA owns (likelihood sums: lik.part)
    const int* cta_row_ptr;  // [n_cta + 1] CSR over cta_row_gene
    const int* cta_row_gene; // gene ids of the rows each CTA owns (genes complete inside one of its SNP ranges)
    int trace_cta;           // tools: the CTA whose phase stamps are recorded
    int trace_first;         // tools: first traced tick
    double* k6;              // optional [3][D][Cp]: sum beta, sum beta^2, count(beta - mi > 0) of the recorded draws
    double* head_trace;      // optional [draws][C][row0]: the shared parameters of every recorded draw, fp64
    T2Peer peer;
    unsigned long long* phase_trace;   // tools: [64 ticks][16] globaltimer stamps of CTA trace_cta, or NULL
};

// sharded evaluation with the per-chain peer exchange (contexts with replicated genes)
struct Eval2Params {
    LikParams lik;
    FinParams fin;
    unsigned long long* bar;
    unsigned long long bar_base;
    int* status;
    int chain_blocks;
    int C;
    T2Peer peer;
    unsigned epoch;          // exchange counter of the evaluation mailbox
};
cudaError_t launch_eval2(const Eval2Params& e, int CL, int CPL, int n_cta, cudaStream_t s);

cudaError_t tick2_configure();
size_t tick2_smem_bytes();
cudaError_t launch_tick2(const Tick2Params& t, int CL, int CPL, int n_cta, cudaStream_t s);

}  // namespace bgh
