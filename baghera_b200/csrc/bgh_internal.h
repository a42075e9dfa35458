// Internal declarations shared by the translation units of libbaghera_b200.so.
// Not part of the public ABI (see include/baghera_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "baghera_b200.h"

namespace bgh {

constexpr int kLikThreads = 512;               // one CTA per SM, 16 warps
constexpr int kLikWarps = kLikThreads / 32;
constexpr int kChunk = 128;                    // SNPs staged per warp per chunk
#ifndef BGH_LIK_BATCH
#define BGH_LIK_BATCH 8
#endif
constexpr int kLikBatch = BGH_LIK_BATCH;       // SNPs per branch-free inner-loop batch
constexpr int kFinThreads = 1024;
constexpr int kFinConstChains = 128;           // chains closed per pass of finalize CTA 0
constexpr int kShortSplit = 32;                // split lists up to this length are summed by one warp
constexpr double kLog2Pi = 1.8378770664093454835606594728112;

enum : int { kHeadPartial = 1, kTailPartial = 2, kExclusive = 4, kOwner = 8 };

// Work decomposition of the gene-sorted SNP array (host side, no CUDA needed).
// A "group" is CL consecutive lanes of a warp that walk one contiguous SNP range
// sequentially; gene boundaries are therefore group-uniform and per-gene sums are plain
// sequential fp64 accumulations (deterministic, no atomics).
struct Plan {
    int n_groups = 0;
    int n_cta = 0;
    std::vector<int32_t> off;          // [n_groups+1] range offsets
    std::vector<int32_t> flags;        // [n_groups]   kHeadPartial | kTailPartial
    // genes whose sum is assembled from slot partials: CSR over slot indices.
    // slot index s <  n_groups              : head slot of group s
    // slot index n_groups <= s < 2 n_groups : tail slot of group s - n_groups
    // slot index s >= 2 n_groups            : CTA slot of likelihood CTA s - 2 n_groups (exclusive CTAs)
    int n_long = 0;                    // split genes with > kShortSplit slots come first
    std::vector<int32_t> split_gene;   // [ns]
    std::vector<int32_t> split_row;    // [ns] payload row (>=0) if the gene is rank-shared, else -1
    std::vector<int32_t> split_ptr;    // [ns+1]
    std::vector<int32_t> split_slots;  // [split_ptr[ns]]
};

// gid must be sorted ascending.  Returns 0 or BGH_EINVAL (message in err).
// `align` > 1: every range boundary that is not a gene boundary is rounded down to a multiple of `align`
// (the upload pads every gene to a multiple of the kernel's batch, so gene boundaries already are).
int build_plan(int64_t M, const int32_t* gid, int32_t G, int n_cta, int groups_per_cta,
               bool first_partial, bool last_partial, int first_row, int last_row, double kappa,
               Plan* plan, std::string* err, int align = 1, int n_replicated = 0, bool owns_replicated = true, double snap = 0.0,
               double kappa_warp = -1.0);   // >= 0: the cuts between the warps of a CTA use this boundary cost (sampler plan)

struct LikParams {
    const double* sp;     // prepared SoA, gene sorted: s / sqrt(l)
    const double* wp;     //                            1 / sqrt(l)
    const double* lp;     //                            sqrt(l)
    const int* gid;
    const int4* groups;   // {a, b, flags, gene id of SNP a}
    const double* q;      // element (row d, chain c) at q[d * ds + c * cs]: [D][Cp] (ds = Cp, cs = 1) for the
    double* grad;         // evaluation API, chain-major [Cp][D] (ds = 1, cs = D) inside the tick kernel
    long long ds, cs;
    const int* woff;      // optional [Cp]: per-chain element offset added to every q / grad access (sampler kernel), or NULL
    double* slots;        // [2*NG + nCTA][Cp]
    double* part;         // [nCTA][4][Cp]
    double* cc;           // [2][Cp] per-chain constants {1/W^2, mi} written for finalize
    unsigned long long* trace;   // optional [chain_blocks*nCTA*kLikWarps][8] phase timestamps, or NULL
    int Cp;
    int G;
    int NG;
    int n_cta;
    int model;
    int fix;
    double c;
    double gamma_rate;    // gamma model: rate of the Gamma prior (N_1kG), gene_regression.py:231
};

struct FinParams {
    const double* q;      // strides as in LikParams
    double* grad;
    long long ds, cs;
    double* logp;
    const double* slots;
    const double* part;
    const double* cc;     // [2][Cp] from the likelihood kernel
    double* payload;      // [(4+n_shared)][Cp]
    const int* split_gene;
    const int* split_row;
    const int* split_ptr;
    const int* split_slots;
    int n_split;
    int n_long;           // split genes [0, n_long) have > kShortSplit slots (one CTA each)
    int n_cta;            // CTAs of the likelihood kernel (rows of part)
    int Cp;
    int G_total;
    int model;
    int fix;
    int finish_inline;    // 1: single rank, write logp / grad rows 0..2 here
    double c;
    double lik_const;
    double gamma_rate;
    // finish-only
    int first_gene_row, last_gene_row, G_local;
    int n_shared_rows;    // payload rows 4.. : a row of a rank-shared gene this rank does NOT hold is zeroed every launch
    int n_replicated;     // genes [0, n_replicated) are held by every rank: payload row 4 + j <-> gene j
};

// kernel launchers (bgh_lik.cu)
cudaError_t launch_lik(const LikParams& p, int CL, int CPL, int n_cta, int chain_blocks, cudaStream_t s);
cudaError_t launch_finalize(const FinParams& p, cudaStream_t s);
cudaError_t launch_finish(const FinParams& p, cudaStream_t s);
cudaError_t launch_transpose_in(const double* q_cd, double* q_dc, int C, int Cp, int D, cudaStream_t s);
cudaError_t launch_transpose_out(const double* x_dc, double* x_cd, int C, int Cp, int D, cudaStream_t s);
cudaError_t launch_dfma_probe(double* out, int iters, int n_cta, cudaStream_t s);
cudaError_t lik_configure();   // opt-in dynamic shared memory, once per process
size_t lik_smem_bytes();

}  // namespace bgh
