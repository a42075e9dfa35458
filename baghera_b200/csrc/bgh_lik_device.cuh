// bgh_lik_device.cuh — device code of the hot path shared by the stand-alone likelihood kernel
// (bgh_lik.cu) and the fused / persistent kernels (bgh_fused.cu).
//
// lik_phase<CL, CPL>: fused per-SNP likelihood + analytic gradient for a batch of chains.
// Replaces the Theano graph PyMC3 generates from ref: baghera_tool/gene_regression.py:77-98
// (gather beta[cat] -> Elemwise{Composite} -> Sum -> AdvancedIncSubtensor1 scatter-add)
// and ref: baghera_tool/heritability.py:44-55 (genome-wide model) with one streaming pass.
//
// Mapping (sm_100a, no tensor cores: the work is FP64-pipe / HBM bound, not a contraction):
//   * SNPs are sorted by gene on upload and stored as a *prepared* fp64 SoA
//         sp_j = s_j / sqrt(l_j)     wp_j = 1 / sqrt(l_j)     lp_j = sqrt(l_j)     gid_j (int32)
//     so that the standardised residual of SNP j for a chain with intercept e and slope b = c*beta is
//         rho = (s - e - l b) / sqrt(l) = sp - e wp - b lp                       (2 DFMA)
//     and the three sums every evaluation needs are plain FMA accumulations (3 DFMA):
//         sum r^2 / l = sum rho^2        sum r / l = sum rho wp        sum_{j in g} r = sum rho lp
//     => 5 FP64 instructions per (SNP, chain); nothing is converted inside the kernel.
//   * A WARP owns one contiguous SNP range and walks it sequentially in batches of 8 SNPs (genes are padded to
//     whole batches, so a gene boundary only occurs between batches and is warp-uniform).  Lanes are laid out as
//     SL = 32 / CL "SNP lanes" x CL chain lanes, each lane carrying CPL chains: with 64 chains (CL = 32, CPL = 2)
//     every lane sees every SNP; with few chains (C = 4: CL = 4, SL = 8) the 8 SNPs of a batch are spread over the
//     8 SNP lanes of each chain — all 32 lanes work, and the range count does not grow with 32 / CL (a range per
//     CL lanes, the round-1 mapping, cut the 1.1M SNPs into 19k ranges at C = 4: nearly every gene was split over
//     several ranges and its sum re-assembled from slots).  The per-gene residual sum is a sequential fp64
//     accumulation per lane, folded over the SNP lanes by a fixed xor-shuffle tree when the gene ends and flushed
//     with one coalesced store: no atomics, bit-reproducible.
//   * Staging: one elected lane per warp moves each 128-SNP chunk of the four arrays global -> shared with
//     bulk asynchronous copies (cp.async.bulk, the 1-D TMA path) that complete on a warp-private
//     mbarrier; two stages, so the copy of chunk k+2 is issued as soon as chunk k is consumed.  All lanes then
//     read the chunk from shared memory (CL = 32: broadcast LDS.128, 2 SNPs per load).
//   * ranges that start or end inside a gene (NonCoding spans ~45% of all SNPs) write "slot"
//     partials; the finalize step sums the slots of each such gene in a fixed order.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "bgh_internal.h"

namespace bgh {

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk asynchronous copy (PTX ISA 8.x, sm_90+)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// The SNP stream is read exactly once per evaluation while the kernel is FP64-bound: mark it evict-first
// in L2 so that it does not push the sampler state (which IS re-read every tick) out of the 126 MB L2.
__device__ __forceinline__ unsigned long long l2_evict_first_policy()
{
    unsigned long long pol;
#ifdef BGH_SNP_EVICT_NORMAL
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
#else
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
    return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, unsigned long long pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ double sigmoid_d(double x)
{
    // stable for both signs
    if (x >= 0.0) {
        const double z = exp(-x);
        return 1.0 / (1.0 + z);
    }
    const double z = exp(x);
    return z / (1.0 + z);
}

// gw model with a "fixed" intercept: e = Uniform(0.9999, 1.0000001), interval-transformed
// (ref: baghera_tool/heritability.py:45-46)
constexpr double kGwFixLo = 0.9999;
constexpr double kGwFixHi = 1.0000001;

// Shared-memory staging, per warp and stage, structure of arrays: sp[kStride], wp[kStride],
// lp[kStride] (fp64) and gid[kStrideG].  Group grp of a warp owns elements [grp*(CH+2), +CH): the +2
// doubles of padding shift consecutive groups by 16 bytes so that the per-group broadcast LDS.128 of
// the SL groups fall into different banks (and keep every bulk-copy destination 16-byte aligned).
constexpr int kStages = 2;
constexpr int kStride = kChunk + 2 * 8;    // fp64 arrays (the slack keeps the three arrays of a stage in different banks)
constexpr int kStrideG = kChunk + 4 * 8;   // gene ids
constexpr size_t kWarpStageBytes = (size_t)kStride * 3 * sizeof(double) + (size_t)kStrideG * sizeof(int);
static_assert(kWarpStageBytes % 16 == 0, "warp staging area must keep 16-byte alignment");
constexpr size_t kLikStageBytes = (size_t)kLikWarps * kStages * kWarpStageBytes;
constexpr size_t kLikBarBytes = (size_t)kLikWarps * kStages * 8 * sizeof(unsigned long long);   // [warp][stage][group]
constexpr size_t kLikSmemBytes = kLikStageBytes + kLikBarBytes;
static_assert((size_t)kLikWarps * 5 * 64 * sizeof(double) <= kLikStageBytes, "epilogue reduction reuses the staging area");

// Once per kernel, before the first lik_phase: initialise the group-private mbarriers.
// The caller follows with a CTA barrier.
__device__ __forceinline__ void lik_stage_init(unsigned char* smem_raw)
{
    const uint32_t bars = smem_u32(smem_raw + kLikStageBytes);
    for (int i = threadIdx.x; i < kLikWarps * kStages * 8; i += blockDim.x) mbar_init(bars + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned sm_id()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
// optional per-warp phase timestamps (tools/trace_lik.py); p.trace is NULL in production
#define BGH_TRACE(slot)                                                                        \
    do {                                                                                       \
        if (p.trace != nullptr && lane == 0)                                                   \
            p.trace[((size_t)(chain_block * p.n_cta + cta) * kLikWarps + warp) * 8 + (slot)] = global_timer_ns(); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// K1: fused likelihood + gradient for likelihood CTA `cta` (0 .. n_cta-1) and chain block
// `chain_block`.  `par` carries the mbarrier phase parities of this thread's group across calls
// (bit s = parity the next wait on stage s expects); start it at 0 right after lik_stage_init.
// Must be called by all kLikThreads threads of the CTA (contains CTA barriers).
// ---------------------------------------------------------------------------------------------
template <int CL, int CPL, bool GAMMA>
__device__ __forceinline__ void lik_phase(const LikParams& p, unsigned char* smem_raw, const int cta, const int chain_block,
                                          uint32_t& par)
{
    constexpr int SL = 32 / CL;     // SNP lanes per chain lane
    constexpr int CH = kChunk;      // SNPs per chunk
    constexpr int CG = CL * CPL;    // chains per chain block
    constexpr int B = kLikBatch;    // SNPs per branch-free batch
    static_assert(CH % B == 0, "batch must divide the chunk");
    static_assert(SL <= B && B % SL == 0, "the SNP lanes share a batch evenly");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int grp = lane / CL;      // SNP lane: takes SNPs grp, grp + SL, ... of every batch
    const int lc = lane % CL;
    constexpr unsigned gmask = 0xffffffffu;
    const int v = cta * kLikWarps + warp;
    const int Cp = p.Cp;
    const int ch0 = chain_block * CG + lc;   // chain of slot k is ch0 + k*CL
    constexpr bool gamma_model = GAMMA;     // compile-time: keeps exp() and the extra selects out of the normal model's code
    const bool gene_model = p.model == BGH_MODEL_GENE_NORMAL || gamma_model;
    const int row0 = gamma_model ? 2 : 3;   // first per-gene row of q / grad

    unsigned char* const warp_stage = smem_raw + (size_t)warp * kStages * kWarpStageBytes;
    const uint32_t bar0 = smem_u32(smem_raw + kLikStageBytes) + 8u * (uint32_t)((warp * kStages) * 8);   // stage s: + 64 s
    BGH_TRACE(0);
    if (p.trace != nullptr && lane == 0)
        p.trace[((size_t)(chain_block * p.n_cta + cta) * kLikWarps + warp) * 8 + 7] = sm_id();

    // ---- preamble: every global load is issued before the first dependent use ------------------
    // (cold L2: descriptor -> first beta rows is a dependent chain of two HBM round trips; the shared rows of q
    // and the exp()/sigmoid() of the per-chain constants run underneath it)
    const int4 desc = p.groups[v];
    const long long ds = p.ds, cs = p.cs;
    double q0[CPL], q1[CPL], q2[CPL];
    // sampler kernel (bgh_tick2.cu): every chain reads its position from / writes its gradient to one of two
    // role-swapped buffers; wo[k] is the chain's element offset into q / grad (0 when p.woff is NULL)
    long long wo[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int ch = ch0 + k * CL;
        wo[k] = p.woff != nullptr ? (long long)p.woff[ch] : 0ll;
        q0[k] = p.q[0 * ds + ch * cs + wo[k]];
        q1[k] = p.q[1 * ds + ch * cs + wo[k]];
        q2[k] = (gene_model && !gamma_model) ? p.q[2 * ds + ch * cs + wo[k]] : 0.0;
    }
    const int ra = desc.x, rb = desc.y, flags = desc.z, g_first = desc.w;
    const bool exclusive = (flags & kExclusive) != 0;
    const bool active = ra < rb;
    const double* const beta = p.q + (long long)row0 * ds + (long long)ch0 * cs;   // per-gene rows (this lane's chains)

    const int base0 = ra & ~(CH - 1);   // CH aligned: every bulk copy source is 16-byte aligned
    const int n_chunks = active ? (rb - base0 + CH - 1) / CH : 0;
    const unsigned long long pol = l2_evict_first_policy();
    auto issue = [&](int k) {   // lane lc == 0 of the group
        const int st = k & (kStages - 1);
        const size_t b0 = (size_t)base0 + (size_t)k * CH;
        const uint32_t bar = bar0 + 64u * st;
        unsigned char* const sb = warp_stage + (size_t)st * kWarpStageBytes;
        const uint32_t d_s = smem_u32(reinterpret_cast<double*>(sb));
        const uint32_t d_g = smem_u32(reinterpret_cast<int*>(reinterpret_cast<double*>(sb) + 3 * kStride));
        mbar_expect_tx(bar, CH * 28);
        bulk_g2s(d_s, p.sp + b0, CH * 8, bar, pol);
        bulk_g2s(d_s + kStride * 8, p.wp + b0, CH * 8, bar, pol);
        bulk_g2s(d_s + 2 * kStride * 8, p.lp + b0, CH * 8, bar, pol);
        bulk_g2s(d_g, p.gid + b0, CH * 4, bar, pol);
    };
    if (lane == 0) {
        if (n_chunks > 0) issue(0);
        if (n_chunks > 1) issue(1);
    }

    double braw[CPL], bn[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) braw[k] = bn[k] = 0.0;
    if (active && gene_model) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            braw[k] = beta[(long long)g_first * ds + (long long)(k * CL) * cs + wo[k]];
            bn[k] = (g_first + 1 < p.G) ? beta[(long long)(g_first + 1) * ds + (long long)(k * CL) * cs + wo[k]] : 0.0;
        }
    }

    // ---- per-chain constants ----------------------------------------------------------------
    double ne[CPL], mi[CPL], iW2[CPL];   // ne = -e
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        if (gamma_model) {
            // ref: gene_regression.py:229-248 — q = [e, mi_logodds__, beta_log__...]
            iW2[k] = 0.0;
            mi[k] = sigmoid_d(q1[k]);
            ne[k] = p.fix ? -1.0 : -q0[k];           // ref: gene_regression.py:234-240
        } else if (gene_model) {
            iW2[k] = exp(-2.0 * q0[k]);
            mi[k] = sigmoid_d(q2[k]);
            ne[k] = p.fix ? -1.0 : -q1[k];           // ref: gene_regression.py:85-90
        } else {
            iW2[k] = 0.0;
            mi[k] = sigmoid_d(q1[k]);
            ne[k] = p.fix ? -(kGwFixLo + (kGwFixHi - kGwFixLo) * sigmoid_d(q0[k])) : -q0[k];
            braw[k] = bn[k] = mi[k];                 // the single slope of heritability.py:50-55
        }
    }

    if (cta == 0 && warp == 0 && grp == 0) {   // hand the per-chain constants to finalize (one SNP lane suffices)
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            p.cc[0 * Cp + ch0 + k * CL] = iW2[k];
            p.cc[1 * Cp + ch0 + k * CL] = mi[k];
        }
    }

    double nb[CPL], db[CPL];   // nb = -c * beta
    double ag[CPL], ae[CPL], al[CPL], s1[CPL], s2[CPL], ax[CPL];
#pragma unroll
    // normal model: slope = c * beta_g,             db = beta_g - mi   (prior Normal(mi, W))
    // gamma model : slope = c * exp(beta_log_g),    db = beta_log_g    (prior Gamma(mi, rate) on exp(.))
    for (int k = 0; k < CPL; ++k) {
        nb[k] = -p.c * (gamma_model ? exp(braw[k]) : braw[k]);
        db[k] = gamma_model ? braw[k] : braw[k] - mi[k];
        ag[k] = ae[k] = al[k] = s1[k] = s2[k] = ax[k] = 0.0;
    }
    const double neg_inv_c = -1.0 / p.c;

    BGH_TRACE(1);
    if (active) {
        const bool head_partial = (flags & kHeadPartial) != 0;
        const bool tail_partial = (flags & kTailPartial) != 0;
        int cur = g_first;

        auto load_beta = [&](int g, double (&dst)[CPL]) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                if (gene_model)
                    dst[k] = (g < p.G) ? beta[(long long)g * ds + (long long)(k * CL) * cs + wo[k]] : 0.0;
                else
                    dst[k] = mi[k];
            }
        };
        // gene `cur` is complete within this range (or the range ends): emit its sum
        auto flush = [&](bool final_flush) {
            bool owner;
            if (SL > 1 && !exclusive) {     // fold the SNP lanes' partial sums of the gene (fixed xor tree)
#pragma unroll
                for (int o = CL; o < 32; o <<= 1)
#pragma unroll
                    for (int k = 0; k < CPL; ++k) ag[k] += __shfl_xor_sync(0xffffffffu, ag[k], o);
            }
            if (exclusive) {
                // every group of this CTA accumulates the same gene: reduced in the CTA epilogue
#pragma unroll
                for (int k = 0; k < CPL; ++k) ax[k] = ag[k];
                owner = (flags & kOwner) != 0;
            } else {
                const bool to_head = (cur == g_first) && head_partial;
                const bool to_tail = !to_head && final_flush && tail_partial;
                if (grp != 0) {
                    // the other SNP lanes hold copies of the folded sum: SNP lane 0 stores
                } else if (to_head || to_tail) {
                    double* dst = p.slots + (size_t)(to_head ? v : p.NG + v) * Cp + ch0;
#pragma unroll
                    for (int k = 0; k < CPL; ++k) dst[k * CL] = ag[k];
                } else if (gamma_model) {
                    // d/d beta_log = beta (c sum r - rate) + mi       (prior + log-Jacobian + likelihood)
                    double* dst = p.grad + (long long)(row0 + cur) * ds + (long long)ch0 * cs;
#pragma unroll
                    for (int k = 0; k < CPL; ++k) dst[(long long)(k * CL) * cs + wo[k]] = fma(nb[k] * neg_inv_c, fma(p.c, ag[k], -p.gamma_rate), mi[k]);
                } else if (gene_model) {
                    double* dst = p.grad + (long long)(row0 + cur) * ds + (long long)ch0 * cs;
#pragma unroll
                    for (int k = 0; k < CPL; ++k) dst[(long long)(k * CL) * cs + wo[k]] = fma(p.c, ag[k], -db[k] * iW2[k]);
                }
                owner = !to_head;   // this group holds the gene's first SNP: it owns the prior terms
            }
            if (owner && grp == 0) {    // (the epilogue adds the SNP lanes: the prior terms are counted by SNP lane 0)
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    s1[k] += db[k];
                    s2[k] = gamma_model ? s2[k] + nb[k] * neg_inv_c : fma(db[k], db[k], s2[k]);   // sum beta / sum (beta-mi)^2
                }
            }
        };
        // switch to gene g (its raw beta in src), prefetch the row after it
        auto enter_gene = [&](int g, const double (&src)[CPL]) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                nb[k] = -p.c * (gamma_model ? exp(src[k]) : src[k]);
                db[k] = gamma_model ? src[k] : src[k] - mi[k];
            }
            cur = g;
            load_beta(g + 1, bn);   // genes are consecutive after the sort
        };
        // Operand order matters: the FP64 pipe issues a DFMA every 2 cycles only if at most TWO of its source operands
        // have to be read from the register file (tools/probe_rf.cu: three fresh 64-bit operands -> 3 cycles, 67% of the
        // DFMA peak); an operand the warp's previous instruction used in the same slot comes from the operand reuse
        // cache.  ptxas schedules SNP-major with the chains of a lane adjacent, so the sums are written quantity-major
        // over the chains (al, ae, ag): 40 of the 80 DFMAs of a batch read three operands instead of 52 with the
        // chain-major order (tools/sass_dfma_reads.py on the built kernel; each accumulator still sees its terms in
        // the same order, results are bit-identical).
        auto accumulate = [&](const double spv, const double wpv, const double lpv) {
            double rho[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) rho[k] = fma(nb[k], lpv, fma(ne[k], wpv, spv));
#pragma unroll
            for (int k = 0; k < CPL; ++k) al[k] = fma(rho[k], rho[k], al[k]);
#pragma unroll
            for (int k = 0; k < CPL; ++k) ae[k] = fma(rho[k], wpv, ae[k]);
#pragma unroll
            for (int k = 0; k < CPL; ++k) ag[k] = fma(rho[k], lpv, ag[k]);
        };

        for (int kc = 0; kc < n_chunks; ++kc) {
            const int st = kc & (kStages - 1);
            const int base = base0 + kc * CH;
            unsigned char* const sb = warp_stage + (size_t)st * kWarpStageBytes;
            // SNP i (0 <= i < CH) of this group lives at index i of the group's slice of each array
            const double* const sm_s = reinterpret_cast<const double*>(sb);
            const double* const sm_w = sm_s + kStride;
            const double* const sm_l = sm_w + kStride;
            const int* const sm_g = reinterpret_cast<const int*>(reinterpret_cast<const double*>(sb) + 3 * kStride);
            mbar_wait(bar0 + 64u * st, (par >> st) & 1u);
            par ^= 1u << st;

            // The upload pads every gene to a multiple of B SNPs (zero-weight padding) and the plan cuts
            // ranges at multiples of B: a batch never straddles a gene boundary, so the only boundary work
            // is "flush the finished gene, switch the slope" in front of a batch.
            const int i1 = min(rb - base, CH);
            for (int i = max(ra - base, 0); i < i1; i += B) {
                const int g_i = sm_g[i];
                if (g_i != cur) {                                   // group-uniform
                    flush(false);
#pragma unroll
                    for (int k = 0; k < CPL; ++k) ag[k] = 0.0;
                    double tmp[CPL];
                    if (g_i == cur + 1) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) tmp[k] = bn[k];
                    } else {
                        load_beta(g_i, tmp);
                    }
                    enter_gene(g_i, tmp);
                }
                if constexpr (SL == 1) {
                    double2 s2v[B / 2], w2v[B / 2], l2v[B / 2];
#pragma unroll
                    for (int t = 0; t < B / 2; ++t) {
                        s2v[t] = *reinterpret_cast<const double2*>(sm_s + i + 2 * t);
                        w2v[t] = *reinterpret_cast<const double2*>(sm_w + i + 2 * t);
                        l2v[t] = *reinterpret_cast<const double2*>(sm_l + i + 2 * t);
                    }
#pragma unroll
                    for (int t = 0; t < B / 2; ++t) {
                        accumulate(s2v[t].x, w2v[t].x, l2v[t].x);
                        accumulate(s2v[t].y, w2v[t].y, l2v[t].y);
                    }
                } else {
                    // SNP lane grp takes SNPs grp, grp + SL, ... of the batch (consecutive doubles across the SNP lanes)
                    double sv[B / SL], wv[B / SL], lv[B / SL];
#pragma unroll
                    for (int t = 0; t < B / SL; ++t) {
                        sv[t] = sm_s[i + grp + t * SL];
                        wv[t] = sm_w[i + grp + t * SL];
                        lv[t] = sm_l[i + grp + t * SL];
                    }
#pragma unroll
                    for (int t = 0; t < B / SL; ++t) accumulate(sv[t], wv[t], lv[t]);
                }
            }
            __syncwarp(gmask);
            if (lane == 0 && kc + kStages < n_chunks) issue(kc + kStages);
        }
        flush(true);
    }

    BGH_TRACE(2);
    // ---- CTA reduction of the per-chain scalars (fixed order -> deterministic) ---------------
    // 1) across the SL groups of a warp (same lc): xor-shuffle tree
#pragma unroll
    for (int o = CL; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            al[k] += __shfl_xor_sync(0xffffffffu, al[k], o);
            ae[k] += __shfl_xor_sync(0xffffffffu, ae[k], o);
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
            s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
            ax[k] += __shfl_xor_sync(0xffffffffu, ax[k], o);
        }
    }
    // 2) across warps through shared memory (reuses the staging area: every copy has landed)
    __syncthreads();
    BGH_TRACE(3);
    double* const red = reinterpret_cast<double*>(smem_raw);   // [kLikWarps][5][CG]
    if (grp == 0) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int cc = lc + k * CL;
            red[(warp * 5 + 0) * CG + cc] = al[k];
            red[(warp * 5 + 1) * CG + cc] = ae[k];
            red[(warp * 5 + 2) * CG + cc] = s1[k];
            red[(warp * 5 + 3) * CG + cc] = s2[k];
            red[(warp * 5 + 4) * CG + cc] = ax[k];
        }
    }
    __syncthreads();
    const bool cta_exclusive = (p.groups[cta * kLikWarps].z & kExclusive) != 0;
    for (int o = threadIdx.x; o < 5 * CG; o += blockDim.x) {
        const int which = o / CG, cc = o % CG;
        if (which == 4 && !cta_exclusive) continue;
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < kLikWarps; ++w) acc += red[w * 5 * CG + o];
        if (which < 4)
            p.part[((size_t)cta * 4 + which) * Cp + chain_block * CG + cc] = acc;
        else
            p.slots[((size_t)2 * p.NG + cta) * Cp + chain_block * CG + cc] = acc;
    }
    // the staging area is handed back to the async proxy (next lik_phase of a persistent kernel)
    __syncthreads();
    fence_proxy_async();
    BGH_TRACE(4);
}

}  // namespace bgh
