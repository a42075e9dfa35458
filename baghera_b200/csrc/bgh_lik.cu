// bgh_lik.cu — the hot path: fused per-SNP likelihood + analytic gradient for a batch of chains.
//
// Replaces the Theano graph PyMC3 generates from ref: baghera_tool/gene_regression.py:77-98
// (gather beta[cat] -> Elemwise{Composite} -> Sum -> AdvancedIncSubtensor1 scatter-add)
// and ref: baghera_tool/heritability.py:44-55 (genome-wide model) with one streaming pass.
//
// Mapping (sm_100a, no tensor cores: the work is FP64-pipe / HBM bound, not a contraction):
//   * SNPs are sorted by gene on upload.  A *group* of CL lanes owns one contiguous SNP range and
//     walks it sequentially; each lane of the group carries CPL chains.  Because every lane of a
//     group sees the same SNP, a gene boundary is group-uniform: the per-gene residual sum is a
//     plain sequential fp64 accumulation that is flushed with one coalesced store when the gene id
//     changes — the "segmented reduction" needs no shuffles and no atomics and is bit-reproducible.
//   * the SoA (chi2 fp32, ld fp32, gene_id int32) stream is read with 128-bit non-allocating loads,
//     converted once per warp (fp32->fp64, w = 1/l) and staged in warp-private shared memory, from
//     which all lanes broadcast-read it (2 x LDS.128 per SNP) — the conversion cost is amortised
//     over CL*CPL chains.  The next chunk's global loads are in flight while the current one is
//     being consumed (register double-buffering).
//   * per (SNP, chain) the inner loop is 6 FP64 instructions:
//         d = s - e;  r = fma(-l, c*beta, d);  acc_g += r;  t = r*w;  acc_e += t;  acc_ll = fma(r, t, acc_ll)
//   * ranges that start or end inside a gene (NonCoding spans ~45% of all SNPs) write "slot"
//     partials; bgh_finalize sums the slots of each such gene in a fixed order.
#include <cuda_runtime.h>
#include <math.h>

#include "bgh_internal.h"

namespace bgh {

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ int4 ldg_stream_i4(const int* p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
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
// log(sigmoid(x)) and log(1 - sigmoid(x)) without cancellation
__device__ __forceinline__ double log_sigmoid_d(double x)
{
    return x >= 0.0 ? -log1p(exp(-x)) : x - log1p(exp(x));
}
__device__ __forceinline__ double log1m_sigmoid_d(double x)
{
    return x >= 0.0 ? -x - log1p(exp(-x)) : -log1p(exp(x));
}

// gw model with a "fixed" intercept: e = Uniform(0.9999, 1.0000001), interval-transformed
// (ref: baghera_tool/heritability.py:45-46)
constexpr double kGwFixLo = 0.9999;
constexpr double kGwFixHi = 1.0000001;

struct __align__(16) SnpA {
    double s;    // chi2
    double nl;   // -ld
};
struct __align__(16) SnpB {
    double w;    // 1/ld
    int g;       // gene id
    int pad;
};

size_t lik_smem_bytes()
{
    const size_t stage = (size_t)kLikWarps * kChunk * (sizeof(SnpA) + sizeof(SnpB));
    return stage;  // the end-of-kernel reduction reuses the staging area
}

// ---------------------------------------------------------------------------------------------
// K1: fused likelihood + gradient
// ---------------------------------------------------------------------------------------------
template <int CL, int CPL>
__global__ void __launch_bounds__(kLikThreads, 1) lik_kernel(const LikParams p)
{
    constexpr int SL = 32 / CL;     // groups per warp
    constexpr int CH = 4 * CL;      // SNPs per group per chunk (SL * CH == kChunk)
    constexpr int CG = CL * CPL;    // chains per chain block
    static_assert(SL * CH == kChunk, "chunk geometry");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    SnpA* const smA_all = reinterpret_cast<SnpA*>(smem_raw);
    SnpB* const smB_all = reinterpret_cast<SnpB*>(smem_raw + (size_t)kLikWarps * kChunk * sizeof(SnpA));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int grp = lane / CL;
    const int lc = lane % CL;
    const unsigned gmask = (CL == 32) ? 0xffffffffu : (((1u << CL) - 1u) << (grp * CL));
    const int v = (blockIdx.x * kLikWarps + warp) * SL + grp;
    const int Cp = p.Cp;
    const int ch0 = blockIdx.y * CG + lc;   // chain of slot k is ch0 + k*CL

    SnpA* const smA = smA_all + warp * kChunk;
    SnpB* const smB = smB_all + warp * kChunk;

    // ---- per-chain constants ----------------------------------------------------------------
    double e[CPL], mi[CPL], iW2[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int ch = ch0 + k * CL;
        if (p.model == BGH_MODEL_GENE_NORMAL) {
            const double xW = p.q[0 * Cp + ch];
            const double ev = p.q[1 * Cp + ch];
            const double xm = p.q[2 * Cp + ch];
            iW2[k] = exp(-2.0 * xW);
            mi[k] = sigmoid_d(xm);
            e[k] = p.fix ? 1.0 : ev;                 // ref: gene_regression.py:85-90
        } else {
            const double x0 = p.q[0 * Cp + ch];
            const double xm = p.q[1 * Cp + ch];
            iW2[k] = 0.0;
            mi[k] = sigmoid_d(xm);
            e[k] = p.fix ? (kGwFixLo + (kGwFixHi - kGwFixLo) * sigmoid_d(x0)) : x0;
        }
    }

    double b[CPL], db[CPL], bn[CPL];
    double ag[CPL], ae[CPL], al[CPL], s1[CPL], s2[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        b[k] = db[k] = bn[k] = 0.0;
        ag[k] = ae[k] = al[k] = s1[k] = s2[k] = 0.0;
    }

    const int4 desc = p.groups[v];
    const int ra = desc.x, rb = desc.y, flags = desc.z;

    if (ra < rb) {
        const bool head_partial = (flags & kHeadPartial) != 0;
        const bool tail_partial = (flags & kTailPartial) != 0;
        const double* const beta = p.q + (size_t)3 * Cp;   // gene model rows 3..
        const int g_first = __ldg(p.gid + ra);
        int cur = g_first;

        auto load_beta = [&](int g, double (&dst)[CPL]) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                if (p.model == BGH_MODEL_GENE_NORMAL)
                    dst[k] = (g < p.G) ? beta[(size_t)g * Cp + ch0 + k * CL] : 0.0;
                else
                    dst[k] = mi[k];
            }
        };
        auto set_beta = [&](const double (&src)[CPL]) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                b[k] = p.c * src[k];
                db[k] = src[k] - mi[k];
            }
        };
        // gene `cur` is complete within this range (or the range ends): emit its sum
        auto flush = [&](bool final_flush) {
            const bool to_head = (cur == g_first) && head_partial;
            const bool to_tail = !to_head && final_flush && tail_partial;
            if (to_head || to_tail) {
                double* dst = p.slots + (size_t)(to_head ? v : p.NG + v) * Cp;
#pragma unroll
                for (int k = 0; k < CPL; ++k) dst[ch0 + k * CL] = ag[k];
            } else if (p.model == BGH_MODEL_GENE_NORMAL) {
                double* dst = p.grad + (size_t)(3 + cur) * Cp;
#pragma unroll
                for (int k = 0; k < CPL; ++k) dst[ch0 + k * CL] = fma(p.c, ag[k], -db[k] * iW2[k]);
            }
            if (!to_head) {   // this group holds the gene's first SNP: it owns the prior terms
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    s1[k] += db[k];
                    s2[k] = fma(db[k], db[k], s2[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < CPL; ++k) ag[k] = 0.0;
        };

        {
            double tmp[CPL];
            load_beta(cur, tmp);
            set_beta(tmp);
            load_beta(cur + 1, bn);
        }

        // chunk loop; `base` is CH aligned so every lane's 4-vector is 16-byte aligned
        int base = ra & ~(CH - 1);
        float4 rs = ldg_stream_f4(p.chi2 + base + lc * 4);
        float4 rl = ldg_stream_f4(p.ld + base + lc * 4);
        int4 rg = ldg_stream_i4(p.gid + base + lc * 4);
        for (; base < rb; base += CH) {
            // convert + stage: SNP i of this group lives at index i*SL + grp
            {
                const float sv[4] = {rs.x, rs.y, rs.z, rs.w};
                const float lv[4] = {rl.x, rl.y, rl.z, rl.w};
                const int gv[4] = {rg.x, rg.y, rg.z, rg.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int idx = (lc * 4 + t) * SL + grp;
                    const double l = (double)lv[t];
                    SnpA a;
                    a.s = (double)sv[t];
                    a.nl = -l;
                    SnpB bb;
                    bb.w = 1.0 / l;
                    bb.g = gv[t];
                    bb.pad = 0;
                    smA[idx] = a;
                    smB[idx] = bb;
                }
            }
            __syncwarp(gmask);
            if (base + CH < rb) {
                rs = ldg_stream_f4(p.chi2 + base + CH + lc * 4);
                rl = ldg_stream_f4(p.ld + base + CH + lc * 4);
                rg = ldg_stream_i4(p.gid + base + CH + lc * 4);
            }
            const int i0 = max(ra - base, 0);
            const int i1 = min(rb - base, CH);
#pragma unroll 4
            for (int i = i0; i < i1; ++i) {
                const SnpA a = smA[i * SL + grp];
                const SnpB bb = smB[i * SL + grp];
                if (bb.g != cur) {
                    flush(false);
                    if (bb.g == cur + 1) {
                        set_beta(bn);
                    } else {
                        double tmp[CPL];
                        load_beta(bb.g, tmp);
                        set_beta(tmp);
                    }
                    cur = bb.g;
                    load_beta(cur + 1, bn);   // prefetch: genes are consecutive after the sort
                }
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const double d = a.s - e[k];
                    const double r = fma(a.nl, b[k], d);
                    ag[k] += r;
                    const double t = r * bb.w;
                    ae[k] += t;
                    al[k] = fma(r, t, al[k]);
                }
            }
            __syncwarp(gmask);
        }
        flush(true);
    }

    // ---- CTA reduction of the four per-chain scalars (fixed order -> deterministic) ----------
    // 1) across the SL groups of a warp (same lc): xor-shuffle tree
#pragma unroll
    for (int o = CL; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            al[k] += __shfl_xor_sync(0xffffffffu, al[k], o);
            ae[k] += __shfl_xor_sync(0xffffffffu, ae[k], o);
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
            s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
        }
    }
    // 2) across warps through shared memory (reuses the staging area)
    __syncthreads();
    double* const red = reinterpret_cast<double*>(smem_raw);   // [kLikWarps][4][CG]
    if (grp == 0) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int cc = lc + k * CL;
            red[(warp * 4 + 0) * CG + cc] = al[k];
            red[(warp * 4 + 1) * CG + cc] = ae[k];
            red[(warp * 4 + 2) * CG + cc] = s1[k];
            red[(warp * 4 + 3) * CG + cc] = s2[k];
        }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 4 * CG; o += blockDim.x) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < kLikWarps; ++w) acc += red[w * 4 * CG + o];
        const int which = o / CG, cc = o % CG;
        p.part[((size_t)blockIdx.x * 4 + which) * Cp + blockIdx.y * CG + cc] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// per-chain closing formulas (priors, transforms, Jacobians) — ref: gene_regression.py:78-81,
// heritability.py:45-49; PyMC3 3.6 distributions/continuous.py + transforms.py
// ---------------------------------------------------------------------------------------------
struct ChainSums {
    double A_ll, A_e, S1, S2;
};

__device__ void finish_chain(const FinParams& p, int ch, const ChainSums& s)
{
    const int Cp = p.Cp;
    if (p.model == BGH_MODEL_GENE_NORMAL) {
        const double xW = p.q[0 * Cp + ch], e = p.q[1 * Cp + ch], xm = p.q[2 * Cp + ch];
        const double iW = exp(-xW), iW2 = exp(-2.0 * xW);
        const double mi = sigmoid_d(xm);
        const double G = (double)p.G_total;
        double lp = (-iW - 2.0 * xW) + xW;                        // InverseGamma(1,1) + log-Jacobian
        lp += -0.5 * (e - 1.0) * (e - 1.0) - 0.5 * kLog2Pi;       // Normal(1,1)
        lp += log_sigmoid_d(xm) + log1m_sigmoid_d(xm);            // Beta(1,1)=0 ; logodds Jacobian
        lp += -0.5 * s.S2 * iW2 - G * xW - 0.5 * G * kLog2Pi;     // Normal(beta | mi, W)
        lp += -0.5 * s.A_ll + p.lik_const;                        // likelihood
        p.logp[ch] = lp;
        p.grad[0 * Cp + ch] = iW - 1.0 + s.S2 * iW2 - G;
        p.grad[1 * Cp + ch] = -(e - 1.0) + (p.fix ? 0.0 : s.A_e);
        p.grad[2 * Cp + ch] = 1.0 - 2.0 * mi + mi * (1.0 - mi) * s.S1 * iW2;
    } else {
        const double x0 = p.q[0 * Cp + ch], xm = p.q[1 * Cp + ch];
        double lp = 0.0, g0;
        if (p.fix) {
            // Uniform(lo,hi) logp = -log(hi-lo); interval Jacobian = log(hi-lo) + log sigmoid'(x0)
            const double sg = sigmoid_d(x0);
            lp += log_sigmoid_d(x0) + log1m_sigmoid_d(x0);
            g0 = s.A_e * (kGwFixHi - kGwFixLo) * sg * (1.0 - sg) + (1.0 - 2.0 * sg);
        } else {
            lp += -0.5 * (x0 - 1.0) * (x0 - 1.0) - 0.5 * kLog2Pi;
            g0 = -(x0 - 1.0) + s.A_e;
        }
        lp += log1m_sigmoid_d(xm) + 0.69314718055994530942;       // Beta(1,2)
        lp += log_sigmoid_d(xm) + log1m_sigmoid_d(xm);            // logodds Jacobian
        lp += -0.5 * s.A_ll + p.lik_const;
        p.logp[ch] = lp;
        p.grad[0 * Cp + ch] = g0;
    }
}

// gradient row of a gene from its total residual sum
__device__ __forceinline__ void finish_gene(const FinParams& p, int gene, int ch, double sum_r)
{
    const int Cp = p.Cp;
    if (p.model == BGH_MODEL_GENE_NORMAL) {
        const double xW = p.q[0 * Cp + ch], xm = p.q[2 * Cp + ch];
        const double iW2 = exp(-2.0 * xW);
        const double mi = sigmoid_d(xm);
        const double bta = p.q[(size_t)(3 + gene) * Cp + ch];
        p.grad[(size_t)(3 + gene) * Cp + ch] = fma(p.c, sum_r, -(bta - mi) * iW2);
    } else {
        const double xm = p.q[1 * Cp + ch];
        const double mi = sigmoid_d(xm);
        // (c*sum_r - 1/(1-mi)) * mi(1-mi) + (1 - 2mi)
        p.grad[1 * Cp + ch] = p.c * sum_r * mi * (1.0 - mi) + 1.0 - 3.0 * mi;
    }
}

// ---------------------------------------------------------------------------------------------
// K2/K3: finalize — CTA 0 reduces the per-CTA scalar partials (and, single rank, closes the
// formulas); CTAs 1..n_long sum the slot partials of one long split gene each (NonCoding spans
// ~45% of all groups); the remaining CTAs handle 32 short split genes each, one warp per gene.
// All sums run in a fixed order -> bit-reproducible.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFinThreads) finalize_kernel(const FinParams p)
{
    __shared__ double sm[32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cp = p.Cp;

    if (blockIdx.x == 0) {
        const int n_out = 4 * Cp;
        const int nob = (n_out + 31) / 32;                 // output blocks of 32
        const int wpb = nob >= 32 ? 1 : 32 / nob;          // warps cooperating per block
        for (int ob0 = 0; ob0 < nob; ob0 += 32 / wpb) {
            const int ob = ob0 + warp / wpb;
            const int sub = warp % wpb;
            const int o = ob * 32 + lane;
            double acc = 0.0;
            if (ob < nob && o < n_out) {
#pragma unroll 4
                for (int cta = sub; cta < p.n_cta; cta += wpb) acc += p.part[(size_t)cta * n_out + o];
            }
            sm[warp][lane] = acc;
            __syncthreads();
            if (sub == 0 && ob < nob && o < n_out) {
                double tot = 0.0;
                for (int w = 0; w < wpb; ++w) tot += sm[warp + w][lane];
                p.payload[o] = tot;   // rows 0..3 of the payload are [4][Cp]
            }
            __syncthreads();
        }
        if (p.finish_inline) {
            __threadfence_block();
            __syncthreads();
            for (int ch = threadIdx.x; ch < Cp; ch += blockDim.x) {
                ChainSums s;
                s.A_ll = p.payload[0 * Cp + ch];
                s.A_e = p.payload[1 * Cp + ch];
                s.S1 = p.payload[2 * Cp + ch];
                s.S2 = p.payload[3 * Cp + ch];
                finish_chain(p, ch, s);
            }
        }
        return;
    }

    if ((int)blockIdx.x > p.n_long) {
        // short lists (<= kShortSplit slots): one warp per split gene, lanes = chains
        const int k = p.n_long + ((int)blockIdx.x - 1 - p.n_long) * 32 + warp;
        if (k >= p.n_split) return;
        const int s0 = p.split_ptr[k], s1 = p.split_ptr[k + 1];
        const int gene = p.split_gene[k], row = p.split_row[k];
        for (int cb = 0; cb < Cp; cb += 32) {
            const int ch = cb + lane;
            if (ch >= Cp) break;
            double tot = 0.0;
            for (int i = s0; i < s1; ++i) tot += p.slots[(size_t)p.split_slots[i] * Cp + ch];
            if (row >= 0)
                p.payload[(size_t)(4 + row) * Cp + ch] = tot;
            else
                finish_gene(p, gene, ch, tot);
        }
        return;
    }

    // long lists: one CTA per split gene, warp w takes slots w, w+32, ...; fixed-order combine
    const int k = blockIdx.x - 1;
    const int s0 = p.split_ptr[k], s1 = p.split_ptr[k + 1];
    const int gene = p.split_gene[k], row = p.split_row[k];
    for (int cb = 0; cb < Cp; cb += 32) {
        const int ch = cb + lane;
        double acc = 0.0;
        if (ch < Cp) {
#pragma unroll 4
            for (int i = s0 + warp; i < s1; i += 32) acc += p.slots[(size_t)p.split_slots[i] * Cp + ch];
        }
        sm[warp][lane] = acc;
        __syncthreads();
        if (warp == 0 && ch < Cp) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < 32; ++w) tot += sm[w][lane];
            if (row >= 0)
                p.payload[(size_t)(4 + row) * Cp + ch] = tot;   // rank-shared: finished after all-reduce
            else
                finish_gene(p, gene, ch, tot);
        }
        __syncthreads();
    }
}

// multi-rank: after the payload all-reduce
__global__ void finish_kernel(const FinParams p)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.Cp) return;
    const int Cp = p.Cp;
    ChainSums s;
    s.A_ll = p.payload[0 * Cp + ch];
    s.A_e = p.payload[1 * Cp + ch];
    s.S1 = p.payload[2 * Cp + ch];
    s.S2 = p.payload[3 * Cp + ch];
    finish_chain(p, ch, s);
    if (p.first_gene_row >= 0) finish_gene(p, 0, ch, p.payload[(size_t)(4 + p.first_gene_row) * Cp + ch]);
    if (p.last_gene_row >= 0 && !(p.G_local == 1 && p.first_gene_row >= 0))
        finish_gene(p, p.G_local - 1, ch, p.payload[(size_t)(4 + p.last_gene_row) * Cp + ch]);
}

// ---------------------------------------------------------------------------------------------
// layout helpers: chain-major [C][D] (PyMC3 / host) <-> gene-major [D][Cp] (device)
// ---------------------------------------------------------------------------------------------
__global__ void transpose_in_kernel(const double* __restrict__ q_cd, double* __restrict__ q_dc, int C, int Cp, int D)
{
    __shared__ double tile[32][33];
    const int d0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r;
        const int d = d0 + threadIdx.x;
        if (c >= C) c = 0;   // padded chains replicate chain 0
        tile[r][threadIdx.x] = (d < D && c0 + r < Cp) ? q_cd[(size_t)c * D + d] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int d = d0 + r, c = c0 + threadIdx.x;
        if (d < D && c < Cp) q_dc[(size_t)d * Cp + c] = tile[threadIdx.x][r];
    }
}

__global__ void transpose_out_kernel(const double* __restrict__ x_dc, double* __restrict__ x_cd, int C, int Cp, int D)
{
    __shared__ double tile[32][33];
    const int d0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int d = d0 + r, c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (d < D && c < Cp) ? x_dc[(size_t)d * Cp + c] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, d = d0 + threadIdx.x;
        if (d < D && c < C) x_cd[(size_t)c * D + d] = tile[threadIdx.x][r];
    }
}

// ---------------------------------------------------------------------------------------------
// FP64 pipe probe: 8 independent DFMA chains per thread, no memory traffic
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
    double a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999, c = 1e-9 * (blockIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) out[0] = r;   // never true; keeps the chains alive
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <int CL, int CPL>
static cudaError_t launch_lik_t(const LikParams& p, int n_cta, int chain_blocks, cudaStream_t s)
{
    dim3 grid(n_cta, chain_blocks);
    lik_kernel<CL, CPL><<<grid, kLikThreads, lik_smem_bytes(), s>>>(p);
    return cudaGetLastError();
}

template <int CL, int CPL>
static cudaError_t configure_t()
{
    return cudaFuncSetAttribute(lik_kernel<CL, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)lik_smem_bytes());
}

cudaError_t lik_configure()
{
    cudaError_t e;
    if ((e = configure_t<4, 1>()) != cudaSuccess) return e;
    if ((e = configure_t<8, 1>()) != cudaSuccess) return e;
    if ((e = configure_t<16, 1>()) != cudaSuccess) return e;
    if ((e = configure_t<32, 1>()) != cudaSuccess) return e;
    if ((e = configure_t<32, 2>()) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t launch_lik(const LikParams& p, int CL, int CPL, int n_cta, int chain_blocks, cudaStream_t s)
{
    if (CL == 4 && CPL == 1) return launch_lik_t<4, 1>(p, n_cta, chain_blocks, s);
    if (CL == 8 && CPL == 1) return launch_lik_t<8, 1>(p, n_cta, chain_blocks, s);
    if (CL == 16 && CPL == 1) return launch_lik_t<16, 1>(p, n_cta, chain_blocks, s);
    if (CL == 32 && CPL == 1) return launch_lik_t<32, 1>(p, n_cta, chain_blocks, s);
    if (CL == 32 && CPL == 2) return launch_lik_t<32, 2>(p, n_cta, chain_blocks, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_finalize(const FinParams& p, cudaStream_t s)
{
    const int n_short = p.n_split - p.n_long;
    finalize_kernel<<<1 + p.n_long + (n_short + 31) / 32, kFinThreads, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_finish(const FinParams& p, cudaStream_t s)
{
    const int t = 128;
    finish_kernel<<<(p.Cp + t - 1) / t, t, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_transpose_in(const double* q_cd, double* q_dc, int C, int Cp, int D, cudaStream_t s)
{
    dim3 grid((D + 31) / 32, (Cp + 31) / 32), block(32, 8);
    transpose_in_kernel<<<grid, block, 0, s>>>(q_cd, q_dc, C, Cp, D);
    return cudaGetLastError();
}

cudaError_t launch_transpose_out(const double* x_dc, double* x_cd, int C, int Cp, int D, cudaStream_t s)
{
    dim3 grid((D + 31) / 32, (Cp + 31) / 32), block(32, 8);
    transpose_out_kernel<<<grid, block, 0, s>>>(x_dc, x_cd, C, Cp, D);
    return cudaGetLastError();
}

cudaError_t launch_dfma_probe(double* out, int iters, int n_cta, cudaStream_t s)
{
    dfma_probe_kernel<<<n_cta, 256, 0, s>>>(out, iters);
    return cudaGetLastError();
}

}  // namespace bgh
