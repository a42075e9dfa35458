// bgh_sampler.cu — batched-chain NUTS transition on the device (K4 of SURVEY §2).
//
// Replaces `pm.sample(..., nuts_kwargs=dict(target_accept=0.90))` at
// ref: baghera_tool/gene_regression.py:100-106 / baghera_tool/heritability.py:56-62.
// Semantics follow PyMC3 3.6 [recalled, SURVEY §8a]: step_methods/hmc/{nuts,integration,
// quadpotential,step_size}.py —
//   * leapfrog: p += eps/2 g ; q += eps * (var . p) ; (logp, g) = f(q) ; p += eps/2 g ;
//     energy = 1/2 p.(var . p) - logp ; one f call per leapfrog
//   * multinomial NUTS: each doubling picks a direction with prob 1/2 and builds a subtree of 2^depth
//     leaves; inside a subtree the proposal is drawn with prob. proportional to exp(-dE) (here by a
//     sequential reservoir draw, equal in distribution to the recursive logbern(log_size2 - log_size)
//     merge); at the top level the new subtree's proposal is taken with prob. min(1, w_new / w_old);
//     U-turn when p_sum . v_left <= 0 or p_sum . v_right <= 0 for every aligned sub-block and for
//     the whole tree; divergence when |E - E0| >= 1000; a diverging / turning subtree is discarded;
//     accept_stat = sum min(1, exp(E0 - E)) / n_leaves.
//   * all chains advance in lock step (one batched logp+grad per leaf); chains whose tree has
//     stopped are masked.  The recursion is unrolled iteratively: sub-block U-turn checks use
//     per-level checkpoints of (p, running p_sum), so memory is O(max_depth * D) per chain.
//   * dual-averaging step size (gamma .05, kappa .75, t0 10, mu = log(10 eps0)) and the windowed
//     diagonal mass matrix (window 101, initial weight 10) are per chain, as PyMC3 runs one
//     independent adaptation per chain.
// Randomness: counter-based Philox4x32-10 keyed by (seed, chain); counters carry (draw, purpose,
// element), never the rank, so every rank of a sharded run makes identical decisions.
#include <cuda_runtime.h>
#include <math.h>

#include "bgh_internal.h"
#include "bgh_sampler.h"
#include "bgh_peer_device.cuh"
#include "bgh_sampler_device.cuh"

namespace bgh {

// Work distribution of the vector kernels: a CTA owns kRowsPerCta-row strips; thread -> (row-in-strip, chain).
// Per-chain partial sums are reduced inside the CTA in a fixed order and written to
// partials[blockIdx][k][Cp]; the scalar kernels add the CTA partials in block order (deterministic).
constexpr int kVecThreads = 256;

struct VecIdx {
    int chain, r0, rstep;   // this thread handles rows r0, r0+rstep, ... of chain `chain`
    bool on;
};
__device__ __forceinline__ VecIdx vec_index(const NutsParams& p)
{
    VecIdx v;
    const int tpr = p.Cp < kVecThreads ? p.Cp : kVecThreads;   // threads along the chain axis
    const int rpb = kVecThreads / tpr;                        // rows per CTA pass
    v.chain = (threadIdx.x % tpr) + blockIdx.y * tpr;
    v.r0 = blockIdx.x * rpb + threadIdx.x / tpr;
    v.rstep = gridDim.x * rpb;
    v.on = v.chain < p.Cp;
    return v;
}

// sharded runs: global coordinate of local row d, and whether this rank counts row d in the per-chain sums
__device__ __forceinline__ unsigned global_coord(const NutsParams& p, int d)
{
    if (p.gcoord != nullptr) return (unsigned)p.gcoord[d];
    return (unsigned)(d >= p.head_rows ? d + p.coord_offset : d);
}
__device__ __forceinline__ bool row_owned(const NutsParams& p, int d)
{
    return d < p.head_rows ? p.own_head != 0 : (d == p.head_rows ? p.own_first != 0 : true);
}

template <int NK>
__device__ __forceinline__ void cta_reduce_store(const NutsParams& p, const VecIdx& v, const double (&acc)[NK], int nk)
{
    __shared__ double red[kVecThreads];
    const int tpr = p.Cp < kVecThreads ? p.Cp : kVecThreads;
    const int rpb = kVecThreads / tpr;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        if (k >= nk) break;
        red[threadIdx.x] = acc[k];
        __syncthreads();
        if (threadIdx.x < tpr) {
            double t = 0.0;
            for (int r = 0; r < rpb; ++r) t += red[r * tpr + threadIdx.x];
            if (v.on) p.partials[((size_t)blockIdx.x * kNutsMaxPartials + k) * p.Cp + v.chain] = t;
        }
        __syncthreads();
    }
}

// Single-CTA (kScThreads threads) reduction of the CTA partials of quantities 0..nk-1 into
// p.sums[k][Cp]: R row groups load independent strips, then one thread per output adds the R strip
// sums in a fixed order.  Ends with a CTA barrier, after which every thread may read p.sums.
constexpr int kScThreads = 1024;
__device__ void reduce_partials(const NutsParams& p, int nk)
{
    __shared__ double stage[kScThreads];
    const int tid = threadIdx.x;
    const int n_out = nk * p.Cp;
    for (int o0 = 0; o0 < n_out; o0 += kScThreads) {
        const int nb = min(kScThreads, n_out - o0);
        const int R = kScThreads / nb;
        const int oo = tid % nb, sub = tid / nb;
        const int k = (o0 + oo) / p.Cp, c = (o0 + oo) % p.Cp;
        double acc = 0.0;
        if (sub < R) {
#pragma unroll 8
            for (int b = sub; b < p.n_vec_blocks; b += R)
                acc += p.partials[((size_t)b * kNutsMaxPartials + k) * p.Cp + c];
        }
        stage[tid] = acc;
        __syncthreads();
        if (tid < nb) {
            double t = 0.0;
            for (int r = 0; r < R; ++r) t += stage[r * nb + tid];
            p.sums[(size_t)k * p.Cp + c] = t;
        }
        __syncthreads();
    }
    if (p.peer.world > 1) {
        // every rank holds the sums over ITS rows: add the ranks' contributions in rank order (bit-identical
        // totals everywhere, so all ranks take the same tree decisions)
        PeerParams pr = p.peer;
        pr.n = n_out;
        __threadfence();
        peer_allreduce(pr, p.sums, p.status);
    }
}
__device__ __forceinline__ double sum_partials(const NutsParams& p, int k, int c) { return p.sums[(size_t)k * p.Cp + c]; }

// ---------------------------------------------------------------------------------------------
// transition start: momentum refresh, tree initialisation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kVecThreads) nuts_begin_kernel(const NutsParams p)
{
    const VecIdx v = vec_index(p);
    double acc[1] = {0.0};
    if (v.on) {
        const uint2 key = chain_key(p, v.chain);
        double* const P_L = slot(p, kSlotPL);
        double* const P_R = slot(p, kSlotPR);
        double* const PS = slot(p, kSlotPSumTree);
        const double* const VAR = slot(p, kSlotVar);
        const double* const QP = slot(p, kSlotQP);
        const double* const GP = slot(p, kSlotGP);
        double* const QL = slot(p, kSlotQL);
        double* const QR = slot(p, kSlotQR);
        double* const GL = slot(p, kSlotGL);
        double* const GR = slot(p, kSlotGR);
        for (int d = v.r0; d < p.D; d += v.rstep) {
            const size_t i = (size_t)d * p.Cp + v.chain;
            // one normal per (coordinate, chain): Box-Muller on a Philox block keyed by the GLOBAL coordinate
            const double z = momentum_normal(key, p.draw, global_coord(p, d));
            const double var = VAR[i];
            const double mom = z * rsqrt(var);          // p ~ N(0, 1/var)   (quadpotential.random)
            P_L[i] = mom;
            P_R[i] = mom;
            PS[i] = mom;
            const double q = QP[i], g = GP[i];
            QL[i] = q; QR[i] = q;
            GL[i] = g; GR[i] = g;
            if (row_owned(p, d)) acc[0] = fma(0.5 * var * mom, mom, acc[0]);
        }
    }
    cta_reduce_store<1>(p, v, acc, 1);
}

__global__ void __launch_bounds__(kScThreads) nuts_begin_scalar_kernel(const NutsParams p)
{
    reduce_partials(p, 1);
    for (int c = threadIdx.x; c < p.Cp; c += blockDim.x) {
    const double kin = sum_partials(p, 0, c);
    const double e0 = kin - sc(p, kScLogpProp, c);
    sc(p, kScE0, c) = e0;
    sc(p, kScEnergyProp, c) = e0;
    sc(p, kScLogSizeTree, c) = 0.0;
    sc(p, kScAcceptSum, c) = 0.0;
    sc(p, kScNProp, c) = 0.0;
    sc(p, kScMaxDE, c) = 0.0;
    fl(p, kFlDepth, c) = 0;
    fl(p, kFlDone, c) = 0;
    fl(p, kFlDiverging, c) = 0;
    fl(p, kFlTurning, c) = 0;
    fl(p, kFlBadStart, c) = isfinite(e0) ? 0 : 1;
    }
}

// per doubling: direction draw, subtree reset
__global__ void nuts_dir_kernel(const NutsParams p)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.Cp) return;
    const int depth = fl(p, kFlDepth, c);
    const double u = chain_uniform(p, c, kPurposeDir, (unsigned)depth);
    fl(p, kFlDir, c) = (u < 0.5) ? 1 : -1;     // logbern(log 0.5) * 2 - 1
    fl(p, kFlActive, c) = fl(p, kFlDone, c) ? 0 : 1;
    fl(p, kFlSubValid, c) = 1;
    fl(p, kFlTake, c) = 0;
    fl(p, kFlTakeTop, c) = 0;
    fl(p, kFlMerge, c) = 0;
    sc(p, kScLogSizeSub, c) = -INFINITY;
}

// ---------------------------------------------------------------------------------------------
// leapfrog, first half:  [pending proposal copy]  p_half = p + h g ; q += 2h var p_half
// first leaf of a doubling reads the tree edge in the chain's direction instead of the working state
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kVecThreads) nuts_leap1_kernel(const NutsParams p, int first)
{
    const VecIdx v = vec_index(p);
    if (!v.on) return;
    const int c = v.chain;
    if (!fl(p, kFlActive, c)) return;
    const int dir = fl(p, kFlDir, c);
    const double h = 0.5 * sc(p, kScEps, c) * (double)dir;
    const bool take = fl(p, kFlTake, c) != 0;
    const double* const VAR = slot(p, kSlotVar);
    double* const QW = slot(p, kSlotQW);
    double* const PW = slot(p, kSlotPW);
    double* const GW = slot(p, kSlotGW);
    const double* const QE = slot(p, dir > 0 ? kSlotQR : kSlotQL);
    const double* const PE = slot(p, dir > 0 ? kSlotPR : kSlotPL);
    const double* const GE = slot(p, dir > 0 ? kSlotGR : kSlotGL);
    double* const QS = slot(p, kSlotQS);
    double* const GS = slot(p, kSlotGS);
    for (int d = v.r0; d < p.D; d += v.rstep) {
        const size_t i = (size_t)d * p.Cp + c;
        double q, mom, g;
        if (first) {
            q = QE[i]; mom = PE[i]; g = GE[i];
        } else {
            q = QW[i]; mom = PW[i]; g = GW[i];
            if (take) { QS[i] = q; GS[i] = g; }   // the previous leaf was drawn as the subtree proposal
        }
        const double ph = fma(h, g, mom);
        PW[i] = ph;
        QW[i] = fma(2.0 * h * VAR[i], ph, q);
    }
}

// ---------------------------------------------------------------------------------------------
// leapfrog, second half + per-chain reductions:  p = p_half + h g' ; kinetic ; running p_sum ;
// checkpoint stores (even leaf) and U-turn dot products for the listed checkpoint levels (odd leaf)
// ---------------------------------------------------------------------------------------------
template <int NL>
__global__ void __launch_bounds__(kVecThreads) nuts_leap2_kernel(const NutsParams p, int leaf, int store_level,
                                                                 int lvl_min)
{
    const VecIdx v = vec_index(p);
    double acc[1 + 2 * NL];
#pragma unroll
    for (int k = 0; k < 1 + 2 * NL; ++k) acc[k] = 0.0;
    const int c = v.chain;
    const bool on = v.on && fl(p, kFlActive, c);
    if (on) {
        const int dir = fl(p, kFlDir, c);
        const double h = 0.5 * sc(p, kScEps, c) * (double)dir;
        const double* const VAR = slot(p, kSlotVar);
        double* const PW = slot(p, kSlotPW);
        const double* const GW = slot(p, kSlotGW);
        double* const PSR = slot(p, kSlotPSumRun);
        for (int d = v.r0; d < p.D; d += v.rstep) {
            const size_t i = (size_t)d * p.Cp + c;
            const double var = VAR[i];
            const double mom = fma(h, GW[i], PW[i]);
            PW[i] = mom;
            const double vel = var * mom;
            const bool own = row_owned(p, d);
            if (own) acc[0] = fma(0.5 * vel, mom, acc[0]);
            const double ps = (leaf == 0) ? mom : PSR[i] + mom;
            PSR[i] = ps;
            if (store_level >= 0) {
                *ckpt2(p, store_level, d, v.chain) = make_double2(mom, ps);
            }
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                const double2 ck = *ckpt2(p, lvl_min + l, d, v.chain);
                const double pc = ck.x;
                const double sub = ps - ck.y + pc;   // p_sum of the aligned block
                if (own) {
                    acc[1 + 2 * l] = fma(sub, var * pc, acc[1 + 2 * l]);       // . v_left
                    acc[2 + 2 * l] = fma(sub, vel, acc[2 + 2 * l]);            // . v_right
                }
            }
        }
    }
    cta_reduce_store<1 + 2 * NL>(p, v, acc, 1 + 2 * NL);
}

// per-leaf tree logic (one thread per chain)
__global__ void __launch_bounds__(kScThreads) nuts_leaf_scalar_kernel(const NutsParams p, int leaf, int n_lvl)
{
    reduce_partials(p, 1 + 2 * n_lvl);
    for (int c = threadIdx.x; c < p.Cp; c += blockDim.x) {
    fl(p, kFlTake, c) = 0;
    if (!fl(p, kFlActive, c)) continue;
    const double kin = sum_partials(p, 0, c);
    const double logp = p.logp_w[c];
    const double energy = kin - logp;
    double dE = energy - sc(p, kScE0, c);
    if (isnan(dE)) dE = INFINITY;
    if (fabs(dE) > fabs(sc(p, kScMaxDE, c))) sc(p, kScMaxDE, c) = dE;
    sc(p, kScNProp, c) += 1.0;
    if (fabs(dE) < p.emax) {
        sc(p, kScAcceptSum, c) += fmin(1.0, exp(-dE));
        const double lw = -dE;
        bool take;
        if (leaf == 0) {
            take = true;
            sc(p, kScLogSizeSub, c) = lw;
        } else {
            const double ls = logaddexp_d(sc(p, kScLogSizeSub, c), lw);
            const int depth = fl(p, kFlDepth, c);
            const double u = chain_uniform(p, c, kPurposeLeaf, ((unsigned)depth << 16) | (unsigned)leaf);
            take = log(u) < lw - ls;
            sc(p, kScLogSizeSub, c) = ls;
        }
        if (take) {
            fl(p, kFlTake, c) = 1;
            sc(p, kScLogpSub, c) = logp;
            sc(p, kScEnergySub, c) = energy;
        }
        bool turning = false;
        for (int l = 0; l < n_lvl; ++l) {
            const double dl = sum_partials(p, 1 + 2 * l, c), dr = sum_partials(p, 2 + 2 * l, c);
            turning = turning || (dl <= 0.0) || (dr <= 0.0);
        }
        if (turning) {
            fl(p, kFlTurning, c) = 1;
            fl(p, kFlSubValid, c) = 0;
            fl(p, kFlActive, c) = 0;
        }
    } else {
        fl(p, kFlDiverging, c) = 1;
        fl(p, kFlSubValid, c) = 0;
        fl(p, kFlActive, c) = 0;
    }
    }
}

// end of a doubling, scalar part 1: top-level biased progressive sampling
__global__ void nuts_top1_kernel(const NutsParams p)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.Cp) return;
    fl(p, kFlMerge, c) = 0;
    fl(p, kFlTakeTop, c) = 0;
    if (fl(p, kFlDone, c)) return;
    const int depth = fl(p, kFlDepth, c);
    fl(p, kFlDepth, c) = depth + 1;
    if (fl(p, kFlSubValid, c)) {
        const double s1 = sc(p, kScLogSizeTree, c), s2 = sc(p, kScLogSizeSub, c);
        const double u = chain_uniform(p, c, kPurposeTop, (unsigned)depth);
        if (log(u) < s2 - s1) {                      // logbern(size2 - size1)
            fl(p, kFlTakeTop, c) = 1;
            sc(p, kScLogpProp, c) = sc(p, kScLogpSub, c);
            sc(p, kScEnergyProp, c) = sc(p, kScEnergySub, c);
        }
        sc(p, kScLogSizeTree, c) = logaddexp_d(s1, s2);
        fl(p, kFlMerge, c) = 1;
    } else {
        fl(p, kFlDone, c) = 1;                       // diverging / turning subtree: discarded, tree stops
    }
}

// end of a doubling, vector part: pending proposal copy, p_sum merge, edge update, tree proposal,
// whole-tree U-turn dots
__global__ void __launch_bounds__(kVecThreads) nuts_end_doubling_kernel(const NutsParams p)
{
    const VecIdx v = vec_index(p);
    double acc[2] = {0.0, 0.0};
    const int c = v.chain;
    if (v.on && fl(p, kFlMerge, c)) {
        const int dir = fl(p, kFlDir, c);
        const bool take = fl(p, kFlTake, c) != 0, take_top = fl(p, kFlTakeTop, c) != 0;
        const double* const VAR = slot(p, kSlotVar);
        const double* const QW = slot(p, kSlotQW);
        const double* const PW = slot(p, kSlotPW);
        const double* const GW = slot(p, kSlotGW);
        double* const QS = slot(p, kSlotQS);
        double* const GS = slot(p, kSlotGS);
        double* const QP = slot(p, kSlotQP);
        double* const GP = slot(p, kSlotGP);
        double* const PST = slot(p, kSlotPSumTree);
        const double* const PSR = slot(p, kSlotPSumRun);
        double* const QE = slot(p, dir > 0 ? kSlotQR : kSlotQL);
        double* const PE = slot(p, dir > 0 ? kSlotPR : kSlotPL);
        double* const GE = slot(p, dir > 0 ? kSlotGR : kSlotGL);
        const double* const PO = slot(p, dir > 0 ? kSlotPL : kSlotPR);   // the other edge's momentum
        for (int d = v.r0; d < p.D; d += v.rstep) {
            const size_t i = (size_t)d * p.Cp + c;
            const double qw = QW[i], pw = PW[i], gw = GW[i];
            double qs, gs;
            if (take) { qs = qw; gs = gw; QS[i] = qs; GS[i] = gs; }
            if (take_top) {
                if (!take) { qs = QS[i]; gs = GS[i]; }
                QP[i] = qs;
                GP[i] = gs;
            }
            QE[i] = qw; PE[i] = pw; GE[i] = gw;          // the far end of the new subtree becomes the edge
            const double pst = PST[i] + PSR[i];
            PST[i] = pst;
            const double var = VAR[i];
            if (row_owned(p, d)) {
                acc[0] = fma(pst, var * pw, acc[0]);
                acc[1] = fma(pst, var * PO[i], acc[1]);
            }
        }
    }
    cta_reduce_store<2>(p, v, acc, 2);
}

__global__ void __launch_bounds__(kScThreads) nuts_top2_kernel(const NutsParams p, int max_depth)
{
    // single CTA; counts the chains whose tree keeps growing
    __shared__ int n_active;
    if (threadIdx.x == 0) n_active = 0;
    reduce_partials(p, 2);
    for (int c = threadIdx.x; c < p.Cp; c += blockDim.x) {
        if (!fl(p, kFlDone, c)) {
            if (fl(p, kFlMerge, c)) {
                const double d0 = sum_partials(p, 0, c), d1 = sum_partials(p, 1, c);
                if (d0 <= 0.0 || d1 <= 0.0) {
                    fl(p, kFlTurning, c) = 1;
                    fl(p, kFlDone, c) = 1;
                }
            }
            if (fl(p, kFlDepth, c) >= max_depth) fl(p, kFlDone, c) = 1;
            if (!fl(p, kFlDone, c) && c < p.C) atomicAdd(&n_active, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *p.n_active = n_active;
}

// ---------------------------------------------------------------------------------------------
// end of transition: adaptation (tuning only) and statistics
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kVecThreads) nuts_adapt_mass_kernel(const NutsParams p, double fg_w, double bg_w,
                                                                      int switch_window)
{
    // quadpotential._WeightedVariance.add_sample(x, 1) for the foreground and background estimators,
    // var = foreground.raw_var / foreground.w_sum, optional window switch (foreground <- background)
    const VecIdx v = vec_index(p);
    if (!v.on) return;
    const int c = v.chain;
    const double* const QP = slot(p, kSlotQP);
    double* const VAR = slot(p, kSlotVar);
    double* const FM = slot(p, kSlotFgMean);
    double* const FR = slot(p, kSlotFgRaw);
    double* const BM = slot(p, kSlotBgMean);
    double* const BR = slot(p, kSlotBgRaw);
    const double inv_fg = 1.0 / fg_w, inv_bg = 1.0 / bg_w;
    for (int d = v.r0; d < p.D; d += v.rstep) {
        const size_t i = (size_t)d * p.Cp + c;
        const double x = QP[i];
        double fm = FM[i], fr = FR[i], bm = BM[i], br = BR[i];
        welford_add(x, inv_fg, fm, fr);      // fg_w, bg_w are the weights AFTER adding this sample
        welford_add(x, inv_bg, bm, br);
        VAR[i] = fr * inv_fg;
        if (switch_window) {
            FM[i] = bm; FR[i] = br;
            BM[i] = 0.0; BR[i] = 0.0;
        } else {
            FM[i] = fm; FR[i] = fr;
            BM[i] = bm; BR[i] = br;
        }
    }
}

__global__ void nuts_finish_scalar_kernel(const NutsParams p, int tune, double* stats)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.Cp) return;
    const double n = fmax(sc(p, kScNProp, c), 1.0);
    const double accept = sc(p, kScAcceptSum, c) / n;
    // step_size.DualAverageAdaptation.update
    if (tune) {
        const double count = sc(p, kScDaCount, c);
        const double w = 1.0 / (count + p.da_t0);
        const double hbar = (1.0 - w) * sc(p, kScDaHbar, c) + w * (p.target_accept - accept);
        const double log_step = sc(p, kScDaMu, c) - hbar * sqrt(count) / p.da_gamma;
        const double mk = pow(count, -p.da_kappa);
        sc(p, kScDaLogBar, c) = mk * log_step + (1.0 - mk) * sc(p, kScDaLogBar, c);
        sc(p, kScDaHbar, c) = hbar;
        sc(p, kScDaCount, c) = count + 1.0;
        sc(p, kScDaLogStep, c) = log_step;
        sc(p, kScEps, c) = exp(log_step);
    }
    if (stats) {
        stats[0 * p.Cp + c] = (double)fl(p, kFlDepth, c);
        stats[1 * p.Cp + c] = accept;
        stats[2 * p.Cp + c] = sc(p, kScNProp, c);
        stats[3 * p.Cp + c] = (double)fl(p, kFlDiverging, c);
        stats[4 * p.Cp + c] = sc(p, kScEnergyProp, c);
        stats[5 * p.Cp + c] = sc(p, kScEps, c);
        stats[6 * p.Cp + c] = sc(p, kScLogpProp, c);
        stats[7 * p.Cp + c] = sc(p, kScMaxDE, c);
    }
}

// switch from the adapting step size to the averaged one (step_adapt.current(tune=False))
__global__ void nuts_stop_tuning_kernel(const NutsParams p)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.Cp) return;
    sc(p, kScEps, c) = exp(sc(p, kScDaLogBar, c));
}

__global__ void nuts_init_scalar_kernel(const NutsParams p, double step0)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.Cp) return;
    sc(p, kScEps, c) = step0;
    sc(p, kScDaLogStep, c) = log(step0);
    sc(p, kScDaLogBar, c) = log(step0);
    sc(p, kScDaHbar, c) = 0.0;
    sc(p, kScDaCount, c) = 1.0;
    sc(p, kScDaMu, c) = log(10.0 * step0);
    sc(p, kScLogpProp, c) = p.logp_w[c];
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static dim3 vec_grid(const NutsParams& p)
{
    const int tpr = p.Cp < kVecThreads ? p.Cp : kVecThreads;
    return dim3(p.n_vec_blocks, (p.Cp + tpr - 1) / tpr);
}
static dim3 sc_grid(const NutsParams& p) { return dim3((p.Cp + 127) / 128); }

#define LAUNCH_CHECK() return cudaGetLastError()
cudaError_t nuts_launch_begin(const NutsParams& p, cudaStream_t s)
{
    nuts_begin_kernel<<<vec_grid(p), kVecThreads, 0, s>>>(p);
    nuts_begin_scalar_kernel<<<1, kScThreads, 0, s>>>(p);
    LAUNCH_CHECK();
}
cudaError_t nuts_launch_dir(const NutsParams& p, cudaStream_t s)
{
    nuts_dir_kernel<<<sc_grid(p), 128, 0, s>>>(p);
    LAUNCH_CHECK();
}
cudaError_t nuts_launch_leap1(const NutsParams& p, int first, cudaStream_t s)
{
    nuts_leap1_kernel<<<vec_grid(p), kVecThreads, 0, s>>>(p, first);
    LAUNCH_CHECK();
}
template <int NL>
static void leap2_t(const NutsParams& p, int leaf, int store_level, int lvl_min, cudaStream_t s)
{
    nuts_leap2_kernel<NL><<<vec_grid(p), kVecThreads, 0, s>>>(p, leaf, store_level, lvl_min);
}
cudaError_t nuts_launch_leap2(const NutsParams& p, int leaf, int store_level, int lvl_min, int lvl_max, cudaStream_t s)
{
    const int n_lvl = lvl_max >= lvl_min ? lvl_max - lvl_min + 1 : 0;
    switch (n_lvl) {
        case 0: leap2_t<0>(p, leaf, store_level, lvl_min, s); break;
        case 1: leap2_t<1>(p, leaf, store_level, lvl_min, s); break;
        case 2: leap2_t<2>(p, leaf, store_level, lvl_min, s); break;
        case 3: leap2_t<3>(p, leaf, store_level, lvl_min, s); break;
        case 4: leap2_t<4>(p, leaf, store_level, lvl_min, s); break;
        case 5: leap2_t<5>(p, leaf, store_level, lvl_min, s); break;
        case 6: leap2_t<6>(p, leaf, store_level, lvl_min, s); break;
        case 7: leap2_t<7>(p, leaf, store_level, lvl_min, s); break;
        case 8: leap2_t<8>(p, leaf, store_level, lvl_min, s); break;
        case 9: leap2_t<9>(p, leaf, store_level, lvl_min, s); break;
        case 10: leap2_t<10>(p, leaf, store_level, lvl_min, s); break;
        default: return cudaErrorInvalidValue;
    }
    nuts_leaf_scalar_kernel<<<1, kScThreads, 0, s>>>(p, leaf, n_lvl);
    LAUNCH_CHECK();
}
cudaError_t nuts_launch_end_doubling(const NutsParams& p, int max_depth, cudaStream_t s)
{
    nuts_top1_kernel<<<sc_grid(p), 128, 0, s>>>(p);
    nuts_end_doubling_kernel<<<vec_grid(p), kVecThreads, 0, s>>>(p);
    nuts_top2_kernel<<<1, kScThreads, 0, s>>>(p, max_depth);
    LAUNCH_CHECK();
}
cudaError_t nuts_launch_adapt_mass(const NutsParams& p, double fg_w, double bg_w, int switch_window, cudaStream_t s)
{
    nuts_adapt_mass_kernel<<<vec_grid(p), kVecThreads, 0, s>>>(p, fg_w, bg_w, switch_window);
    LAUNCH_CHECK();
}
cudaError_t nuts_launch_finish(const NutsParams& p, int tune, double* stats, cudaStream_t s)
{
    nuts_finish_scalar_kernel<<<sc_grid(p), 128, 0, s>>>(p, tune, stats);
    LAUNCH_CHECK();
}
cudaError_t nuts_launch_stop_tuning(const NutsParams& p, cudaStream_t s)
{
    nuts_stop_tuning_kernel<<<sc_grid(p), 128, 0, s>>>(p);
    LAUNCH_CHECK();
}
cudaError_t nuts_launch_init(const NutsParams& p, double step0, cudaStream_t s)
{
    nuts_init_scalar_kernel<<<sc_grid(p), 128, 0, s>>>(p, step0);
    LAUNCH_CHECK();
}

}  // namespace bgh

// =============================================================================================
// Asynchronous per-chain state machine (bgh_nuts_run)
// =============================================================================================
namespace bgh {

__device__ __forceinline__ int& afl(const AsyncParams& p, int k, int c) { return p.afl[(size_t)k * p.n.Cp + c]; }

// Philox draws of the async machine use the chain's OWN draw index, so that a chain's random stream
// does not depend on what the other chains of the batch are doing.
__device__ __forceinline__ double chain_uniform_at(const NutsParams& p, int chain, long long draw, unsigned purpose,
                                                   unsigned idx)
{
    const uint4 r = philox4x32_10(make_uint4((unsigned)draw, purpose, idx, (unsigned)(draw >> 32)), chain_key(p, chain));
    return u01_open(r.x, r.y);
}

__device__ __forceinline__ void schedule_leaf(const AsyncParams& a, int c, int leaf, int depth)
{
    // checkpoint bookkeeping of the iterative tree for leaf index `leaf` of a doubling of depth `depth`
    int idx_max = 0;
    for (int x = leaf >> 1; x > 0; x >>= 1) idx_max += x & 1;
    int store_level = -1, lvl_min = 0, n_lvl = 0;
    if ((leaf & 1) == 0) {
        store_level = idx_max;
    } else {
        int ones = 0;
        for (int x = leaf; x & 1; x >>= 1) ++ones;
        n_lvl = ones;
        lvl_min = idx_max - ones + 1;
    }
    afl(a, kAfStoreLevel, c) = store_level;
    afl(a, kAfLvlMin, c) = lvl_min;
    afl(a, kAfNLvl, c) = n_lvl;
    afl(a, kAfLastLeaf, c) = (leaf == (1 << depth) - 1) ? 1 : 0;
}

// The tick itself (pre / likelihood / finalize / post / scalar) is the persistent kernel in bgh_fused.cu;
// only the initialisation of the per-chain machine state lives here.
__global__ void nuts_async_init_kernel(const AsyncParams a, double step0, double initial_weight)
{
    const NutsParams& p = a.n;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.Cp) return;
    for (int k = 0; k < kAfCount; ++k) afl(a, k, c) = 0;
    sc(p, kScEps, c) = step0;            // tune == 0 -> exp(log_bar) == step0 as well
    sc(p, kScDaLogStep, c) = log(step0);
    sc(p, kScDaLogBar, c) = log(step0);
    sc(p, kScDaHbar, c) = 0.0;
    sc(p, kScDaCount, c) = 1.0;
    sc(p, kScDaMu, c) = log(10.0 * step0);
    sc(p, kScLogpProp, c) = p.logp_w[c];
    sc(p, kScFgW, c) = initial_weight;
    sc(p, kScBgW, c) = 0.0;
    sc(p, kScLogSizeSub, c) = -INFINITY;
    {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        sc(p, kScRunStartUs, c) = (double)(now >> 10);
    }
    afl(a, kAfRecSlot, c) = -1;
    const bool real = c < p.C && (a.tune + a.draws) > 0;
    afl(a, kAfPhase, c) = real ? kPhaseRun : kPhaseFinished;
    if (real) {
        afl(a, kAfNew, c) = 1;
        const double ud = chain_uniform_at(p, c, 0, kPurposeDir, 0u);
        const int dir = (ud < 0.5) ? 1 : -1;
        afl(a, kAfDir, c) = dir;
        sc(p, kScH, c) = 0.5 * step0 * (double)dir;
        schedule_leaf(a, c, 0, 0);
    }
    if (c == 0) *a.n_finished = 0;
}

// standard normals of every running chain's FIRST transition -> slot Z (later transitions: the tick
// kernel refills Z cooperatively, see tick_post in bgh_fused.cu)
__global__ void nuts_async_zgen_kernel(const AsyncParams a)
{
    const NutsParams& p = a.n;
    double* const Z = slot(p, kSlotZ);
    const size_t n = (size_t)p.D * p.Cp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / p.D), d = (int)(i % p.D);     // chain-major, as the tick kernel keeps its state
        Z[i] = momentum_normal(chain_key(p, c), 0, (unsigned)(p.coord_offset + d));
    }
}

cudaError_t nuts_async_launch_init(const AsyncParams& a, double step0, double initial_weight, cudaStream_t s, bool fill_z)
{
    if (fill_z) nuts_async_zgen_kernel<<<a.n.n_vec_blocks, 256, 0, s>>>(a);   // nuts_tick_kernel only (bgh_tick2.cu draws its normals in place)
    nuts_async_init_kernel<<<(a.n.Cp + 127) / 128, 128, 0, s>>>(a, step0, initial_weight);
    return cudaGetLastError();
}

}  // namespace bgh
