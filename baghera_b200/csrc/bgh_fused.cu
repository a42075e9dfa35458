// bgh_fused.cu — cooperative (one CTA per SM, all co-resident) kernels that run the whole hot path
// without leaving the device:
//
//   logp_grad_fused_kernel : likelihood+gradient pass | grid barrier | finalize   (one launch per
//                            evaluation; replaces lik_kernel + finalize_kernel back to back)
//   nuts_tick_kernel       : the asynchronous per-chain NUTS state machine as a PERSISTENT kernel.
//                            Every tick = pre (end-of-doubling ops, momentum refresh, half kick +
//                            drift) | likelihood+gradient | finalize | post (half kick, kinetic energy,
//                            U-turn dots) | scalar (per-chain tree / adaptation state machine), the
//                            phases separated by grid barriers instead of kernel boundaries; the host
//                            launches it once per `n_ticks` ticks and only polls the number of
//                            finished chains.
//
// Replaces pm.sample's per-leapfrog Python/Theano round trip (ref: baghera_tool/gene_regression.py:100-106,
// PyMC3 3.6 step_methods/hmc/nuts.py [recalled, SURVEY §8a]).  Semantics of the tick are those of the
// lock-step kernels in bgh_sampler.cu (same Philox counters, same decisions).
#include <cuda_runtime.h>
#include <math.h>

#include "bgh_fin_device.cuh"
#include "bgh_fused.h"
#include "bgh_internal.h"
#include "bgh_lik_device.cuh"
#include "bgh_sampler.h"
#include "bgh_sampler_device.cuh"

#ifndef BGH_POST_HOIST
#define BGH_POST_HOIST 6
#endif
#ifndef BGH_POST_J
#define BGH_POST_J 2
#endif
#ifndef BGH_POST_HL
#define BGH_POST_HL 2
#endif

namespace bgh {

// ---------------------------------------------------------------------------------------------
// grid barrier: one monotonic 64-bit arrival counter (release-add, acquire-spin).  Requires all CTAs to
// be co-resident (cooperative launch).  A CTA that
// waits longer than kBarrierTimeoutNs raises *status and returns so that a logic error cannot hang
// the GPU; the host then reports the failure and resets the barrier words.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kBarrierTimeoutNs = 2000000000ull;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_i32(const int* p)
{
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// `target` is thread 0's running arrival target: the host passes the counter value at launch
// (FusedParams.bar_base, it knows how many barriers every launch executes) and every barrier adds gridDim.x.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long& target, int* status)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        unsigned long long old;
        asm volatile("atom.add.release.gpu.global.u64 %0, [%1], %2;" : "=l"(old) : "l"(bar), "l"(1ull) : "memory");
        if (old + 1 < target) {
            const unsigned long long t0 = global_timer_ns();
            unsigned spins = 0;
            while (true) {
                unsigned long long v;
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
                if (v >= target) break;
                if ((++spins & 1023u) == 0 && (global_timer_ns() - t0 > kBarrierTimeoutNs || ld_volatile_i32(status) != 0)) {
                    atomicExch(status, 1);
                    break;
                }
            }
        } else {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// finalize, spread over the resident CTAs (all sums in a fixed order -> bit-reproducible):
//   * chain quads : one CTA task per 4 chains; warp w sums quantity (w & 3) of chain (w >> 2) over the
//                   per-CTA partials part[n_cta][4][Cp] (lane-strided loads + xor tree) while the
//                   sum-independent transcendental terms of the chain are evaluated; then the closing
//                   formulas (logp, grad rows 0..2) — or, multi-rank, the payload rows 0..3.
//   * long split genes (> kShortSplit slots, e.g. NonCoding): one CTA task per gene.
//   * short split genes: one warp task per (gene, block of 32 chains).
// Tasks are dealt round-robin over the CTAs, short tasks starting after the block-level ones.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fin_phase(const FinParams& p, unsigned char* smem_raw)
{
    double* const sm = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Cp = p.Cp;
    const int n_quads = (Cp + 3) / 4;
    const int n_blk = (int)gridDim.x;

    // ---- chain quads
    for (int task = blockIdx.x; task < n_quads; task += n_blk) {
        const int which = warp & 3, ch = task * 4 + (warp >> 2);
        const bool on = ch < Cp;
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = lane + 32 * i;
            v[i] = (on && row < p.n_cta) ? p.part[((size_t)row * 4 + which) * Cp + ch] : 0.0;
        }
        ChainConst k;
        k.iW = k.iW2 = k.mi = k.lp0 = 0.0;
        if (p.finish_inline && which == 0 && on) k = chain_const(p, ch);   // overlaps the loads above
        double acc = 0.0;
        for (int i = 0; i * 32 < p.n_cta; i += 8) {
            if (i > 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int row = lane + 32 * (i + j);
                    v[j] = (on && row < p.n_cta) ? p.part[((size_t)row * 4 + which) * Cp + ch] : 0.0;
                }
            }
            acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            sm[warp] = acc;
            if (on) p.payload[(size_t)which * Cp + ch] = acc;   // rows 0..3 of the payload are [4][Cp]
        }
        __syncthreads();
        if (p.finish_inline && which == 0 && lane == 0 && on) {
            ChainSums cs;
            cs.A_ll = sm[warp + 0];
            cs.A_e = sm[warp + 1];
            cs.S1 = sm[warp + 2];
            cs.S2 = sm[warp + 3];
            finish_chain(p, ch, cs, k);
        }
        __syncthreads();
    }

    // ---- long split genes: warp w takes slots w, w+16, ...; fixed-order combine over the 16 warps
    for (int k = (((int)blockIdx.x - n_quads) % n_blk + n_blk) % n_blk; k < p.n_long; k += n_blk) {
        const int s0 = p.split_ptr[k], s1 = p.split_ptr[k + 1];
        const int gene = p.split_gene[k], row = p.split_row[k];
        for (int cb = 0; cb < Cp; cb += 32) {
            const int ch = cb + lane;
            const bool on = ch < Cp;
            GeneCoef gc = load_gene_coef(p, gene, on ? ch : 0, row < 0 && warp == 0);
            double acc = 0.0;
#pragma unroll 4
            for (int i = s0 + warp; i < s1; i += kLikWarps) {
                const int slot = p.split_slots[i];
                if (on) acc += p.slots[(size_t)slot * Cp + ch];
            }
            sm[warp * 33 + lane] = acc;
            __syncthreads();
            if (warp == 0 && on) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < kLikWarps; ++w) tot += sm[w * 33 + lane];
                if (row >= 0)
                    p.payload[(size_t)(4 + row) * Cp + ch] = tot;   // rank-shared: finished after the exchange
                else
                    finish_gene(p, gene, ch, tot, gc);
            }
            __syncthreads();
        }
    }

    // ---- short split genes (<= 32 slots): one warp per (gene, block of 32 chains)
    {
        const int ncb = (Cp + 31) / 32;
        const int n_tasks = (p.n_split - p.n_long) * ncb;
        const int total_warps = n_blk * kLikWarps;
        const int first = ((n_quads + p.n_long) % n_blk) * kLikWarps;
        int t = ((int)blockIdx.x * kLikWarps + warp - first + total_warps) % total_warps;
        for (; t < n_tasks; t += total_warps) {
            const int k = p.n_long + t / ncb, cb = t % ncb;
            const int s0 = p.split_ptr[k], n = p.split_ptr[k + 1] - s0;
            const int gene = p.split_gene[k], row = p.split_row[k];
            const int ch = cb * 32 + lane;
            const bool on = ch < Cp;
            const int my_slot = (lane < n) ? p.split_slots[s0 + lane] : 0;
            GeneCoef gc = load_gene_coef(p, gene, on ? ch : 0, row < 0);
            double tot = 0.0;
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
                const int slot = __shfl_sync(0xffffffffu, my_slot, i);
                if (on) tot += p.slots[(size_t)slot * Cp + ch];
            }
            if (on) {
                if (row >= 0)
                    p.payload[(size_t)(4 + row) * Cp + ch] = tot;
                else
                    finish_gene(p, gene, ch, tot, gc);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// one launch per evaluation
// ---------------------------------------------------------------------------------------------
template <int CL, int CPL>
__global__ void __launch_bounds__(kLikThreads, 1) logp_grad_fused_kernel(const FusedParams f)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    lik_stage_init(smem_raw);
    __syncthreads();
    uint32_t par = 0;
    unsigned long long target = f.bar_base;
    for (int cb = 0; cb < f.chain_blocks; ++cb) lik_phase<CL, CPL>(f.lik, smem_raw, blockIdx.x, cb, par);
    grid_barrier(f.bar, target, f.status);
    fin_phase(f.fin, smem_raw);
}

// ---------------------------------------------------------------------------------------------
// sharded evaluation: likelihood pass | finalize (payload) | ONE-SHOT ALL-REDUCE over peer memory | finish,
// all inside one launch per rank.  The payload ([4 + shared genes][Cp] fp64, a few KB) is latency bound:
// instead of a separate NCCL launch, CTA 0 stores its payload straight into every peer's mailbox over
// NVLink, releases a system-scope flag, waits for the peers' flags and adds the R contributions in RANK
// ORDER (every rank obtains bit-identical totals, hence identical NUTS decisions).  Mailboxes are
// double buffered by launch parity: a rank can only be one launch ahead of its slowest peer, because it
// needs that peer's flag of launch e to finish launch e.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// executed by one whole CTA; on return payload[0..n) holds the rank-ordered total on every rank
__device__ __forceinline__ void peer_allreduce(const PeerParams& pr, double* payload, int* status)
{
    const int tid = threadIdx.x;
    const int par = (int)(pr.epoch & 1ull);
    const size_t my_box = ((size_t)par * pr.world + pr.rank) * pr.n;
    for (int r = 0; r < pr.world; ++r) {
        if (r == pr.rank) continue;
        double* const dst = pr.mail[r] + my_box;
        for (int i = tid; i < pr.n; i += blockDim.x) dst[i] = payload[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < pr.world && tid != pr.rank) {
        st_release_sys_u64(pr.flag[tid] + ((size_t)par * pr.world + pr.rank) * 16, pr.epoch);
        const unsigned long long* const mine = pr.flag[pr.rank] + ((size_t)par * pr.world + tid) * 16;
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys_u64(mine) < pr.epoch) {
            if ((++spins & 255u) == 0 && (global_timer_ns() - t0 > kBarrierTimeoutNs || ld_volatile_i32(status) != 0)) {
                atomicExch(status, 2);
                break;
            }
        }
    }
    __syncthreads();
    const double* const box = pr.mail[pr.rank] + (size_t)par * pr.world * pr.n;
    for (int i = tid; i < pr.n; i += blockDim.x) {
        const double own = payload[i];
        double t = 0.0;
        for (int r = 0; r < pr.world; ++r) t += (r == pr.rank) ? own : ld_relaxed_sys_f64(box + (size_t)r * pr.n + i);
        payload[i] = t;
    }
    __syncthreads();
}

template <int CL, int CPL>
__global__ void __launch_bounds__(kLikThreads, 1) logp_grad_sharded_kernel(const FusedParams f, const PeerParams pr)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    lik_stage_init(smem_raw);
    __syncthreads();
    uint32_t par = 0;
    unsigned long long target = f.bar_base;
    for (int cb = 0; cb < f.chain_blocks; ++cb) lik_phase<CL, CPL>(f.lik, smem_raw, blockIdx.x, cb, par);
    grid_barrier(f.bar, target, f.status);
    fin_phase(f.fin, smem_raw);            // finish_inline == 0: payload rows + local split genes
    grid_barrier(f.bar, target, f.status);
    if (blockIdx.x != 0) return;
    peer_allreduce(pr, f.fin.payload, f.status);
    // finish (what finish_kernel does after an NCCL all-reduce)
    const FinParams& p = f.fin;
    for (int ch = threadIdx.x; ch < p.Cp; ch += blockDim.x) {
        const int Cp = p.Cp;
        ChainSums s;
        s.A_ll = p.payload[0 * Cp + ch];
        s.A_e = p.payload[1 * Cp + ch];
        s.S1 = p.payload[2 * Cp + ch];
        s.S2 = p.payload[3 * Cp + ch];
        finish_chain(p, ch, s, chain_const(p, ch));
        if (p.first_gene_row >= 0)
            finish_gene(p, 0, ch, p.payload[(size_t)(4 + p.first_gene_row) * Cp + ch], load_gene_coef(p, 0, ch, true));
        if (p.last_gene_row >= 0 && !(p.G_local == 1 && p.first_gene_row >= 0))
            finish_gene(p, p.G_local - 1, ch, p.payload[(size_t)(4 + p.last_gene_row) * Cp + ch],
                        load_gene_coef(p, p.G_local - 1, ch, true));
    }
}

// ---------------------------------------------------------------------------------------------
// persistent NUTS tick kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int& afl(const AsyncParams& p, int k, int c) { return p.afl[(size_t)k * p.n.Cp + c]; }

// Philox draws of the async machine use the chain's OWN draw index, so that a chain's random stream
// does not depend on what the other chains of the batch are doing.
__device__ __forceinline__ double chain_uniform_at(const NutsParams& p, int chain, long long draw, unsigned purpose,
                                                   unsigned idx)
{
    const uint4 r = philox4x32_10(make_uint4((unsigned)draw, purpose, idx, (unsigned)(draw >> 32)), chain_key(p, chain));
    return u01_open(r.x, r.y);
}

// tools only: globaltimer stamps of thread 0 of every CTA, 16 slots per (tick, CTA), first 8 traced ticks
struct PhaseTrace {
    unsigned long long* buf;   // NULL in production
    int row;                   // (traced tick index * n_cta + cta) or -1
};
__device__ __forceinline__ void stamp(const PhaseTrace& pt, int k)
{
    if (pt.buf != nullptr && pt.row >= 0 && threadIdx.x == 0) pt.buf[(size_t)pt.row * 16 + k] = global_timer_ns();
}

struct TickIdx {
    int chain, r0, rstep, tpr, rpb;
    bool on;
};
__device__ __forceinline__ TickIdx tick_index(const NutsParams& p)
{
    TickIdx v;
    v.tpr = p.Cp < kLikThreads ? p.Cp : kLikThreads;   // threads along the chain axis (Cp <= kLikThreads)
    v.rpb = kLikThreads / v.tpr;                       // rows per CTA pass
    v.chain = threadIdx.x % v.tpr;
    v.r0 = blockIdx.x * v.rpb + threadIdx.x / v.tpr;
    v.rstep = gridDim.x * v.rpb;
    v.on = v.chain < p.Cp && threadIdx.x < v.tpr * v.rpb;
    return v;
}

// ---- tick, vector part 1: pending end-of-doubling ops, trace record, mass-matrix adaptation, momentum
// refresh of chains that start a transition, then the first half of the leapfrog for every running chain.
// Two rows per iteration, every load of both rows issued before the first store (the phase is bound by
// memory latency x serial iterations, not by bytes).
struct PreLoads {
    double qw, pw, gw, var, qs, gs, pst, psr, qp, gp, fm, fr, bm, br, qe, pe, ge, z;
};

__device__ __forceinline__ void tick_pre(const AsyncParams& a, unsigned char* smem_raw, const PhaseTrace& pt)
{
    const NutsParams& p = a.n;
    const TickIdx v = tick_index(p);
    double kin0 = 0.0;
    const int c = v.chain;
    const int phase = v.on ? afl(a, kAfPhase, c) : kPhaseFinished;
    if (phase != kPhaseFinished) {
        const bool merge = afl(a, kAfMerge, c) != 0, take = afl(a, kAfTake, c) != 0, take_top = afl(a, kAfTakeTop, c) != 0;
        const bool is_new = afl(a, kAfNew, c) != 0, is_first = afl(a, kAfFirst, c) != 0;
        const bool adapt = afl(a, kAfAdapt, c) != 0, sw = afl(a, kAfSwitch, c) != 0;
        const int rec_slot = afl(a, kAfRecSlot, c);
        const int dir = afl(a, kAfDir, c), dir_prev = afl(a, kAfDirPrev, c);
        const double h = sc(p, kScH, c);
        double inv_fg = 0.0, inv_bg = 0.0;
        if (is_new && adapt) { inv_fg = 1.0 / sc(p, kScFgUse, c); inv_bg = 1.0 / sc(p, kScBgUse, c); }
        double* const VAR = slot(p, kSlotVar);
        double* const QW = slot(p, kSlotQW);
        double* const PW = slot(p, kSlotPW);
        double* const GW = slot(p, kSlotGW);
        double* const QS = slot(p, kSlotQS);
        double* const GS = slot(p, kSlotGS);
        double* const QP = slot(p, kSlotQP);
        double* const GP = slot(p, kSlotGP);
        double* const PST = slot(p, kSlotPSumTree);
        const double* const PSR = slot(p, kSlotPSumRun);
        double* const QL = slot(p, kSlotQL);
        double* const PL = slot(p, kSlotPL);
        double* const GL = slot(p, kSlotGL);
        double* const QR = slot(p, kSlotQR);
        double* const PR = slot(p, kSlotPR);
        double* const GR = slot(p, kSlotGR);
        double* const FM = slot(p, kSlotFgMean);
        double* const FR = slot(p, kSlotFgRaw);
        double* const BM = slot(p, kSlotBgMean);
        double* const BR = slot(p, kSlotBgRaw);
        const double* const Z = slot(p, kSlotZ);
        // per-chain, loop-invariant load predicates
        const bool flush = phase == kPhaseFlush;
        const bool ld_s = merge && take_top && !take;                 // proposal of the finished subtree
        const bool ld_prop = (is_new || flush) && !(merge && take_top);
        const bool ld_adapt = is_new && adapt && !flush;
        const bool ld_z = is_new && !flush;
        const bool ld_edge = !flush && !is_new && is_first && !(merge && dir == dir_prev);
        const double* const QE = dir > 0 ? QR : QL;
        const double* const PE = dir > 0 ? PR : PL;
        const double* const GE = dir > 0 ? GR : GL;

        auto load = [&](int d, PreLoads& L) {
            const size_t i = (size_t)d * p.Cp + c;
            L.qw = QW[i]; L.pw = PW[i]; L.gw = GW[i]; L.var = VAR[i];
            L.qs = L.qw; L.gs = L.gw;
            L.pst = L.psr = L.qp = L.gp = L.fm = L.fr = L.bm = L.br = L.qe = L.pe = L.ge = L.z = 0.0;
            if (ld_s) { L.qs = QS[i]; L.gs = GS[i]; }
            if (merge) { L.pst = PST[i]; L.psr = PSR[i]; }
            if (ld_prop) { L.qp = QP[i]; L.gp = GP[i]; }
            if (ld_adapt) { L.fm = FM[i]; L.fr = FR[i]; L.bm = BM[i]; L.br = BR[i]; }
            if (ld_z) L.z = Z[i];
            if (ld_edge) { L.qe = QE[i]; L.pe = PE[i]; L.ge = GE[i]; }
        };
        auto process = [&](int d, PreLoads& L) {
            const size_t i = (size_t)d * p.Cp + c;
            double qp = L.qp, gp = L.gp;
            // ---- the doubling that ended last tick was valid: fold it into the tree
            if (merge) {
                if (take) { QS[i] = L.qw; GS[i] = L.gw; }
                if (take_top) { QP[i] = L.qs; GP[i] = L.gs; qp = L.qs; gp = L.gs; }
                if (dir_prev > 0) { QR[i] = L.qw; PR[i] = L.pw; GR[i] = L.gw; }
                else { QL[i] = L.qw; PL[i] = L.pw; GL[i] = L.gw; }
                PST[i] = L.pst + L.psr;
            }
            if ((is_new || flush) && rec_slot >= 0 && c < p.C) {
                const size_t t = ((size_t)rec_slot * p.D + d) * p.C + c;
                if (a.trace_f32) reinterpret_cast<float*>(a.trace)[t] = (float)qp;
                else reinterpret_cast<double*>(a.trace)[t] = qp;
            }
            if (flush) return;
            double q, mom, g, var = L.var;
            if (is_new) {
                if (adapt) {   // quadpotential.QuadPotentialDiagAdapt.update with the sample just produced
                    double fm = L.fm, fr = L.fr, bm = L.bm, br = L.br;
                    welford_add(qp, inv_fg, fm, fr);
                    welford_add(qp, inv_bg, bm, br);
                    var = fr * inv_fg;
                    VAR[i] = var;
                    if (sw) { FM[i] = bm; FR[i] = br; BM[i] = 0.0; BR[i] = 0.0; }
                    else { FM[i] = fm; FR[i] = fr; BM[i] = bm; BR[i] = br; }
                }
                mom = L.z * rsqrt(var);          // p ~ N(0, 1/var); z was drawn ahead of time (slot Z)
                PL[i] = mom; PR[i] = mom; PST[i] = mom;
                QL[i] = qp; QR[i] = qp; GL[i] = gp; GR[i] = gp;
                kin0 = fma(0.5 * var * mom, mom, kin0);
                q = qp; g = gp;
            } else if (is_first) {
                if (merge && dir == dir_prev) { q = L.qw; mom = L.pw; g = L.gw; }
                else { q = L.qe; mom = L.pe; g = L.ge; }
            } else {
                q = L.qw; mom = L.pw; g = L.gw;
                if (take) { QS[i] = q; GS[i] = g; }
            }
            const double ph = fma(h, g, mom);
            PW[i] = ph;
            QW[i] = fma(2.0 * h * var, ph, q);
        };
        int d = v.r0;
        for (; d + v.rstep < p.D; d += 2 * v.rstep) {
            PreLoads L0, L1;
            load(d, L0);
            load(d + v.rstep, L1);
            process(d, L0);
            process(d + v.rstep, L1);
        }
        if (d < p.D) {
            PreLoads L0;
            load(d, L0);
            process(d, L0);
        }
    }
    // kinetic energy of the fresh momenta -> quantity kAqKin0 (fixed-order CTA reduction)
    {
        double* const red = reinterpret_cast<double*>(smem_raw);
        red[threadIdx.x] = kin0;
        __syncthreads();
        if (threadIdx.x < v.tpr && v.on) {
            double t = 0.0;
            for (int r = 0; r < v.rpb; ++r) t += red[r * v.tpr + threadIdx.x];
            p.partials[((size_t)blockIdx.x * kAsyncPartials + kAqKin0) * p.Cp + c] = t;
        }
        __syncthreads();
    }
}

// ---- tick, vector part 2: second half kick + per-chain reductions (kinetic, sub-block U-turn dots for
// the chain's own checkpoint levels, speculative whole-tree dots on the last leaf of a doubling).
//
// Every CTA owns a contiguous block of rows and walks it in tiles of 32 rows x all chains:
//   core pass   lane <-> chain (the [D][Cp] state arrays are chain-minor: coalesced): half kick, running
//               p_sum, kinetic and whole-tree dots; {p, p_sum, var} of the tile go to shared memory
//   level pass  lane <-> row, warp <-> chain (transposed through shared memory): the checkpoints are
//               chain-major [Cp][level][D], so the store of the chain's new checkpoint (even leaves) and
//               the loads of the levels it checks (odd leaves) are contiguous 512-byte warp accesses no
//               matter which level each chain is at; the per-(chain, level) dots are reduced by a
//               shuffle tree and accumulated in shared memory, tile after tile (fixed order).
constexpr int kTileRows = 32;
constexpr int kTickMaxChains = 128;   // shared-memory budget of the post phase
struct PostLayout {
    int stride;            // tile row stride in doubles (tpr + 1: conflict-free transposed reads)
    double* t_mom;         // [kTileRows][stride]
    double* t_ps;
    double* t_var;
    double* l_acc;         // [tpr][2 * kNutsMaxDepth]
    int4* cmd;             // [tpr] {n_lvl, lvl_min, store_level, running}
    double* red;           // [3][kLikThreads] (reuses the tile area after the last tile)
};
__device__ __forceinline__ PostLayout post_layout(unsigned char* smem_raw, int tpr)
{
    PostLayout L;
    L.stride = tpr + 1;
    double* base = reinterpret_cast<double*>(smem_raw);
    L.t_mom = base;
    L.t_ps = L.t_mom + kTileRows * L.stride;
    L.t_var = L.t_ps + kTileRows * L.stride;
    L.l_acc = base + (size_t)3 * kTileRows * (kTickMaxChains + 1);          // fixed offsets: the tile area is
    L.cmd = reinterpret_cast<int4*>(L.l_acc + (size_t)kTickMaxChains * 2 * kNutsMaxDepth);   // sized for kTickMaxChains
    L.red = base;
    return L;
}
constexpr size_t kPostSmemBytes = ((size_t)3 * kTileRows * (kTickMaxChains + 1) + (size_t)kTickMaxChains * 2 * kNutsMaxDepth) * sizeof(double) +
                                  (size_t)kTickMaxChains * sizeof(int4);
constexpr size_t kZSnapOffset = kPostSmemBytes;   // snapshot of the chains that began a transition (Z refill)
static_assert(kZSnapOffset + kTickMaxChains * sizeof(int) <= kLikStageBytes, "post-phase shared memory");
static_assert((size_t)3 * kLikThreads * sizeof(double) <= (size_t)3 * kTileRows * (kTickMaxChains + 1) * sizeof(double), "reduction area inside the tile area");

__device__ __forceinline__ void tick_post(const AsyncParams& a, unsigned char* smem_raw, const PhaseTrace& pt)
{
    const NutsParams& p = a.n;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tpr = p.Cp;                       // threads along the chain axis (Cp <= kTickMaxChains)
    const int rpb = kLikThreads / tpr;          // rows per CTA pass of the core mapping
    const int c = tid % tpr, rs = tid / tpr;
    const bool t_on = rs < rpb;
    const PostLayout L = post_layout(smem_raw, tpr);
    // ---- per-chain commands of this tick -> shared memory; snapshot for the Z refill (tick_scalar)
    {
        int* const s_new = reinterpret_cast<int*>(smem_raw + kZSnapOffset);
        for (int cc = tid; cc < tpr; cc += kLikThreads) {
            const bool run = afl(a, kAfPhase, cc) == kPhaseRun;
            L.cmd[cc] = make_int4(run ? afl(a, kAfNLvl, cc) : 0, afl(a, kAfLvlMin, cc), run ? afl(a, kAfStoreLevel, cc) : -1, run ? 1 : 0);
            s_new[cc] = (cc < p.C && run && afl(a, kAfNew, cc) != 0) ? afl(a, kAfDraw, cc) + 1 : -1;
        }
        for (int o = tid; o < tpr * 2 * kNutsMaxDepth; o += kLikThreads) L.l_acc[o] = 0.0;
    }
    const bool on = t_on && afl(a, kAfPhase, c) == kPhaseRun;
    const int dir = afl(a, kAfDir, c), leaf = afl(a, kAfLeaf, c);
    const bool last = afl(a, kAfLastLeaf, c) != 0;
    const double h = sc(p, kScH, c);
    const double* const VAR = slot(p, kSlotVar);
    double* const PW = slot(p, kSlotPW);
    const double* const GW = slot(p, kSlotGW);
    double* const PSR = slot(p, kSlotPSumRun);
    const double* const PST = slot(p, kSlotPSumTree);
    const double* const PO = slot(p, dir > 0 ? kSlotPL : kSlotPR);
    double a_kin = 0.0, a_t0 = 0.0, a_t1 = 0.0;
    const int rows_per_cta = (p.D + (int)gridDim.x - 1) / (int)gridDim.x;
    const int r_begin = (int)blockIdx.x * rows_per_cta;
    const int r_end = min(p.D, r_begin + rows_per_cta);
    __syncthreads();
    stamp(pt, 8);
    constexpr int U = 4;    // rows per thread and tile at 64 chains (more chains: fewer rows per thread)
    const int n_u = (kTileRows + rpb - 1) / rpb;             // core rows per thread and tile (<= U for Cp >= 64)
    const int buf_doubles = 3 * kTileRows * L.stride;
    double var[U], gwv[U], pwv[U], psr[U], pstv[U], pov[U];

    // core pass, part 1: all global loads of a tile's rows of this thread (kept in registers)
    auto core_load = [&](int tile0, int u0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rt = rs + (u0 + u) * rpb, d = tile0 + rt;
            var[u] = gwv[u] = pwv[u] = psr[u] = pstv[u] = pov[u] = 0.0;
            if (on && rt < kTileRows && d < r_end) {
                const size_t i = (size_t)d * p.Cp + c;
                var[u] = VAR[i]; gwv[u] = GW[i]; pwv[u] = PW[i];
                if (leaf != 0) psr[u] = PSR[i];
                if (last) { pstv[u] = PST[i]; pov[u] = PO[i]; }
            }
        }
    };
    // core pass, part 2: half kick, running p_sum, kinetic / whole-tree dots; tile -> shared memory
    auto core_compute = [&](int tile0, int u0, int buf) {
        double* const tm = L.t_mom + buf * buf_doubles;
        double* const tp = L.t_ps + buf * buf_doubles;
        double* const tv = L.t_var + buf * buf_doubles;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rt = rs + (u0 + u) * rpb, d = tile0 + rt;
            if (t_on && rt < kTileRows) {
                double mom = 0.0, ps = 0.0;
                if (on && d < r_end) {
                    const size_t i = (size_t)d * p.Cp + c;
                    mom = fma(h, gwv[u], pwv[u]);
                    PW[i] = mom;
                    const double vel = var[u] * mom;
                    a_kin = fma(0.5 * vel, mom, a_kin);
                    ps = psr[u] + mom;
                    PSR[i] = ps;
                    if (last) {
                        const double pst = pstv[u] + ps;
                        a_t0 = fma(pst, vel, a_t0);
                        a_t1 = fma(pst, var[u] * pov[u], a_t1);
                    }
                }
                tm[rt * L.stride + c] = mom;     // rows past the end / idle chains: zeros
                tp[rt * L.stride + c] = ps;
                tv[rt * L.stride + c] = var[u];
            }
        }
    };
    // level pass (lane <-> row of the tile, warp <-> chain): 4 chains per warp and round, the
    // checkpoint loads of all four in flight together (2 levels hoisted, deeper levels are rare)
    auto level_pass = [&](int tile0, int buf) {
        const double* const tm = L.t_mom + buf * buf_doubles;
        const double* const tp = L.t_ps + buf * buf_doubles;
        const double* const tv = L.t_var + buf * buf_doubles;
        const int d = tile0 + lane;
        const bool row_ok = d < r_end;
        constexpr int J = BGH_POST_J, HL = BGH_POST_HL;
        auto reduce_add = [&](int cc, int l, double d0, double d1) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                d0 += __shfl_xor_sync(0xffffffffu, d0, o);
                d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            }
            if (lane == 0) {
                L.l_acc[cc * 2 * kNutsMaxDepth + 2 * l] += d0;
                L.l_acc[cc * 2 * kNutsMaxDepth + 2 * l + 1] += d1;
            }
        };
        for (int cc0 = warp; cc0 < tpr; cc0 += J * kLikWarps) {
            int4 cm[J];
            double mom[J], ps[J], vr[J];
            double2 ck[J][HL];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int cc = cc0 + j * kLikWarps;
                cm[j] = cc < tpr ? L.cmd[cc] : make_int4(0, 0, -1, 0);
                mom[j] = ps[j] = vr[j] = 0.0;
                if (cm[j].w != 0 && (cm[j].x > 0 || cm[j].z >= 0)) {         // warp-uniform
                    mom[j] = tm[lane * L.stride + cc];
                    ps[j] = tp[lane * L.stride + cc];
                    vr[j] = tv[lane * L.stride + cc];
                    if (cm[j].z >= 0 && row_ok) *ckpt2(p, cm[j].z, d, cc) = make_double2(mom[j], ps[j]);
                }
#pragma unroll
                for (int l = 0; l < HL; ++l) {
                    ck[j][l] = make_double2(0.0, 0.0);
                    if (l < cm[j].x && row_ok) ck[j][l] = *ckpt2(p, cm[j].y + l, d, cc);
                }
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int cc = cc0 + j * kLikWarps;
                const double vel = vr[j] * mom[j];
#pragma unroll
                for (int l = 0; l < HL; ++l) {
                    if (l < cm[j].x) {                                       // warp-uniform
                        const double pc = ck[j][l].x;
                        const double sub = ps[j] - ck[j][l].y + pc;          // rows past the end: exactly 0
                        reduce_add(cc, l, sub * (vr[j] * pc), sub * vel);
                    }
                }
                for (int l = HL; l < cm[j].x; ++l) {
                    double2 k2 = make_double2(0.0, 0.0);
                    if (row_ok) k2 = *ckpt2(p, cm[j].y + l, d, cc);
                    const double sub = ps[j] - k2.y + k2.x;
                    reduce_add(cc, l, sub * (vr[j] * k2.x), sub * vel);
                }
            }
        }
    };

    // (a software pipeline over the tiles — core loads of tile t+1 in flight during the level pass of tile t,
    // two tile buffers — was measured slower: it pushes the kernel over the 128-register budget)
    for (int tile0 = r_begin; tile0 < r_end; tile0 += kTileRows) {
        for (int u0 = 0; u0 < n_u; u0 += U) {
            core_load(tile0, u0);
            core_compute(tile0, u0, 0);
        }
        __syncthreads();
        level_pass(tile0, 0);
        __syncthreads();
    }
    stamp(pt, 9);
    // ---- CTA reduction of the three core sums in a fixed order, then the partials of this CTA
    L.red[0 * kLikThreads + tid] = a_kin;
    L.red[1 * kLikThreads + tid] = a_t0;
    L.red[2 * kLikThreads + tid] = a_t1;
    __syncthreads();
    stamp(pt, 10);
    for (int o = tid; o < (kAqTree0 + 2) * tpr; o += kLikThreads) {
        const int k = o / tpr, cc = o % tpr;
        double t = 0.0;
        if (k == 0 || k >= kAqTree0) {
            const int q = k == 0 ? 0 : 1 + (k - kAqTree0);
            for (int r = 0; r < rpb; ++r) t += L.red[q * kLikThreads + r * tpr + cc];
        } else {
            t = L.l_acc[cc * 2 * kNutsMaxDepth + (k - 1)];
        }
        p.partials[((size_t)blockIdx.x * kAsyncPartials + k) * p.Cp + cc] = t;
    }
    __syncthreads();
}


__device__ __forceinline__ void schedule_leaf_sm(int* fl, int leaf, int depth)
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
    fl[kAfStoreLevel] = store_level;
    fl[kAfLvlMin] = lvl_min;
    fl[kAfNLvl] = n_lvl;
    fl[kAfLastLeaf] = (leaf == (1 << depth) - 1) ? 1 : 0;
}

// ---- tick, scalar part: the per-chain NUTS state machine, one CTA per chain.  The chain's partial
// sums are reduced over the CTA partials in a fixed order (16 strips per quantity, then the strips
// in order); its scalars / flags are staged in shared memory, updated by one thread, written back.
__device__ __forceinline__ void tick_scalar(const AsyncParams& a, unsigned char* smem_raw)
{
    const NutsParams& p = a.n;
    double* const strip = reinterpret_cast<double*>(smem_raw);          // [32][16]
    double* const s_sum = strip + 32 * 16;                               // [kAsyncPartials]
    double* const s_sc = s_sum + 32;                                     // [kAscCount]
    int* const s_fl = reinterpret_cast<int*>(s_sc + 32);                 // [kAfCount]
    const int tid = threadIdx.x;
    const int n_blk = (int)gridDim.x;
    if ((int)blockIdx.x >= p.C || n_blk <= p.C) {
        // CTAs without a chain of their own (all CTAs if there are more chains than CTAs): draw the
        // standard normals of the NEXT momentum refresh of every chain that began a transition this tick
        // into slot Z (Philox is counter based: the values do not depend on who computes them or when)
        const int* const s_new = reinterpret_cast<const int*>(smem_raw + kZSnapOffset);
        double* const Z = slot(p, kSlotZ);
        const int first = n_blk > p.C ? p.C : 0, n_help = n_blk - first;
        const int gtid = ((int)blockIdx.x - first) * kLikThreads + tid, gn = n_help * kLikThreads;
        for (int cc = 0; cc < p.C; ++cc) {
            const int draw = s_new[cc];
            if (draw < 0 || draw >= a.tune + a.draws) continue;      // CTA-uniform
            const uint2 key = chain_key(p, cc);
            for (int d = gtid; d < p.D; d += gn)
                Z[(size_t)d * p.Cp + cc] = momentum_normal(key, draw, (unsigned)(p.coord_offset + d));
        }
        __syncthreads();   // the snapshot / reduction area is reused below
    }
    for (int c = blockIdx.x; c < p.C; c += n_blk) {
        int phase = afl(a, kAfPhase, c);
        if (phase == kPhaseFinished) continue;                           // CTA-uniform
        if (phase == kPhaseFlush) {
            if (tid == 0) {
                afl(a, kAfPhase, c) = kPhaseFinished;
                afl(a, kAfMerge, c) = 0; afl(a, kAfTake, c) = 0; afl(a, kAfTakeTop, c) = 0; afl(a, kAfRecSlot, c) = -1;
                atomicAdd(a.n_finished, 1);
            }
            continue;
        }
        {
            const int k = tid >> 4, j = tid & 15;
            double acc = 0.0;
            if (k < kAsyncPartials)
                for (int b = j; b < n_blk; b += 16) acc += p.partials[((size_t)b * kAsyncPartials + k) * p.Cp + c];
            if (k < 32) strip[k * 16 + j] = acc;
            if (tid >= 384 && tid < 384 + kAscCount) s_sc[tid - 384] = sc(p, tid - 384, c);
            if (tid >= 416 && tid < 416 + kAfCount) s_fl[tid - 416] = afl(a, tid - 416, c);
        }
        __syncthreads();
        if (tid < kAsyncPartials) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < 16; ++j) t += strip[tid * 16 + j];
            s_sum[tid] = t;
        }
        __syncthreads();
        if (tid == 0) {
            long long draw = s_fl[kAfDraw];
            int depth = s_fl[kAfDepth], leaf = s_fl[kAfLeaf], dir = s_fl[kAfDir];
            if (s_fl[kAfNew]) {      // the transition began this tick
                const double e0 = s_sum[kAqKin0] - s_sc[kScLogpProp];
                s_sc[kScE0] = e0;
                s_sc[kScEnergyProp] = e0;
                s_sc[kScLogSizeTree] = 0.0;
                s_sc[kScAcceptSum] = 0.0;
                s_sc[kScNProp] = 0.0;
                s_sc[kScMaxDE] = 0.0;
                s_fl[kAfDiverging] = 0;
                s_fl[kAfTurning] = 0;
                if (!isfinite(e0)) s_fl[kAfBadStart] = 1;
            }
            s_fl[kAfNew] = 0; s_fl[kAfFirst] = 0; s_fl[kAfMerge] = 0; s_fl[kAfTakeTop] = 0;
            s_fl[kAfTake] = 0; s_fl[kAfAdapt] = 0; s_fl[kAfSwitch] = 0; s_fl[kAfRecSlot] = -1;

            // ---- the leaf just integrated (same logic as nuts_leaf_scalar_kernel)
            const double kin = s_sum[0];
            const double logp = p.logp_w[c];
            const double energy = kin - logp;
            double dE = energy - s_sc[kScE0];
            if (isnan(dE)) dE = INFINITY;
            if (fabs(dE) > fabs(s_sc[kScMaxDE])) s_sc[kScMaxDE] = dE;
            s_sc[kScNProp] += 1.0;
            bool valid = true;
            if (fabs(dE) < p.emax) {
                s_sc[kScAcceptSum] += fmin(1.0, exp(-dE));
                const double lw = -dE;
                bool take;
                if (leaf == 0) {
                    take = true;
                    s_sc[kScLogSizeSub] = lw;
                } else {
                    const double ls = logaddexp_d(s_sc[kScLogSizeSub], lw);
                    const double u = chain_uniform_at(p, c, draw, kPurposeLeaf, ((unsigned)depth << 16) | (unsigned)leaf);
                    take = log(u) < lw - ls;
                    s_sc[kScLogSizeSub] = ls;
                }
                if (take) {
                    s_fl[kAfTake] = 1;
                    s_sc[kScLogpSub] = logp;
                    s_sc[kScEnergySub] = energy;
                }
                const int n_lvl = s_fl[kAfNLvl];
                bool turning = false;
                for (int l = 0; l < n_lvl; ++l)
                    turning = turning || (s_sum[1 + 2 * l] <= 0.0) || (s_sum[2 + 2 * l] <= 0.0);
                if (turning) {
                    s_fl[kAfTurning] = 1;
                    valid = false;
                }
            } else {
                s_fl[kAfDiverging] = 1;
                valid = false;
            }

            const int max_depth = (draw < a.tune && draw < a.early_draws) ? a.early_depth : a.max_depth;
            bool end_transition = false;
            if (!valid) {
                end_transition = true;
                s_fl[kAfTake] = 0;
                depth += 1;                                       // PyMC3 counts the discarded doubling in `depth`
            } else if (leaf == (1 << depth) - 1) {
                const double s1 = s_sc[kScLogSizeTree], s2 = s_sc[kScLogSizeSub];
                const double u = chain_uniform_at(p, c, draw, kPurposeTop, (unsigned)depth);
                if (log(u) < s2 - s1) {
                    s_fl[kAfTakeTop] = 1;
                    s_sc[kScLogpProp] = s_sc[kScLogpSub];
                    s_sc[kScEnergyProp] = s_sc[kScEnergySub];
                }
                s_sc[kScLogSizeTree] = logaddexp_d(s1, s2);
                s_fl[kAfMerge] = 1;
                s_fl[kAfDirPrev] = dir;
                depth += 1;
                if (s_sum[kAqTree0] <= 0.0 || s_sum[kAqTree0 + 1] <= 0.0) {
                    s_fl[kAfTurning] = 1;
                    end_transition = true;
                }
                if (depth >= max_depth) end_transition = true;
                if (!end_transition) {
                    const double ud = chain_uniform_at(p, c, draw, kPurposeDir, (unsigned)depth);
                    dir = (ud < 0.5) ? 1 : -1;
                    leaf = 0;
                    s_fl[kAfFirst] = 1;
                    s_sc[kScLogSizeSub] = -INFINITY;
                }
            } else {
                leaf += 1;
            }

            if (end_transition) {
                const bool tuning = draw < a.tune;
                const double n = fmax(s_sc[kScNProp], 1.0);
                const double accept = s_sc[kScAcceptSum] / n;
                const double eps_used = s_sc[kScEps];
                if (tuning) {     // step_size.DualAverageAdaptation.update
                    const double count = s_sc[kScDaCount];
                    const double w = 1.0 / (count + p.da_t0);
                    const double hbar = (1.0 - w) * s_sc[kScDaHbar] + w * (p.target_accept - accept);
                    const double log_step = s_sc[kScDaMu] - hbar * sqrt(count) / p.da_gamma;
                    const double mk = pow(count, -p.da_kappa);
                    s_sc[kScDaLogBar] = mk * log_step + (1.0 - mk) * s_sc[kScDaLogBar];
                    s_sc[kScDaHbar] = hbar;
                    s_sc[kScDaCount] = count + 1.0;
                    s_sc[kScDaLogStep] = log_step;
                    s_sc[kScEps] = exp(log_step);
                }
                if (a.stats) {
                    double* st = a.stats + (size_t)draw * BGH_NUTS_N_STATS * p.C;
                    st[0 * p.C + c] = (double)depth;
                    st[1 * p.C + c] = accept;
                    st[2 * p.C + c] = s_sc[kScNProp];
                    st[3 * p.C + c] = (double)s_fl[kAfDiverging];
                    st[4 * p.C + c] = s_sc[kScEnergyProp];
                    st[5 * p.C + c] = tuning ? s_sc[kScEps] : eps_used;
                    st[6 * p.C + c] = s_sc[kScLogpProp];
                    st[7 * p.C + c] = s_sc[kScMaxDE];
                }
                s_fl[kAfRecSlot] = (draw >= a.tune) ? (int)(draw - a.tune) : -1;
                if (tuning) {     // QuadPotentialDiagAdapt.update schedule
                    const int n_adapt = s_fl[kAfNAdapt];
                    double fg = s_sc[kScFgW] + 1.0, bg = s_sc[kScBgW] + 1.0;
                    s_sc[kScFgUse] = fg;
                    s_sc[kScBgUse] = bg;
                    s_fl[kAfAdapt] = 1;
                    const int sw = (n_adapt > 0 && n_adapt % a.window == 0) ? 1 : 0;
                    s_fl[kAfSwitch] = sw;
                    if (sw) { fg = bg; bg = 0.0; }
                    s_sc[kScFgW] = fg;
                    s_sc[kScBgW] = bg;
                    s_fl[kAfNAdapt] = n_adapt + 1;
                }
                draw += 1;
                if (draw == a.tune) s_sc[kScEps] = exp(s_sc[kScDaLogBar]);    // stop tuning
                if (draw >= (long long)a.tune + a.draws) {
                    phase = kPhaseFlush;
                } else {
                    s_fl[kAfNew] = 1;
                    depth = 0;
                    leaf = 0;
                    const double ud = chain_uniform_at(p, c, draw, kPurposeDir, 0u);
                    dir = (ud < 0.5) ? 1 : -1;
                    s_sc[kScLogSizeSub] = -INFINITY;
                }
            }
            s_fl[kAfPhase] = phase;
            s_fl[kAfDraw] = (int)draw;
            s_fl[kAfDepth] = depth;
            s_fl[kAfLeaf] = leaf;
            s_fl[kAfDir] = dir;
            s_sc[kScH] = 0.5 * s_sc[kScEps] * (double)dir;
            schedule_leaf_sm(s_fl, leaf, depth);
        }
        __syncthreads();
        if (tid < kAscCount) sc(p, tid, c) = s_sc[tid];
        if (tid >= 32 && tid < 32 + kAfCount) afl(a, tid - 32, c) = s_fl[tid - 32];
        __syncthreads();
    }
}

template <int CL, int CPL>
__global__ void __launch_bounds__(kLikThreads, 1) nuts_tick_kernel(const TickParams t)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_stop;
    lik_stage_init(smem_raw);
    __syncthreads();
    uint32_t par = 0;
    const FusedParams& f = t.f;
    unsigned long long target = f.bar_base;
    for (int tick = 0; tick < t.n_ticks; ++tick) {
        PhaseTrace pt;
        pt.buf = t.phase_trace;
        {
            const int gt_ = t.tick_base + tick - t.trace_first;
            pt.row = (gt_ >= 0 && gt_ < 8) ? gt_ * (int)gridDim.x + (int)blockIdx.x : -1;
        }
#define BGH_PHASE_STAMP(k) stamp(pt, k)
        BGH_PHASE_STAMP(0);
        tick_pre(t.a, smem_raw, pt);
        BGH_PHASE_STAMP(1);
        grid_barrier(f.bar, target, f.status);
        fence_proxy_async();   // the staging area was written through the generic proxy (reductions)
        BGH_PHASE_STAMP(2);
        for (int cb = 0; cb < f.chain_blocks; ++cb) lik_phase<CL, CPL>(f.lik, smem_raw, blockIdx.x, cb, par);
        BGH_PHASE_STAMP(3);
        grid_barrier(f.bar, target, f.status);
        fin_phase(f.fin, smem_raw);
        BGH_PHASE_STAMP(4);
        grid_barrier(f.bar, target, f.status);
        tick_post(t.a, smem_raw, pt);
        BGH_PHASE_STAMP(5);
        grid_barrier(f.bar, target, f.status);
        BGH_PHASE_STAMP(6);
        tick_scalar(t.a, smem_raw);
        BGH_PHASE_STAMP(7);
        grid_barrier(f.bar, target, f.status);
        if (threadIdx.x == 0) {
            s_stop = (ld_volatile_i32(t.a.n_finished) >= t.a.n.C || ld_volatile_i32(f.status) != 0) ? 1 : 0;
            if (blockIdx.x == 0) *t.ticks_done = *t.ticks_done + 1;
        }
        __syncthreads();
        if (s_stop) break;
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <int CL, int CPL>
static cudaError_t configure_fused_t()
{
    cudaError_t e = cudaFuncSetAttribute(logp_grad_fused_kernel<CL, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kLikSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(logp_grad_sharded_kernel<CL, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLikSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(nuts_tick_kernel<CL, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLikSmemBytes);
}

cudaError_t fused_configure()
{
    cudaError_t e;
    if ((e = configure_fused_t<4, 1>()) != cudaSuccess) return e;
    if ((e = configure_fused_t<8, 1>()) != cudaSuccess) return e;
    if ((e = configure_fused_t<16, 1>()) != cudaSuccess) return e;
    if ((e = configure_fused_t<32, 1>()) != cudaSuccess) return e;
    if ((e = configure_fused_t<32, 2>()) != cudaSuccess) return e;
    return cudaSuccess;
}

template <int CL, int CPL>
static cudaError_t launch_fused_t(const FusedParams& f, int n_cta, cudaStream_t s)
{
    void* args[] = {const_cast<FusedParams*>(&f)};
    return cudaLaunchCooperativeKernel((const void*)logp_grad_fused_kernel<CL, CPL>, dim3(n_cta), dim3(kLikThreads), args,
                                       kLikSmemBytes, s);
}
template <int CL, int CPL>
static cudaError_t launch_sharded_t(const FusedParams& f, const PeerParams& pr, int n_cta, cudaStream_t s)
{
    void* args[] = {const_cast<FusedParams*>(&f), const_cast<PeerParams*>(&pr)};
    return cudaLaunchCooperativeKernel((const void*)logp_grad_sharded_kernel<CL, CPL>, dim3(n_cta), dim3(kLikThreads), args,
                                       kLikSmemBytes, s);
}
template <int CL, int CPL>
static cudaError_t launch_tick_t(const TickParams& t, int n_cta, cudaStream_t s)
{
    void* args[] = {const_cast<TickParams*>(&t)};
    return cudaLaunchCooperativeKernel((const void*)nuts_tick_kernel<CL, CPL>, dim3(n_cta), dim3(kLikThreads), args,
                                       kLikSmemBytes, s);
}

cudaError_t launch_logp_grad_fused(const FusedParams& f, int CL, int CPL, int n_cta, cudaStream_t s)
{
    if (CL == 4 && CPL == 1) return launch_fused_t<4, 1>(f, n_cta, s);
    if (CL == 8 && CPL == 1) return launch_fused_t<8, 1>(f, n_cta, s);
    if (CL == 16 && CPL == 1) return launch_fused_t<16, 1>(f, n_cta, s);
    if (CL == 32 && CPL == 1) return launch_fused_t<32, 1>(f, n_cta, s);
    if (CL == 32 && CPL == 2) return launch_fused_t<32, 2>(f, n_cta, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_logp_grad_sharded(const FusedParams& f, const PeerParams& pr, int CL, int CPL, int n_cta, cudaStream_t s)
{
    if (CL == 4 && CPL == 1) return launch_sharded_t<4, 1>(f, pr, n_cta, s);
    if (CL == 8 && CPL == 1) return launch_sharded_t<8, 1>(f, pr, n_cta, s);
    if (CL == 16 && CPL == 1) return launch_sharded_t<16, 1>(f, pr, n_cta, s);
    if (CL == 32 && CPL == 1) return launch_sharded_t<32, 1>(f, pr, n_cta, s);
    if (CL == 32 && CPL == 2) return launch_sharded_t<32, 2>(f, pr, n_cta, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_nuts_ticks(const TickParams& t, int CL, int CPL, int n_cta, cudaStream_t s)
{
    if (CL == 4 && CPL == 1) return launch_tick_t<4, 1>(t, n_cta, s);
    if (CL == 8 && CPL == 1) return launch_tick_t<8, 1>(t, n_cta, s);
    if (CL == 16 && CPL == 1) return launch_tick_t<16, 1>(t, n_cta, s);
    if (CL == 32 && CPL == 1) return launch_tick_t<32, 1>(t, n_cta, s);
    if (CL == 32 && CPL == 2) return launch_tick_t<32, 2>(t, n_cta, s);
    return cudaErrorInvalidValue;
}

}  // namespace bgh
