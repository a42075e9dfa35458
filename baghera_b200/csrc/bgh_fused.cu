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
#include "bgh_grid_device.cuh"
#include "bgh_internal.h"
#include "bgh_lik_device.cuh"
#include "bgh_peer_device.cuh"
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
            // the all-reduce leaves TOTALS in the payload: rows of rank-shared genes this rank does not hold
            // must contribute zero to the next exchange
            if (on && which == 0 && !p.finish_inline && p.n_replicated == 0)
                for (int r = 0; r < p.n_shared_rows; ++r)
                    if (r != p.first_gene_row && r != p.last_gene_row) p.payload[(size_t)(4 + r) * Cp + ch] = 0.0;
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
template <int CL, int CPL, bool GAMMA>
__global__ void __launch_bounds__(kLikThreads, 1) logp_grad_fused_kernel(const FusedParams f)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    lik_stage_init(smem_raw);
    __syncthreads();
    uint32_t par = 0;
    unsigned long long target = f.bar_base;
    for (int cb = 0; cb < f.chain_blocks; ++cb) lik_phase<CL, CPL, GAMMA>(f.lik, smem_raw, blockIdx.x, cb, par);
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
template <int CL, int CPL, bool GAMMA>
__global__ void __launch_bounds__(kLikThreads, 1) logp_grad_sharded_kernel(const FusedParams f, const PeerParams pr)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    lik_stage_init(smem_raw);
    __syncthreads();
    uint32_t par = 0;
    unsigned long long target = f.bar_base;
    for (int cb = 0; cb < f.chain_blocks; ++cb) lik_phase<CL, CPL, GAMMA>(f.lik, smem_raw, blockIdx.x, cb, par);
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
        for (int j = 0; j < p.n_replicated; ++j)
            finish_gene(p, j, ch, p.payload[(size_t)(4 + j) * Cp + ch], load_gene_coef(p, j, ch, true));
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

// ---------------------------------------------------------------------------------------------
// Vector phases of the tick.  Inside bgh_nuts_run the sampler state is CHAIN-MAJOR ([slot][Cp][D], the
// checkpoints [Cp][level][D] and the trace [draw][C][D]); only the two vectors exchanged with the
// likelihood pass — working position QW and its gradient GW — stay gene-major [D][Cp].
// Every CTA owns a contiguous block of rows (coordinates) of ALL chains, so all CTAs carry the same load
// whatever the chains are doing.  Inside the CTA a warp works on one chain at a time (lane <-> row):
// the chain's tree state is warp-uniform (no divergence, no predication) and every access to the ~20
// per-chain vectors is a contiguous full-sector stream.  QW / GW pass through a shared-memory tile that
// is loaded / stored with lane <-> chain (coalesced on the gene-major arrays) and read transposed.
// Per-chain sums are reduced by a shuffle tree into partials[cta][quantity][chain]; the scalar phase adds
// the CTAs in a fixed order.
// ---------------------------------------------------------------------------------------------
constexpr int kTickMaxChains = 128;
constexpr size_t kZSnapOffset = 0;                                   // int[kTickMaxChains]: chains that began a transition
constexpr size_t kCmdOffset = 1024;                                   // per-phase copy of the chains' flags and scalars
// shared-memory copy of afl[kAfCount][Cp] and sc[kAscCount][Cp] (one coalesced load per phase instead of
// a dependent global load per chain and flag), followed by two tiles [rows][Cp + 1] of doubles
struct ChainCmd {
    const int* fl;       // [kAfCount][Cp]
    const double* sc;    // [kAscCount][Cp]
    int Cp;
    __device__ __forceinline__ int f(int k, int c) const { return fl[k * Cp + c]; }
    __device__ __forceinline__ double s(int k, int c) const { return sc[k * Cp + c]; }
};
__device__ __forceinline__ size_t tile_offset(int Cp)
{
    return (kCmdOffset + (size_t)Cp * (kAfCount * sizeof(int) + kAscCount * sizeof(double)) + 127) & ~(size_t)127;
}
__device__ __forceinline__ ChainCmd load_cmd(const AsyncParams& a, unsigned char* smem_raw)
{
    const NutsParams& p = a.n;
    double* const ssc = reinterpret_cast<double*>(smem_raw + kCmdOffset);
    int* const sfl = reinterpret_cast<int*>(ssc + (size_t)kAscCount * p.Cp);
    for (int i = threadIdx.x; i < kAscCount * p.Cp; i += kLikThreads) ssc[i] = p.sc[i];
    for (int i = threadIdx.x; i < kAfCount * p.Cp; i += kLikThreads) sfl[i] = a.afl[i];
    ChainCmd cm;
    cm.fl = sfl; cm.sc = ssc; cm.Cp = p.Cp;
    return cm;     // the caller synchronises the CTA before the first use
}

struct RowTile {
    int r0, nr;            // rows [r0, r0 + nr) of this CTA in the current tile
    int stride;            // Cp + 1 doubles (conflict-free transposed reads)
    double* tq;            // working position tile
    double* tg;            // gradient tile
};
__device__ __forceinline__ int tile_rows_max(int Cp)
{
    return (int)((kLikStageBytes - tile_offset(Cp)) / 2 / ((size_t)(Cp + 1) * sizeof(double)));
}
__device__ __forceinline__ double* slotc(const NutsParams& p, int s, int chain)
{
    return p.state + ((size_t)s * p.Cp + chain) * p.D;     // chain-major slot: [Cp][D]
}
// gene-major [D][Cp] -> tile (lane <-> chain: coalesced)
__device__ __forceinline__ void tile_load(const NutsParams& p, const double* __restrict__ src, double* tile, const RowTile& T)
{
    const int Cp = p.Cp;
    const int c = threadIdx.x % Cp, r = threadIdx.x / Cp, rstep = kLikThreads / Cp;
    if (r < rstep)
        for (int rr = r; rr < T.nr; rr += rstep) tile[rr * T.stride + c] = src[(size_t)(T.r0 + rr) * Cp + c];
}
__device__ __forceinline__ void tile_store(const NutsParams& p, double* __restrict__ dst, const double* tile, const RowTile& T)
{
    const int Cp = p.Cp;
    const int c = threadIdx.x % Cp, r = threadIdx.x / Cp, rstep = kLikThreads / Cp;
    if (r < rstep)
        for (int rr = r; rr < T.nr; rr += rstep) dst[(size_t)(T.r0 + rr) * Cp + c] = tile[rr * T.stride + c];
}
// Accesses of the RARE per-chain vectors (tree edges, proposal, adaptation windows, Z, trace) carry an L2
// evict-first policy so that they stream through without displacing the vectors every tick re-reads
// (working position / momentum / gradient / mass matrix / running p_sum / low checkpoint levels).
__device__ __forceinline__ double ldg_ef(const double* p, unsigned long long pol)
{
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void stg_ef(double* p, double v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- tick, vector part 1: pending end-of-doubling ops, trace record, mass-matrix adaptation, momentum
// refresh of chains that start a transition, then the first half of the leapfrog for every running chain
struct PreLoads {
    double pw, var, qs, gs, pst, psr, qp, gp, fm, fr, bm, br, qe, pe, ge, z;
};

__device__ __forceinline__ void pre_chain(const AsyncParams& a, const ChainCmd& cm, const RowTile& T, const int c, const int lane,
                                          const bool first_tile)
{
    const NutsParams& p = a.n;
    const int phase = cm.f(kAfPhase, c);
    if (phase == kPhaseFinished) return;                           // warp-uniform
    const bool merge = cm.f(kAfMerge, c) != 0, take = cm.f(kAfTake, c) != 0, take_top = cm.f(kAfTakeTop, c) != 0;
    const bool is_new = cm.f(kAfNew, c) != 0, is_first = cm.f(kAfFirst, c) != 0;
    const bool adapt = cm.f(kAfAdapt, c) != 0, sw = cm.f(kAfSwitch, c) != 0;
    const int rec_slot = cm.f(kAfRecSlot, c);
    const int dir = cm.f(kAfDir, c), dir_prev = cm.f(kAfDirPrev, c);
    const double h = cm.s(kScH, c);
    const bool flush = phase == kPhaseFlush;
    // slot s of this chain lives at base + s * sstride (addresses are formed at the access)
    double* const base = p.state + (size_t)c * p.D;
    const size_t sstride = (size_t)p.Cp * p.D;
#define SLOT(s) (base + (size_t)(s) * sstride)
    double* const tq = T.tq + c;
    const double* const tg = T.tg + c;
    if (!merge && !is_new && !is_first && !flush) {
        // ---- plain leaf inside a subtree (most chain-ticks): 2 global loads per row, 4 rows in flight
        constexpr int U = 4;
        for (int rr0 = lane; rr0 < T.nr; rr0 += 32 * U) {
            double pw[U], var[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int rr = rr0 + 32 * u;
                pw[u] = var[u] = 0.0;
                if (rr < T.nr) { pw[u] = SLOT(kSlotPW)[T.r0 + rr]; var[u] = SLOT(kSlotVar)[T.r0 + rr]; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int rr = rr0 + 32 * u;
                if (rr < T.nr) {
                    const int d = T.r0 + rr;
                    const double q = tq[rr * T.stride], g = tg[rr * T.stride];
                    if (take) { SLOT(kSlotQS)[d] = q; SLOT(kSlotGS)[d] = g; }
                    const double ph = fma(h, g, pw[u]);
                    SLOT(kSlotPW)[d] = ph;
                    tq[rr * T.stride] = fma(2.0 * h * var[u], ph, q);
                }
            }
        }
        return;
    }
    // ---- general path (end of a doubling, start of a transition, flush)
    const unsigned long long pol = l2_evict_first_policy();
    double inv_fg = 0.0, inv_bg = 0.0;
    if (is_new && adapt) { inv_fg = 1.0 / cm.s(kScFgUse, c); inv_bg = 1.0 / cm.s(kScBgUse, c); }
    const bool ld_s = merge && take_top && !take;                 // proposal of the finished subtree
    const bool ld_prop = (is_new || flush) && !(merge && take_top);
    const bool ld_adapt = is_new && adapt && !flush;
    const bool ld_z = is_new && !flush;
    const bool ld_edge = !flush && !is_new && is_first && !(merge && dir == dir_prev);
    const int eQ = dir > 0 ? kSlotQR : kSlotQL, eP = dir > 0 ? kSlotPR : kSlotPL, eG = dir > 0 ? kSlotGR : kSlotGL;
    const bool rec = (is_new || flush) && rec_slot >= 0;
    double kin0 = 0.0;
    auto load = [&](int d, PreLoads& L) {
        L.pw = SLOT(kSlotPW)[d]; L.var = SLOT(kSlotVar)[d];
        L.qs = L.gs = L.pst = L.psr = L.qp = L.gp = L.fm = L.fr = L.bm = L.br = L.qe = L.pe = L.ge = L.z = 0.0;
        if (ld_s) { L.qs = ldg_ef(SLOT(kSlotQS) + d, pol); L.gs = ldg_ef(SLOT(kSlotGS) + d, pol); }
        if (merge) { L.pst = ldg_ef(SLOT(kSlotPSumTree) + d, pol); L.psr = SLOT(kSlotPSumRun)[d]; }
        if (ld_prop) { L.qp = ldg_ef(SLOT(kSlotQP) + d, pol); L.gp = ldg_ef(SLOT(kSlotGP) + d, pol); }
        if (ld_adapt) { L.fm = ldg_ef(SLOT(kSlotFgMean) + d, pol); L.fr = ldg_ef(SLOT(kSlotFgRaw) + d, pol); L.bm = ldg_ef(SLOT(kSlotBgMean) + d, pol); L.br = ldg_ef(SLOT(kSlotBgRaw) + d, pol); }
        if (ld_z) L.z = ldg_ef(SLOT(kSlotZ) + d, pol);
        if (ld_edge) { L.qe = ldg_ef(SLOT(eQ) + d, pol); L.pe = ldg_ef(SLOT(eP) + d, pol); L.ge = ldg_ef(SLOT(eG) + d, pol); }
    };
    auto process = [&](int rr, PreLoads& L) {
        const int d = T.r0 + rr;
        const double qw = tq[rr * T.stride], gw = tg[rr * T.stride];
        double qp = L.qp, gp = L.gp;
        if (merge) {     // the doubling that ended last tick was valid: fold it into the tree
            const double qs = ld_s ? L.qs : qw, gs = ld_s ? L.gs : gw;
            if (take) { stg_ef(SLOT(kSlotQS) + d, qw, pol); stg_ef(SLOT(kSlotGS) + d, gw, pol); }
            if (take_top) { stg_ef(SLOT(kSlotQP) + d, qs, pol); stg_ef(SLOT(kSlotGP) + d, gs, pol); qp = qs; gp = gs; }
            if (dir_prev > 0) { stg_ef(SLOT(kSlotQR) + d, qw, pol); stg_ef(SLOT(kSlotPR) + d, L.pw, pol); stg_ef(SLOT(kSlotGR) + d, gw, pol); }
            else { stg_ef(SLOT(kSlotQL) + d, qw, pol); stg_ef(SLOT(kSlotPL) + d, L.pw, pol); stg_ef(SLOT(kSlotGL) + d, gw, pol); }
            stg_ef(SLOT(kSlotPSumTree) + d, L.pst + L.psr, pol);
        }
        if (rec) {
            const size_t t = ((size_t)rec_slot * p.C + c) * p.D + d;          // trace [draw][C][D]
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
                SLOT(kSlotVar)[d] = var;
                if (sw) { stg_ef(SLOT(kSlotFgMean) + d, bm, pol); stg_ef(SLOT(kSlotFgRaw) + d, br, pol); stg_ef(SLOT(kSlotBgMean) + d, 0.0, pol); stg_ef(SLOT(kSlotBgRaw) + d, 0.0, pol); }
                else { stg_ef(SLOT(kSlotFgMean) + d, fm, pol); stg_ef(SLOT(kSlotFgRaw) + d, fr, pol); stg_ef(SLOT(kSlotBgMean) + d, bm, pol); stg_ef(SLOT(kSlotBgRaw) + d, br, pol); }
            }
            mom = L.z * rsqrt(var);          // p ~ N(0, 1/var); z was drawn ahead of time (slot Z)
            stg_ef(SLOT(kSlotPL) + d, mom, pol); stg_ef(SLOT(kSlotPR) + d, mom, pol); stg_ef(SLOT(kSlotPSumTree) + d, mom, pol);
            stg_ef(SLOT(kSlotQL) + d, qp, pol); stg_ef(SLOT(kSlotQR) + d, qp, pol); stg_ef(SLOT(kSlotGL) + d, gp, pol); stg_ef(SLOT(kSlotGR) + d, gp, pol);
            kin0 = fma(0.5 * var * mom, mom, kin0);
            q = qp; g = gp;
        } else if (is_first) {
            if (merge && dir == dir_prev) { q = qw; mom = L.pw; g = gw; }
            else { q = L.qe; mom = L.pe; g = L.ge; }
        } else {
            q = qw; mom = L.pw; g = gw;
            if (take) { stg_ef(SLOT(kSlotQS) + d, q, pol); stg_ef(SLOT(kSlotGS) + d, g, pol); }
        }
        const double ph = fma(h, g, mom);
        SLOT(kSlotPW)[d] = ph;
        tq[rr * T.stride] = fma(2.0 * h * var, ph, q);
    };
    for (int rr0 = lane; rr0 < T.nr; rr0 += 64) {
        PreLoads L0, L1;
        const bool two = rr0 + 32 < T.nr;
        load(T.r0 + rr0, L0);
        if (two) load(T.r0 + rr0 + 32, L1);
        process(rr0, L0);
        if (two) process(rr0 + 32, L1);
    }
#undef SLOT
    if (is_new && !flush) {     // kinetic energy of the fresh momenta -> quantity kAqKin0 (added over the tiles)
        const double t = warp_sum(kin0);
        if (lane == 0) {
            double* const dst = p.partials + ((size_t)blockIdx.x * kAsyncPartials + kAqKin0) * p.Cp + c;
            *dst = (first_tile ? 0.0 : *dst) + t;
        }
    }
}

__device__ __forceinline__ void tick_pre(const AsyncParams& a, unsigned char* smem_raw, const PhaseTrace& pt)
{
    const NutsParams& p = a.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rpc = (p.D + (int)gridDim.x - 1) / (int)gridDim.x;
    const int r_begin = (int)blockIdx.x * rpc, r_end = min(p.D, r_begin + rpc);
    const int trmax = tile_rows_max(p.Cp);
    const int n_tiles = (rpc + trmax - 1) / trmax, tr = (rpc + n_tiles - 1) / n_tiles;
    RowTile T;
    T.stride = p.Cp + 1;
    T.tq = reinterpret_cast<double*>(smem_raw + tile_offset(p.Cp));
    T.tg = T.tq + (size_t)tr * T.stride;
    const double* const GWg = slot(p, kSlotGW);     // gene-major [D][Cp]
    double* const QWg = slot(p, kSlotQW);
    const ChainCmd cm = load_cmd(a, smem_raw);
    for (int t0 = r_begin; t0 < r_end; t0 += tr) {
        T.r0 = t0;
        T.nr = min(tr, r_end - t0);
        tile_load(p, QWg, T.tq, T);
        tile_load(p, GWg, T.tg, T);
        __syncthreads();
        for (int c = warp; c < p.C; c += kLikWarps) pre_chain(a, cm, T, c, lane, t0 == r_begin);
        __syncthreads();
        tile_store(p, QWg, T.tq, T);
        __syncthreads();
    }
}

// second half kick + reductions for the rows of the tile of one chain that checks NL checkpoint levels
// (kinetic, NL pairs of sub-block U-turn dots, whole-tree dots on the last leaf of a doubling);
// U rows per lane and iteration with all loads first.  Everything but the row index is warp-uniform.
template <int NL, int U>
__device__ __forceinline__ void post_rows(const NutsParams& p, const RowTile& T, const int c, const int lane, const int leaf,
                                          const int store_level, const int lvl_min, const bool last, const int dir,
                                          const double h, const bool first_tile)
{
    const double* const VAR = slotc(p, kSlotVar, c);
    double* const PW = slotc(p, kSlotPW, c);
    double* const PSR = slotc(p, kSlotPSumRun, c);
    const double* const PST = slotc(p, kSlotPSumTree, c);
    const double* const PO = slotc(p, dir > 0 ? kSlotPL : kSlotPR, c);
    const double* const tg = T.tg + c;
    double a_kin = 0.0, a_t0 = 0.0, a_t1 = 0.0;
    double lv[2 * (NL > 0 ? NL : 1)];
#pragma unroll
    for (int k = 0; k < 2 * (NL > 0 ? NL : 1); ++k) lv[k] = 0.0;
    for (int rr0 = lane; rr0 < T.nr; rr0 += 32 * U) {
        double var[U], pwv[U], psr[U], pstv[U], pov[U];
        double2 ck[U][NL > 0 ? NL : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rr = rr0 + 32 * u, d = T.r0 + rr;
            var[u] = pwv[u] = psr[u] = pstv[u] = pov[u] = 0.0;
#pragma unroll
            for (int l = 0; l < NL; ++l) ck[u][l] = make_double2(0.0, 0.0);
            if (rr < T.nr) {
                var[u] = VAR[d]; pwv[u] = PW[d];
                if (leaf != 0) psr[u] = PSR[d];
                if (last) { pstv[u] = PST[d]; pov[u] = PO[d]; }
#pragma unroll
                for (int l = 0; l < NL; ++l) ck[u][l] = *ckpt2(p, lvl_min + l, d, c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rr = rr0 + 32 * u, d = T.r0 + rr;
            if (rr < T.nr) {
                const double mom = fma(h, tg[rr * T.stride], pwv[u]);
                PW[d] = mom;
                const double vel = var[u] * mom;
                a_kin = fma(0.5 * vel, mom, a_kin);
                const double ps = psr[u] + mom;
                PSR[d] = ps;
                if (store_level >= 0) *ckpt2(p, store_level, d, c) = make_double2(mom, ps);
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    const double pc = ck[u][l].x;
                    const double sub = ps - ck[u][l].y + pc;
                    lv[2 * l] = fma(sub, var[u] * pc, lv[2 * l]);
                    lv[2 * l + 1] = fma(sub, vel, lv[2 * l + 1]);
                }
                if (last) {
                    const double pst = pstv[u] + ps;
                    a_t0 = fma(pst, vel, a_t0);
                    a_t1 = fma(pst, var[u] * pov[u], a_t1);
                }
            }
        }
    }
    // warp sums -> partials[cta][quantity][chain] (accumulated over the tiles of this CTA, in order)
    double* const out = p.partials + (size_t)blockIdx.x * kAsyncPartials * p.Cp + c;
    auto emit = [&](int k, double v) {
        const double t = warp_sum(v);
        if (lane == 0) out[(size_t)k * p.Cp] = (first_tile ? 0.0 : out[(size_t)k * p.Cp]) + t;
    };
    emit(0, a_kin);
#pragma unroll
    for (int k = 0; k < 2 * NL; ++k) emit(1 + k, lv[k]);
    if (last) {
        emit(kAqTree0, a_t0);
        emit(kAqTree0 + 1, a_t1);
    }
}

// ---- tick, vector part 2: second half kick + per-chain reductions
__device__ __forceinline__ void tick_post(const AsyncParams& a, unsigned char* smem_raw, const PhaseTrace& pt)
{
    const NutsParams& p = a.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // snapshot, for the Z refill done later in the scalar phase, of the chains that began a transition this tick
    {
        int* const s_new = reinterpret_cast<int*>(smem_raw + kZSnapOffset);
        for (int cc = threadIdx.x; cc < p.C; cc += kLikThreads)
            s_new[cc] = (afl(a, kAfPhase, cc) == kPhaseRun && afl(a, kAfNew, cc) != 0) ? afl(a, kAfDraw, cc) + 1 : -1;
    }
    stamp(pt, 8);
    static_assert(kZSnapOffset + kTickMaxChains * sizeof(int) <= kCmdOffset, "Z snapshot area");
    const int rpc = (p.D + (int)gridDim.x - 1) / (int)gridDim.x;
    const int r_begin = (int)blockIdx.x * rpc, r_end = min(p.D, r_begin + rpc);
    const int trmax = 2 * tile_rows_max(p.Cp);           // only the gradient tile is needed
    const int n_tiles = (rpc + trmax - 1) / trmax, tr = (rpc + n_tiles - 1) / n_tiles;
    RowTile T;
    T.stride = p.Cp + 1;
    T.tq = nullptr;
    T.tg = reinterpret_cast<double*>(smem_raw + tile_offset(p.Cp));
    const double* const GWg = slot(p, kSlotGW);
    const ChainCmd cm = load_cmd(a, smem_raw);
    for (int t0 = r_begin; t0 < r_end; t0 += tr) {
        T.r0 = t0;
        T.nr = min(tr, r_end - t0);
        const bool first_tile = t0 == r_begin;
        tile_load(p, GWg, T.tg, T);
        __syncthreads();
        for (int c = warp; c < p.C; c += kLikWarps) {
            if (cm.f(kAfPhase, c) != kPhaseRun) continue;              // warp-uniform
            const int dir = cm.f(kAfDir, c), leaf = cm.f(kAfLeaf, c), store_level = cm.f(kAfStoreLevel, c);
            const int lvl_min = cm.f(kAfLvlMin, c), n_lvl = cm.f(kAfNLvl, c);
            const bool last = cm.f(kAfLastLeaf, c) != 0;
            const double h = cm.s(kScH, c);
            switch (n_lvl) {
                case 0: post_rows<0, 4>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
                case 1: post_rows<1, 4>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
                case 2: post_rows<2, 2>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
                case 3: post_rows<3, 2>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
                case 4: post_rows<4, 2>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
                case 5: post_rows<5, 1>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
                case 6: post_rows<6, 1>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
                case 7: post_rows<7, 1>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
                case 8: post_rows<8, 1>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
                case 9: post_rows<9, 1>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
                default: post_rows<10, 1>(p, T, c, lane, leaf, store_level, lvl_min, last, dir, h, first_tile); break;
            }
        }
        __syncthreads();
    }
    stamp(pt, 9);
    stamp(pt, 10);
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
            double* const Zc = Z + (size_t)cc * p.D;                  // chain-major
            for (int d = gtid; d < p.D; d += gn) Zc[d] = momentum_normal(key, draw, (unsigned)(p.coord_offset + d));
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
                if (draw == (long long)a.tune - 1) s_fl[kAfTuneEndUs] = (int)(unsigned)(global_timer_ns() >> 10);
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

template <int CL, int CPL, bool GAMMA>
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
        for (int cb = 0; cb < f.chain_blocks; ++cb) lik_phase<CL, CPL, GAMMA>(f.lik, smem_raw, blockIdx.x, cb, par);
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
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* f) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLikSmemBytes);
    };
    set((const void*)logp_grad_fused_kernel<CL, CPL, false>);
    set((const void*)logp_grad_fused_kernel<CL, CPL, true>);
    set((const void*)logp_grad_sharded_kernel<CL, CPL, false>);
    set((const void*)nuts_tick_kernel<CL, CPL, false>);
    set((const void*)nuts_tick_kernel<CL, CPL, true>);
    return e;
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

// BGH_PLAIN_LAUNCH=1 (experiment): ordinary launch instead of cudaLaunchCooperativeKernel.  The grid is one CTA
// per SM, so the CTAs are co-resident as soon as the SMs are free; the grid barrier has a timeout either way.
static bool plain_launch()
{
    static const bool v = [] { const char* e = getenv("BGH_PLAIN_LAUNCH"); return e && *e == '1'; }();
    return v;
}
static cudaError_t launch_grid(const void* fn, int n_cta, void** args, cudaStream_t s)
{
    if (plain_launch()) return cudaLaunchKernel(fn, dim3(n_cta), dim3(kLikThreads), args, kLikSmemBytes, s);
    return cudaLaunchCooperativeKernel(fn, dim3(n_cta), dim3(kLikThreads), args, kLikSmemBytes, s);
}

template <int CL, int CPL>
static cudaError_t launch_fused_t(const FusedParams& f, int n_cta, cudaStream_t s)
{
    void* args[] = {const_cast<FusedParams*>(&f)};
    const void* fn = f.lik.model == BGH_MODEL_GENE_GAMMA ? (const void*)logp_grad_fused_kernel<CL, CPL, true>
                                                         : (const void*)logp_grad_fused_kernel<CL, CPL, false>;
    return launch_grid(fn, n_cta, args, s);
}
template <int CL, int CPL>
static cudaError_t launch_sharded_t(const FusedParams& f, const PeerParams& pr, int n_cta, cudaStream_t s)
{
    void* args[] = {const_cast<FusedParams*>(&f), const_cast<PeerParams*>(&pr)};
    return launch_grid((const void*)logp_grad_sharded_kernel<CL, CPL, false>, n_cta, args, s);   // the gamma model runs on single-rank contexts
}
template <int CL, int CPL>
static cudaError_t launch_tick_t(const TickParams& t, int n_cta, cudaStream_t s)
{
    void* args[] = {const_cast<TickParams*>(&t)};
    const void* fn = t.f.lik.model == BGH_MODEL_GENE_GAMMA ? (const void*)nuts_tick_kernel<CL, CPL, true>
                                                           : (const void*)nuts_tick_kernel<CL, CPL, false>;
    return cudaLaunchCooperativeKernel(fn, dim3(n_cta), dim3(kLikThreads), args, kLikSmemBytes, s);
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
