// bgh_tick2.cu — the batched NUTS sampler as ONE persistent kernel (engine of bgh_nuts_run).
//
// Replaces pm.sample's per-leapfrog Python/Theano round trip (ref: baghera_tool/gene_regression.py:100-106,
// baghera_tool/heritability.py:56-62; PyMC3 3.6 step_methods/hmc/nuts.py [recalled, SURVEY §8a]).
// Decisions are those of the lock-step kernels in bgh_sampler.cu (same Philox counters, same tree logic); the
// bookkeeping of the tree is reorganised so that rare events cost (almost) no memory traffic:
//
//   * Two role-swapped state buffers X0 / X1 = (q, p, grad).  The working state of a chain IS the tree edge in
//     the direction it is integrating; the other buffer holds the opposite edge.  A doubling that reverses the
//     direction flips the chain's buffer index — tree edges are never copied (the lock-step kernels copy three
//     vectors per doubling and read three back).
//   * One running momentum total T per coordinate instead of {running sum of the doubling, sum of the tree}:
//     sub-block sums are differences of T (checkpoints hold {p, T}), the whole-tree sum is T itself, and the
//     end of a doubling needs no vector work at all.
//   * What remains of the rare events (proposal copies when a subtree proposal is taken, trace record, mass
//     matrix update and momentum refresh at the start of a transition) runs as (chain, row) items spread over all
//     threads of the CTA.
//
// A CTA OWNS the rows (coordinates) of the genes that lie completely inside one of its SNP ranges: its likelihood
// pass streams their SNPs and writes their gradient, so the leapfrog of those rows needs no grid-wide
// synchronisation.  Per tick (one leapfrog of every chain):
//
//   F    rare-event items of the CTA's rows
//   PRE  first half of the leapfrog (p += h g; q += 2h var p), thread <-> (chain, rows), coalesced [D][Cp] rows
//   L    likelihood + gradient pass (lik_phase, bgh_lik_device.cuh), unchanged; reads q / writes grad of the
//        chain's working buffer
//   POST second half (p += h g'), kinetic energy, running total, checkpoint store / sub-block U-turn dots,
//        per-chain sums of the CTA's rows
//   B    grid barrier
//   S    scalar phase, one CTA per chain: fixed-order reduction of the CTA partials (likelihood sums and sampler
//        sums), closing formulas, the "late rows" whose gradient needs a grid-wide sum (shared parameters, genes
//        split over several SNP ranges such as NonCoding), the per-chain NUTS state machine, and the late rows'
//        first half step of the NEXT tick.  Sharded contexts exchange the chain's sums with the peer GPUs HERE:
//        one LL-protocol message (data + tag in one 16-byte store: no fence, no flag round trip) per chain, rank
//        and tick over NVLink peer memory, issued by all chain CTAs in parallel.
//   B    grid barrier
//
// Two grid barriers and one peer exchange per leapfrog (nuts_tick_kernel: five barriers; the lock-step driver on
// sharded contexts: three exchanges and a dozen launches).
#include <cuda_runtime.h>
#include <math.h>

#include "bgh_fin_device.cuh"
#include "bgh_grid_device.cuh"
#include "bgh_internal.h"
#include "bgh_lik_device.cuh"
#include "bgh_sampler.h"
#include "bgh_sampler_device.cuh"
#include "bgh_tick2.h"

namespace bgh {

constexpr int kT2MaxChains = 128;
// shared memory: [likelihood staging + mbarriers][command words][kinetic energy of fresh momenta, rare-chain list][row list]
constexpr size_t kT2CmdOffset = kLikSmemBytes;
constexpr size_t kT2CmdBytes = kT2MaxChains * (sizeof(double) + sizeof(int));
constexpr size_t kT2Kin0Offset = kT2CmdOffset + kT2CmdBytes;                  // double[kT2MaxChains] + int[kT2MaxChains]
constexpr int kT2MaxRowsSmem = 8192;
constexpr size_t kT2RowsOffset = kT2Kin0Offset + kT2MaxChains * (sizeof(double) + sizeof(int));   // int[kT2MaxRowsSmem]
constexpr size_t kT2LateOffset = kT2RowsOffset + kT2MaxRowsSmem * sizeof(int);   // int[kLikThreads] late-row table
constexpr int kT2MaxSlotsSmem = 1024;
constexpr size_t kT2SplitOffset = kT2LateOffset + kLikThreads * sizeof(int);     // int[kLikThreads + 1] list starts + int[kT2MaxSlotsSmem] slot ids
constexpr size_t kT2SmemBytes = kT2SplitOffset + (kLikThreads + 1 + kT2MaxSlotsSmem) * sizeof(int);
static_assert(kT2SmemBytes <= 227 * 1024, "shared memory of the fused sampler kernel");

size_t tick2_smem_bytes() { return kT2SmemBytes; }

__device__ __forceinline__ int& t2_afl(const AsyncParams& a, int k, int c) { return a.afl[(size_t)k * a.n.Cp + c]; }

__device__ __forceinline__ double t2_chain_uniform(const NutsParams& p, int chain, long long draw, unsigned purpose, unsigned idx)
{
    const uint4 r = philox4x32_10(make_uint4((unsigned)draw, purpose, idx, (unsigned)(draw >> 32)), chain_key(p, chain));
    return u01_open(r.x, r.y);
}

__device__ __forceinline__ unsigned t2_global_coord(const NutsParams& p, int d)
{
    if (p.gcoord != nullptr) return (unsigned)p.gcoord[d];
    return (unsigned)(d >= p.head_rows ? d + p.coord_offset : d);
}

// State slots of the fused kernel (numbering of bgh_sampler.h reused): the two role-swapped buffers are
// X0 = (kSlotQL, kSlotPL, kSlotGL) and X1 = (kSlotQR, kSlotPR, kSlotGR), i.e. X1 = X0 + 3 slots for q, p and grad
// alike; the running momentum total lives in kSlotPSumRun.  All vectors are gene-major [D][Cp].
constexpr int kXStride = kSlotQR - kSlotQL;     // slots between the buffers
static_assert(kSlotPR - kSlotPL == kXStride && kSlotGR - kSlotGL == kXStride, "buffer layout");
constexpr int kSlotT = kSlotPSumRun;

// checkpoints of the fused kernel: [level][D][Cp] pairs {p, running total}; level l is re-read 2^(l+1) ticks after it was
// stored and half as often as level l-1
__device__ __forceinline__ double2* t2_ckpt(const NutsParams& p, int level, size_t i)
{
    return reinterpret_cast<double2*>(p.ckpt) + (size_t)level * p.D * p.Cp + i;
}
__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

// The vectors every tick re-reads (the two state buffers, mass matrix, running total) carry an L2 evict-last policy: with the
// SNP stream (evict-first) and the checkpoints flowing through, plain LRU throws them out between two ticks although they
// would fit (the host reserves the persisting share of the L2 for the run, bgh_api.cu).
__device__ __forceinline__ unsigned long long l2_keep_policy()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double ldk(const double* p, unsigned long long pol)
{
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void stk(double* p, double v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Row operations.  A "row" is one coordinate d of one chain (index i = d * Cp + chain).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cmd_rare(int bits)
{
    return ((bits & kCbRun) && !(bits & kCbPlain)) || (bits & kCbFlush);
}

// Rare events of a chain at the start of a tick: proposal copies of the doubling that just ended (the subtree's
// proposal is the last leaf / becomes the tree's proposal), trace record and streaming statistics of a finished
// transition, and the start of a transition (mass-matrix update with the sample just produced, momentum refresh,
// both buffers <- start point).  Returns the kinetic energy term of a fresh momentum.
// Same operations as nuts_end_doubling / nuts_adapt_mass / nuts_begin of bgh_sampler.cu.
__device__ __forceinline__ double row_rare(const Tick2Params* tp, int bits, int ch, int d)
{
    const AsyncParams& a = tp->a;
    const NutsParams& p = a.n;
    const size_t vec = (size_t)p.D * p.Cp;
    const size_t i = (size_t)d * p.Cp + ch;
    double* const st = p.state + i;
#define T2SLOT(s) st[(size_t)(s) * vec]
    const bool merge = (bits & kCbMerge) != 0, take = (bits & kCbTake) != 0, take_top = (bits & kCbTakeTop) != 0;
    const bool is_new = (bits & kCbNew) != 0, flush = (bits & kCbFlush) != 0;
    // the buffer the last leaf was integrated in (the scalar phase may have flipped the index for the next doubling)
    const int w_prev = ((bits & kCbW) ? 1 : 0) ^ ((bits & kCbFlip) ? 1 : 0);
    double qp = 0.0, gp = 0.0;
    bool have_p = false;
    if (merge && (take || take_top)) {
        double qs, gs;
        if (take) {      // the last leaf is the proposal of its subtree
            qs = T2SLOT(kSlotQL + kXStride * w_prev);
            gs = T2SLOT(kSlotGL + kXStride * w_prev);
            T2SLOT(kSlotQS) = qs; T2SLOT(kSlotGS) = gs;
        } else {
            qs = T2SLOT(kSlotQS); gs = T2SLOT(kSlotGS);
        }
        if (take_top) { T2SLOT(kSlotQP) = qs; T2SLOT(kSlotGP) = gs; qp = qs; gp = gs; have_p = true; }
    }
    if (!(is_new || flush)) return 0.0;
    if (!have_p) { qp = T2SLOT(kSlotQP); gp = T2SLOT(kSlotGP); }
    const int rec_slot = t2_afl(a, kAfRecSlot, ch);
    if (rec_slot >= 0) {
        if (a.trace != nullptr) {
            const size_t t = ((size_t)rec_slot * p.C + ch) * p.D + d;          // trace [draw][C][D]
            if (a.trace_f32) reinterpret_cast<float*>(a.trace)[t] = (float)qp;
            else reinterpret_cast<double*>(a.trace)[t] = qp;
        }
        if (tp->head_trace != nullptr && d < tp->row0) tp->head_trace[((size_t)rec_slot * p.C + ch) * tp->row0 + d] = qp;
        if (tp->k6 != nullptr && d >= tp->row0) {
            // streaming posterior statistics of the per-gene effects (ref: gene_regression.py:165-174, :283, :314)
            const bool gamma = tp->lik.model == BGH_MODEL_GENE_GAMMA;
            const double v = gamma ? tp->lik.gamma_rate * exp(qp) : qp;
            const double mi = sigmoid_d(p.state[(size_t)kSlotQP * vec + (size_t)(tp->row0 - 1) * p.Cp + ch]);
            double* const k6 = tp->k6 + i;
            k6[0] += v;
            k6[vec] = fma(v, v, k6[vec]);
            k6[2 * vec] += (v - mi > 0.0) ? 1.0 : 0.0;
        }
    }
    if (flush) return 0.0;
    // ---- start of a transition
    double var = T2SLOT(kSlotVar);
    if (bits & kCbAdapt) {   // quadpotential.QuadPotentialDiagAdapt.update with the sample just produced
        const double inv_fg = 1.0 / sc(p, kScFgUse, ch), inv_bg = 1.0 / sc(p, kScBgUse, ch);
        double fm = T2SLOT(kSlotFgMean), fr = T2SLOT(kSlotFgRaw), bm = T2SLOT(kSlotBgMean), br = T2SLOT(kSlotBgRaw);
        welford_add(qp, inv_fg, fm, fr);
        welford_add(qp, inv_bg, bm, br);
        var = fr * inv_fg;
        T2SLOT(kSlotVar) = var;
        if (bits & kCbSwitch) { T2SLOT(kSlotFgMean) = bm; T2SLOT(kSlotFgRaw) = br; T2SLOT(kSlotBgMean) = 0.0; T2SLOT(kSlotBgRaw) = 0.0; }
        else { T2SLOT(kSlotFgMean) = fm; T2SLOT(kSlotFgRaw) = fr; T2SLOT(kSlotBgMean) = bm; T2SLOT(kSlotBgRaw) = br; }
    }
    const int draw = t2_afl(a, kAfDraw, ch);
    const double z = momentum_normal(chain_key(p, ch), draw, t2_global_coord(p, d));
    const double mom = z * rsqrt(var);          // p ~ N(0, 1/var)
    T2SLOT(kSlotQL) = qp; T2SLOT(kSlotPL) = mom; T2SLOT(kSlotGL) = gp;
    T2SLOT(kSlotQR) = qp; T2SLOT(kSlotPR) = mom; T2SLOT(kSlotGR) = gp;
    T2SLOT(kSlotT) = mom;
#undef T2SLOT
    return 0.5 * var * mom * mom;
}

// first half of the leapfrog for one row of a running chain (q, mom, g: the chain's working buffer)
__device__ __forceinline__ void row_half_step(const NutsParams& p, int bits, double h, size_t i, size_t wo, double q, double mom, double g,
                                              double var, unsigned long long pol)
{
    const size_t vec = (size_t)p.D * p.Cp;
    if ((bits & (kCbPlain | kCbTake)) == (kCbPlain | kCbTake)) {   // the previous leaf was drawn as the subtree proposal
        p.state[(size_t)kSlotQS * vec + i] = q;
        p.state[(size_t)kSlotGS * vec + i] = g;
    }
    const double ph = fma(h, g, mom);
    stk(p.state + (size_t)kSlotPL * vec + wo + i, ph, pol);
    stk(p.state + (size_t)kSlotQL * vec + wo + i, fma(2.0 * h * var, ph, q), pol);
}

// second half of the leapfrog for one row whose new gradient is known.  acc[j * stride] accumulates quantity j
// (kAq* numbering of bgh_sampler.h) of the row's chain; the kinetic energy (quantity 0) stays in a register.
// ck0 = checkpoint of the first checked level, tot = running momentum total (both loaded by the caller).
__device__ __forceinline__ double row_finish(const NutsParams& p, int bits, double h, size_t i, size_t wo, double gnew, double ph, double var,
                                             double tot, double2 ck0, double& kin, double* acc, int stride, unsigned long long pol)
{
    const size_t vec = (size_t)p.D * p.Cp;
    const double mom = fma(h, gnew, ph);
    stk(p.state + (size_t)kSlotPL * vec + wo + i, mom, pol);
    const double vel = var * mom;
    kin = fma(0.5 * vel, mom, kin);
    const double ps = tot + mom;
    stk(p.state + (size_t)kSlotT * vec + i, ps, pol);
    const int sl = (bits >> kCbStoreShift) & 15;
    if (sl) *t2_ckpt(p, sl - 1, i) = make_double2(mom, ps);
    const int nl = (bits >> kCbNLvlShift) & 15, lm = (bits >> kCbLvlMinShift) & 15;
#pragma unroll 1
    for (int l = 0; l < nl; ++l) {
        const double2 ck = l == 0 ? ck0 : *t2_ckpt(p, lm + l, i);
        const double sub = ps - ck.y + ck.x;       // momentum sum of the aligned block
        acc[(1 + 2 * l) * stride] = fma(sub, var * ck.x, acc[(1 + 2 * l) * stride]);     // . v_left
        acc[(2 + 2 * l) * stride] = fma(sub, vel, acc[(2 + 2 * l) * stride]);            // . v_right
    }
    if (bits & kCbLast) {   // whole tree: its momentum sum is the running total; po = momentum at the opposite edge
        const size_t wo_other = wo ? 0 : (size_t)kXStride * vec;
        const double po = p.state[(size_t)kSlotPL * vec + wo_other + i];
        acc[kAqTree0 * stride] = fma(ps, vel, acc[kAqTree0 * stride]);
        acc[(kAqTree0 + 1) * stride] = fma(ps, var * po, acc[(kAqTree0 + 1) * stride]);
    }
    return mom;
}

// which of the kAsyncPartials per-chain quantities a chain running under command `bits` produces this tick
__device__ __forceinline__ unsigned used_quantities(int bits)
{
    if (!(bits & kCbRun)) return 0u;
    const int nl = (bits >> kCbNLvlShift) & 15;
    unsigned m = (1u << (1 + 2 * nl)) - 1u;                       // kinetic + 2 dots per checked level
    if (bits & kCbLast) m |= 3u << kAqTree0;
    if (bits & kCbNew) m |= 1u << kAqKin0;
    return m;
}

// ---------------------------------------------------------------------------------------------
// Vector passes of a CTA over the rows it OWNS, all chains.  Thread t handles chain t % Cp of rows t / Cp,
// t / Cp + 512 / Cp, ... of the CTA's row list: a warp access is a contiguous 256-byte piece of a [D][Cp] row, a
// thread keeps ONE chain for the whole tick (command word and accumulators are thread constants), and several rows
// are in flight per thread.
// ---------------------------------------------------------------------------------------------
constexpr int kUPre = 8, kUPost = 4;
static_assert(kUPost % 2 == 0, "the level-major POST pass takes the rows of a batch in pairs");

struct T2Own {
    const int* rows;     // gene ids of the CTA's rows (shared memory, or global when the list is too long)
    int n;
};

// F: rare-event items (chain with rare-event work, row).  s_k0: scratch for the kinetic energy of fresh momenta
// (staging area, idle here); s_kin0c[chain] receives the CTA's sum in a fixed order.
__device__ __forceinline__ void tick2_rare(const Tick2Params& t, const T2Own& own, const int* s_bits, const int* s_np, int n_np,
                                           double* s_k0, double* s_kin0c, unsigned long long* tr)
{
    const int tid = threadIdx.x;
    constexpr int kCap = (int)(kLikStageBytes / sizeof(double));
    // chains are taken in groups so that the kinetic-energy scratch [chain in group][row] fits the staging area
    const int per_pass = max(1, min(n_np, kCap / own.n));
    for (int j0 = 0; j0 < n_np; j0 += per_pass) {
        const int nj = min(per_pass, n_np - j0);
        const int n_items = nj * own.n;
        for (int item = tid; item < n_items; item += kLikThreads) {
            const int jj = item / own.n, r = item - jj * own.n;
            const int ch = s_np[j0 + jj];
            const int bits = s_bits[ch];
            const double k0 = row_rare(&t, bits, ch, t.row0 + own.rows[r]);
            if ((bits & kCbNew) && item < kCap) s_k0[item] = k0;
        }
        if (tr) tr[12] = global_timer_ns();
        __syncthreads();
        if (tr) tr[13] = global_timer_ns();
        const int warp = tid >> 5, lane = tid & 31;
        for (int jj = warp; jj < nj; jj += kLikWarps) {
            const int ch = s_np[j0 + jj];
            if (!(s_bits[ch] & kCbNew)) continue;
            double x = 0.0;
            for (int r = lane; r < own.n; r += 32) x += s_k0[jj * own.n + r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) s_kin0c[ch] = x;
        }
        __syncthreads();
    }
}

// PRE: the first half of the leapfrog for every running chain
__device__ __forceinline__ void tick2_pre(const Tick2Params& t, const T2Own& own, const int* s_bits, const double* s_h)
{
    const NutsParams& p = t.a.n;
    const int Cp = p.Cp;
    const int ch = threadIdx.x % Cp;
    // the 512 / Cp threads of a chain take contiguous blocks of the row list (consecutive rows: checkpoint sectors are re-used)
    const int blk = (own.n + kLikThreads / Cp - 1) / (kLikThreads / Cp), r0 = (threadIdx.x / Cp) * blk, r1 = min(own.n, r0 + blk);
    const int bits = s_bits[ch];
    const double h = s_h[ch];
    const size_t vec = (size_t)p.D * Cp;
    if (!(bits & kCbRun)) return;
    const size_t wo = (bits & kCbW) ? (size_t)kXStride * vec : 0;
    const unsigned long long pol = l2_keep_policy();
    for (int r = r0; r < r1; r += kUPre) {
        size_t i[kUPre];
        double qw[kUPre], pw[kUPre], gw[kUPre], var[kUPre];
#pragma unroll
        for (int u = 0; u < kUPre; ++u) {
            const int rr = r + u;
            i[u] = (size_t)(t.row0 + own.rows[rr < r1 ? rr : r]) * Cp + ch;
            qw[u] = ldk(p.state + (size_t)kSlotQL * vec + wo + i[u], pol);
            pw[u] = ldk(p.state + (size_t)kSlotPL * vec + wo + i[u], pol);
            gw[u] = ldk(p.state + (size_t)kSlotGL * vec + wo + i[u], pol);
            var[u] = ldk(p.state + (size_t)kSlotVar * vec + i[u], pol);
        }
#pragma unroll
        for (int u = 0; u < kUPre; ++u) {
            if (r + u >= r1) break;
            row_half_step(p, bits, h, i[u], wo, qw[u], pw[u], gw[u], var[u], pol);
        }
    }
}

// POST: second half of the leapfrog with the gradient the likelihood pass just wrote, per-chain sums of the
// CTA's rows -> part[cta][quantity][chain].  Accumulators live in the (idle) likelihood staging area,
// thread-interleaved [quantity][thread], which is also the layout the CTA reduction reads.
__device__ __forceinline__ void tick2_post(const Tick2Params& t, const T2Own& own, const int* s_bits, const double* s_h, const double* s_kin0c,
                                           unsigned need, unsigned char* smem_raw)
{
    const NutsParams& p = t.a.n;
    const int Cp = p.Cp;
    const int tid = threadIdx.x;
    const int ch = tid % Cp;
    const int blk = (own.n + kLikThreads / Cp - 1) / (kLikThreads / Cp), r0 = (tid / Cp) * blk, r1 = min(own.n, r0 + blk);
    const int bits = s_bits[ch];
    const double h = s_h[ch];
    const size_t vec = (size_t)p.D * Cp;
    double* const acc = reinterpret_cast<double*>(smem_raw) + tid;     // quantity j at acc[j * kLikThreads]
    static_assert((size_t)kAsyncPartials * kLikThreads * sizeof(double) <= kLikStageBytes, "accumulators reuse the staging area");
#pragma unroll 1
    for (int j = 1; j < kAsyncPartials; ++j)
        if ((need >> j) & 1u) acc[j * kLikThreads] = 0.0;
    double kin = 0.0;
    const unsigned long long pol = l2_keep_policy();
    if (bits & kCbRun) {
        const size_t wo = (bits & kCbW) ? (size_t)kXStride * vec : 0;
        const int nl = (bits >> kCbNLvlShift) & 15, lm = (bits >> kCbLvlMinShift) & 15;
        for (int r = r0; r < r1; r += kUPost) {
            size_t i[kUPost];
            double gn[kUPost], ph[kUPost], var[kUPost], tot[kUPost];
            double2 ck0[kUPost];
#pragma unroll
            for (int u = 0; u < kUPost; ++u) {
                const int rr = r + u;
                i[u] = (size_t)(t.row0 + own.rows[rr < r1 ? rr : r]) * Cp + ch;
                gn[u] = ldk(p.state + (size_t)kSlotGL * vec + wo + i[u], pol);
                ph[u] = ldk(p.state + (size_t)kSlotPL * vec + wo + i[u], pol);
                var[u] = ldk(p.state + (size_t)kSlotVar * vec + i[u], pol);
                tot[u] = ldk(p.state + (size_t)kSlotT * vec + i[u], pol);
                ck0[u] = nl > 0 ? *t2_ckpt(p, lm, i[u]) : make_double2(0.0, 0.0);
                for (int l = 1; l < nl; ++l) prefetch_l1(t2_ckpt(p, lm + l, i[u]));     // deeper levels: in flight with the row's other data
            }
            // row_finish for the kUPost rows, levels outermost: the checkpoint loads of one level are in flight for all
            // rows at once (a warp holds 32 chains at different tree positions, so it walks max(nl) levels; one row at
            // a time that is max(nl) dependent round trips PER ROW).  Every accumulator still receives the rows in
            // ascending order: sums are bit-identical to row_finish.
            double ps[kUPost], vel[kUPost];
            const int sl = (bits >> kCbStoreShift) & 15;
#pragma unroll
            for (int u = 0; u < kUPost; ++u) {
                if (r + u >= r1) break;
                const double mom = fma(h, gn[u], ph[u]);
                stk(p.state + (size_t)kSlotPL * vec + wo + i[u], mom, pol);
                vel[u] = var[u] * mom;
                kin = fma(0.5 * vel[u], mom, kin);
                ps[u] = tot[u] + mom;
                stk(p.state + (size_t)kSlotT * vec + i[u], ps[u], pol);
                if (sl) *t2_ckpt(p, sl - 1, i[u]) = make_double2(mom, ps[u]);
                if (nl > 0) {
                    const double sub = ps[u] - ck0[u].y + ck0[u].x;
                    acc[1 * kLikThreads] = fma(sub, var[u] * ck0[u].x, acc[1 * kLikThreads]);
                    acc[2 * kLikThreads] = fma(sub, vel[u], acc[2 * kLikThreads]);
                }
            }
#pragma unroll 1
            for (int l = 1; l < nl; ++l) {
                double a1 = acc[(1 + 2 * l) * kLikThreads], a2 = acc[(2 + 2 * l) * kLikThreads];
#pragma unroll
                for (int u0 = 0; u0 < kUPost; u0 += 2) {
                    const double2 cka = *t2_ckpt(p, lm + l, i[u0]), ckb = *t2_ckpt(p, lm + l, i[u0 + 1]);   // a row past the end repeats row r
                    if (r + u0 < r1) {
                        const double sub = ps[u0] - cka.y + cka.x;
                        a1 = fma(sub, var[u0] * cka.x, a1);
                        a2 = fma(sub, vel[u0], a2);
                    }
                    if (r + u0 + 1 < r1) {
                        const double sub = ps[u0 + 1] - ckb.y + ckb.x;
                        a1 = fma(sub, var[u0 + 1] * ckb.x, a1);
                        a2 = fma(sub, vel[u0 + 1], a2);
                    }
                }
                acc[(1 + 2 * l) * kLikThreads] = a1; acc[(2 + 2 * l) * kLikThreads] = a2;
            }
            if (bits & kCbLast) {   // whole tree: its momentum sum is the running total; po = momentum at the opposite edge
                const size_t wo_other = wo ? 0 : (size_t)kXStride * vec;
                double a1 = acc[kAqTree0 * kLikThreads], a2 = acc[(kAqTree0 + 1) * kLikThreads];
#pragma unroll
                for (int u = 0; u < kUPost; ++u) {
                    if (r + u >= r1) break;
                    const double po = p.state[(size_t)kSlotPL * vec + wo_other + i[u]];
                    a1 = fma(ps[u], vel[u], a1);
                    a2 = fma(ps[u], var[u] * po, a2);
                }
                acc[kAqTree0 * kLikThreads] = a1; acc[(kAqTree0 + 1) * kLikThreads] = a2;
            }
        }
    }
    acc[0] = kin;
    __syncthreads();
    // CTA reduction per chain: the 512 / Cp threads of a chain in a fixed order
    const double* const red = reinterpret_cast<const double*>(smem_raw);
    const int per_chain = kLikThreads / Cp;
    for (int o = tid; o < kAsyncPartials * Cp; o += kLikThreads) {
        const int j = o / Cp, c = o % Cp;
        if (!((need >> j) & 1u)) continue;
        double x = 0.0;
        if (j == kAqKin0) x = s_kin0c[c];
        else
            for (int m = 0; m < per_chain; ++m) x += red[j * kLikThreads + c + m * Cp];
        t.part[((size_t)blockIdx.x * kAsyncPartials + j) * Cp + c] = x;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// S: scalar phase of chain c, executed by one whole CTA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int pack_cmd(const int* s_fl, int phase)
{
    int b = 0;
    if (phase == kPhaseRun) b |= kCbRun;
    if (phase == kPhaseFlush) b |= kCbFlush;
    if (s_fl[kAfTake]) b |= kCbTake;
    if (s_fl[kAfMerge]) b |= kCbMerge;
    if (s_fl[kAfTakeTop]) b |= kCbTakeTop;
    if (s_fl[kAfNew]) b |= kCbNew;
    if (s_fl[kAfAdapt]) b |= kCbAdapt;
    if (s_fl[kAfSwitch]) b |= kCbSwitch;
    if (s_fl[kAfLastLeaf]) b |= kCbLast;
    if (s_fl[kAfW]) b |= kCbW;
    if (s_fl[kAfFlip]) b |= kCbFlip;
    const bool copies = s_fl[kAfMerge] && (s_fl[kAfTake] || s_fl[kAfTakeTop]);
    if (phase == kPhaseRun && !s_fl[kAfNew] && !copies) b |= kCbPlain;
    b |= (s_fl[kAfStoreLevel] + 1) << kCbStoreShift;
    b |= s_fl[kAfLvlMin] << kCbLvlMinShift;
    b |= s_fl[kAfNLvl] << kCbNLvlShift;
    return b;
}

__device__ __forceinline__ void t2_schedule_leaf(int* fl, int leaf, int depth)
{
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

// CTA-wide sum of one double per thread in a fixed order (xor tree inside the warp, warps in order)
__device__ __forceinline__ double cta_sum(double v, double* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tsum = 0.0;
#pragma unroll
    for (int w = 0; w < kLikWarps; ++w) tsum += s_red[w];
    return tsum;
}

// LL-protocol exchange of n doubles per rank for chain c: on return vals[0..n) holds the rank-ordered total.
__device__ __forceinline__ void t2_exchange(const T2Peer& pr, int Cp, int c, unsigned seq, double* vals, double* s_tmp, int* status)
{
    // seq: exchange counter of this mailbox (identical on all ranks): tag = epoch_base + seq + 1, parity = seq & 1
    const int tid = threadIdx.x;
    const int n = pr.n_pay;
    const unsigned tag = pr.epoch_base + seq + 1u;
    const int par = (int)(seq & 1u);
    // send: one 16-byte store {lo, tag, hi, tag} per value and peer
    for (int o = tid; o < pr.world * n; o += blockDim.x) {
        const int r = o / n, i = o % n;
        if (r == pr.rank) continue;
        const unsigned long long u = (unsigned long long)__double_as_longlong(vals[i]);
        uint4* const dst = pr.box[r] + (((size_t)par * pr.world + pr.rank) * Cp + c) * n + i;
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"((unsigned)u), "r"(tag), "r"((unsigned)(u >> 32)), "r"(tag)
                     : "memory");
    }
    // receive
    for (int o = tid; o < pr.world * n; o += blockDim.x) {
        const int r = o / n, i = o % n;
        double x;
        if (r == pr.rank || pr.send_only) {
            x = r == pr.rank ? vals[i] : 0.0;
        } else {
            const uint4* const src = pr.box[pr.rank] + (((size_t)par * pr.world + r) * Cp + c) * n + i;
            uint4 w;
            const unsigned long long t0 = global_timer_ns();
            unsigned spins = 0;
            while (true) {
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(src) : "memory");
                if (w.y == tag && w.w == tag) break;
                if ((++spins & 255u) == 0 && (global_timer_ns() - t0 > kBarrierTimeoutNs || ld_volatile_i32(status) != 0)) {
                    atomicExch(status, 2);
                    break;
                }
            }
            x = __longlong_as_double((long long)(((unsigned long long)w.z << 32) | (unsigned long long)w.x));
        }
        s_tmp[r * n + i] = x;
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        double tot = 0.0;
        for (int r = 0; r < pr.world; ++r) tot += s_tmp[r * n + i];
        vals[i] = tot;
    }
    __syncthreads();
}

// shared-memory layout of the scalar phase (inside the likelihood staging area, which is idle here)
struct T2Scalar {
    double* strip;     // [32][16]
    double* sum;       // [kT2Nq + spare]  totals of the chain: [0,4) likelihood sums, [4, 4+kAsyncPartials) sampler sums
    double* pay;       // [kT2Nq + kMaxPay] payload view for the exchange (sum followed by replicated-gene sums)
    double* sc;        // [kAscCount]
    int* fl;           // [kAfCount]
    double* red;       // [kLikWarps]
    double* red2;      // [kLikWarps][kAsyncPartials]
    double* tmp;       // [kMaxPeers][n_pay] exchange staging
    double* sr;        // [n_split] residual sums of the split genes
    int* misc;         // [0] new command bits
};
constexpr int kT2MaxPay = 96;
constexpr int kT2MaxSplit = 8192;

__device__ __forceinline__ T2Scalar t2_scalar_smem(unsigned char* smem_raw)
{
    T2Scalar s;
    double* b = reinterpret_cast<double*>(smem_raw);
    s.strip = b; b += 32 * 16;
    s.sum = b; b += kT2MaxPay;          // sum and pay are the same array: payload = sums followed by replicated-gene sums
    s.pay = s.sum;
    s.sc = b; b += 32;
    s.fl = reinterpret_cast<int*>(b); b += 16;
    s.red = b; b += kLikWarps;
    s.red2 = b; b += kLikWarps * kAsyncPartials;
    s.tmp = b; b += kMaxPeers * kT2MaxPay;
    s.misc = reinterpret_cast<int*>(b); b += 2;
    s.sr = b;
    return s;
}
static_assert((32 * 16 + kT2MaxPay + 32 + 16 + kLikWarps + kLikWarps * kAsyncPartials + kMaxPeers * kT2MaxPay + 2 + kT2MaxSplit) * sizeof(double) <=
                  kLikStageBytes, "scalar phase shared memory");

// residual sums of the split genes of chain c in fin_phase's summation order (bit-identical to the evaluation kernels):
// short lists (<= kShortSplit slots) one THREAD per gene, sequentially in list order — all genes in flight at once, three
// dependent loads deep; long lists (NonCoding) one warp per gene: 16 strided sequential sums added in order.
__device__ __forceinline__ void split_gene_sums(const FinParams& f, int c, double* s_sr)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Cp = f.Cp;
    for (int k = f.n_long + tid; k < f.n_split; k += blockDim.x) {
        const int s0 = f.split_ptr[k], n = f.split_ptr[k + 1] - s0;
        double tot = 0.0;
        for (int i = 0; i < n; ++i) tot += f.slots[(size_t)f.split_slots[s0 + i] * Cp + c];
        s_sr[k] = tot;
    }
    for (int k = warp; k < f.n_long; k += kLikWarps) {
        const int s0 = f.split_ptr[k], n = f.split_ptr[k + 1] - s0;
        double acc = 0.0, tot = 0.0;
        if (lane < kLikWarps)
            for (int i = lane; i < n; i += kLikWarps) acc += f.slots[(size_t)f.split_slots[s0 + i] * Cp + c];
        for (int w = 0; w < kLikWarps; ++w) tot += __shfl_sync(0xffffffffu, acc, w);
        if (lane == 0) s_sr[k] = tot;
    }
}

// the same sums with the (constant) slot lists cached in shared memory: one global round trip instead of three
// (k_first: first short-list gene the per-thread loop still has to sum; the caller may have done one gene per thread itself)
__device__ __forceinline__ void split_gene_sums_cached(const FinParams& f, int c, double* s_sr, const int* s_sptr, const int* s_sidx, int k_first)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Cp = f.Cp;
    for (int k = k_first + tid; k < f.n_split; k += blockDim.x) {
        const int s0 = s_sptr[k], n = s_sptr[k + 1] - s0;
        double tot = 0.0;
        for (int i = 0; i < n; ++i) tot += f.slots[(size_t)s_sidx[s0 + i] * Cp + c];
        s_sr[k] = tot;
    }
    for (int k = warp; k < f.n_long; k += kLikWarps) {
        const int s0 = s_sptr[k], n = s_sptr[k + 1] - s0;
        double acc = 0.0, tot = 0.0;
        if (lane < kLikWarps)
            for (int i = lane; i < n; i += kLikWarps) acc += f.slots[(size_t)s_sidx[s0 + i] * Cp + c];
        for (int w = 0; w < kLikWarps; ++w) tot += __shfl_sync(0xffffffffu, acc, w);
        if (lane == 0) s_sr[k] = tot;
    }
}

// late rows of chain c: local index r in [0, n_late): head rows first, then the split genes in list order
__device__ __forceinline__ int t2_n_late(const Tick2Params& t)
{
    return t.lik.model == BGH_MODEL_GW_NORMAL ? t.a.n.D : t.row0 + t.fin.n_split;
}
__device__ __forceinline__ int t2_late_row(const Tick2Params& t, int r)
{
    return (r < t.row0 || t.lik.model == BGH_MODEL_GW_NORMAL) ? r : t.row0 + t.fin.split_gene[r - t.row0];
}
// 0: rank-local row (its sums are complete on this rank), 1: replicated row (needs the exchanged totals)
__device__ __forceinline__ bool t2_late_replicated(const Tick2Params& t, int r)
{
    if (r < t.row0 || t.lik.model == BGH_MODEL_GW_NORMAL) return true;
    return t.fin.split_row[r - t.row0] >= 0;
}

// second half of the leapfrog for a set of late rows of chain c; adds their contributions to s.sum
__device__ __forceinline__ void t2_late_finish(const Tick2Params& t, const T2Scalar& s, int c, int bits, double h, bool replicated_set)
{
    const NutsParams& p = t.a.n;
    const int n_late = t2_n_late(t);
    const size_t vec = (size_t)p.D * p.Cp;
    double kin = 0.0;
    double acc[kAsyncPartials];
#pragma unroll 1
    for (int j = 0; j < kAsyncPartials; ++j) acc[j] = 0.0;
    const bool single = t.peer.world <= 1;
    const size_t wo = (bits & kCbW) ? (size_t)kXStride * vec : 0;
    for (int r = threadIdx.x; r < n_late; r += blockDim.x) {
        if (!single && t2_late_replicated(t, r) != replicated_set) continue;
        const int d = t2_late_row(t, r);
        const size_t i = (size_t)d * p.Cp + c;
        const double gnew = p.state[(size_t)kSlotGL * vec + wo + i];     // written by the closing formulas (same thread block)
        const double ph = p.state[(size_t)kSlotPL * vec + wo + i], var = p.state[(size_t)kSlotVar * vec + i];
        const double tot = p.state[(size_t)kSlotT * vec + i];
        const int nl = (bits >> kCbNLvlShift) & 15, lm = (bits >> kCbLvlMinShift) & 15;
        const double2 ck0 = nl > 0 ? *t2_ckpt(p, lm, i) : make_double2(0.0, 0.0);
        row_finish(p, bits, h, i, wo, gnew, ph, var, tot, ck0, kin, acc, 1, l2_keep_policy());
    }
    acc[0] = kin;
    const unsigned used = used_quantities(bits) & ~(1u << kAqKin0);
    // fixed-order CTA sums of the quantities this chain uses: xor tree inside the warp, then the warps in order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
    for (int j = 0; j < kAsyncPartials; ++j) {
        if (!((used >> j) & 1u)) continue;
        double x = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s.red2[warp * kAsyncPartials + j] = x;
    }
    __syncthreads();
    if (threadIdx.x < kAsyncPartials && ((used >> threadIdx.x) & 1u)) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < kLikWarps; ++w) tot += s.red2[w * kAsyncPartials + threadIdx.x];
        s.sum[kT2Lik + threadIdx.x] += tot;
    }
    __syncthreads();
}

// rare-event work + first half step of the NEXT leapfrog for the late rows of chain c (new command bits)
__device__ __forceinline__ void t2_late_enter(const Tick2Params& t, const T2Scalar& s, int c, int bits, double h)
{
    const NutsParams& p = t.a.n;
    const int n_late = t2_n_late(t);
    const size_t vec = (size_t)p.D * p.Cp;
    const bool rare = cmd_rare(bits);
    const size_t wo = (bits & kCbW) ? (size_t)kXStride * vec : 0;
    double k0_loc = 0.0, k0_rep = 0.0;
    // two passes: the head rows first (a recorded draw's statistics need the NEW value of mi), then the gene rows
    for (int pass = 0; pass < 2; ++pass) {
        for (int r = threadIdx.x; r < n_late; r += blockDim.x) {
            const bool head = r < t.row0 || t.lik.model == BGH_MODEL_GW_NORMAL;
            if (head != (pass == 0)) continue;
            const int d = t2_late_row(t, r);
            const size_t i = (size_t)d * p.Cp + c;
            if (rare) {
                const double k0 = row_rare(&t, bits, c, d);
                if (t2_late_replicated(t, r)) k0_rep += k0; else k0_loc += k0;
            }
            if (bits & kCbRun) {
                const double qw = p.state[(size_t)kSlotQL * vec + wo + i], pw = p.state[(size_t)kSlotPL * vec + wo + i];
                const double gw = p.state[(size_t)kSlotGL * vec + wo + i], var = p.state[(size_t)kSlotVar * vec + i];
                row_half_step(p, bits, h, i, wo, qw, pw, gw, var, l2_keep_policy());
            }
        }
        __syncthreads();
    }
    if (bits & kCbNew) {     // CTA-uniform
        const double a_loc = cta_sum(k0_loc, s.red);
        const double a_rep = cta_sum(k0_rep, s.red);
        if (threadIdx.x == 0) { s.sc[kScKin0LateLocal] = a_loc; s.sc[kScKin0LateRep] = a_rep; }
    } else if (threadIdx.x == 0) {
        s.sc[kScKin0LateLocal] = 0.0;
        s.sc[kScKin0LateRep] = 0.0;
    }
    __syncthreads();
}

__device__ __forceinline__ void t2_store_chain(const Tick2Params& t, const T2Scalar& s, int c, int bits)
{
    const NutsParams& p = t.a.n;
    const int tid = threadIdx.x;
    if (tid < kAscCount) sc(p, tid, c) = s.sc[tid];
    if (tid >= 32 && tid < 32 + kAfCount) t2_afl(t.a, tid - 32, c) = s.fl[tid - 32];
    if (tid == 64) {
        t.cmd_bits[c] = bits;
        t.cmd_h[c] = s.sc[kScH];
        t.cmd_woff[c] = (bits & kCbW) ? kXStride * p.D * p.Cp : 0;
    }
    __syncthreads();
}

// before the first tick of a run: the late rows take their first half step (the rows the likelihood pass
// owns take it on entering their gene)
__device__ __forceinline__ void tick2_prologue(const Tick2Params& t, unsigned char* smem_raw)
{
    const NutsParams& p = t.a.n;
    const T2Scalar s = t2_scalar_smem(smem_raw);
    const int tid = threadIdx.x;
    for (int c = blockIdx.x; c < p.Cp; c += (int)gridDim.x) {
        if (tid < kAscCount) s.sc[tid] = sc(p, tid, c);
        if (tid >= 32 && tid < 32 + kAfCount) s.fl[tid - 32] = t2_afl(t.a, tid - 32, c);
        __syncthreads();
        const int bits = (c < p.C) ? pack_cmd(s.fl, s.fl[kAfPhase]) : 0;
        t2_late_enter(t, s, c, bits, s.sc[kScH]);
        t2_store_chain(t, s, c, bits);
    }
}

__device__ __noinline__ void tick2_scalar_generic(const Tick2Params& t, unsigned char* smem_raw, unsigned tick_global, unsigned long long* tr)
{
    const AsyncParams& a = t.a;
    const NutsParams& p = a.n;
    const FinParams& f = t.fin;
    const T2Scalar s = t2_scalar_smem(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_blk = (int)gridDim.x;
    const bool single = t.peer.world <= 1;
    const bool gw = t.lik.model == BGH_MODEL_GW_NORMAL;
    for (int c = blockIdx.x; c < p.C; c += n_blk) {
        // ---- A: chain state to shared memory; fixed-order reduction of the CTA partials
        if (tid < kAscCount) s.sc[tid] = sc(p, tid, c);
        if (tid >= 32 && tid < 32 + kAfCount) s.fl[tid - 32] = t2_afl(a, tid - 32, c);
        __syncthreads();
        {
            const unsigned used = used_quantities(pack_cmd(s.fl, s.fl[kAfPhase]));
            // likelihood sums: in fin_phase's order (bgh_fused.cu), so logp and the head gradients are bit-identical
            // to the evaluation kernel's; sampler sums: lane-strided over the CTAs + xor tree
            for (int k = warp; k < kT2Nq; k += kLikWarps) {
                const bool lik = k < kT2Lik;
                if (!lik && !((used >> (k - kT2Lik)) & 1u)) {
                    if (lane == 0) s.sum[k] = 0.0;
                    continue;
                }
                const double* const src = lik ? t.lik.part + (size_t)k * p.Cp + c : t.part + (size_t)(k - kT2Lik) * p.Cp + c;
                const size_t stride = (size_t)(lik ? kT2Lik : kAsyncPartials) * p.Cp;
                double acc = 0.0;
                for (int i0 = 0; i0 * 32 < f.n_cta; i0 += 8) {
                    double v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int row = lane + 32 * (i0 + j);
                        v[j] = row < f.n_cta ? src[(size_t)row * stride] : 0.0;
                    }
                    acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) s.sum[k] = acc;
            }
        }
        split_gene_sums(f, c, s.sr);
        __syncthreads();
        if (tr) tr[5] = global_timer_ns();
        int phase = s.fl[kAfPhase];
        const int bits_old = pack_cmd(s.fl, phase);     // the command this tick's vector work ran under
        const double h_old = s.sc[kScH];
        if (tid == 0) {
            // kinetic energy of the fresh momenta of the late rows was computed when they were refreshed (last tick)
            s.sum[kT2Lik + kAqKin0] += s.sc[kScKin0LateLocal];
        }
        __syncthreads();

        // the closing formulas read the position / write the gradient of the chain's working buffer
        FinParams fc = f;
        {
            const long long wo = (bits_old & kCbW) ? (long long)kXStride * p.D * p.Cp : 0ll;
            fc.q = f.q + wo;
            fc.grad = f.grad + wo;
        }
        const ChainConst cc = chain_const(fc, c);        // every thread: cheap, avoids a broadcast
        if (!single) {
            // ---- B: rank-local late rows (their residual sums are complete on this rank), then the exchange
            for (int k = tid; k < f.n_split; k += blockDim.x)
                if (f.split_row[k] < 0) finish_gene(fc, f.split_gene[k], c, s.sr[k], load_gene_coef(fc, f.split_gene[k], c, true));
            __syncthreads();
            if (phase == kPhaseRun) t2_late_finish(t, s, c, bits_old, h_old, false);
            // payload = the chain's sums followed by the residual sums of the replicated genes (payload row 4 + j)
            for (int k = tid; k < f.n_split; k += blockDim.x)
                if (f.split_row[k] >= 0) s.pay[kT2Nq + f.split_row[k]] = s.sr[k];
            __syncthreads();
            t2_exchange(t.peer, p.Cp, c, tick_global, s.pay, s.tmp, t.status);
            for (int k = tid; k < f.n_split; k += blockDim.x)
                if (f.split_row[k] >= 0) s.sr[k] = s.pay[kT2Nq + f.split_row[k]];
            __syncthreads();
        }
        if (tid == 0) s.sum[kT2Lik + kAqKin0] += s.sc[kScKin0LateRep];
        // ---- C: closing formulas (logp, gradient of the head rows), gradients of the (remaining) split genes
        if (tid == 0) {
            ChainSums cs;
            cs.A_ll = s.sum[0]; cs.A_e = s.sum[1]; cs.S1 = s.sum[2]; cs.S2 = s.sum[3];
            finish_chain(fc, c, cs, cc);
        }
        for (int k = tid; k < f.n_split; k += blockDim.x)
            if (single || f.split_row[k] >= 0 || gw) finish_gene(fc, f.split_gene[k], c, s.sr[k], load_gene_coef(fc, f.split_gene[k], c, true));
        __syncthreads();
        if (phase == kPhaseFinished) continue;                            // CTA-uniform
        if (phase == kPhaseFlush) {
            if (tid == 0) {
                s.fl[kAfPhase] = kPhaseFinished;
                s.fl[kAfMerge] = 0; s.fl[kAfTake] = 0; s.fl[kAfTakeTop] = 0; s.fl[kAfRecSlot] = -1;
                atomicAdd(a.n_finished, 1);
            }
            __syncthreads();
            t2_store_chain(t, s, c, 0);
            continue;
        }
        // ---- D: second half of the leapfrog for the late rows (replicated ones when sharded)
        t2_late_finish(t, s, c, bits_old, h_old, true);
        if (tr) tr[6] = global_timer_ns();

        // ---- E: the per-chain NUTS state machine (same logic as nuts_leaf_scalar / top1 / top2 / finish kernels)
        if (tid == 0) {
            const double* const ssum = s.sum + kT2Lik;
            double* const s_sc = s.sc;
            int* const s_fl = s.fl;
            long long draw = s_fl[kAfDraw];
            int depth = s_fl[kAfDepth], leaf = s_fl[kAfLeaf], dir = s_fl[kAfDir];
            if (s_fl[kAfNew]) {      // the transition began this tick
                const double e0 = ssum[kAqKin0] - s_sc[kScLogpProp];
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
            s_fl[kAfTake] = 0; s_fl[kAfAdapt] = 0; s_fl[kAfSwitch] = 0; s_fl[kAfRecSlot] = -1; s_fl[kAfFlip] = 0;

            const double kin = ssum[0];
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
                    const double u = t2_chain_uniform(p, c, draw, kPurposeLeaf, ((unsigned)depth << 16) | (unsigned)leaf);
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
                for (int l = 0; l < n_lvl; ++l) turning = turning || (ssum[1 + 2 * l] <= 0.0) || (ssum[2 + 2 * l] <= 0.0);
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
                const double z1 = s_sc[kScLogSizeTree], z2 = s_sc[kScLogSizeSub];
                const double u = t2_chain_uniform(p, c, draw, kPurposeTop, (unsigned)depth);
                if (log(u) < z2 - z1) {
                    s_fl[kAfTakeTop] = 1;
                    s_sc[kScLogpProp] = s_sc[kScLogpSub];
                    s_sc[kScEnergyProp] = s_sc[kScEnergySub];
                }
                s_sc[kScLogSizeTree] = logaddexp_d(z1, z2);
                s_fl[kAfMerge] = 1;
                s_fl[kAfDirPrev] = dir;
                depth += 1;
                if (ssum[kAqTree0] <= 0.0 || ssum[kAqTree0 + 1] <= 0.0) {
                    s_fl[kAfTurning] = 1;
                    end_transition = true;
                }
                if (depth >= max_depth) end_transition = true;
                if (!end_transition) {
                    const double ud = t2_chain_uniform(p, c, draw, kPurposeDir, (unsigned)depth);
                    const int dir_new = (ud < 0.5) ? 1 : -1;
                    if (dir_new != dir) {     // the next doubling grows the other end of the tree: its edge is the other buffer
                        s_fl[kAfFlip] = 1;
                        s_fl[kAfW] ^= 1;
                    }
                    dir = dir_new;
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
                    const double ud = t2_chain_uniform(p, c, draw, kPurposeDir, 0u);
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
            t2_schedule_leaf(s_fl, leaf, depth);
            s.misc[0] = pack_cmd(s_fl, phase);
        }
        __syncthreads();
        // ---- F: rare-event work and first half step of the next leapfrog for the late rows
        if (tr) tr[7] = global_timer_ns();
        const int bits_new = s.misc[0];
        // the rare-event path reads these from global memory
        if (tid == 0) {
            t2_afl(a, kAfRecSlot, c) = s.fl[kAfRecSlot];
            t2_afl(a, kAfDraw, c) = s.fl[kAfDraw];
            sc(p, kScFgUse, c) = s.sc[kScFgUse];
            sc(p, kScBgUse, c) = s.sc[kScBgUse];
        }
        __syncthreads();
        t2_late_enter(t, s, c, bits_new, s.sc[kScH]);
        t2_store_chain(t, s, c, bits_new);
    }
}

// ---------------------------------------------------------------------------------------------
// Scalar phase, latency-optimised path (at most one late row per thread, i.e. n_late <= 512 — always, unless hundreds of
// genes are larger than an SNP range).  Same arithmetic and summation orders as tick2_scalar_generic; what differs is that
// nothing waits for a global round trip it does not need: the tick's command comes from the CTA's shared-memory copy, every
// thread fetches the state of "its" late row while the partial sums are being reduced, gradients and logp travel through
// shared memory instead of global memory, and a late row that simply continues its trajectory starts the next half step
// from registers.
// s_lrow: late-row table (persistent shared memory): local row | replicated << 30.
// ---------------------------------------------------------------------------------------------
constexpr int kLateRep = 1 << 30;

__device__ __forceinline__ void tick2_scalar(const Tick2Params& t, unsigned char* smem_raw, unsigned tick_global, unsigned long long* tr,
                                             const int* s_bits, const double* s_h, const int* s_lrow, int n_late, const int* s_sptr,
                                             const int* s_sidx, bool split_cached)
{
    const AsyncParams& a = t.a;
    const NutsParams& p = a.n;
    const FinParams& f = t.fin;
    const T2Scalar s = t2_scalar_smem(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_blk = (int)gridDim.x;
    const bool single = t.peer.world <= 1;
    const bool gw = t.lik.model == BGH_MODEL_GW_NORMAL;
    const size_t vec = (size_t)p.D * p.Cp;
    const unsigned long long pol = l2_keep_policy();
    for (int c = blockIdx.x; c < p.C; c += n_blk) {
        const int bits_old = s_bits[c];                  // the command this tick's vector work ran under
        const double h_old = s_h[c];
        const bool run = (bits_old & kCbRun) != 0;
        const size_t wo = (bits_old & kCbW) ? (size_t)kXStride * vec : 0;
        const unsigned used = used_quantities(bits_old);
        const int nl = (bits_old >> kCbNLvlShift) & 15, lm = (bits_old >> kCbLvlMinShift) & 15;
        // ---- A: everything this phase reads from global memory is requested now
        // (a warp issues in order: a load whose value is STORED right away would hold back every later load of the warp by a
        // round trip, so the per-chain scalars / flags and the first slots of the thread's split gene are only loaded here and
        // consumed after the partial sums' loads have been issued)
        const double sc_v = tid < kAscCount ? sc(p, tid, c) : 0.0;
        const int fl_v = (tid >= 32 && tid < 32 + kAfCount) ? t2_afl(a, tid - 32, c) : 0;
        constexpr int kSlotsAhead = 4;
        double sv[kSlotsAhead];
        const int k_mine = f.n_long + tid;
        const bool slots_mine = split_cached && k_mine < f.n_split;
        const int sl0 = slots_mine ? s_sptr[k_mine] : 0, sln = slots_mine ? s_sptr[k_mine + 1] - sl0 : 0;
#pragma unroll
        for (int j = 0; j < kSlotsAhead; ++j) sv[j] = j < sln ? f.slots[(size_t)s_sidx[sl0 + j] * p.Cp + c] : 0.0;
        const bool has_row = tid < n_late;
        const int lr = has_row ? s_lrow[tid] : 0;
        const int d = lr & (kLateRep - 1);
        const bool rep_row = single || (lr & kLateRep) != 0;           // its gradient needs the exchanged totals
        const size_t i = (size_t)d * p.Cp + c;
        double qv = 0.0, pw = 0.0, var = 0.0, tot = 0.0;
        double2 ck0 = make_double2(0.0, 0.0);
        if (has_row && run) {
            qv = ldk(p.state + (size_t)kSlotQL * vec + wo + i, pol);
            pw = ldk(p.state + (size_t)kSlotPL * vec + wo + i, pol);
            var = ldk(p.state + (size_t)kSlotVar * vec + i, pol);
            tot = ldk(p.state + (size_t)kSlotT * vec + i, pol);
            if (nl > 0) ck0 = *t2_ckpt(p, lm, i);
            for (int l = 1; l < nl; ++l) prefetch_l1(t2_ckpt(p, lm + l, i));     // deep checks: every level in flight now
            if (bits_old & kCbLast) prefetch_l1(p.state + (size_t)kSlotPL * vec + (wo ? 0 : (size_t)kXStride * vec) + i);
        }
        FinParams fc = f;                                // closing formulas: the chain's working buffer
        fc.q = f.q + wo;
        fc.grad = f.grad + wo;
        // the chain's closing constants (two exp / log1p chains on the FP64 pipe, needed by thread 0 only, in part C): ONE thread
        // of the warp with the lightest reduction duty computes them under the loads of part A and hands them over in shared
        // memory (all 16 warps computing them cost ~1.5 us of this latency-bound phase)
        if (tid == kLikThreads - 32) {
            const ChainConst k = chain_const(fc, c);
            s.red2[0] = k.iW; s.red2[1] = k.iW2; s.red2[2] = k.mi; s.red2[3] = k.lp0; s.red2[4] = k.x0; s.red2[5] = k.x1;
        }
        GeneCoef gc;
        gc.iW2 = gc.mi = gc.beta = 0.0;
        const int k_split = (!gw && d >= t.row0) ? tid - t.row0 : (gw && tid == 1 ? 0 : -1);     // split-gene index of this thread's row
        if (has_row && k_split >= 0) gc = load_gene_coef(fc, f.split_gene[k_split], c, true);
        // likelihood sums in fin_phase's order (bit-identical logp / head gradients); sampler sums lane-strided + xor tree.
        // Warp w reduces quantities w and w + 16: the loads of both are issued before either is summed (one round trip).
        {
            const int k0 = warp, k1 = warp + kLikWarps;
            const bool on0 = k0 < kT2Lik || ((used >> (k0 - kT2Lik)) & 1u);
            const bool on1 = k1 < kT2Nq && ((used >> (k1 - kT2Lik)) & 1u);
            const double* const src0 = k0 < kT2Lik ? t.lik.part + (size_t)k0 * p.Cp + c : t.part + (size_t)(k0 - kT2Lik) * p.Cp + c;
            const size_t st0 = (size_t)(k0 < kT2Lik ? kT2Lik : kAsyncPartials) * p.Cp;
            const double* const src1 = t.part + (size_t)(k1 - kT2Lik) * p.Cp + c;
            const size_t st1 = (size_t)kAsyncPartials * p.Cp;
            double a0 = 0.0, a1 = 0.0;
            for (int i0 = 0; i0 * 32 < f.n_cta; i0 += 8) {
                double v[8], w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int row = lane + 32 * (i0 + j);
                    v[j] = (on0 && row < f.n_cta) ? src0[(size_t)row * st0] : 0.0;
                    w[j] = (on1 && row < f.n_cta) ? src1[(size_t)row * st1] : 0.0;
                }
                if (i0 == 0) {      // every load of part A is in flight: hand over the early ones
                    if (tid < kAscCount) s.sc[tid] = sc_v;
                    if (tid >= 32 && tid < 32 + kAfCount) s.fl[tid - 32] = fl_v;
                }
                a0 += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
                a1 += ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            }
            if (lane == 0) {
                s.sum[k0] = a0;
                if (k1 < kT2Nq) s.sum[k1] = a1;
            }
        }
        if (split_cached) {
            if (slots_mine) {      // the slots in list order, as split_gene_sums_cached adds them
                double tot = 0.0;
#pragma unroll
                for (int j = 0; j < kSlotsAhead; ++j)
                    if (j < sln) tot += sv[j];
                for (int j = kSlotsAhead; j < sln; ++j) tot += f.slots[(size_t)s_sidx[sl0 + j] * p.Cp + c];
                s.sr[k_mine] = tot;
            }
            split_gene_sums_cached(f, c, s.sr, s_sptr, s_sidx, f.n_long + (int)blockDim.x);
        } else {
            split_gene_sums(f, c, s.sr);
        }
        __syncthreads();
        if (tr) tr[5] = global_timer_ns();
        if (tid == 0) s.sum[kT2Lik + kAqKin0] += s.sc[kScKin0LateLocal];   // fresh momenta of the late rows (refreshed last tick)

        double kin = 0.0, gnew = 0.0, mom_new = pw;
        // per-thread accumulators of the row's contributions: shared memory, thread-interleaved [quantity][thread] behind the
        // (at most 512) split-gene sums — a local-memory array costs a cache round trip per quantity on this latency path
        double* const acc = s.sr + kLikThreads + tid;
        constexpr int kAs = kLikThreads;
        static_assert((32 * 16 + kT2MaxPay + 32 + 16 + kLikWarps + kLikWarps * kAsyncPartials + kMaxPeers * kT2MaxPay + 2 + kLikThreads +
                       kAsyncPartials * kLikThreads) * sizeof(double) <= kLikStageBytes, "scalar phase accumulators");
        const unsigned u_acc = used & ~(1u << kAqKin0);
#pragma unroll 1
        for (int j = 1; j < kAsyncPartials; ++j)
            if ((u_acc >> j) & 1u) acc[j * kAs] = 0.0;
        // fixed-order CTA sum of the used quantities of acc[] into s.sum,
        // the quantities in parallel: warp w sums quantities w and w + 16 over the 512 per-thread slots (lane-strided in a fixed
        // order, then an xor tree) — one pass for all used quantities instead of a shuffle tree per quantity in sequence
        // (a deep tree check uses 23 of them: that sequence was the tail of the slowest chain's scalar phase)
        auto reduce_acc = [&]() {
            const unsigned u = u_acc;
            acc[0] = kin;
            __syncthreads();
#pragma unroll 1
            for (int j = warp; j < kAsyncPartials; j += kLikWarps) {
                if (!((u >> j) & 1u)) continue;
                const double* const src = s.sr + kLikThreads + (size_t)j * kAs;
                double v[kLikThreads / 32];
#pragma unroll
                for (int k = 0; k < kLikThreads / 32; ++k) v[k] = src[lane + 32 * k];
                double x = 0.0;
#pragma unroll
                for (int k = 0; k < kLikThreads / 32; ++k) x += v[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                if (lane == 0) s.sum[kT2Lik + j] += x;
            }
            __syncthreads();
            kin = 0.0;
#pragma unroll 1
            for (int j = 1; j < kAsyncPartials; ++j)
                if ((u >> j) & 1u) acc[j * kAs] = 0.0;
        };
        if (!single) {
            // ---- B: rank-local late rows (their residual sums are complete on this rank), then the exchange
            if (has_row && !rep_row) {
                gnew = gene_grad(fc, s.sr[k_split], gc);
                stk(p.state + (size_t)kSlotGL * vec + wo + i, gnew, pol);
                if (run) mom_new = row_finish(p, bits_old, h_old, i, wo, gnew, pw, var, tot, ck0, kin, acc, kAs, pol);
            }
            if (run) reduce_acc();
            // payload = the chain's sums followed by the residual sums of the replicated genes (payload row 4 + j)
            for (int k = tid; k < f.n_split; k += blockDim.x)
                if (f.split_row[k] >= 0) s.pay[kT2Nq + f.split_row[k]] = s.sr[k];
            __syncthreads();
            if (tr) tr[11] = global_timer_ns();
            t2_exchange(t.peer, p.Cp, c, tick_global, s.pay, s.tmp, t.status);
            if (tr) tr[12] = global_timer_ns();
            for (int k = tid; k < f.n_split; k += blockDim.x)
                if (f.split_row[k] >= 0) s.sr[k] = s.pay[kT2Nq + f.split_row[k]];
            __syncthreads();
        }
        // ---- C: closing formulas (logp, gradient of the head rows) through shared memory
        double* const s_hg = s.red;                      // [0..2] head gradients, [3] logp (s.red is free here)
        if (tid == 0) {
            s.sum[kT2Lik + kAqKin0] += s.sc[kScKin0LateRep];
            ChainSums cs;
            cs.A_ll = s.sum[0]; cs.A_e = s.sum[1]; cs.S1 = s.sum[2]; cs.S2 = s.sum[3];
            double lp, g[3];
            ChainConst cc;
            cc.iW = s.red2[0]; cc.iW2 = s.red2[1]; cc.mi = s.red2[2]; cc.lp0 = s.red2[3]; cc.x0 = s.red2[4]; cc.x1 = s.red2[5];
            chain_closing(fc, c, cs, cc, lp, g);
            f.logp[c] = lp;
            s_hg[0] = g[0]; s_hg[1] = g[1]; s_hg[2] = g[2]; s_hg[3] = lp;
        }
        __syncthreads();
        if (!run) {                                                       // CTA-uniform: finished or flushing chain
            if (has_row && rep_row) {
                const double g = k_split >= 0 ? gene_grad(fc, s.sr[k_split], gc) : s_hg[d];
                stk(p.state + (size_t)kSlotGL * vec + wo + i, g, pol);
            }
            if (bits_old & kCbFlush) {
                if (tid == 0) {
                    s.fl[kAfPhase] = kPhaseFinished;
                    s.fl[kAfMerge] = 0; s.fl[kAfTake] = 0; s.fl[kAfTakeTop] = 0; s.fl[kAfRecSlot] = -1;
                    atomicAdd(a.n_finished, 1);
                }
                __syncthreads();
                t2_store_chain(t, s, c, 0);
            }
            __syncthreads();
            continue;
        }
        // ---- D: second half of the leapfrog for the late rows (all of them on a single rank, the replicated ones when sharded)
        if (has_row && rep_row) {
            gnew = k_split >= 0 ? gene_grad(fc, s.sr[k_split], gc) : s_hg[d];
            stk(p.state + (size_t)kSlotGL * vec + wo + i, gnew, pol);
            mom_new = row_finish(p, bits_old, h_old, i, wo, gnew, pw, var, tot, ck0, kin, acc, kAs, pol);
        }
        const double logp_now = s_hg[3];
        reduce_acc();
        if (tr) tr[6] = global_timer_ns();

        // ---- E: the per-chain NUTS state machine (same logic as nuts_leaf_scalar / top1 / top2 / finish kernels)
        if (tid == 0) {
            int phase = kPhaseRun;
            const double* const ssum = s.sum + kT2Lik;
            double* const s_sc = s.sc;
            int* const s_fl = s.fl;
            long long draw = s_fl[kAfDraw];
            int depth = s_fl[kAfDepth], leaf = s_fl[kAfLeaf], dir = s_fl[kAfDir];
            if (s_fl[kAfNew]) {      // the transition began this tick
                const double e0 = ssum[kAqKin0] - s_sc[kScLogpProp];
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
            s_fl[kAfTake] = 0; s_fl[kAfAdapt] = 0; s_fl[kAfSwitch] = 0; s_fl[kAfRecSlot] = -1; s_fl[kAfFlip] = 0;

            const double kin_tot = ssum[0];
            const double logp = logp_now;
            const double energy = kin_tot - logp;
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
                    const double u = t2_chain_uniform(p, c, draw, kPurposeLeaf, ((unsigned)depth << 16) | (unsigned)leaf);
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
                for (int l = 0; l < n_lvl; ++l) turning = turning || (ssum[1 + 2 * l] <= 0.0) || (ssum[2 + 2 * l] <= 0.0);
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
                const double z1 = s_sc[kScLogSizeTree], z2 = s_sc[kScLogSizeSub];
                const double u = t2_chain_uniform(p, c, draw, kPurposeTop, (unsigned)depth);
                if (log(u) < z2 - z1) {
                    s_fl[kAfTakeTop] = 1;
                    s_sc[kScLogpProp] = s_sc[kScLogpSub];
                    s_sc[kScEnergyProp] = s_sc[kScEnergySub];
                }
                s_sc[kScLogSizeTree] = logaddexp_d(z1, z2);
                s_fl[kAfMerge] = 1;
                s_fl[kAfDirPrev] = dir;
                depth += 1;
                if (ssum[kAqTree0] <= 0.0 || ssum[kAqTree0 + 1] <= 0.0) {
                    s_fl[kAfTurning] = 1;
                    end_transition = true;
                }
                if (depth >= max_depth) end_transition = true;
                if (!end_transition) {
                    const double ud = t2_chain_uniform(p, c, draw, kPurposeDir, (unsigned)depth);
                    const int dir_new = (ud < 0.5) ? 1 : -1;
                    if (dir_new != dir) {     // the next doubling grows the other end of the tree: its edge is the other buffer
                        s_fl[kAfFlip] = 1;
                        s_fl[kAfW] ^= 1;
                    }
                    dir = dir_new;
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
                    const double ud = t2_chain_uniform(p, c, draw, kPurposeDir, 0u);
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
            t2_schedule_leaf(s_fl, leaf, depth);
            s.misc[0] = pack_cmd(s_fl, phase);
            if (cmd_rare(s.misc[0])) {      // the rare-event path reads these from global memory
                t2_afl(a, kAfRecSlot, c) = s_fl[kAfRecSlot];
                t2_afl(a, kAfDraw, c) = s_fl[kAfDraw];
                sc(p, kScFgUse, c) = s_sc[kScFgUse];
                sc(p, kScBgUse, c) = s_sc[kScBgUse];
            }
        }
        __syncthreads();
        if (tr) tr[7] = global_timer_ns();
        // ---- F: rare-event work and first half step of the next leapfrog for the late rows
        const int bits_new = s.misc[0];
        const double h_new = s.sc[kScH];
        if (cmd_rare(bits_new) || (bits_new & kCbFlip)) {
            t2_late_enter(t, s, c, bits_new, h_new);          // generic path: reloads the rows from global memory
        } else {
            // the chain simply continues: the row's state is still in registers
            if (has_row) row_half_step(p, bits_new, h_new, i, wo, qv, mom_new, gnew, var, pol);
            if (tid == 0) { s.sc[kScKin0LateLocal] = 0.0; s.sc[kScKin0LateRep] = 0.0; }
            __syncthreads();
        }
        t2_store_chain(t, s, c, bits_new);
    }
}

// ---------------------------------------------------------------------------------------------
// Sharded evaluation (bgh_logp_grad_sharded on contexts with replicated genes): likelihood pass | grid barrier | per-chain
// closing with the peer exchange, one launch per rank.  The chain CTAs reduce the CTA partials and the slot sums of the
// split genes themselves, so the payload of a chain (4 sums + the replicated genes' residual sums) is complete inside
// ONE CTA: no second barrier, no single-CTA all-reduce — every chain's LL message leaves as soon as its CTA is done.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void eval2_close(const Eval2Params& e, unsigned char* smem_raw)
{
    const FinParams& f = e.fin;
    const T2Scalar s = t2_scalar_smem(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Cp = f.Cp;
    const bool gw = f.model == BGH_MODEL_GW_NORMAL;
    for (int c = blockIdx.x; c < e.C; c += (int)gridDim.x) {
        // likelihood sums in fin_phase's order (bit-identical to the single-rank kernels for a single rank)
        if (warp < kT2Lik) {
            double acc = 0.0;
            for (int i0 = 0; i0 * 32 < f.n_cta; i0 += 8) {
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int row = lane + 32 * (i0 + j);
                    v[j] = row < f.n_cta ? f.part[((size_t)row * kT2Lik + warp) * Cp + c] : 0.0;
                }
                acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) s.sum[warp] = acc;
        }
        split_gene_sums(f, c, s.sr);
        __syncthreads();
        const bool single = e.peer.world <= 1;
        if (!single) {
            for (int k = tid; k < f.n_split; k += blockDim.x) {
                if (f.split_row[k] < 0) finish_gene(f, f.split_gene[k], c, s.sr[k], load_gene_coef(f, f.split_gene[k], c, true));
                else s.pay[kT2Lik + f.split_row[k]] = s.sr[k];
            }
            __syncthreads();
            t2_exchange(e.peer, Cp, c, e.epoch, s.pay, s.tmp, e.status);
            for (int k = tid; k < f.n_split; k += blockDim.x)
                if (f.split_row[k] >= 0) s.sr[k] = s.pay[kT2Lik + f.split_row[k]];
            __syncthreads();
        }
        if (tid == 0) {
            ChainSums cs;
            cs.A_ll = s.sum[0]; cs.A_e = s.sum[1]; cs.S1 = s.sum[2]; cs.S2 = s.sum[3];
            finish_chain(f, c, cs, chain_const(f, c));
        }
        for (int k = tid; k < f.n_split; k += blockDim.x)
            if (single || f.split_row[k] >= 0 || gw) finish_gene(f, f.split_gene[k], c, s.sr[k], load_gene_coef(f, f.split_gene[k], c, true));
        __syncthreads();
    }
}

template <int CL, int CPL, bool GAMMA>
__global__ void __launch_bounds__(kLikThreads, 1) logp_grad_chainx_kernel(const __grid_constant__ Eval2Params e)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    lik_stage_init(smem_raw);
    __syncthreads();
    uint32_t par = 0;
    unsigned long long target = e.bar_base;
    for (int cb = 0; cb < e.chain_blocks; ++cb) lik_phase<CL, CPL, GAMMA>(e.lik, smem_raw, blockIdx.x, cb, par);
    grid_barrier(e.bar, target, e.status);
    eval2_close(e, smem_raw);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int CL, int CPL, bool GAMMA>
__global__ void __launch_bounds__(kLikThreads, 1) nuts_tick2_kernel(const __grid_constant__ Tick2Params t)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_stop;
    __shared__ unsigned s_need;
    __shared__ int s_nnp;
    lik_stage_init(smem_raw);
    uint32_t par = 0;
    unsigned long long target = 0;      // the host zeroes the barrier word before every launch
    const NutsParams& p = t.a.n;
    double* const s_h = reinterpret_cast<double*>(smem_raw + kT2CmdOffset);
    int* const s_bits = reinterpret_cast<int*>(s_h + kT2MaxChains);
    double* const s_kin0c = reinterpret_cast<double*>(smem_raw + kT2Kin0Offset);     // [kT2MaxChains]
    int* const s_np = reinterpret_cast<int*>(s_kin0c + kT2MaxChains);                  // [kT2MaxChains]
    int* const s_rows = reinterpret_cast<int*>(smem_raw + kT2RowsOffset);
    int* const s_lrow = reinterpret_cast<int*>(smem_raw + kT2LateOffset);
    const int n_late = t2_n_late(t);
    if (n_late <= kLikThreads)
        for (int r = threadIdx.x; r < n_late; r += kLikThreads) s_lrow[r] = t2_late_row(t, r) | (t2_late_replicated(t, r) ? kLateRep : 0);
    // slot lists of the split genes (constant for the run)
    int* const s_sptr = reinterpret_cast<int*>(smem_raw + kT2SplitOffset);
    int* const s_sidx = s_sptr + kLikThreads + 1;
    const int n_split = t.fin.n_split;
    const bool split_cached = n_split <= kLikThreads && t.fin.split_ptr[n_split] <= kT2MaxSlotsSmem;
    if (split_cached) {
        for (int k = threadIdx.x; k <= n_split; k += kLikThreads) s_sptr[k] = t.fin.split_ptr[k];
        for (int i = threadIdx.x; i < t.fin.split_ptr[n_split]; i += kLikThreads) s_sidx[i] = t.fin.split_slots[i];
    }
    // the rows this CTA owns (constant for the whole run): gene ids into shared memory
    T2Own own;
    {
        const int i0 = t.cta_row_ptr[blockIdx.x];
        own.n = t.cta_row_ptr[blockIdx.x + 1] - i0;
        if (own.n <= kT2MaxRowsSmem) {
            for (int i = threadIdx.x; i < own.n; i += kLikThreads) s_rows[i] = t.cta_row_gene[i0 + i];
            own.rows = s_rows;
        } else {
            own.rows = t.cta_row_gene + i0;
        }
    }
    __syncthreads();
    if (t.first_launch) {
        tick2_prologue(t, smem_raw);
        grid_barrier(t.bar, target, t.status);
    }
    for (int tick = 0; tick < t.n_ticks; ++tick) {
        const int traced = t.tick_base + tick - t.trace_first;
        unsigned long long* const tr = (t.phase_trace != nullptr && blockIdx.x == t.trace_cta && threadIdx.x == 0 && traced >= 0 && traced < 64)
                                           ? t.phase_trace + (size_t)traced * 16 : nullptr;
        if (tr) tr[0] = global_timer_ns();
        // command words of the chains -> shared memory (written by the scalar phase before the last barrier)
        if (threadIdx.x == 0) { s_need = 0u; s_nnp = 0; }
        __syncthreads();
        for (int c = threadIdx.x; c < p.Cp; c += kLikThreads) {
            const int b = c < p.C ? t.cmd_bits[c] : 0;
            s_bits[c] = b;
            s_h[c] = t.cmd_h[c];
            s_kin0c[c] = 0.0;
            const unsigned u = used_quantities(b);
            if (u) atomicOr(&s_need, u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {     // chains with rare-event work, in chain order (ballot + prefix count per 32 chains)
            int n = 0;
            for (int c0 = 0; c0 < p.C; c0 += 32) {
                const int c = c0 + threadIdx.x;
                const bool rare = c < p.C && cmd_rare(s_bits[c]);
                const unsigned m = __ballot_sync(0xffffffffu, rare);
                if (rare) s_np[n + __popc(m & ((1u << threadIdx.x) - 1u))] = c;
                n += __popc(m);
            }
            if (threadIdx.x == 0) s_nnp = n;
        }
        __syncthreads();
        if (tr) { tr[11] = global_timer_ns(); tr[14] = (unsigned long long)s_nnp; tr[15] = (unsigned long long)own.n; }
        if (s_nnp > 0 && own.n > 0) tick2_rare(t, own, s_bits, s_np, s_nnp, reinterpret_cast<double*>(smem_raw), s_kin0c, tr);
        if (tr) tr[10] = global_timer_ns();
        tick2_pre(t, own, s_bits, s_h);
        __syncthreads();
        if (tr) tr[8] = global_timer_ns();
        fence_proxy_async();   // the staging area was written through the generic proxy (scalar phase, accumulators)
        for (int cb = 0; cb < t.chain_blocks; ++cb) lik_phase<CL, CPL, GAMMA>(t.lik, smem_raw, blockIdx.x, cb, par);
        if (tr) tr[9] = global_timer_ns();
        tick2_post(t, own, s_bits, s_h, s_kin0c, s_need, smem_raw);
        if (tr) tr[1] = global_timer_ns();
        grid_barrier(t.bar, target, t.status);
        if (tr) tr[2] = global_timer_ns();
        if (n_late <= kLikThreads) tick2_scalar(t, smem_raw, (unsigned)(t.tick_base + tick), tr, s_bits, s_h, s_lrow, n_late, s_sptr, s_sidx, split_cached);
        else tick2_scalar_generic(t, smem_raw, (unsigned)(t.tick_base + tick), tr);
        if (tr) tr[3] = global_timer_ns();
        grid_barrier(t.bar, target, t.status);
        if (tr) tr[4] = global_timer_ns();
        if (threadIdx.x == 0) {
            s_stop = (ld_volatile_i32(t.a.n_finished) >= p.C || ld_volatile_i32(t.status) != 0) ? 1 : 0;
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
static cudaError_t configure_tick2_t()
{
    cudaError_t e = cudaFuncSetAttribute((const void*)nuts_tick2_kernel<CL, CPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kT2SmemBytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute((const void*)nuts_tick2_kernel<CL, CPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kT2SmemBytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute((const void*)logp_grad_chainx_kernel<CL, CPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLikSmemBytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute((const void*)logp_grad_chainx_kernel<CL, CPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLikSmemBytes);
    return e;
}

cudaError_t tick2_configure()
{
    cudaError_t e;
    if ((e = configure_tick2_t<4, 1>()) != cudaSuccess) return e;
    if ((e = configure_tick2_t<8, 1>()) != cudaSuccess) return e;
    if ((e = configure_tick2_t<16, 1>()) != cudaSuccess) return e;
    if ((e = configure_tick2_t<32, 1>()) != cudaSuccess) return e;
    if ((e = configure_tick2_t<32, 2>()) != cudaSuccess) return e;
    return cudaSuccess;
}

template <int CL, int CPL>
static cudaError_t launch_tick2_t(const Tick2Params& t, int n_cta, cudaStream_t s)
{
    void* args[] = {const_cast<Tick2Params*>(&t)};
    const void* fn = t.lik.model == BGH_MODEL_GENE_GAMMA ? (const void*)nuts_tick2_kernel<CL, CPL, true>
                                                         : (const void*)nuts_tick2_kernel<CL, CPL, false>;
    return cudaLaunchCooperativeKernel(fn, dim3(n_cta), dim3(kLikThreads), args, kT2SmemBytes, s);
}

template <int CL, int CPL>
static cudaError_t launch_eval2_t(const Eval2Params& e, int n_cta, cudaStream_t s)
{
    void* args[] = {const_cast<Eval2Params*>(&e)};
    const void* fn = e.lik.model == BGH_MODEL_GENE_GAMMA ? (const void*)logp_grad_chainx_kernel<CL, CPL, true>
                                                         : (const void*)logp_grad_chainx_kernel<CL, CPL, false>;
    return cudaLaunchCooperativeKernel(fn, dim3(n_cta), dim3(kLikThreads), args, kLikSmemBytes, s);
}

cudaError_t launch_eval2(const Eval2Params& e, int CL, int CPL, int n_cta, cudaStream_t s)
{
    if (CL == 4 && CPL == 1) return launch_eval2_t<4, 1>(e, n_cta, s);
    if (CL == 8 && CPL == 1) return launch_eval2_t<8, 1>(e, n_cta, s);
    if (CL == 16 && CPL == 1) return launch_eval2_t<16, 1>(e, n_cta, s);
    if (CL == 32 && CPL == 1) return launch_eval2_t<32, 1>(e, n_cta, s);
    if (CL == 32 && CPL == 2) return launch_eval2_t<32, 2>(e, n_cta, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_tick2(const Tick2Params& t, int CL, int CPL, int n_cta, cudaStream_t s)
{
    if (CL == 4 && CPL == 1) return launch_tick2_t<4, 1>(t, n_cta, s);
    if (CL == 8 && CPL == 1) return launch_tick2_t<8, 1>(t, n_cta, s);
    if (CL == 16 && CPL == 1) return launch_tick2_t<16, 1>(t, n_cta, s);
    if (CL == 32 && CPL == 1) return launch_tick2_t<32, 1>(t, n_cta, s);
    if (CL == 32 && CPL == 2) return launch_tick2_t<32, 2>(t, n_cta, s);
    return cudaErrorInvalidValue;
}

}  // namespace bgh
