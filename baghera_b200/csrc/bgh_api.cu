// bgh_api.cu — context management, host-side work planning and the C ABI (include/baghera_b200.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <chrono>
#include <string>
#include <vector>

#include "bgh_fused.h"
#include "bgh_internal.h"
#include "bgh_sampler.h"
#include "bgh_tick2.h"

namespace bgh {

// ---------------------------------------------------------------------------------------------
// work planning (pure host code)
// ---------------------------------------------------------------------------------------------
// BGH_PLAN_SNAP overrides (tuning only): fraction of a range's length within which a cut snaps to a gene boundary
static double plan_snap()
{
    static const double v = [] { const char* e = getenv("BGH_PLAN_SNAP"); return (e && *e) ? atof(e) : 0.0625; }();
    return v;
}

int build_plan(int64_t M, const int32_t* gid, int32_t G, int n_cta, int groups_per_cta,
               bool first_partial, bool last_partial, int first_row, int last_row, double kappa,
               Plan* plan, std::string* err, int align, int n_replicated, bool owns_replicated, double snap, double kappa_warp)
{
    const int n_groups = n_cta * groups_per_cta;
    if (M < 0 || M > 0x7fffff00LL || n_cta <= 0 || groups_per_cta <= 0 || G <= 0) {
        *err = "build_plan: bad sizes";
        return BGH_EINVAL;
    }
    // gene offsets (CSR) + validation: sorted, compact ids
    std::vector<int32_t> goff((size_t)G + 1, 0);
    for (int64_t j = 0; j < M; ++j) {
        const int32_t g = gid[j];
        if (g < 0 || g >= G) {
            *err = "build_plan: gene id out of range";
            return BGH_EINVAL;
        }
        if (j > 0 && g < gid[j - 1]) {
            *err = "build_plan: gene ids not sorted";
            return BGH_EINVAL;
        }
        goff[(size_t)g + 1]++;
    }
    for (int32_t g = 0; g < G; ++g) {
        if (M > 0 && goff[(size_t)g + 1] == 0) {
            *err = "build_plan: gene without SNPs (ids must be compact)";
            return BGH_EINVAL;
        }
        goff[(size_t)g + 1] += goff[g];
    }

    Plan& P = *plan;
    P = Plan();
    P.n_groups = n_groups;
    P.n_cta = n_cta;
    P.off.assign((size_t)n_groups + 1, 0);
    P.flags.assign((size_t)n_groups, 0);

    // ---- regions: a "huge" gene (>= 1.5 CTAs' worth of SNPs, e.g. NonCoding = 45% of all SNPs)
    // gets CTAs of its own, whose groups all accumulate the same gene and are reduced inside the
    // CTA (one partial per CTA instead of one per group); maximal runs of the other genes form
    // "mixed" regions.  CTAs are apportioned to regions by cost = SNPs + kappa * gene boundaries
    // (a boundary costs a flush + a slope reload, about kappa SNPs' worth of issue slots).
    struct Region { int64_t a, b; int32_t gene; int n; int32_t g0, g1; double w; };   // gene = -1 for mixed
    std::vector<Region> regions;
    if (!(kappa >= 0.0)) kappa = 0.0;
    const double per_cta = (double)M / (double)n_cta;
    {
        int64_t run_start = 0;
        int32_t run_g = 0;
        for (int32_t g = 0; g < G; ++g) {
            const int64_t sz = goff[(size_t)g + 1] - goff[g];
            // a replicated gene (sharded runs) always gets CTAs of its own: its sum then travels through a CTA slot into the payload
            if (g < n_replicated || (n_cta > 1 && (double)sz >= 1.5 * per_cta && sz >= 4 * (int64_t)groups_per_cta)) {
                if (goff[g] > run_start) regions.push_back({run_start, goff[g], -1, 0, run_g, g, 0.0});
                regions.push_back({goff[g], goff[(size_t)g + 1], g, 0, g, g + 1, 0.0});
                run_start = goff[(size_t)g + 1];
                run_g = g + 1;
            }
        }
        if (M > run_start || regions.empty()) regions.push_back({run_start, M, -1, 0, run_g, G, 0.0});
        if ((int)regions.size() > n_cta) {   // degenerate (tiny n_cta): one mixed region
            if (n_replicated > 0) {
                *err = "build_plan: more replicated genes than CTAs";
                return BGH_EINVAL;
            }
            regions.clear();
            regions.push_back({0, M, -1, 0, 0, G, 0.0});
        }
    }
    {
        double w_total = 0.0;
        for (auto& r : regions) {
            r.w = (double)(r.b - r.a) + (r.gene >= 0 ? 0.0 : kappa * (double)(r.g1 - r.g0));
            w_total += r.w;
        }
        const double w_per_cta = std::max(w_total / (double)n_cta, 1e-9);
        int total = 0;
        for (auto& r : regions) {
            r.n = std::max(1, (int)llround(r.w / w_per_cta));
            total += r.n;
        }
        while (total > n_cta) {   // take from the most over-provisioned region
            int best = -1;
            double best_v = -1.0;
            for (size_t i = 0; i < regions.size(); ++i) {
                if (regions[i].n <= 1) continue;
                const double v = (double)regions[i].n / std::max(regions[i].w, 1e-9);
                if (v > best_v) { best_v = v; best = (int)i; }
            }
            regions[(size_t)best].n--;
            total--;
        }
        while (total < n_cta) {   // give to the most loaded region
            int best = 0;
            double best_v = -1.0;
            for (size_t i = 0; i < regions.size(); ++i) {
                const double v = regions[i].w / (double)regions[i].n;
                if (v > best_v) { best_v = v; best = (int)i; }
            }
            regions[(size_t)best].n++;
            total++;
        }
    }

    // ---- group ranges
    std::vector<std::vector<int32_t>> lists;   // per split gene, in gene order
    std::vector<int32_t> list_gene;
    auto add_slot = [&](int32_t gene, int32_t slot) {
        if (list_gene.empty() || list_gene.back() != gene) {
            list_gene.push_back(gene);
            lists.emplace_back();
        }
        lists.back().push_back(slot);
    };
    int v0 = 0, cta0 = 0;
    for (const Region& r : regions) {
        const int ng = r.n * groups_per_cta;
        const double len = (double)(r.b - r.a) / (double)ng;
        if (r.gene >= 0) {
            // exclusive CTAs: equal cuts, every group flagged kExclusive; one slot per CTA
            const bool owned_here = !(r.gene == 0 && first_partial) && !(r.gene < n_replicated && !owns_replicated);
            for (int k = 0; k < ng; ++k) {
                const int64_t x = r.a + (int64_t)llround((double)k * len);
                P.off[(size_t)(v0 + k)] = (int32_t)x;
                P.flags[(size_t)(v0 + k)] = kExclusive;
            }
            // the prior terms of the gene belong to the (first non-empty) group holding its first SNP
            for (int k = 0; k < ng && owned_here; ++k) {
                const int64_t xa = P.off[(size_t)(v0 + k)];
                const int64_t xb = (k + 1 < ng) ? (int64_t)P.off[(size_t)(v0 + k + 1)] : r.b;
                if (xb > xa) {
                    P.flags[(size_t)(v0 + k)] |= kOwner;
                    break;
                }
            }
            for (int c = 0; c < r.n; ++c) add_slot(r.gene, 2 * n_groups + cta0 + c);
        } else {
            // mixed region: ideal equal-cost cuts snapped to the nearest gene boundary within snap * len, so that only
            // genes longer than 2 * snap * len are ever split across groups (a split gene is a "late row" of the
            // sampler kernel, bgh_tick2.cu: its leapfrog runs in the per-chain scalar phase)
            const int64_t tol = (int64_t)floor(len * (snap > 0.0 ? snap : plan_snap()));
            int64_t prev = r.a;
            // cost of [r.a, x): SNPs + kappa * genes started; cuts equalise the cost
            auto cost = [&](int64_t x) {
                const int64_t ng_started = std::lower_bound(goff.begin() + r.g0, goff.begin() + r.g1, (int32_t)x) - (goff.begin() + r.g0);
                return (double)(x - r.a) + kappa * (double)ng_started;
            };
            // Two-level cuts (kappa_warp >= 0, the sampler plan): the CTA cuts equalise SNPs + kappa * genes (a gene also costs
            // the leapfrog of its row, which ALL threads of the CTA share), the cuts between the warps of a CTA equalise
            // SNPs + kappa_warp * genes (what a boundary costs inside the likelihood pass) — with one kappa for both a warp
            // holding one long gene streams three times the SNPs of a warp holding six short ones and the CTA waits for it.
            const bool two_level = kappa_warp >= 0.0 && kappa_warp != kappa;
            int64_t cta_a = r.a, cta_b = r.b;      // span of the CTA being cut (two-level)
            double cta_w = 0.0;
            auto started = [&](int64_t x) {
                return (int64_t)(std::lower_bound(goff.begin() + r.g0, goff.begin() + r.g1, (int32_t)x) - (goff.begin() + r.g0));
            };
            for (int k = 0; k < ng; ++k) {
                int64_t x;
                if (!two_level || k % groups_per_cta == 0) {
                    const double target = r.w * (double)k / (double)ng;
                    int64_t lo = r.a, hi = r.b;
                    while (lo < hi) {
                        const int64_t mid = lo + (hi - lo) / 2;
                        if (cost(mid) < target) lo = mid + 1; else hi = mid;
                    }
                    x = lo;
                } else {
                    const double target = cta_w * (double)(k % groups_per_cta) / (double)groups_per_cta;
                    const int64_t s0 = started(cta_a);
                    int64_t lo = cta_a, hi = cta_b;
                    while (lo < hi) {
                        const int64_t mid = lo + (hi - lo) / 2;
                        if ((double)(mid - cta_a) + kappa_warp * (double)(started(mid) - s0) < target) lo = mid + 1; else hi = mid;
                    }
                    x = lo;
                }
                if (k == 0) x = r.a;
                if (x < prev) x = prev;
                if (x > r.b) x = r.b;
                if (k > 0 && tol > 0) {
                    auto it = std::lower_bound(goff.begin(), goff.end(), (int32_t)x);
                    int64_t best = -1;
                    if (it != goff.end() && llabs((int64_t)*it - x) <= tol) best = *it;
                    if (it != goff.begin()) {
                        const int64_t lo = *(it - 1);
                        if (llabs(lo - x) <= tol && (best < 0 || llabs(lo - x) < llabs(best - x))) best = lo;
                    }
                    if (best >= prev && best <= r.b) x = best;
                }
                P.off[(size_t)(v0 + k)] = (int32_t)x;
                prev = x;
                if (two_level && k % groups_per_cta == 0) {
                    // the CTA's span ends at the (unsnapped) ideal cut of the next CTA: close enough for the warp targets
                    cta_a = x;
                    if (k + groups_per_cta >= ng) cta_b = r.b;
                    else {
                        const double target = r.w * (double)(k + groups_per_cta) / (double)ng;
                        int64_t lo = r.a, hi = r.b;
                        while (lo < hi) {
                            const int64_t mid = lo + (hi - lo) / 2;
                            if (cost(mid) < target) lo = mid + 1; else hi = mid;
                        }
                        cta_b = std::max(lo, cta_a);
                    }
                    cta_w = (double)(cta_b - cta_a) + kappa_warp * (double)(started(cta_b) - started(cta_a));
                }
            }
        }
        v0 += ng;
        cta0 += r.n;
        P.off[(size_t)v0] = (int32_t)r.b;
    }
    P.off[(size_t)n_groups] = (int32_t)M;
    if (align > 1) {
        // batch-aligned ranges: a cut inside a gene moves down to a multiple of `align` (gene boundaries of the
        // padded array are multiples already); monotonicity is preserved by rounding down
        for (int v = 1; v < n_groups; ++v) P.off[(size_t)v] = (P.off[(size_t)v] / align) * align;
        // the owner of an exclusive gene's prior terms is its first NON-EMPTY group: recompute (tiny inputs can
        // lose a group to the rounding)
        int32_t owned_gene = -1;
        for (int v = 0; v < n_groups; ++v) {
            if (!(P.flags[(size_t)v] & kExclusive)) continue;
            P.flags[(size_t)v] &= ~kOwner;
            if (P.off[(size_t)v] >= P.off[(size_t)v + 1]) continue;
            const int32_t g = gid[P.off[(size_t)v]];
            if (g != owned_gene && !(g == 0 && first_partial) && !(g < n_replicated && !owns_replicated)) P.flags[(size_t)v] |= kOwner;
            owned_gene = g;
        }
    }

    // flags + slot lists of the mixed groups, mirroring the kernel's flush rule
    std::vector<std::pair<int32_t, int32_t>> mixed_slots;   // (gene, slot)
    for (int v = 0; v < n_groups; ++v) {
        if (P.flags[(size_t)v] & kExclusive) continue;
        const int32_t a = P.off[(size_t)v], b = P.off[(size_t)v + 1];
        if (a >= b) continue;
        const int32_t gf = gid[a], gl = gid[b - 1];
        const bool head = (a > 0) ? (gid[a - 1] == gf) : first_partial;
        const bool tail = (b < M) ? (gid[b] == gl) : last_partial;
        P.flags[(size_t)v] = (head ? kHeadPartial : 0) | (tail ? kTailPartial : 0);
        if (head) mixed_slots.emplace_back(gf, v);
        if (tail && !(gl == gf && head)) mixed_slots.emplace_back(gl, n_groups + v);
    }
    // merge exclusive lists and mixed slots in gene order (a gene is either exclusive or mixed)
    {
        std::vector<std::vector<int32_t>> all_lists;
        std::vector<int32_t> all_gene;
        size_t ie = 0, im = 0;
        while (ie < lists.size() || im < mixed_slots.size()) {
            const int32_t ge = ie < lists.size() ? list_gene[ie] : INT32_MAX;
            const int32_t gm = im < mixed_slots.size() ? mixed_slots[im].first : INT32_MAX;
            if (ge <= gm) {
                all_gene.push_back(ge);
                all_lists.push_back(lists[ie]);
                ++ie;
            } else {
                all_gene.push_back(gm);
                all_lists.emplace_back();
                while (im < mixed_slots.size() && mixed_slots[im].first == gm) all_lists.back().push_back(mixed_slots[im++].second);
            }
        }
        lists.swap(all_lists);
        list_gene.swap(all_gene);
    }
    // long lists first (one CTA each in finalize), then the short ones (one warp each)
    std::vector<size_t> order;
    for (size_t i = 0; i < lists.size(); ++i)
        if ((int)lists[i].size() > kShortSplit) order.push_back(i);
    P.n_long = (int)order.size();
    for (size_t i = 0; i < lists.size(); ++i)
        if ((int)lists[i].size() <= kShortSplit) order.push_back(i);
    P.split_ptr.push_back(0);
    for (size_t i : order) {
        const int32_t g = list_gene[i];
        P.split_gene.push_back(g);
        int row = -1;
        if (g == 0 && first_partial) row = first_row;
        if (g == G - 1 && last_partial && last_row >= 0) row = last_row;
        if (g < n_replicated) row = g;
        P.split_row.push_back(row);
        for (int32_t sl : lists[i]) P.split_slots.push_back(sl);
        P.split_ptr.push_back((int32_t)P.split_slots.size());
    }
    return BGH_OK;
}

}  // namespace bgh

using namespace bgh;

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

struct bgh_ctx {
    bgh_cfg cfg;
    std::string err;
    int D = 0, Cp = 0, CL = 0, CPL = 0, chain_blocks = 0;
    int n_cta = 0, n_groups = 0, sm_count = 0;
    bool uploaded = false;
    double lik_const = 0.0;
    Plan plan;
    int64_t stream_len = 0;   // padded length of the device stream
    // device data
    double* d_xs = nullptr;   // prepared fp64 SoA {s/sqrt(l), 1/sqrt(l), sqrt(l)} (see bgh_lik_device.cuh)
    double* d_xw = nullptr;
    double* d_xl = nullptr;
    int* d_gid = nullptr;
    int4* d_groups = nullptr;
    double* d_slots = nullptr;
    double* d_part = nullptr;
    double* d_payload = nullptr;
    double* d_cc = nullptr;
    int* d_split_gene = nullptr;
    int* d_split_row = nullptr;
    int* d_split_ptr = nullptr;
    int* d_split_slots = nullptr;
    unsigned long long* d_trace = nullptr;   // tools only (bgh_debug_trace)
    // cooperative kernels (bgh_fused.cu): grid barrier words, status flag, tick counter
    unsigned long long* d_bar = nullptr;
    unsigned long long bar_count = 0;        // arrivals so far == value of *d_bar once queued launches finish
    int* d_status = nullptr;
    int* d_ticks = nullptr;
    int* h_poll = nullptr;                   // pinned: {n_finished, status, ticks}
    // fused persistent sampler (bgh_tick2.cu): command words, CTA partials, its own barrier word
    unsigned char* d_t2 = nullptr;
    size_t t2_bytes = 0;
    // the sampler kernel's own work plan (bgh_tick2.cu): a gene boundary costs a leapfrog of the gene's row in all chains on
    // top of the flush, and a gene split over SNP ranges is a "late row" handled by one CTA per chain, so its plan weighs
    // boundaries higher (BGH_NUTS_KAPPA) and snaps every cut to a gene boundary (BGH_NUTS_SNAP) unless the gene is huge
    Plan plan2;
    int4* d_groups2 = nullptr;
    int* d_split_gene2 = nullptr;
    int* d_split_row2 = nullptr;
    int* d_split_ptr2 = nullptr;
    int* d_split_slots2 = nullptr;
    std::vector<int32_t> own_ptr, own_gene;  // per CTA: the genes complete inside one of its SNP ranges (rows the CTA owns)
    int* d_own_ptr = nullptr;
    int* d_own_gene = nullptr;
    unsigned int t2_epoch = 0;               // LL tags used so far by the per-chain peer exchange of the sampler ticks
    unsigned int ll_eval_seq = 0;            // exchanges so far on the evaluation mailbox
    bool peer_local = false;                 // peers are contexts of this process (bgh_peer_connect_local): nothing to unmap
    bool dbg_send_only = false;              // tests: bgh_debug_peer_mode
    int* d_gcoord = nullptr;                 // optional global coordinate of every local row (bgh_set_row_coords)
    bool fused = true;                       // BGH_UNFUSED=1 -> separate likelihood / finalize launches
    // peer-memory exchange of sharded contexts (bgh_peer_export / bgh_peer_connect)
    unsigned char* d_mailbox = nullptr;      // [flags 2 x kMaxPeers x 128 B][mail 2 x kMaxPeers x n doubles]
    unsigned char* peer_base[kMaxPeers] = {nullptr};
    int peer_rank = -1, peer_world = 0;
    unsigned long long peer_epoch = 0;       // launches that used mailbox region A (evaluation payload)
    unsigned long long nuts_epoch = 0;       // launches that used mailbox region B (sampler sums)
    // host-entry staging (lazily allocated)
    static constexpr int kPipe = 4;          // groups of evaluations in flight: upload | compute | download overlap
    int group_max = 1;                       // most evaluations per group (one copy each way per group), set with the staging
    double* d_q_cd[kPipe] = {};
    double* d_q_dc[kPipe] = {};
    double* d_g_dc[kPipe] = {};
    double* d_g_cd[kPipe] = {};
    double* d_lp[kPipe] = {};
    cudaStream_t st_h2d[2] = {}, st_cmp = nullptr, st_d2h[2] = {}, st_graph = nullptr;   // copies alternate between two streams per direction
    cudaEvent_t ev_up[kPipe] = {}, ev_cmp[kPipe] = {}, ev_dn[kPipe] = {};
    // profiling
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;   // triples: before lik, after lik, after finalize
    size_t prof_n = 0;
    static constexpr size_t kProfCap = 2048;
    int64_t launches = 0;
};

static int fail(bgh_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg;
    g_last_error = msg;
    return code;
}

#define BGH_CUDA(ctx, expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            char _buf[512];                                                                    \
            snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                      \
            return fail(ctx, BGH_ECUDA, _buf);                                                 \
        }                                                                                      \
    } while (0)

// cost of one gene boundary in SNP-equivalents for the work plan (BGH_PLAN_KAPPA overrides; tuning only)
static double plan_kappa()
{
    const char* e = getenv("BGH_PLAN_KAPPA");
    if (e && *e) return atof(e);
    return 5.0;     // a gene boundary (flush + slope switch in front of a batch) costs about 5 SNPs' worth (measured)
}

static void pick_chain_geometry(int C, int* CL, int* CPL, int* Cp)
{
    if (C <= 4) { *CL = 4; *CPL = 1; *Cp = 4; }
    else if (C <= 8) { *CL = 8; *CPL = 1; *Cp = 8; }
    else if (C <= 16) { *CL = 16; *CPL = 1; *Cp = 16; }
    else if (C <= 32) { *CL = 32; *CPL = 1; *Cp = 32; }
    else { *CL = 32; *CPL = 2; *Cp = ((C + 63) / 64) * 64; }
}

// ---------------------------------------------------------------------------------------------
// bgh_upload on the device: the gene-sorted, padded, prepared fp64 stream is built from the raw fp32 columns.
// The host keeps the one integer pass the work plan needs anyway (gene histogram; dest[j] = padded start of the
// gene + number of earlier SNPs of the gene in file order, i.e. a STABLE counting sort); the scatter, the three
// fp64 arrays (IEEE sqrt / divide: the same bits the host formulas give), the input validation and sum log(l) run
// here.  ref: snps.py:117-134 hands (z^2, l, cat) to the model build, gene_regression.py:61-98.
// ---------------------------------------------------------------------------------------------
constexpr int kPrepThreads = 256;
__global__ void __launch_bounds__(kPrepThreads) prepare_stream_kernel(const float* __restrict__ chi2, const float* __restrict__ ld,
                                                                      const int* __restrict__ dest, long long M, double* __restrict__ sp,
                                                                      double* __restrict__ wp, double* __restrict__ lp,
                                                                      double* __restrict__ log_part, int* __restrict__ bad)
{
    __shared__ double s_red[kPrepThreads / 32];
    // block b owns the contiguous slice [b * per, (b + 1) * per): its sum of log(l) is a fixed-order sum (thread-strided
    // within the slice, xor tree, warps in order), so the total is reproducible for a given M
    const long long per = (M + gridDim.x - 1) / gridDim.x;
    const long long j0 = (long long)blockIdx.x * per, j1 = min(M, j0 + per);
    double acc = 0.0;
    for (long long j = j0 + threadIdx.x; j < j1; j += kPrepThreads) {
        const float lf = ld[j], sf = chi2[j];
        if (!(lf > 0.0f) || !isfinite(lf) || !isfinite(sf)) { *bad = 1; continue; }
        const double l = (double)lf, sq = sqrt(l);
        const int d = dest[j];
        sp[d] = (double)sf / sq;
        wp[d] = 1.0 / sq;
        lp[d] = sq;
        acc += log(l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kPrepThreads / 32; ++w) t += s_red[w];
        log_part[blockIdx.x] = t;
    }
}

extern "C" {

void bgh_cfg_default(bgh_cfg* cfg)
{
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->abi_version = BGH_ABI_VERSION;
    cfg->model = BGH_MODEL_GENE_NORMAL;
    cfg->n_chains = 4;
    cfg->first_gene_row = -1;
    cfg->last_gene_row = -1;
}

const char* bgh_version(void)
{
    return "baghera_b200 abi=1 arch=sm_100a build=" __DATE__ " " __TIME__;
}

const char* bgh_last_error(const bgh_ctx* ctx)
{
    return ctx ? ctx->err.c_str() : g_last_error.c_str();
}

int bgh_create(bgh_ctx** out, const bgh_cfg* cfg)
{
    if (!out || !cfg) return fail(nullptr, BGH_EINVAL, "bgh_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != BGH_ABI_VERSION) return fail(nullptr, BGH_EINVAL, "bgh_create: ABI version mismatch");
    if (cfg->n_snps <= 0 || cfg->n_snps > 0x7fffff00LL) return fail(nullptr, BGH_EINVAL, "bgh_create: n_snps out of range");
    if (cfg->n_chains <= 0 || cfg->n_chains > 4096) return fail(nullptr, BGH_EINVAL, "bgh_create: n_chains out of range");
    if (cfg->model != BGH_MODEL_GENE_NORMAL && cfg->model != BGH_MODEL_GW_NORMAL && cfg->model != BGH_MODEL_GENE_GAMMA)
        return fail(nullptr, BGH_EINVAL, "bgh_create: unknown model");
    if (cfg->model == BGH_MODEL_GENE_GAMMA && (!(cfg->gamma_rate > 0.0) || !isfinite(cfg->gamma_rate)))
        return fail(nullptr, BGH_EINVAL, "bgh_create: the gamma model needs gamma_rate = N_1kG > 0");
    if (cfg->model == BGH_MODEL_GENE_GAMMA && cfg->n_shared_rows != cfg->n_replicated)
        return fail(nullptr, BGH_EINVAL, "bgh_create: the gamma model shards with replicated genes only");
    if (cfg->n_genes <= 0 || (cfg->model == BGH_MODEL_GW_NORMAL && cfg->n_genes != 1))
        return fail(nullptr, BGH_EINVAL, "bgh_create: bad n_genes");
    if (!(cfg->c > 0.0) || !isfinite(cfg->c)) return fail(nullptr, BGH_EINVAL, "bgh_create: c must be positive");
    if (cfg->n_shared_rows < 0 || cfg->first_gene_row >= cfg->n_shared_rows || cfg->last_gene_row >= cfg->n_shared_rows)
        return fail(nullptr, BGH_EINVAL, "bgh_create: inconsistent shared-gene rows");
    if (cfg->n_replicated < 0 || cfg->n_replicated > cfg->n_genes || cfg->n_replicated > 64 ||
        (cfg->n_replicated > 0 && (cfg->n_shared_rows != cfg->n_replicated || cfg->first_gene_partial || cfg->last_gene_partial)))
        return fail(nullptr, BGH_EINVAL, "bgh_create: inconsistent replicated genes (n_shared_rows must equal n_replicated <= 64)");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        return fail(nullptr, BGH_ENODEV, "bgh_create: no usable CUDA device (this library has no CPU fallback)");
    }
    bgh_ctx* ctx = new (std::nothrow) bgh_ctx();
    if (!ctx) return fail(nullptr, BGH_ENOMEM, "bgh_create: out of host memory");
    ctx->cfg = *cfg;
    if (ctx->cfg.n_genes_total <= 0) ctx->cfg.n_genes_total = cfg->n_genes;
    ctx->D = (cfg->model == BGH_MODEL_GENE_NORMAL) ? cfg->n_genes + 3 : (cfg->model == BGH_MODEL_GENE_GAMMA ? cfg->n_genes + 2 : 2);
    pick_chain_geometry(cfg->n_chains, &ctx->CL, &ctx->CPL, &ctx->Cp);
    ctx->chain_blocks = ctx->Cp / (ctx->CL * ctx->CPL);

    cudaDeviceProp prop;
    if (cudaSetDevice(cfg->device) != cudaSuccess || cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) {
        std::string m = std::string("bgh_create: cudaSetDevice failed: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return fail(nullptr, BGH_ECUDA, m);
    }
    if (prop.major < 10) {
        delete ctx;
        return fail(nullptr, BGH_ENODEV, "bgh_create: device is not sm_100 class (kernels are built for sm_100a only)");
    }
    ctx->sm_count = cfg->sm_count_override > 0 ? cfg->sm_count_override : prop.multiProcessorCount;
    ctx->n_cta = ctx->sm_count;
    ctx->n_groups = ctx->n_cta * kLikWarps;     // one SNP range per warp, whatever the chain geometry
    e = lik_configure();
    if (e == cudaSuccess) e = fused_configure();
    if (e == cudaSuccess) e = tick2_configure();
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_bar, 256);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_bar, 0, 256);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_status, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_status, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_ticks, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_ticks, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_poll, 4 * sizeof(int));
    {
        const char* uf = getenv("BGH_UNFUSED");
        ctx->fused = !(uf && *uf == '1') && prop.cooperativeLaunch != 0 && ctx->n_cta <= prop.multiProcessorCount;
    }
    if (e != cudaSuccess) {
        std::string m = std::string("bgh_create: kernel configuration failed: ") + cudaGetErrorString(e);
        delete ctx;
        return fail(nullptr, BGH_ECUDA, m);
    }
    *out = ctx;
    return BGH_OK;
}

void bgh_destroy(bgh_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    cudaFree(ctx->d_xs); cudaFree(ctx->d_xw); cudaFree(ctx->d_xl); cudaFree(ctx->d_gid); cudaFree(ctx->d_groups);
    cudaFree(ctx->d_slots); cudaFree(ctx->d_part); cudaFree(ctx->d_payload); cudaFree(ctx->d_cc);
    cudaFree(ctx->d_trace);
    cudaFree(ctx->d_bar); cudaFree(ctx->d_status); cudaFree(ctx->d_ticks); cudaFree(ctx->d_t2);
    cudaFree(ctx->d_own_ptr); cudaFree(ctx->d_own_gene); cudaFree(ctx->d_gcoord);
    cudaFree(ctx->d_groups2); cudaFree(ctx->d_split_gene2); cudaFree(ctx->d_split_row2); cudaFree(ctx->d_split_ptr2); cudaFree(ctx->d_split_slots2);
    if (ctx->h_poll) cudaFreeHost(ctx->h_poll);
    for (int r = 0; r < kMaxPeers; ++r)
        if (ctx->peer_base[r] && r != ctx->peer_rank && !ctx->peer_local) cudaIpcCloseMemHandle(ctx->peer_base[r]);
    cudaFree(ctx->d_mailbox);
    cudaFree(ctx->d_split_gene); cudaFree(ctx->d_split_row); cudaFree(ctx->d_split_ptr); cudaFree(ctx->d_split_slots);
    for (int i = 0; i < bgh_ctx::kPipe; ++i) {
        cudaFree(ctx->d_q_cd[i]); cudaFree(ctx->d_q_dc[i]); cudaFree(ctx->d_g_dc[i]); cudaFree(ctx->d_g_cd[i]);
        cudaFree(ctx->d_lp[i]);
        if (ctx->ev_up[i]) cudaEventDestroy(ctx->ev_up[i]);
        if (ctx->ev_cmp[i]) cudaEventDestroy(ctx->ev_cmp[i]);
        if (ctx->ev_dn[i]) cudaEventDestroy(ctx->ev_dn[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (ctx->st_h2d[i]) cudaStreamDestroy(ctx->st_h2d[i]);
        if (ctx->st_d2h[i]) cudaStreamDestroy(ctx->st_d2h[i]);
    }
    if (ctx->st_cmp) cudaStreamDestroy(ctx->st_cmp);
    if (ctx->st_graph) cudaStreamDestroy(ctx->st_graph);
    for (cudaEvent_t ev : ctx->prof_ev) cudaEventDestroy(ev);
    delete ctx;
}

int32_t bgh_dim(const bgh_ctx* ctx) { return ctx ? ctx->D : 0; }
int32_t bgh_chain_stride(const bgh_ctx* ctx) { return ctx ? ctx->Cp : 0; }
int32_t bgh_payload_rows(const bgh_ctx* ctx) { return ctx ? 4 + ctx->cfg.n_shared_rows : 0; }
double bgh_lik_const(const bgh_ctx* ctx) { return ctx ? ctx->lik_const : 0.0; }
int bgh_set_lik_const(bgh_ctx* ctx, double total)
{
    if (!ctx || !isfinite(total)) return fail(ctx, BGH_EINVAL, "bgh_set_lik_const: bad argument");
    if (!ctx->uploaded) return fail(ctx, BGH_ESTATE, "bgh_set_lik_const: call bgh_upload first");
    ctx->lik_const = total;
    return BGH_OK;
}
int32_t bgh_n_groups(const bgh_ctx* ctx) { return ctx ? ctx->n_groups : 0; }
int32_t bgh_n_split_genes(const bgh_ctx* ctx) { return ctx ? (int32_t)ctx->plan.split_gene.size() : 0; }
int64_t bgh_launch_count(const bgh_ctx* ctx) { return ctx ? ctx->launches : 0; }

int bgh_get_partition(const bgh_ctx* ctx, int32_t* offsets, int32_t n)
{
    if (!ctx || !offsets) return BGH_EINVAL;
    if (!ctx->uploaded) return BGH_ESTATE;
    if (n != ctx->n_groups + 1) return BGH_EINVAL;
    memcpy(offsets, ctx->plan.off.data(), sizeof(int32_t) * (size_t)n);
    return BGH_OK;
}

int bgh_get_sampler_partition(const bgh_ctx* ctx, int32_t* offsets, int32_t n)
{
    if (!ctx || !offsets) return BGH_EINVAL;
    if (!ctx->uploaded) return BGH_ESTATE;
    if (n != ctx->n_groups + 1) return BGH_EINVAL;
    memcpy(offsets, ctx->plan2.off.data(), sizeof(int32_t) * (size_t)n);
    return BGH_OK;
}

// Host-only planning entry point used by the CPU tests (no device needed): sorted gene ids in,
// offsets[n_groups+1], flags[n_groups] out; split CSR returned through caller-sized buffers.
int bgh_debug_plan(int64_t M, const int32_t* gid_sorted, int32_t G, int32_t n_cta, int32_t groups_per_cta, int32_t first_partial,
                   int32_t last_partial, int32_t* offsets, int32_t* flags, int32_t* n_split,
                   int32_t* split_gene, int32_t* split_ptr, int32_t* split_slots, int32_t cap)
{
    Plan P;
    std::string err;
    int rc = build_plan(M, gid_sorted, G, n_cta, groups_per_cta, first_partial != 0, last_partial != 0,
                        first_partial ? 0 : -1, last_partial ? (first_partial && G == 1 ? 0 : (first_partial ? 1 : 0)) : -1,
                        plan_kappa(), &P, &err);
    if (rc != BGH_OK) return fail(nullptr, rc, err);
    memcpy(offsets, P.off.data(), sizeof(int32_t) * P.off.size());
    memcpy(flags, P.flags.data(), sizeof(int32_t) * P.flags.size());
    const int ns = (int)P.split_gene.size();
    if (ns > cap || (int)P.split_slots.size() > cap) return fail(nullptr, BGH_EINVAL, "bgh_debug_plan: cap too small");
    *n_split = ns;
    memcpy(split_gene, P.split_gene.data(), sizeof(int32_t) * (size_t)ns);
    memcpy(split_ptr, P.split_ptr.data(), sizeof(int32_t) * (size_t)(ns + 1));
    memcpy(split_slots, P.split_slots.data(), sizeof(int32_t) * P.split_slots.size());
    return BGH_OK;
}

// The sampler kernel's plan (defaults of bgh_upload: kappa 60, snap 0.5, kappa_warp 8, ranges aligned to the batch) on the host.
int bgh_debug_sampler_plan(int64_t M, const int32_t* gid_sorted, int32_t G, int32_t n_cta, int32_t groups_per_cta, int32_t two_level,
                           int32_t* offsets, int32_t* flags)
{
    Plan P;
    std::string err;
    int rc = build_plan(M, gid_sorted, G, n_cta, groups_per_cta, false, false, -1, -1, 60.0, &P, &err, kLikBatch, 0, true, 0.5,
                        two_level ? 8.0 : -1.0);
    if (rc != BGH_OK) return fail(nullptr, rc, err);
    memcpy(offsets, P.off.data(), sizeof(int32_t) * P.off.size());
    memcpy(flags, P.flags.data(), sizeof(int32_t) * P.flags.size());
    return BGH_OK;
}

// Tools only: enable (buf != NULL -> allocate) per-warp phase timestamps of the likelihood kernel and
// copy the last launch's trace [chain_blocks * n_cta * 16 warps][8] u64 to `out` (may be NULL).
int64_t bgh_stream_len(const bgh_ctx* ctx) { return (ctx && ctx->uploaded) ? ctx->stream_len : 0; }

int bgh_debug_stream(bgh_ctx* ctx, int32_t which, void* out, int64_t n)
{
    if (!ctx || !ctx->uploaded || !out || which < 0 || which > 3 || n < 0 || n > ctx->stream_len)
        return fail(ctx, BGH_EINVAL, "bgh_debug_stream: bad argument");
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const void* src = which == 0 ? (const void*)ctx->d_xs : which == 1 ? (const void*)ctx->d_xw : which == 2 ? (const void*)ctx->d_xl : (const void*)ctx->d_gid;
    BGH_CUDA(ctx, cudaMemcpy(out, src, (size_t)n * (which == 3 ? sizeof(int) : sizeof(double)), cudaMemcpyDeviceToHost));
    return BGH_OK;
}

int bgh_debug_trace(bgh_ctx* ctx, int32_t enable, uint64_t* out, int64_t n_words)
{
    if (!ctx) return BGH_EINVAL;
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t words = (size_t)ctx->chain_blocks * ctx->n_cta * kLikWarps * 8;
    if (enable && !ctx->d_trace) {
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_trace, words * sizeof(unsigned long long)));
        BGH_CUDA(ctx, cudaMemset(ctx->d_trace, 0, words * sizeof(unsigned long long)));
    }
    if (out) {
        if (!ctx->d_trace || (size_t)n_words != words) return fail(ctx, BGH_EINVAL, "bgh_debug_trace: bad buffer size");
        BGH_CUDA(ctx, cudaDeviceSynchronize());
        BGH_CUDA(ctx, cudaMemcpy(out, ctx->d_trace, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    if (!enable && ctx->d_trace) {
        cudaFree(ctx->d_trace);
        ctx->d_trace = nullptr;
    }
    return BGH_OK;
}

int bgh_upload(bgh_ctx* ctx, const float* chi2, const float* ld, const int32_t* gene_id)
{
    if (!ctx || !chi2 || !ld) return fail(ctx, BGH_EINVAL, "bgh_upload: null argument");
    const int64_t M = ctx->cfg.n_snps;
    const int32_t G = ctx->cfg.n_genes;
    if (!gene_id && ctx->cfg.model != BGH_MODEL_GW_NORMAL) return fail(ctx, BGH_EINVAL, "bgh_upload: gene_id required");
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));

    // stable counting sort by gene (ref: gene_regression.py:61-71 defines the id; Q6: file order is unsorted)
    std::vector<int64_t> cnt((size_t)G + 1, 0);
    if (gene_id) {
        for (int64_t j = 0; j < M; ++j) {
            const int32_t g = gene_id[j];
            if (g < 0 || g >= G) return fail(ctx, BGH_EINVAL, "bgh_upload: gene id out of range");
            cnt[(size_t)g + 1]++;
        }
    } else {
        cnt[1] = M;
    }
    for (int32_t g = 0; g < G; ++g) cnt[(size_t)g + 1] += cnt[g];
    // Every gene is padded to a multiple of the kernel's batch (kLikBatch SNPs) with zero-weight SNPs
    // (sp = wp = lp = 0 -> rho = 0: they add exactly nothing), so a batch never straddles a gene boundary
    // and the inner loop needs no per-SNP boundary handling.  Mp = padded SNP count.
    std::vector<int64_t> pcnt((size_t)G + 1, 0);
    for (int32_t g = 0; g < G; ++g) {
        const int64_t n = cnt[(size_t)g + 1] - cnt[g];
        pcnt[(size_t)g + 1] = pcnt[g] + ((n + kLikBatch - 1) / kLikBatch) * kLikBatch;
    }
    const int64_t Mp = pcnt[(size_t)G];
    if (Mp > 0x7fffff00LL) return fail(ctx, BGH_EINVAL, "bgh_upload: too many SNPs after padding");
    const int64_t Mpad = ((Mp + kChunk - 1) / kChunk) * kChunk + kChunk;
    std::vector<int32_t> g_sorted((size_t)Mpad, G - 1);
    for (int32_t g = 0; g < G; ++g)
        for (int64_t j = pcnt[g]; j < pcnt[(size_t)g + 1]; ++j) g_sorted[(size_t)j] = g;
    // stable destination of every SNP in the padded, gene-sorted stream (the scatter itself runs on the device)
    std::vector<int32_t> dest((size_t)std::max<int64_t>(M, 1));
    {
        std::vector<int64_t> pos(pcnt.begin(), pcnt.end() - 1);
        for (int64_t j = 0; j < M; ++j) dest[(size_t)j] = (int32_t)pos[(size_t)(gene_id ? gene_id[j] : 0)]++;
    }

    const bool gw = ctx->cfg.model == BGH_MODEL_GW_NORMAL;
    const bool fp = gw ? true : ctx->cfg.first_gene_partial != 0;
    const bool lp = gw ? true : ctx->cfg.last_gene_partial != 0;
    std::string err;
    const char* ekw = getenv("BGH_PLAN_KAPPA_WARP");       // tuning only: two-level cuts for the evaluation plan as well
    int rc = build_plan(Mp, g_sorted.data(), G, ctx->n_cta, kLikWarps, fp, lp, ctx->cfg.first_gene_row,
                        ctx->cfg.last_gene_row, plan_kappa(), &ctx->plan, &err, kLikBatch, ctx->cfg.n_replicated,
                        ctx->cfg.owns_replicated != 0, 0.0, (ekw && *ekw) ? atof(ekw) : -1.0);
    if (rc != BGH_OK) return fail(ctx, rc, err);

    {
        const char* ek = getenv("BGH_NUTS_KAPPA");
        const char* es = getenv("BGH_NUTS_SNAP");
        const char* ew = getenv("BGH_NUTS_KAPPA_WARP");
        rc = build_plan(Mp, g_sorted.data(), G, ctx->n_cta, kLikWarps, fp, lp, ctx->cfg.first_gene_row,
                        ctx->cfg.last_gene_row, (ek && *ek) ? atof(ek) : 60.0, &ctx->plan2, &err, kLikBatch, ctx->cfg.n_replicated,
                        ctx->cfg.owns_replicated != 0, (es && *es) ? atof(es) : 0.5, (ew && *ew) ? atof(ew) : 8.0);
        if (rc != BGH_OK) return fail(ctx, rc, err);
    }
    // rows each CTA owns in the fused sampler kernel (bgh_tick2.cu): genes complete inside one SNP range of the CTA
    {
        const Plan& P = ctx->plan2;
        const int gpc = kLikWarps;
        ctx->own_ptr.assign((size_t)ctx->n_cta + 1, 0);
        ctx->own_gene.clear();
        for (int cta = 0; cta < ctx->n_cta; ++cta) {
            for (int v = cta * gpc; v < (cta + 1) * gpc && !gw; ++v) {
                const int32_t a = P.off[(size_t)v], b = P.off[(size_t)v + 1];
                if (a >= b || (P.flags[(size_t)v] & kExclusive)) continue;
                const int32_t gf = g_sorted[(size_t)a], gl = g_sorted[(size_t)b - 1];
                for (int32_t g = gf; g <= gl; ++g) {
                    if (g == gf && (P.flags[(size_t)v] & kHeadPartial)) continue;
                    if (g == gl && (P.flags[(size_t)v] & kTailPartial)) continue;
                    ctx->own_gene.push_back(g);
                }
            }
            ctx->own_ptr[(size_t)cta + 1] = (int32_t)ctx->own_gene.size();
        }
    }

    auto dfree = [](auto*& p) { cudaFree(p); p = nullptr; };
    dfree(ctx->d_own_ptr); dfree(ctx->d_own_gene);
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_own_ptr, sizeof(int) * ctx->own_ptr.size()));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_own_gene, sizeof(int) * std::max<size_t>(ctx->own_gene.size(), 1)));
    BGH_CUDA(ctx, cudaMemcpy(ctx->d_own_ptr, ctx->own_ptr.data(), sizeof(int) * ctx->own_ptr.size(), cudaMemcpyHostToDevice));
    if (!ctx->own_gene.empty())
        BGH_CUDA(ctx, cudaMemcpy(ctx->d_own_gene, ctx->own_gene.data(), sizeof(int) * ctx->own_gene.size(), cudaMemcpyHostToDevice));
    dfree(ctx->d_xs); dfree(ctx->d_xw); dfree(ctx->d_xl); dfree(ctx->d_gid); dfree(ctx->d_groups); dfree(ctx->d_slots);
    dfree(ctx->d_part); dfree(ctx->d_payload); dfree(ctx->d_cc); dfree(ctx->d_split_gene); dfree(ctx->d_split_row);
    dfree(ctx->d_split_ptr); dfree(ctx->d_split_slots);

    const Plan& P = ctx->plan;
    std::vector<int4> groups((size_t)ctx->n_groups);
    for (int v = 0; v < ctx->n_groups; ++v)
        groups[(size_t)v] = make_int4(P.off[(size_t)v], P.off[(size_t)v + 1], P.flags[(size_t)v],
                                      P.off[(size_t)v] < P.off[(size_t)v + 1] ? g_sorted[(size_t)P.off[(size_t)v]] : 0);
    const size_t ns = P.split_gene.size();
    const size_t nsl = std::max<size_t>(P.split_slots.size(), 1);

    BGH_CUDA(ctx, cudaMalloc(&ctx->d_xs, sizeof(double) * (size_t)Mpad));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_xw, sizeof(double) * (size_t)Mpad));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_xl, sizeof(double) * (size_t)Mpad));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_gid, sizeof(int) * (size_t)Mpad));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_groups, sizeof(int4) * groups.size()));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_slots, sizeof(double) * (2 * (size_t)ctx->n_groups + ctx->n_cta) * ctx->Cp));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_part, sizeof(double) * 4 * (size_t)ctx->n_cta * ctx->Cp));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_cc, sizeof(double) * 2 * (size_t)ctx->Cp));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_payload, sizeof(double) * (size_t)(4 + ctx->cfg.n_shared_rows) * ctx->Cp));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_gene, sizeof(int) * std::max<size_t>(ns, 1)));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_row, sizeof(int) * std::max<size_t>(ns, 1)));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_ptr, sizeof(int) * (ns + 1)));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_slots, sizeof(int) * nsl));
    {
        // prepared fp64 SoA: rho = sp - e wp - (c beta) lp is the standardised residual (bgh_lik_device.cuh); padding
        // SNPs keep sp = wp = lp = 0.  Built by prepare_stream_kernel from the raw fp32 columns.
        float *d_chi2 = nullptr, *d_ld = nullptr;
        int *d_dest = nullptr, *d_bad = nullptr;
        double* d_logp = nullptr;
        const int n_blk = (int)std::min<int64_t>(1184, std::max<int64_t>(1, (M + kPrepThreads - 1) / kPrepThreads));
        auto release = [&]() { cudaFree(d_chi2); cudaFree(d_ld); cudaFree(d_dest); cudaFree(d_bad); cudaFree(d_logp); };
        cudaError_t e = cudaMalloc(&d_chi2, sizeof(float) * dest.size());
        if (e == cudaSuccess) e = cudaMalloc(&d_ld, sizeof(float) * dest.size());
        if (e == cudaSuccess) e = cudaMalloc(&d_dest, sizeof(int) * dest.size());
        if (e == cudaSuccess) e = cudaMalloc(&d_bad, sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&d_logp, sizeof(double) * n_blk);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_chi2, chi2, sizeof(float) * (size_t)M, cudaMemcpyHostToDevice, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ld, ld, sizeof(float) * (size_t)M, cudaMemcpyHostToDevice, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_dest, dest.data(), sizeof(int) * (size_t)M, cudaMemcpyHostToDevice, 0);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, sizeof(int), 0);
        if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_xs, 0, sizeof(double) * (size_t)Mpad, 0);
        if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_xw, 0, sizeof(double) * (size_t)Mpad, 0);
        if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_xl, 0, sizeof(double) * (size_t)Mpad, 0);
        std::vector<double> log_part((size_t)n_blk, 0.0);
        int bad = 0;
        if (e == cudaSuccess && M > 0) {
            prepare_stream_kernel<<<n_blk, kPrepThreads>>>(d_chi2, d_ld, d_dest, (long long)M, ctx->d_xs, ctx->d_xw, ctx->d_xl, d_logp, d_bad);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpy(log_part.data(), d_logp, sizeof(double) * n_blk, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
        }
        release();
        BGH_CUDA(ctx, e);
        if (bad) return fail(ctx, BGH_EINVAL, "bgh_upload: ld must be finite and > 0, chi2 finite (snps.py:121 gives l >= 1)");
        double acc = 0.0;
        for (int b = 0; b < n_blk; ++b) acc += log_part[(size_t)b];
        ctx->lik_const = -0.5 * acc - 0.5 * (double)M * kLog2Pi;
        ctx->stream_len = Mpad;
    }
    BGH_CUDA(ctx, cudaMemcpy(ctx->d_gid, g_sorted.data(), sizeof(int) * (size_t)Mpad, cudaMemcpyHostToDevice));
    BGH_CUDA(ctx, cudaMemcpy(ctx->d_groups, groups.data(), sizeof(int4) * groups.size(), cudaMemcpyHostToDevice));
    BGH_CUDA(ctx, cudaMemset(ctx->d_slots, 0, sizeof(double) * (2 * (size_t)ctx->n_groups + ctx->n_cta) * ctx->Cp));
    BGH_CUDA(ctx, cudaMemset(ctx->d_payload, 0, sizeof(double) * (size_t)(4 + ctx->cfg.n_shared_rows) * ctx->Cp));
    if (ns) {
        BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_gene, P.split_gene.data(), sizeof(int) * ns, cudaMemcpyHostToDevice));
        BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_row, P.split_row.data(), sizeof(int) * ns, cudaMemcpyHostToDevice));
        BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_slots, P.split_slots.data(), sizeof(int) * P.split_slots.size(), cudaMemcpyHostToDevice));
    }
    BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_ptr, P.split_ptr.data(), sizeof(int) * (ns + 1), cudaMemcpyHostToDevice));
    {
        const Plan& P2 = ctx->plan2;
        dfree(ctx->d_groups2); dfree(ctx->d_split_gene2); dfree(ctx->d_split_row2); dfree(ctx->d_split_ptr2); dfree(ctx->d_split_slots2);
        for (int v = 0; v < ctx->n_groups; ++v)
            groups[(size_t)v] = make_int4(P2.off[(size_t)v], P2.off[(size_t)v + 1], P2.flags[(size_t)v],
                                          P2.off[(size_t)v] < P2.off[(size_t)v + 1] ? g_sorted[(size_t)P2.off[(size_t)v]] : 0);
        const size_t ns2 = P2.split_gene.size();
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_groups2, sizeof(int4) * groups.size()));
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_gene2, sizeof(int) * std::max<size_t>(ns2, 1)));
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_row2, sizeof(int) * std::max<size_t>(ns2, 1)));
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_ptr2, sizeof(int) * (ns2 + 1)));
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_slots2, sizeof(int) * std::max<size_t>(P2.split_slots.size(), 1)));
        BGH_CUDA(ctx, cudaMemcpy(ctx->d_groups2, groups.data(), sizeof(int4) * groups.size(), cudaMemcpyHostToDevice));
        if (ns2) {
            BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_gene2, P2.split_gene.data(), sizeof(int) * ns2, cudaMemcpyHostToDevice));
            BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_row2, P2.split_row.data(), sizeof(int) * ns2, cudaMemcpyHostToDevice));
            BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_slots2, P2.split_slots.data(), sizeof(int) * P2.split_slots.size(), cudaMemcpyHostToDevice));
        }
        BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_ptr2, P2.split_ptr.data(), sizeof(int) * (ns2 + 1), cudaMemcpyHostToDevice));
    }
    ctx->uploaded = true;
    return BGH_OK;
}

static void fill_params(bgh_ctx* ctx, const double* q, double* grad, double* logp, double* payload,
                        LikParams* lp, FinParams* fp, bool nuts_plan = false)
{
    lp->sp = ctx->d_xs; lp->wp = ctx->d_xw; lp->lp = ctx->d_xl; lp->gid = ctx->d_gid; lp->groups = ctx->d_groups;
    lp->q = q; lp->grad = grad; lp->ds = ctx->Cp; lp->cs = 1; lp->woff = nullptr; fp->ds = ctx->Cp; fp->cs = 1; lp->slots = ctx->d_slots; lp->part = ctx->d_part; lp->cc = ctx->d_cc; lp->trace = ctx->d_trace;
    lp->Cp = ctx->Cp; lp->G = ctx->cfg.n_genes; lp->NG = ctx->n_groups; lp->n_cta = ctx->n_cta; lp->model = ctx->cfg.model;
    lp->fix = ctx->cfg.fix_intercept; lp->c = ctx->cfg.c; lp->gamma_rate = ctx->cfg.gamma_rate;
    fp->q = q; fp->grad = grad; fp->logp = logp; fp->slots = ctx->d_slots; fp->part = ctx->d_part; fp->cc = ctx->d_cc;
    fp->payload = payload;
    fp->split_gene = ctx->d_split_gene; fp->split_row = ctx->d_split_row; fp->split_ptr = ctx->d_split_ptr;
    fp->split_slots = ctx->d_split_slots; fp->n_split = (int)ctx->plan.split_gene.size(); fp->n_long = ctx->plan.n_long;
    fp->n_cta = ctx->n_cta; fp->Cp = ctx->Cp; fp->G_total = ctx->cfg.n_genes_total; fp->model = ctx->cfg.model;
    fp->fix = ctx->cfg.fix_intercept; fp->finish_inline = 0; fp->c = ctx->cfg.c; fp->lik_const = ctx->lik_const; fp->gamma_rate = ctx->cfg.gamma_rate;
    fp->first_gene_row = ctx->cfg.first_gene_row; fp->last_gene_row = ctx->cfg.last_gene_row;
    fp->G_local = ctx->cfg.n_genes;
    fp->n_shared_rows = ctx->cfg.n_shared_rows;
    fp->n_replicated = ctx->cfg.n_replicated;
    if (nuts_plan) {
        lp->groups = ctx->d_groups2;
        fp->split_gene = ctx->d_split_gene2; fp->split_row = ctx->d_split_row2; fp->split_ptr = ctx->d_split_ptr2;
        fp->split_slots = ctx->d_split_slots2; fp->n_split = (int)ctx->plan2.split_gene.size(); fp->n_long = ctx->plan2.n_long;
    }
}

static int run_sharded(bgh_ctx* ctx, const double* q, double* logp, double* grad, void* stream);

static int run_local(bgh_ctx* ctx, const double* q, double* logp, double* grad, double* payload,
                     bool finish_inline, cudaStream_t s)
{
    LikParams lp;
    FinParams fp;
    fill_params(ctx, q, grad, logp, payload, &lp, &fp);
    fp.finish_inline = finish_inline ? 1 : 0;
    cudaEvent_t* ev = nullptr;
    if (ctx->profiling && ctx->prof_n < bgh_ctx::kProfCap) {
        ev = &ctx->prof_ev[3 * ctx->prof_n];
        ctx->prof_n++;
        BGH_CUDA(ctx, cudaEventRecord(ev[0], s));
    }
    if (ctx->fused) {
        // one cooperative launch: likelihood pass | grid barrier | finalize (bgh_fused.cu)
        FusedParams f;
        f.lik = lp; f.fin = fp; f.bar = ctx->d_bar; f.status = ctx->d_status; f.chain_blocks = ctx->chain_blocks;
        f.bar_base = ctx->bar_count;
        ctx->bar_count += (unsigned long long)ctx->n_cta;   // one grid barrier per launch
        BGH_CUDA(ctx, launch_logp_grad_fused(f, ctx->CL, ctx->CPL, ctx->n_cta, s));
        if (ev) {
            BGH_CUDA(ctx, cudaEventRecord(ev[1], s));
            BGH_CUDA(ctx, cudaEventRecord(ev[2], s));
        }
        ctx->launches += 1;
        return BGH_OK;
    }
    BGH_CUDA(ctx, launch_lik(lp, ctx->CL, ctx->CPL, ctx->n_cta, ctx->chain_blocks, s));
    if (ev) BGH_CUDA(ctx, cudaEventRecord(ev[1], s));
    BGH_CUDA(ctx, launch_finalize(fp, s));
    if (ev) BGH_CUDA(ctx, cudaEventRecord(ev[2], s));
    ctx->launches += 2;
    return BGH_OK;
}

int bgh_logp_grad(bgh_ctx* ctx, const double* q, double* logp, double* grad, void* stream)
{
    if (!ctx || !q || !logp || !grad) return fail(ctx, BGH_EINVAL, "bgh_logp_grad: null argument");
    if (!ctx->uploaded) return fail(ctx, BGH_ESTATE, "bgh_logp_grad: call bgh_upload first");
    if (ctx->cfg.n_shared_rows != 0) return fail(ctx, BGH_ESTATE, "bgh_logp_grad: sharded context, use _local/_finish");
    return run_local(ctx, q, logp, grad, ctx->d_payload, true, (cudaStream_t)stream);
}

int bgh_logp_grad_local(bgh_ctx* ctx, const double* q, double* grad, double* payload, void* stream)
{
    if (!ctx || !q || !grad || !payload) return fail(ctx, BGH_EINVAL, "bgh_logp_grad_local: null argument");
    if (!ctx->uploaded) return fail(ctx, BGH_ESTATE, "bgh_logp_grad_local: call bgh_upload first");
    return run_local(ctx, q, nullptr, grad, payload, false, (cudaStream_t)stream);
}

int bgh_logp_grad_finish(bgh_ctx* ctx, const double* q, const double* payload, double* logp, double* grad,
                         void* stream)
{
    if (!ctx || !q || !payload || !logp || !grad) return fail(ctx, BGH_EINVAL, "bgh_logp_grad_finish: null argument");
    if (!ctx->uploaded) return fail(ctx, BGH_ESTATE, "bgh_logp_grad_finish: call bgh_upload first");
    LikParams lp;
    FinParams fp;
    fill_params(ctx, q, grad, logp, const_cast<double*>(payload), &lp, &fp);
    BGH_CUDA(ctx, launch_finish(fp, (cudaStream_t)stream));
    ctx->launches += 1;
    return BGH_OK;
}

int bgh_to_device_layout(bgh_ctx* ctx, const double* q_cd, double* q_dc, void* stream)
{
    if (!ctx || !q_cd || !q_dc) return fail(ctx, BGH_EINVAL, "bgh_to_device_layout: null argument");
    BGH_CUDA(ctx, launch_transpose_in(q_cd, q_dc, ctx->cfg.n_chains, ctx->Cp, ctx->D, (cudaStream_t)stream));
    ctx->launches += 1;
    return BGH_OK;
}

int bgh_from_device_layout(bgh_ctx* ctx, const double* x_dc, double* x_cd, void* stream)
{
    if (!ctx || !x_dc || !x_cd) return fail(ctx, BGH_EINVAL, "bgh_from_device_layout: null argument");
    BGH_CUDA(ctx, launch_transpose_out(x_dc, x_cd, ctx->cfg.n_chains, ctx->Cp, ctx->D, (cudaStream_t)stream));
    ctx->launches += 1;
    return BGH_OK;
}

static int ensure_host_staging(bgh_ctx* ctx)
{
    if (ctx->st_cmp) return BGH_OK;
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t n_cd = (size_t)ctx->cfg.n_chains * ctx->D, n_dc = (size_t)ctx->Cp * ctx->D;
    // a group moves about 16 MB each way: two evaluations at 64 chains x 15k genes, more when a rank's block
    // of q is small (sharded contexts, genome-wide model), so that a copy always outweighs the hand-off bubbles
    const double eval_bytes = (double)(sizeof(double) * n_cd);
    int gm = (int)(16.0e6 / eval_bytes + 0.5);
    ctx->group_max = gm < 1 ? 1 : (gm > 16 ? 16 : gm);
    for (int i = 0; i < bgh_ctx::kPipe; ++i) {
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_q_cd[i], sizeof(double) * n_cd * ctx->group_max));
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_g_cd[i], sizeof(double) * n_cd * ctx->group_max));
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_lp[i], sizeof(double) * (size_t)ctx->Cp * ctx->group_max));
        BGH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_up[i], cudaEventDisableTiming));
        BGH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_cmp[i], cudaEventDisableTiming));
        BGH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_dn[i], cudaEventDisableTiming));
    }
    // the device-layout buffers are touched by the compute stream only: one of each is enough
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_q_dc[0], sizeof(double) * n_dc));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_g_dc[0], sizeof(double) * n_dc));
    BGH_CUDA(ctx, cudaMemset(ctx->d_g_dc[0], 0, sizeof(double) * n_dc));
    for (int i = 0; i < 2; ++i) {
        BGH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->st_h2d[i], cudaStreamNonBlocking));
        BGH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->st_d2h[i], cudaStreamNonBlocking));
    }
    BGH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->st_cmp, cudaStreamNonBlocking));
    return BGH_OK;
}

// Host entry (B1' with host buffers).  Stages: upload | compute | download.  A stream-ordered event wait in
// front of a copy leaves a bubble of tens of microseconds on that stream's timeline (tools/event_probe.py:
// 249 vs 161 us per upload+download pair), so (1) each copy direction ALTERNATES between two streams - one
// stream's bubble hides under the other's copy - and (2) evaluations can travel in GROUPS of g (one H2D copy,
// g x {transpose, evaluation, transpose}, one D2H copy per group).  kPipe groups are in flight.  Tools: BGH_E2E_GROUP overrides g, BGH_E2E_TRACE=1 prints the stage timeline of every group
// (2: host enqueue time only), BGH_E2E_SKIP masks stages (bit 0 evaluation, 1 upload, 2 download, 3 transposes).
int bgh_logp_grad_host_stream(bgh_ctx* ctx, int32_t n_evals, const double* q, double* logp, double* grad)
{
    if (!ctx || !q || !logp || !grad || n_evals <= 0) return fail(ctx, BGH_EINVAL, "bgh_logp_grad_host_stream: bad argument");
    if (!ctx->uploaded) return fail(ctx, BGH_ESTATE, "bgh_logp_grad_host_stream: call bgh_upload first");
    // sharded contexts (peer mailboxes connected): q / grad are this rank's rows [C][3 + local genes]; every
    // rank must make the same sequence of calls (each evaluation exchanges its payload with the peers)
    const bool sharded = ctx->peer_world > 1;
    if (!sharded && ctx->cfg.n_shared_rows != 0)
        return fail(ctx, BGH_ESTATE, "bgh_logp_grad_host_stream: sharded context, call bgh_peer_connect first");
    int rc = ensure_host_staging(ctx);
    if (rc != BGH_OK) return rc;
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const int C = ctx->cfg.n_chains, D = ctx->D, Cp = ctx->Cp;
    const size_t n_cd = (size_t)C * D;
    // group size: enough groups to keep the three stages overlapped, at most group_max evaluations per group
    int g = n_evals / (2 * bgh_ctx::kPipe);
    if (const char* e = getenv("BGH_E2E_GROUP")) g = atoi(e);
    g = g < 1 ? 1 : (g > ctx->group_max ? ctx->group_max : g);
    const int n_groups = (n_evals + g - 1) / g;
    const char* tr_env = getenv("BGH_E2E_TRACE");
    const bool trace = tr_env && *tr_env == '1' && n_groups <= 256;
    const char* sk_env = getenv("BGH_E2E_SKIP");
    const int skip = sk_env ? atoi(sk_env) : 0;
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize((size_t)n_groups * 6);
        for (auto& e : tev) BGH_CUDA(ctx, cudaEventCreate(&e));
    }
    double host_us = 0.0;
    for (int k = 0; k < n_groups; ++k) {
        const auto h0 = std::chrono::steady_clock::now();
        const int b = k % bgh_ctx::kPipe;
        cudaStream_t up = ctx->st_h2d[k & 1], dn = ctx->st_d2h[k & 1];
        const int i0 = k * g, ng = (n_evals - i0 < g) ? n_evals - i0 : g;
        // slot reuse: the upload of group k may overwrite q_cd[b] only after the compute of group k-kPipe has
        // consumed it, compute may overwrite g_cd[b] / lp[b] only after the download of group k-kPipe finished
        if (k >= bgh_ctx::kPipe) {
            BGH_CUDA(ctx, cudaStreamWaitEvent(up, ctx->ev_cmp[b], 0));
            BGH_CUDA(ctx, cudaStreamWaitEvent(ctx->st_cmp, ctx->ev_dn[b], 0));
        }
        if (trace) BGH_CUDA(ctx, cudaEventRecord(tev[6 * k + 0], up));
        if (!(skip & 2))
            BGH_CUDA(ctx, cudaMemcpyAsync(ctx->d_q_cd[b], q + (size_t)i0 * n_cd, sizeof(double) * n_cd * ng,
                                          cudaMemcpyHostToDevice, up));
        if (trace) BGH_CUDA(ctx, cudaEventRecord(tev[6 * k + 1], up));
        BGH_CUDA(ctx, cudaEventRecord(ctx->ev_up[b], up));
        BGH_CUDA(ctx, cudaStreamWaitEvent(ctx->st_cmp, ctx->ev_up[b], 0));
        if (trace) BGH_CUDA(ctx, cudaEventRecord(tev[6 * k + 2], ctx->st_cmp));
        for (int j = 0; j < ng; ++j) {
            if (!(skip & 8))
                BGH_CUDA(ctx, launch_transpose_in(ctx->d_q_cd[b] + (size_t)j * n_cd, ctx->d_q_dc[0], C, Cp, D, ctx->st_cmp));
            if (!(skip & 1)) {
                rc = sharded ? run_sharded(ctx, ctx->d_q_dc[0], ctx->d_lp[b] + (size_t)j * Cp, ctx->d_g_dc[0], ctx->st_cmp)
                             : run_local(ctx, ctx->d_q_dc[0], ctx->d_lp[b] + (size_t)j * Cp, ctx->d_g_dc[0], ctx->d_payload,
                                         true, ctx->st_cmp);
                if (rc != BGH_OK) return rc;
            }
            if (!(skip & 8))
                BGH_CUDA(ctx, launch_transpose_out(ctx->d_g_dc[0], ctx->d_g_cd[b] + (size_t)j * n_cd, C, Cp, D, ctx->st_cmp));
            ctx->launches += 2;
        }
        if (trace) BGH_CUDA(ctx, cudaEventRecord(tev[6 * k + 3], ctx->st_cmp));
        BGH_CUDA(ctx, cudaEventRecord(ctx->ev_cmp[b], ctx->st_cmp));
        BGH_CUDA(ctx, cudaStreamWaitEvent(dn, ctx->ev_cmp[b], 0));
        if (trace) BGH_CUDA(ctx, cudaEventRecord(tev[6 * k + 4], dn));
        if (!(skip & 4))
            BGH_CUDA(ctx, cudaMemcpyAsync(grad + (size_t)i0 * n_cd, ctx->d_g_cd[b], sizeof(double) * n_cd * ng,
                                          cudaMemcpyDeviceToHost, dn));
        BGH_CUDA(ctx, cudaMemcpy2DAsync(logp + (size_t)i0 * C, sizeof(double) * C, ctx->d_lp[b], sizeof(double) * Cp,
                                        sizeof(double) * C, (size_t)ng, cudaMemcpyDeviceToHost, dn));
        if (trace) BGH_CUDA(ctx, cudaEventRecord(tev[6 * k + 5], dn));
        BGH_CUDA(ctx, cudaEventRecord(ctx->ev_dn[b], dn));
        host_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h0).count();
    }
    for (int i = 0; i < 2; ++i) BGH_CUDA(ctx, cudaStreamSynchronize(ctx->st_d2h[i]));
    BGH_CUDA(ctx, cudaStreamSynchronize(ctx->st_cmp));
    for (int i = 0; i < 2; ++i) BGH_CUDA(ctx, cudaStreamSynchronize(ctx->st_h2d[i]));
    if (tr_env && (*tr_env == '1' || *tr_env == '2'))
        fprintf(stderr, "# %d evaluations in %d groups of %d; host enqueue time per evaluation %.1f us\n", n_evals, n_groups, g,
                host_us / n_evals);
    if (trace) {
        fprintf(stderr, "# group  h2d[start end]  compute[start end]  d2h[start end]   (us since the first upload)\n");
        for (int k = 0; k < n_groups; ++k) {
            float t[6];
            for (int j = 0; j < 6; ++j) cudaEventElapsedTime(&t[j], tev[0], tev[6 * k + j]);
            fprintf(stderr, "%4d  %8.1f %8.1f  %8.1f %8.1f  %8.1f %8.1f\n", k, t[0] * 1e3f, t[1] * 1e3f, t[2] * 1e3f,
                    t[3] * 1e3f, t[4] * 1e3f, t[5] * 1e3f);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    return BGH_OK;
}

int bgh_logp_grad_host(bgh_ctx* ctx, const double* q, double* logp, double* grad)
{
    return bgh_logp_grad_host_stream(ctx, 1, q, logp, grad);
}

int bgh_set_profiling(bgh_ctx* ctx, int32_t enable)
{
    if (!ctx) return BGH_EINVAL;
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    if (enable && ctx->prof_ev.empty()) {
        ctx->prof_ev.resize(3 * bgh_ctx::kProfCap, nullptr);
        for (auto& ev : ctx->prof_ev) BGH_CUDA(ctx, cudaEventCreate(&ev));
    }
    ctx->profiling = enable != 0;
    ctx->prof_n = 0;
    return BGH_OK;
}

int bgh_kernel_time_ms(bgh_ctx* ctx, double* lik_ms, double* finalize_ms, int64_t* n, int32_t reset)
{
    if (!ctx || !lik_ms || !finalize_ms || !n) return fail(ctx, BGH_EINVAL, "bgh_kernel_time_ms: null argument");
    double a = 0.0, b = 0.0;
    for (size_t i = 0; i < ctx->prof_n; ++i) {
        cudaEvent_t* ev = &ctx->prof_ev[3 * i];
        BGH_CUDA(ctx, cudaEventSynchronize(ev[2]));
        float t0 = 0.f, t1 = 0.f;
        BGH_CUDA(ctx, cudaEventElapsedTime(&t0, ev[0], ev[1]));
        BGH_CUDA(ctx, cudaEventElapsedTime(&t1, ev[1], ev[2]));
        a += t0;
        b += t1;
    }
    *n = (int64_t)ctx->prof_n;
    *lik_ms = ctx->prof_n ? a / (double)ctx->prof_n : 0.0;
    *finalize_ms = ctx->prof_n ? b / (double)ctx->prof_n : 0.0;
    if (reset) ctx->prof_n = 0;
    return BGH_OK;
}

int bgh_dfma_peak(bgh_ctx* ctx, double* ginstr_per_s)
{
    if (!ctx || !ginstr_per_s) return fail(ctx, BGH_EINVAL, "bgh_dfma_peak: null argument");
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    double* d_out = nullptr;
    BGH_CUDA(ctx, cudaMalloc(&d_out, sizeof(double)));
    cudaEvent_t e0, e1;
    BGH_CUDA(ctx, cudaEventCreate(&e0));
    BGH_CUDA(ctx, cudaEventCreate(&e1));
    const int iters = 1 << 15, n_cta = ctx->sm_count * 8;
    BGH_CUDA(ctx, launch_dfma_probe(d_out, iters / 8, n_cta, nullptr));   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        BGH_CUDA(ctx, cudaEventRecord(e0, nullptr));
        BGH_CUDA(ctx, launch_dfma_probe(d_out, iters, n_cta, nullptr));
        BGH_CUDA(ctx, cudaEventRecord(e1, nullptr));
        BGH_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        BGH_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double lanes = (double)n_cta * 256.0 * 8.0 * (double)iters;
        best = std::max(best, lanes / (ms * 1e-3) * 1e-9);
    }
    ctx->launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *ginstr_per_s = best;
    return BGH_OK;
}

// ---------------------------------------------------------------------------------------------
// sharded evaluation with the payload all-reduce fused into the kernel (peer-mapped mailboxes)
// ---------------------------------------------------------------------------------------------
// Mailbox of one rank (one allocation, one IPC handle): region A = evaluation payload, region B = the sharded
// sampler's per-chain sums; each region is [flags 2 x kMaxPeers x 128 B][mail 2 x kMaxPeers x n doubles].
static constexpr size_t kFlagBytes = 2 * (size_t)kMaxPeers * 128;
static size_t mailbox_a_bytes(const bgh_ctx* ctx)
{
    const size_t b = kFlagBytes + 2 * (size_t)kMaxPeers * (size_t)(4 + ctx->cfg.n_shared_rows) * ctx->Cp * sizeof(double);
    return (b + 127) / 128 * 128;
}
static size_t mailbox_b_bytes(const bgh_ctx* ctx)
{
    const size_t b = kFlagBytes + 2 * (size_t)kMaxPeers * (size_t)kNutsMaxPartials * ctx->Cp * sizeof(double);
    return (b + 127) / 128 * 128;
}
// LL-protocol regions of the per-chain exchange (bgh_tick2.cu): C = evaluation (4 + replicated genes values per chain),
// D = sampler tick (kT2Nq + replicated genes values per chain); cells of 16 bytes, [2 parities][kMaxPeers][Cp][n_pay]
static int ll_eval_pay(const bgh_ctx* ctx) { return kT2Lik + ctx->cfg.n_replicated; }
static int ll_tick_pay(const bgh_ctx* ctx) { return kT2Nq + ctx->cfg.n_replicated; }
static size_t mailbox_c_bytes(const bgh_ctx* ctx) { return 2 * (size_t)kMaxPeers * ctx->Cp * ll_eval_pay(ctx) * 16; }
static size_t mailbox_d_bytes(const bgh_ctx* ctx) { return 2 * (size_t)kMaxPeers * ctx->Cp * ll_tick_pay(ctx) * 16; }
static size_t mailbox_bytes(const bgh_ctx* ctx)
{
    return mailbox_a_bytes(ctx) + mailbox_b_bytes(ctx) + mailbox_c_bytes(ctx) + mailbox_d_bytes(ctx);
}
static void fill_ll_peer(const bgh_ctx* ctx, bool tick, T2Peer* pr)
{
    memset(pr, 0, sizeof(*pr));
    const size_t off = mailbox_a_bytes(ctx) + mailbox_b_bytes(ctx) + (tick ? mailbox_c_bytes(ctx) : 0);
    for (int r = 0; r < ctx->peer_world; ++r) pr->box[r] = reinterpret_cast<uint4*>(ctx->peer_base[r] + off);
    pr->rank = ctx->peer_rank;
    pr->world = ctx->peer_world < 1 ? 1 : ctx->peer_world;
    pr->n_pay = tick ? ll_tick_pay(ctx) : ll_eval_pay(ctx);
}

int32_t bgh_peer_handle_bytes(void) { return (int32_t)sizeof(cudaIpcMemHandle_t); }

int bgh_peer_export(bgh_ctx* ctx, void* handle_out)
{
    if (!ctx || !handle_out) return fail(ctx, BGH_EINVAL, "bgh_peer_export: null argument");
    if (!ctx->uploaded) return fail(ctx, BGH_ESTATE, "bgh_peer_export: call bgh_upload first");
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    if (!ctx->d_mailbox) {
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_mailbox, mailbox_bytes(ctx)));
        BGH_CUDA(ctx, cudaMemset(ctx->d_mailbox, 0, mailbox_bytes(ctx)));
        BGH_CUDA(ctx, cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    BGH_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->d_mailbox));
    memcpy(handle_out, &h, sizeof(h));
    return BGH_OK;
}

int bgh_peer_connect(bgh_ctx* ctx, int32_t rank, int32_t world, const void* handles)
{
    if (!ctx || !handles || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(ctx, BGH_EINVAL, "bgh_peer_connect: bad argument (at most 8 ranks)");
    if (!ctx->d_mailbox) return fail(ctx, BGH_ESTATE, "bgh_peer_connect: call bgh_peer_export first");
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            ctx->peer_base[r] = ctx->d_mailbox;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char*)handles + (size_t)r * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        BGH_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_base[r] = (unsigned char*)ptr;
    }
    ctx->peer_rank = rank;
    ctx->peer_world = world;
    ctx->peer_epoch = 0;
    ctx->nuts_epoch = 0;
    ctx->t2_epoch = 0;
    ctx->ll_eval_seq = 0;
    return BGH_OK;
}

int bgh_peer_connect_local(bgh_ctx* ctx, int32_t rank, int32_t world, bgh_ctx* const* peers)
{
    if (!ctx || !peers || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(ctx, BGH_EINVAL, "bgh_peer_connect_local: bad argument (at most 8 ranks)");
    for (int r = 0; r < world; ++r) {
        if (!peers[r] || !peers[r]->d_mailbox) return fail(ctx, BGH_ESTATE, "bgh_peer_connect_local: call bgh_peer_export on every rank first");
        if (peers[r]->Cp != ctx->Cp || peers[r]->cfg.n_replicated != ctx->cfg.n_replicated || peers[r]->cfg.n_shared_rows != ctx->cfg.n_shared_rows)
            return fail(ctx, BGH_EINVAL, "bgh_peer_connect_local: ranks disagree on the payload geometry");
        ctx->peer_base[r] = peers[r]->d_mailbox;
    }
    if (peers[rank] != ctx) return fail(ctx, BGH_EINVAL, "bgh_peer_connect_local: peers[rank] must be this context");
    ctx->peer_rank = rank;
    ctx->peer_world = world;
    ctx->peer_local = true;
    ctx->peer_epoch = 0;
    ctx->nuts_epoch = 0;
    ctx->t2_epoch = 0;
    ctx->ll_eval_seq = 0;
    return BGH_OK;
}

int bgh_debug_peer_mode(bgh_ctx* ctx, int32_t send_only, int32_t seq)
{
    if (!ctx) return BGH_EINVAL;
    ctx->dbg_send_only = send_only != 0;
    if (seq >= 0) ctx->ll_eval_seq = (unsigned int)seq;
    return BGH_OK;
}

int bgh_set_row_coords(bgh_ctx* ctx, const int32_t* global_coord)
{
    if (!ctx || !global_coord) return fail(ctx, BGH_EINVAL, "bgh_set_row_coords: null argument");
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    if (!ctx->d_gcoord) BGH_CUDA(ctx, cudaMalloc(&ctx->d_gcoord, sizeof(int) * (size_t)ctx->D));
    BGH_CUDA(ctx, cudaMemcpy(ctx->d_gcoord, global_coord, sizeof(int) * (size_t)ctx->D, cudaMemcpyHostToDevice));
    return BGH_OK;
}

static int run_sharded(bgh_ctx* ctx, const double* q, double* logp, double* grad, void* stream)
{
    if (ctx->peer_world < 1) return fail(ctx, BGH_ESTATE, "bgh_logp_grad_sharded: call bgh_peer_connect first");
    if (!ctx->fused) return fail(ctx, BGH_ESTATE, "bgh_logp_grad_sharded: needs the cooperative kernels");
    LikParams lp;
    FinParams fp;
    fill_params(ctx, q, grad, logp, ctx->d_payload, &lp, &fp);
    fp.finish_inline = 0;
    if (ctx->cfg.n_replicated > 0 || ctx->cfg.n_shared_rows == 0) {
        // replicated-gene sharding: one barrier, per-chain LL exchange issued by the chain CTAs (bgh_tick2.cu)
        Eval2Params e;
        e.lik = lp; e.fin = fp; e.bar = ctx->d_bar; e.status = ctx->d_status; e.chain_blocks = ctx->chain_blocks;
        e.C = ctx->cfg.n_chains;
        e.bar_base = ctx->bar_count;
        ctx->bar_count += (unsigned long long)ctx->n_cta;
        fill_ll_peer(ctx, false, &e.peer);
        e.peer.epoch_base = 0;
        e.peer.send_only = ctx->dbg_send_only ? 1 : 0;
        e.epoch = ctx->ll_eval_seq++;
        cudaStream_t s2 = (cudaStream_t)stream;
        cudaEvent_t* ev2 = nullptr;
        if (ctx->profiling && ctx->prof_n < bgh_ctx::kProfCap) {
            ev2 = &ctx->prof_ev[3 * ctx->prof_n];
            ctx->prof_n++;
            BGH_CUDA(ctx, cudaEventRecord(ev2[0], s2));
        }
        BGH_CUDA(ctx, launch_eval2(e, ctx->CL, ctx->CPL, ctx->n_cta, s2));
        if (ev2) {
            BGH_CUDA(ctx, cudaEventRecord(ev2[1], s2));
            BGH_CUDA(ctx, cudaEventRecord(ev2[2], s2));
        }
        ctx->launches += 1;
        return BGH_OK;
    }
    FusedParams f;
    f.lik = lp; f.fin = fp; f.bar = ctx->d_bar; f.status = ctx->d_status; f.chain_blocks = ctx->chain_blocks;
    f.bar_base = ctx->bar_count;
    ctx->bar_count += 2ull * (unsigned long long)ctx->n_cta;   // two grid barriers per launch
    PeerParams pr;
    memset(&pr, 0, sizeof(pr));
    for (int r = 0; r < ctx->peer_world; ++r) {
        pr.flag[r] = reinterpret_cast<unsigned long long*>(ctx->peer_base[r]);
        pr.mail[r] = reinterpret_cast<double*>(ctx->peer_base[r] + kFlagBytes);
    }
    pr.rank = ctx->peer_rank; pr.world = ctx->peer_world;
    pr.n = (4 + ctx->cfg.n_shared_rows) * ctx->Cp;
    pr.epoch = ++ctx->peer_epoch;
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t* ev = nullptr;
    if (ctx->profiling && ctx->prof_n < bgh_ctx::kProfCap) {
        ev = &ctx->prof_ev[3 * ctx->prof_n];
        ctx->prof_n++;
        BGH_CUDA(ctx, cudaEventRecord(ev[0], s));
    }
    BGH_CUDA(ctx, launch_logp_grad_sharded(f, pr, ctx->CL, ctx->CPL, ctx->n_cta, s));
    if (ev) {
        BGH_CUDA(ctx, cudaEventRecord(ev[1], s));
        BGH_CUDA(ctx, cudaEventRecord(ev[2], s));
    }
    ctx->launches += 1;
    return BGH_OK;
}

int bgh_logp_grad_sharded(bgh_ctx* ctx, const double* q, double* logp, double* grad, void* stream)
{
    if (!ctx || !q || !logp || !grad) return fail(ctx, BGH_EINVAL, "bgh_logp_grad_sharded: null argument");
    if (!ctx->uploaded) return fail(ctx, BGH_ESTATE, "bgh_logp_grad_sharded: call bgh_upload first");
    return run_sharded(ctx, q, logp, grad, stream);
}

int bgh_peer_status(bgh_ctx* ctx, int32_t* status)
{
    if (!ctx || !status) return BGH_EINVAL;
    BGH_CUDA(ctx, cudaMemcpy(status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    return BGH_OK;
}

// ---------------------------------------------------------------------------------------------
// batched NUTS: host-side transition loop
// ---------------------------------------------------------------------------------------------
void bgh_nuts_cfg_default(bgh_nuts_cfg* cfg)
{
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->seed = 20211;
    cfg->target_accept = 0.90;
    cfg->step_scale = 0.25;
    cfg->emax = 1000.0;
    cfg->da_gamma = 0.05;
    cfg->da_kappa = 0.75;
    cfg->da_t0 = 10.0;
}

int32_t bgh_nuts_vec_blocks(const bgh_ctx* ctx) { return ctx ? 2 * ctx->sm_count : 0; }

static int nuts_params(bgh_ctx* ctx, const bgh_nuts_buffers* buf, const bgh_nuts_cfg* cfg, int64_t draw, NutsParams* p)
{
    if (!ctx || !buf || !cfg) return fail(ctx, BGH_EINVAL, "bgh_nuts: null argument");
    if (!ctx->uploaded) return fail(ctx, BGH_ESTATE, "bgh_nuts: call bgh_upload first");
    const bool sharded = ctx->cfg.n_genes_total != ctx->cfg.n_genes || ctx->cfg.n_shared_rows != 0 || ctx->peer_world > 1;
    if (sharded && ctx->peer_world < 2)
        return fail(ctx, BGH_ESTATE, "bgh_nuts: sharded context, call bgh_peer_connect first");
    if (!buf->state || !buf->ckpt || !buf->sc || !buf->fl || !buf->partials || !buf->sums || !buf->logp_w ||
        !buf->n_active_dev || !buf->n_active_host)
        return fail(ctx, BGH_EINVAL, "bgh_nuts: null buffer");
    p->state = buf->state; p->ckpt = buf->ckpt; p->sc = buf->sc; p->fl = buf->fl; p->partials = buf->partials;
    p->sums = buf->sums; p->logp_w = buf->logp_w; p->n_active = buf->n_active_dev;
    p->D = ctx->D; p->Cp = ctx->Cp; p->C = ctx->cfg.n_chains; p->n_vec_blocks = 2 * ctx->sm_count;
    p->coord_offset = cfg->coord_offset; p->seed = cfg->seed; p->draw = draw; p->chain_offset = cfg->chain_offset;
    p->head_rows = 0; p->own_head = 1; p->own_first = 1; p->status = ctx->d_status; p->gcoord = ctx->d_gcoord;
    memset(&p->peer, 0, sizeof(p->peer));
    if (sharded) {
        // rows in front of the per-gene block (the whole vector in the genome-wide model) are replicated
        p->head_rows = ctx->cfg.model == BGH_MODEL_GW_NORMAL ? ctx->D : (ctx->cfg.model == BGH_MODEL_GENE_GAMMA ? 2 : 3);
        if (ctx->cfg.model != BGH_MODEL_GW_NORMAL) p->head_rows += ctx->cfg.n_replicated;   // replicated genes are replicated rows
        p->own_head = ctx->peer_rank == 0 ? 1 : 0;
        p->own_first = ctx->cfg.first_gene_partial ? 0 : 1;
        for (int r = 0; r < ctx->peer_world; ++r) {
            unsigned char* const base = ctx->peer_base[r] + mailbox_a_bytes(ctx);
            p->peer.flag[r] = reinterpret_cast<unsigned long long*>(base);
            p->peer.mail[r] = reinterpret_cast<double*>(base + kFlagBytes);
        }
        p->peer.rank = ctx->peer_rank; p->peer.world = ctx->peer_world;
    }
    p->emax = cfg->emax; p->target_accept = cfg->target_accept; p->da_gamma = cfg->da_gamma;
    p->da_kappa = cfg->da_kappa; p->da_t0 = cfg->da_t0;
    return BGH_OK;
}

int bgh_nuts_init(bgh_ctx* ctx, const bgh_nuts_buffers* buf, const bgh_nuts_cfg* cfg, void* stream)
{
    NutsParams p;
    int rc = nuts_params(ctx, buf, cfg, 0, &p);
    if (rc != BGH_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t vec = (size_t)ctx->D * ctx->Cp;
    rc = p.peer.world > 1 ? run_sharded(ctx, buf->state + kSlotQP * vec, buf->logp_w, buf->state + kSlotGP * vec, s)
                          : run_local(ctx, buf->state + kSlotQP * vec, buf->logp_w, buf->state + kSlotGP * vec, ctx->d_payload, true, s);
    if (rc != BGH_OK) return rc;
    const int Dtot = cfg->n_dim_total > 0 ? cfg->n_dim_total : ctx->D;
    const double step0 = cfg->step_scale / pow((double)Dtot, 0.25);
    BGH_CUDA(ctx, nuts_launch_init(p, step0, s));
    ctx->launches += 1;
    return BGH_OK;
}

int bgh_nuts_stop_tuning(bgh_ctx* ctx, const bgh_nuts_buffers* buf, const bgh_nuts_cfg* cfg, void* stream)
{
    NutsParams p;
    int rc = nuts_params(ctx, buf, cfg, 0, &p);
    if (rc != BGH_OK) return rc;
    BGH_CUDA(ctx, nuts_launch_stop_tuning(p, (cudaStream_t)stream));
    ctx->launches += 1;
    return BGH_OK;
}

int bgh_nuts_transition(bgh_ctx* ctx, const bgh_nuts_buffers* buf, const bgh_nuts_cfg* cfg, int64_t draw,
                        int32_t tune, int32_t max_treedepth, double fg_weight, double bg_weight,
                        int32_t switch_window, void* stream, int32_t* n_leapfrogs)
{
    NutsParams p;
    int rc = nuts_params(ctx, buf, cfg, draw, &p);
    if (rc != BGH_OK) return rc;
    if (max_treedepth < 1 || max_treedepth > kNutsMaxDepth) return fail(ctx, BGH_EINVAL, "bgh_nuts_transition: bad max_treedepth");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t vec = (size_t)ctx->D * ctx->Cp;
    double* const qw = buf->state + kSlotQW * vec;
    double* const gw = buf->state + kSlotGW * vec;
    int leapfrogs = 0;

    // sharded: every scalar kernel that consumes per-chain sums all-reduces them over mailbox region B
    auto next_epoch = [&] { if (p.peer.world > 1) p.peer.epoch = ++ctx->nuts_epoch; };
    next_epoch();
    BGH_CUDA(ctx, nuts_launch_begin(p, s));
    ctx->launches += 2;
    for (int depth = 0; depth < max_treedepth; ++depth) {
        BGH_CUDA(ctx, nuts_launch_dir(p, s));
        ctx->launches += 1;
        const int n_leaves = 1 << depth;
        for (int leaf = 0; leaf < n_leaves; ++leaf) {
            BGH_CUDA(ctx, nuts_launch_leap1(p, leaf == 0 ? 1 : 0, s));
            rc = p.peer.world > 1 ? run_sharded(ctx, qw, buf->logp_w, gw, s)
                                  : run_local(ctx, qw, buf->logp_w, gw, ctx->d_payload, true, s);
            if (rc != BGH_OK) return rc;
            // checkpoint bookkeeping of the iterative tree (see bgh_sampler.cu)
            int idx_max = 0;
            for (int x = leaf >> 1; x > 0; x >>= 1) idx_max += x & 1;
            int store_level = -1, lvl_min = 0, lvl_max = -1;
            if ((leaf & 1) == 0) {
                store_level = idx_max;
            } else {
                int ones = 0;
                for (int x = leaf; x & 1; x >>= 1) ++ones;
                lvl_max = idx_max;
                lvl_min = idx_max - ones + 1;
            }
            next_epoch();
            BGH_CUDA(ctx, nuts_launch_leap2(p, leaf, store_level, lvl_min, lvl_max, s));
            ctx->launches += kNutsLaunchesPerLeaf;
            ++leapfrogs;
        }
        next_epoch();
        BGH_CUDA(ctx, nuts_launch_end_doubling(p, max_treedepth, s));
        ctx->launches += 3;
        BGH_CUDA(ctx, cudaMemcpyAsync(buf->n_active_host, buf->n_active_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        BGH_CUDA(ctx, cudaStreamSynchronize(s));
        if (*buf->n_active_host <= 0) break;
    }
    if (tune) {
        BGH_CUDA(ctx, nuts_launch_adapt_mass(p, fg_weight, bg_weight, switch_window, s));
        ctx->launches += 1;
    }
    BGH_CUDA(ctx, nuts_launch_finish(p, tune, buf->stats, s));
    ctx->launches += 1;
    if (n_leapfrogs) *n_leapfrogs = leapfrogs;
    return BGH_OK;
}

// The fused persistent sampler (bgh_tick2.cu): state stays gene-major [slot][D][Cp]; two grid barriers per tick.
static int nuts_run_tick2(bgh_ctx* ctx, const bgh_nuts_async_buffers* buf, const bgh_nuts_cfg* cfg, AsyncParams a, int poll_ticks,
                          cudaStream_t s, int64_t* n_ticks)
{
    const size_t vec = (size_t)ctx->D * ctx->Cp;
    const int Dtot = cfg->n_dim_total > 0 ? cfg->n_dim_total : ctx->D;
    const double step0 = cfg->step_scale / pow((double)Dtot, 0.25);
    if ((int)ctx->plan2.split_gene.size() > 8192)
        return fail(ctx, BGH_EINVAL, "bgh_nuts_run: more than 8192 genes are split over several SNP ranges");
    // scratch: [barrier 256 B][cmd_h Cp doubles][cmd_bits Cp ints][cmd_woff Cp ints][partials n_cta x kAsyncPartials x Cp doubles]
    const size_t off_h = 256, off_b = off_h + sizeof(double) * ctx->Cp, off_w = off_b + sizeof(int) * ctx->Cp;
    const size_t off_p = (off_w + sizeof(int) * ctx->Cp + 255) / 256 * 256;
    const size_t need = off_p + sizeof(double) * (size_t)ctx->n_cta * kAsyncPartials * ctx->Cp;
    if (ctx->t2_bytes < need) {
        cudaFree(ctx->d_t2);
        ctx->d_t2 = nullptr;
        ctx->t2_bytes = 0;
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_t2, need));
        ctx->t2_bytes = need;
    }
    BGH_CUDA(ctx, cudaMemsetAsync(ctx->d_t2, 0, need, s));
    BGH_CUDA(ctx, nuts_async_launch_init(a, step0, 10.0, s, false));
    ctx->launches += 1;
    // the two role-swapped buffers X0 = (QL, PL, GL), X1 = (QR, PR, GR): a finite start for every chain (padded chains
    // included); the first tick initialises both from Q / GRAD
    double* const q0 = buf->base.state + kSlotQL * vec;
    double* const g0 = buf->base.state + kSlotGL * vec;
    for (int which = 0; which < 2; ++which) {
        const size_t o = (size_t)which * (kSlotQR - kSlotQL) * vec;
        BGH_CUDA(ctx, cudaMemcpyAsync(q0 + o, buf->base.state + kSlotQP * vec, sizeof(double) * vec, cudaMemcpyDeviceToDevice, s));
        BGH_CUDA(ctx, cudaMemcpyAsync(g0 + o, buf->base.state + kSlotGP * vec, sizeof(double) * vec, cudaMemcpyDeviceToDevice, s));
        BGH_CUDA(ctx, cudaMemsetAsync(buf->base.state + kSlotPL * vec + o, 0, sizeof(double) * vec, s));
    }

    // reserve the persisting share of the L2 for the vectors every tick re-reads (evict-last hints in bgh_tick2.cu)
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->cfg.device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
            const char* e = getenv("BGH_L2_PERSIST_MB");
            size_t want = e ? (size_t)atoi(e) << 20 : 0;    // measured: reserving the persisting share makes the tick slower
            if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            if (getenv("BGH_T2_VERBOSE")) fprintf(stderr, "# persisting L2: max %d MB, reserved %zu MB, L2 %d MB\n",
                                                  prop.persistingL2CacheMaxSize >> 20, want >> 20, prop.l2CacheSize >> 20);
        }
        cudaGetLastError();
    }
    Tick2Params t;
    memset(&t, 0, sizeof(t));
    fill_params(ctx, q0, g0, buf->base.logp_w, ctx->d_payload, &t.lik, &t.fin, true);
    t.fin.finish_inline = 1;
    t.lik.trace = nullptr;
    t.a = a;
    t.a.n.n_vec_blocks = ctx->n_cta;
    t.bar = reinterpret_cast<unsigned long long*>(ctx->d_t2);
    t.status = ctx->d_status;
    t.chain_blocks = ctx->chain_blocks;
    t.n_ticks = poll_ticks;
    t.ticks_done = ctx->d_ticks;
    t.first_launch = 1;
    t.tick_base = 0;
    t.row0 = ctx->cfg.model == BGH_MODEL_GENE_NORMAL ? 3 : 2;
    t.n_unowned = 0;
    t.cmd_h = reinterpret_cast<double*>(ctx->d_t2 + off_h);
    t.cmd_bits = reinterpret_cast<int*>(ctx->d_t2 + off_b);
    t.cmd_woff = reinterpret_cast<int*>(ctx->d_t2 + off_w);
    t.lik.woff = t.cmd_woff;
    t.part = reinterpret_cast<double*>(ctx->d_t2 + off_p);
    t.cta_row_ptr = ctx->d_own_ptr;
    t.cta_row_gene = ctx->d_own_gene;
    { const char* e = getenv("BGH_T2_TRACE_CTA"); t.trace_cta = e ? atoi(e) : 0; }
    { const char* e = getenv("BGH_TRACE_TICK"); t.trace_first = e ? atoi(e) : 0; }
    t.k6 = buf->k6;
    t.head_trace = buf->head_trace;
    fill_ll_peer(ctx, true, &t.peer);
    t.peer.epoch_base = ctx->t2_epoch;
    t.phase_trace = ctx->d_trace;   // NULL unless bgh_debug_trace enabled it (tools/time_tick2.py)
    if (ctx->d_trace && getenv("BGH_T2_LIK_TRACE")) {   // tools: the per-warp stamps of the likelihood pass (last tick) instead
        t.lik.trace = ctx->d_trace;
        t.phase_trace = nullptr;
    }
    if (ctx->d_trace)               // the evaluation kernels share the buffer (per-warp stamps): start from zeros
        BGH_CUDA(ctx, cudaMemsetAsync(ctx->d_trace, 0, sizeof(unsigned long long) * (size_t)ctx->chain_blocks * ctx->n_cta * kLikWarps * 8, s));

    const int64_t max_ticks = (int64_t)(a.tune + a.draws) * 1024 + 1024;
    int64_t ticks = 0;
    int status = BGH_OK;
    BGH_CUDA(ctx, cudaMemsetAsync(ctx->d_ticks, 0, sizeof(int), s));
    while (ticks < max_ticks) {
        cudaError_t e = cudaMemsetAsync(t.bar, 0, 8, s);
        if (e == cudaSuccess) e = launch_tick2(t, ctx->CL, ctx->CPL, ctx->n_cta, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&ctx->h_poll[0], buf->base.n_active_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&ctx->h_poll[1], ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&ctx->h_poll[2], ctx->d_ticks, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { status = fail(ctx, BGH_ECUDA, std::string("bgh_nuts_run: ") + cudaGetErrorString(e)); break; }
        ctx->launches += 1;
        ticks = ctx->h_poll[2];
        t.tick_base = (int)ticks;
        t.first_launch = 0;
        if (ctx->h_poll[1] != 0) {
            cudaMemset(ctx->d_status, 0, sizeof(int));
            status = fail(ctx, BGH_ECUDA, ctx->h_poll[1] == 2 ? "bgh_nuts_run: a peer GPU did not answer inside the tick kernel"
                                                               : "bgh_nuts_run: grid barrier timed out inside the tick kernel");
            break;
        }
        *buf->base.n_active_host = ctx->h_poll[0];
        if (ctx->h_poll[0] >= ctx->cfg.n_chains) break;
    }
    ctx->t2_epoch += (unsigned int)ticks + 1u;
    cudaCtxResetPersistingL2Cache();
    cudaGetLastError();
    if (status == BGH_OK && ticks >= max_ticks) status = fail(ctx, BGH_ESTATE, "bgh_nuts_run: tick budget exhausted");
    if (n_ticks) *n_ticks = ticks;
    return status;
}

// Which persistent kernel runs bgh_nuts_run.  BGH_TICK=1 / 2 forces nuts_tick_kernel (bgh_fused.cu) / the fused kernel
// (bgh_tick2.cu).  Default: the fused kernel — sharded contexts and streaming statistics need it, and it wins wherever a
// rank's sampler state fits the L2 — except on a single rank whose state is far larger than the L2 (1.1M x 15k x 64 chains:
// both kernels stream the state from HBM there and the round-1 kernel's chain-major sweeps are ~15% faster).
static bool tick_v1(const bgh_ctx* ctx = nullptr, bool needs_v2 = false)
{
    const char* e = getenv("BGH_TICK");
    if (e && *e == '1') return true;
    if (e && *e == '2') return false;
    const char* o = getenv("BGH_TICK_V1");
    if (o && *o == '1') return true;
    if (o && *o == '0') return false;
    if (!ctx || needs_v2) return false;
    return sizeof(double) * (size_t)ctx->D * ctx->Cp > (size_t)4 << 20;
}

int bgh_nuts_run(bgh_ctx* ctx, const bgh_nuts_async_buffers* buf, const bgh_nuts_cfg* cfg, int32_t tune,
                 int32_t draws, int32_t poll_ticks, void* stream, int64_t* n_ticks)
{
    if (!buf) return fail(ctx, BGH_EINVAL, "bgh_nuts_run: null argument");
    AsyncParams a;
    int rc = nuts_params(ctx, &buf->base, cfg, 0, &a.n);
    if (rc != BGH_OK) return rc;
    const bool needs_v2 = a.n.peer.world > 1 || buf->k6 != nullptr || buf->head_trace != nullptr || !buf->trace;
    if (a.n.peer.world > 1 && (tick_v1(ctx, true) || (ctx->cfg.n_shared_rows != ctx->cfg.n_replicated)))
        return fail(ctx, BGH_ESTATE, "bgh_nuts_run: sharded contexts need replicated-gene sharding (n_replicated) and the fused "
                                     "sampler kernel; other sharded contexts use bgh_nuts_transition");
    if (!buf->afl || (draws > 0 && !buf->trace && !buf->k6) || tune < 0 || draws < 0 || tune + draws <= 0)
        return fail(ctx, BGH_EINVAL, "bgh_nuts_run: bad argument");
    if (ctx->Cp > 128) return fail(ctx, BGH_EINVAL, "bgh_nuts_run: at most 128 chains per context (use several contexts / GPUs)");
    if (poll_ticks <= 0) poll_ticks = 64;
    a.afl = buf->afl; a.trace = buf->trace; a.stats = buf->run_stats; a.n_finished = buf->base.n_active_dev;
    a.trace_f32 = buf->trace_f32; a.tune = tune; a.draws = draws;
    a.early_draws = 200; a.early_depth = 8; a.max_depth = kNutsMaxDepth; a.window = 101;
    // stream capture is not allowed on the legacy default stream, which is what callers usually pass:
    // run on the context's own stream, ordered after everything queued on the caller's stream
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    if (!ctx->st_graph) BGH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->st_graph, cudaStreamNonBlocking));
    BGH_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    cudaStream_t s = ctx->st_graph;
    const size_t vec = (size_t)ctx->D * ctx->Cp;
    double* const qw = buf->base.state + kSlotQW * vec;
    double* const gw = buf->base.state + kSlotGW * vec;
    const int Dtot = cfg->n_dim_total > 0 ? cfg->n_dim_total : ctx->D;
    const double step0 = cfg->step_scale / pow((double)Dtot, 0.25);
    if (ctx->cfg.n_chains > ctx->n_cta) return fail(ctx, BGH_EINVAL, "bgh_nuts_run: more chains than CTAs");
    if (!ctx->fused) return fail(ctx, BGH_ESTATE, "bgh_nuts_run: needs the cooperative kernels (BGH_UNFUSED is set or the device cannot co-schedule the grid)");
    if (!tick_v1(ctx, needs_v2)) return nuts_run_tick2(ctx, buf, cfg, a, poll_ticks, s, n_ticks);
    if (buf->k6 || !buf->trace) return fail(ctx, BGH_EINVAL, "bgh_nuts_run: BGH_TICK_V1 needs a trace and has no streaming statistics");
    BGH_CUDA(ctx, nuts_async_launch_init(a, step0, 10.0, s));
    ctx->launches += 1;
    // Inside the run the sampler state is chain-major ([slot][Cp][D], see bgh_fused.cu): convert the slots
    // the caller filled ([D][Cp]) — current point, its gradient, mass matrix and the adaptation windows —
    // and give the working position a finite start (a copy of the current point) for every chain.
    static const int kIoSlots[] = {kSlotQP, kSlotGP, kSlotVar, kSlotFgMean, kSlotFgRaw, kSlotBgMean, kSlotBgRaw};
    double* const scratch = buf->base.state + kSlotPW * vec;
    BGH_CUDA(ctx, cudaMemcpyAsync(qw, buf->base.state + kSlotQP * vec, sizeof(double) * vec, cudaMemcpyDeviceToDevice, s));   // [D][Cp]
    for (int sl : kIoSlots) {
        double* const x = buf->base.state + (size_t)sl * vec;
        BGH_CUDA(ctx, launch_transpose_out(x, scratch, ctx->Cp, ctx->Cp, ctx->D, s));
        BGH_CUDA(ctx, cudaMemcpyAsync(x, scratch, sizeof(double) * vec, cudaMemcpyDeviceToDevice, s));
        ctx->launches += 1;
    }

    // a transition needs at most 2^10 - 1 leaves; bound the loop so that a logic error cannot hang the GPU
    const int64_t max_ticks = (int64_t)(tune + draws) * 1024 + 1024;
    int64_t ticks = 0;
    int status = BGH_OK;
    // persistent tick kernel (bgh_fused.cu): pre | likelihood | finalize | post | scalar per tick, grid
    // barriers in between; one cooperative launch runs up to `poll_ticks` ticks, then the host polls
    TickParams t;
    fill_params(ctx, qw, gw, buf->base.logp_w, ctx->d_payload, &t.f.lik, &t.f.fin);
    t.f.fin.finish_inline = 1;
    t.f.lik.trace = nullptr;
    // (the working position / gradient exchanged with the likelihood pass stay gene-major [D][Cp])
    t.f.bar = ctx->d_bar; t.f.status = ctx->d_status; t.f.chain_blocks = ctx->chain_blocks;
    t.a = a;
    t.a.n.n_vec_blocks = ctx->n_cta;
    t.n_ticks = poll_ticks;
    t.ticks_done = ctx->d_ticks;
    t.phase_trace = ctx->d_trace;   // NULL unless bgh_debug_trace enabled it (tools/trace_ticks.py)
    { const char* tf = getenv("BGH_TRACE_TICK"); t.trace_first = tf ? atoi(tf) : 0; }
    t.tick_base = 0;
    BGH_CUDA(ctx, cudaMemsetAsync(ctx->d_ticks, 0, sizeof(int), s));
    while (ticks < max_ticks) {
        t.f.bar_base = ctx->bar_count;
        cudaError_t e = launch_nuts_ticks(t, ctx->CL, ctx->CPL, ctx->n_cta, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&ctx->h_poll[0], buf->base.n_active_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&ctx->h_poll[1], ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&ctx->h_poll[2], ctx->d_ticks, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { status = fail(ctx, BGH_ECUDA, std::string("bgh_nuts_run: ") + cudaGetErrorString(e)); break; }
        ctx->launches += 1;
        ctx->bar_count += 5ull * (unsigned long long)ctx->n_cta * (unsigned long long)(ctx->h_poll[2] - ticks);
        ticks = ctx->h_poll[2];
        t.tick_base = (int)ticks;
        if (ctx->h_poll[1] != 0) {
            cudaMemset(ctx->d_bar, 0, 256);
            cudaMemset(ctx->d_status, 0, sizeof(int));
            ctx->bar_count = 0;
            status = fail(ctx, BGH_ECUDA, "bgh_nuts_run: grid barrier timed out inside the tick kernel");
            break;
        }
        *buf->base.n_active_host = ctx->h_poll[0];
        if (ctx->h_poll[0] >= ctx->cfg.n_chains) break;
    }
    // back to the caller's [D][Cp] layout
    for (int sl : kIoSlots) {
        double* const x = buf->base.state + (size_t)sl * vec;
        cudaError_t e = launch_transpose_in(x, scratch, ctx->Cp, ctx->Cp, ctx->D, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(x, scratch, sizeof(double) * vec, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess && status == BGH_OK) status = fail(ctx, BGH_ECUDA, std::string("bgh_nuts_run: ") + cudaGetErrorString(e));
        ctx->launches += 1;
    }
    {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess && status == BGH_OK) status = fail(ctx, BGH_ECUDA, std::string("bgh_nuts_run: ") + cudaGetErrorString(e));
    }
    if (status == BGH_OK && ticks >= max_ticks) status = fail(ctx, BGH_ESTATE, "bgh_nuts_run: tick budget exhausted");
    if (n_ticks) *n_ticks = ticks;
    return status;
}

}  // extern "C"
