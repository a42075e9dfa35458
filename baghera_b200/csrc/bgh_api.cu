// bgh_api.cu — context management, host-side work planning and the C ABI (include/baghera_b200.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "bgh_internal.h"

namespace bgh {

// ---------------------------------------------------------------------------------------------
// work planning (pure host code)
// ---------------------------------------------------------------------------------------------
int build_plan(int64_t M, const int32_t* gid, int32_t G, int n_groups, bool first_partial,
               bool last_partial, int first_row, int last_row, Plan* plan, std::string* err)
{
    if (M < 0 || M > 0x7fffff00LL || n_groups <= 0 || G <= 0) {
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
    P.off.assign((size_t)n_groups + 1, 0);
    P.flags.assign((size_t)n_groups, 0);

    // cuts: ideal equal-SNP cuts snapped to the nearest gene boundary when one lies within
    // 1/16 of a range length, so that only genes longer than a range are ever split.
    const double len = (double)M / (double)n_groups;
    const int64_t tol = (int64_t)floor(len / 16.0);
    P.off[0] = 0;
    P.off[(size_t)n_groups] = (int32_t)M;
    for (int k = 1; k < n_groups; ++k) {
        int64_t x = (int64_t)llround((double)k * len);
        if (x < P.off[(size_t)k - 1]) x = P.off[(size_t)k - 1];
        if (x > M) x = M;
        if (tol > 0) {
            // nearest gene boundary to x
            auto it = std::lower_bound(goff.begin(), goff.end(), (int32_t)x);
            int64_t best = -1;
            if (it != goff.end() && llabs((int64_t)*it - x) <= tol) best = *it;
            if (it != goff.begin()) {
                const int64_t lo = *(it - 1);
                if (llabs(lo - x) <= tol && (best < 0 || llabs(lo - x) < llabs(best - x))) best = lo;
            }
            if (best >= P.off[(size_t)k - 1]) x = best;
        }
        P.off[(size_t)k] = (int32_t)x;
    }

    // flags + slot lists, mirroring the kernel's flush rule
    std::vector<std::vector<int32_t>> lists;       // per split gene, in discovery (= gene) order
    std::vector<int32_t> list_gene;
    auto add_slot = [&](int32_t gene, int32_t slot) {
        if (list_gene.empty() || list_gene.back() != gene) {
            list_gene.push_back(gene);
            lists.emplace_back();
        }
        lists.back().push_back(slot);
    };
    for (int v = 0; v < n_groups; ++v) {
        const int32_t a = P.off[(size_t)v], b = P.off[(size_t)v + 1];
        if (a >= b) continue;
        const int32_t gf = gid[a], gl = gid[b - 1];
        const bool head = (a > 0) ? (gid[a - 1] == gf) : first_partial;
        const bool tail = (b < M) ? (gid[b] == gl) : last_partial;
        P.flags[(size_t)v] = (head ? kHeadPartial : 0) | (tail ? kTailPartial : 0);
        if (head) add_slot(gf, v);
        if (tail && !(gl == gf && head)) add_slot(gl, n_groups + v);
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
        P.split_row.push_back(row);
        for (int32_t s : lists[i]) P.split_slots.push_back(s);
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
    // device data
    float* d_chi2 = nullptr;
    float* d_ld = nullptr;
    int* d_gid = nullptr;
    int4* d_groups = nullptr;
    double* d_slots = nullptr;
    double* d_part = nullptr;
    double* d_payload = nullptr;
    int* d_split_gene = nullptr;
    int* d_split_row = nullptr;
    int* d_split_ptr = nullptr;
    int* d_split_slots = nullptr;
    // host-entry staging (lazily allocated)
    static constexpr int kPipe = 2;
    double* d_q_cd[kPipe] = {nullptr, nullptr};
    double* d_q_dc[kPipe] = {nullptr, nullptr};
    double* d_g_dc[kPipe] = {nullptr, nullptr};
    double* d_g_cd[kPipe] = {nullptr, nullptr};
    double* d_lp[kPipe] = {nullptr, nullptr};
    cudaStream_t st_h2d = nullptr, st_cmp = nullptr, st_d2h = nullptr;
    cudaEvent_t ev_up[kPipe] = {nullptr, nullptr}, ev_cmp[kPipe] = {nullptr, nullptr},
                ev_dn[kPipe] = {nullptr, nullptr};
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

static void pick_chain_geometry(int C, int* CL, int* CPL, int* Cp)
{
    if (C <= 4) { *CL = 4; *CPL = 1; *Cp = 4; }
    else if (C <= 8) { *CL = 8; *CPL = 1; *Cp = 8; }
    else if (C <= 16) { *CL = 16; *CPL = 1; *Cp = 16; }
    else if (C <= 32) { *CL = 32; *CPL = 1; *Cp = 32; }
    else { *CL = 32; *CPL = 2; *Cp = ((C + 63) / 64) * 64; }
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
    if (cfg->model != BGH_MODEL_GENE_NORMAL && cfg->model != BGH_MODEL_GW_NORMAL)
        return fail(nullptr, BGH_EINVAL, "bgh_create: unknown model");
    if (cfg->n_genes <= 0 || (cfg->model == BGH_MODEL_GW_NORMAL && cfg->n_genes != 1))
        return fail(nullptr, BGH_EINVAL, "bgh_create: bad n_genes");
    if (!(cfg->c > 0.0) || !isfinite(cfg->c)) return fail(nullptr, BGH_EINVAL, "bgh_create: c must be positive");
    if (cfg->n_shared_rows < 0 || cfg->first_gene_row >= cfg->n_shared_rows || cfg->last_gene_row >= cfg->n_shared_rows)
        return fail(nullptr, BGH_EINVAL, "bgh_create: inconsistent shared-gene rows");

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
    ctx->D = (cfg->model == BGH_MODEL_GENE_NORMAL) ? cfg->n_genes + 3 : 2;
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
    ctx->n_groups = ctx->n_cta * kLikWarps * (32 / ctx->CL);
    e = lik_configure();
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
    cudaFree(ctx->d_chi2); cudaFree(ctx->d_ld); cudaFree(ctx->d_gid); cudaFree(ctx->d_groups);
    cudaFree(ctx->d_slots); cudaFree(ctx->d_part); cudaFree(ctx->d_payload);
    cudaFree(ctx->d_split_gene); cudaFree(ctx->d_split_row); cudaFree(ctx->d_split_ptr); cudaFree(ctx->d_split_slots);
    for (int i = 0; i < bgh_ctx::kPipe; ++i) {
        cudaFree(ctx->d_q_cd[i]); cudaFree(ctx->d_q_dc[i]); cudaFree(ctx->d_g_dc[i]); cudaFree(ctx->d_g_cd[i]);
        cudaFree(ctx->d_lp[i]);
        if (ctx->ev_up[i]) cudaEventDestroy(ctx->ev_up[i]);
        if (ctx->ev_cmp[i]) cudaEventDestroy(ctx->ev_cmp[i]);
        if (ctx->ev_dn[i]) cudaEventDestroy(ctx->ev_dn[i]);
    }
    if (ctx->st_h2d) cudaStreamDestroy(ctx->st_h2d);
    if (ctx->st_cmp) cudaStreamDestroy(ctx->st_cmp);
    if (ctx->st_d2h) cudaStreamDestroy(ctx->st_d2h);
    for (cudaEvent_t ev : ctx->prof_ev) cudaEventDestroy(ev);
    delete ctx;
}

int32_t bgh_dim(const bgh_ctx* ctx) { return ctx ? ctx->D : 0; }
int32_t bgh_chain_stride(const bgh_ctx* ctx) { return ctx ? ctx->Cp : 0; }
int32_t bgh_payload_rows(const bgh_ctx* ctx) { return ctx ? 4 + ctx->cfg.n_shared_rows : 0; }
double bgh_lik_const(const bgh_ctx* ctx) { return ctx ? ctx->lik_const : 0.0; }
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

// Host-only planning entry point used by the CPU tests (no device needed): sorted gene ids in,
// offsets[n_groups+1], flags[n_groups] out; split CSR returned through caller-sized buffers.
int bgh_debug_plan(int64_t M, const int32_t* gid_sorted, int32_t G, int32_t n_groups, int32_t first_partial,
                   int32_t last_partial, int32_t* offsets, int32_t* flags, int32_t* n_split,
                   int32_t* split_gene, int32_t* split_ptr, int32_t* split_slots, int32_t cap)
{
    Plan P;
    std::string err;
    int rc = build_plan(M, gid_sorted, G, n_groups, first_partial != 0, last_partial != 0,
                        first_partial ? 0 : -1, last_partial ? (first_partial && G == 1 ? 0 : (first_partial ? 1 : 0)) : -1,
                        &P, &err);
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
    const int64_t Mpad = ((M + kChunk - 1) / kChunk) * kChunk + kChunk;
    std::vector<float> s_sorted((size_t)Mpad, 0.0f), l_sorted((size_t)Mpad, 1.0f);
    std::vector<int32_t> g_sorted((size_t)Mpad, G - 1);
    {
        std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
        for (int64_t j = 0; j < M; ++j) {
            const int32_t g = gene_id ? gene_id[j] : 0;
            const int64_t d = pos[(size_t)g]++;
            s_sorted[(size_t)d] = chi2[j];
            l_sorted[(size_t)d] = ld[j];
            g_sorted[(size_t)d] = g;
        }
    }
    double acc = 0.0;
    for (int64_t j = 0; j < M; ++j) {
        const float l = l_sorted[(size_t)j];
        if (!(l > 0.0f) || !isfinite(l) || !isfinite(s_sorted[(size_t)j]))
            return fail(ctx, BGH_EINVAL, "bgh_upload: ld must be finite and > 0, chi2 finite (snps.py:121 gives l >= 1)");
        acc += log((double)l);
    }
    ctx->lik_const = -0.5 * acc - 0.5 * (double)M * kLog2Pi;

    const bool gw = ctx->cfg.model == BGH_MODEL_GW_NORMAL;
    const bool fp = gw ? true : ctx->cfg.first_gene_partial != 0;
    const bool lp = gw ? true : ctx->cfg.last_gene_partial != 0;
    std::string err;
    int rc = build_plan(M, g_sorted.data(), G, ctx->n_groups, fp, lp, ctx->cfg.first_gene_row,
                        ctx->cfg.last_gene_row, &ctx->plan, &err);
    if (rc != BGH_OK) return fail(ctx, rc, err);

    auto dfree = [](auto*& p) { cudaFree(p); p = nullptr; };
    dfree(ctx->d_chi2); dfree(ctx->d_ld); dfree(ctx->d_gid); dfree(ctx->d_groups); dfree(ctx->d_slots);
    dfree(ctx->d_part); dfree(ctx->d_payload); dfree(ctx->d_split_gene); dfree(ctx->d_split_row);
    dfree(ctx->d_split_ptr); dfree(ctx->d_split_slots);

    const Plan& P = ctx->plan;
    std::vector<int4> groups((size_t)ctx->n_groups);
    for (int v = 0; v < ctx->n_groups; ++v)
        groups[(size_t)v] = make_int4(P.off[(size_t)v], P.off[(size_t)v + 1], P.flags[(size_t)v], 0);
    const size_t ns = P.split_gene.size();
    const size_t nsl = std::max<size_t>(P.split_slots.size(), 1);

    BGH_CUDA(ctx, cudaMalloc(&ctx->d_chi2, sizeof(float) * (size_t)Mpad));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_ld, sizeof(float) * (size_t)Mpad));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_gid, sizeof(int) * (size_t)Mpad));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_groups, sizeof(int4) * groups.size()));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_slots, sizeof(double) * 2 * (size_t)ctx->n_groups * ctx->Cp));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_part, sizeof(double) * 4 * (size_t)ctx->n_cta * ctx->Cp));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_payload, sizeof(double) * (size_t)(4 + ctx->cfg.n_shared_rows) * ctx->Cp));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_gene, sizeof(int) * std::max<size_t>(ns, 1)));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_row, sizeof(int) * std::max<size_t>(ns, 1)));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_ptr, sizeof(int) * (ns + 1)));
    BGH_CUDA(ctx, cudaMalloc(&ctx->d_split_slots, sizeof(int) * nsl));
    BGH_CUDA(ctx, cudaMemcpy(ctx->d_chi2, s_sorted.data(), sizeof(float) * (size_t)Mpad, cudaMemcpyHostToDevice));
    BGH_CUDA(ctx, cudaMemcpy(ctx->d_ld, l_sorted.data(), sizeof(float) * (size_t)Mpad, cudaMemcpyHostToDevice));
    BGH_CUDA(ctx, cudaMemcpy(ctx->d_gid, g_sorted.data(), sizeof(int) * (size_t)Mpad, cudaMemcpyHostToDevice));
    BGH_CUDA(ctx, cudaMemcpy(ctx->d_groups, groups.data(), sizeof(int4) * groups.size(), cudaMemcpyHostToDevice));
    BGH_CUDA(ctx, cudaMemset(ctx->d_slots, 0, sizeof(double) * 2 * (size_t)ctx->n_groups * ctx->Cp));
    BGH_CUDA(ctx, cudaMemset(ctx->d_payload, 0, sizeof(double) * (size_t)(4 + ctx->cfg.n_shared_rows) * ctx->Cp));
    if (ns) {
        BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_gene, P.split_gene.data(), sizeof(int) * ns, cudaMemcpyHostToDevice));
        BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_row, P.split_row.data(), sizeof(int) * ns, cudaMemcpyHostToDevice));
        BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_slots, P.split_slots.data(), sizeof(int) * P.split_slots.size(), cudaMemcpyHostToDevice));
    }
    BGH_CUDA(ctx, cudaMemcpy(ctx->d_split_ptr, P.split_ptr.data(), sizeof(int) * (ns + 1), cudaMemcpyHostToDevice));
    ctx->uploaded = true;
    return BGH_OK;
}

static void fill_params(bgh_ctx* ctx, const double* q, double* grad, double* logp, double* payload,
                        LikParams* lp, FinParams* fp)
{
    lp->chi2 = ctx->d_chi2; lp->ld = ctx->d_ld; lp->gid = ctx->d_gid; lp->groups = ctx->d_groups;
    lp->q = q; lp->grad = grad; lp->slots = ctx->d_slots; lp->part = ctx->d_part;
    lp->Cp = ctx->Cp; lp->G = ctx->cfg.n_genes; lp->NG = ctx->n_groups; lp->model = ctx->cfg.model;
    lp->fix = ctx->cfg.fix_intercept; lp->c = ctx->cfg.c;
    fp->q = q; fp->grad = grad; fp->logp = logp; fp->slots = ctx->d_slots; fp->part = ctx->d_part;
    fp->payload = payload;
    fp->split_gene = ctx->d_split_gene; fp->split_row = ctx->d_split_row; fp->split_ptr = ctx->d_split_ptr;
    fp->split_slots = ctx->d_split_slots; fp->n_split = (int)ctx->plan.split_gene.size(); fp->n_long = ctx->plan.n_long;
    fp->n_cta = ctx->n_cta; fp->Cp = ctx->Cp; fp->G_total = ctx->cfg.n_genes_total; fp->model = ctx->cfg.model;
    fp->fix = ctx->cfg.fix_intercept; fp->finish_inline = 0; fp->c = ctx->cfg.c; fp->lik_const = ctx->lik_const;
    fp->first_gene_row = ctx->cfg.first_gene_row; fp->last_gene_row = ctx->cfg.last_gene_row;
    fp->G_local = ctx->cfg.n_genes;
}

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
    for (int i = 0; i < bgh_ctx::kPipe; ++i) {
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_q_cd[i], sizeof(double) * n_cd));
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_g_cd[i], sizeof(double) * n_cd));
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_q_dc[i], sizeof(double) * n_dc));
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_g_dc[i], sizeof(double) * n_dc));
        BGH_CUDA(ctx, cudaMalloc(&ctx->d_lp[i], sizeof(double) * (size_t)ctx->Cp));
        BGH_CUDA(ctx, cudaMemset(ctx->d_g_dc[i], 0, sizeof(double) * n_dc));
        BGH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_up[i], cudaEventDisableTiming));
        BGH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_cmp[i], cudaEventDisableTiming));
        BGH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_dn[i], cudaEventDisableTiming));
    }
    BGH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->st_h2d, cudaStreamNonBlocking));
    BGH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->st_cmp, cudaStreamNonBlocking));
    BGH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->st_d2h, cudaStreamNonBlocking));
    return BGH_OK;
}

int bgh_logp_grad_host_stream(bgh_ctx* ctx, int32_t n_evals, const double* q, double* logp, double* grad)
{
    if (!ctx || !q || !logp || !grad || n_evals <= 0) return fail(ctx, BGH_EINVAL, "bgh_logp_grad_host_stream: bad argument");
    if (!ctx->uploaded) return fail(ctx, BGH_ESTATE, "bgh_logp_grad_host_stream: call bgh_upload first");
    if (ctx->cfg.n_shared_rows != 0) return fail(ctx, BGH_ESTATE, "bgh_logp_grad_host_stream: single-rank contexts only");
    int rc = ensure_host_staging(ctx);
    if (rc != BGH_OK) return rc;
    BGH_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const int C = ctx->cfg.n_chains, D = ctx->D;
    const size_t n_cd = (size_t)C * D;
    for (int i = 0; i < n_evals; ++i) {
        const int b = i % bgh_ctx::kPipe;
        // buffer reuse: upload of eval i may overwrite q_cd[b] only after compute of eval i-2 consumed it,
        // compute may overwrite g_cd[b] only after the download of eval i-2 finished.
        if (i >= bgh_ctx::kPipe) {
            BGH_CUDA(ctx, cudaStreamWaitEvent(ctx->st_h2d, ctx->ev_cmp[b], 0));
            BGH_CUDA(ctx, cudaStreamWaitEvent(ctx->st_cmp, ctx->ev_dn[b], 0));
        }
        BGH_CUDA(ctx, cudaMemcpyAsync(ctx->d_q_cd[b], q + (size_t)i * n_cd, sizeof(double) * n_cd,
                                      cudaMemcpyHostToDevice, ctx->st_h2d));
        BGH_CUDA(ctx, cudaEventRecord(ctx->ev_up[b], ctx->st_h2d));
        BGH_CUDA(ctx, cudaStreamWaitEvent(ctx->st_cmp, ctx->ev_up[b], 0));
        BGH_CUDA(ctx, launch_transpose_in(ctx->d_q_cd[b], ctx->d_q_dc[b], C, ctx->Cp, D, ctx->st_cmp));
        rc = run_local(ctx, ctx->d_q_dc[b], ctx->d_lp[b], ctx->d_g_dc[b], ctx->d_payload, true, ctx->st_cmp);
        if (rc != BGH_OK) return rc;
        BGH_CUDA(ctx, launch_transpose_out(ctx->d_g_dc[b], ctx->d_g_cd[b], C, ctx->Cp, D, ctx->st_cmp));
        ctx->launches += 2;
        BGH_CUDA(ctx, cudaEventRecord(ctx->ev_cmp[b], ctx->st_cmp));
        BGH_CUDA(ctx, cudaStreamWaitEvent(ctx->st_d2h, ctx->ev_cmp[b], 0));
        BGH_CUDA(ctx, cudaMemcpyAsync(grad + (size_t)i * n_cd, ctx->d_g_cd[b], sizeof(double) * n_cd,
                                      cudaMemcpyDeviceToHost, ctx->st_d2h));
        BGH_CUDA(ctx, cudaMemcpyAsync(logp + (size_t)i * C, ctx->d_lp[b], sizeof(double) * (size_t)C,
                                      cudaMemcpyDeviceToHost, ctx->st_d2h));
        BGH_CUDA(ctx, cudaEventRecord(ctx->ev_dn[b], ctx->st_d2h));
    }
    BGH_CUDA(ctx, cudaStreamSynchronize(ctx->st_d2h));
    BGH_CUDA(ctx, cudaStreamSynchronize(ctx->st_cmp));
    BGH_CUDA(ctx, cudaStreamSynchronize(ctx->st_h2d));
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

}  // extern "C"
