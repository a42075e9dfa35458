/*
 * baghera_b200.h — C ABI of the B200-native BAGHERA hot path.
 *
 * The reference (stracquadaniolab/baghera v2.2.0) has no FFI: the path sits behind a Python
 * operator boundary and, inside PyMC3, a numerical callback.  Each entry point below cites
 * the reference interface it replaces (paths are relative to the reference checkout):
 *
 *   B1  analyse_normal(snps_object, ..., SWEEPS, TUNE, CHAINS, CORES, N_1kG, fix_intercept)
 *       baghera_tool/gene_regression.py:30-40   (model build :77-98, pm.sample :100-106)
 *   B1' model.logp_dlogp_function()(q[D]) -> (logp, grad[D])     [PyMC3 3.6, pymc3/model.py]
 *       parameter order = variable creation order, gene_regression.py:78-81:
 *           q = [W_log__, e, mi_logodds__, beta_0 .. beta_{G-1}]         D = G + 3
 *       genome-wide model, baghera_tool/heritability.py:44-55:
 *           q = [e, mi_logodds__]                                        D = 2
 *
 * Conventions
 *   - every function returns 0 on success or a negative BGH_E* code; the message is available
 *     from bgh_last_error(ctx) (bgh_last_error(NULL) for failures of bgh_create).  No C++
 *     exception crosses this boundary.
 *   - a context is bound to one CUDA device, owns the uploaded SNP arrays and all scratch, and is
 *     NOT thread-safe: one context per GPU / rank.
 *   - "device" pointers are plain CUDA device pointers owned by the caller (e.g. torch tensors'
 *     data_ptr()); `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Device entry points are asynchronous on that stream and never synchronise.
 *   - device parameter layout is gene-major / chain-minor:  q[d * Cp + chain],  d in [0, D),
 *     Cp = bgh_chain_stride(ctx) >= C (chains padded to the kernel's chain-group width; padded
 *     chains must hold finite values, e.g. a copy of chain 0; their outputs are unspecified).
 *   - host entry points use PyMC3's chain-major layout q[chain * D + d] and exclude padding.
 */
#ifndef BAGHERA_B200_H
#define BAGHERA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGH_ABI_VERSION 1

enum {
    BGH_OK = 0,
    BGH_EINVAL = -1,   /* bad argument / inconsistent configuration       */
    BGH_ECUDA = -2,    /* CUDA runtime error (message has the CUDA string) */
    BGH_ENOMEM = -3,   /* host allocation failed                           */
    BGH_ESTATE = -4,   /* call order violated (e.g. logp before upload)    */
    BGH_ENODEV = -5    /* no usable sm_100 device                          */
};

enum {
    BGH_MODEL_GENE_NORMAL = 0, /* gene_regression.py:77-98   */
    BGH_MODEL_GW_NORMAL = 1,   /* heritability.py:44-55      */
    BGH_MODEL_GENE_GAMMA = 2   /* gene_regression.py:229-248: q = [e, mi_logodds__, beta_log___0..G-1], D = G + 2;
                                  pass c = n_patients (the likelihood mean is n * beta * l + e) and gamma_rate = N_1kG */
};

typedef struct bgh_ctx bgh_ctx;

typedef struct bgh_cfg {
    int32_t abi_version;   /* BGH_ABI_VERSION */
    int32_t device;        /* CUDA device ordinal */
    int32_t model;         /* BGH_MODEL_* */
    int32_t fix_intercept; /* gene_regression.py:84-90 (e still sampled from its prior, Q8) */
    int64_t n_snps;        /* M: SNPs in THIS context (this rank's shard) */
    int32_t n_genes;       /* G: gene labels present in this shard, ids 0..G-1 (gw model: 1) */
    int32_t n_chains;      /* C */
    double  c;             /* n_patients / N_1kG  (gene_regression.py:95) */
    /* --- gene-block sharding (single GPU: leave zero / -1 as set by bgh_cfg_default) --- */
    int32_t n_genes_total;     /* G over all ranks; 0 => n_genes */
    int32_t first_gene_partial;/* 1: this shard's first gene started on a previous rank */
    int32_t last_gene_partial; /* 1: this shard's last gene continues on the next rank  */
    int32_t first_gene_row;    /* payload row (>=0) of the first gene if it is rank-shared, else -1 */
    int32_t last_gene_row;     /* payload row of the last gene if rank-shared, else -1 */
    int32_t n_shared_rows;     /* rank-shared genes over all ranks (payload rows 4 .. 4+n-1) */
    int32_t sm_count_override; /* 0 = query the device; tests may set it to build partitions on CPU */
    int32_t n_replicated;      /* gene-block sharding with REPLICATED genes: this shard's genes [0, n_replicated) are held (with
                                * a share of their SNPs) by EVERY rank; gene j travels in payload row j (n_shared_rows must
                                * equal n_replicated, first/last_gene_* stay 0 / -1).  Their rows of q are replicated too. */
    double  gamma_rate;        /* BGH_MODEL_GENE_GAMMA: rate of the Gamma(mi, rate) prior = N_1kG (gene_regression.py:231) */
    int32_t owns_replicated;   /* 1 on the one rank (rank 0) that counts the replicated genes' prior terms */
    int32_t reserved;
} bgh_cfg;

/* Fills *cfg with single-GPU defaults. */
void bgh_cfg_default(bgh_cfg* cfg);

/* Library / build identification ("baghera_b200 <abi> sm_100a ..."). */
const char* bgh_version(void);

/* Creates a context on cfg->device.  Fails with BGH_ENODEV when no CUDA device is usable:
 * there is no CPU fallback.  Replaces the `with pm.Model()` block, gene_regression.py:77. */
int bgh_create(bgh_ctx** out, const bgh_cfg* cfg);
void bgh_destroy(bgh_ctx* ctx);
const char* bgh_last_error(const bgh_ctx* ctx);

/* Geometry the caller needs to size its tensors. */
int32_t bgh_dim(const bgh_ctx* ctx);           /* D  (G+3 or 2) */
int32_t bgh_chain_stride(const bgh_ctx* ctx);  /* Cp */
int32_t bgh_payload_rows(const bgh_ctx* ctx);  /* 4 + n_shared_rows: rows of the all-reduce payload */

/* Uploads the observed data in FILE order (host pointers): chi2 = stats = z^2 (snps.py:78),
 * ld = transformed LD score >= 1 (snps.py:121), gene_id = rank of the SNP's gene label in sorted
 * label order (gene_regression.py:61-71), unsorted (Q6).  The library stable-sorts by gene,
 * builds the warp partition and precomputes sum_j -1/2 log(2 pi l_j).  gene_id may be NULL for
 * the gw model.  Replaces the observed-variable construction gene_regression.py:93-98. */
int bgh_upload(bgh_ctx* ctx, const float* chi2, const float* ld, const int32_t* gene_id);

/* sum_j -1/2 log(2 pi l_j), the chain-independent constant of the likelihood (after upload). */
double bgh_lik_const(const bgh_ctx* ctx);
/* Sharded runs: every rank uploads its shard, the ranks all-reduce (sum) bgh_lik_const once and
 * install the total here so that each rank's logp carries the full constant. */
int bgh_set_lik_const(bgh_ctx* ctx, double total);

/* B1' on the device: q, grad are [D][Cp], logp is [Cp]; async on `stream`.
 * Single-rank contexts only (n_shared_rows == 0). */
int bgh_logp_grad(bgh_ctx* ctx, const double* q, double* logp, double* grad, void* stream);

/* Sharded form: local pass writes this rank's gradient rows for its non-shared genes and the
 * payload [bgh_payload_rows][Cp] = {sum r^2/l, sum r/l, sum(beta-mi), sum(beta-mi)^2, shared-gene
 * sums ...}; the caller all-reduces (sum) the payload over ranks (NCCL) and calls _finish, which
 * writes logp[Cp], grad rows 0..2 and the gradient rows of the rank-shared genes. */
int bgh_logp_grad_local(bgh_ctx* ctx, const double* q, double* grad, double* payload, void* stream);
int bgh_logp_grad_finish(bgh_ctx* ctx, const double* q, const double* payload, double* logp,
                         double* grad, void* stream);

/* Sharded form with the exchange fused into the kernel (one launch per rank and evaluation): after
 * the local pass CTA 0 stores the payload into every peer's mailbox over NVLink (peer-mapped memory),
 * raises a system-scope flag, waits for the peers and adds the contributions in rank order, then
 * closes the formulas — no NCCL call on the hot path.  Setup (once): every rank calls
 * bgh_peer_export (allocates its mailbox, returns a bgh_peer_handle_bytes()-byte CUDA IPC handle), the
 * ranks all-gather the handles (torch.distributed) and call bgh_peer_connect(rank, world, handles);
 * a barrier must separate the last connect from the first evaluation.  Every rank must call
 * bgh_logp_grad_sharded the same number of times; one process per GPU, at most 8 ranks of one node.
 * bgh_peer_status returns the device status word (0 ok, 1 grid barrier timeout, 2 peer timeout). */
int32_t bgh_peer_handle_bytes(void);
int bgh_peer_export(bgh_ctx* ctx, void* handle_out);
int bgh_peer_connect(bgh_ctx* ctx, int32_t rank, int32_t world, const void* handles);
int bgh_logp_grad_sharded(bgh_ctx* ctx, const double* q, double* logp, double* grad, void* stream);
/* Ranks as contexts of ONE process (several GPUs driven by one process, or tests that emulate the ranks one after the other
 * on one GPU): peers[r] is rank r's context (bgh_peer_export already called on each). */
int bgh_peer_connect_local(bgh_ctx* ctx, int32_t rank, int32_t world, bgh_ctx* const* peers);
/* Tests only: send_only != 0 makes the next evaluations deposit their payload in the peers' mailboxes without waiting for
 * the peers' (results are garbage) — a single GPU can then run the ranks one after the other: every rank once send-only,
 * then every rank normally with the same sequence number.  seq >= 0 sets the mailbox sequence number of the next evaluation. */
int bgh_debug_peer_mode(bgh_ctx* ctx, int32_t send_only, int32_t seq);
int bgh_peer_status(bgh_ctx* ctx, int32_t* status);

/* B1' with HOST buffers in PyMC3's layout: q[C][D] -> logp[C], grad[C][D].  Copies H2D, transposes
 * on the device, evaluates, transposes back and copies D2H; synchronous.  Buffers may be pageable
 * or pinned (pinned is faster). */
int bgh_logp_grad_host(bgh_ctx* ctx, const double* q, double* logp, double* grad);

/* Pipelined host form: evaluates `n_evals` parameter batches q[n][C][D] -> logp[n][C], grad[n][C][D]
 * with upload / compute / download of consecutive batches overlapped (copies alternate between two
 * streams per direction, batches travel in groups of about 16 MB per copy, four groups in flight).
 * Synchronous on return.  Sharded contexts (after bgh_peer_connect): q / grad are this rank's rows
 * [C][3 + local genes]; every rank must make the same calls with the same n_evals. */
int bgh_logp_grad_host_stream(bgh_ctx* ctx, int32_t n_evals, const double* q, double* logp, double* grad);

/* Layout helpers on the device: chain-major [C][D] <-> gene-major [D][Cp]. */
int bgh_to_device_layout(bgh_ctx* ctx, const double* q_cd, double* q_dc, void* stream);
int bgh_from_device_layout(bgh_ctx* ctx, const double* x_dc, double* x_cd, void* stream);

/* Introspection for tests: the SNP-range partition (n_groups+1 offsets into the gene-sorted SNP
 * array) and kernel launch counters (how many of this library's kernels were launched). */
int32_t bgh_n_groups(const bgh_ctx* ctx);
int bgh_get_partition(const bgh_ctx* ctx, int32_t* offsets, int32_t n);
/* the same for the sampler kernel's own work plan (bgh_nuts_run): CTA cuts weigh a gene by the leapfrog of its row,
 * the cuts between the warps of a CTA by what a gene boundary costs inside the likelihood pass */
int bgh_get_sampler_partition(const bgh_ctx* ctx, int32_t* offsets, int32_t n);
int32_t bgh_n_split_genes(const bgh_ctx* ctx);
/* Introspection for tests: the gene-sorted, padded stream bgh_upload built on the device.  which = 0 / 1 / 2:
 * s/sqrt(l), 1/sqrt(l), sqrt(l) as double[n]; which = 3: compact gene id per position as int32[n].
 * bgh_stream_len = padded length (every gene a multiple of the kernel's batch of 8 SNPs, plus chunk slack). */
int64_t bgh_stream_len(const bgh_ctx* ctx);
int bgh_debug_stream(bgh_ctx* ctx, int32_t which, void* out, int64_t n);
int64_t bgh_launch_count(const bgh_ctx* ctx);
/* average device time (ms) of the likelihood kernel since the last call with reset != 0; timing is
 * only collected after bgh_set_profiling(ctx, 1) (adds CUDA events around each launch). */
int bgh_set_profiling(bgh_ctx* ctx, int32_t enable);
int bgh_kernel_time_ms(bgh_ctx* ctx, double* lik_ms, double* finalize_ms, int64_t* n, int32_t reset);

/* Host-only planning (no device needed; used by the CPU tests): gid_sorted[M] ascending compact
 * gene ids -> offsets[n_groups+1], flags[n_groups] (bit0 head-partial, bit1 tail-partial, bit2
 * exclusive CTA, bit3 prior owner) with n_groups = n_cta * groups_per_cta, and the CSR of genes
 * whose sum is assembled from slot partials (slot s < n_groups: head slot of group s; s < 2
 * n_groups: tail slot of group s - n_groups; else CTA slot s - 2 n_groups).  cap = capacity of
 * the split_* buffers. */
int bgh_debug_plan(int64_t M, const int32_t* gid_sorted, int32_t G, int32_t n_cta,
                   int32_t groups_per_cta, int32_t first_partial, int32_t last_partial, int32_t* offsets, int32_t* flags,
                   int32_t* n_split, int32_t* split_gene, int32_t* split_ptr, int32_t* split_slots,
                   int32_t cap);

/* ------------------------------------------------------------------------------------------
 * Batched-chain NUTS (replaces pm.sample(..., nuts_kwargs=dict(target_accept=0.90)),
 * gene_regression.py:100-106 / heritability.py:56-62; PyMC3 3.6 semantics).
 * The caller (the PyTorch host driver) owns all sampler memory and passes it as plain pointers:
 *   state    double[BGH_NUTS_N_SLOTS][D][Cp]   vector slots; the caller reads/writes slots
 *                                              Q (current position), GRAD, VAR (mass-matrix diagonal)
 *                                              and the four adaptation slots, the rest is scratch
 *   ckpt     double[2][BGH_NUTS_MAX_DEPTH][D][Cp]
 *   sc       double[BGH_NUTS_N_SC][Cp]         per-chain scalars (EPS = step size, LOGP = logp(Q))
 *   fl       int32 [BGH_NUTS_N_FL][Cp]
 *   partials double[bgh_nuts_vec_blocks()][BGH_NUTS_MAX_PARTIALS][Cp]
 *   sums     double[BGH_NUTS_MAX_PARTIALS][Cp]
 *   logp_w   double[Cp]
 *   stats    double[BGH_NUTS_N_STATS][Cp]      {depth, mean_tree_accept, tree_size, diverging, energy,
 *                                               step_size, model_logp, max_energy_error} of the last transition
 *   n_active_dev / n_active_host: one int32 on the device / in pinned host memory
 * ------------------------------------------------------------------------------------------ */
#define BGH_NUTS_MAX_DEPTH 10
#define BGH_NUTS_MAX_PARTIALS (1 + 2 * BGH_NUTS_MAX_DEPTH)
#define BGH_NUTS_N_SLOTS 21
#define BGH_NUTS_SLOT_Q 0
#define BGH_NUTS_SLOT_GRAD 1
#define BGH_NUTS_SLOT_VAR 2
#define BGH_NUTS_SLOT_FG_MEAN 16
#define BGH_NUTS_SLOT_FG_RAW 17
#define BGH_NUTS_SLOT_BG_MEAN 18
#define BGH_NUTS_SLOT_BG_RAW 19
#define BGH_NUTS_N_SC 16
#define BGH_NUTS_SC_EPS 0
#define BGH_NUTS_SC_LOGP 1
#define BGH_NUTS_N_FL 11
#define BGH_NUTS_N_STATS 8
#define BGH_NUTS_ASYNC_N_FL 26
#define BGH_NUTS_ASYNC_N_SC 24
#define BGH_NUTS_ASYNC_PARTIALS (2 * BGH_NUTS_MAX_DEPTH + 4)

typedef struct bgh_nuts_buffers {
    double* state;
    double* ckpt;
    double* sc;
    int32_t* fl;
    double* partials;
    double* sums;
    double* logp_w;
    double* stats;
    int32_t* n_active_dev;
    int32_t* n_active_host;
} bgh_nuts_buffers;

typedef struct bgh_nuts_cfg {
    uint64_t seed;          /* Philox key; chains and draws select the counter */
    double target_accept;   /* 0.90 in the reference (gene_regression.py:105) */
    double step_scale;      /* 0.25: initial step = step_scale / D^(1/4) [PyMC3 BaseHMC] */
    double emax;            /* 1000: divergence threshold */
    double da_gamma, da_kappa, da_t0;   /* 0.05, 0.75, 10 */
    int32_t n_dim_total;    /* D over all ranks (step-size rule); 0 => bgh_dim(ctx) */
    int32_t coord_offset;   /* sharded contexts: genes in front of this rank's gene block (local per-gene row d is
                             * global row d + coord_offset; the shared rows keep their index); 0 on a single rank */
    int32_t chain_offset;   /* index of this context's chain 0 among all chains of the analysis (chains dealt to several
                             * devices): the Philox key of a chain is derived from (seed, chain_offset + chain), so the
                             * draws do not depend on how the chains are spread over devices */
    int32_t reserved[3];
} bgh_nuts_cfg;

void bgh_nuts_cfg_default(bgh_nuts_cfg* cfg);
int32_t bgh_nuts_vec_blocks(const bgh_ctx* ctx);

/* Evaluates logp/grad at state[Q] into state[GRAD] / sc[LOGP] and initialises the step-size
 * adaptation.  The caller has filled Q, VAR (ones), FG_MEAN (mean of the chains' starts), FG_RAW
 * (10 * ones), BG_* (zeros) — PyMC3's init='jitter+adapt_diag'. */
int bgh_nuts_init(bgh_ctx* ctx, const bgh_nuts_buffers* buf, const bgh_nuts_cfg* cfg, void* stream);

/* One NUTS transition of every chain.  tune != 0: step-size and mass-matrix adaptation run after
 * the transition; fg_weight / bg_weight are the estimator weights after adding this draw and
 * switch_window != 0 swaps the estimators (quadpotential.QuadPotentialDiagAdapt.update, driven by
 * the host because the schedule is identical for all chains).  Synchronises the stream once per
 * tree doubling (to learn whether any chain is still growing its tree).  *n_leapfrogs receives the
 * number of batched logp+grad evaluations.
 * Sharded contexts (after bgh_peer_connect; every rank calls with the same arguments, cfg->coord_offset =
 * genes in front of the rank's block, cfg->n_dim_total = global D): every rank advances its rows of q; the
 * per-chain sums (kinetic energy, U-turn dots) are all-reduced over peer memory inside the scalar kernels
 * in rank order, so all ranks take bit-identical decisions and reproduce the single-rank chains. */
int bgh_nuts_transition(bgh_ctx* ctx, const bgh_nuts_buffers* buf, const bgh_nuts_cfg* cfg, int64_t draw,
                        int32_t tune, int32_t max_treedepth, double fg_weight, double bg_weight,
                        int32_t switch_window, void* stream, int32_t* n_leapfrogs);

/* Sharded contexts: global coordinate of every local row of q (host int32[D]); the momentum normals are Philox draws keyed
 * by the GLOBAL coordinate, so a sharded run reproduces the single-GPU chains.  Default: row d -> d (+ coord_offset behind
 * the head rows). */
int bgh_set_row_coords(bgh_ctx* ctx, const int32_t* global_coord);

/* End of tuning: step size <- exp(log_step_bar) (DualAverageAdaptation.current(tune=False)). */
int bgh_nuts_stop_tuning(bgh_ctx* ctx, const bgh_nuts_buffers* buf, const bgh_nuts_cfg* cfg, void* stream);

/* ---- asynchronous batched NUTS: the whole pm.sample(draws, tune) run on the device --------------
 * Every "tick" (one batched logp+grad evaluation) advances every chain by one leapfrog; a chain whose
 * tree stops starts its next transition in the next tick, so no chain waits for the deepest tree of
 * the batch.  Per-chain decisions depend only on the chain's own state and on Philox counters keyed by
 * its own draw index, hence the result equals the lock-step bgh_nuts_transition loop.  One tick is
 * captured in a CUDA graph; the host only polls the number of finished chains every `poll_ticks`.
 * Buffers: as bgh_nuts_buffers, but sc has BGH_NUTS_ASYNC_N_SC rows and partials
 * BGH_NUTS_ASYNC_PARTIALS quantities per block; plus
 *   afl      int32[BGH_NUTS_ASYNC_N_FL][Cp]
 *   trace    [draws][C][D] fp64 (trace_f32 = 0) or fp32: positions after each post-tuning transition
 *   stats    double[tune + draws][BGH_NUTS_N_STATS][C]
 * Call bgh_nuts_init first (start point, mass-matrix slots); then bgh_nuts_run once. */
typedef struct bgh_nuts_async_buffers {
    bgh_nuts_buffers base;
    int32_t* afl;
    void* trace;
    double* run_stats;
    int32_t trace_f32;
    int32_t reserved;
    double* k6;     /* optional [3][D][Cp] (zeroed by the caller): streaming per-coordinate sums over the recorded draws
                     * {sum v, sum v^2, count(v - mi > 0)}, v = beta_g (normal model) or N_1kG * beta_g (gamma model) for
                     * the per-gene rows: the per-gene posterior table of gene_regression.py:165-174 without the trace.
                     * With k6 set, `trace` may be NULL (no positions stored). */
    double* head_trace; /* optional [draws][C][H] fp64, H = rows in front of the per-gene rows (3 normal, 2 gamma / genome-wide):
                         * the shared parameters of every recorded draw in full precision when `trace` is fp32 or absent */
} bgh_nuts_async_buffers;

int bgh_nuts_run(bgh_ctx* ctx, const bgh_nuts_async_buffers* buf, const bgh_nuts_cfg* cfg, int32_t tune,
                 int32_t draws, int32_t poll_ticks, void* stream, int64_t* n_ticks);

/* Host-only: the sampler kernel's work plan with bgh_upload's defaults (gid_sorted padded to the batch of 8 as on the
 * device); two_level = 0 cuts the warps of a CTA with the CTA-level boundary cost as well (round-2 behaviour, for tests). */
int bgh_debug_sampler_plan(int64_t M, const int32_t* gid_sorted, int32_t G, int32_t n_cta, int32_t groups_per_cta, int32_t two_level,
                           int32_t* offsets, int32_t* flags);

/* Tools only: per-warp phase timestamps (globaltimer ns) of the likelihood kernel.  enable != 0
 * allocates the trace buffer; out (optional) receives the last launch's
 * [chain_blocks * n_cta * 16][8] u64 words {t_enter, t_setup_done, t_stream_done, t_barrier,
 * t_exit, -, -, smid}. */
int bgh_debug_trace(bgh_ctx* ctx, int32_t enable, uint64_t* out, int64_t n_words);

/* FP64 pipe microbenchmark (dependent-free DFMA streams on every SM): returns achieved
 * G DFMA-lane-instructions / s, the compute roof of the likelihood kernel at large C. */
int bgh_dfma_peak(bgh_ctx* ctx, double* ginstr_per_s);

/* ------------------------------------------------------------------------------------------
 * Native reader of the annotated SNP table (host only; replaces the pd.read_csv of snps.py:32-64 for
 * the schema of docs/createdata.rst:45-47: chr, rs_id, position, cm, maf, l, gene, sample_size, z).
 * The file is memory-mapped and parsed by n_threads threads (0 = all cores) into typed columns:
 * kind 0 = float64 (NaN for pandas' NA spellings), kind 1 = dictionary-encoded labels ("gene", "chr";
 * code -1 = missing; the dictionary is SORTED, so a gene code is the `cat` index of
 * gene_regression.py:61-71), kind 2 = skipped ("rs_id").  sep is the separator character.
 * Errors: BGH_EINVAL (unreadable file, ragged row, unparsable number) with bgh_ingest_last_error(). */
typedef struct bgh_table bgh_table;
int bgh_csv_read(const char* path, int32_t sep, int32_t n_threads, bgh_table** out);
const char* bgh_ingest_last_error(void);
int64_t bgh_table_rows(const bgh_table* t);
int32_t bgh_table_n_columns(const bgh_table* t);
const char* bgh_table_column_name(const bgh_table* t, int32_t column);
int32_t bgh_table_column_kind(const bgh_table* t, int32_t column);
const double* bgh_table_f64(const bgh_table* t, int32_t column);
const int32_t* bgh_table_codes(const bgh_table* t, int32_t column);
int32_t bgh_table_n_labels(const bgh_table* t, int32_t column);
const char* bgh_table_label(const bgh_table* t, int32_t column, int32_t i);
void bgh_table_free(bgh_table* t);

#ifdef __cplusplus
}
#endif
#endif /* BAGHERA_B200_H */
