/*
 * oracle_c.c — plain-C, fp64, scalar CPU restatement of the reference's logp+grad path.
 *
 * TEST INFRASTRUCTURE / REPORTED CPU BASELINE ONLY.  Never linked into the product library
 * (baghera_b200/csrc); loaded only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference evaluates this density through pymc3==3.6 + theano
 * (ref: requirements.txt:7), absent from /root/reference and not installable here; there are
 * no reference tests or golden vectors for the path.  This file follows the op sequence that
 * Theano generates from ref: baghera_tool/gene_regression.py:77-98:
 *
 *   forward : eta = beta[cat]                (AdvancedSubtensor1, gather through UNSORTED cat)
 *             mu  = c*eta*l + e              (Elemwise{Composite}, M-long fp64 temporary)
 *             ll  = sum_j (-tau_j (s_j-mu_j)^2 + log(tau_j/2pi))/2     (Sum{acc_dtype=float64})
 *   backward: d ll/d mu_j elementwise (M-long temporary) -> sum (d/de)
 *             d/d beta = scatter-add over cat  (AdvancedIncSubtensor1, serial loop)
 *   priors + Jacobians: O(G)                 (ref: gene_regression.py:78-81)
 *
 * One chain per call, single threaded (Theano's openmp flag defaults off);
 * bgh_oracle_logp_grad_batch runs one chain per pthread the way pm.sample(cores=N)
 * runs one chain per process (ref: gene_regression.py:100-106).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LOG_2PI 1.8378770664093454835606594728112

typedef struct {
    int64_t M, G;
    const double *chi2, *ld;   /* fp64 copies of the fp32-rounded inputs */
    const int64_t *cat;        /* file order, unsorted (Q6) */
    double c;                  /* n_patients / N_1kG  (ref: gene_regression.py:95) */
    double lik_const;          /* sum_j -1/2 log(2 pi l_j) */
    int fix_intercept;
} oracle_data;

double bgh_oracle_lik_const(int64_t M, const double *ld)
{
    double acc = 0.0;
    for (int64_t j = 0; j < M; ++j) acc += log(ld[j]);
    return -0.5 * acc - 0.5 * (double)M * LOG_2PI;
}

/* work: 2*M doubles of scratch (the M-long temporaries Theano materialises). */
int bgh_oracle_logp_grad(int64_t M, int64_t G, const double *chi2, const double *ld,
                         const int64_t *cat, double c, double lik_const, int fix_intercept,
                         const double *q, double *logp, double *grad, double *work)
{
    const double xW = q[0], e = q[1], xm = q[2];
    const double *beta = q + 3;
    const double W = exp(xW);
    const double mi = 1.0 / (1.0 + exp(-xm));
    const double icpt = fix_intercept ? 1.0 : e;
    double *mu = work, *dmu = work + M;

    /* pass 1: gather + elementwise mean */
    for (int64_t j = 0; j < M; ++j) mu[j] = c * beta[cat[j]] * ld[j] + icpt;
    /* pass 2: residual / tau, log-likelihood sum */
    double ll = 0.0;
    for (int64_t j = 0; j < M; ++j) {
        const double r = chi2[j] - mu[j];
        const double t = r / ld[j];
        dmu[j] = t;                 /* d ll / d mu_j = tau_j (s_j - mu_j) */
        ll += r * t;
    }
    ll = -0.5 * ll + lik_const;
    /* pass 3: d/de = sum_j dmu_j */
    double de = 0.0;
    for (int64_t j = 0; j < M; ++j) de += dmu[j];
    /* pass 4: scatter-add d/d beta_g += c * l_j * dmu_j */
    for (int64_t g = 0; g < G; ++g) grad[3 + g] = 0.0;
    for (int64_t j = 0; j < M; ++j) grad[3 + cat[j]] += c * ld[j] * dmu[j];

    /* priors + Jacobians */
    double S1 = 0.0, S2 = 0.0;
    const double iW2 = 1.0 / (W * W);
    for (int64_t g = 0; g < G; ++g) {
        const double d = beta[g] - mi;
        S1 += d;
        S2 += d * d;
        grad[3 + g] -= d * iW2;
    }
    double lp = (-1.0 / W - 2.0 * log(W)) + xW;
    lp += -0.5 * (e - 1.0) * (e - 1.0) - 0.5 * LOG_2PI;
    lp += log(mi) + log1p(-mi);
    lp += -0.5 * S2 * iW2 - (double)G * log(W) - 0.5 * (double)G * LOG_2PI;
    lp += ll;
    grad[0] = 1.0 / W - 1.0 + S2 * iW2 - (double)G;
    grad[1] = -(e - 1.0) + (fix_intercept ? 0.0 : de);
    grad[2] = 1.0 - 2.0 * mi + mi * (1.0 - mi) * S1 * iW2;
    *logp = lp;
    return 0;
}

/* Genome-wide model, ref: baghera_tool/heritability.py:44-55 ; q = [e, mi_logodds__]. */
int bgh_oracle_gw_logp_grad(int64_t M, const double *chi2, const double *ld, double c,
                            double lik_const, const double *q, double *logp, double *grad,
                            double *work)
{
    const double e = q[0], xm = q[1];
    const double mi = 1.0 / (1.0 + exp(-xm));
    double *mu = work, *dmu = work + M;
    for (int64_t j = 0; j < M; ++j) mu[j] = c * mi * ld[j] + e;
    double ll = 0.0;
    for (int64_t j = 0; j < M; ++j) {
        const double r = chi2[j] - mu[j];
        const double t = r / ld[j];
        dmu[j] = t;
        ll += r * t;
    }
    double de = 0.0, dmi = 0.0;
    for (int64_t j = 0; j < M; ++j) de += dmu[j];
    for (int64_t j = 0; j < M; ++j) dmi += c * ld[j] * dmu[j];
    double lp = -0.5 * (e - 1.0) * (e - 1.0) - 0.5 * LOG_2PI;
    lp += log1p(-mi) + log(2.0);
    lp += log(mi) + log1p(-mi);
    lp += -0.5 * ll + lik_const;
    grad[0] = -(e - 1.0) + de;
    grad[1] = (dmi - 1.0 / (1.0 - mi)) * mi * (1.0 - mi) + (1.0 - 2.0 * mi);
    *logp = lp;
    return 0;
}

/* Q: [C][D] chain-major; one chain per worker thread, the way pm.sample(cores=N) runs one
 * chain per process.  n_threads <= 0 -> number of online cores. */
#include <pthread.h>
#include <unistd.h>

typedef struct {
    int64_t M, G, C;
    const double *chi2, *ld;
    const int64_t *cat;
    double c, lik_const;
    int fix_intercept;
    const double *Q;
    double *logp, *grad;
    int64_t *next;            /* shared work counter */
    pthread_mutex_t *mu;
    int rc;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *job = (batch_job *)arg;
    const int64_t D = job->G + 3;
    double *work = (double *)malloc(sizeof(double) * 2 * (size_t)job->M);
    if (!work) { job->rc = -1; return NULL; }
    for (;;) {
        pthread_mutex_lock(job->mu);
        const int64_t ch = (*job->next)++;
        pthread_mutex_unlock(job->mu);
        if (ch >= job->C) break;
        bgh_oracle_logp_grad(job->M, job->G, job->chi2, job->ld, job->cat, job->c,
                             job->lik_const, job->fix_intercept, job->Q + ch * D,
                             job->logp + ch, job->grad + ch * D, work);
    }
    free(work);
    return NULL;
}

int bgh_oracle_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int bgh_oracle_logp_grad_batch(int64_t M, int64_t G, const double *chi2, const double *ld,
                               const int64_t *cat, double c, double lik_const,
                               int fix_intercept, int64_t C, const double *Q, double *logp,
                               double *grad, int n_threads)
{
    if (n_threads <= 0) n_threads = bgh_oracle_num_threads();
    if (n_threads > C) n_threads = (int)C;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t tid[1024];
    batch_job jobs[1024];
    int64_t next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    for (int t = 0; t < n_threads; ++t) {
        batch_job j = {M, G, C, chi2, ld, cat, c, lik_const, fix_intercept, Q, logp, grad,
                       &next, &mu, 0};
        jobs[t] = j;
    }
    int started = 0;
    for (int t = 0; t < n_threads; ++t) {
        if (pthread_create(&tid[t], NULL, batch_worker, &jobs[t]) != 0) break;
        ++started;
    }
    if (started == 0) { batch_worker(&jobs[0]); return jobs[0].rc; }
    int rc = 0;
    for (int t = 0; t < started; ++t) { pthread_join(tid[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
    return rc;
}
