// bgh_fin_device.cuh — per-chain closing formulas (priors, transforms, Jacobians) shared by the
// stand-alone finalize / finish kernels (bgh_lik.cu) and the fused kernels (bgh_fused.cu).
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "bgh_internal.h"
#include "bgh_lik_device.cuh"

namespace bgh {

// ---------------------------------------------------------------------------------------------
// per-chain closing formulas (priors, transforms, Jacobians) — ref: gene_regression.py:78-81,
// heritability.py:45-49; PyMC3 3.6 distributions/continuous.py + transforms.py
// ---------------------------------------------------------------------------------------------
// digamma for x > 0: recurrence up to x >= 10, then the asymptotic series (|error| < 1e-15 relative)
__device__ __forceinline__ double digamma_d(double x)
{
    double r = 0.0;
    while (x < 10.0) {
        r -= 1.0 / x;
        x += 1.0;
    }
    const double f = 1.0 / (x * x);
    const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 +
                     f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return r + log(x) - 0.5 / x + t;
}

constexpr double kGammaETau = 1.0e6;   // e ~ Normal(1, sd=0.001), ref: gene_regression.py:229

struct ChainSums {
    double A_ll, A_e, S1, S2;
};
// Sum-independent per-chain terms of the closing formulas: computed once per evaluation, in
// parallel with the partial-sum reduction (finalize) or directly (finish kernel).
//   gene model: {iW, iW2, mi, lp0}   lp0 = all logp terms that do not involve the four sums
//   gw model  : {sg0 (sigmoid of x0 when the intercept is "fixed"), -, mi, lp0}
struct ChainConst {
    double iW, iW2, mi, lp0;
    double x0, x1;      // rows 0 and 1 of q as read for the constants (the closing formulas need e / x0 again)
};

__device__ __forceinline__ ChainConst chain_const(const FinParams& p, int ch)
{
    const int Cp = p.Cp;
    ChainConst c;
    if (p.model == BGH_MODEL_GENE_GAMMA) {
        // ref: gene_regression.py:229-231.  e ~ Normal(1, 0.001); mi ~ Beta(1,1) (logodds);
        // beta_g ~ Gamma(alpha = mi, beta = rate) sampled as beta_log__:
        //   sum_g [ -lgamma(mi) + mi log(rate) - rate beta_g + (mi - 1) x_g + x_g ]   (x_g = log beta_g, Jacobian x_g)
        const double e = p.q[0 * p.ds + ch * p.cs], xm = p.q[1 * p.ds + ch * p.cs];
        c.x0 = e; c.x1 = xm;
        const double ax = fabs(xm);
        const double u = exp(-ax);
        const double L = log1p(u);
        c.mi = (xm >= 0.0 ? 1.0 : u) / (1.0 + u);
        const double G = (double)p.G_total;
        const double lrate = log(p.gamma_rate);
        double lp = -0.5 * kGammaETau * (e - 1.0) * (e - 1.0) + 0.5 * (log(kGammaETau) - kLog2Pi);
        lp += -ax - 2.0 * L;                                       // Beta(1,1)=0 ; logodds Jacobian
        lp += G * (-lgamma(c.mi) + c.mi * lrate);
        lp += p.lik_const;
        c.lp0 = lp;
        c.iW = G * (lrate - digamma_d(c.mi));                      // d/d mi of the gene-independent prior terms
        c.iW2 = 0.0;
        return c;
    }
    if (p.model == BGH_MODEL_GENE_NORMAL) {
        const double xW = p.q[0 * p.ds + ch * p.cs], e = p.q[1 * p.ds + ch * p.cs], xm = p.q[2 * p.ds + ch * p.cs];
        c.x0 = xW; c.x1 = e;
        const double ax = fabs(xm);
        const double u = exp(-ax);                 // one exp + one log1p serve sigmoid and both logs
        const double L = log1p(u);
        c.iW = exp(-xW);
        c.iW2 = c.iW * c.iW;
        c.mi = (xm >= 0.0 ? 1.0 : u) / (1.0 + u);
        const double G = (double)p.G_total;
        // log sigmoid(x) + log(1 - sigmoid(x)) = -|x| - 2 log1p(exp(-|x|))
        double lp = (-c.iW - 2.0 * xW) + xW;                       // InverseGamma(1,1) + log-Jacobian
        lp += -0.5 * (e - 1.0) * (e - 1.0) - 0.5 * kLog2Pi;        // Normal(1,1)
        lp += -ax - 2.0 * L;                                       // Beta(1,1)=0 ; logodds Jacobian
        lp += -G * xW - 0.5 * G * kLog2Pi;                         // Normal(beta | mi, W) constants
        lp += p.lik_const;
        c.lp0 = lp;
    } else {
        const double x0 = p.q[0 * p.ds + ch * p.cs], xm = p.q[1 * p.ds + ch * p.cs];
        c.x0 = x0; c.x1 = xm;
        const double ax = fabs(xm);
        const double u = exp(-ax);
        const double L = log1p(u);
        c.mi = (xm >= 0.0 ? 1.0 : u) / (1.0 + u);
        const double l1m = (xm >= 0.0 ? -xm : 0.0) - L;            // log(1 - mi)
        double lp = l1m + 0.69314718055994530942;                  // Beta(1,2)
        lp += -ax - 2.0 * L;                                       // logodds Jacobian
        c.iW = 0.0;
        if (p.fix) {
            // Uniform(lo,hi) logp = -log(hi-lo); interval Jacobian = log(hi-lo) + log sigmoid'(x0)
            const double a0 = fabs(x0), u0 = exp(-a0);
            c.iW = (x0 >= 0.0 ? 1.0 : u0) / (1.0 + u0);            // sigmoid(x0)
            lp += -a0 - 2.0 * log1p(u0);
        } else {
            lp += -0.5 * (x0 - 1.0) * (x0 - 1.0) - 0.5 * kLog2Pi;
        }
        lp += p.lik_const;
        c.iW2 = 0.0;
        c.lp0 = lp;
    }
    return c;
}

// logp and the gradients of the head rows of one chain (g[0..2]; the genome-wide model's g[1] comes from finish_gene)
__device__ __forceinline__ void chain_closing(const FinParams& p, int ch, const ChainSums& s, const ChainConst& c, double& logp, double (&g)[3])
{
    g[0] = g[1] = g[2] = 0.0;
    if (p.model == BGH_MODEL_GENE_GAMMA) {
        // S1 = sum_g x_g, S2 = sum_g beta_g (owned genes)
        const double e = c.x0;
        logp = c.lp0 + c.mi * s.S1 - p.gamma_rate * s.S2 - 0.5 * s.A_ll;
        g[0] = -kGammaETau * (e - 1.0) + (p.fix ? 0.0 : s.A_e);
        g[1] = 1.0 - 2.0 * c.mi + c.mi * (1.0 - c.mi) * (c.iW + s.S1);
        return;
    }
    if (p.model == BGH_MODEL_GENE_NORMAL) {
        const double e = c.x1;
        const double G = (double)p.G_total;
        logp = c.lp0 - 0.5 * s.S2 * c.iW2 - 0.5 * s.A_ll;
        g[0] = c.iW - 1.0 + s.S2 * c.iW2 - G;
        g[1] = -(e - 1.0) + (p.fix ? 0.0 : s.A_e);
        g[2] = 1.0 - 2.0 * c.mi + c.mi * (1.0 - c.mi) * s.S1 * c.iW2;
    } else {
        const double x0 = c.x0;
        logp = c.lp0 - 0.5 * s.A_ll;
        if (p.fix) {
            const double sg = c.iW;
            g[0] = s.A_e * (kGwFixHi - kGwFixLo) * sg * (1.0 - sg) + (1.0 - 2.0 * sg);
        } else {
            g[0] = -(x0 - 1.0) + s.A_e;
        }
    }
}

__device__ __forceinline__ void finish_chain(const FinParams& p, int ch, const ChainSums& s, const ChainConst& c)
{
    double lp, g[3];
    chain_closing(p, ch, s, c, lp, g);
    p.logp[ch] = lp;
    p.grad[0 * p.ds + ch * p.cs] = g[0];
    if (p.model == BGH_MODEL_GENE_GAMMA) p.grad[1 * p.ds + ch * p.cs] = g[1];
    if (p.model == BGH_MODEL_GENE_NORMAL) {
        p.grad[1 * p.ds + ch * p.cs] = g[1];
        p.grad[2 * p.ds + ch * p.cs] = g[2];
    }
}

// gradient row of a gene from its total residual sum.  The per-chain constants (1/W^2, mi) were
// computed by the likelihood kernel and left in p.cc, so no exp() chain sits on this path.
struct GeneCoef {
    double iW2, mi, beta;
};
__device__ __forceinline__ GeneCoef load_gene_coef(const FinParams& p, int gene, int ch, bool needed)
{
    GeneCoef gc;
    gc.iW2 = gc.mi = gc.beta = 0.0;
    if (needed) {
        gc.iW2 = p.cc[0 * p.Cp + ch];
        gc.mi = p.cc[1 * p.Cp + ch];
        if (p.model == BGH_MODEL_GENE_NORMAL) gc.beta = p.q[(long long)(3 + gene) * p.ds + ch * p.cs];
        if (p.model == BGH_MODEL_GENE_GAMMA) gc.beta = exp(p.q[(long long)(2 + gene) * p.ds + ch * p.cs]);
    }
    return gc;
}
// gradient of a gene's row (genome-wide model: of the slope row 1) from its total residual sum
__device__ __forceinline__ double gene_grad(const FinParams& p, double sum_r, const GeneCoef& gc)
{
    if (p.model == BGH_MODEL_GENE_GAMMA) return fma(gc.beta, fma(p.c, sum_r, -p.gamma_rate), gc.mi);
    if (p.model == BGH_MODEL_GENE_NORMAL) return fma(p.c, sum_r, -(gc.beta - gc.mi) * gc.iW2);
    return p.c * sum_r * gc.mi * (1.0 - gc.mi) + 1.0 - 3.0 * gc.mi;      // (c*sum_r - 1/(1-mi)) * mi(1-mi) + (1 - 2mi)
}
__device__ __forceinline__ void finish_gene(const FinParams& p, int gene, int ch, double sum_r, const GeneCoef& gc)
{
    const double g = gene_grad(p, sum_r, gc);
    if (p.model == BGH_MODEL_GENE_GAMMA) p.grad[(long long)(2 + gene) * p.ds + ch * p.cs] = g;
    else if (p.model == BGH_MODEL_GENE_NORMAL) p.grad[(long long)(3 + gene) * p.ds + ch * p.cs] = g;
    else p.grad[1 * p.ds + ch * p.cs] = g;
}

}  // namespace bgh
