// bgh_sampler_device.cuh — device helpers of the NUTS kernels (Philox4x32-10, state accessors) shared
// by bgh_sampler.cu (lock-step transition kernels) and bgh_fused.cu (persistent tick kernel).
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "bgh_sampler.h"

namespace bgh {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}
__device__ __forceinline__ double u01_open(unsigned a, unsigned b)
{
    // 53-bit uniform in (0, 1)
    const unsigned long long x = ((unsigned long long)a << 21) ^ (unsigned long long)(b >> 11);
    return ((double)(x & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ uint2 chain_key(const NutsParams& p, int chain)
{
    const unsigned g = (unsigned)(chain + p.chain_offset);
    return make_uint2((unsigned)(p.seed & 0xffffffffull) ^ (0x9E3779B9u * (g + 1u)), (unsigned)(p.seed >> 32) + g);
}
enum : unsigned { kPurposeMomentum = 1, kPurposeDir = 2, kPurposeLeaf = 3, kPurposeTop = 4 };

// one standard normal per (chain, draw, GLOBAL coordinate): Box-Muller on a Philox block
__device__ __forceinline__ double momentum_normal(const uint2 key, long long draw, unsigned global_coord)
{
    const uint4 r = philox4x32_10(make_uint4((unsigned)draw, kPurposeMomentum, global_coord, (unsigned)(draw >> 32)), key);
    const double u1 = u01_open(r.x, r.y), u2 = u01_open(r.z, r.w);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// quadpotential._WeightedVariance.add_sample with the weight AFTER adding the sample given as its
// reciprocal (one division per chain and transition instead of three per coordinate)
__device__ __forceinline__ void welford_add(double x, double inv_w, double& mean, double& raw)
{
    const double od = x - mean;
    mean = fma(od, inv_w, mean);
    raw = fma(od, x - mean, raw);
}

__device__ __forceinline__ double chain_uniform(const NutsParams& p, int chain, unsigned purpose, unsigned idx)
{
    const uint4 r = philox4x32_10(make_uint4((unsigned)p.draw, purpose, idx, (unsigned)(p.draw >> 32)), chain_key(p, chain));
    return u01_open(r.x, r.y);
}

__device__ __forceinline__ double logaddexp_d(double a, double b)
{
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    const double m = fmax(a, b);
    return m + log1p(exp(-fabs(a - b)));
}

// vector slot accessor: state is [n_slots][D][Cp]
__device__ __forceinline__ double* slot(const NutsParams& p, int s) { return p.state + (size_t)s * p.D * p.Cp; }
// checkpoints: chain-major [Cp][kNutsMaxDepth][D] pairs {p, running p_sum}.  One (chain, level) is a
// contiguous stream over the coordinates, so the tick kernel — which transposes a tile of rows through
// shared memory — reads and writes checkpoints fully coalesced even though every chain sits at its own
// tree position (its own level).
__device__ __forceinline__ double2* ckpt2(const NutsParams& p, int level, int d, int c)
{
    return reinterpret_cast<double2*>(p.ckpt) + ((size_t)c * kNutsMaxDepth + level) * p.D + d;
}
__device__ __forceinline__ double& sc(const NutsParams& p, int k, int c) { return p.sc[(size_t)k * p.Cp + c]; }
__device__ __forceinline__ int& fl(const NutsParams& p, int k, int c) { return p.fl[(size_t)k * p.Cp + c]; }

}  // namespace bgh
