// Device side of the one-shot peer-memory all-reduce (see bgh_peer.h).  The message is a few KB, i.e. latency
// bound: instead of a separate NCCL launch, one CTA stores its contribution straight into every peer's mailbox
// over NVLink, releases a system-scope flag, waits for the peers' flags and adds the R contributions in RANK
// ORDER, so every rank obtains bit-identical totals (hence identical NUTS decisions).  Mailboxes are double
// buffered by launch parity: a rank can only be one launch ahead of its slowest peer, because it needs that
// peer's flag of launch e to finish launch e.
#pragma once

#include <cuda_runtime.h>

#include "bgh_peer.h"

namespace bgh {

constexpr unsigned long long kPeerTimeoutNs = 2000000000ull;

__device__ __forceinline__ unsigned long long peer_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int peer_ld_volatile_i32(const int* p)
{
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
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

// executed by one whole CTA; on return payload[0..n) holds the rank-ordered total on every rank.
// A peer that does not answer within kPeerTimeoutNs raises *status (2) instead of hanging the GPU.
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
        const unsigned long long t0 = peer_timer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys_u64(mine) < pr.epoch) {
            if ((++spins & 255u) == 0 && (peer_timer_ns() - t0 > kPeerTimeoutNs || peer_ld_volatile_i32(status) != 0)) {
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

}  // namespace bgh
