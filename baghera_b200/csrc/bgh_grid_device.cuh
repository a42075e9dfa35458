// bgh_grid_device.cuh — grid barrier of the cooperative (one CTA per SM, all co-resident) kernels.
#pragma once

#include <cuda_runtime.h>

#include "bgh_lik_device.cuh"

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

}  // namespace bgh
