// Internal declarations of the cooperative kernels (bgh_fused.cu) used by bgh_api.cu.
#pragma once

#include <cuda_runtime.h>

#include "bgh_internal.h"
#include "bgh_peer.h"
#include "bgh_sampler.h"

namespace bgh {

struct FusedParams {
    LikParams lik;
    FinParams fin;
    unsigned long long* bar;       // grid barrier: monotonic arrival counter (zeroed at creation)
    unsigned long long bar_base;   // value of *bar when this launch starts (tracked by the host)
    int* status;        // device flag: 0 ok, 1 = a grid barrier timed out
    int chain_blocks;   // chain blocks of CL*CPL chains, walked sequentially by every CTA
};

struct TickParams {
    FusedParams f;
    AsyncParams a;
    int n_ticks;        // ticks per launch (the kernel stops early once every chain has finished)
    int* ticks_done;    // device counter, incremented once per executed tick
    unsigned long long* phase_trace;   // tools only: [16 ticks][n_cta][8] globaltimer stamps per phase, or NULL
    int trace_first;    // first traced tick (global tick index)
    int tick_base;      // global index of this launch's first tick
};

cudaError_t fused_configure();   // opt-in dynamic shared memory, once per process
cudaError_t launch_logp_grad_fused(const FusedParams& f, int CL, int CPL, int n_cta, cudaStream_t s);
cudaError_t launch_logp_grad_sharded(const FusedParams& f, const PeerParams& pr, int CL, int CPL, int n_cta, cudaStream_t s);
cudaError_t launch_nuts_ticks(const TickParams& t, int CL, int CPL, int n_cta, cudaStream_t s);

}  // namespace bgh
