// One-shot all-reduce over peer-mapped memory (NVLink P2P stores + system-scope flags): parameters shared by
// the sharded evaluation kernel (bgh_fused.cu) and the sharded sampler's scalar kernels (bgh_sampler.cu).
#pragma once

namespace bgh {

constexpr int kMaxPeers = 8;
struct PeerParams {
    double* mail[kMaxPeers];                  // mail[r]: rank r's mailbox [2 parities][world][n] (mail[rank] is local)
    unsigned long long* flag[kMaxPeers];      // flag[r]: rank r's flags  [2 parities][world] x 16 words (128-byte lines)
    int rank, world;
    int n;                                    // doubles exchanged by this launch (identical on every rank)
    unsigned long long epoch;                 // launch counter of this mailbox, >= 1, identical on every rank
};

}  // namespace bgh
