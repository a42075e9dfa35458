// Inner-loop probe for the 5-DFMA likelihood loop (bgh_lik_device.cuh): smem-fed SoA {sp, wp, lp},
// LDS.128 per 2 SNPs, CPL chains per lane, B SNPs per branch-free batch, WARPS warps per SM.
// Reports achieved G FP64 lane-instr/s (5 per SNP x chain).  Build: nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

template <int CPL, int B, int WARPS, int GCHK>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(double* out, int chunks, const double* in)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* soa = reinterpret_cast<double*>(sm) + warp * 128 * 4;   // sp[128], wp[128], lp[128], gid
    int* gid = reinterpret_cast<int*>(soa + 384);
    for (int i = lane; i < 128; i += 32) { soa[i] = 1.0 + i; soa[128 + i] = 1.0 / (2.0 + i); soa[256 + i] = 2.0 + i; gid[i] = 0; }
    __syncwarp();
    double ne[CPL], nb[CPL], ag[CPL], ae[CPL], al[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) { ne[k] = in[k] + lane; nb[k] = in[8 + k]; ag[k] = ae[k] = al[k] = 0; }
    int cur = 0;
    for (int c = 0; c < chunks; ++c) {
        for (int i = 0; i < 128; i += B) {
            if (GCHK) { if (gid[i + B - 1] != cur) { cur++; ag[0] += 1.0; } }
            double2 s2[B / 2], w2[B / 2], l2[B / 2];
#pragma unroll
            for (int t = 0; t < B / 2; ++t) {
                s2[t] = *reinterpret_cast<const double2*>(soa + i + 2 * t);
                w2[t] = *reinterpret_cast<const double2*>(soa + 128 + i + 2 * t);
                l2[t] = *reinterpret_cast<const double2*>(soa + 256 + i + 2 * t);
            }
#pragma unroll
            for (int t = 0; t < B / 2; ++t) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double s = h ? s2[t].y : s2[t].x, w = h ? w2[t].y : w2[t].x, l = h ? l2[t].y : l2[t].x;
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        const double rho = fma(nb[k], l, fma(ne[k], w, s));
                        ag[k] = fma(rho, l, ag[k]);
                        ae[k] = fma(rho, w, ae[k]);
                        al[k] = fma(rho, rho, al[k]);
                    }
                }
            }
        }
        __syncwarp();
    }
    double r = 0;
#pragma unroll
    for (int k = 0; k < CPL; ++k) r += ag[k] + ae[k] + al[k];
    if (r == 123.456) out[0] = r;
}

template <int CPL, int B, int WARPS, int GCHK>
void run(double* out, const double* in)
{
    const int chunks = 64, sms = 148;
    const size_t smem = WARPS * 128 * 32;
    cudaFuncSetAttribute(probe<CPL, B, WARPS, GCHK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<CPL, B, WARPS, GCHK><<<sms, WARPS * 32, smem>>>(out, chunks, in);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<CPL, B, WARPS, GCHK><<<sms, WARPS * 32, smem>>>(out, chunks, in);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double lanes = (double)sms * WARPS * 32 * chunks * 128.0 * CPL * 5.0;
    printf("warps/SM %2d CPL %d B %2d gchk %d : %.0f G lane-instr/s  (%s)\n", WARPS, CPL, B, GCHK, lanes / ms * 1e-6,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double* out; cudaMalloc(&out, 8);
    double* in; cudaMalloc(&in, 8 * 128); cudaMemset(in, 0, 8 * 128);
    run<2, 8, 16, 1>(out, in);
    run<2, 8, 16, 0>(out, in);
    run<2, 16, 16, 1>(out, in);
    run<2, 4, 16, 1>(out, in);
    run<1, 8, 16, 1>(out, in);
    run<1, 8, 32, 1>(out, in);
    run<4, 8, 16, 1>(out, in);
    run<4, 8, 8, 1>(out, in);
    run<4, 4, 8, 1>(out, in);
    run<2, 8, 8, 1>(out, in);
    run<2, 8, 12, 1>(out, in);
    run<2, 8, 20, 1>(out, in);
    run<2, 8, 24, 1>(out, in);
    run<3, 8, 16, 1>(out, in);
    return 0;
}
