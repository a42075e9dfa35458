// Inner-loop probe: the likelihood fast path fed from shared memory, in several layouts.
//   V=0: AoS {s,nl} LDS.128 + {w} LDS.64 broadcast per SNP, CPL chains per lane (current kernel)
//   V=1: SoA doubles: for 2 consecutive SNPs LDS.128 s[i..i+1], nl[i..i+1], w[i..i+1]
//   V=2: like V=0 but {s,l} as fp32 pair (LDS.64) + w double (LDS.64); converts in registers
// Reports achieved G FP64 lane-instr/s (6 per SNP x chain) for 16 warps/SM.
#include <cstdio>
#include <cuda_runtime.h>

struct __align__(16) A2 { double s, nl; };
struct __align__(16) B2 { double w; int g; int pad; };

template <int V, int CPL, int B, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(double* out, int chunks, const double* in)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    A2* smA = reinterpret_cast<A2*>(sm) + warp * 128;
    B2* smB = reinterpret_cast<B2*>(sm + WARPS * 128 * 16) + warp * 128;
    double* soa = reinterpret_cast<double*>(sm) + warp * 128 * 4;   // s[128], nl[128], w[128], (g)
    float2* smF = reinterpret_cast<float2*>(sm) + warp * 128;
    for (int i = lane; i < 128; i += 32) {
        if (V == 0) { smA[i].s = 1.0 + i; smA[i].nl = -2.0 - i; smB[i].w = 1.0 / (2.0 + i); smB[i].g = 0; }
        if (V == 1) { soa[i] = 1.0 + i; soa[128 + i] = -2.0 - i; soa[256 + i] = 1.0 / (2.0 + i); }
        if (V == 2) { smF[i] = make_float2(1.0f + i, -2.0f - i); smB[i].w = 1.0 / (2.0 + i); smB[i].g = 0; }
    }
    __syncwarp();
    double e[CPL], b[CPL], ag[CPL], ae[CPL], al[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) { e[k] = in[k] + lane; b[k] = in[8 + k]; ag[k] = ae[k] = al[k] = 0; }
    int cur = 0;
    for (int c = 0; c < chunks; ++c) {
        for (int i = 0; i < 128; i += B) {
            if (V != 1) { if (smB[i + B - 1].g != cur) { cur++; ag[0] += 1.0; } }
#pragma unroll
            for (int t = 0; t < B; ++t) {
                double s, nl, w;
                if (V == 0) { A2 a = smA[i + t]; s = a.s; nl = a.nl; w = smB[i + t].w; }
                if (V == 1) { s = soa[i + t]; nl = soa[128 + i + t]; w = soa[256 + i + t]; }
                if (V == 2) { float2 f = smF[i + t]; s = (double)f.x; nl = (double)f.y; w = smB[i + t].w; }
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const double d = s - e[k];
                    const double r = fma(nl, b[k], d);
                    ag[k] += r;
                    const double tt = r * w;
                    ae[k] += tt;
                    al[k] = fma(r, tt, al[k]);
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

template <int V, int CPL, int B, int WARPS>
void run(const char* name, double* out, const double* in)
{
    const int chunks = 64, sms = 148;
    const size_t smem = WARPS * 128 * 32;
    cudaFuncSetAttribute(probe<V, CPL, B, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<V, CPL, B, WARPS><<<sms, WARPS * 32, smem>>>(out, chunks, in);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<V, CPL, B, WARPS><<<sms, WARPS * 32, smem>>>(out, chunks, in);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double lanes = (double)sms * WARPS * 32 * chunks * 128.0 * CPL * 6.0;
    printf("%-28s warps/SM %2d CPL %d B %d : %.0f G lane-instr/s  (%s)\n", name, WARPS, CPL, B, lanes / ms * 1e-6,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double* out; cudaMalloc(&out, 8);
    double* in; cudaMalloc(&in, 8 * 128); cudaMemset(in, 0, 8 * 128);
    run<0, 2, 8, 16>("AoS LDS.128+LDS.64", out, in);
    run<0, 2, 8, 8>("AoS LDS.128+LDS.64", out, in);
    run<0, 2, 4, 16>("AoS LDS.128+LDS.64", out, in);
    run<0, 1, 8, 16>("AoS LDS.128+LDS.64", out, in);
    run<0, 4, 8, 16>("AoS LDS.128+LDS.64", out, in);
    run<0, 4, 4, 16>("AoS LDS.128+LDS.64", out, in);
    run<0, 4, 8, 12>("AoS LDS.128+LDS.64", out, in);
    run<1, 2, 8, 16>("SoA 3x LDS.128 / 2 SNPs", out, in);
    run<1, 2, 8, 8>("SoA 3x LDS.128 / 2 SNPs", out, in);
    run<1, 4, 8, 16>("SoA 3x LDS.128 / 2 SNPs", out, in);
    run<2, 2, 8, 16>("fp32 pair + w", out, in);
    run<2, 4, 8, 16>("fp32 pair + w", out, in);
    return 0;
}
