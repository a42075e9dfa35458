// FP64 pipe probe at different occupancies / ILP: how many warps per SMSP saturate the DP pipe?
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double m, double c)
{
    double a[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) a[j] = threadIdx.x * 1e-3 + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) a[j] = fma(a[j], m, c);
    }
    double r = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) r += a[j];
    if (r == 123.456) out[0] = r;
}
// 3-register-operand mix like the likelihood inner loop: d = s - e; r = fma(nl, b, d); ag += r; t = r*w; ae += t; al = fma(r,t,al)
template <int NS>
__global__ void mix(double* out, int iters, const double* in)
{
    double s[NS], nl[NS], w[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) { s[j] = in[j]; nl[j] = in[NS + j]; w[j] = in[2 * NS + j]; }
    double e0 = in[100] + threadIdx.x, e1 = in[101], b0 = in[102], b1 = in[103];
    double ag0 = 0, ag1 = 0, ae0 = 0, ae1 = 0, al0 = 0, al1 = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            double d0 = s[j] - e0, d1 = s[j] - e1;
            double r0 = fma(nl[j], b0, d0), r1 = fma(nl[j], b1, d1);
            ag0 += r0; ag1 += r1;
            double t0 = r0 * w[j], t1 = r1 * w[j];
            ae0 += t0; ae1 += t1;
            al0 = fma(r0, t0, al0); al1 = fma(r1, t1, al1);
        }
        e0 += 1e-9; e1 -= 1e-9;
    }
    double r = ag0 + ag1 + ae0 + ae1 + al0 + al1;
    if (r == 123.456) out[0] = r;
}
template <typename F>
float timeit(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main()
{
    double* out; cudaMalloc(&out, 8);
    double* in; cudaMalloc(&in, 8 * 128); cudaMemset(in, 0, 8 * 128);
    int sms = 148, iters = 1 << 14;
    for (int warps : {4, 8, 16, 32, 64}) {
        int threads = warps * 32 > 1024 ? 1024 : warps * 32, ctas = sms * (warps * 32 / threads);
        float m1 = timeit([&] { k<1><<<ctas, threads>>>(out, iters, 0.999999, 1e-9); });
        float m2 = timeit([&] { k<2><<<ctas, threads>>>(out, iters, 0.999999, 1e-9); });
        float m4 = timeit([&] { k<4><<<ctas, threads>>>(out, iters, 0.999999, 1e-9); });
        float m8 = timeit([&] { k<8><<<ctas, threads>>>(out, iters, 0.999999, 1e-9); });
        double lanes = (double)ctas * threads * iters;
        printf("warps/SM %2d  G lane-instr/s: ILP1 %.0f ILP2 %.0f ILP4 %.0f ILP8 %.0f\n", warps,
               lanes * 1 / m1 * 1e-6, lanes * 2 / m2 * 1e-6, lanes * 4 / m4 * 1e-6, lanes * 8 / m8 * 1e-6);
    }
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32 > 1024 ? 1024 : warps * 32, ctas = sms * (warps * 32 / threads);
        int it = 1 << 11;
        float m4 = timeit([&] { mix<4><<<ctas, threads>>>(out, it, in); });
        float m8 = timeit([&] { mix<8><<<ctas, threads>>>(out, it, in); });
        double lanes = (double)ctas * threads * it * 12.0;
        printf("mix warps/SM %2d  G lane-instr/s: NS4 %.0f NS8 %.0f\n", warps, lanes * 4 / m4 * 1e-6, lanes * 8 / m8 * 1e-6);
    }
    return 0;
}
