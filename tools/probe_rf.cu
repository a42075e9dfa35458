// Register-file read bandwidth probe for the FP64 pipe (sm_100a): DFMA throughput as a function of the number of
// source operands that have to be READ per instruction (operands the previous instruction of the warp used in the
// same slot come from the operand reuse cache).  Build: nvcc -O3 -arch=sm_100a -o probe_rf probe_rf.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 8;
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters, const double* in)
{
    double a[N], x[N], y[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { a[j] = in[j] + threadIdx.x; x[j] = in[N + j]; y[j] = in[2 * N + j]; }
    const double m = in[100], c = in[101];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (MODE == 1) a[j] = fma(a[j], m, c);            // 1 read  (m, c reused)
            if (MODE == 2) a[j] = fma(x[j], m, a[j]);         // 2 reads (m reused)
            if (MODE == 3) a[j] = fma(x[j], y[j], a[j]);      // 3 reads
            if (MODE == 4) a[j] = fma(x[j], x[j], a[j]);      // the same register in two slots
            if (MODE == 5) a[j] = fma(x[j], y[(j + 1) & (N - 1)], a[j]);   // 3 reads, another pairing
        }
    }
    double r = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) r += a[j];
    if (r == 123.456) out[0] = r;
}
template <int MODE>
void run(double* out, const double* in, int warps)
{
    const int iters = 1 << 14, sms = 148;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<sms, warps * 32>>>(out, iters, in);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<sms, warps * 32>>>(out, iters, in);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("mode %d warps/SM %2d: %.0f G lane-instr/s (%s)\n", MODE, warps, (double)sms * warps * 32 * iters * N / ms * 1e-6,
           cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    double* out; cudaMalloc(&out, 8);
    double* in; cudaMalloc(&in, 8 * 128); cudaMemset(in, 0, 8 * 128);
    for (int w : {8, 16}) { run<1>(out, in, w); run<2>(out, in, w); run<3>(out, in, w); run<4>(out, in, w); run<5>(out, in, w); }
    return 0;
}
