// Inner-loop probe, operand-order variants of the 5-DFMA likelihood loop (see probe_rf.cu: a DFMA that has to READ three
// source operands costs 3 issue cycles of the FP64 pipe instead of 2; operands the warp's previous instruction used in the
// same slot come from the operand reuse cache).  smem-fed SoA {sp, wp, lp}, LDS.128 per 2 SNPs, 2 chains per lane,
// 8 SNPs per batch, 16 warps per SM.  Build: nvcc -O3 -arch=sm_100a -o probe_inner6 probe_inner6.cu
#include <cstdio>
#include <cuda_runtime.h>

#define FMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))

template <int ORDER>
__device__ __forceinline__ void pair(const double2 s, const double2 w, const double2 l, const double ne0, const double ne1, const double nb0,
                                     const double nb1, double& ag0, double& ag1, double& ae0, double& ae1, double& al0, double& al1)
{
    if (ORDER == 0) {          // the shipping order: chain-major per SNP, compiler-scheduled
        for (int h = 0; h < 2; ++h) {
            const double sv = h ? s.y : s.x, wv = h ? w.y : w.x, lv = h ? l.y : l.x;
            const double r0 = fma(nb0, lv, fma(ne0, wv, sv));
            ag0 = fma(r0, lv, ag0); ae0 = fma(r0, wv, ae0); al0 = fma(r0, r0, al0);
            const double r1 = fma(nb1, lv, fma(ne1, wv, sv));
            ag1 = fma(r1, lv, ag1); ae1 = fma(r1, wv, ae1); al1 = fma(r1, r1, al1);
        }
    } else if (ORDER == 1) {   // reuse-chained order, plain C (the compiler may reorder)
        const double t0a = fma(ne0, w.x, s.x), t1a = fma(ne1, w.x, s.x), t1b = fma(ne1, w.y, s.y), t0b = fma(ne0, w.y, s.y);
        const double r0b = fma(nb0, l.y, t0b), r0a = fma(nb0, l.x, t0a), r1a = fma(nb1, l.x, t1a), r1b = fma(nb1, l.y, t1b);
        al1 = fma(r1b, r1b, al1); ae1 = fma(r1b, w.y, ae1); ag1 = fma(r1b, l.y, ag1);
        ag0 = fma(r0b, l.y, ag0); ae0 = fma(r0b, w.y, ae0); al0 = fma(r0b, r0b, al0);
        al1 = fma(r1a, r1a, al1); ae1 = fma(r1a, w.x, ae1); ag1 = fma(r1a, l.x, ag1);
        ag0 = fma(r0a, l.x, ag0); ae0 = fma(r0a, w.x, ae0); al0 = fma(r0a, r0a, al0);
    } else {                   // the same order pinned with volatile asm
        double t0a, t1a, t1b, t0b, r0a, r0b, r1a, r1b;
        FMA(t0a, ne0, w.x, s.x); FMA(t1a, ne1, w.x, s.x); FMA(t1b, ne1, w.y, s.y); FMA(t0b, ne0, w.y, s.y);
        FMA(r0b, nb0, l.y, t0b); FMA(r0a, nb0, l.x, t0a); FMA(r1a, nb1, l.x, t1a); FMA(r1b, nb1, l.y, t1b);
        FMA(al1, r1b, r1b, al1); FMA(ae1, r1b, w.y, ae1); FMA(ag1, r1b, l.y, ag1);
        FMA(ag0, r0b, l.y, ag0); FMA(ae0, r0b, w.y, ae0); FMA(al0, r0b, r0b, al0);
        FMA(al1, r1a, r1a, al1); FMA(ae1, r1a, w.x, ae1); FMA(ag1, r1a, l.x, ag1);
        FMA(ag0, r0a, l.x, ag0); FMA(ae0, r0a, w.x, ae0); FMA(al0, r0a, r0a, al0);
    }
}

template <int ORDER, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(double* out, int chunks, const double* in)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* soa = reinterpret_cast<double*>(sm) + warp * 128 * 4;   // sp[128], wp[128], lp[128], gid
    int* gid = reinterpret_cast<int*>(soa + 384);
    for (int i = lane; i < 128; i += 32) { soa[i] = 1.0 + i; soa[128 + i] = 1.0 / (2.0 + i); soa[256 + i] = 2.0 + i; gid[i] = 0; }
    __syncwarp();
    double ne0 = in[0] + lane, ne1 = in[1] + lane, nb0 = in[8], nb1 = in[9], ag0 = 0, ag1 = 0, ae0 = 0, ae1 = 0, al0 = 0, al1 = 0;
    int cur = 0;
    for (int c = 0; c < chunks; ++c) {
        for (int i = 0; i < 128; i += 8) {
            if (gid[i + 7] != cur) { cur++; ag0 += 1.0; }
            double2 s2[4], w2[4], l2[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                s2[t] = *reinterpret_cast<const double2*>(soa + i + 2 * t);
                w2[t] = *reinterpret_cast<const double2*>(soa + 128 + i + 2 * t);
                l2[t] = *reinterpret_cast<const double2*>(soa + 256 + i + 2 * t);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) pair<ORDER>(s2[t], w2[t], l2[t], ne0, ne1, nb0, nb1, ag0, ag1, ae0, ae1, al0, al1);
        }
        __syncwarp();
    }
    const double r = ag0 + ag1 + ae0 + ae1 + al0 + al1;
    if (r == 123.456) out[0] = r;
}

template <int ORDER, int WARPS>
void run(double* out, const double* in)
{
    const int chunks = 64, sms = 148;
    const size_t smem = WARPS * 128 * 32;
    cudaFuncSetAttribute(probe<ORDER, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<ORDER, WARPS><<<sms, WARPS * 32, smem>>>(out, chunks, in);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<ORDER, WARPS><<<sms, WARPS * 32, smem>>>(out, chunks, in);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double lanes = (double)sms * WARPS * 32 * chunks * 128.0 * 2 * 5.0;
    printf("order %d warps/SM %2d : %.0f G lane-instr/s  (%s)\n", ORDER, WARPS, lanes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

// ORDER 3/4: software pipeline over SNP pairs: stage A (t of pair p+2), stage B (rho of pair p+1), stage C (sums of pair p),
// each stage written in a reuse-chained operand order (3: plain C, 4: volatile asm)
template <int ORDER, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe_sp(double* out, int chunks, const double* in)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* soa = reinterpret_cast<double*>(sm) + warp * 128 * 4;
    int* gid = reinterpret_cast<int*>(soa + 384);
    for (int i = lane; i < 128; i += 32) { soa[i] = 1.0 + i; soa[128 + i] = 1.0 / (2.0 + i); soa[256 + i] = 2.0 + i; gid[i] = 0; }
    __syncwarp();
    double ne0 = in[0] + lane, ne1 = in[1] + lane, nb0 = in[8], nb1 = in[9], ag0 = 0, ag1 = 0, ae0 = 0, ae1 = 0, al0 = 0, al1 = 0;
    int cur = 0;
    for (int c = 0; c < chunks; ++c) {
        for (int i = 0; i < 128; i += 8) {
            if (gid[i + 7] != cur) { cur++; ag0 += 1.0; }
            double2 s2[4], w2[4], l2[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                s2[t] = *reinterpret_cast<const double2*>(soa + i + 2 * t);
                w2[t] = *reinterpret_cast<const double2*>(soa + 128 + i + 2 * t);
                l2[t] = *reinterpret_cast<const double2*>(soa + 256 + i + 2 * t);
            }
            double t0a[4], t1a[4], t0b[4], t1b[4], r0a[4], r1a[4], r0b[4], r1b[4];
#pragma unroll
            for (int st = 0; st < 6; ++st) {
                const int pa = st, pb = st - 1, pc = st - 2;
                if (ORDER == 3) {
                    if (pa < 4) {
                        t0a[pa] = fma(ne0, w2[pa].x, s2[pa].x); t1a[pa] = fma(ne1, w2[pa].x, s2[pa].x);
                        t1b[pa] = fma(ne1, w2[pa].y, s2[pa].y); t0b[pa] = fma(ne0, w2[pa].y, s2[pa].y);
                    }
                    if (pb >= 0 && pb < 4) {
                        r0b[pb] = fma(nb0, l2[pb].y, t0b[pb]); r0a[pb] = fma(nb0, l2[pb].x, t0a[pb]);
                        r1a[pb] = fma(nb1, l2[pb].x, t1a[pb]); r1b[pb] = fma(nb1, l2[pb].y, t1b[pb]);
                    }
                    if (pc >= 0) {
                        al1 = fma(r1b[pc], r1b[pc], al1); ae1 = fma(r1b[pc], w2[pc].y, ae1); ag1 = fma(r1b[pc], l2[pc].y, ag1);
                        ag0 = fma(r0b[pc], l2[pc].y, ag0); ae0 = fma(r0b[pc], w2[pc].y, ae0); al0 = fma(r0b[pc], r0b[pc], al0);
                        al1 = fma(r1a[pc], r1a[pc], al1); ae1 = fma(r1a[pc], w2[pc].x, ae1); ag1 = fma(r1a[pc], l2[pc].x, ag1);
                        ag0 = fma(r0a[pc], l2[pc].x, ag0); ae0 = fma(r0a[pc], w2[pc].x, ae0); al0 = fma(r0a[pc], r0a[pc], al0);
                    }
                } else {
                    if (pa < 4) {
                        FMA(t0a[pa], ne0, w2[pa].x, s2[pa].x); FMA(t1a[pa], ne1, w2[pa].x, s2[pa].x);
                        FMA(t1b[pa], ne1, w2[pa].y, s2[pa].y); FMA(t0b[pa], ne0, w2[pa].y, s2[pa].y);
                    }
                    if (pb >= 0 && pb < 4) {
                        FMA(r0b[pb], nb0, l2[pb].y, t0b[pb]); FMA(r0a[pb], nb0, l2[pb].x, t0a[pb]);
                        FMA(r1a[pb], nb1, l2[pb].x, t1a[pb]); FMA(r1b[pb], nb1, l2[pb].y, t1b[pb]);
                    }
                    if (pc >= 0) {
                        FMA(al1, r1b[pc], r1b[pc], al1); FMA(ae1, r1b[pc], w2[pc].y, ae1); FMA(ag1, r1b[pc], l2[pc].y, ag1);
                        FMA(ag0, r0b[pc], l2[pc].y, ag0); FMA(ae0, r0b[pc], w2[pc].y, ae0); FMA(al0, r0b[pc], r0b[pc], al0);
                        FMA(al1, r1a[pc], r1a[pc], al1); FMA(ae1, r1a[pc], w2[pc].x, ae1); FMA(ag1, r1a[pc], l2[pc].x, ag1);
                        FMA(ag0, r0a[pc], l2[pc].x, ag0); FMA(ae0, r0a[pc], w2[pc].x, ae0); FMA(al0, r0a[pc], r0a[pc], al0);
                    }
                }
            }
        }
        __syncwarp();
    }
    const double r = ag0 + ag1 + ae0 + ae1 + al0 + al1;
    if (r == 123.456) out[0] = r;
}
template <int ORDER, int WARPS>
void run_sp(double* out, const double* in)
{
    const int chunks = 64, sms = 148;
    const size_t smem = WARPS * 128 * 32;
    cudaFuncSetAttribute(probe_sp<ORDER, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe_sp<ORDER, WARPS><<<sms, WARPS * 32, smem>>>(out, chunks, in);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe_sp<ORDER, WARPS><<<sms, WARPS * 32, smem>>>(out, chunks, in);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double lanes = (double)sms * WARPS * 32 * chunks * 128.0 * 2 * 5.0;
    printf("order %d warps/SM %2d : %.0f G lane-instr/s  (%s)\n", ORDER, WARPS, lanes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

// ORDER 5..7: chain-outer over the 8 SNPs of a batch (ne_k / nb_k shared by consecutive instructions)
template <int ORDER, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe_co(double* out, int chunks, const double* in)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* soa = reinterpret_cast<double*>(sm) + warp * 128 * 4;
    int* gid = reinterpret_cast<int*>(soa + 384);
    for (int i = lane; i < 128; i += 32) { soa[i] = 1.0 + i; soa[128 + i] = 1.0 / (2.0 + i); soa[256 + i] = 2.0 + i; gid[i] = 0; }
    __syncwarp();
    double ne[2] = {in[0] + lane, in[1] + lane}, nb[2] = {in[8], in[9]}, ag[2] = {0, 0}, ae[2] = {0, 0}, al[2] = {0, 0};
    int cur = 0;
    for (int c = 0; c < chunks; ++c) {
        for (int i = 0; i < 128; i += 8) {
            if (gid[i + 7] != cur) { cur++; ag[0] += 1.0; }
            double s[8], w[8], l[8];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const double2 a = *reinterpret_cast<const double2*>(soa + i + 2 * t);
                const double2 b = *reinterpret_cast<const double2*>(soa + 128 + i + 2 * t);
                const double2 d = *reinterpret_cast<const double2*>(soa + 256 + i + 2 * t);
                s[2 * t] = a.x; s[2 * t + 1] = a.y; w[2 * t] = b.x; w[2 * t + 1] = b.y; l[2 * t] = d.x; l[2 * t + 1] = d.y;
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                double r[8];
                if (ORDER == 5) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = fma(w[j], ne[k], s[j]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = fma(l[j], nb[k], r[j]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { al[k] = fma(r[j], r[j], al[k]); ae[k] = fma(r[j], w[j], ae[k]); ag[k] = fma(r[j], l[j], ag[k]); }
                } else if (ORDER == 6) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) FMA(r[j], w[j], ne[k], s[j]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) FMA(r[j], l[j], nb[k], r[j]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { FMA(al[k], r[j], r[j], al[k]); FMA(ae[k], r[j], w[j], ae[k]); FMA(ag[k], r[j], l[j], ag[k]); }
                } else {
                    // 7: accumulators split in two (even / odd SNPs) so that consecutive sums are independent
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = fma(w[j], ne[k], s[j]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = fma(l[j], nb[k], r[j]);
                    double al2 = 0, ae2 = 0, ag2 = 0;
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        al[k] = fma(r[j], r[j], al[k]); ae[k] = fma(r[j], w[j], ae[k]); ag[k] = fma(r[j], l[j], ag[k]);
                        al2 = fma(r[j + 1], r[j + 1], al2); ae2 = fma(r[j + 1], w[j + 1], ae2); ag2 = fma(r[j + 1], l[j + 1], ag2);
                    }
                    al[k] += al2; ae[k] += ae2; ag[k] += ag2;
                }
            }
        }
        __syncwarp();
    }
    const double r = ag[0] + ag[1] + ae[0] + ae[1] + al[0] + al[1];
    if (r == 123.456) out[0] = r;
}
template <int ORDER, int WARPS>
void run_co(double* out, const double* in)
{
    const int chunks = 64, sms = 148;
    const size_t smem = WARPS * 128 * 32;
    cudaFuncSetAttribute(probe_co<ORDER, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe_co<ORDER, WARPS><<<sms, WARPS * 32, smem>>>(out, chunks, in);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe_co<ORDER, WARPS><<<sms, WARPS * 32, smem>>>(out, chunks, in);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double lanes = (double)sms * WARPS * 32 * chunks * 128.0 * 2 * 5.0;
    printf("order %d warps/SM %2d : %.0f G lane-instr/s  (%s)\n", ORDER, WARPS, lanes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double* out; cudaMalloc(&out, 8);
    double* in; cudaMalloc(&in, 8 * 128); cudaMemset(in, 0, 8 * 128);
    run<0, 16>(out, in); run<1, 16>(out, in); run<2, 16>(out, in);
    run_sp<3, 16>(out, in); run_sp<4, 16>(out, in);
    run_co<5, 16>(out, in); run_co<6, 16>(out, in); run_co<7, 16>(out, in);
    run<0, 8>(out, in); run<1, 8>(out, in); run<2, 8>(out, in);
    return 0;
}
