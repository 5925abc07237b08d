// Do NVLink peer reads overlap local HBM traffic inside one kernel?  (single process, 2 GPUs)
// A: local copy alone, B: peer read alone, C: both in one kernel with R dedicated "remote" CTAs,
// D: both in every thread (3 local + 1 remote 16-byte load per iteration), E: two kernels, two streams.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__global__ void k_copy(const float4 *src, float4 *dst, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) dst[i] = src[i];
}
__global__ void k_split(const float4 *loc, const float4 *peer, float4 *dst_l, float4 *dst_p, size_t nl, size_t np, int R) {
    if ((int)blockIdx.x < R) {
        size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)R * blockDim.x;
        for (; i < np; i += s) dst_p[i] = peer[i];
    } else {
        size_t i = (blockIdx.x - R) * (size_t)blockDim.x + threadIdx.x, s = (size_t)(gridDim.x - R) * blockDim.x;
        for (; i < nl; i += s) dst_l[i] = loc[i];
    }
}
__global__ void k_inter(const float4 *loc, const float4 *peer, float4 *dst_l, float4 *dst_p, size_t nl, size_t np) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    size_t ratio = nl / np;     // local loads per remote load
    for (size_t j = t; j < np; j += s) {
        float4 r = peer[j];
        for (size_t k = 0; k < ratio; k++) dst_l[j * ratio + k] = loc[j * ratio + k];
        dst_p[j] = r;
    }
}
template <typename F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t lb = (size_t)2 << 30, pb = (size_t)512 << 20, nl = lb / 16, np = pb / 16;
    float4 *L0, *L0b, *D0, *P1;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&P1, pb)); CK(cudaMemset(P1, 0, pb));
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&L0, lb)); CK(cudaMalloc(&L0b, lb)); CK(cudaMalloc(&D0, pb)); CK(cudaMemset(L0, 0, lb));
    float a = timeit([&] { k_copy<<<148 * 8, 256>>>(L0, L0b, nl); });
    float b = timeit([&] { k_copy<<<148 * 8, 256>>>(P1, D0, np); });
    printf("A local copy 2 GiB: %.3f ms (%.0f GB/s r+w)   B peer read 0.5 GiB: %.3f ms (%.0f GB/s)   sum %.3f  max %.3f\n", a, 2 * lb / a / 1e6, b, pb / b / 1e6, a + b, a > b ? a : b);
    for (int R : {8, 16, 32, 64, 148, 296, 592})
        printf("C one kernel, %4d remote CTAs of %d: %.3f ms\n", R, 148 * 8, timeit([&] { k_split<<<148 * 8, 256>>>(L0, P1, L0b, D0, nl, np, R); }));
    for (int g : {148 * 2, 148 * 8})
        printf("D interleaved in every thread, grid %d: %.3f ms\n", g, timeit([&] { k_inter<<<g, 256>>>(L0, P1, L0b, D0, nl, np); }));
    cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    for (int R : {16, 32, 64, 148}) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
            k_copy<<<R, 256, 0, s2>>>(P1, D0, np); cudaEventRecord(e2, s2);
            k_copy<<<148 * 8, 256, 0, s1>>>(L0, L0b, nl);
            cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("E two streams, peer-read kernel with %3d CTAs launched first: %.3f ms\n", R, best);
    }
    return 0;
}
