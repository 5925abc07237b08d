// NVLink peer-memory microbenchmark (single process, 2 GPUs): kernel loads/stores on a peer mapping.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__global__ void k_copy(const float4 *src, float4 *dst, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) dst[i] = src[i];
}
// tile store: each warp writes `rows` rows x 512 B at row stride ld4 (float4 units), like couple/expand
__global__ void k_tile(const float4 *src, float4 *dst, size_t ld4, int rows_per_warp, size_t nrows, int seg4) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    size_t row0 = (size_t)warp * rows_per_warp;
    if (row0 >= nrows) return;
    size_t col = (size_t)blockIdx.y * seg4;
    for (int r = 0; r < rows_per_warp; r++)
        for (int c = lane; c < seg4; c += 32) dst[(row0 + r) * ld4 + col + c] = src[(row0 + r) * ld4 + col + c];
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
    size_t bytes = (size_t)2 << 30, n = bytes / 16;
    float4 *L0, *L0b, *P1;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&P1, bytes)); CK(cudaMemset(P1, 0, bytes));
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&L0, bytes)); CK(cudaMalloc(&L0b, bytes)); CK(cudaMemset(L0, 0, bytes));
    for (int blocks : {148 * 2, 148 * 8, 148 * 32}) {
        float rd = timeit([&] { k_copy<<<blocks, 256>>>(P1, L0, n); });     // peer read -> local write
        float wr = timeit([&] { k_copy<<<blocks, 256>>>(L0, P1, n); });     // local read -> peer write
        float lo = timeit([&] { k_copy<<<blocks, 256>>>(L0, L0b, n); });
        printf("grid %5d: peer-read %.0f GB/s  peer-write %.0f GB/s  local copy %.0f GB/s (one-way bytes)\n", blocks, bytes / rd / 1e6, bytes / wr / 1e6, bytes / lo / 1e6);
    }
    // tile patterns: matrix 8192 rows x 65536 floats (ld = 65536 floats = 16384 float4) = 2 GiB
    size_t ld4 = 16384, nrows = 8192;
    for (int seg4 : {8, 32, 128}) {          // 128 B, 512 B, 2 KB row segments
        int rows = 8;
        dim3 g((unsigned)((nrows / rows + 3) / 4), (unsigned)(ld4 / seg4));
        float wr = timeit([&] { k_tile<<<g, 128>>>(L0, P1, ld4, rows, nrows, seg4); });
        float rd = timeit([&] { k_tile<<<g, 128>>>(P1, L0, ld4, rows, nrows, seg4); });
        printf("tile %4d B segments x 8 rows/warp: peer-write %.0f GB/s  peer-read %.0f GB/s\n", seg4 * 16, bytes / wr / 1e6, bytes / rd / 1e6);
    }
    float mc = timeit([&] { cudaMemcpyPeerAsync(P1, 1, L0, 0, bytes); });
    printf("cudaMemcpyPeer 0->1: %.0f GB/s\n", bytes / mc / 1e6);
    return 0;
}
