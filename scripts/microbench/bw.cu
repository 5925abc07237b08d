// Microbenchmarks of the raw HBM paths the kinship kernels rely on (B200, sm_100a).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void k_write(float4 *p, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    float4 v = make_float4(1.f, 2.f, 3.f, (float)i);
    for (; i < n; i += s) p[i] = v;
}
__global__ void k_read(const float4 *p, size_t n, float *out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    float acc = 0;
    for (; i < n; i += s) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) *out = acc;
}
__global__ void k_copy(const float4 *a, float4 *b, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) b[i] = a[i];
}
// tile writes: each warp writes ROWS rows x SEG bytes (row stride ld bytes), like the expand kernel
template <int SEGF4>   // float4 per row segment per warp (32 => 512 B)
__global__ void k_tile_write(float4 *p, size_t ld4, int rows_per_warp, int nseg, size_t nrows) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    size_t row0 = (size_t)warp * rows_per_warp;
    if (row0 >= nrows) return;
    float4 v = make_float4(1.f, 2.f, 3.f, (float)lane);
    for (int s = 0; s < nseg; s++) {
        size_t col = ((size_t)blockIdx.y * nseg + s) * SEGF4;
        for (int r = 0; r < rows_per_warp; r++)
            for (int c = lane; c < SEGF4; c += 32) p[(row0 + r) * ld4 + col + c] = v;
    }
}
__global__ void k_cvt(const float *a, float *b, size_t n, int iters) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = a[i];
    double acc = 0;
    for (int k = 0; k < iters; k++) { acc = fma(0.5, (double)x, 0.5 * acc); x = (float)acc + 1e-30f; }
    b[i] = x;
}
template <typename F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    size_t bytes = (size_t)8 << 30;            // 8 GiB buffers
    float4 *A, *B; float *out;
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&out, 4));
    size_t n = bytes / 16;
    CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes));
    for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        float w = timeit([&] { k_write<<<blocks, 256>>>(A, n); });
        float r = timeit([&] { k_read<<<blocks, 256>>>(A, n, out); });
        float c = timeit([&] { k_copy<<<blocks, 256>>>(A, B, n); });
        printf("grid %5d: write %.0f GB/s  read %.0f GB/s  copy %.0f GB/s (r+w)\n", blocks, bytes / w / 1e6, bytes / r / 1e6, 2.0 * bytes / c / 1e6);
    }
    float m = timeit([&] { cudaMemsetAsync(A, 1, bytes); });
    printf("cudaMemset: %.0f GB/s\n", bytes / m / 1e6);
    float mc = timeit([&] { cudaMemcpyAsync(B, A, bytes, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D: %.0f GB/s (r+w)\n", 2.0 * bytes / mc / 1e6);
    // tile writes into a 40000 x 40000 float matrix region (ld = 80000 floats)
    size_t ldf = 80000, nrows = 20000; size_t ld4 = ldf / 4;   // 20000 x 320 KB = 6.4 GB < 8 GiB
    {
        // 512 B segments, 8 rows per warp, 16 segments per CTA-y (like expand v5)
        int rows = 8, nseg = 16; int cols4 = 40000 / 4; int gy = cols4 / (32 * nseg);
        dim3 g((unsigned)((nrows / rows + 3) / 4), gy);
        float t = timeit([&] { k_tile_write<32><<<g, 128>>>(A, ld4, rows, nseg, nrows); });
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        printf("tile write 8 rows x 512B x16: %.0f GB/s\n", (double)nrows * gy * nseg * 512 / t / 1e6);
        rows = 8; nseg = 4; gy = cols4 / (128 * nseg);
        dim3 g2((unsigned)((nrows / rows + 3) / 4), gy);
        t = timeit([&] { k_tile_write<128><<<g2, 128>>>(A, ld4, rows, nseg, nrows); });
        printf("tile write 8 rows x 2048B x4: %.0f GB/s\n", (double)nrows * gy * nseg * 2048 / t / 1e6);
        rows = 1; nseg = 1; gy = 1;
        dim3 g3((unsigned)((nrows / rows + 3) / 4), gy);
        t = timeit([&] { k_tile_write<10000><<<g3, 128>>>(A, ld4, rows, nseg, nrows); });
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        printf("row write 1 row x 160000B: %.0f GB/s\n", (double)nrows * 40000 * 4 / t / 1e6);
    }
    {
        size_t ne = (size_t)1 << 26;
        for (int iters : {16, 64}) {
            float t = timeit([&] { k_cvt<<<(unsigned)(ne / 256), 256>>>((float *)A, (float *)B, ne, iters); });
            printf("cvt chain iters=%d: %.1f G(f32->f64 + dfma + dmul + f64->f32)/s\n", iters, (double)ne * iters / t / 1e6);
        }
    }
    return 0;
}
