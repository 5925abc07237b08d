// Does a producer/consumer strip buffer stay in the B200's 126 MB L2?
//
// Emulates the layer kernel's data flow: per iteration a PRODUCER reads S bytes of frontier rows from
// DRAM (a fresh region every iteration) and writes S bytes of transposed pairs into a strip buffer; a
// CONSUMER reads that strip (2 random 512-byte row segments per work item) and writes S bytes of new
// frontier rows to DRAM (fresh region).  With NBUF strip buffers rotating, the strip traffic stays in L2
// when NBUF * S fits; with many buffers it spills to DRAM (today's Rt scratch).  Reports the DRAM-side
// copy bandwidth 2 S / t.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ float4 ld_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float4 *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one launch = producer CTAs for strip (it + 1) and consumer CTAs for strip it
template <bool HINT>
__global__ void __launch_bounds__(256) k_iter(const float4 *src, float4 *dst, float4 *strip_w, const float4 *strip_r,
                                              size_t n4, int n_prod, const int *perm, int rows) {
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < n_prod) {
        size_t i = blockIdx.x * (size_t)256 + tid, s = (size_t)n_prod * 256;
        for (; i < n4; i += s) {
            float4 v = HINT ? ld_stream(src + i) : src[i];
            v.x += 1.f;
            strip_w[i] = v;
        }
    } else {
        // consumer: work item = output row r (512 B = 32 float4); reads strip rows perm[2r], perm[2r+1]
        const int nc = gridDim.x - n_prod, cb = blockIdx.x - n_prod;
        const int warp = tid >> 5, lane = tid & 31;
        for (int r = cb * 8 + warp; r < rows; r += nc * 8) {
            const float4 a = strip_r[(size_t)perm[2 * r] * 32 + lane], b = strip_r[(size_t)perm[2 * r + 1] * 32 + lane];
            float4 v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            if (HINT) st_stream(dst + (size_t)r * 32 + lane, v); else dst[(size_t)r * 32 + lane] = v;
        }
    }
}
__global__ void k_read_l2(const float4 *p, size_t n4, int passes, float *out) {
    float acc = 0;
    for (int k = 0; k < passes; k++) {
        size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
        for (; i < n4; i += s) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
    }
    if (acc == 123.456f) *out = acc;
}

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int maxPersist = 0, maxWindow = 0;
    cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    printf("%s: L2 %d MB, max persisting L2 %d MB, max access-policy window %d MB\n", pr.name, pr.l2CacheSize >> 20, maxPersist >> 20, maxWindow >> 20);
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float *out; CK(cudaMalloc(&out, 4));
    // ---- plain L2 read bandwidth vs working-set size ----
    {
        float4 *buf; CK(cudaMalloc(&buf, (size_t)512 << 20)); CK(cudaMemset(buf, 0, (size_t)512 << 20));
        for (int mb : {8, 16, 24, 32, 48, 64, 96, 128, 256}) {
            size_t n4 = ((size_t)mb << 20) / 16;
            const int passes = 40;
            k_read_l2<<<148 * 8, 256, 0, st>>>(buf, n4, 2, out);
            cudaEventRecord(e0, st);
            k_read_l2<<<148 * 8, 256, 0, st>>>(buf, n4, passes, out);
            cudaEventRecord(e1, st); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("read %4d MB x %d passes: %.0f GB/s\n", mb, passes, (double)mb * 1048576.0 * passes / ms / 1e6);
        }
        cudaFree(buf);
    }
    // ---- strip pattern ----
    const size_t big = (size_t)6 << 30;
    float4 *src, *dst, *strips; int *perm;
    CK(cudaMalloc(&src, big)); CK(cudaMalloc(&dst, big)); CK(cudaMemset(src, 0, big)); CK(cudaMemset(dst, 0, big));
    const size_t strips_bytes = (size_t)4 << 30;
    CK(cudaMalloc(&strips, strips_bytes)); CK(cudaMemset(strips, 0, strips_bytes));
    for (int smb : {10, 20, 34, 48}) {
        const size_t S = (size_t)smb << 20, n4 = S / 16;
        const int rows = (int)(S / 512);
        std::vector<int> h(2 * (size_t)rows);
        unsigned long long x = 88172645463325252ULL;
        for (auto &v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = (int)(x % (unsigned long long)rows); }
        CK(cudaMalloc(&perm, h.size() * 4)); CK(cudaMemcpy(perm, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        for (int nbuf : {2, 3, 64}) {
            if ((size_t)nbuf * S > strips_bytes) continue;
            for (int mode = 0; mode < 3; mode++) {          // 0: no hints, 1: streaming hints on DRAM side, 2: + persisting window
                if (mode == 2 && nbuf > 3) continue;
                cudaStreamAttrValue attr = {};
                if (mode == 2) {
                    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxPersist));
                    attr.accessPolicyWindow.base_ptr = strips;
                    attr.accessPolicyWindow.num_bytes = std::min<size_t>((size_t)nbuf * S, (size_t)maxWindow);
                    attr.accessPolicyWindow.hitRatio = 1.0f;
                    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
                }
                const int iters = (int)std::min<size_t>(200, big / S - 1);
                const int n_prod = 148 * 2, n_cons = 148 * 2;
                auto run = [&](int it0, int n) {
                    for (int it = it0; it < it0 + n; it++) {
                        float4 *w = strips + (size_t)((it + 1) % nbuf) * n4;
                        const float4 *r = strips + (size_t)(it % nbuf) * n4;
                        if (mode == 0) k_iter<false><<<n_prod + n_cons, 256, 0, st>>>(src + (size_t)(it % iters) * n4, dst + (size_t)(it % iters) * n4, w, r, n4, n_prod, perm, rows);
                        else k_iter<true><<<n_prod + n_cons, 256, 0, st>>>(src + (size_t)(it % iters) * n4, dst + (size_t)(it % iters) * n4, w, r, n4, n_prod, perm, rows);
                    }
                };
                run(0, 10); CK(cudaStreamSynchronize(st));
                cudaEventRecord(e0, st); run(0, iters); cudaEventRecord(e1, st); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("strip %2d MB x %2d buffers, mode %d: %.2f us/iter, DRAM-side copy %.0f GB/s (2S/t)\n", smb, nbuf, mode,
                       ms * 1e3 / iters, 2.0 * S * iters / ms / 1e6);
                if (mode == 2) {
                    attr.accessPolicyWindow.num_bytes = 0;
                    CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
                    CK(cudaCtxResetPersistingL2Cache());
                }
            }
        }
        cudaFree(perm);
    }
    // reference: plain copy of the same volume per launch
    for (int smb : {20, 48}) {
        const size_t S = (size_t)smb << 20, n4 = S / 16;
        const int iters = 200;
        cudaEventRecord(e0, st);
        for (int it = 0; it < iters; it++) k_iter<true><<<148 * 4, 256, 0, st>>>(src + (size_t)it * n4, dst, dst + (size_t)it * n4, src, n4, 148 * 4, nullptr, 0);
        cudaEventRecord(e1, st); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("plain copy %d MB per launch: %.2f us/iter, %.0f GB/s (r+w)\n", smb, ms * 1e3 / iters, 2.0 * S * iters / ms / 1e6);
    }
    return 0;
}
