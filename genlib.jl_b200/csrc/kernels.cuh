// kernels.cuh -- sm_100a kernels of the level-synchronous kinship sweep.
//
// One generation step (= one cut-vertex step of the reference, src/compute.jl:276-302)
// turns the live frontier matrix Psi (symmetric, stored in full, indexed by SLOT) into the
// next one IN PLACE: rows/columns of individuals that stay live are untouched (the
// reference re-copies them, compute.jl:108-110), rows of evicted individuals are recycled.
//
//   cross_kernel  R[F, p] = 1/2 Psi[f_F, p] + 1/2 Psi[m_F, p]        (compute.jl:111-126)
//                 for every couple F of the layer and every live column p; written
//                   - rounded, as the rows/columns (new member x carried individual), and
//                   - unrounded fp64, transposed, into the scratch block Rt[p, F].
//   intra_kernel  V[F, G] = 1/2 Rt[f_F, G] + 1/2 Rt[m_F, G]         (compute.jl:130-147)
//                 = the kinship of a member of F with a member of G when the F member has
//                 the larger rank (it is "climbed first"); expanded to members, with the
//                 diagonal 1/2 + 1/2 Psi[f, m] (compute.jl:148-155).
//
// Arithmetic is binary64 in the reference's grouping; storage type T is float
// (GENLIB_NUMERICS_REFERENCE: one RN32 per step, like compute.jl:296) or double.
// Both kernels are HBM-bound streaming kernels: 128-bit loads of contiguous row
// segments, shared-memory tile transposes, coalesced stores; no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "plan.hpp"

namespace genlib {

struct LayerArgs {
    int32_t n_new, n_fam, rt_lo, rt_rows, nf_pad, any_carried;
    const int32_t *mem_ind, *mem_slot, *mem_fam;
    const int32_t *fam_pf, *fam_pm, *fam_start;
    const uint8_t *flags;
    const int32_t *mt_minrank, *mt_maxrank;
};

constexpr int kThreads = 256;
constexpr int kSRStride = kPTile + 2;   // doubles; even => 16-byte aligned rows
constexpr int kVStride = kMTile + 1;

// ---- 4-wide row-segment access ------------------------------------------------------
__device__ __forceinline__ void load4(const float *p, double (&d)[4]) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}
__device__ __forceinline__ void load4(const double *p, double (&d)[4]) {
    double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
}
__device__ __forceinline__ void store4(float *p, const double (&d)[4]) {
    // (float)double is cvt.rn.f32.f64: round-to-nearest-even, subnormals kept (no -ftz)
    float4 v = make_float4((float)d[0], (float)d[1], (float)d[2], (float)d[3]);
    *reinterpret_cast<float4 *>(p) = v;
}
__device__ __forceinline__ void store4(double *p, const double (&d)[4]) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(d[0], d[1]);
    reinterpret_cast<double2 *>(p)[1] = make_double2(d[2], d[3]);
}
// 1/2 x + 1/2 y with ONE binary64 rounding (1/2 y is exact), = Julia's `0 + x/2 + y/2`
__device__ __forceinline__ double half_sum(double x, double y) { return fma(0.5, x, 0.5 * y); }

// =====================================================================================
// cross_kernel: grid (live column tiles, family tiles), 256 threads.
// Tile = kFTile couples x kPTile live columns.  Warp w owns couples 4w..4w+3, lane l
// owns columns 4l..4l+3 of the tile (one 128-bit load per parent row).
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads)
cross_kernel(T *__restrict__ A, int64_t ld, double *__restrict__ Rt, LayerArgs L) {
    extern __shared__ double sR[];                       // [kFTile][kSRStride]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int F0 = blockIdx.y * kFTile;
    const int pt = blockIdx.x;
    const int p0 = L.rt_lo + pt * kPTile;
    const uint8_t *fl = L.flags + (size_t)pt * kPTile;
    const uint8_t myflag = fl[threadIdx.x & (kPTile - 1)];
    const int live_here = __syncthreads_or(myflag & kFlagLive);
    if (!live_here) return;                              // hole in a fragmented slot range
    const int carried_here = L.any_carried ? __syncthreads_or(myflag & kFlagCarried) : 0;

    // ---- gather-average of the two parent rows ----
    int pf[4], pm[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int F = F0 + warp * 4 + q;
        const bool ok = F < L.n_fam;
        pf[q] = ok ? L.fam_pf[F] : -1;
        pm[q] = ok ? L.fam_pm[F] : -1;
    }
    double x[4][4], y[4][4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
        for (int k = 0; k < 4; k++) { x[q][k] = 0.0; y[q][k] = 0.0; }
        if (pf[q] >= 0) load4(A + (int64_t)pf[q] * ld + p0 + 4 * lane, x[q]);
        if (pm[q] >= 0) load4(A + (int64_t)pm[q] * ld + p0 + 4 * lane, y[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        double r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = half_sum(x[q][k], y[q][k]);
        double2 *dst = reinterpret_cast<double2 *>(sR + (warp * 4 + q) * kSRStride + 4 * lane);
        dst[0] = make_double2(r[0], r[1]);
        dst[1] = make_double2(r[2], r[3]);
        if (carried_here) {
            // rows of the new members against this tile's columns (rounded once, compute.jl:296).
            // Columns that are not carried receive values nobody reads; new x new is
            // rewritten by intra_kernel afterwards.
            const int F = F0 + warp * 4 + q;
            if (F < L.n_fam) {
                const int m1 = L.fam_start[F + 1];
                for (int m = L.fam_start[F]; m < m1; m++)
                    store4(A + (int64_t)L.mem_slot[m] * ld + p0 + 4 * lane, r);
            }
        }
    }
    __syncthreads();

    // ---- transposed, unrounded: Rt[p, F] for every live column p of the tile ----
    for (int pl = warp; pl < kPTile; pl += kThreads / 32) {
        if (fl[pl] & kFlagLive)
            Rt[((size_t)pt * kPTile + pl) * L.nf_pad + F0 + lane] = sR[lane * kSRStride + pl];
    }
    // ---- mirror: columns of the new members in the rows of carried individuals ----
    if (carried_here) {
        const int m0 = L.fam_start[F0];
        const int Fe = min(F0 + kFTile, L.n_fam);
        const int m1 = L.fam_start[Fe];
        for (int pl = warp; pl < kPTile; pl += kThreads / 32) {
            if (!(fl[pl] & kFlagCarried)) continue;
            T *row = A + (int64_t)(p0 + pl) * ld;
            for (int m = m0 + lane; m < m1; m += 32)
                row[L.mem_slot[m]] = (T)sR[(L.mem_fam[m] - F0) * kSRStride + pl];
        }
    }
}

// =====================================================================================
// intra_kernel: one CTA per pair (I >= J) of member tiles (kMTile members each).
// =====================================================================================
template <typename T>
struct IntraSmem {
    T Vab[kMTile * kVStride];   // [family of I][family of J]: the I member is climbed first
    T Vba[kMTile * kVStride];   // [family of J][family of I]: the J member is climbed first
    T diag[kMTile];
    int32_t fam[2][kMTile], rank[2][kMTile], slot[2][kMTile];
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
intra_kernel(T *__restrict__ A, int64_t ld, const double *__restrict__ Rt, LayerArgs L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IntraSmem<T> &S = *reinterpret_cast<IntraSmem<T> *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // linear index over the lower triangle -> (I, J), I >= J
    const long long t = blockIdx.x;
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long long)I * (I + 1) / 2 > t) I--;
    while ((long long)(I + 1) * (I + 2) / 2 <= t) I++;
    const int J = (int)(t - (long long)I * (I + 1) / 2);
    const int mI0 = I * kMTile, mJ0 = J * kMTile;
    const int cI = min(kMTile, L.n_new - mI0), cJ = min(kMTile, L.n_new - mJ0);

    if (threadIdx.x < 2 * kMTile) {
        const int side = threadIdx.x / kMTile, q = threadIdx.x % kMTile;
        const int c = side ? cJ : cI, m = (side ? mJ0 : mI0) + min(q, c - 1);
        S.fam[side][q] = L.mem_fam[m];
        S.rank[side][q] = L.mem_ind[m];
        S.slot[side][q] = L.mem_slot[m];
    }
    __syncthreads();
    const int fI0 = S.fam[0][0], nfI = S.fam[0][cI - 1] - fI0 + 1;
    const int fJ0 = S.fam[1][0], nfJ = S.fam[1][cJ - 1] - fJ0 + 1;
    const bool same = (I == J);
    // rank ranges decide which orientation can occur at all
    const bool need_ab = same || L.mt_maxrank[I] > L.mt_minrank[J];
    const bool need_ba = !same && L.mt_maxrank[J] > L.mt_minrank[I];

    if (need_ab) {
        for (int f = warp; f < nfI; f += kThreads / 32) {
            const int pf = L.fam_pf[fI0 + f], pm = L.fam_pm[fI0 + f];
            const double *rf = Rt + (size_t)max(pf - L.rt_lo, 0) * L.nf_pad + fJ0;
            const double *rm = Rt + (size_t)max(pm - L.rt_lo, 0) * L.nf_pad + fJ0;
            for (int g = lane; g < nfJ; g += 32) {
                const double a = pf >= 0 ? __ldg(rf + g) : 0.0;
                const double b = pm >= 0 ? __ldg(rm + g) : 0.0;
                S.Vab[f * kVStride + g] = (T)half_sum(a, b);
            }
        }
    }
    if (need_ba) {
        for (int g = warp; g < nfJ; g += kThreads / 32) {
            const int pf = L.fam_pf[fJ0 + g], pm = L.fam_pm[fJ0 + g];
            const double *rf = Rt + (size_t)max(pf - L.rt_lo, 0) * L.nf_pad + fI0;
            const double *rm = Rt + (size_t)max(pm - L.rt_lo, 0) * L.nf_pad + fI0;
            for (int f = lane; f < nfI; f += 32) {
                const double a = pf >= 0 ? __ldg(rf + f) : 0.0;
                const double b = pm >= 0 ? __ldg(rm + f) : 0.0;
                S.Vba[g * kVStride + f] = (T)half_sum(a, b);
            }
        }
    }
    if (same && threadIdx.x < nfI) {
        // compute.jl:148-155: 1/2 + 1/2 Psi[father, mother] when both parents are known
        const int pf = L.fam_pf[fI0 + threadIdx.x], pm = L.fam_pm[fI0 + threadIdx.x];
        double v = 0.5;
        if (pf >= 0 && pm >= 0) v = fma(0.5, (double)A[(int64_t)pf * ld + pm], 0.5);
        S.diag[threadIdx.x] = (T)v;
    }
    __syncthreads();

    // ---- expand couples to members: rows of I ----
    for (int il = warp; il < cI; il += kThreads / 32) {
        const int f = S.fam[0][il] - fI0, ri = S.rank[0][il];
        T *row = A + (int64_t)S.slot[0][il] * ld;
        for (int jl = lane; jl < cJ; jl += 32) {
            const int g = S.fam[1][jl] - fJ0, rj = S.rank[1][jl];
            T v;
            if (same) v = (ri == rj) ? S.diag[f] : (ri > rj ? S.Vab[f * kVStride + g] : S.Vab[g * kVStride + f]);
            else v = ri > rj ? S.Vab[f * kVStride + g] : S.Vba[g * kVStride + f];
            row[S.slot[1][jl]] = v;
        }
    }
    // ---- and the symmetric block: rows of J ----
    if (!same) {
        for (int jl = warp; jl < cJ; jl += kThreads / 32) {
            const int g = S.fam[1][jl] - fJ0, rj = S.rank[1][jl];
            T *row = A + (int64_t)S.slot[1][jl] * ld;
            for (int il = lane; il < cI; il += 32) {
                const int f = S.fam[0][il] - fI0, ri = S.rank[0][il];
                row[S.slot[0][il]] = ri > rj ? S.Vab[f * kVStride + g] : S.Vba[g * kVStride + f];
            }
        }
    }
}

// =====================================================================================
// proband gather (compute.jl:303: the last frontier, rows/columns in probandIDs order)
// =====================================================================================
template <typename T, typename O>
__global__ void gather_kernel(const T *__restrict__ A, int64_t ld, const int32_t *__restrict__ slots,
                              int32_t P, int32_t row0, int32_t nrows, O *__restrict__ out) {
    const int u = row0 + blockIdx.y;
    if (blockIdx.y >= (unsigned)nrows) return;
    const T *row = A + (int64_t)slots[u] * ld;
    O *dst = out + (size_t)blockIdx.y * P;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < P; v += gridDim.x * blockDim.x)
        dst[v] = (O)row[slots[v]];
}

// sum and trace of the proband block, binary64 accumulation (phiMean, compute.jl:454-459)
template <typename T>
__global__ void mean_kernel(const T *__restrict__ A, int64_t ld, const int32_t *__restrict__ slots,
                            int32_t P, double *__restrict__ acc /* [0]=sum, [1]=trace */) {
    __shared__ double ssum[kThreads / 32], str[kThreads / 32];
    double s = 0.0, tr = 0.0;
    for (int u = blockIdx.x; u < P; u += gridDim.x) {
        const T *row = A + (int64_t)slots[u] * ld;
        for (int v = threadIdx.x; v < P; v += blockDim.x) {
            const double x = (double)row[slots[v]];
            s += x;
            if (u == v) tr += x;
        }
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); tr += __shfl_down_sync(0xffffffffu, tr, o); }
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = s; str[threadIdx.x >> 5] = tr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; w++) { s += ssum[w]; tr += str[w]; }
        atomicAdd(acc, s);
        atomicAdd(acc + 1, tr);
    }
}

}  // namespace genlib
