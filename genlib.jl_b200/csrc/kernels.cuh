// kernels.cuh -- sm_100a kernels of the level-synchronous kinship sweep.
//
// One generation step (= one cut-vertex step of the reference, src/compute.jl:276-302)
// turns the live frontier matrix Psi (symmetric, stored in full, indexed by SLOT) into the
// next one IN PLACE: rows/columns of individuals that stay live are untouched (the
// reference re-copies them, compute.jl:108-110), rows of evicted individuals are recycled.
//
//   cross_kernel  R[F, p] = 1/2 Psi[f_F, p] + 1/2 Psi[m_F, p]        (compute.jl:111-126)
//                 for every couple F of the layer and every live column p; written
//                   - rounded, as the rows (new member x carried individual), and
//                   - unrounded fp64, transposed, into the scratch block Rt[p, F].
//   mirror_kernel the same values as the columns of the new members in the carried rows.
//   couple_kernel V[F, G] = 1/2 Rt[f_F, G] + 1/2 Rt[m_F, G]         (compute.jl:130-147)
//                 = the kinship of a member of F with a member of G when the F member has
//                 the larger rank (it is "climbed first"); plus the diagonal value
//                 1/2 + 1/2 Psi[f, m] of the couple's members (compute.jl:148-155).
//   expand_kernel couples -> members: entry (i, j) = V[F_i, G_j] or V[G_j, F_i] by rank.
//
// Arithmetic is binary64 in the reference's grouping; storage type T is float
// (GENLIB_NUMERICS_REFERENCE: one RN32 per step, like compute.jl:296) or double.
// STORED = true selects the arithmetic of gen.sparse_phi instead (compute.jl:321-447: every
// stored kinship is a Float32 and is halved in Float32; the plan then orders the layers and the
// members by sparse_phi's queue, see plan.hpp).
// All are HBM-bound streaming kernels: TMA bulk copies (cross) / 128-bit loads of contiguous
// row segments, shared-memory tile transposes, coalesced stores; no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <climits>
#include <cstdint>

#include "plan.hpp"

namespace genlib {

constexpr int kMaxWorld = 16;

// Where every rank keeps its frontier rows and its couple-matrix rows.  One process per GPU:
// entries of other ranks are CUDA-IPC mappings of their arenas (NVLink peer memory).
struct PeerTable {
    void *A[kMaxWorld];       // rows_cap[g] x ld frontier rows of rank g
    void *Vrow[kMaxWorld];    // own couples x nf_pad rows of V on rank g
};

struct LayerArgs {
    int32_t n_new, n_fam, rt_lo, rt_rows, nf_pad, any_carried;
    // row sharding: this rank owns couples [own_f0, own_f0 + own_nf) and members
    // [own_m0, own_m0 + own_nm); nfo_pad = row stride of its transposed cross block
    int32_t rank, world, own_f0, own_nf, own_m0, own_nm, nfo_pad, ftile_shift;
    int32_t fam_base[kMaxWorld + 1];
    const int32_t *mem_ind, *mem_slot, *mem_fam, *mem_lrow;
    const int32_t *fam_pf, *fam_pm, *fam_start;
    const int8_t *fam_pf_owner, *fam_pm_owner, *live_owner, *mem_gowner;
    const int32_t *fam_pf_lrow, *fam_pm_lrow, *live_lrow, *mem_glrow;
    const uint8_t *flags;
    const int32_t *fam_minrank, *fam_maxrank;
    const int32_t *mt_minrank, *mt_maxrank, *mt_fam0, *mt_nfam, *mt_m0, *mt_cnt;
    int32_t n_mtiles;
    int32_t vstride;            // staged couple segment of expand_kernel: row stride (elements)
    int32_t pchunk;             // live-column tiles per cross_kernel CTA (<= kMaxPChunk)
    int32_t ctile0;             // first own-couple tile of this launch: cross_kernel tiles of kFTile (blockIdx.y + ctile0),
                                // couple_kernel column tiles of kCTile (blockIdx.x + ctile0 * kFTile / kCTile)
};

constexpr int kThreads = 256;

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
__device__ __forceinline__ void store_vec4(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store_vec4(double *p, const double (&v)[4]) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(v[0], v[1]);
    reinterpret_cast<double2 *>(p)[1] = make_double2(v[2], v[3]);
}
// 1/2 x + 1/2 y with ONE binary64 rounding (1/2 y is exact), = Julia's `0 + x/2 + y/2`
__device__ __forceinline__ double half_sum(double x, double y) { return fma(0.5, x, 0.5 * y); }
// sparse_phi's `phi[..] / 2` is a FLOAT32 division of a stored Float32 (compute.jl:350-389): it rounds
// when the stored value is a subnormal with an odd last bit; the sum is Float64 as above.
__device__ __forceinline__ double half_sum_stored(double x, double y) {
    return (double)__fmul_rn((float)x, 0.5f) + (double)__fmul_rn((float)y, 0.5f);
}
template <bool STORED>
__device__ __forceinline__ double half_sum_mode(double x, double y) {
    if constexpr (STORED) return half_sum_stored(x, y);
    else return half_sum(x, y);
}

// =====================================================================================
// cross_kernel: grid (chunks of L.pchunk live-column tiles, own couple tiles), 256 threads.
// Tile = kFTile couples x kPTile live columns.  A CTA keeps ONE couple tile (the 64 parent rows
// are resolved once: local HBM or a peer's, read through NVLink) and streams the live column
// tiles of its chunk through a kCrossStages-deep shared-memory ring with 16-byte cp.async, so
// the bytes in flight per SM are a design parameter (2 stages x 32 KB x 2 CTAs) instead of a
// consequence of occupancy -- the previous one-tile-per-CTA version waited on its own loads
// (long-scoreboard stalls, 51 % of DRAM peak in profiles/r01/ncu_full_c3_summary.json).
// The stage holds the RAW parent rows; the unrounded sums are formed when they are written:
//   (a) transposed into Rt[p, F]: lane = couple, the column rotates with lane/8 so that the
//       32 shared loads of a warp hit 32 banks (rows are padded by 16 bytes);
//   (b) only in tiles with carried columns: rounded, as the rows of the couple's members.
// =====================================================================================
constexpr int kCrossStages = 3;
constexpr int kMaxPChunk = 32;               // column tiles per CTA (upper bound of L.pchunk)
template <typename T> __host__ __device__ constexpr int cross_row_bytes() { return kPTile * (int)sizeof(T) + 16; }
template <typename T> __host__ __device__ constexpr int cross_stage_bytes() { return 2 * kFTile * cross_row_bytes<T>(); }
template <typename T> constexpr size_t cross_smem_bytes() { return (size_t)kCrossStages * cross_stage_bytes<T>(); }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16_to(unsigned smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem), "l"(gmem) : "memory");
}
// mbarrier + bulk-copy (TMA, SASS UBLKCP) helpers
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16), completes on `bar`
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void zero16_shared(unsigned smem) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};\n" ::"r"(smem), "r"(0) : "memory");
}
__device__ __forceinline__ void lds4(const float *p, double (&d)[4]) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}
__device__ __forceinline__ void lds4(const double *p, double (&d)[4]) {
    const double2 a = reinterpret_cast<const double2 *>(p)[0], b = reinterpret_cast<const double2 *>(p)[1];
    d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
}

template <typename T, bool STORED>     // STORED: the sparse_phi schedule (Float32 halves of stored values)
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 2 : 1)
cross_kernel(T *__restrict__ A, int64_t ld, double *__restrict__ Rt, PeerTable PT, LayerArgs L) {
    extern __shared__ __align__(16) unsigned char cross_smem[];      // [stage][father rows | mother rows][row bytes]
    __shared__ const T *s_row[2 * kFTile];                           // parent rows at the chunk's first column
    __shared__ __align__(16) uint8_t s_flag[kMaxPChunk * kPTile];    // column flags of the chunk
    __shared__ int s_tile[kMaxPChunk];                               // live tiles: index | carried << 8
    __shared__ int s_ntile;
    __shared__ __align__(8) unsigned long long s_bar[kCrossStages];  // "stage filled" mbarriers
    constexpr int RB = cross_row_bytes<T>(), STAGE = cross_stage_bytes<T>();
    constexpr unsigned ROWB = kPTile * (unsigned)sizeof(T);          // bytes of one row segment
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int F0 = (blockIdx.y + L.ctile0) * kFTile;                 // local couple index (own couples only)
    const int t0 = blockIdx.x * L.pchunk;
    const int nt_all = min(L.pchunk, L.rt_rows / kPTile - t0);
    const int c0 = t0 * kPTile;                                      // first column of the chunk, from rt_lo

    {
        const uint4 *src = reinterpret_cast<const uint4 *>(L.flags + c0);
        uint4 *dst = reinterpret_cast<uint4 *>(s_flag);
        for (int i = tid; i < nt_all * (kPTile / 16); i += kThreads) dst[i] = __ldg(src + i);
    }
    if (tid < 2 * kFTile) {
        const int Fl = F0 + (tid & (kFTile - 1));
        const T *row = nullptr;
        if (Fl < L.own_nf) {
            const int F = L.own_f0 + Fl;
            const int o = tid < kFTile ? L.fam_pf_owner[F] : L.fam_pm_owner[F];
            if (o >= 0) {
                const int lr = tid < kFTile ? L.fam_pf_lrow[F] : L.fam_pm_lrow[F];
                row = static_cast<const T *>(PT.A[o]) + (int64_t)lr * ld + L.rt_lo + c0;
            }
        }
        s_row[tid] = row;
    }
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < kCrossStages; st++) mbar_init((unsigned)__cvta_generic_to_shared(&s_bar[st]), 2 * kFTile);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {                                                 // holes of a fragmented slot range are skipped
        int info = 0;
        if (lane < nt_all) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(s_flag + lane * kPTile);
            uint32_t acc = 0;
            for (int k = 0; k < kPTile / 4; k++) acc |= w[(k + lane) & (kPTile / 4 - 1)];
            info = ((acc & 0x01010101u * kFlagLive) ? 1 : 0) | ((acc & 0x01010101u * kFlagCarried) ? 0x100 : 0);
        }
        const unsigned m = __ballot_sync(0xffffffffu, info & 1);
        if (info & 1) s_tile[__popc(m & ((1u << lane) - 1u))] = lane | (info & 0x100);
        if (lane == 0) s_ntile = __popc(m);
    }
    __syncthreads();
    const int nt = s_ntile;
    if (nt == 0) return;

    int mb[4], me[4];                                                // members of the warp's 4 couples (rows to write)
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int Fl = F0 + warp * 4 + q;
        mb[q] = 0; me[q] = 0;
        if (L.any_carried && Fl < L.own_nf) { mb[q] = L.fam_start[L.own_f0 + Fl]; me[q] = L.fam_start[L.own_f0 + Fl + 1]; }
    }

    const unsigned sbase = (unsigned)__cvta_generic_to_shared(cross_smem);
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&s_bar[0]);
    // One bulk copy (TMA) per parent row and tile: 512-byte (1 KB) requests, which is what NVLink
    // wants for the rows that live on a peer -- 16-byte cp.async reached only ~300 GB/s there.
    // Lanes 0..7 of every warp issue one row each and arrive on the stage's mbarrier.
    auto issue = [&](int k) {                                        // k-th live tile -> stage k % kCrossStages
        if (lane < 2 * kFTile / (kThreads / 32)) {
            const int row = warp * (2 * kFTile / (kThreads / 32)) + lane;
            const int ti = s_tile[k] & 0xff;
            const unsigned bar = bar0 + 8u * (unsigned)(k % kCrossStages);
            const unsigned dst = sbase + (unsigned)(k % kCrossStages) * STAGE + (unsigned)(row * RB);
            const T *src = s_row[row];
            if (src) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the stage
                mbar_arrive_expect_tx(bar, ROWB);
                bulk_g2s(dst, src + ti * kPTile, ROWB, bar);
            } else {                                                 // unknown parent: contributes 0 (compute.jl:111-126)
                for (unsigned c = 0; c < ROWB; c += 16) zero16_shared(dst + c);
                mbar_arrive_expect_tx(bar, 0);
            }
        }
    };
#pragma unroll
    for (int k = 0; k < kCrossStages - 1; k++)
        if (k < nt) issue(k);
    const int rot0 = lane >> 3;
    for (int k = 0; k < nt; k++) {
        mbar_wait(bar0 + 8u * (unsigned)(k % kCrossStages), (unsigned)(k / kCrossStages) & 1u);   // tile k landed
        __syncthreads();                                             // everybody is done with stage (k-1) % S
        if (k + kCrossStages - 1 < nt) issue(k + kCrossStages - 1);
        const int tinfo = s_tile[k], ti = tinfo & 0xff;
        const unsigned char *st = cross_smem + (k % kCrossStages) * STAGE;
        const uint8_t *fl = s_flag + ti * kPTile;
        // ---- transposed, unrounded: Rt[p, F] for every live column p of the tile ----
        {
            const T *xr = reinterpret_cast<const T *>(st + lane * RB);
            const T *yr = reinterpret_cast<const T *>(st + (kFTile + lane) * RB);
            double *rt = Rt + (size_t)(c0 + ti * kPTile) * L.nfo_pad + F0 + lane;
#pragma unroll
            for (int g = 0; g < kPTile / 32; g++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int col = warp * (kPTile / 8) + g * 4 + ((j + rot0) & 3);
                    if (fl[col] & kFlagLive) rt[(size_t)col * L.nfo_pad] = half_sum_mode<STORED>((double)xr[col], (double)yr[col]);
                }
            }
        }
        // ---- rows of the new members against this tile's columns (rounded once, compute.jl:296).
        //      Columns that are not carried receive values nobody reads; new x new is rewritten
        //      by expand_kernel afterwards. ----
        if (tinfo & 0x100) {
            const int64_t col0 = (int64_t)L.rt_lo + c0 + ti * kPTile + 4 * lane;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (me[q] <= mb[q]) continue;
                double x[4], y[4], r[4];
                lds4(reinterpret_cast<const T *>(st + (warp * 4 + q) * RB) + 4 * lane, x);
                lds4(reinterpret_cast<const T *>(st + (kFTile + warp * 4 + q) * RB) + 4 * lane, y);
#pragma unroll
                for (int e = 0; e < 4; e++) r[e] = half_sum_mode<STORED>(x[e], y[e]);
                for (int m = mb[q]; m < me[q]; m++) {
                    store4(A + (int64_t)L.mem_lrow[m] * ld + col0, r);
                    if (L.world > 1) {
                        const int go = L.mem_gowner[m];          // guest copy of the new row (GENLIB_GUESTS=1)
                        if (go >= 0) store4(static_cast<T *>(PT.A[go]) + (int64_t)L.mem_glrow[m] * ld + col0, r);
                    }
                }
            }
        }
    }
}

// =====================================================================================
// mirror_kernel: the columns of the new members in the rows of the CARRIED individuals,
// Psi[c, i] = RN(R[F_i, c]) (compute.jl:119-126 by symmetry).  One warp per carried row: it reads
// that row of the transposed cross block Rt[c, own couples] (contiguous) and writes the members'
// columns (contiguous slots) -- a peer store when the carried row lives on another GPU.
// grid (member chunks of 32 x kMirrorCols, live rows / 8).
// =====================================================================================
constexpr int kMirrorCols = 8;     // members per lane and CTA column chunk

template <typename T>
__global__ void __launch_bounds__(kThreads)
mirror_kernel(const double *__restrict__ Rt, int64_t ld, PeerTable PT, LayerArgs L) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.y * (kThreads / 32) + warp;          // row of the live slot range
    if (r >= L.rt_rows || !(L.flags[r] & kFlagCarried)) return;
    T *row = static_cast<T *>(PT.A[L.live_owner[r]]) + (int64_t)L.live_lrow[r] * ld;
    const double *src = Rt + (size_t)r * L.nfo_pad - L.own_f0;  // indexed by global couple
    const int m0 = L.own_m0 + blockIdx.x * 32 * kMirrorCols, m1 = min(L.own_m0 + L.own_nm, m0 + 32 * kMirrorCols);
#pragma unroll
    for (int k = 0; k < kMirrorCols; k++) {
        const int m = m0 + k * 32 + lane;
        if (m < m1) row[L.mem_slot[m]] = (T)src[L.mem_fam[m]];
    }
}

// =====================================================================================
// couple_kernel: V[F, G] = 1/2 Rt[f_F, G] + 1/2 Rt[m_F, G]  and its transpose Vt[G, F].
// Same shape as cross_kernel: tile = kFTile couple rows x kCTile couple columns, warp w owns
// rows 4w..4w+3, lane l owns columns 4l..4l+3 (two 128-bit loads per parent row); the tile
// goes out row-major (V) and, through shared memory, transposed (Vt), so that expand_kernel
// finds both orientations of a couple pair at the SAME offset of two row-major matrices.
// Rows whose members cannot outrank any member of the tile's columns are skipped.
// Also Dg[F] = 1/2 + 1/2 Psi[f_F, m_F], the diagonal of the couple's members (compute.jl:148-155).
// =====================================================================================
constexpr int kCTile = 128;
constexpr int kCStride = kCTile + 1;
// couple rows per CTA: several passes of kFTile rows, so that the transposed copy goes out in 512-byte runs
template <typename T> __host__ __device__ constexpr int couple_rows() { return (sizeof(T) == 4 ? 4 : 2) * kFTile; }

template <typename T, bool STORED>
__global__ void __launch_bounds__(kThreads, 3)
couple_kernel(int64_t ld, const double *__restrict__ Rt, T *__restrict__ Vt, T *__restrict__ Dg, PeerTable PT,
              LayerArgs L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kCRows = couple_rows<T>();
    T *sV = reinterpret_cast<T *>(smem_raw);                  // [kCRows][kCStride]
    __shared__ int s_skip[kCRows];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // rows: ALL couples F of the layer; columns: this rank's own couples G (local index gl).
    // The row tiles are visited starting behind this rank's own range: at any moment the ranks
    // push to DIFFERENT owners instead of all hitting the same GPU's NVLink ingress.
    const int ytile = (int)((blockIdx.y + (unsigned)L.ftile_shift * kFTile / kCRows) % gridDim.y);
    const int F0 = ytile * kCRows, G0 = (blockIdx.x + L.ctile0 * kFTile / kCTile) * kCTile;
    const int gl = G0 + 4 * lane;
    int minG = INT_MAX;                                       // lowest rank among the members of the tile's column couples
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (gl + k < L.own_nf) minG = min(minG, L.fam_minrank[L.own_f0 + gl + k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) minG = min(minG, __shfl_xor_sync(0xffffffffu, minG, o));
    const bool col_ok = gl < L.nfo_pad;                       // nfo_pad is a multiple of 4
#pragma unroll 1
    for (int half = 0; half < kCRows / kFTile; half++) {
    // phase 1: the four rows of this warp -- couple, parents, skip; the couple diagonal
    int pfs[4], pms[4];
    bool skips[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int fl = half * kFTile + warp * 4 + q, F = F0 + fl;
        skips[q] = true; pfs[q] = -1; pms[q] = -1;
        if (F < L.n_fam && G0 < L.own_nf) {
            pfs[q] = L.fam_pf[F]; pms[q] = L.fam_pm[F];
            skips[q] = L.fam_maxrank[F] <= minG;              // nobody in F outranks anybody in the tile
        }
        if (lane == 0) s_skip[fl] = skips[q];
    }
    if (G0 == 0 && lane < 4) {                                // diagonal of the couple's members
        const int F = F0 + half * kFTile + warp * 4 + lane;
        if (F >= L.own_f0 && F < L.own_f0 + L.own_nf) {
            const int pf = L.fam_pf[F], pm = L.fam_pm[F];
            double d = 0.5;
            if (pf >= 0 && pm >= 0)
                d = half_sum_mode<STORED>((double)(static_cast<const T *>(PT.A[L.fam_pf_owner[F]]) + (int64_t)L.fam_pf_lrow[F] * ld)[pm], 1.0);
            Dg[F] = (T)d;
        }
    }
    // phase 2: all parent-row loads of the warp in flight together (two 128-bit loads per row and parent)
    double a[4][4], b[4][4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
        for (int k = 0; k < 4; k++) { a[q][k] = 0.0; b[q][k] = 0.0; }
        if (!skips[q] && col_ok) {
            if (pfs[q] >= 0) load4(Rt + (size_t)(pfs[q] - L.rt_lo) * L.nfo_pad + gl, a[q]);
            if (pms[q] >= 0) load4(Rt + (size_t)(pms[q] - L.rt_lo) * L.nfo_pad + gl, b[q]);
            if constexpr (STORED) {                              // the STORED (Float32) cross values (compute.jl:331, 363-395)
#pragma unroll
                for (int k = 0; k < 4; k++) { a[q][k] = (double)(T)a[q][k]; b[q][k] = (double)(T)b[q][k]; }
            }
        }
    }
    // phase 3: V[F, own G] goes to the rank that owns couple F (its row block of V): local, or a
    // 16-byte peer store over NVLink; the tile is kept in shared memory for the transposed copy
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int fl = half * kFTile + warp * 4 + q, F = F0 + fl;
        if (!skips[q] && col_ok) {
            T v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                v[k] = (T)half_sum_mode<STORED>(a[q][k], b[q][k]);
                sV[fl * kCStride + 4 * lane + k] = v[k];
            }
            int o = 0;
            while (o + 1 < L.world && F >= L.fam_base[o + 1]) o++;
            T *vrow = static_cast<T *>(PT.Vrow[o]) + (size_t)(F - L.fam_base[o]) * L.nf_pad + L.own_f0;
            if (gl < L.own_nf) store_vec4(vrow + gl, v);       // never into the next rank's couple columns
        }
    }
    }
    __syncthreads();
    // transposed, local: Vt[own G, F0 .. F0 + kCRows), one couple column per warp iteration, lane = couple
    // row (the passes back to back: one contiguous run per column)
    bool row_ok[kCRows / kFTile];
#pragma unroll
    for (int h = 0; h < kCRows / kFTile; h++) row_ok[h] = (F0 + h * kFTile + lane < L.n_fam) && !s_skip[h * kFTile + lane];
    const int gl_end = min(kCTile, L.own_nf - G0);
    for (int g = warp; g < gl_end; g += kThreads / 32) {
        T *dst = Vt + (size_t)(G0 + g) * L.nf_pad + F0 + lane;
#pragma unroll
        for (int h = 0; h < kCRows / kFTile; h++)
            if (row_ok[h]) dst[h * kFTile] = sV[(h * kFTile + lane) * kCStride + g];
    }
}

// =====================================================================================
// expand_kernel: couples -> members.  Every WARP owns kERows consecutive member rows and sweeps
// the member columns in steps of kMTile (= 4 per lane), writing only its own rows (one
// contiguous 512-byte segment per row and step); the symmetric partner block is written by
// the warp that owns those rows -- no transposed stores, no block-wide barriers.
//   entry (i, j), i in couple F, j in couple G:  rank_i > rank_j ? V[F, G] : V[G, F] = Vt[F, G]
//   (compute.jl:130-147: the higher rank is climbed first);  i == j: Dg[F].
// The <= kERows couple rows of V and Vt that a warp needs for the next column step, and that
// step's column metadata, stream into warp-private shared memory with cp.async while the
// current step is expanded (double buffered).
// =====================================================================================
constexpr int kERows = 8;          // member rows per warp
constexpr int kEWarps = 4;         // warps per CTA
constexpr int kEChunk = 16;        // column steps per CTA
constexpr int kExpandThreads = kEWarps * 32;

template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(s), "l"(gmem), "n"(BYTES) : "memory");
}

// per warp and stage: Vab[kERows][vstride], Vba[kERows][vstride]
template <typename T>
__host__ __device__ inline size_t expand_stage_bytes(int vstride) {
    return 2 * (size_t)kERows * vstride * sizeof(T);
}

template <typename T>
__device__ __forceinline__ T lds(unsigned addr) {
    T v;
    if constexpr (sizeof(T) == 4) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    else asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int4 lds_int4(unsigned addr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async16_s(unsigned smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem), "l"(gmem) : "memory");
}

// One column step of one warp.  FAST: all kERows rows exist, the column tile is complete and the
// lane's four column slots are consecutive and 16-byte aligned (one 128-bit store per row).
template <typename T, bool FAST, bool DIAG, bool GUESTS>
__device__ __forceinline__ void expand_step(const unsigned (&go)[4], const int (&rj)[4], const int (&sj)[4],
                                            const unsigned (&roff)[kERows], const int (&rrank)[kERows],
                                            T *const (&rptr)[kERows], T *const (&gptr)[kERows], unsigned vba_off,
                                            int nr, int ncol, int dk0, const T *__restrict__ Dg,
                                            const int (&rfam)[kERows]) {
#pragma unroll
    for (int r = 0; r < kERows; r++) {
        if (FAST || r < nr) {
            T v[4];
#pragma unroll
            for (int k = 0; k < 4; k++)                    // the higher rank is climbed first: V[F, G] if the row
                v[k] = lds<T>(go[k] + roff[r] + (rrank[r] > rj[k] ? 0u : vba_off));   // outranks the column, else Vt[F, G]
            if (DIAG && (unsigned)(dk0 + r) < 4u) {        // own diagonal entry (compute.jl:148-155)
                const T d = Dg[rfam[r]];
#pragma unroll
                for (int k = 0; k < 4; k++) if (dk0 + r == k) v[k] = d;
            }
            if (FAST) store_vec4(rptr[r] + sj[0], v);
            else {
#pragma unroll
                for (int k = 0; k < 4; k++) if (k < ncol) rptr[r][sj[k]] = v[k];
            }
            if (GUESTS && gptr[r]) {                       // the same row, into the guest copy on another GPU
                if (FAST) store_vec4(gptr[r] + sj[0], v);
                else {
#pragma unroll
                    for (int k = 0; k < 4; k++) if (k < ncol) gptr[r][sj[k]] = v[k];
                }
            }
        }
    }
}

template <typename T, bool GUESTS>
__global__ void __launch_bounds__(kExpandThreads, 6)
expand_kernel(T *__restrict__ A, int64_t ld, const T *__restrict__ V, const T *__restrict__ Vt,
              const T *__restrict__ Dg, PeerTable PT, LayerArgs L) {
    constexpr int kVec = 16 / sizeof(T);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = column chunk (fastest): CTAs that run together stream the SAME couple rows of
    // V / Vt and the same output rows at adjacent columns
    const int row0 = L.own_m0 + (blockIdx.y * kEWarps + warp) * kERows;   // this rank's member rows only
    if (row0 >= L.own_m0 + L.own_nm) return;                // no block-wide barrier below
    const int nr = min(kERows, L.own_m0 + L.own_nm - row0);
    const unsigned stage_bytes = (unsigned)expand_stage_bytes<T>(L.vstride);
    const unsigned mine = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)warp * 2u * stage_bytes;
    const unsigned row_bytes = (unsigned)L.vstride * (unsigned)sizeof(T);
    const unsigned vba_off = kERows * row_bytes;

    // ---- the warp's rows: couple, rank, row pointer (registers) ----
    const int mrow = row0 + min(lane, nr - 1);
    const int myfam = L.mem_fam[mrow], myrank = L.mem_ind[mrow], myslot = L.mem_lrow[mrow];   // local row
    const int f0 = __shfl_sync(0xffffffffu, myfam, 0);
    const int nfr = __shfl_sync(0xffffffffu, myfam, nr - 1) - f0 + 1;      // <= kERows couples
    int minI = myrank, maxI = myrank;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {                       // lanes >= nr repeat the last row
        minI = min(minI, __shfl_xor_sync(0xffffffffu, minI, o));
        maxI = max(maxI, __shfl_xor_sync(0xffffffffu, maxI, o));
    }
    minI = __shfl_sync(0xffffffffu, minI, 0); maxI = __shfl_sync(0xffffffffu, maxI, 0);
    int mygo = -1, mygl = 0;
    if (GUESTS) { mygo = L.mem_gowner[mrow]; mygl = L.mem_glrow[mrow]; }
    unsigned roff[kERows];
    int rrank[kERows], rfam[kERows];
    T *rptr[kERows], *gptr[kERows];
#pragma unroll
    for (int r = 0; r < kERows; r++) {
        gptr[r] = nullptr;
        if (GUESTS) {
            const int go = __shfl_sync(0xffffffffu, mygo, r), gl = __shfl_sync(0xffffffffu, mygl, r);
            if (go >= 0 && r < nr) {
                unsigned long long q = (unsigned long long)(static_cast<T *>(PT.A[go]) + (int64_t)gl * ld);
                asm volatile("" : "+l"(q));
                gptr[r] = reinterpret_cast<T *>(q);
            }
        }
        rfam[r] = __shfl_sync(0xffffffffu, myfam, r);
        roff[r] = (unsigned)(rfam[r] - f0) * row_bytes;
        rrank[r] = __shfl_sync(0xffffffffu, myrank, r);
        unsigned long long p = (unsigned long long)(A + (int64_t)__shfl_sync(0xffffffffu, myslot, r) * ld);
        asm volatile("" : "+l"(p));                        // keep the pointer; do not recompute it per store
        rptr[r] = reinterpret_cast<T *>(p);
    }
    // ---- the column steps of this CTA, one per lane: aligned first couple column, 16-byte chunks
    //      per couple row, which orientations can be selected at all ----
    const int Jbeg = blockIdx.x * kEChunk, Jend = min(L.n_mtiles, Jbeg + kEChunk);
    int t_c0, t_info, t_m0, t_cnt;
    {
        const int Jl = min(Jbeg + lane, L.n_mtiles - 1);
        t_m0 = L.mt_m0[Jl]; t_cnt = L.mt_cnt[Jl];
        const int fJ0 = L.mt_fam0[Jl], nfJ = L.mt_nfam[Jl];
        t_c0 = fJ0 & ~(kVec - 1);
        const int nchunk = (fJ0 + nfJ - t_c0 + kVec - 1) / kVec;
        t_info = nchunk | (maxI > L.mt_minrank[Jl] ? 0x100 : 0)      // some row outranks some column
                        | (L.mt_maxrank[Jl] > minI ? 0x200 : 0);
    }
    const T *vsrc = V + (size_t)(f0 - L.own_f0) * L.nf_pad + lane * kVec;     // V, Vt hold own couple rows
    const T *vtsrc = Vt + (size_t)(f0 - L.own_f0) * L.nf_pad + lane * kVec;
    // column metadata of a step (couple, rank, slot of the lane's four members) goes straight into
    // registers, one step ahead: three 128-bit loads per lane instead of a round trip through shared
    // memory (the shared-memory pipe is this kernel's busiest unit)
    auto load_meta = [&](int J, int4 &g, int4 &r, int4 &sl) {
        const int mJ0 = __shfl_sync(0xffffffffu, t_m0, J - Jbeg), mJ1 = mJ0 + __shfl_sync(0xffffffffu, t_cnt, J - Jbeg);
        const int j0 = mJ0 + 4 * lane;
        if (j0 + 3 < mJ1) {
            g = __ldg(reinterpret_cast<const int4 *>(L.mem_fam + j0));
            r = __ldg(reinterpret_cast<const int4 *>(L.mem_ind + j0));
            sl = __ldg(reinterpret_cast<const int4 *>(L.mem_slot + j0));
        } else if (j0 < mJ1) {                              // ragged end of the tile: clamped
            const int m1 = min(j0 + 1, mJ1 - 1), m2 = min(j0 + 2, mJ1 - 1), m3 = min(j0 + 3, mJ1 - 1);
            g = make_int4(L.mem_fam[j0], L.mem_fam[m1], L.mem_fam[m2], L.mem_fam[m3]);
            r = make_int4(L.mem_ind[j0], L.mem_ind[m1], L.mem_ind[m2], L.mem_ind[m3]);
            sl = make_int4(L.mem_slot[j0], L.mem_slot[m1], L.mem_slot[m2], L.mem_slot[m3]);
        }
    };

    auto prefetch = [&](int J, int buf) {
        const unsigned stage = mine + (unsigned)buf * stage_bytes;
        const int c0 = __shfl_sync(0xffffffffu, t_c0, J - Jbeg), info = __shfl_sync(0xffffffffu, t_info, J - Jbeg);
        const int nchunk = info & 0xff;
        for (int c = lane; c < nchunk; c += 32) {           // one pass (two for the 8-byte type)
            unsigned dst = stage + (unsigned)c * 16u;
            const T *a = vsrc + c0 + (c - lane) * kVec, *b = vtsrc + c0 + (c - lane) * kVec;
            for (int f = 0; f < nfr; f++, dst += row_bytes, a += L.nf_pad, b += L.nf_pad) {
                if (info & 0x100) cp_async16_s(dst, a);
                if (info & 0x200) cp_async16_s(dst + vba_off, b);
            }
        }
        cp_async_commit();
    };

    prefetch(Jbeg, 0);
    int4 gj = make_int4(0, 0, 0, 0), rj4 = gj, sj4 = gj, gjn = gj, rjn = gj, sjn = gj;
    load_meta(Jbeg, gj, rj4, sj4);
    int buf = 0;
    for (int J = Jbeg; J < Jend; J++, buf ^= 1) {
        if (J + 1 < Jend) { prefetch(J + 1, buf ^ 1); load_meta(J + 1, gjn, rjn, sjn); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();                                      // other lanes' copies are visible
        const unsigned stage = mine + (unsigned)buf * stage_bytes;
        const int mJ0 = __shfl_sync(0xffffffffu, t_m0, J - Jbeg), mJ1 = mJ0 + __shfl_sync(0xffffffffu, t_cnt, J - Jbeg);
        const int j0 = mJ0 + 4 * lane;
        const int c0 = __shfl_sync(0xffffffffu, t_c0, J - Jbeg);
        if (j0 < mJ1) {
            const unsigned go[4] = {stage + (unsigned)(gj.x - c0) * (unsigned)sizeof(T), stage + (unsigned)(gj.y - c0) * (unsigned)sizeof(T),
                                    stage + (unsigned)(gj.z - c0) * (unsigned)sizeof(T), stage + (unsigned)(gj.w - c0) * (unsigned)sizeof(T)};
            const int rj[4] = {rj4.x, rj4.y, rj4.z, rj4.w}, sj[4] = {sj4.x, sj4.y, sj4.z, sj4.w};
            const int ncol = min(4, mJ1 - j0);
            const bool vec = ncol == 4 && ((sj[0] & 3) == 0) && sj[1] == sj[0] + 1 && sj[2] == sj[0] + 2 &&
                             sj[3] == sj[0] + 3;
            const int dk0 = row0 - j0;                     // the diagonal crosses this lane's columns?
            const bool diag_tile = row0 < mJ1 && row0 + kERows > mJ0;   // warp-uniform: the rows meet the columns
            if (nr == kERows && vec) {
                if (diag_tile) expand_step<T, true, true, GUESTS>(go, rj, sj, roff, rrank, rptr, gptr, vba_off, nr, ncol, dk0, Dg, rfam);
                else expand_step<T, true, false, GUESTS>(go, rj, sj, roff, rrank, rptr, gptr, vba_off, nr, ncol, dk0, Dg, rfam);
            } else {
                expand_step<T, false, true, GUESTS>(go, rj, sj, roff, rrank, rptr, gptr, vba_off, nr, ncol, dk0, Dg, rfam);
            }
        }
        __syncwarp();                                      // stage free before it is refilled
        gj = gjn; rj4 = rjn; sj4 = sjn;
    }
}

// =====================================================================================
// misfile_kernel (GENLIB_SCHEDULE_SPARSE_PHI only).  sparse_phi FILES a kinship under
// phi[earlier processed][later processed] (compute.jl:393) but every look-up reads
// phi[lower rank][higher rank] (:36-40, :350-358, :367-389): between two individuals whose queue
// order inverts their rank order the stored value is never found again and the reference goes on
// as if the pair were unrelated.  Both orders follow the depth, so this only happens between the
// members of ONE layer; the frontier keeps what a look-up finds, 0.  grid (own rows, column chunks).
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads)
misfile_kernel(T *__restrict__ A, int64_t ld, const int32_t *__restrict__ mem_rank, LayerArgs L) {
    const int i = L.own_m0 + blockIdx.x;
    const int seq_i = L.mem_ind[i], rank_i = mem_rank[i];
    T *row = A + (int64_t)L.mem_lrow[i] * ld;
    for (int j = blockIdx.y * kThreads + threadIdx.x; j < L.n_new; j += gridDim.y * kThreads)
        if (j != i && ((seq_i > L.mem_ind[j]) != (rank_i > mem_rank[j]))) row[L.mem_slot[j]] = (T)0;
}

// =====================================================================================
// proband gather (compute.jl:303: the last frontier, rows/columns in probandIDs order).
// rows[] are LOCAL rows of this rank's probands, cols[] the global slots of all probands.
// =====================================================================================
template <typename T, typename O>
__global__ void gather_kernel(const T *__restrict__ A, int64_t ld, const int32_t *__restrict__ rows,
                              const int32_t *__restrict__ cols, int32_t P, int32_t row0, int32_t nrows,
                              O *__restrict__ out) {
    if (blockIdx.y >= (unsigned)nrows) return;
    const T *row = A + (int64_t)rows[row0 + blockIdx.y] * ld;
    O *dst = out + (size_t)blockIdx.y * P;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < P; v += gridDim.x * blockDim.x)
        dst[v] = (O)row[cols[v]];
}

// sum and trace of the proband block, binary64 accumulation (phiMean, compute.jl:454-459);
// one rank: rows == cols == the proband slots
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

// =====================================================================================
// barrier_kernel: all ranks of one box meet here (one process per GPU, so every rank's
// kernel is resident on its own device).  Rank r stores `epoch` into slot r of every peer's
// flag array (release, system scope, through the NVLink mapping) and waits until all slots of
// its own array have reached `epoch`.  A rank that waits longer than ~10 s records an error
// instead of hanging.
// =====================================================================================
struct BarrierTable {
    unsigned *flags[kMaxWorld];   // [world] words on each rank; word kMaxWorld = error flag
};

__global__ void barrier_kernel(BarrierTable B, int rank, int world, unsigned epoch, long long timeout_cycles) {
    const int peer = threadIdx.x;
    if (peer >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(B.flags[peer] + rank), "r"(epoch) : "memory");
    const unsigned *mine = B.flags[rank] + peer;
    const long long t0 = clock64();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (clock64() - t0 > timeout_cycles) { B.flags[rank][kMaxWorld] = 1u; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace genlib
