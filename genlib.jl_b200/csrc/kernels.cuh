// kernels.cuh -- device-side vocabulary shared by the sm_100a kernels of the kinship sweep, and the
// small kernels around the layer kernel (layer_kernel.cuh): proband gather, phiMean reduction, the
// inter-GPU barrier and sparse_phi's misfiled pairs.
//
// Arithmetic is binary64 in the reference's grouping; storage type T is float
// (GENLIB_NUMERICS_REFERENCE: one RN32 per step, like src/compute.jl:296) or double.
// STORED = true selects the arithmetic of gen.sparse_phi instead (compute.jl:321-447: every
// stored kinship is a Float32 and is halved in Float32; the plan then orders the layers and the
// members by sparse_phi's queue, see plan.hpp).  No -use_fast_math / -ftz: Float32 subnormals occur.
#pragma once
#include <cuda_runtime.h>
#include <climits>
#include <cstdint>

#include "plan.hpp"

namespace genlib {

constexpr int kMaxWorld = 16;

// Where every rank keeps its frontier rows.  Entries of other ranks are NVLink peer memory: CUDA-IPC
// mappings of their arenas (one process per GPU) or plain peer pointers (one process, several devices).
struct PeerTable {
    void *A[kMaxWorld];       // rows_cap[g] x ld frontier rows of rank g
};

struct LayerArgs {
    int32_t n_new, n_fam, rt_lo, rt_rows, nf_pad, any_carried;
    // row sharding: this rank owns couples [own_f0, own_f0 + own_nf) and members [own_m0, own_m0 + own_nm)
    int32_t rank, world, own_f0, own_nf, own_m0, own_nm;
    int32_t fam_base[kMaxWorld + 1];
    const int32_t *mem_ind, *mem_slot, *mem_fam, *mem_lrow;     // members: rank (or queue position), column slot, couple, local row
    const int32_t *fam_pf, *fam_pm, *fam_start;                  // couples: parent slots (-1 none), first member
    const int2 *fam_q;                                           // couples: the parents' rows in a strip buffer (-1 none)
    const int8_t *fam_pf_owner, *fam_pm_owner, *live_owner;      // where the parents' / the live individuals' rows are
    const int32_t *fam_pf_lrow, *fam_pm_lrow, *live_lrow;
    const uint8_t *flags;                                        // per slot of the live range: kFlagLive | kFlagCarried
    const int32_t *tile_map;                                     // per tile of the live range: index among the live tiles (-1: hole)
    const int4 *mt_desc;                                         // member tiles: first couple, couples, first member, members
    int32_t n_mtiles;
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
// New frontier rows are written once and read by the NEXT step's kernel at the earliest: streaming stores
// (st.global.cs, evict-first) keep them from pushing the strip buffers out of L2.
__device__ __forceinline__ void store4(float *p, const double (&d)[4], int) {
    __stcs(reinterpret_cast<float4 *>(p), make_float4((float)d[0], (float)d[1], (float)d[2], (float)d[3]));
}
__device__ __forceinline__ void store4(double *p, const double (&d)[4], int) {
    __stcs(reinterpret_cast<double2 *>(p), make_double2(d[0], d[1]));
    __stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(d[2], d[3]));
}
__device__ __forceinline__ void store_vec4(float *p, const float (&v)[4]) {
    __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
}
__device__ __forceinline__ void store_vec4(double *p, const double (&v)[4]) {
    __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1]));
    __stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(v[2], v[3]));
}
// an L2 cache policy for data that is read once (the parents' rows): evict first
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
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

constexpr int kMaxPChunk = 32;               // live-column tiles per producer unit (upper bound)

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
// 16 bytes global -> shared, cached in L2 only (coherent with what other SMs wrote before they signalled)
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
// returns false when the phase did not complete within ~2^26 polls (a lost bulk copy must not hang the device)
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    for (unsigned spins = 0; spins < (1u << 26); spins++) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16), completes on `bar`
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(unsigned dst, const void *src, unsigned bytes, unsigned bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
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
template <typename T>
__device__ __forceinline__ T lds(unsigned addr) {
    T v;
    if constexpr (sizeof(T) == 4) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    else asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
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

// The same for ANY proband row of a sharded frontier: row u lives on rank owner[u] at local row lrow[u]
// (a peer read over NVLink unless it is this rank's own).  One process, several devices: every device
// assembles a contiguous block of output rows and copies it to the host over its own PCIe link.
template <typename T, typename O>
__global__ void gather_peer_kernel(PeerTable PT, int64_t ld, const int8_t *__restrict__ owner, const int32_t *__restrict__ lrow,
                                   const int32_t *__restrict__ cols, int32_t P, int32_t row0, int32_t nrows,
                                   O *__restrict__ out) {
    if (blockIdx.y >= (unsigned)nrows) return;
    const int u = row0 + blockIdx.y;
    const T *row = static_cast<const T *>(PT.A[owner[u]]) + (int64_t)lrow[u] * ld;
    O *dst = out + (size_t)blockIdx.y * P;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < P; v += gridDim.x * blockDim.x)
        dst[v] = (O)row[cols[v]];
}

// Row sums of the proband block for phiMean (src/compute.jl:454-459), deterministic: one CTA per own
// proband row, every thread adds its columns in ascending order in binary64, then a fixed shuffle /
// shared-memory tree.  out[2 r] = sum over all proband columns of row r, out[2 r + 1] = its diagonal entry.
// The host adds the rows in proband order, so the mean does not depend on the number of ranks.
template <typename T>
__global__ void __launch_bounds__(kThreads)
rowsum_kernel(const T *__restrict__ A, int64_t ld, const int32_t *__restrict__ own_rows, const int32_t *__restrict__ own_index,
              const int32_t *__restrict__ slots, int32_t P, double *__restrict__ out) {
    __shared__ double part[kThreads / 32];
    const int r = blockIdx.x;
    const T *row = A + (int64_t)own_rows[r] * ld;
    double s = 0.0;
    for (int v = threadIdx.x; v < P; v += kThreads) s += (double)row[slots[v]];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; w++) s += part[w];
        out[2 * r] = s;
        out[2 * r + 1] = (double)row[slots[own_index[r]]];
    }
}

// =====================================================================================
// barrier_kernel: all ranks of one box meet here (one process per GPU, so every rank's
// kernel is resident on its own device).  Rank r stores `epoch` into slot r of every peer's
// flag array (release, system scope, through the NVLink mapping) and waits until all slots of
// its own array have reached `epoch`.  A rank that waits longer than ~10 s (GENLIB_BARRIER_TIMEOUT_S)
// records an error in every rank's flag block instead of hanging: all ranks' genlib_engine_run fail.
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
        if (clock64() - t0 > timeout_cycles) {               // tell EVERY rank: the peers that arrive later pass this barrier,
            for (int g = 0; g < world; g++)                   // but their genlib_engine_run must fail as well
                asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(B.flags[g] + kMaxWorld), "r"(1u) : "memory");
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace genlib
