// layer_kernel.cuh -- ONE persistent sm_100a kernel per generation step.
//
// A step (one cut-vertex step of the reference, src/compute.jl:276-302) turns the frontier Psi into
//   phi[i, j] = 1/2 (1/2 Psi[f_hi, f_lo] + 1/2 Psi[f_hi, m_lo]) + 1/2 (1/2 Psi[m_hi, f_lo] + 1/2 Psi[m_hi, m_lo])
// for two new individuals (hi = the one with the larger rank is climbed first, compute.jl:130-147) and
//   phi[i, c] = 1/2 Psi[f_i, c] + 1/2 Psi[m_i, c]    against a carried individual c (compute.jl:111-126).
// Full siblings share everything but the diagonal, so the work is done per COUPLE (F = (f, m)).
//
// The four frontier entries of a couple pair sit at the crossing of two parent ROWS (contiguous) and two
// parent COLUMNS (scattered): the step is a gather along both axes.  It is done as two coalesced passes
// through a transposed scratch -- but the scratch never leaves the 126 MB L2:
//
//   strip     = up to 64 of this rank's couples.
//   PRODUCER  reads the strip's parent rows (TMA bulk copies into a shared-memory ring, local HBM or a
//             peer's over NVLink) and writes them transposed and interleaved,
//                 Q[p][F] = (Psi[f_F, p], Psi[m_F, p])          for every live column p,
//             into one of a few strip buffers that are pinned in L2 (persisting access-policy window);
//             where the step carries columns over it also writes the members' rows against them.
//   CONSUMER  for a tile of couples G (ALL couples of the layer) reads Q[f_G][strip], Q[m_G][strip] -- two
//             contiguous segments, L2 hits -- which hold all four entries of every (F, G) pair in BOTH
//             groupings,
//                 a = (Psi[f_F,f_G], Psi[m_F,f_G]), c = (Psi[f_F,m_G], Psi[m_F,m_G])
//                 F climbed: hs(hs(a.x, c.x), hs(a.y, c.y))     G climbed: hs(hs(a.x, a.y), hs(c.x, c.y))
//             (hs(x, y) = 1/2 x + 1/2 y, one binary64 rounding), rounds ONCE to the storage type
//             (compute.jl:296) and writes the strip members' rows over the tile's member columns, picking
//             the grouping by rank; the diagonal is 1/2 + 1/2 Psi[f, m] (compute.jl:148-155).
//   MIRROR    the strip members' columns in the rows of the carried individuals, from Q[c][strip].
//
// Every entry of the step is written exactly once, in contiguous row segments, by the rank that owns the
// row; nothing but stored frontier rows crosses NVLink.  DRAM sees the compulsory traffic only: the parent
// rows once, the new rows once.  Units (producer chunks, consumer tile groups, mirror row blocks) are
// handed out in one global order by an atomic counter; a consumer unit waits for its strip's producers,
// a producer for the consumers of the strip that used its buffer before -- always units handed out
// earlier, so the kernel cannot deadlock whatever the number of resident CTAs.
#pragma once
#include "kernels.cuh"

namespace genlib {

constexpr int kLayerThreads = 256;
constexpr int kMaxStrip = 64;                   // couples per strip (upper bound of StripArgs::sw)
constexpr int kVPitch = kMaxTileFam + 1;        // row pitch of the staged couple tile (65: conflict-free)
constexpr int kMaxStages = 4;

struct StripArgs {
    int32_t sw;          // strip width: couples per strip (8, 16, 32 or 64)
    int32_t ft;          // couples per producer tile (8, 16 or 32; divides sw)
    int32_t n_strips;    // strips of this rank's couples
    int32_t pchunk;      // live-column tiles per producer unit (<= kMaxPChunk)
    int32_t n_pchunks;   // producer column chunks per tile row; 0 when nothing is live
    int32_t gt;          // member tiles per consumer unit
    int32_t n_cunits;    // consumer units per strip
    int32_t mrows;       // live-range rows per mirror unit
    int32_t n_munits;    // mirror units per strip (0 when nothing is carried)
    int32_t nbuf;        // strip buffers in rotation
    int32_t stages;      // ring stages
    int32_t n_units;     // all units of the layer
    int64_t qstride;     // pairs per strip buffer (live tiles * kPTile * sw)
    void *Q;             // strip buffers
    int32_t *sync;       // [0] next unit, [1] error, [2 + s] producer units done, [2 + n_strips + s] consumer units done
    long long timeout_cycles;
};

template <typename T> struct PairOf;
template <> struct PairOf<float> { using type = float2; };
template <> struct PairOf<double> { using type = double2; };

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// the two groupings of the four frontier entries of a couple pair (see the header)
template <typename T, bool STORED>
__device__ __forceinline__ void couple_pair(double ax, double ay, double cx, double cy, T &f_climbed, T &g_climbed) {
    if constexpr (STORED) {      // sparse_phi: every intermediate kinship is a stored Float32 (compute.jl:331, 363-395)
        f_climbed = (T)half_sum_stored((double)(T)half_sum_stored(ax, cx), (double)(T)half_sum_stored(ay, cy));
        g_climbed = (T)half_sum_stored((double)(T)half_sum_stored(ax, ay), (double)(T)half_sum_stored(cx, cy));
    } else {
        f_climbed = (T)half_sum(half_sum(ax, cx), half_sum(ay, cy));
        g_climbed = (T)half_sum(half_sum(ax, ay), half_sum(cx, cy));
    }
}

template <typename T> constexpr size_t layer_ring_bytes(int ft, int stages) {
    return (size_t)stages * 2 * ft * (kPTile * sizeof(T) + 16);
}
// consumer: staged parent-row segments of a couple tile (2 x kMaxTileFam rows x sw pairs) + Va | Vb
inline size_t layer_consumer_bytes(int sw, size_t es) {
    return (size_t)2 * kMaxTileFam * sw * 2 * es + (size_t)2 * sw * kVPitch * es;
}

template <typename T, bool STORED>
__global__ void __launch_bounds__(kLayerThreads, 2)
layer_kernel(T *__restrict__ A, int64_t ld, PeerTable PT, LayerArgs L, StripArgs S) {
    using P2 = typename PairOf<T>::type;
    extern __shared__ __align__(16) unsigned char dyn_smem[];     // producer: the ring; consumer: staged segments | Va | Vb
    __shared__ const T *s_row[2 * kFTile];                        // producer: parent rows at the chunk's first column
    __shared__ __align__(16) uint8_t s_flag[kMaxPChunk * kPTile]; // producer: column flags of the chunk
    __shared__ int s_tile[kMaxPChunk];                            // producer: live tiles, index | carried << 8
    __shared__ int s_tq[kMaxPChunk];                              // producer: their rows in the strip buffer
    __shared__ int s_ntile;
    __shared__ __align__(8) unsigned long long s_bar[kMaxStages]; // "stage filled" mbarriers
    __shared__ int s_unit;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sw = S.sw, ft = S.ft, NS = S.n_strips;
    const int RB = kPTile * (int)sizeof(T) + 16, STAGE = 2 * ft * RB;
    const unsigned ROWB = kPTile * (unsigned)sizeof(T);
    const int nP = (sw / ft) * S.n_pchunks, nCM = S.n_cunits + S.n_munits;
    const int look = S.nbuf - 1;                                   // strips the producers run ahead
    int *const done_p = S.sync + 2, *const done_c = S.sync + 2 + NS;
    P2 *const Qall = static_cast<P2 *>(S.Q);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(dyn_smem);
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&s_bar[0]);

    if (tid == 0) {
        for (int st = 0; st < S.stages; st++) mbar_init(bar0 + 8u * st, 2 * ft);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned kk = 0;                                               // tiles this CTA has sent through its ring
    int next = 0;
    if (tid == 0) next = atomicAdd(S.sync, 1);

    auto wait_for = [&](const int *counter, int target) {          // thread 0 spins, everybody follows
        if (tid == 0 && target > 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(counter) < target) {
                if (clock64() - t0 > S.timeout_cycles) { atomicExch(S.sync + 1, 1); break; }
                __nanosleep(100);
            }
        }
        __syncthreads();
    };

    for (;;) {
        __syncthreads();                                           // the previous unit is done with shared memory
        if (tid == 0) s_unit = next;
        __syncthreads();
        const int u = s_unit;
        if (u >= S.n_units) break;
        if (tid == 0) next = atomicAdd(S.sync, 1);                 // fetched while this unit runs
        // ---- which unit: P(0 .. look-1), then C(s) followed by P(s + look) ----
        int s, r;
        bool producer;
        {
            const int pro = min(look, NS) * nP;
            if (u < pro) { producer = true; s = u / nP; r = u - s * nP; }
            else {
                int v = u - pro;
                const int B = nCM + nP, full = max(0, NS - look);
                if (v < full * B) {
                    s = v / B; r = v - s * B;
                    producer = r >= nCM;
                    if (producer) { s += look; r -= nCM; }
                } else {
                    v -= full * B;
                    s = full + v / nCM; r = v % nCM; producer = false;
                }
            }
        }
        const int F0l = s * sw;                                    // first couple of the strip, local to this rank
        const int nFs = min(sw, L.own_nf - F0l);
        P2 *const Q = Qall + (size_t)(s % S.nbuf) * S.qstride;

        if (producer) {
            // ================= producer: parent rows -> Q[p][F] pairs (+ member rows x carried columns) =================
            const int pt = r / S.n_pchunks, chunk = r - pt * S.n_pchunks;
            const int Fp0 = F0l + pt * ft;                         // first couple of this tile (local)
            const int t0 = chunk * S.pchunk;
            const int nt_all = min(S.pchunk, L.rt_rows / kPTile - t0);
            const int c0 = t0 * kPTile;
            {
                const uint4 *src = reinterpret_cast<const uint4 *>(L.flags + c0);
                uint4 *dst = reinterpret_cast<uint4 *>(s_flag);
                for (int i = tid; i < nt_all * (kPTile / 16); i += kLayerThreads) dst[i] = __ldg(src + i);
            }
            if (tid < 2 * ft) {
                const bool mo = tid >= ft;
                const int Fl = Fp0 + (mo ? tid - ft : tid);
                const T *row = nullptr;
                if (Fl < L.own_nf) {
                    const int F = L.own_f0 + Fl;
                    const int o = mo ? L.fam_pm_owner[F] : L.fam_pf_owner[F];
                    if (o >= 0) row = static_cast<const T *>(PT.A[o]) + (int64_t)(mo ? L.fam_pm_lrow[F] : L.fam_pf_lrow[F]) * ld + L.rt_lo + c0;
                }
                s_row[tid] = row;
            }
            __syncthreads();
            if (warp == 0) {                                       // holes of a fragmented slot range are skipped
                int info = 0;
                if (lane < nt_all) {
                    const uint32_t *w = reinterpret_cast<const uint32_t *>(s_flag + lane * kPTile);
                    uint32_t acc = 0;
                    for (int k = 0; k < kPTile / 4; k++) acc |= w[(k + lane) & (kPTile / 4 - 1)];
                    info = ((acc & 0x01010101u * kFlagLive) ? 1 : 0) | ((acc & 0x01010101u * kFlagCarried) ? 0x100 : 0);
                }
                const unsigned m = __ballot_sync(0xffffffffu, info & 1);
                if (info & 1) {
                    const int at = __popc(m & ((1u << lane) - 1u));
                    s_tile[at] = lane | (info & 0x100);
                    s_tq[at] = L.tile_map[t0 + lane] * kPTile;
                }
                if (lane == 0) s_ntile = Fp0 < L.own_nf ? __popc(m) : 0;     // a tile of couples past the last one: nothing to do
            }
            wait_for(done_c + (s - S.nbuf), s >= S.nbuf ? nCM : 0);    // the strip that used this buffer is consumed
            const int nt = s_ntile;
            const int rpw = 2 * ft / (kLayerThreads / 32);         // rows each warp issues
            auto issue = [&](int k) {                              // k-th live tile of the unit -> ring
                if (lane < rpw) {
                    const unsigned slot = (kk + (unsigned)k) % (unsigned)S.stages;
                    const int row = warp * rpw + lane;
                    const int ti = s_tile[k] & 0xff;
                    const unsigned bar = bar0 + 8u * slot;
                    const unsigned dst = sbase + slot * (unsigned)STAGE + (unsigned)(row * RB);
                    const T *src = s_row[row];
                    if (src) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of the stage
                        mbar_arrive_expect_tx(bar, ROWB);
                        bulk_g2s(dst, src + ti * kPTile, ROWB, bar);
                    } else {                                       // unknown parent: contributes 0 (compute.jl:111-126)
                        for (unsigned c = 0; c < ROWB; c += 16) zero16_shared(dst + c);
                        mbar_arrive_expect_tx(bar, 0);
                    }
                }
            };
            for (int k = 0; k < S.stages - 1; k++)
                if (k < nt) issue(k);
            const int cpw = ft / (kLayerThreads / 32);             // couples per warp for the member rows
            const int f = lane % ft, hi = lane >> 3;
            for (int k = 0; k < nt; k++) {
                const unsigned g = kk + (unsigned)k, slot = g % (unsigned)S.stages;
                mbar_wait(bar0 + 8u * slot, (g / (unsigned)S.stages) & 1u);
                __syncthreads();                                   // everybody is done with the stage refilled next
                if (k + S.stages - 1 < nt) issue(k + S.stages - 1);
                const int tinfo = s_tile[k], ti = tinfo & 0xff;
                const unsigned char *st = dyn_smem + slot * STAGE;
                const uint8_t *fl = s_flag + ti * kPTile;
                // ---- transposed and interleaved: Q[p][F] = (father row, mother row) at column p.  Lane = couple
                //      (ft of them) x column; the column rotates with lane / 8 so that the 32 shared loads of a warp
                //      hit 32 banks (rows are padded by 16 bytes). ----
                {
                    const T *xr = reinterpret_cast<const T *>(st + f * RB);
                    const T *yr = reinterpret_cast<const T *>(st + (ft + f) * RB);
                    P2 *q = Q + (size_t)s_tq[k] * sw + pt * ft + f;
                    for (int gq = 0; gq < kPTile / 32; gq++) {
                        for (int j = 0; j < ft / 8; j++) {
                            const int col = warp * (kPTile / 8) + gq * 4 + ((j + hi) & 3);
                            if (fl[col] & kFlagLive) {
                                P2 v; v.x = xr[col]; v.y = yr[col];
                                q[(size_t)col * sw] = v;
                            }
                        }
                    }
                }
                // ---- rows of the new members against this tile's carried columns (rounded once, compute.jl:296).
                //      Columns that are not carried receive values nobody reads. ----
                if (tinfo & 0x100) {
                    const int64_t col0 = (int64_t)L.rt_lo + c0 + ti * kPTile + 4 * lane;
                    for (int qd = 0; qd < cpw; qd++) {
                        const int fi = warp * cpw + qd, Fl = Fp0 + fi;
                        if (Fl >= L.own_nf) continue;
                        const int mb = L.fam_start[L.own_f0 + Fl], me = L.fam_start[L.own_f0 + Fl + 1];
                        if (me <= mb) continue;
                        double x[4], y[4], rr[4];
                        lds4(reinterpret_cast<const T *>(st + fi * RB) + 4 * lane, x);
                        lds4(reinterpret_cast<const T *>(st + (ft + fi) * RB) + 4 * lane, y);
#pragma unroll
                        for (int e = 0; e < 4; e++) rr[e] = half_sum_mode<STORED>(x[e], y[e]);
                        for (int m = mb; m < me; m++) store4(A + (int64_t)L.mem_lrow[m] * ld + col0, rr);
                    }
                }
            }
            kk += (unsigned)nt;
            __syncthreads();
            if (tid == 0) { __threadfence(); atomicAdd(done_p + s, 1); }
            continue;
        }

        // strip members (rows of this unit): [ms0, ms1) in the layer's member order
        const int F0 = L.own_f0 + F0l;
        const int ms0 = L.fam_start[F0], ms1 = L.fam_start[F0 + nFs];
        wait_for(done_p + s, nP);                                  // the strip's pairs are complete (in L2)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the ring's bulk copies are behind us

        if (r >= S.n_cunits) {
            // ================= mirror: the strip members' columns in the carried rows =================
            const int r0 = (r - S.n_cunits) * S.mrows, r1 = min(L.rt_rows, r0 + S.mrows);
            for (int row = r0 + warp; row < r1; row += kLayerThreads / 32) {
                if (!(L.flags[row] & kFlagCarried)) continue;
                T *dst = static_cast<T *>(PT.A[L.live_owner[row]]) + (int64_t)L.live_lrow[row] * ld;
                const P2 *q = Q + ((size_t)L.tile_map[row / kPTile] * kPTile + (size_t)(row % kPTile)) * sw;
                T v0 = (T)0, v1 = (T)0;
                if (lane < nFs) { const P2 p = __ldcg(q + lane); v0 = (T)half_sum_mode<STORED>((double)p.x, (double)p.y); }
                if (lane + 32 < nFs) { const P2 p = __ldcg(q + lane + 32); v1 = (T)half_sum_mode<STORED>((double)p.x, (double)p.y); }
                for (int m = ms0 + lane; m < ((ms1 - ms0 + 31) & ~31) + ms0; m += 32) {
                    const int mm = min(m, ms1 - 1);
                    const int fi = L.mem_fam[mm] - F0;
                    const T a = __shfl_sync(0xffffffffu, v0, fi & 31), b = __shfl_sync(0xffffffffu, v1, fi & 31);
                    if (m < ms1) dst[L.mem_slot[mm]] = fi < 32 ? a : b;
                }
            }
            __syncthreads();
            if (tid == 0) { __threadfence(); atomicAdd(done_c + s, 1); }
            continue;
        }

        // ================= consumer: couple tiles -> the strip members' rows =================
        // Per member tile J: the two parent-row segments Q[f_G][strip], Q[m_G][strip] of its couples are
        // staged in shared memory with 16-byte cp.async (issued while the previous tile is expanded, so
        // the L2 latency is hidden), both groupings of every (F, G) pair go to Va | Vb, and the warps
        // write the strip members' rows over the tile's member columns.
        P2 *const stg = reinterpret_cast<P2 *>(dyn_smem);                               // [2 g + parent][f]
        T *const Va = reinterpret_cast<T *>(dyn_smem + (size_t)2 * kMaxTileFam * sw * sizeof(P2));   // [f][g]: F climbed first
        T *const Vb = Va + (size_t)sw * kVPitch;                                        // [f][g]: G climbed first
        const unsigned stg_s = (unsigned)__cvta_generic_to_shared(stg);
        const unsigned va_s = (unsigned)__cvta_generic_to_shared(Va), vb_s = (unsigned)__cvta_generic_to_shared(Vb);
        const int J0 = r * S.gt, J1 = min(L.n_mtiles, J0 + S.gt);
        const int cpr = sw * (int)sizeof(P2) / 16;                 // 16-byte chunks per staged row
        auto stage_tile = [&](int J) {
            const int fJ0 = L.mt_fam0[J], nfJ = L.mt_nfam[J];
            for (int c = tid; c < 2 * nfJ * cpr; c += kLayerThreads) {
                const int row = c / cpr, part = c - row * cpr;
                const int G = fJ0 + (row >> 1);
                const int p = (row & 1) ? L.fam_pm[G] : L.fam_pf[G];
                const unsigned dst = stg_s + (unsigned)(row * sw * (int)sizeof(P2) + part * 16);
                if (p >= 0) {
                    const int rel = p - L.rt_lo;
                    const P2 *src = Q + ((size_t)L.tile_map[rel / kPTile] * kPTile + (size_t)(rel % kPTile)) * sw;
                    cp_async16_to(dst, reinterpret_cast<const unsigned char *>(src) + part * 16);
                } else zero16_shared(dst);                         // unknown parent: contributes 0
            }
            cp_async_commit();
        };
        // the warp's rows: 32 at a time, their metadata spread over the lanes; a strip of up to 256 rows
        // (the usual case) keeps them in registers for the whole unit
        const int n_rows = ms1 - ms0;
        int my_f = 0, my_rank = 0, my_lrow = 0;
        auto load_rows = [&](int i0) {
            const int im = min(i0 + lane, ms1 - 1);
            my_f = L.mem_fam[im] - F0; my_rank = L.mem_ind[im]; my_lrow = L.mem_lrow[im];
        };
        if (n_rows > 0) load_rows(ms0 + warp * 32);
        stage_tile(J0);
        for (int J = J0; J < J1; J++) {
            const int fJ0 = L.mt_fam0[J], nfJ = L.mt_nfam[J], mJ0 = L.mt_m0[J], cntJ = L.mt_cnt[J];
            // the lane's four member columns of this tile (loaded before the wait: the latencies overlap)
            const int j0 = mJ0 + 4 * lane, ncol = min(4, mJ0 + cntJ - j0);
            int gj[4], rj[4], sj[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int j = min(j0 + k, mJ0 + cntJ - 1);
                gj[k] = L.mem_fam[j] - fJ0; rj[k] = L.mem_ind[j]; sj[k] = L.mem_slot[j];
            }
            cp_async_wait<0>();
            __syncthreads();                                       // the tile's segments are staged; the previous tile is expanded
            for (int e = tid; e < nfJ * nFs; e += kLayerThreads) {
                const int g = e / nFs, fl = e - g * nFs;
                const P2 a = stg[(2 * g) * sw + fl], c = stg[(2 * g + 1) * sw + fl];
                T vf, vg;
                couple_pair<T, STORED>((double)a.x, (double)a.y, (double)c.x, (double)c.y, vf, vg);
                Va[fl * kVPitch + g] = vf;
                Vb[fl * kVPitch + g] = vg;
            }
            __syncthreads();                                       // Va | Vb complete, the staging area is free
            if (J + 1 < J1) stage_tile(J + 1);
            const bool vec = ncol == 4 && ((sj[0] & 3) == 0) && sj[1] == sj[0] + 1 && sj[2] == sj[0] + 2 && sj[3] == sj[0] + 3;
            for (int i0 = ms0 + warp * 32; i0 < ms1; i0 += kLayerThreads) {
                if (n_rows > kLayerThreads) load_rows(i0);
                const int nrow = min(32, ms1 - i0);
                for (int rr = 0; rr < nrow; rr++) {
                    const int fi = __shfl_sync(0xffffffffu, my_f, rr), ri = __shfl_sync(0xffffffffu, my_rank, rr);
                    const int lrow = __shfl_sync(0xffffffffu, my_lrow, rr);
                    if (ncol <= 0) continue;
                    T v[4];
#pragma unroll
                    for (int k = 0; k < 4; k++)                    // the higher rank is climbed first (compute.jl:130-147)
                        v[k] = lds<T>((ri > rj[k] ? va_s : vb_s) + (unsigned)((fi * kVPitch + gj[k]) * (int)sizeof(T)));
                    const int dk = i0 + rr - j0;
                    if ((unsigned)dk < 4u) {                       // own diagonal entry (compute.jl:148-155)
                        const int F = F0 + fi, pf = L.fam_pf[F], pm = L.fam_pm[F];
                        double d = 0.5;
                        if (pf >= 0 && pm >= 0)
                            d = half_sum_mode<STORED>((double)(static_cast<const T *>(PT.A[L.fam_pf_owner[F]]) + (int64_t)L.fam_pf_lrow[F] * ld)[pm], 1.0);
#pragma unroll
                        for (int k = 0; k < 4; k++) if (dk == k) v[k] = (T)d;
                    }
                    T *row = A + (int64_t)lrow * ld;
                    if (vec) store_vec4(row + sj[0], v);
                    else {
#pragma unroll
                        for (int k = 0; k < 4; k++) if (k < ncol) row[sj[k]] = v[k];
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicAdd(done_c + s, 1); }
    }
}

}  // namespace genlib
