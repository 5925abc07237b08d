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
//   strip     = up to 64 of this rank's couples; the strips are processed in order.
//   PRODUCER  CTAs (one per SM) read the strip's parent rows -- TMA bulk copies into a shared-memory ring that
//             runs on across strips, from local HBM or a peer's over NVLink -- and write them transposed and
//             interleaved,
//                 Q[p][F] = (Psi[f_F, p], Psi[m_F, p])          for every live column p,
//             into one of a few strip buffers that are pinned in L2 (persisting access-policy window);
//             where the step carries columns over they also write the members' rows against them.
//   CONSUMER  CTAs (the other one on every SM) take tiles of couples G (ALL couples of the layer), stage
//             Q[f_G][strip], Q[m_G][strip] -- two contiguous segments per couple, L2 hits -- which hold all
//             four entries of every (F, G) pair in BOTH groupings,
//                 a = (Psi[f_F,f_G], Psi[m_F,f_G]), c = (Psi[f_F,m_G], Psi[m_F,m_G])
//                 F climbed: hs(hs(a.x, c.x), hs(a.y, c.y))     G climbed: hs(hs(a.x, a.y), hs(c.x, c.y))
//             (hs(x, y) = 1/2 x + 1/2 y, one binary64 rounding), round ONCE to the storage type
//             (compute.jl:296) and write the strip members' rows over the tile's member columns, picking
//             the grouping by rank; the diagonal is 1/2 + 1/2 Psi[f, m] (compute.jl:148-155).  They also
//             write the strip members' columns into the rows of the carried individuals (from Q[c][strip]).
//
// Every entry of the step is written exactly once, in contiguous row segments, by the rank that owns the
// row; nothing but stored frontier rows crosses NVLink.  DRAM sees the compulsory traffic only: the parent
// rows once, the new rows once.
//
// Flow control is static: within a strip the items (producer: column tiles, consumer: member tiles and
// blocks of carried rows) are dealt round-robin to the CTAs of the role, the deal rotating from strip to
// strip so that remainders even out.  Every CTA of a role counts itself off on the strip's counter when
// its share is done; consumers start a strip when all producers have counted off, producers reuse a strip
// buffer when all consumers of the strip that used it before have.  All CTAs are resident (two per SM), and
// a wait that lasts seconds raises the layer's error word instead of hanging the device.
#pragma once
#include "kernels.cuh"

namespace genlib {

constexpr int kLayerThreads = 256;
constexpr int kLayerWarps = kLayerThreads / 32;
constexpr int kMaxStrip = 64;                   // couples per strip (upper bound of StripArgs::sw)
constexpr int kVPitch = kMaxTileFam + 1;        // row pitch of the staged couple tile (65: conflict-free)
constexpr int kMaxStages = 4;


struct StripArgs {
    int32_t sw;          // strip width: couples per strip (8, 16, 32 or 64)
    int32_t ft;          // couples per producer item (8, 16 or 32; divides sw)
    int32_t n_strips;    // strips of this rank's couples
    int32_t nbuf;        // strip buffers in rotation
    int32_t stages;      // ring stages
    int32_t n_prod;      // producer CTAs (0 when nothing is live); the grid is n_prod + n_cons
    int32_t n_cons;      // consumer CTAs
    int32_t n_pitems;    // producer items per strip: (sw / ft) * live tiles
    int32_t n_citems;    // consumer items per strip: member tiles + blocks of carried rows
    int32_t mrows;       // live-range rows per block of carried rows
    int32_t rot_p;       // rotation of the deal per strip (items % CTAs of the role)
    int32_t rot_c;
    int64_t qstride;     // pairs per strip buffer (live tiles * kPTile * sw)
    void *Q;             // strip buffers
    int32_t *sync;       // [1] error word, [2 + s] producers done with strip s, [2 + n_strips + s] consumers done
    const int32_t *live_tiles;   // the live tiles of the layer's slot range: index | kTileCarried
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

inline size_t layer_ring_bytes(int ft, int stages, size_t es) { return (size_t)stages * 2 * ft * (kPTile * es + 16); }
// consumer: staged parent-row segments of a couple tile (2 x kMaxTileFam rows x sw pairs) + Va | Vb
inline size_t layer_consumer_bytes(int sw, size_t es) {
    return (size_t)2 * kMaxTileFam * sw * 2 * es + (size_t)2 * sw * kVPitch * es;
}

// The items of strip s that CTA k of a role with n CTAs takes: first, first + n, ... below n_items.
__device__ __forceinline__ int first_item(int k, int s, int rot, int n) {
    int f = (k - (int)(((long long)s * rot) % n)) % n;
    return f < 0 ? f + n : f;
}

template <typename T, bool STORED>
__global__ void __launch_bounds__(kLayerThreads, 2)
layer_kernel(T *__restrict__ A, int64_t ld, PeerTable PT, LayerArgs L, StripArgs S) {
    using P2 = typename PairOf<T>::type;
    extern __shared__ __align__(16) unsigned char dyn_smem[];     // producer: the ring; consumer: staged segments | Va | Vb
    __shared__ const T *s_row[2 * kMaxStrip];                     // producer: parent rows of the strip being read
    __shared__ __align__(8) unsigned long long s_bar[kMaxStages]; // producer: "stage filled" mbarriers
    __shared__ int s_ready;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sw = S.sw, ft = S.ft, NS = S.n_strips;
    int *const err = S.sync + 1, *const done_p = S.sync + 2, *const done_c = S.sync + 2 + NS;
    P2 *const Qall = static_cast<P2 *>(S.Q);

    // Thread 0 spins, everybody follows.  A dependency that does not arrive in time sets the layer's error
    // word (genlib_engine_run then fails with GENLIB_ECUDA); once it is set nobody waits any more, so a
    // broken schedule drains in seconds instead of hanging the device.
    auto wait_for = [&](const int *counter, int target) {
        if (tid == 0 && target > 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(counter) < target) {
                if (clock64() - t0 > S.timeout_cycles || ld_acquire_gpu(err) != 0) { atomicExch(err, 1); break; }
                __nanosleep(100);
            }
        }
        __syncthreads();
    };
    auto count_off = [&](int *counter) {                           // this CTA's share of the strip is done
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicAdd(counter, 1); }
    };

    if ((int)blockIdx.x < S.n_prod) {
        // =============== producer: parent rows -> Q[p][F] pairs (+ member rows x carried columns) ===============
        const int k = blockIdx.x, NP = S.n_prod, NI = S.n_pitems;
        const int npt = sw / ft;                                   // items per live tile
        const int RB = kPTile * (int)sizeof(T) + 16, STAGE = 2 * ft * RB;
        const unsigned ROWB = kPTile * (unsigned)sizeof(T);
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(dyn_smem);
        const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&s_bar[0]);
        if (tid == 0) {
            for (int st = 0; st < S.stages; st++) mbar_init(bar0 + 8u * st, 2 * ft);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // two cursors over this CTA's items, strip after strip: `is/ii` is being read (TMA issued), `cs/ci` written
        int is = 0, ii = first_item(k, 0, S.rot_p, NP);
        while (is < NS && ii >= NI) { is++; ii = first_item(k, is, S.rot_p, NP); }
        int cs = is, ci = ii;
        int rows_of = -1;                                          // strip whose parent rows are in s_row
        unsigned n_issued = 0, n_done = 0;
        const int rpw = 2 * ft / kLayerWarps;                      // rows each warp issues
        const int cpw = ft / kLayerWarps;                          // couples per warp for the member rows
        const int f = lane % ft, hi = lane >> 3;
        auto row_of = [&](int s) -> const T * {                    // thread tid < 2 sw: parent row tid of strip s
            const bool mo = tid >= sw;
            const int Fl = s * sw + (mo ? tid - sw : tid);
            if (s >= NS || Fl >= L.own_nf) return nullptr;
            const int F = L.own_f0 + Fl;
            const int o = mo ? L.fam_pm_owner[F] : L.fam_pf_owner[F];
            if (o < 0) return nullptr;
            return static_cast<const T *>(PT.A[o]) + (int64_t)(mo ? L.fam_pm_lrow[F] : L.fam_pf_lrow[F]) * ld + L.rt_lo;
        };
        const T *pre_row = nullptr;                                // the same for strip pre_s, fetched a strip ahead
        int pre_s = -1;
        auto load_rows = [&](int s) {                              // all threads; the caller syncs
            if (tid < 2 * sw) {
                s_row[tid] = pre_s == s ? pre_row : row_of(s);
                pre_row = row_of(s + 1);                           // in flight while strip s is read
            }
            pre_s = s + 1;
        };
        auto issue = [&]() {                                       // item (is, ii) -> ring slot n_issued % stages
            if (lane < rpw) {
                const unsigned slot = n_issued % (unsigned)S.stages;
                const int pt = ii % npt, tile = S.live_tiles[ii / npt] & (kTileCarried - 1);
                const int row = warp * rpw + lane;                 // 0 .. 2 ft - 1: fathers, then mothers
                const unsigned bar = bar0 + 8u * slot;
                const unsigned dst = sbase + slot * (unsigned)STAGE + (unsigned)(row * RB);
                const int fi = pt * ft + (row < ft ? row : row - ft);
                const T *src = s_row[(row < ft ? 0 : sw) + fi];
                if (src) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of the stage
                    mbar_arrive_expect_tx(bar, ROWB);
                    bulk_g2s(dst, src + tile * kPTile, ROWB, bar);
                } else {                                           // unknown parent: contributes 0 (compute.jl:111-126)
                    for (unsigned c = 0; c < ROWB; c += 16) zero16_shared(dst + c);
                    mbar_arrive_expect_tx(bar, 0);
                }
            }
            n_issued++;
            ii += NP;
            while (is < NS && ii >= NI) { is++; ii = first_item(k, is, S.rot_p, NP); }
        };
        auto try_issue = [&]() {                                   // uniform over the CTA
            if (is >= NS) return;
            if (rows_of != is) { __syncthreads(); load_rows(is); rows_of = is; __syncthreads(); }
            issue();
        };
        __syncthreads();
        for (int n = 0; n < S.stages - 1; n++) try_issue();
        int cur = -1;                                              // strip this CTA is writing
        // the item being written: its tile (| kTileCarried) and the flags of the warp's 16 columns (lane & 3 holds
        // the word of column group lane & 3); both are fetched one item ahead
        auto tile_flags = [&](int tinfo) {
            return __ldg(reinterpret_cast<const uint32_t *>(L.flags + (tinfo & (kTileCarried - 1)) * kPTile) + warp * 4 + (lane & 3));
        };
        int tinfo = cs < NS ? S.live_tiles[ci / npt] : 0;
        uint32_t live4 = cs < NS ? tile_flags(tinfo) : 0u;
        while (cs < NS) {
            if (cs != cur) {                                       // count off the strips that are behind us
                for (int s = max(cur, 0); s < cs; s++) count_off(done_p + s);
                cur = cs;
                wait_for(done_c + (cs - S.nbuf), cs >= S.nbuf ? S.n_cons : 0);    // the strip that used this buffer is consumed
            }
            int ncs = cs, nci = ci + NP;                           // the item after this one
            while (ncs < NS && nci >= NI) { ncs++; nci = first_item(k, ncs, S.rot_p, NP); }
            const int tinfo_next = ncs < NS ? S.live_tiles[nci / npt] : 0;
            const unsigned slot = n_done % (unsigned)S.stages;
            if (!mbar_wait(bar0 + 8u * slot, (n_done / (unsigned)S.stages) & 1u)) atomicExch(err, 2);
            __syncthreads();                                       // everybody is done with the stage refilled next
            try_issue();
            const int pt = ci % npt, lt = ci / npt;
            const int tile = tinfo & (kTileCarried - 1);
            const unsigned char *st = dyn_smem + slot * STAGE;
            P2 *const Q = Qall + (size_t)(cs % S.nbuf) * S.qstride;
            // ---- transposed and interleaved: Q[p][F] = (father row, mother row) at column p.  Lane = couple
            //      (ft of them) x column; the column rotates with lane / 8 so that the 32 shared loads of a warp
            //      hit 32 banks (rows are padded by 16 bytes). ----
            {
                const T *xr = reinterpret_cast<const T *>(st + f * RB);
                const T *yr = reinterpret_cast<const T *>(st + (ft + f) * RB);
                P2 *q = Q + (size_t)lt * kPTile * sw + pt * ft + f;
#pragma unroll
                for (int gq = 0; gq < kPTile / 32; gq++) {
                    const uint32_t w = __shfl_sync(0xffffffffu, live4, gq);
                    for (int j = 0; j < ft / 8; j++) {
                        const int c4 = (j + hi) & 3;
                        const int col = warp * (kPTile / 8) + gq * 4 + c4;
                        if ((w >> (8 * c4)) & kFlagLive) {
                            P2 v; v.x = xr[col]; v.y = yr[col];
                            q[(size_t)col * sw] = v;
                        }
                    }
                }
            }
            // ---- rows of the new members against this tile's carried columns (rounded once, compute.jl:296).
            //      Columns that are not carried receive values nobody reads. ----
            if (tinfo & kTileCarried) {
                const int64_t col0 = (int64_t)L.rt_lo + tile * kPTile + 4 * lane;
                for (int qd = 0; qd < cpw; qd++) {
                    const int fi = warp * cpw + qd, Fl = cs * sw + pt * ft + fi;
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
            n_done++;
            cs = ncs; ci = nci;
            tinfo = tinfo_next;
            live4 = cs < NS ? tile_flags(tinfo) : 0u;
        }
        for (int s = max(cur, 0); s < NS; s++) count_off(done_p + s);
        return;
    }

    // ===== consumer: couple tiles -> the strip members' rows; carried rows <- the strip members' columns =====
    const int k = blockIdx.x - S.n_prod, NC = S.n_cons, NI = S.n_citems;
    P2 *const stg = reinterpret_cast<P2 *>(dyn_smem);                               // [2 g + parent][f]
    T *const Va = reinterpret_cast<T *>(dyn_smem + (size_t)2 * kMaxTileFam * sw * sizeof(P2));   // [f][g]: F climbed first
    T *const Vb = Va + (size_t)sw * kVPitch;                                        // [f][g]: G climbed first
    const unsigned stg_s = (unsigned)__cvta_generic_to_shared(stg);
    const unsigned va_s = (unsigned)__cvta_generic_to_shared(Va), vb_s = (unsigned)__cvta_generic_to_shared(Vb);
    const int lsw = 31 - __clz(sw);                                // sw is a power of two
    // Two threads per staged row (one parent of one couple of the tile): one look-up, then the 16-byte copies.
    const int row_bytes = sw * (int)sizeof(P2), half_bytes = row_bytes / 2;
    auto stage_tile = [&](const P2 *Q, int fJ0, int nfJ) {
        const int row = tid >> 1, part = tid & 1;
        if (row < 2 * nfJ) {
            const int G = fJ0 + (row >> 1);
            const int q = (row & 1) ? L.fam_qm[G] : L.fam_qf[G];
            const unsigned dst = stg_s + (unsigned)(row * row_bytes + part * half_bytes);
            if (q >= 0) {
                const unsigned char *src = reinterpret_cast<const unsigned char *>(Q + (size_t)q * sw) + part * half_bytes;
                for (int c = 0; c < half_bytes; c += 16) cp_async16_to(dst + c, src + c);
            } else {
                for (int c = 0; c < half_bytes; c += 16) zero16_shared(dst + c);   // unknown parent: contributes 0
            }
        }
        cp_async_commit();
    };
    // a member tile's descriptor and the lane's four member columns in it
    struct TileCtx { int fJ0, nfJ, mJ0, cntJ, gj[4], rj[4], sj[4]; };
    auto load_ctx = [&](TileCtx &c, int J) {
        c.fJ0 = L.mt_fam0[J]; c.nfJ = L.mt_nfam[J]; c.mJ0 = L.mt_m0[J]; c.cntJ = L.mt_cnt[J];
    };
    auto load_cols = [&](TileCtx &c) {
        const int j0 = c.mJ0 + 4 * lane;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = min(j0 + q, c.mJ0 + c.cntJ - 1);
            c.gj[q] = L.mem_fam[j] - c.fJ0; c.rj[q] = L.mem_ind[j]; c.sj[q] = L.mem_slot[j];
        }
    };
    int cur = -1;                                                  // strip this CTA is in
    int F0 = 0, nFs = 0, ms0 = 0, ms1 = 0, n_rows = 0, share = 1, pass_rows = kLayerWarps;
    int my_f = 0, my_rank = 0, my_lrow = 0;
    auto load_member_rows = [&](int i0) {
        const int im = min(i0 + lane, ms1 - 1);
        my_f = L.mem_fam[im] - F0; my_rank = L.mem_ind[im]; my_lrow = L.mem_lrow[im];
    };
    int s = 0, it = first_item(k, 0, S.rot_c, NC);
    while (s < NS && it >= NI) { s++; it = first_item(k, s, S.rot_c, NC); }
    bool staged = false;                                           // the current item (a tile) has been staged ahead (cx is loaded)
    TileCtx cx, nx;
    while (s < NS) {
        if (s != cur) {
            for (int t = max(cur, 0); t < s; t++) count_off(done_c + t);
            cur = s;
            F0 = L.own_f0 + s * sw;
            nFs = min(sw, L.own_nf - s * sw);
            ms0 = L.fam_start[F0]; ms1 = L.fam_start[F0 + nFs];
            n_rows = ms1 - ms0;
            // the warp's rows: an even share of the strip's member rows (at most 32 at a time), their metadata
            // spread over the lanes; a strip of up to 256 rows (the usual case) keeps them in registers
            share = max(1, min(32, (n_rows + kLayerWarps - 1) / kLayerWarps));
            pass_rows = share * kLayerWarps;
            if (n_rows > 0) load_member_rows(ms0 + warp * share);
            if (!staged) wait_for(done_p + s, S.n_prod);          // the strip's pairs are complete (in L2)
        }
        const P2 *const Q = Qall + (size_t)(s % S.nbuf) * S.qstride;
        // the item after this one (maybe in the next strip)
        int ns = s, nit = it + NC;
        while (ns < NS && nit >= NI) { ns++; nit = first_item(k, ns, S.rot_c, NC); }

        if (it >= L.n_mtiles) {
            // ---- a block of carried rows: the strip members' columns, Psi[c, i] = RN(hs(Q[c][F_i])) ----
            const int r0 = (it - L.n_mtiles) * S.mrows, r1 = min(L.rt_rows, r0 + S.mrows);
            for (int row = r0 + warp; row < r1; row += kLayerWarps) {
                if (!(L.flags[row] & kFlagCarried)) continue;
                T *dst = static_cast<T *>(PT.A[L.live_owner[row]]) + (int64_t)L.live_lrow[row] * ld;
                const P2 *q = Q + ((size_t)L.tile_map[row / kPTile] * kPTile + (size_t)(row % kPTile)) * sw;
                T v0 = (T)0, v1 = (T)0;
                if (lane < nFs) { const P2 p = __ldcg(q + lane); v0 = (T)half_sum_mode<STORED>((double)p.x, (double)p.y); }
                if (lane + 32 < nFs) { const P2 p = __ldcg(q + lane + 32); v1 = (T)half_sum_mode<STORED>((double)p.x, (double)p.y); }
                for (int m = ms0 + lane; m < ((n_rows + 31) & ~31) + ms0; m += 32) {
                    const int mm = min(m, ms1 - 1);
                    const int fi = L.mem_fam[mm] - F0;
                    const T a = __shfl_sync(0xffffffffu, v0, fi & 31), b = __shfl_sync(0xffffffffu, v1, fi & 31);
                    if (m < ms1) dst[L.mem_slot[mm]] = fi < 32 ? a : b;
                }
            }
            staged = false;
        } else {
            // ---- a member tile J: stage its couples' segments (unless done ahead), form Va | Vb, expand ----
            if (!staged) { load_ctx(cx, it); stage_tile(Q, cx.fJ0, cx.nfJ); load_cols(cx); }
            const int nfJ = cx.nfJ;
            const int j0 = cx.mJ0 + 4 * lane, ncol = min(4, cx.mJ0 + cx.cntJ - j0);
            cp_async_wait<0>();
            __syncthreads();                                       // the tile's segments are staged; the previous item is written
#pragma unroll 2
            for (int e = tid; e < (nfJ << lsw); e += kLayerThreads) {
                const int g = e >> lsw, fl = e & (sw - 1);
                if (fl < nFs) {
                    const P2 a = stg[(2 * g) * sw + fl], c = stg[(2 * g + 1) * sw + fl];
                    T vf, vg;
                    couple_pair<T, STORED>((double)a.x, (double)a.y, (double)c.x, (double)c.y, vf, vg);
                    Va[fl * kVPitch + g] = vf;
                    Vb[fl * kVPitch + g] = vg;
                }
            }
            __syncthreads();                                       // Va | Vb complete, the staging area is free
            // stage the next tile, and fetch its columns, while this one is expanded (same strip, or the
            // next one if its producers are done)
            staged = false;
            if (ns < NS && nit < L.n_mtiles) {
                bool ready = ns == s;
                if (!ready) {                                      // peek: do not wait here, the expansion comes first
                    if (tid == 0) s_ready = ld_acquire_gpu(done_p + ns) >= S.n_prod;
                    __syncthreads();
                    ready = s_ready != 0;
                }
                if (ready) {
                    load_ctx(nx, nit);
                    stage_tile(Qall + (size_t)(ns % S.nbuf) * S.qstride, nx.fJ0, nx.nfJ);
                    load_cols(nx);
                    staged = true;
                }
            }
            int gj[4], rj[4], sj[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { gj[q] = cx.gj[q]; rj[q] = cx.rj[q]; sj[q] = cx.sj[q]; }
            const bool vec = ncol == 4 && ((sj[0] & 3) == 0) && sj[1] == sj[0] + 1 && sj[2] == sj[0] + 2 && sj[3] == sj[0] + 3;
            unsigned off[4];                                       // the lane's columns inside a couple-tile row
#pragma unroll
            for (int q = 0; q < 4; q++) off[q] = (unsigned)(gj[q] * (int)sizeof(T));
            for (int i0 = ms0 + warp * share; i0 < ms1; i0 += pass_rows) {
                if (n_rows > pass_rows) load_member_rows(i0);
                const int nrow = min(share, ms1 - i0);
                const int dk0 = i0 - j0;                           // row rr meets the lane's column q when dk0 + rr == q
#pragma unroll 2
                for (int rr = 0; rr < nrow; rr++) {                // (every lane takes part in the shuffles)
                    const int fi = __shfl_sync(0xffffffffu, my_f, rr), ri = __shfl_sync(0xffffffffu, my_rank, rr);
                    const int lrow = __shfl_sync(0xffffffffu, my_lrow, rr);
                    if (ncol <= 0) continue;
                    const unsigned rowoff = (unsigned)(fi * kVPitch * (int)sizeof(T));
                    T v[4];
#pragma unroll
                    for (int q = 0; q < 4; q++)                    // the higher rank is climbed first (compute.jl:130-147)
                        v[q] = lds<T>((ri > rj[q] ? va_s : vb_s) + rowoff + off[q]);
                    if ((unsigned)(dk0 + rr) < 4u) {               // own diagonal entry (compute.jl:148-155)
                        const int F = F0 + fi, pf = L.fam_pf[F], pm = L.fam_pm[F];
                        double d = 0.5;
                        if (pf >= 0 && pm >= 0)
                            d = half_sum_mode<STORED>((double)(static_cast<const T *>(PT.A[L.fam_pf_owner[F]]) + (int64_t)L.fam_pf_lrow[F] * ld)[pm], 1.0);
#pragma unroll
                        for (int q = 0; q < 4; q++) if (dk0 + rr == q) v[q] = (T)d;
                    }
                    T *row = A + (int64_t)lrow * ld;
                    if (vec) store_vec4(row + sj[0], v);
                    else {
#pragma unroll
                        for (int q = 0; q < 4; q++) if (q < ncol) row[sj[q]] = v[q];
                    }
                }
            }
        }
        if (staged) cx = nx;
        s = ns; it = nit;
    }
    for (int t = max(cur, 0); t < NS; t++) count_off(done_c + t);
}

}  // namespace genlib
