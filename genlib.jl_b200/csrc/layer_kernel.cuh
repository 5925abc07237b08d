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
#if defined(GENLIB_CHECK)
#include <cassert>
#define CHECK(cond) assert(cond)
#elif defined(GENLIB_SOFTCHECK)       // record the source line of a violated bound in the layer's error word, go on
#define CHECK(cond) do { if (!(cond)) atomicMax(S.sync + 1, 100000 + __LINE__); } while (0)
#else
#define CHECK(cond) do { } while (0)
#endif

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
    long long *prof;             // -DGENLIB_PROFILE: 8 cycle counters per CTA (else unused)
};

template <typename T> struct PairOf;
template <> struct PairOf<float> { using type = float2; };
template <> struct PairOf<double> { using type = double2; };

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The two groupings of the four frontier entries of a couple pair (see the header).  hs(x, y) = RN(x/2 + y/2)
// = RN(x + y) / 2 (scaling by a power of two commutes with rounding: binary64 cannot underflow here), so
// hs(hs(a, b), hs(c, d)) = RN(RN(a + b) + RN(c + d)) / 4 bit for bit: three additions and one exact product.
template <typename T, bool STORED>
__device__ __forceinline__ void couple_pair(double ax, double ay, double cx, double cy, T &f_climbed, T &g_climbed) {
    if constexpr (STORED) {      // sparse_phi: every intermediate kinship is a stored Float32 (compute.jl:331, 363-395)
        f_climbed = (T)half_sum_stored((double)(T)half_sum_stored(ax, cx), (double)(T)half_sum_stored(ay, cy));
        g_climbed = (T)half_sum_stored((double)(T)half_sum_stored(ax, ay), (double)(T)half_sum_stored(cx, cy));
    } else {
        f_climbed = (T)__dmul_rn(0.25, __dadd_rn(__dadd_rn(ax, cx), __dadd_rn(ay, cy)));
        g_climbed = (T)__dmul_rn(0.25, __dadd_rn(__dadd_rn(ax, ay), __dadd_rn(cx, cy)));
    }
}

inline size_t layer_ring_bytes(int ft, int stages, size_t es) { return (size_t)stages * 2 * ft * (kPTile * es + 16); }
// consumer: staged parent-row segments of a couple tile (2 x kMaxTileFam rows x sw pairs), Va | Vb, a ring of
// three tiles' metadata, the strip's member rows
inline size_t layer_consumer_bytes(int sw, size_t es) {
    return (size_t)2 * kMaxTileFam * sw * 2 * es + (size_t)2 * sw * kVPitch * es + (size_t)3 * 4 * kMTile * 4 + (size_t)kLayerThreads * 16;
}

// The items of strip s that CTA k of a role with n CTAs takes: first, first + n, ... below n_items.
__device__ __forceinline__ int first_item(int k, int s, int rot, int n) {
    int f = (k - (int)(((long long)s * rot) % n)) % n;
    return f < 0 ? f + n : f;
}

// Optional cycle accounting per CTA and phase (-DGENLIB_PROFILE): thread 0's clock64 between phase marks,
// summed into S.prof[blockIdx.x * 8 + phase]; scripts/prof_layers.py prints them.
#ifdef GENLIB_PROFILE
#define PROF_DECL long long prof_t = clock64(), prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROF_MARK(ph) do { const long long now_ = clock64(); prof_acc[ph] += now_ - prof_t; prof_t = now_; } while (0)
#define PROF_FLUSH() do { if (tid == 0 && S.prof) for (int i_ = 0; i_ < 8; i_++) S.prof[(size_t)blockIdx.x * 8 + i_] = prof_acc[i_]; } while (0)
#else
#define PROF_DECL
#define PROF_MARK(ph) do { } while (0)
#define PROF_FLUSH() do { } while (0)
#endif

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
    PROF_DECL

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
    // A producer's share of the strip is done: its pairs were written with ordinary stores and will be read with
    // bulk copies (async proxy), so every writer orders its stores against that proxy before the CTA's release.
    auto produced = [&](int *counter) {
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicAdd(counter, 1); }
    };
    // A consumer's share is done: its copies of the strip's pairs have landed in shared memory (it waited for
    // them), which is all the producers that will overwrite the buffer need to know -- no fence: the rows it
    // wrote are read by the next kernel at the earliest.
    auto consumed = [&](int *counter) {
        __syncthreads();
        if (tid == 0) atomicAdd(counter, 1);
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
        int issue_tile = is < NS ? S.live_tiles[ii / npt] & (kTileCarried - 1) : 0;   // tile of the item at the issue cursor
        int rows_of = -1;                                          // strip whose parent rows are in s_row
        unsigned n_issued = 0, n_done = 0;
        const int rpw = 2 * ft / kLayerWarps;                      // rows each warp issues
        const int cpw = ft / kLayerWarps;                          // couples per warp for the member rows
        const int f = lane % ft;
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
                const int pt = ii % npt, tile = issue_tile;
                const int row = warp * rpw + lane;                 // 0 .. 2 ft - 1: fathers, then mothers
                const unsigned bar = bar0 + 8u * slot;
                const unsigned dst = sbase + slot * (unsigned)STAGE + (unsigned)(row * RB);
                const int fi = pt * ft + (row < ft ? row : row - ft);
                const T *src = s_row[(row < ft ? 0 : sw) + fi];
                // (the stage was last READ with ordinary loads, before the barrier we come from: a bulk copy may
                //  overwrite it without a proxy fence; only the zero fill below WRITES it through the generic proxy)
                if (src) {
                    mbar_arrive_expect_tx(bar, ROWB);
                    bulk_g2s(dst, src + tile * kPTile, ROWB, bar);
                } else {                                           // unknown parent: contributes 0 (compute.jl:111-126)
                    for (unsigned c = 0; c < ROWB; c += 16) zero16_shared(dst + c);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive_expect_tx(bar, 0);
                }
            }
            n_issued++;
            ii += NP;
            while (is < NS && ii >= NI) { is++; ii = first_item(k, is, S.rot_p, NP); }
            if (is < NS) issue_tile = S.live_tiles[ii / npt] & (kTileCarried - 1);   // in flight until the next issue
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
                for (int s = max(cur, 0); s < cs; s++) produced(done_p + s);
                cur = cs;
                PROF_MARK(0);
                wait_for(done_c + (cs - S.nbuf), cs >= S.nbuf ? S.n_cons : 0);    // the strip that used this buffer is consumed
                PROF_MARK(1);
            }
            int ncs = cs, nci = ci + NP;                           // the item after this one
            while (ncs < NS && nci >= NI) { ncs++; nci = first_item(k, ncs, S.rot_p, NP); }
            const int tinfo_next = ncs < NS ? S.live_tiles[nci / npt] : 0;
            const unsigned slot = n_done % (unsigned)S.stages;
            PROF_MARK(0);
            if (!mbar_wait(bar0 + 8u * slot, (n_done / (unsigned)S.stages) & 1u)) atomicExch(err, 2);
            __syncthreads();                                       // everybody is done with the stage refilled next
            PROF_MARK(2);
            try_issue();
            PROF_MARK(3);
            const int pt = ci % npt, lt = ci / npt;
            const int tile = tinfo & (kTileCarried - 1);
            const unsigned char *st = dyn_smem + slot * STAGE;
            P2 *const Q = Qall + (size_t)(cs % S.nbuf) * S.qstride;
            // ---- transposed and interleaved: Q[p][F] = (father row, mother row) at column p.  A lane takes one
            //      couple and four columns at a time: two 128-bit shared loads (a quarter warp spans the 32 banks:
            //      rows are padded by 16 bytes), four 8-byte stores that the lanes of a warp lay side by side. ----
            {
                const int gpi = 32 / ft;                           // column groups a warp handles per step
                const unsigned char *xr = st + f * RB, *yr = st + (ft + f) * RB;
                P2 *q = Q + (size_t)lt * kPTile * sw + pt * ft + f;
                for (int gq = lane / ft; gq < 4; gq += gpi) {
                    const uint32_t w = __shfl_sync(0xffffffffu, live4, gq);
                    const int col = warp * (kPTile / 8) + gq * 4;
                    T x[4], y[4];
                    if constexpr (sizeof(T) == 4) {
                        const float4 a4 = *reinterpret_cast<const float4 *>(xr + col * 4), b4 = *reinterpret_cast<const float4 *>(yr + col * 4);
                        x[0] = a4.x; x[1] = a4.y; x[2] = a4.z; x[3] = a4.w; y[0] = b4.x; y[1] = b4.y; y[2] = b4.z; y[3] = b4.w;
                    } else {
                        const double2 a0 = *reinterpret_cast<const double2 *>(xr + col * 8), a1 = *reinterpret_cast<const double2 *>(xr + col * 8 + 16);
                        const double2 b0 = *reinterpret_cast<const double2 *>(yr + col * 8), b1 = *reinterpret_cast<const double2 *>(yr + col * 8 + 16);
                        x[0] = a0.x; x[1] = a0.y; x[2] = a1.x; x[3] = a1.y; y[0] = b0.x; y[1] = b0.y; y[2] = b1.x; y[3] = b1.y;
                    }
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++)
                        if ((w >> (8 * c4)) & kFlagLive) {
                            CHECK(((size_t)lt * kPTile + col + c4) * sw + pt * ft + f < (size_t)S.qstride);
                            P2 v; v.x = x[c4]; v.y = y[c4];
                            q[(size_t)(col + c4) * sw] = v;
                        }
                }
            }
            PROF_MARK(4);
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
                    CHECK(col0 + 3 < ld && mb >= 0 && me <= L.n_new);
                    for (int m = mb; m < me; m++) store4(A + (int64_t)L.mem_lrow[m] * ld + col0, rr);
                }
            }
            PROF_MARK(5);
            n_done++;
            cs = ncs; ci = nci;
            tinfo = tinfo_next;
            live4 = cs < NS ? tile_flags(tinfo) : 0u;
        }
        for (int s = max(cur, 0); s < NS; s++) produced(done_p + s);
        PROF_MARK(0);
        PROF_FLUSH();
        return;
    }

    // ===== consumer: couple tiles -> the strip members' rows; carried rows <- the strip members' columns =====
    // Nothing the inner loops need comes from a dependent global load: tile descriptors travel three items
    // ahead in registers, a tile's metadata (its couples' strip-buffer rows, its members' couple / rank /
    // slot) two items ahead into a shared-memory ring with cp.async, its segments one item ahead, and the
    // strip's member rows one strip ahead.
    const int k = blockIdx.x - S.n_prod, NC = S.n_cons, NI = S.n_citems;
    constexpr int kMetaSlots = 3, kMetaInts = 4 * kMTile;          // per slot: qrow[128] | couple[128] | rank[128] | slot[128]
    constexpr int kRowCache = kLayerThreads;                       // strip member rows kept in shared memory at a time
    struct RowMeta { unsigned rowoff; int rank; long long bytes; };   // Va row offset, rank, byte offset of the frontier row
    unsigned char *sm = dyn_smem;
    P2 *const stg = reinterpret_cast<P2 *>(sm);                    // [2 g + parent][f]
    sm += (size_t)2 * kMaxTileFam * sw * sizeof(P2);
    T *const Va = reinterpret_cast<T *>(sm);                       // [f][g]: the strip couple's member is climbed first
    T *const Vb = Va + (size_t)sw * kVPitch;                       // [f][g]: the tile couple's member is climbed first
    sm += (size_t)2 * sw * kVPitch * sizeof(T);
    int *const meta = reinterpret_cast<int *>(sm);
    sm += (size_t)kMetaSlots * kMetaInts * sizeof(int);
    RowMeta *const rowmeta = reinterpret_cast<RowMeta *>(sm);
    const unsigned stg_s = (unsigned)__cvta_generic_to_shared(stg), meta_s = (unsigned)__cvta_generic_to_shared(meta);
    const unsigned va_s = (unsigned)__cvta_generic_to_shared(Va), vb_s = (unsigned)__cvta_generic_to_shared(Vb);
    const int lsw = 31 - __clz(sw);                                // sw is a power of two
    const int row_bytes = sw * (int)sizeof(P2);

    struct Cur { int s, it; };                                     // an item of this CTA: strip, index in the strip
    auto advance = [&](Cur c) {
        c.it += NC;
        while (c.s < NS && c.it >= NI) { c.s++; c.it = first_item(k, c.s, S.rot_c, NC); }
        return c;
    };
    auto is_tile = [&](Cur c) { return c.s < NS && c.it < L.n_mtiles; };
    auto fetch_desc = [&](Cur c) { return is_tile(c) ? __ldg(L.mt_desc + c.it) : make_int4(0, 0, 0, 0); };
    // metadata of tile `d` -> ring slot (cp.async; the caller commits)
    auto fetch_meta = [&](int4 d, int slot) {
        CHECK(d.y >= 1 && d.y <= kMaxTileFam && d.w >= 1 && d.w <= kMTile && d.x >= 0 && d.x + d.y <= L.n_fam && d.z >= 0 && d.z + d.w <= L.n_new);
        const unsigned base = meta_s + (unsigned)(slot * kMetaInts * (int)sizeof(int));
        if (tid < d.y)                                             // the couples' (father, mother) rows in the strip buffers
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(base + 8u * tid), "l"(L.fam_q + d.x + tid) : "memory");
        const int nchunk = (d.w + 3) >> 2;                         // member columns: couple, rank, slot in 16-byte chunks
        if (tid < 3 * nchunk) {
            const int arr = tid / nchunk, c = tid - arr * nchunk;
            const int32_t *src = (arr == 0 ? L.mem_fam : arr == 1 ? L.mem_ind : L.mem_slot) + d.z + 4 * c;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(base + (unsigned)((kMTile * (1 + arr) + 4 * c) * 4)), "l"(src) : "memory");
        }
    };
    // the two parent-row segments of every couple of the tile -> stg: one TMA bulk copy per segment, issued by
    // the first 2 * kMaxTileFam threads (one each), all completing on one mbarrier phase per tile
    const unsigned sbar = (unsigned)__cvta_generic_to_shared(&s_bar[0]);
    if (tid == 0) {
        mbar_init(sbar, 2 * kMaxTileFam);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned n_staged = 0, n_landed = 0;                           // tiles whose copies were issued / awaited
#ifdef GENLIB_STAGE_LDGSTS
    // (16-byte cp.async through the LSU: the TMA unit is left to the producers, whose requests are as small)
    auto stage_tile = [&](const P2 *Q, int nfJ, int slot) {
        const int *qrow = meta + slot * kMetaInts;
        const int per = row_bytes / 16;                            // lanes per segment (4 ... 32)
        const int seg_per_step = 32 / per, sub = lane / per, part = lane - sub * per;
        for (int r0 = warp * seg_per_step; r0 < 2 * nfJ; r0 += kLayerWarps * seg_per_step) {
            const int r = r0 + sub;
            if (r < 2 * nfJ) {
                const int q = qrow[r];
                CHECK(q >= -1 && (long long)q * sw < S.qstride);
                const unsigned dst = stg_s + (unsigned)(r * row_bytes + part * 16);
                if (q >= 0) cp_async16_to(dst, reinterpret_cast<const unsigned char *>(Q + (size_t)q * sw) + part * 16);
                else zero16_shared(dst);                           // unknown parent: contributes 0
            }
        }
        n_staged++;
    };
    auto await_tile = [&]() { n_landed++; };                       // (cp_async_wait<0> by the caller covers the segments)
#else
    auto stage_tile = [&](const P2 *Q, int nfJ, int slot) {
        if (lane < 2 * kMaxTileFam / kLayerWarps) {                // 16 lanes of every warp: the issue is serial per warp
            const int r = warp * (2 * kMaxTileFam / kLayerWarps) + lane;
            const int q = r < 2 * nfJ ? meta[slot * kMetaInts + r] : -1;
            CHECK(q >= -1 && (long long)q * sw < S.qstride);
            const unsigned dst = stg_s + (unsigned)(r * row_bytes);
            if (q >= 0) {                                          // (write after read of the staging area: no proxy fence)
                mbar_arrive_expect_tx(sbar, (unsigned)row_bytes);
                bulk_g2s(dst, Q + (size_t)q * sw, (unsigned)row_bytes, sbar);
            } else {
                if (r < 2 * nfJ) {                                 // unknown parent: contributes 0
                    for (int c = 0; c < row_bytes; c += 16) zero16_shared(dst + c);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                mbar_arrive_expect_tx(sbar, 0);
            }
        }
        n_staged++;
    };
    auto await_tile = [&]() {                                      // all threads
        if (!mbar_wait(sbar, n_landed & 1u)) atomicExch(err, 2);
        n_landed++;
    };
#endif

    // ---- the strip's member rows: prefetched one strip ahead (registers), cached in shared memory ----
    int cur = -1;                                                  // strip this CTA is in
    int F0 = 0, nFs = 0, ms0 = 0, ms1 = 0, n_rows = 0;
    int pre_s = -1, pre_ms0 = 0, pre_ms1 = 0, pre_f = 0, pre_rank = 0, pre_lrow = 0;   // strip pre_s: bounds, row ms0 + tid
    int pre2_s = -1, pre2_ms0 = 0, pre2_ms1 = 0;                   // strip pre2_s: bounds only
    auto strip_bounds = [&](int st, int &b0, int &b1) {
        const int f0 = L.own_f0 + st * sw, nf = min(sw, L.own_nf - st * sw);
        b0 = L.fam_start[f0]; b1 = L.fam_start[f0 + nf];
    };
    auto strip_row = [&](int st, int b0, int b1, int first, int &f, int &rk, int &lr) {   // row first + tid of strip st
        const int im = min(b0 + first + tid, b1 - 1);
        f = L.mem_fam[im] - (L.own_f0 + st * sw); rk = L.mem_ind[im]; lr = L.mem_lrow[im];
    };
    auto put_row = [&](int f, int rk, int lr) {
        RowMeta m; m.rowoff = (unsigned)(f * kVPitch * (int)sizeof(T)); m.rank = rk; m.bytes = (long long)lr * ld * (long long)sizeof(T);
        rowmeta[tid] = m;
    };

    Cur c0; c0.s = 0; c0.it = first_item(k, 0, S.rot_c, NC);
    while (c0.s < NS && c0.it >= NI) { c0.s++; c0.it = first_item(k, c0.s, S.rot_c, NC); }
    Cur c1 = advance(c0), c2 = advance(c1), c3 = advance(c2);
    int4 d0 = fetch_desc(c0), d1 = fetch_desc(c1), d2 = fetch_desc(c2);
    unsigned n_item = 0;                                           // items done: item j's metadata sits in ring slot j % 3
    if (is_tile(c0)) fetch_meta(d0, 0);
    if (is_tile(c1)) fetch_meta(d1, 1);
    cp_async_commit();
    bool staged = false;                                           // the segments of c0 are on their way (or there)
    while (c0.s < NS) {
        const int s = c0.s, it = c0.it;
        if (s != cur) {
            for (int t = max(cur, 0); t < s; t++) consumed(done_c + t);
            cur = s;
            F0 = L.own_f0 + s * sw;
            nFs = min(sw, L.own_nf - s * sw);
            int rf, rk, rl;
            if (pre_s == s) { ms0 = pre_ms0; ms1 = pre_ms1; rf = pre_f; rk = pre_rank; rl = pre_lrow; }
            else { strip_bounds(s, ms0, ms1); strip_row(s, ms0, ms1, 0, rf, rk, rl); }
            n_rows = ms1 - ms0;
            put_row(rf, rk, rl);                                   // (read after the barriers below)
            // the next strip's rows, and the bounds of the one after, are fetched now and used a strip later
            if (s + 1 < NS) {
                if (pre2_s == s + 1) { pre_ms0 = pre2_ms0; pre_ms1 = pre2_ms1; } else strip_bounds(s + 1, pre_ms0, pre_ms1);
                strip_row(s + 1, pre_ms0, pre_ms1, 0, pre_f, pre_rank, pre_lrow);
                pre_s = s + 1;
                if (s + 2 < NS) { strip_bounds(s + 2, pre2_ms0, pre2_ms1); pre2_s = s + 2; }
            }
            PROF_MARK(2);
            if (!staged) wait_for(done_p + s, S.n_prod);          // the strip's pairs are complete (in L2)
            else __syncthreads();
            PROF_MARK(1);
        }
        const P2 *const Q = Qall + (size_t)(s % S.nbuf) * S.qstride;
        const int slot0 = (int)(n_item % kMetaSlots);

        if (it >= L.n_mtiles) {
            // ---- a block of carried rows: the strip members' columns, Psi[c, i] = RN(hs(Q[c][F_i])) ----
            const int r0 = (it - L.n_mtiles) * S.mrows, r1 = min(L.rt_rows, r0 + S.mrows);
            for (int row = r0 + warp; row < r1; row += kLayerWarps) {
                if (!(L.flags[row] & kFlagCarried)) continue;
                T *dst = static_cast<T *>(PT.A[L.live_owner[row]]) + (int64_t)L.live_lrow[row] * ld;
                CHECK(L.tile_map[row / kPTile] >= 0 && L.live_owner[row] >= 0 && L.live_lrow[row] >= 0);
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
            // keep the pipeline of descriptors and metadata moving
            const int4 d3 = fetch_desc(c3);
            if (is_tile(c2)) fetch_meta(d2, (int)((n_item + 2) % kMetaSlots));
            cp_async_commit();
            staged = false;
            c0 = c1; c1 = c2; c2 = c3; c3 = advance(c3);
            d0 = d1; d1 = d2; d2 = d3;
            n_item++;
            PROF_MARK(7);
            continue;
        }

        // ---- a member tile: its couples' segments (staged ahead unless the strip was not ready), Va | Vb, expansion ----
        const int fJ0 = d0.x, nfJ = d0.y, mJ0 = d0.z, cntJ = d0.w;
        if (!staged) {
            cp_async_wait<0>();
            __syncthreads();                                       // the tile's metadata is in the ring
            stage_tile(Q, nfJ, slot0);
            cp_async_commit();
        }
        PROF_MARK(0);
        cp_async_wait<0>();                                        // the next tile's metadata (own copies) ...
        await_tile();                                              // ... and this tile's segments have landed
        __syncthreads();                                           // for everybody; the previous item is written
        PROF_MARK(3);
        {
            const int fl = tid & (sw - 1), g0 = tid >> lsw, gstep = kLayerThreads >> lsw;
            if (fl < nFs) {
#pragma unroll 4
                for (int g = g0; g < nfJ; g += gstep) {
                    const P2 a = stg[(2 * g) * sw + fl], c = stg[(2 * g + 1) * sw + fl];
                    T vf, vg;
                    couple_pair<T, STORED>((double)a.x, (double)a.y, (double)c.x, (double)c.y, vf, vg);
                    Va[fl * kVPitch + g] = vf;
                    Vb[fl * kVPitch + g] = vg;
                }
            }
        }
        __syncthreads();                                           // Va | Vb complete, the staging area is free
        PROF_MARK(4);
        // the pipeline: descriptor of item +3, metadata of item +2, segments of item +1 (if its strip is produced)
        const int4 d3 = fetch_desc(c3);
        if (is_tile(c2)) fetch_meta(d2, (int)((n_item + 2) % kMetaSlots));
        staged = false;
        if (is_tile(c1)) {
            bool ready = c1.s == s;
            if (!ready) {                                          // peek: do not wait here, the expansion comes first
                if (tid == 0) s_ready = ld_acquire_gpu(done_p + c1.s) >= S.n_prod;
                __syncthreads();
                ready = s_ready != 0;
            }
            if (ready) { stage_tile(Qall + (size_t)(c1.s % S.nbuf) * S.qstride, d1.y, (int)((n_item + 1) % kMetaSlots)); staged = true; }
        }
        cp_async_commit();
        PROF_MARK(5);
        // ---- expansion: the lane's four member columns, the warp's share of the strip's rows, four rows at a time ----
        {
            const int *mcol = meta + slot0 * kMetaInts + kMTile;
            const int4 cg = *reinterpret_cast<const int4 *>(mcol + 4 * lane);
            const int4 cr = *reinterpret_cast<const int4 *>(mcol + kMTile + 4 * lane);
            const int4 cs4 = *reinterpret_cast<const int4 *>(mcol + 2 * kMTile + 4 * lane);
            const int ncol = min(4, cntJ - 4 * lane);
            // columns past the end of the tile repeat the lane's first one (what the ring holds there is stale)
            const int gq[4] = {cg.x, ncol > 1 ? cg.y : cg.x, ncol > 2 ? cg.z : cg.x, ncol > 3 ? cg.w : cg.x};
            const int rj[4] = {cr.x, ncol > 1 ? cr.y : cr.x, ncol > 2 ? cr.z : cr.x, ncol > 3 ? cr.w : cr.x};
            const int sj[4] = {cs4.x, cs4.y, cs4.z, cs4.w};
            unsigned off[4];
#pragma unroll
            for (int q = 0; q < 4; q++) off[q] = ncol > 0 ? (unsigned)((gq[q] - fJ0) * (int)sizeof(T)) : 0u;
            const bool vec = ncol == 4 && ((sj[0] & 3) == 0) && sj[1] == sj[0] + 1 && sj[2] == sj[0] + 2 && sj[3] == sj[0] + 3;
            unsigned char *const Ab = reinterpret_cast<unsigned char *>(A);
            for (int pass0 = 0; pass0 < n_rows; pass0 += kRowCache) {
                if (pass0 > 0 || n_rows > kRowCache) {             // a strip with more rows than the cache: reload it per pass
                    __syncthreads();
                    int rf, rk, rl;
                    strip_row(s, ms0, ms1, pass0, rf, rk, rl);
                    put_row(rf, rk, rl);
                    __syncthreads();
                }
                const int nrp = min(kRowCache, n_rows - pass0);
                const int share = (nrp + kLayerWarps - 1) / kLayerWarps;
                const int rbeg = warp * share, rend = min(nrp, rbeg + share);
                if (ncol > 0) {
                    int r = rbeg;
                    for (; r + 4 <= rend; r += 4) {
                        RowMeta m[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            m[u] = rowmeta[r + u];
                            CHECK(m[u].rowoff < (unsigned)(sw * kVPitch * (int)sizeof(T)) && m[u].bytes >= 0);
                        }
                        CHECK(off[0] < 260u && off[1] < 260u && off[2] < 260u && off[3] < 260u);
                        CHECK(sj[0] >= 0 && sj[0] < ld);
                        T v[4][4];
#pragma unroll
                        for (int u = 0; u < 4; u++)
#pragma unroll
                            for (int q = 0; q < 4; q++)            // the higher rank is climbed first (compute.jl:130-147)
                                v[u][q] = lds<T>((m[u].rank > rj[q] ? va_s : vb_s) + m[u].rowoff + off[q]);
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            T *row = reinterpret_cast<T *>(Ab + m[u].bytes);
                            if (vec) store_vec4(row + sj[0], v[u]);
                            else {
#pragma unroll
                                for (int q = 0; q < 4; q++) if (q < ncol) row[sj[q]] = v[u][q];
                            }
                        }
                    }
                    for (; r < rend; r++) {
                        const RowMeta m = rowmeta[r];
                        CHECK(m.rowoff < (unsigned)(sw * kVPitch * (int)sizeof(T)) && m.bytes >= 0);
                        T *row = reinterpret_cast<T *>(Ab + m.bytes);
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if (q < ncol) row[sj[q]] = lds<T>((m.rank > rj[q] ? va_s : vb_s) + m.rowoff + off[q]);
                    }
                }
                // own diagonal entries (compute.jl:148-155): rows of this pass that are also columns of this tile,
                // written over the row segments above (__syncwarp orders the warp's stores)
                __syncwarp();
                const int dlo = max(ms0 + pass0 + rbeg, mJ0), dhi = min(ms0 + pass0 + rend, mJ0 + cntJ);
                for (int i = dlo + lane; i < dhi; i += 32) {
                    const RowMeta m = rowmeta[i - ms0 - pass0];
                    const int F = F0 + (int)(m.rowoff / (unsigned)(kVPitch * (int)sizeof(T)));
                    CHECK(F >= 0 && F < L.n_fam && i - mJ0 >= 0 && i - mJ0 < kMTile);
                    const int pf = L.fam_pf[F], pm = L.fam_pm[F];
                    double d = 0.5;
                    if (pf >= 0 && pm >= 0)
                        d = half_sum_mode<STORED>((double)(static_cast<const T *>(PT.A[L.fam_pf_owner[F]]) + (int64_t)L.fam_pf_lrow[F] * ld)[pm], 1.0);
                    reinterpret_cast<T *>(Ab + m.bytes)[meta[slot0 * kMetaInts + 3 * kMTile + (i - mJ0)]] = (T)d;
                }
            }
        }
        c0 = c1; c1 = c2; c2 = c3; c3 = advance(c3);
        d0 = d1; d1 = d2; d2 = d3;
        n_item++;
        PROF_MARK(6);
    }
    for (int t = max(cur, 0); t < NS; t++) consumed(done_c + t);
    PROF_MARK(0);
    PROF_FLUSH();
}

}  // namespace genlib
