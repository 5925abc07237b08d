// engine.cu -- device runtime + C ABI of libgenlib_cuda.so (include/genlib_cuda.h).
//
// The ABI stands where the reference's Julia method stands:
//   phi(pedigree, probandIDs; verbose, compute)            src/compute.jl:233-304
// There is no CPU fallback: without a usable CUDA device every compute entry
// point returns GENLIB_ECUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/genlib_cuda.h"
#include "kernels.cuh"
#include "plan.hpp"

using namespace genlib;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) { g_err = msg; return code; }
}  // namespace
int genlib::set_error(int code, const std::string &msg) { return fail(code, msg); }
namespace {

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(e_ == cudaErrorMemoryAllocation ? GENLIB_ENOMEM : GENLIB_ECUDA,       \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    int enter(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return fail(GENLIB_ECUDA, "no usable CUDA device (cudaGetDevice failed)");
        if (dev >= 0 && dev != prev) {
            if (cudaSetDevice(dev) != cudaSuccess) return fail(GENLIB_ECUDA, "cudaSetDevice failed");
            active = true;
        }
        return GENLIB_OK;
    }
    ~DeviceGuard() { if (active) cudaSetDevice(prev); }
};

// ---- device arena: ONE allocation per engine, cached across calls -------------------------
// cudaMalloc / cudaFree of tens of GB cost hundreds of milliseconds; a repeated gen.phi call
// (and the one-shot genlib_phi) reuses the previous arena when it is large enough.
struct Arena {
    void *base = nullptr;
    size_t size = 0;
    int device = -1;
};

class ArenaCache {
    std::mutex mu_;
    std::vector<Arena> free_;
public:
    cudaError_t acquire(int device, size_t bytes, Arena &out) {
        {
            std::lock_guard<std::mutex> g(mu_);
            int best = -1;
            for (int i = 0; i < (int)free_.size(); i++)
                if (free_[i].device == device && free_[i].size >= bytes && (best < 0 || free_[i].size < free_[best].size)) best = i;
            if (best >= 0) { out = free_[best]; free_.erase(free_.begin() + best); return cudaSuccess; }
        }
        // Nothing cached fits.  Cached arenas may still be mapped by peer processes (CUDA IPC), and
        // freeing exported memory under an open mapping is undefined, so they are only given back
        // when the device is out of memory.
        out = Arena{nullptr, bytes, device};
        cudaError_t ce = cudaMalloc(&out.base, bytes);
        if (ce == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            release_device(device);
            ce = cudaMalloc(&out.base, bytes);
        }
        return ce;
    }
    void release(Arena a) {
        if (!a.base) return;
        std::lock_guard<std::mutex> g(mu_);
        free_.push_back(a);
    }
    void release_device(int device) {
        std::lock_guard<std::mutex> g(mu_);
        for (size_t i = 0; i < free_.size();) {
            if (device < 0 || free_[i].device == device) {
                int prev = -1; cudaGetDevice(&prev);
                cudaSetDevice(free_[i].device); cudaFree(free_[i].base);
                if (prev >= 0) cudaSetDevice(prev);
                free_.erase(free_.begin() + (long)i);
            } else i++;
        }
    }
};
ArenaCache g_arenas;

// cudaIpcOpenMemHandle / CloseMemHandle of multi-GB arenas cost tens of milliseconds per peer.
// Arenas are cached, so the same handles come back call after call: keep the mappings open.
class PeerMapCache {
    std::mutex mu_;
    struct Entry { unsigned char handle[64]; int device; void *base; };
    std::vector<Entry> open_;
public:
    cudaError_t map(int device, const void *handle64, void **base) {
        std::lock_guard<std::mutex> g(mu_);
        for (const Entry &en : open_)
            if (en.device == device && std::memcmp(en.handle, handle64, 64) == 0) { *base = en.base; return cudaSuccess; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, sizeof h);
        cudaError_t ce = cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess);
        if (ce != cudaSuccess) return ce;
        Entry en; std::memcpy(en.handle, handle64, 64); en.device = device; en.base = *base;
        open_.push_back(en);
        return cudaSuccess;
    }
    void close_all() {
        std::lock_guard<std::mutex> g(mu_);
        for (const Entry &en : open_) {
            int prev = -1; cudaGetDevice(&prev);
            cudaSetDevice(en.device); cudaIpcCloseMemHandle(en.base);
            if (prev >= 0) cudaSetDevice(prev);
        }
        open_.clear();
    }
};
PeerMapCache g_peer_maps;

// non-owning typed view into the arena
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    static size_t padded(size_t count) { return (std::max<size_t>(count, 1) * sizeof(T) + 255) / 256 * 256; }
    void place(unsigned char *&cursor, size_t count) { p = reinterpret_cast<T *>(cursor); n = count; cursor += padded(count); }
    cudaError_t upload(const std::vector<T> &h, cudaStream_t s) {
        if (h.empty()) return cudaSuccess;
        return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

constexpr size_t kFetchStageBytes = (size_t)64 << 20;

}  // namespace

struct genlib_plan {
    Plan p;
    double ms_plan = 0;
};

constexpr int kMaxGroups = 8;    // couple groups per layer (multi-GPU software pipeline of cross and couple)

struct genlib_engine {
    const genlib_plan *plan = nullptr;
    int numerics = 0, device = 0, sm_count = 148;
    int rank = 0, world = 1;
    bool attached = false;                 // peers' arenas mapped (always true for one rank)
    size_t esize = 4;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t side_stream = nullptr;    // multi-GPU: couple_kernel of one couple group runs beside cross_kernel of the next
    cudaEvent_t group_ev[kMaxGroups + 1] = {};
    bool piped = false;
    Arena arena;
    void *A = nullptr;                     // this rank's frontier rows: rows_cap x capacity
    double *Rt = nullptr;                  // transposed cross block: live slots x own couples (fp64)
    void *Vrow = nullptr, *Vt = nullptr, *Dg = nullptr;   // V[own F, G], V[G, own F], couple diagonal
    unsigned *bar_flags = nullptr;
    unsigned char *fetch_stage[2] = {nullptr, nullptr};
    PeerTable peers{};
    BarrierTable bars{};
    void *peer_base[kMaxWorld] = {};       // cudaIpcOpenMemHandle mappings (to close)
    unsigned epoch = 0;
    DevBuf<int32_t> mem_ind, mem_slot, mem_fam, mem_lrow, fam_pf, fam_pm, fam_pf_lrow, fam_pm_lrow, fam_start,
        fam_minrank, fam_maxrank, mt_min, mt_max, mt_fam0, mt_nfam, mt_m0, mt_cnt, pro_slot, own_pro_row, live_lrow;
    DevBuf<int8_t> fam_pf_owner, fam_pm_owner, live_owner, mem_gowner;
    DevBuf<int32_t> mem_glrow, mem_rank;
    DevBuf<uint8_t> flags;
    DevBuf<double> acc;
    std::vector<int32_t> own_pro;          // proband indices (output rows) this rank owns, ascending
    std::vector<genlib_layer_info> info;
    std::vector<cudaEvent_t> events;
    genlib_stats stats{};
    bool ran = false;
    int32_t layer_limit = -1;
    ~genlib_engine() {
        for (auto e : events) cudaEventDestroy(e);
        if (stream) cudaStreamSynchronize(stream);
        if (copy_stream) cudaStreamSynchronize(copy_stream);
        if (side_stream) { cudaStreamSynchronize(side_stream); cudaStreamDestroy(side_stream); }
        for (auto e : group_ev) if (e) cudaEventDestroy(e);
        g_arenas.release(arena);
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
    }
};

namespace {

void fill_info(const Layer &L, genlib_layer_info *o) {
    std::memset(o, 0, sizeof *o);
    o->n_new = L.n_new; o->n_fam = L.n_fam; o->live_before = L.live_before; o->carried = L.carried;
    o->ref_founders = L.ref_founders; o->ref_probands = L.ref_probands; o->ref_both = L.ref_both;
    o->alg_elems = L.alg_elems;
}

size_t pad256(size_t b) { return (std::max<size_t>(b, 1) + 255) / 256 * 256; }

size_t plan_index_bytes(const Plan &P) {
    return (P.mem_ind.size() * 4 + P.fam_pf.size() * 6 + P.fam_start.size() + P.mtile_minrank.size() * 6 +
            P.pro_slot.size() * 2 + P.live_lrow.size()) * sizeof(int32_t) + P.fam_pf.size() * 2 + P.live_owner.size() + P.flags.size();
}

// Arena layout of rank g.  The first three regions are what peers address (barrier flags,
// frontier rows, row block of V), so their offsets must be computable by every rank.
constexpr size_t kFlagBytes = 256;
size_t total_rows(const Plan &P, int g) { return (size_t)P.rows_cap[g] + 2 * (size_t)P.guest_cap[g]; }
size_t a_bytes(const Plan &P, size_t es, int g) { return pad256(total_rows(P, g) * (size_t)P.capacity * es); }
size_t v_bytes(const Plan &P, size_t es, int g) { return pad256(P.rank_v_elems[g] * es); }
size_t off_A() { return kFlagBytes; }
size_t off_Vrow(const Plan &P, size_t es, int g) { return kFlagBytes + a_bytes(P, es, g); }

size_t engine_bytes(const Plan &P, int numerics, int g) {
    const size_t es = numerics == GENLIB_NUMERICS_FP64 ? 8 : 4;
    size_t b = kFlagBytes + a_bytes(P, es, g) + 2 * v_bytes(P, es, g);          // flags, A, Vrow, Vt
    b += pad256(P.rank_rt_elems[g] * sizeof(double));                          // Rt
    b += pad256(P.fam_pf.size() * es);                                         // Dg
    b += 2 * pad256(kFetchStageBytes);                                         // proband staging
    b += 4 * DevBuf<int32_t>::padded(P.mem_ind.size()) + 6 * DevBuf<int32_t>::padded(P.fam_pf.size()) +
         DevBuf<int32_t>::padded(P.fam_start.size()) + 6 * DevBuf<int32_t>::padded(P.mtile_minrank.size()) +
         2 * DevBuf<int32_t>::padded(P.pro_slot.size()) + DevBuf<int32_t>::padded(P.live_lrow.size()) +
         2 * DevBuf<int8_t>::padded(P.fam_pf.size()) + DevBuf<int8_t>::padded(P.live_owner.size()) +
         DevBuf<int8_t>::padded(P.mem_gowner.size()) + DevBuf<int32_t>::padded(P.mem_glrow.size()) +
         DevBuf<int32_t>::padded(P.mem_rank.size()) +
         DevBuf<uint8_t>::padded(P.flags.size()) + DevBuf<double>::padded(2);
    return b;
}

LayerArgs layer_args(const genlib_engine &E, int t) {
    const Plan &P = E.plan->p;
    const Layer &L = P.layers[t];
    LayerArgs a;
    std::memset(&a, 0, sizeof a);
    a.n_new = L.n_new; a.n_fam = L.n_fam; a.rt_lo = L.rt_lo; a.rt_rows = L.rt_rows; a.nf_pad = L.nf_pad;
    a.any_carried = L.carried > 0;
    a.rank = E.rank; a.world = E.world;
    const int32_t *fb = P.fam_base.data() + L.base_off, *mb = P.mem_base.data() + L.base_off;
    for (int g = 0; g <= E.world; g++) a.fam_base[g] = fb[g];
    a.own_f0 = fb[E.rank]; a.own_nf = fb[E.rank + 1] - fb[E.rank];
    a.own_m0 = mb[E.rank]; a.own_nm = mb[E.rank + 1] - mb[E.rank];
    a.nfo_pad = pad32(a.own_nf);
    a.ftile_shift = E.world > 1 ? fb[(E.rank + 1) % E.world] / kFTile : 0;
    a.mem_ind = E.mem_ind.p + L.mem_off; a.mem_slot = E.mem_slot.p + L.mem_off; a.mem_fam = E.mem_fam.p + L.mem_off;
    a.mem_lrow = E.mem_lrow.p + L.mem_off;
    a.mem_gowner = E.mem_gowner.p + L.mem_off; a.mem_glrow = E.mem_glrow.p + L.mem_off;
    a.fam_pf = E.fam_pf.p + L.fam_off; a.fam_pm = E.fam_pm.p + L.fam_off;
    a.fam_pf_owner = E.fam_pf_owner.p + L.fam_off; a.fam_pm_owner = E.fam_pm_owner.p + L.fam_off;
    a.fam_pf_lrow = E.fam_pf_lrow.p + L.fam_off; a.fam_pm_lrow = E.fam_pm_lrow.p + L.fam_off;
    a.fam_start = E.fam_start.p + L.fam_off + t;
    a.flags = E.flags.p + L.flag_off;
    a.live_owner = E.live_owner.p + L.flag_off; a.live_lrow = E.live_lrow.p + L.flag_off;
    a.fam_minrank = E.fam_minrank.p + L.fam_off; a.fam_maxrank = E.fam_maxrank.p + L.fam_off;
    a.mt_minrank = E.mt_min.p + L.mtile_off; a.mt_maxrank = E.mt_max.p + L.mtile_off;
    a.mt_fam0 = E.mt_fam0.p + L.mtile_off; a.mt_nfam = E.mt_nfam.p + L.mtile_off;
    a.mt_m0 = E.mt_m0.p + L.mtile_off; a.mt_cnt = E.mt_cnt.p + L.mtile_off;
    a.n_mtiles = L.n_mtiles;
    const int vec = 16 / (int)E.esize;
    a.vstride = (std::max(L.max_tile_fam, 1) + vec - 1) / vec * vec + vec;
    return a;
}

int launch_barrier(genlib_engine &E) {
    if (E.world == 1) return GENLIB_OK;
    E.epoch++;
    barrier_kernel<<<1, 32, 0, E.stream>>>(E.bars, E.rank, E.world, E.epoch, (long long)20e9);
    return GENLIB_OK;
}

template <typename T>
int launch_layers(genlib_engine &E, bool timed) {
    const Plan &P = E.plan->p;
    T *A = static_cast<T *>(E.A);
    const int64_t ld = P.capacity;
    const size_t cross_smem = cross_smem_bytes<T>();
    const int vec = 16 / (int)sizeof(T);
    int max_tile_fam = 1;
    for (const Layer &L : P.layers) max_tile_fam = std::max(max_tile_fam, L.max_tile_fam);
    const size_t expand_smem_max = (size_t)kEWarps * 2 * expand_stage_bytes<T>((max_tile_fam + vec - 1) / vec * vec + vec);
    if (expand_smem_max > 227 * 1024) return fail(GENLIB_EINVAL, "couple tile too wide for the expand kernel");
    auto expand_fn = E.world > 1 ? expand_kernel<T, true> : expand_kernel<T, false>;
    CU(cudaFuncSetAttribute(expand_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)expand_smem_max));
    const bool stored = sparse_schedule(P.schedule);           // sparse_phi's arithmetic (Float32 halves of stored values)
    auto cross_fn = stored ? cross_kernel<T, true> : cross_kernel<T, false>;
    auto couple_fn = stored ? couple_kernel<T, true> : couple_kernel<T, false>;
    const size_t cross_smem_piped = std::max<size_t>(cross_smem, 120 * 1024);    // one CTA per SM
    CU(cudaFuncSetAttribute(cross_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cross_smem_piped));
    CU(cudaFuncSetAttribute(cross_fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CU(cudaFuncSetAttribute(expand_fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    T *V = static_cast<T *>(E.Vrow), *Vt = static_cast<T *>(E.Vt), *Dg = static_cast<T *>(E.Dg);
    constexpr int kCRows = couple_rows<T>();
    const size_t couple_smem = sizeof(T) * kCRows * kCStride;
    CU(cudaFuncSetAttribute(couple_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)couple_smem));
    CU(cudaFuncSetAttribute(couple_fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int launches = 0;
    size_t ev = 0;
    for (int t = 0; t < (int)P.layers.size(); t++) {
        const Layer &L = P.layers[t];
        if (L.n_new == 0) continue;
        if (E.layer_limit >= 0 && t >= E.layer_limit) break;
        LayerArgs a = layer_args(E, t);
        if (timed) CU(cudaEventRecord(E.events[ev++], E.stream));
        // Couple groups (GENLIB_PIPE=1, several ranks): cross_kernel fills the columns of Rt that belong to a
        // group of own couples and couple_kernel consumes exactly those columns.  Cross is bound by NVLink
        // INGRESS (remote parent rows), couple by NVLink EGRESS (rows of V pushed to their owners), so
        // couple(group g) runs on a second, high-priority stream beside cross(group g + 1); cross then asks
        // for enough shared memory to keep ONE of its CTAs per SM, which leaves room for couple's.
        const int64_t ftiles = (a.own_nf + kFTile - 1) / kFTile;
        if (ftiles > 65535 || (L.n_fam + kCRows - 1) / kCRows > 65535) return fail(GENLIB_EINVAL, "layer too wide for one cross/couple launch");
        const int per_ctile = kCTile / kFTile;
        int groups = 1;
        if (E.piped && L.live_before > 0) groups = (int)std::max<int64_t>(1, std::min<int64_t>(kMaxGroups, ftiles / (16 * per_ctile)));
        const int64_t group_tiles = ((ftiles + groups - 1) / groups + per_ctile - 1) / per_ctile * per_ctile;
        const bool piped = groups > 1;
        for (int g = 0; g < groups && a.own_nf > 0; g++) {
            const int64_t ft0 = g * group_tiles, ft1 = std::min<int64_t>(ftiles, ft0 + group_tiles);
            if (ft0 >= ft1) break;
            const bool last = ft1 >= ftiles;
            a.ctile0 = (int32_t)ft0;
            if (L.live_before > 0) {
                // column tiles per CTA: long chunks amortise the pipeline fill, but keep >= ~4 waves of CTAs
                const int64_t ptiles = L.rt_rows / kPTile;
                const int64_t want = ptiles * (ft1 - ft0) / (4 * 2 * (int64_t)E.sm_count);
                a.pchunk = (int)std::max<int64_t>(std::min<int64_t>(4, ptiles), std::min<int64_t>(kMaxPChunk, want));
                dim3 grid((unsigned)((ptiles + a.pchunk - 1) / a.pchunk), (unsigned)(ft1 - ft0));
                cross_fn<<<grid, kThreads, piped ? cross_smem_piped : cross_smem, E.stream>>>(A, ld, E.Rt, E.peers, a);
                launches++;
            }
            if (last && L.live_before > 0 && L.carried > 0 && a.own_nm > 0) {
                dim3 mgrid((unsigned)((a.own_nm + 32 * kMirrorCols - 1) / (32 * kMirrorCols)), (unsigned)((L.rt_rows + kThreads / 32 - 1) / (kThreads / 32)));
                mirror_kernel<T><<<mgrid, kThreads, 0, E.stream>>>(E.Rt, ld, E.peers, a);
                launches++;
            }
            if (last && timed) CU(cudaEventRecord(E.events[ev++], E.stream));
            cudaStream_t cs = E.stream;
            if (piped) {
                CU(cudaEventRecord(E.group_ev[g], E.stream));
                CU(cudaStreamWaitEvent(E.side_stream, E.group_ev[g], 0));
                cs = E.side_stream;
            }
            const int64_t ct0 = ft0 / per_ctile, ct1 = (std::min<int64_t>(ft1 * kFTile, a.nfo_pad) + kCTile - 1) / kCTile;
            dim3 cgrid((unsigned)(ct1 - ct0), (unsigned)((L.n_fam + kCRows - 1) / kCRows));
            couple_fn<<<cgrid, kThreads, couple_smem, cs>>>(ld, E.Rt, Vt, Dg, E.peers, a);
            launches++;
        }
        if (a.own_nf <= 0 && timed) CU(cudaEventRecord(E.events[ev++], E.stream));
        if (piped) {                                         // join: the layer's V is complete on the main stream
            CU(cudaEventRecord(E.group_ev[kMaxGroups], E.side_stream));
            CU(cudaStreamWaitEvent(E.stream, E.group_ev[kMaxGroups], 0));
        }
        a.ctile0 = 0;
        if (timed) CU(cudaEventRecord(E.events[ev++], E.stream));
        launch_barrier(E);                 // every rank's row block of V is complete (peer stores landed)
        if (timed) CU(cudaEventRecord(E.events[ev++], E.stream));
        if (a.own_nm > 0) {
            const int rows_per_cta = kEWarps * kERows;
            dim3 grid((unsigned)((L.n_mtiles + kEChunk - 1) / kEChunk), (unsigned)((a.own_nm + rows_per_cta - 1) / rows_per_cta));
            if (grid.y > 65535) return fail(GENLIB_EINVAL, "layer too wide for one expand launch");
            const size_t smem = (size_t)kEWarps * 2 * expand_stage_bytes<T>(a.vstride);
            expand_fn<<<grid, kExpandThreads, smem, E.stream>>>(A, ld, V, Vt, Dg, E.peers, a);
            launches++;
            if (P.schedule == kScheduleSparsePhi && L.n_new > 1) {      // the reference's misfiled kinships read as 0
                if (E.world > 1 && P.guest_cap[E.rank] > 0) return fail(GENLIB_EINVAL, "sparse_phi schedule with guest rows is not supported");
                dim3 mgrid((unsigned)a.own_nm, (unsigned)std::min<int64_t>((L.n_new + 4 * kThreads - 1) / (4 * kThreads), 65535));
                misfile_kernel<T><<<mgrid, kThreads, 0, E.stream>>>(A, ld, E.mem_rank.p + L.mem_off, a);
                launches++;
            }
        }
        if (timed) CU(cudaEventRecord(E.events[ev++], E.stream));
        launch_barrier(E);                 // all new rows exist everywhere before the next layer reads them
        if (timed) CU(cudaEventRecord(E.events[ev++], E.stream));
    }
    CU(cudaGetLastError());
    E.stats.kernel_launches = launches;
    return GENLIB_OK;
}

// Rows of this rank's probands (own_pro order), all proband columns, streamed to host memory.
template <typename T, typename O>
int fetch_rows(genlib_engine &E, O *out) {
    const Plan &P = E.plan->p;
    const int32_t n = P.n_unique, nown = (int32_t)E.own_pro.size();
    if (n == 0 || nown == 0) return GENLIB_OK;
    // stream row blocks through two staging buffers so the gather of block b+1
    // overlaps the D2H copy of block b
    const size_t row_bytes = (size_t)n * sizeof(O);
    int32_t rows_per = (int32_t)std::max<size_t>(1, std::min<size_t>((size_t)nown, kFetchStageBytes / row_bytes));
    if ((size_t)rows_per * row_bytes > kFetchStageBytes) return fail(GENLIB_EINVAL, "proband row does not fit the staging buffer");
    rows_per = std::min(rows_per, 65535);
    O *stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2], copied[2];
    for (int b = 0; b < 2; b++) {
        stage[b] = reinterpret_cast<O *>(E.fetch_stage[b]);
        CU(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
    }
    int rc = GENLIB_OK;
    int blk = 0;
    for (int32_t r0 = 0; r0 < nown; r0 += rows_per, blk++) {
        const int b = blk & 1;
        const int32_t nr = std::min(rows_per, nown - r0);
        if (blk >= 2 && cudaStreamWaitEvent(E.stream, copied[b], 0) != cudaSuccess) { rc = GENLIB_ECUDA; break; }
        dim3 grid((unsigned)std::min<int32_t>((n + 255) / 256, 64), (unsigned)nr);
        gather_kernel<T, O><<<grid, 256, 0, E.stream>>>(static_cast<const T *>(E.A), P.capacity, E.own_pro_row.p, E.pro_slot.p,
                                                       n, r0, nr, stage[b]);
        cudaEventRecord(done[b], E.stream);
        cudaStreamWaitEvent(E.copy_stream, done[b], 0);
        if (cudaMemcpyAsync(out + (size_t)r0 * n, stage[b], (size_t)nr * row_bytes, cudaMemcpyDeviceToHost, E.copy_stream) != cudaSuccess) { rc = GENLIB_ECUDA; break; }
        cudaEventRecord(copied[b], E.copy_stream);
    }
    cudaError_t e1 = cudaStreamSynchronize(E.stream), e2 = cudaStreamSynchronize(E.copy_stream);
    for (int b = 0; b < 2; b++) { cudaEventDestroy(done[b]); cudaEventDestroy(copied[b]); }
    if (rc != GENLIB_OK || e1 != cudaSuccess || e2 != cudaSuccess)
        return fail(GENLIB_ECUDA, std::string("proband fetch failed: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    E.stats.d2h_bytes += (int64_t)nown * (int64_t)row_bytes;
    return GENLIB_OK;
}

int create_engine(const genlib_plan *plan, int numerics, int device, int rank, genlib_engine **out) {
    if (!plan || !out) return fail(GENLIB_EINVAL, "genlib_engine_create: null argument");
    *out = nullptr;
    if (numerics != GENLIB_NUMERICS_REFERENCE && numerics != GENLIB_NUMERICS_FP64) return fail(GENLIB_EINVAL, "unknown numerics mode");
    const Plan &P = plan->p;
    if (sparse_schedule(P.schedule) && numerics != GENLIB_NUMERICS_REFERENCE)
        return fail(GENLIB_EINVAL, "the sparse_phi schedule stores Float32 (numerics must be GENLIB_NUMERICS_REFERENCE)");
    if (rank < 0 || rank >= P.world) return fail(GENLIB_EINVAL, "rank outside the plan's world");
    if (P.world > kMaxWorld) return fail(GENLIB_EINVAL, "the engine supports at most 16 ranks");
    if (P.n_unique == 0) return fail(GENLIB_EINVAL, "empty proband list: nothing to run");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(GENLIB_ECUDA, "no CUDA device: libgenlib_cuda has no CPU fallback");
    DeviceGuard guard;
    if (int rc = guard.enter(device)) return rc;
    std::unique_ptr<genlib_engine> E(new (std::nothrow) genlib_engine);
    if (!E) return fail(GENLIB_ENOMEM, "out of host memory");
    E->plan = plan; E->numerics = numerics; E->rank = rank; E->world = P.world;
    E->esize = numerics == GENLIB_NUMERICS_FP64 ? 8 : 4;
    CU(cudaGetDevice(&E->device));
    CU(cudaDeviceGetAttribute(&E->sm_count, cudaDevAttrMultiProcessorCount, E->device));
    const size_t need = engine_bytes(P, numerics, rank);
    const double t0 = now_ms();
    CU(cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&E->copy_stream, cudaStreamNonBlocking));
    {
        const char *env = std::getenv("GENLIB_PIPE");
        E->piped = E->world > 1 && env && env[0] == '1';
    }
    if (E->piped) {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // hi = numerically lowest = highest priority
        CU(cudaStreamCreateWithPriority(&E->side_stream, cudaStreamNonBlocking, hi));
        for (auto &e : E->group_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    {
        cudaError_t ce = g_arenas.acquire(E->device, need, E->arena);
        if (ce != cudaSuccess) {
            E->arena = Arena();
            cudaGetLastError();
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            char msg[256];
            std::snprintf(msg, sizeof msg, "rank %d needs %.2f GB on the device, %.2f GB free (capacity %lld slots, %lld rows): shard over more GPUs",
                          rank, need / 1e9, free_b / 1e9, (long long)P.capacity, (long long)P.rows_cap[rank]);
            return fail(GENLIB_ENOMEM, msg);
        }
        unsigned char *base = static_cast<unsigned char *>(E->arena.base), *cur = base;
        auto take = [&](size_t bytes) { unsigned char *p = cur; cur += pad256(bytes); return p; };
        E->bar_flags = reinterpret_cast<unsigned *>(take(kFlagBytes));
        E->A = take(total_rows(P, rank) * (size_t)P.capacity * E->esize);
        E->Vrow = take(P.rank_v_elems[rank] * E->esize);
        E->Vt = take(P.rank_v_elems[rank] * E->esize);
        E->Rt = reinterpret_cast<double *>(take(P.rank_rt_elems[rank] * sizeof(double)));
        E->Dg = take(P.fam_pf.size() * E->esize);
        E->fetch_stage[0] = take(kFetchStageBytes);
        E->fetch_stage[1] = take(kFetchStageBytes);
        E->mem_ind.place(cur, P.mem_ind.size()); E->mem_slot.place(cur, P.mem_slot.size()); E->mem_fam.place(cur, P.mem_fam.size());
        E->mem_lrow.place(cur, P.mem_lrow.size());
        E->fam_pf.place(cur, P.fam_pf.size()); E->fam_pm.place(cur, P.fam_pm.size());
        E->fam_pf_lrow.place(cur, P.fam_pf_lrow.size()); E->fam_pm_lrow.place(cur, P.fam_pm_lrow.size());
        E->fam_minrank.place(cur, P.fam_minrank.size()); E->fam_maxrank.place(cur, P.fam_maxrank.size());
        E->fam_start.place(cur, P.fam_start.size());
        E->mt_min.place(cur, P.mtile_minrank.size()); E->mt_max.place(cur, P.mtile_maxrank.size());
        E->mt_fam0.place(cur, P.mtile_fam0.size()); E->mt_nfam.place(cur, P.mtile_nfam.size());
        E->mt_m0.place(cur, P.mtile_m0.size()); E->mt_cnt.place(cur, P.mtile_cnt.size());
        E->pro_slot.place(cur, P.pro_slot.size()); E->own_pro_row.place(cur, P.pro_slot.size());
        E->live_lrow.place(cur, P.live_lrow.size());
        E->fam_pf_owner.place(cur, P.fam_pf_owner.size()); E->fam_pm_owner.place(cur, P.fam_pm_owner.size());
        E->live_owner.place(cur, P.live_owner.size());
        E->mem_gowner.place(cur, P.mem_gowner.size()); E->mem_glrow.place(cur, P.mem_glrow.size());
        E->mem_rank.place(cur, P.mem_rank.size());
        E->flags.place(cur, P.flags.size()); E->acc.place(cur, 2);
        if ((size_t)(cur - base) > need) return fail(GENLIB_EINVAL, "internal: arena layout overflow");
        if ((size_t)(static_cast<unsigned char *>(E->A) - base) != off_A() ||
            (size_t)(static_cast<unsigned char *>(E->Vrow) - base) != off_Vrow(P, E->esize, rank))
            return fail(GENLIB_EINVAL, "internal: peer-visible arena offsets drifted");
    }
    // this rank's probands, in output order
    std::vector<int32_t> own_rows;
    for (size_t u = 0; u < P.pro_ind.size(); u++)
        if (P.pro_owner[u] == rank) { E->own_pro.push_back((int32_t)u); own_rows.push_back(P.pro_lrow[u]); }
    CU(cudaMemsetAsync(E->bar_flags, 0, kFlagBytes, E->stream));
    CU(E->mem_ind.upload(P.mem_ind, E->stream));
    CU(E->mem_slot.upload(P.mem_slot, E->stream));
    CU(E->mem_fam.upload(P.mem_fam, E->stream));
    CU(E->mem_lrow.upload(P.mem_lrow, E->stream));
    CU(E->fam_pf.upload(P.fam_pf, E->stream));
    CU(E->fam_pm.upload(P.fam_pm, E->stream));
    CU(E->fam_pf_lrow.upload(P.fam_pf_lrow, E->stream));
    CU(E->fam_pm_lrow.upload(P.fam_pm_lrow, E->stream));
    CU(E->fam_pf_owner.upload(P.fam_pf_owner, E->stream));
    CU(E->fam_pm_owner.upload(P.fam_pm_owner, E->stream));
    CU(E->fam_start.upload(P.fam_start, E->stream));
    CU(E->fam_minrank.upload(P.fam_minrank, E->stream));
    CU(E->fam_maxrank.upload(P.fam_maxrank, E->stream));
    CU(E->mt_min.upload(P.mtile_minrank, E->stream));
    CU(E->mt_max.upload(P.mtile_maxrank, E->stream));
    CU(E->mt_fam0.upload(P.mtile_fam0, E->stream));
    CU(E->mt_nfam.upload(P.mtile_nfam, E->stream));
    CU(E->mt_m0.upload(P.mtile_m0, E->stream));
    CU(E->mt_cnt.upload(P.mtile_cnt, E->stream));
    CU(E->pro_slot.upload(P.pro_slot, E->stream));
    CU(E->own_pro_row.upload(own_rows, E->stream));
    CU(E->live_owner.upload(P.live_owner, E->stream));
    CU(E->live_lrow.upload(P.live_lrow, E->stream));
    CU(E->mem_gowner.upload(P.mem_gowner, E->stream));
    CU(E->mem_glrow.upload(P.mem_glrow, E->stream));
    CU(E->mem_rank.upload(P.mem_rank, E->stream));
    CU(E->flags.upload(P.flags, E->stream));
    CU(cudaStreamSynchronize(E->stream));
    E->peers.A[rank] = E->A; E->peers.Vrow[rank] = E->Vrow; E->bars.flags[rank] = E->bar_flags;
    E->attached = P.world == 1;
    E->info.resize(P.layers.size());
    for (size_t t = 0; t < P.layers.size(); t++) fill_info(P.layers[t], &E->info[t]);
    E->events.resize(P.layers.size() * 6 + 2);
    for (auto &e : E->events) CU(cudaEventCreate(&e));
    genlib_stats &s = E->stats;
    s.n_unique = P.n_unique; s.n_layers = (int32_t)P.layers.size(); s.row_updates = P.row_updates;
    s.capacity = P.capacity; s.device_bytes = (int64_t)need; s.alg_bytes = P.alg_elems * (double)E->esize;
    s.ms_plan = plan->ms_plan; s.ms_upload = now_ms() - t0;
    s.h2d_bytes = (int64_t)plan_index_bytes(P);
    *out = E.release();
    return GENLIB_OK;
}

}  // namespace

// ------------------------------------------------------------------------------ ABI
extern "C" {

int genlib_version(void) { return GENLIB_ABI_VERSION; }
const char *genlib_last_error(void) { return g_err.c_str(); }

int genlib_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { g_err = "cudaGetDeviceCount failed (no driver / no device)"; return -GENLIB_ECUDA; }
    return n;
}

int genlib_release_cache(void) {
    g_peer_maps.close_all();
    g_arenas.release_device(-1);
    release_plan_cache();
    return GENLIB_OK;
}

int genlib_pinned_alloc(size_t bytes, void **out) {
    if (!out) return fail(GENLIB_EINVAL, "null argument");
    *out = nullptr;
    CU(cudaHostAlloc(out, std::max<size_t>(bytes, 16), cudaHostAllocDefault));
    return GENLIB_OK;
}

int genlib_pinned_free(void *ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return GENLIB_OK;
}

int genlib_plan_create(int32_t n, const int32_t *father, const int32_t *mother, int32_t n_pro,
                       const int32_t *proband, int32_t world, genlib_plan **out) {
    return genlib_plan_create_scheduled(n, father, mother, n_pro, proband, world, GENLIB_SCHEDULE_PHI, out);
}

int genlib_plan_create_scheduled(int32_t n, const int32_t *father, const int32_t *mother, int32_t n_pro,
                                 const int32_t *proband, int32_t world, int schedule, genlib_plan **out) {
    return genlib_plan_create_ex(n, father, mother, nullptr, n_pro, proband, world, schedule, out);
}

int genlib_plan_create_ex(int32_t n, const int32_t *father, const int32_t *mother, const int64_t *ids, int32_t n_pro,
                          const int32_t *proband, int32_t world, int schedule, genlib_plan **out) {
    if (!out) return fail(GENLIB_EINVAL, "genlib_plan_create: out is null");
    *out = nullptr;
    std::unique_ptr<genlib_plan> pl(new (std::nothrow) genlib_plan);
    if (!pl) return fail(GENLIB_ENOMEM, "out of host memory");
    const double t0 = now_ms();
    std::string err;
    int rc;
    try {
        adopt_retired_storage(pl->p);        // the arrays of the last destroyed plan, already paged in
        rc = build_plan(n, father, mother, ids, n_pro, proband, world, schedule, pl->p, err);
    } catch (const std::bad_alloc &) {
        return fail(GENLIB_ENOMEM, "out of host memory while planning");
    }
    if (rc != GENLIB_OK) return fail(rc, err);
    pl->ms_plan = now_ms() - t0;
    *out = pl.release();
    return GENLIB_OK;
}

void genlib_plan_destroy(genlib_plan *plan) {
    if (!plan) return;
    try { retire_storage(plan->p); } catch (...) {}
    delete plan;
}
int32_t genlib_plan_n_unique(const genlib_plan *plan) { return plan ? plan->p.n_unique : -1; }
int32_t genlib_plan_schedule(const genlib_plan *plan) { return plan ? plan->p.schedule : -1; }
int32_t genlib_plan_n_layers(const genlib_plan *plan) { return plan ? (int32_t)plan->p.layers.size() : -1; }
int64_t genlib_plan_capacity(const genlib_plan *plan) { return plan ? plan->p.capacity : -1; }
int64_t genlib_plan_row_updates(const genlib_plan *plan) { return plan ? plan->p.row_updates : -1; }

int genlib_plan_layer_info(const genlib_plan *plan, int32_t layer, genlib_layer_info *out) {
    if (!plan || !out || layer < 0 || layer >= (int32_t)plan->p.layers.size()) return fail(GENLIB_EINVAL, "bad layer");
    fill_info(plan->p.layers[layer], out);
    return GENLIB_OK;
}

int64_t genlib_plan_device_bytes(const genlib_plan *plan, int numerics, int32_t rank) {
    if (!plan || rank < 0 || rank >= plan->p.world || plan->p.n_unique == 0) return plan ? 0 : -1;
    return (int64_t)engine_bytes(plan->p, numerics, rank);
}

int genlib_plan_layer_arrays(const genlib_plan *plan, int32_t layer, int32_t *member_ind,
                             int32_t *member_slot, int32_t *member_fam, int32_t *fam_father_slot,
                             int32_t *fam_mother_slot, int32_t *member_owner) {
    if (!plan || layer < 0 || layer >= (int32_t)plan->p.layers.size()) return fail(GENLIB_EINVAL, "bad layer");
    const Plan &P = plan->p;
    const Layer &L = P.layers[layer];
    for (int32_t q = 0; q < L.n_new; q++) {
        if (member_ind) member_ind[q] = P.mem_ind[L.mem_off + q];
        if (member_slot) member_slot[q] = P.mem_slot[L.mem_off + q];
        if (member_fam) member_fam[q] = P.mem_fam[L.mem_off + q];
        if (member_owner) {
            int g = 0;
            while (g + 1 < P.world && q >= P.mem_base[L.base_off + g + 1]) g++;
            member_owner[q] = g;
        }
    }
    for (int32_t f = 0; f < L.n_fam; f++) {
        if (fam_father_slot) fam_father_slot[f] = P.fam_pf[L.fam_off + f];
        if (fam_mother_slot) fam_mother_slot[f] = P.fam_pm[L.fam_off + f];
    }
    return GENLIB_OK;
}

int genlib_plan_layer_ranks(const genlib_plan *plan, int32_t layer, int32_t *member_rank) {
    if (!plan || !member_rank || layer < 0 || layer >= (int32_t)plan->p.layers.size()) return fail(GENLIB_EINVAL, "bad layer");
    const Plan &P = plan->p;
    const Layer &L = P.layers[layer];
    for (int32_t q = 0; q < L.n_new; q++)
        member_rank[q] = P.mem_rank.empty() ? P.mem_ind[L.mem_off + q] : P.mem_rank[L.mem_off + q];
    return GENLIB_OK;
}

int genlib_plan_layer_shard(const genlib_plan *plan, int32_t layer, int32_t *fam_base, int32_t *mem_base,
                            int32_t *member_lrow, int32_t *fam_father_owner, int32_t *fam_father_lrow,
                            int32_t *fam_mother_owner, int32_t *fam_mother_lrow, int32_t *member_guest_owner,
                            int32_t *member_guest_lrow) {
    if (!plan || layer < 0 || layer >= (int32_t)plan->p.layers.size()) return fail(GENLIB_EINVAL, "bad layer");
    const Plan &P = plan->p;
    const Layer &L = P.layers[layer];
    for (int32_t g = 0; g <= P.world; g++) {
        if (fam_base) fam_base[g] = P.fam_base[L.base_off + g];
        if (mem_base) mem_base[g] = P.mem_base[L.base_off + g];
    }
    for (int32_t q = 0; q < L.n_new; q++) {
        if (member_lrow) member_lrow[q] = P.mem_lrow[L.mem_off + q];
        if (member_guest_owner) member_guest_owner[q] = P.mem_gowner[L.mem_off + q];
        if (member_guest_lrow) member_guest_lrow[q] = P.mem_glrow[L.mem_off + q];
    }
    for (int32_t f = 0; f < L.n_fam; f++) {
        if (fam_father_owner) fam_father_owner[f] = P.fam_pf_owner[L.fam_off + f];
        if (fam_father_lrow) fam_father_lrow[f] = P.fam_pf_lrow[L.fam_off + f];
        if (fam_mother_owner) fam_mother_owner[f] = P.fam_pm_owner[L.fam_off + f];
        if (fam_mother_lrow) fam_mother_lrow[f] = P.fam_pm_lrow[L.fam_off + f];
    }
    return GENLIB_OK;
}

int genlib_plan_layer_live_rows(const genlib_plan *plan, int32_t layer, int32_t *live_owner, int32_t *live_lrow) {
    if (!plan || !live_owner || !live_lrow || layer < 0 || layer >= (int32_t)plan->p.layers.size()) return fail(GENLIB_EINVAL, "bad layer");
    const Plan &P = plan->p;
    const Layer &L = P.layers[layer];
    for (int64_t s = 0; s < P.capacity; s++) { live_owner[s] = -1; live_lrow[s] = -1; }
    for (int32_t r = 0; r < L.rt_rows; r++)
        if (P.flags[L.flag_off + r] & kFlagLive) {
            live_owner[L.rt_lo + r] = P.live_owner[L.flag_off + r];
            live_lrow[L.rt_lo + r] = P.live_lrow[L.flag_off + r];
        }
    return GENLIB_OK;
}

int64_t genlib_plan_rank_rows(const genlib_plan *plan, int32_t rank) {
    if (!plan || rank < 0 || rank >= plan->p.world) return -1;
    return plan->p.rows_cap.empty() ? 0 : (int64_t)total_rows(plan->p, rank);
}

int32_t genlib_plan_world(const genlib_plan *plan) { return plan ? plan->p.world : -1; }

int genlib_plan_proband_rows(const genlib_plan *plan, int32_t *owner, int32_t *lrow) {
    if (!plan || !owner || !lrow) return fail(GENLIB_EINVAL, "null argument");
    for (size_t u = 0; u < plan->p.pro_ind.size(); u++) { owner[u] = plan->p.pro_owner[u]; lrow[u] = plan->p.pro_lrow[u]; }
    return GENLIB_OK;
}

int genlib_plan_layer_flags(const genlib_plan *plan, int32_t layer, uint8_t *live_flags) {
    if (!plan || !live_flags || layer < 0 || layer >= (int32_t)plan->p.layers.size()) return fail(GENLIB_EINVAL, "bad layer");
    const Plan &P = plan->p;
    const Layer &L = P.layers[layer];
    std::memset(live_flags, 0, (size_t)P.capacity);
    for (int32_t r = 0; r < L.rt_rows; r++) live_flags[L.rt_lo + r] = P.flags[L.flag_off + r];
    return GENLIB_OK;
}

int genlib_plan_proband_slots(const genlib_plan *plan, int32_t *slots) {
    if (!plan || !slots) return fail(GENLIB_EINVAL, "null argument");
    std::copy(plan->p.pro_slot.begin(), plan->p.pro_slot.end(), slots);
    return GENLIB_OK;
}

int genlib_engine_create(const genlib_plan *plan, int numerics, int device, genlib_engine **out) {
    if (plan && plan->p.world != 1) return fail(GENLIB_EINVAL, "plan was built for several ranks: use genlib_engine_create_dist");
    return create_engine(plan, numerics, device, 0, out);
}

int genlib_engine_create_dist(const genlib_plan *plan, int numerics, int device, int32_t rank, genlib_engine **out) {
    return create_engine(plan, numerics, device, rank, out);
}

int genlib_engine_ipc_export(genlib_engine *eng, void *handle64) {
    if (!eng || !handle64) return fail(GENLIB_EINVAL, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard guard;
    if (int rc = guard.enter(eng->device)) return rc;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, eng->arena.base));
    std::memcpy(handle64, &h, sizeof h);
    return GENLIB_OK;
}

int genlib_engine_ipc_attach(genlib_engine *eng, const void *handles, size_t stride) {
    if (!eng || !handles || stride < 64) return fail(GENLIB_EINVAL, "null argument");
    if (eng->attached) return GENLIB_OK;
    DeviceGuard guard;
    if (int rc = guard.enter(eng->device)) return rc;
    const Plan &P = eng->plan->p;
    for (int g = 0; g < eng->world; g++) {
        if (g == eng->rank) continue;
        void *base = nullptr;
        cudaError_t ce = g_peer_maps.map(eng->device, static_cast<const unsigned char *>(handles) + (size_t)g * stride, &base);
        if (ce != cudaSuccess) {
            cudaGetLastError();
            return fail(GENLIB_ECOMM, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(g) + "): " + cudaGetErrorString(ce));
        }
        eng->peer_base[g] = base;
        unsigned char *b = static_cast<unsigned char *>(base);
        eng->bars.flags[g] = reinterpret_cast<unsigned *>(b);
        eng->peers.A[g] = b + off_A();
        eng->peers.Vrow[g] = b + off_Vrow(P, eng->esize, g);
    }
    eng->attached = true;
    return GENLIB_OK;
}

int32_t genlib_engine_own_probands(const genlib_engine *eng, int32_t *index) {
    if (!eng) return -1;
    if (index) std::copy(eng->own_pro.begin(), eng->own_pro.end(), index);
    return (int32_t)eng->own_pro.size();
}

void genlib_engine_destroy(genlib_engine *eng) {
    if (!eng) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(eng->device);
    delete eng;
    if (prev >= 0) cudaSetDevice(prev);
}

int genlib_engine_run(genlib_engine *eng, int time_layers) {
    if (!eng) return fail(GENLIB_EINVAL, "null engine");
    DeviceGuard guard;
    if (int rc = guard.enter(eng->device)) return rc;
    genlib_engine &E = *eng;
    if (!E.attached) return fail(GENLIB_ECOMM, "genlib_engine_run before genlib_engine_ipc_attach");
    const size_t nev = E.events.size();
    CU(cudaEventRecord(E.events[nev - 2], E.stream));
    int rc = E.numerics == GENLIB_NUMERICS_FP64 ? launch_layers<double>(E, time_layers != 0)
                                                : launch_layers<float>(E, time_layers != 0);
    if (rc != GENLIB_OK) return rc;
    CU(cudaEventRecord(E.events[nev - 1], E.stream));
    CU(cudaStreamSynchronize(E.stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, E.events[nev - 2], E.events[nev - 1]));
    E.stats.ms_kernels = ms;
    if (time_layers) {
        size_t ev = 0;
        for (size_t t = 0; t < E.info.size(); t++) {
            if (E.info[t].n_new == 0) continue;
            if (E.layer_limit >= 0 && (int32_t)t >= E.layer_limit) break;
            float a = 0, b = 0, w1 = 0, c = 0, w2 = 0;
            CU(cudaEventElapsedTime(&a, E.events[ev], E.events[ev + 1]));
            CU(cudaEventElapsedTime(&b, E.events[ev + 1], E.events[ev + 2]));
            CU(cudaEventElapsedTime(&w1, E.events[ev + 2], E.events[ev + 3]));
            CU(cudaEventElapsedTime(&c, E.events[ev + 3], E.events[ev + 4]));
            CU(cudaEventElapsedTime(&w2, E.events[ev + 4], E.events[ev + 5]));
            ev += 6;
            E.info[t].ms_cross = a; E.info[t].ms_couple = b; E.info[t].ms_expand = c;
            E.info[t].ms_wait = w1 + w2;
        }
    }
    if (E.world > 1) {
        unsigned errw = 0;
        CU(cudaMemcpy(&errw, E.bar_flags + kMaxWorld, sizeof errw, cudaMemcpyDeviceToHost));
        if (errw) return fail(GENLIB_ECOMM, "a rank did not reach the inter-GPU barrier within the time limit");
    }
    E.ran = true;
    return GENLIB_OK;
}

int genlib_engine_layer_info(const genlib_engine *eng, int32_t layer, genlib_layer_info *out) {
    if (!eng || !out || layer < 0 || layer >= (int32_t)eng->info.size()) return fail(GENLIB_EINVAL, "bad layer");
    *out = eng->info[layer];
    return GENLIB_OK;
}

int genlib_engine_stats(const genlib_engine *eng, genlib_stats *out) {
    if (!eng || !out) return fail(GENLIB_EINVAL, "null argument");
    *out = eng->stats;
    return GENLIB_OK;
}

int genlib_engine_fetch(genlib_engine *eng, void *out, int out_dtype) {
    if (!eng || (!out && eng->plan->p.n_unique > 0)) return fail(GENLIB_EINVAL, "null argument");
    if (!eng->ran) return fail(GENLIB_EINVAL, "genlib_engine_fetch before genlib_engine_run");
    if (out_dtype != GENLIB_F32 && out_dtype != GENLIB_F64) return fail(GENLIB_EINVAL, "unknown out_dtype");
    DeviceGuard guard;
    if (int rc = guard.enter(eng->device)) return rc;
    const double t0 = now_ms();
    int rc;
    if (eng->numerics == GENLIB_NUMERICS_FP64)
        rc = out_dtype == GENLIB_F64 ? fetch_rows<double, double>(*eng, (double *)out) : fetch_rows<double, float>(*eng, (float *)out);
    else
        rc = out_dtype == GENLIB_F64 ? fetch_rows<float, double>(*eng, (double *)out) : fetch_rows<float, float>(*eng, (float *)out);
    eng->stats.ms_fetch = now_ms() - t0;
    return rc;
}

int genlib_engine_phi_mean(genlib_engine *eng, double *out) {
    if (!eng || !out) return fail(GENLIB_EINVAL, "null argument");
    if (!eng->ran) return fail(GENLIB_EINVAL, "genlib_engine_phi_mean before genlib_engine_run");
    if (eng->world != 1) return fail(GENLIB_EINVAL, "genlib_engine_phi_mean: single-rank engines only");
    DeviceGuard guard;
    if (int rc = guard.enter(eng->device)) return rc;
    const Plan &P = eng->plan->p;
    const int32_t n = P.n_unique;
    if (n < 2) { *out = 0.0; return GENLIB_OK; }
    CU(cudaMemsetAsync(eng->acc.p, 0, 2 * sizeof(double), eng->stream));
    const unsigned grid = (unsigned)std::min<int32_t>(n, 148 * 8);
    if (eng->numerics == GENLIB_NUMERICS_FP64)
        mean_kernel<double><<<grid, kThreads, 0, eng->stream>>>((const double *)eng->A, P.capacity, eng->pro_slot.p, n, eng->acc.p);
    else
        mean_kernel<float><<<grid, kThreads, 0, eng->stream>>>((const float *)eng->A, P.capacity, eng->pro_slot.p, n, eng->acc.p);
    double h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, eng->acc.p, sizeof h, cudaMemcpyDeviceToHost, eng->stream));
    CU(cudaStreamSynchronize(eng->stream));
    *out = (h[0] - h[1]) / ((double)n * n - n);
    return GENLIB_OK;
}

int genlib_engine_set_layer_limit(genlib_engine *eng, int32_t n_layers) {
    if (!eng) return fail(GENLIB_EINVAL, "null engine");
    eng->layer_limit = n_layers;
    return GENLIB_OK;
}

int genlib_engine_read_block(genlib_engine *eng, int32_t n_slots, const int32_t *slots, double *out) {
    if (!eng || !slots || !out || n_slots < 0) return fail(GENLIB_EINVAL, "null argument");
    if (eng->world != 1) return fail(GENLIB_EINVAL, "genlib_engine_read_block: single-rank engines only");
    DeviceGuard guard;
    if (int rc = guard.enter(eng->device)) return rc;
    if (n_slots == 0) return GENLIB_OK;
    const Plan &P = eng->plan->p;
    for (int32_t i = 0; i < n_slots; i++)
        if (slots[i] < 0 || slots[i] >= P.capacity) return fail(GENLIB_EINVAL, "slot out of range");
    int32_t *dslots = nullptr;
    double *dout = nullptr;
    CU(cudaMalloc(&dslots, (size_t)n_slots * sizeof(int32_t)));
    CU(cudaMalloc(&dout, (size_t)n_slots * n_slots * sizeof(double)));
    CU(cudaMemcpyAsync(dslots, slots, (size_t)n_slots * sizeof(int32_t), cudaMemcpyHostToDevice, eng->stream));
    for (int32_t r0 = 0; r0 < n_slots; r0 += 65535) {
        const int32_t nr = std::min(65535, n_slots - r0);
        dim3 grid((unsigned)std::min<int32_t>((n_slots + 255) / 256, 64), (unsigned)nr);
        if (eng->numerics == GENLIB_NUMERICS_FP64)
            gather_kernel<double, double><<<grid, 256, 0, eng->stream>>>((const double *)eng->A, P.capacity, dslots, dslots, n_slots, r0, nr, dout + (size_t)r0 * n_slots);
        else
            gather_kernel<float, double><<<grid, 256, 0, eng->stream>>>((const float *)eng->A, P.capacity, dslots, dslots, n_slots, r0, nr, dout + (size_t)r0 * n_slots);
    }
    CU(cudaMemcpyAsync(out, dout, (size_t)n_slots * n_slots * sizeof(double), cudaMemcpyDeviceToHost, eng->stream));
    CU(cudaStreamSynchronize(eng->stream));
    cudaFree(dslots); cudaFree(dout);
    return GENLIB_OK;
}

int genlib_phi(int32_t n, const int32_t *father, const int32_t *mother, int32_t n_pro,
               const int32_t *proband, void *out, int out_dtype, int numerics, int device,
               genlib_stats *stats) {
    genlib_plan *plan = nullptr;
    int rc = genlib_plan_create(n, father, mother, n_pro, proband, 1, &plan);
    if (rc != GENLIB_OK) return rc;
    std::unique_ptr<genlib_plan> pguard(plan);
    if (plan->p.n_unique == 0) {
        if (stats) { std::memset(stats, 0, sizeof *stats); stats->ms_plan = plan->ms_plan; }
        return GENLIB_OK;                       // 0 x 0 matrix, like the reference
    }
    if (!out) return fail(GENLIB_EINVAL, "genlib_phi: out is null");
    genlib_engine *eng = nullptr;
    rc = genlib_engine_create(plan, numerics, device, &eng);
    if (rc != GENLIB_OK) return rc;
    rc = genlib_engine_run(eng, 0);
    if (rc == GENLIB_OK) rc = genlib_engine_fetch(eng, out, out_dtype);
    if (rc == GENLIB_OK && stats) *stats = eng->stats;
    genlib_engine_destroy(eng);
    return rc;
}

}  // extern "C"
