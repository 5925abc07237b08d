// loader.cpp -- the input side of the path: pedigree file / columns ->
// rank-ordered flat arrays, fast enough that 5 M-row files are not the new bottleneck.
//
// Mirrors (reference, relative to /root/reference):
//   src/create.jl:161-189   genealogy(::String): first line skipped, then whitespace-separated
//                           `ind father mother sex`, 0 = unknown parent
//   src/create.jl:131-146   genealogy(::DataFrame)
//   src/create.jl:196-209   _max_depth!  (founders = 1; here iterative, no recursion limit)
//   src/create.jl:217-227   _ordered_pedigree: STABLE sort by depth (ties keep file order)
//   src/create.jl:234-254   _finalize_pedigree: rank = position; a parent must already exist
//   src/identify.jl:35-39   pro: IDs without children, ascending
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/genlib_cuda.h"
#include "plan.hpp"

// ID -> index.  Pedigree IDs are usually small positive integers: then a direct table (4 bytes per
// possible ID) replaces hashing; otherwise open addressing with one 16-byte record per bucket and
// software prefetch (5 M individuals: 1.4 s with two plain hash maps -> 0.4 s).
struct GenlibIdMap {
    struct Bucket { int64_t key; int32_t val, pad; };
    std::vector<int32_t> direct;             // direct[id] = index or -1 (ids in [0, direct.size()))
    std::vector<Bucket> table;
    uint64_t mask = 0;
    static uint64_t mix(uint64_t x) {
        x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33;
        x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
    }
    bool build(const int64_t *ids, size_t n);                  // false on a duplicate ID
    void prefetch(int64_t id) const {
        if (!direct.empty()) { if (id >= 0 && (uint64_t)id < direct.size()) __builtin_prefetch(&direct[(size_t)id]); }
        else if (!table.empty()) __builtin_prefetch(&table[mix((uint64_t)id) & mask]);
    }
    int32_t get(int64_t id) const {
        if (!direct.empty()) return id >= 0 && (uint64_t)id < direct.size() ? direct[(size_t)id] : -1;
        if (table.empty()) return -1;
        uint64_t h = mix((uint64_t)id) & mask;
        while (table[h].val >= 0) { if (table[h].key == id) return table[h].val; h = (h + 1) & mask; }
        return -1;
    }
};

struct genlib_pedigree {
    std::vector<int64_t> id;                 // by rank
    std::vector<int32_t> father, mother, sex, nchild;
    int32_t depth = 0;                       // number of generations (create.jl:196-209)
    GenlibIdMap by_file;                     // ID -> position in the input
    std::vector<int32_t> pos;                // position in the input -> rank
};

namespace {

struct ErrSink {                       // assignment records the message for genlib_last_error()
    std::string last;
    ErrSink &operator=(const std::string &m) { last = m; genlib::set_error(0, m); return *this; }
    ErrSink &operator=(const char *m) { return *this = std::string(m); }
};
thread_local ErrSink g_lerr;

constexpr int kAhead = 16;                 // software prefetch distance of the ID look-ups

}  // namespace

bool GenlibIdMap::build(const int64_t *ids, size_t n) {
    int64_t lo = 0, hi = -1;
    for (size_t i = 0; i < n; i++) { lo = std::min(lo, ids[i]); hi = std::max(hi, ids[i]); }
    if (lo >= 0 && (uint64_t)hi < 16 * (uint64_t)n + 1024) {       // dense enough: direct table
        direct.assign((size_t)hi + 1, -1);
        for (size_t i = 0; i < n; i++) {
            int32_t &slot = direct[(size_t)ids[i]];
            if (slot >= 0) return false;
            slot = (int32_t)i;
        }
        return true;
    }
    size_t cap = 16;
    while (cap < 2 * n + 2) cap <<= 1;
    table.assign(cap, Bucket{0, -1, 0}); mask = cap - 1;
    for (size_t i = 0; i < n; i++) {
        if (i + kAhead < n) __builtin_prefetch(&table[mix((uint64_t)ids[i + kAhead]) & mask], 1);
        uint64_t h = mix((uint64_t)ids[i]) & mask;
        while (table[h].val >= 0) { if (table[h].key == ids[i]) return false; h = (h + 1) & mask; }
        table[h].key = ids[i]; table[h].val = (int32_t)i;
    }
    return true;
}

namespace {

int build(size_t n, const int64_t *ind, const int64_t *fid, const int64_t *mid, const int32_t *sex, int sort,
          genlib_pedigree **out) {
    if (n > 0x7ffffff0u) { g_lerr = "pedigree too large"; return GENLIB_EINVAL; }
    GenlibIdMap byfile;
    if (!byfile.build(ind, n)) { g_lerr = "duplicate individual ID"; return GENLIB_EINVAL; }
    std::vector<int32_t> f(n), m(n);
    for (size_t i = 0; i < n; i++) {
        if (i + kAhead < n) { byfile.prefetch(fid[i + kAhead]); byfile.prefetch(mid[i + kAhead]); }
        f[i] = fid[i] == 0 ? -1 : byfile.get(fid[i]);
        m[i] = mid[i] == 0 ? -1 : byfile.get(mid[i]);
        if ((fid[i] != 0 && f[i] < 0) || (mid[i] != 0 && m[i] < 0)) {
            g_lerr = "KeyError: parent " + std::to_string(fid[i] != 0 && f[i] < 0 ? fid[i] : mid[i]) + " of individual " +
                     std::to_string(ind[i]) + " is not in the pedigree";
            return GENLIB_EKEY;
        }
    }
    std::vector<int32_t> order(n);
    if (sort) {
        // depth = 1 + max(depth of parents), memoised, with an explicit stack
        std::vector<int32_t> depth(n, 0), stack;
        int32_t maxd = 0;
        for (size_t s = 0; s < n; s++) {
            if (depth[s]) continue;
            depth[s] = -1;                                   // -1 = on the stack
            stack.push_back((int32_t)s);
            while (!stack.empty()) {
                const int32_t x = stack.back();
                const int32_t a = f[x], b = m[x];
                if (a >= 0 && depth[a] == 0) { depth[a] = -1; stack.push_back(a); continue; }
                if (b >= 0 && depth[b] == 0) { depth[b] = -1; stack.push_back(b); continue; }
                const int32_t da = a >= 0 ? depth[a] : 0, db = b >= 0 ? depth[b] : 0;
                if (da < 0 || db < 0) { g_lerr = "pedigree contains a cycle"; return GENLIB_EORDER; }
                depth[x] = std::max(da, db) + 1;
                maxd = std::max(maxd, depth[x]);
                stack.pop_back();
            }
        }
        std::vector<int32_t> cnt((size_t)maxd + 2, 0);               // stable counting sort by depth
        for (size_t i = 0; i < n; i++) cnt[depth[i] + 1]++;
        for (int32_t d = 1; d <= maxd + 1; d++) cnt[d] += cnt[d - 1];
        for (size_t i = 0; i < n; i++) order[cnt[depth[i]]++] = (int32_t)i;
    } else {
        for (size_t i = 0; i < n; i++) order[i] = (int32_t)i;
    }
    std::vector<int32_t> pos(n);
    for (size_t r = 0; r < n; r++) pos[order[r]] = (int32_t)r;
    std::unique_ptr<genlib_pedigree> P(new genlib_pedigree);
    P->id.resize(n); P->father.resize(n); P->mother.resize(n); P->sex.resize(n); P->nchild.assign(n, 0);
    for (size_t r = 0; r < n; r++) {
        const int32_t i = order[r];
        P->id[r] = ind[i];
        P->father[r] = f[i] < 0 ? -1 : pos[f[i]];
        P->mother[r] = m[i] < 0 ? -1 : pos[m[i]];
        P->sex[r] = sex ? sex[i] : 0;
        if (P->father[r] >= (int32_t)r || P->mother[r] >= (int32_t)r) {      // pedigree[father] not defined yet
            g_lerr = "KeyError: a parent of individual " + std::to_string(ind[i]) + " does not precede it (sort = false?)";
            return GENLIB_EKEY;
        }
        if (P->father[r] >= 0) P->nchild[P->father[r]]++;
        if (P->mother[r] >= 0) P->nchild[P->mother[r]]++;
    }
    {   // generations: parents precede children in rank order, one forward sweep
        std::vector<int32_t> d(n, 1);
        for (size_t r = 0; r < n; r++) {
            if (P->father[r] >= 0) d[r] = std::max(d[r], d[P->father[r]] + 1);
            if (P->mother[r] >= 0) d[r] = std::max(d[r], d[P->mother[r]] + 1);
            P->depth = std::max(P->depth, d[r]);
        }
    }
    P->by_file = std::move(byfile);                          // ID -> rank = pos[by_file.get(ID)]
    P->pos.swap(pos);
    *out = P.release();
    return GENLIB_OK;
}

}  // namespace

extern "C" {

int genlib_genealogy_arrays(int64_t n, const int64_t *ind, const int64_t *father, const int64_t *mother,
                            const int32_t *sex, int sort, genlib_pedigree **out) {
    if (!out || n < 0 || (n > 0 && (!ind || !father || !mother))) { g_lerr = "null argument"; return GENLIB_EINVAL; }
    *out = nullptr;
    try {
        return build((size_t)n, ind, father, mother, sex, sort, out);
    } catch (const std::bad_alloc &) { g_lerr = "out of host memory"; return GENLIB_ENOMEM; }
}

int genlib_genealogy_csv(const char *path, int sort, genlib_pedigree **out) {
    if (!path || !out) { g_lerr = "null argument"; return GENLIB_EINVAL; }
    *out = nullptr;
    FILE *fp = std::fopen(path, "rb");
    if (!fp) { g_lerr = std::string("cannot open ") + path; return GENLIB_EINVAL; }
    std::fseek(fp, 0, SEEK_END);
    const long sz = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::vector<char> buf((size_t)std::max<long>(sz, 0) + 1);
    const size_t got = sz > 0 ? std::fread(buf.data(), 1, (size_t)sz, fp) : 0;
    std::fclose(fp);
    buf[got] = '\n';
    const char *p = buf.data(), *end = buf.data() + got;
    while (p < end && *p != '\n') p++;                      // the first line is skipped (create.jl:167-170)
    std::vector<int64_t> ind, fa, mo;
    std::vector<int32_t> sx;
    ind.reserve(got / 16); fa.reserve(got / 16); mo.reserve(got / 16); sx.reserve(got / 16);
    while (p < end) {
        int64_t v[4];
        int k = 0;
        while (k < 4) {
            while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
            if (p >= end || *p == '\n') break;
            bool neg = false;
            if (*p == '-') { neg = true; p++; }
            if (p >= end || *p < '0' || *p > '9') { g_lerr = "malformed pedigree file: expected an integer"; return GENLIB_EINVAL; }
            int64_t x = 0;
            while (p < end && *p >= '0' && *p <= '9') x = x * 10 + (*p++ - '0');
            v[k++] = neg ? -x : x;
        }
        while (p < end && *p != '\n') p++;                  // rest of the line
        p++;
        if (k == 0) continue;                                // blank line
        if (k != 4) { g_lerr = "malformed pedigree file: a line does not have 4 fields"; return GENLIB_EINVAL; }
        ind.push_back(v[0]); fa.push_back(v[1]); mo.push_back(v[2]); sx.push_back((int32_t)v[3]);
    }
    return genlib_genealogy_arrays((int64_t)ind.size(), ind.data(), fa.data(), mo.data(), sx.data(), sort, out);
}

void genlib_pedigree_destroy(genlib_pedigree *ped) { delete ped; }
int64_t genlib_pedigree_n(const genlib_pedigree *ped) { return ped ? (int64_t)ped->id.size() : -1; }
int32_t genlib_pedigree_depth(const genlib_pedigree *ped) { return ped ? ped->depth : -1; }

int genlib_pedigree_arrays(const genlib_pedigree *ped, int64_t *ids, int32_t *father, int32_t *mother, int32_t *sex) {
    if (!ped) { g_lerr = "null pedigree"; return GENLIB_EINVAL; }
    const size_t n = ped->id.size();
    if (ids) std::memcpy(ids, ped->id.data(), n * sizeof(int64_t));
    if (father) std::memcpy(father, ped->father.data(), n * sizeof(int32_t));
    if (mother) std::memcpy(mother, ped->mother.data(), n * sizeof(int32_t));
    if (sex) std::memcpy(sex, ped->sex.data(), n * sizeof(int32_t));
    return GENLIB_OK;
}

int64_t genlib_pedigree_pro(const genlib_pedigree *ped, int64_t *out) {
    if (!ped) return -1;
    int64_t k = 0;
    for (size_t r = 0; r < ped->id.size(); r++)
        if (ped->nchild[r] == 0) { if (out) out[k] = ped->id[r]; k++; }
    if (out) std::sort(out, out + k);
    return k;
}

int genlib_pedigree_ranks(const genlib_pedigree *ped, int64_t n, const int64_t *ids, int32_t *ranks) {
    if (!ped || (n > 0 && (!ids || !ranks))) { g_lerr = "null argument"; return GENLIB_EINVAL; }
    for (int64_t i = 0; i < n; i++) {
        const int32_t at = ped->by_file.get(ids[i]);
        if (at < 0) { g_lerr = "KeyError: " + std::to_string(ids[i]); return GENLIB_EKEY; }
        ranks[i] = ped->pos[(size_t)at];
    }
    return GENLIB_OK;
}

}  // extern "C"
