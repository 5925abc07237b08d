/*
 * genlib_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, CPU restatement of the one GenLib.jl path this repository
 * accelerates: gen.genealogy -> gen.pro -> gen.phi(ped, probands).  It is the
 * checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline / --impl reference leg).  Nothing under genlib.jl_b200/ may
 * include, link or call it; the product path has no CPU fallback.
 *
 * It follows the reference's ALGORITHM (cut vertices, pair-by-pair recursion
 * with founder_index, Float64 accumulator, Float32 store), not the engine's
 * level-synchronous formulation, so that agreement between the two is evidence
 * and not a tautology.  Reference lines (relative to /root/reference):
 *
 *   src/create.jl:161-189   CSV parse                     -> parse_csv()
 *   src/create.jl:196-209   _max_depth!                   -> max_depth()
 *   src/create.jl:217-227   _ordered_pedigree (stable)    -> build_ped()
 *   src/create.jl:234-254   _finalize_pedigree (rank)     -> build_ped()
 *   src/identify.jl:35-39   pro                           -> oracle_pro()
 *   src/compute.jl:66-95    phi(ind, ind) Karigl          -> pair_phi()
 *   src/compute.jl:105-158  phi(ind, ind, Psi)            -> cut_phi()
 *   src/compute.jl:193-207  _previous_generation          -> previous_generation()
 *   src/compute.jl:236-251  cut vertices                  -> build_cuts()
 *   src/compute.jl:271-302  step loop, threaded pair loop -> oracle_phi_ranks()
 *   src/compute.jl:454-459  phiMean                       -> oracle_phi_mean()
 *   src/compute.jl:321-447  sparse_phi ("next" row N2)     -> oracle_sparse_phi_ranks()
 *
 * Parity pinning: geneaJi is pinned by the reference's own known-answer test
 * (test/runtests.jl:47-53, checked in tests/test_oracle.py).  genea140 has no
 * kinship value in the reference's tests; it is pinned by the survey's
 * independent NumPy emulation checksum and by an
 * exact-rational Karigl recursion on sampled pairs -- i.e. "pinned by
 * cross-restatement", not by the reference itself (Julia is not installable
 * here).
 *
 * Arithmetic contract: every '+' is one IEEE binary64
 * round-to-nearest addition in the order written, x/2 is exact, the store into
 * the step matrix is binary64 -> binary32 round-to-nearest-even with gradual
 * underflow.  Compile WITHOUT -ffast-math.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>

#define ORACLE_OK 0
#define ORACLE_EKEY 1   /* unknown ID: the reference raises KeyError (create.jl:70) */
#define ORACLE_EIO 2
#define ORACLE_EORDER 3 /* parent does not precede child */
#define ORACLE_ENOMEM 4

typedef struct oracle_ped {
    int n;
    int64_t *id;      /* by rank (0-based position = rank-1 of create.jl:244) */
    int32_t *father;  /* 0-based rank of the father, -1 = nothing */
    int32_t *mother;
    int32_t *sex;
    int32_t *nchild;  /* length of .children (create.jl:246-251) */
    /* ID -> rank hash (open addressing) */
    int64_t *hkey;
    int32_t *hval;
    uint64_t hmask;
} oracle_ped;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---------------------------------------------------------------- hashing */
static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
static int hash_build(int n, const int64_t *keys, int64_t **hk, int32_t **hv, uint64_t *mask) {
    uint64_t cap = 16; while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
    int64_t *k = malloc(cap * sizeof *k); int32_t *v = malloc(cap * sizeof *v);
    if (!k || !v) { free(k); free(v); return ORACLE_ENOMEM; }
    for (uint64_t i = 0; i < cap; i++) v[i] = -1;
    for (int i = 0; i < n; i++) {
        uint64_t h = mix64((uint64_t)keys[i]) & (cap - 1);
        while (v[h] >= 0 && k[h] != keys[i]) h = (h + 1) & (cap - 1);
        k[h] = keys[i]; v[h] = i;     /* later duplicates overwrite, like Dict setindex! */
    }
    *hk = k; *hv = v; *mask = cap - 1; return ORACLE_OK;
}
static int hash_get(const int64_t *hk, const int32_t *hv, uint64_t mask, int64_t key) {
    uint64_t h = mix64((uint64_t)key) & mask;
    while (hv[h] >= 0) { if (hk[h] == key) return hv[h]; h = (h + 1) & mask; }
    return -1;
}

/* ------------------------------------------------ create.jl: the pedigree */

/* create.jl:196-209 -- memoised recursive depth; founders have depth 1. */
static int max_depth(int i, const int32_t *f, const int32_t *m, int32_t *depth) {
    if (depth[i] == -1) {
        int fd = 0, md = 0;
        if (f[i] >= 0) fd += max_depth(f[i], f, m, depth);
        if (m[i] >= 0) md += max_depth(m[i], f, m, depth);
        depth[i] = (fd > md ? fd : md) + 1;
    }
    return depth[i];
}

void oracle_free(oracle_ped *p) {
    if (!p) return;
    free(p->id); free(p->father); free(p->mother); free(p->sex); free(p->nchild);
    free(p->hkey); free(p->hval); free(p);
}

/* create.jl:131-146 + 217-254: file-order records -> rank-ordered pedigree.
 * ind/father/mother are IDs, 0 = unknown parent (create.jl:240-241). */
static oracle_ped *build_ped(int n, const int64_t *ind, const int64_t *fid, const int64_t *mid,
                             const int32_t *sex, int do_sort, int *status) {
    *status = ORACLE_OK;
    int64_t *hk = NULL; int32_t *hv = NULL; uint64_t mask = 0;
    if (hash_build(n, ind, &hk, &hv, &mask)) { *status = ORACLE_ENOMEM; return NULL; }
    int32_t *f = malloc((size_t)(n + 1) * sizeof *f), *m = malloc((size_t)(n + 1) * sizeof *m);
    int32_t *depth = malloc((size_t)(n + 1) * sizeof *depth), *order = malloc((size_t)(n + 1) * sizeof *order);
    int32_t *pos = malloc((size_t)(n + 1) * sizeof *pos);
    oracle_ped *p = calloc(1, sizeof *p);
    for (int i = 0; i < n; i++) {           /* file index of each parent */
        f[i] = fid[i] == 0 ? -1 : hash_get(hk, hv, mask, fid[i]);
        m[i] = mid[i] == 0 ? -1 : hash_get(hk, hv, mask, mid[i]);
        if ((fid[i] != 0 && f[i] < 0) || (mid[i] != 0 && m[i] < 0)) { *status = ORACLE_EKEY; goto fail; }
        depth[i] = -1;
    }
    if (do_sort) {
        int maxd = 0;
        for (int i = 0; i < n; i++) { int d = max_depth(i, f, m, depth); if (d > maxd) maxd = d; }
        /* sortperm(depths) is stable: counting sort keeps file order inside a depth */
        int32_t *cnt = calloc((size_t)maxd + 2, sizeof *cnt);
        for (int i = 0; i < n; i++) cnt[depth[i] + 1]++;
        for (int d = 1; d <= maxd + 1; d++) cnt[d] += cnt[d - 1];
        for (int i = 0; i < n; i++) order[cnt[depth[i]]++] = i;
        free(cnt);
    } else {
        for (int i = 0; i < n; i++) order[i] = i;
    }
    for (int r = 0; r < n; r++) pos[order[r]] = r;
    p->n = n;
    p->id = malloc((size_t)(n + 1) * sizeof *p->id);
    p->father = malloc((size_t)(n + 1) * sizeof *p->father);
    p->mother = malloc((size_t)(n + 1) * sizeof *p->mother);
    p->sex = malloc((size_t)(n + 1) * sizeof *p->sex);
    p->nchild = calloc((size_t)n + 1, sizeof *p->nchild);
    for (int r = 0; r < n; r++) {           /* create.jl:237-252 */
        int i = order[r];
        p->id[r] = ind[i];
        p->father[r] = f[i] < 0 ? -1 : pos[f[i]];
        p->mother[r] = m[i] < 0 ? -1 : pos[m[i]];
        p->sex[r] = sex ? sex[i] : 0;
        /* pedigree[individual.father] must already exist (KeyError otherwise) */
        if (p->father[r] >= r || p->mother[r] >= r) { *status = ORACLE_EORDER; goto fail; }
        if (p->father[r] >= 0) p->nchild[p->father[r]]++;
        if (p->mother[r] >= 0) p->nchild[p->mother[r]]++;
    }
    free(hk); free(hv); hk = NULL; hv = NULL;
    if (hash_build(n, p->id, &p->hkey, &p->hval, &p->hmask)) { *status = ORACLE_ENOMEM; goto fail; }
    free(f); free(m); free(depth); free(order); free(pos);
    return p;
fail:
    free(hk); free(hv); free(f); free(m); free(depth); free(order); free(pos);
    oracle_free(p);
    return NULL;
}

oracle_ped *oracle_genealogy_arrays(int n, const int64_t *ind, const int64_t *father,
                                    const int64_t *mother, const int32_t *sex, int do_sort,
                                    int *status) {
    return build_ped(n, ind, father, mother, sex, do_sort, status);
}

/* create.jl:161-189: skip the first line, then whitespace-separated
 * "ind father mother sex" per line. */
oracle_ped *oracle_genealogy_csv(const char *path, int do_sort, int *status) {
    FILE *fp = fopen(path, "r");
    if (!fp) { *status = ORACLE_EIO; return NULL; }
    size_t cap = 1 << 16, n = 0;
    int64_t *ind = malloc(cap * sizeof *ind), *fa = malloc(cap * sizeof *fa), *mo = malloc(cap * sizeof *mo);
    int32_t *sx = malloc(cap * sizeof *sx);
    char *line = NULL; size_t lcap = 0; int first = 1;
    while (getline(&line, &lcap, fp) > 0) {
        if (first) { first = 0; continue; }
        long long a, b, c, d;
        if (sscanf(line, "%lld %lld %lld %lld", &a, &b, &c, &d) != 4) continue;
        if (n == cap) {
            cap *= 2;
            ind = realloc(ind, cap * sizeof *ind); fa = realloc(fa, cap * sizeof *fa);
            mo = realloc(mo, cap * sizeof *mo); sx = realloc(sx, cap * sizeof *sx);
        }
        ind[n] = a; fa[n] = b; mo[n] = c; sx[n] = (int32_t)d; n++;
    }
    free(line); fclose(fp);
    oracle_ped *p = build_ped((int)n, ind, fa, mo, sx, do_sort, status);
    free(ind); free(fa); free(mo); free(sx);
    return p;
}

int oracle_ped_n(const oracle_ped *p) { return p->n; }

/* Flat view in rank order (what the C ABI of the product takes). */
void oracle_ped_arrays(const oracle_ped *p, int64_t *ids, int32_t *father, int32_t *mother, int32_t *sex) {
    for (int r = 0; r < p->n; r++) {
        if (ids) ids[r] = p->id[r];
        if (father) father[r] = p->father[r];
        if (mother) mother[r] = p->mother[r];
        if (sex) sex[r] = p->sex[r];
    }
}

int oracle_rank_of(const oracle_ped *p, int64_t id) { return hash_get(p->hkey, p->hval, p->hmask, id); }

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b; return (x > y) - (x < y);
}

/* identify.jl:35-39 -- IDs without children, sorted ascending.  Returns count. */
int oracle_pro(const oracle_ped *p, int64_t *out) {
    int k = 0;
    for (int r = 0; r < p->n; r++) if (p->nchild[r] == 0) { if (out) out[k] = p->id[r]; k++; }
    if (out) qsort(out, (size_t)k, sizeof *out, cmp_i64);
    return k;
}

/* compute.jl:66-95 -- plain Karigl recursion (exponential; small inputs only). */
static double pair_phi(const oracle_ped *p, int i, int j) {
    double value = 0.;
    if (i > j) {
        if (p->father[i] >= 0) value += pair_phi(p, p->father[i], j) / 2;
        if (p->mother[i] >= 0) value += pair_phi(p, p->mother[i], j) / 2;
    } else if (j > i) {
        if (p->father[j] >= 0) value += pair_phi(p, p->father[j], i) / 2;
        if (p->mother[j] >= 0) value += pair_phi(p, p->mother[j], i) / 2;
    } else {
        value += 0.5;
        if (p->father[i] >= 0 && p->mother[i] >= 0) value += pair_phi(p, p->father[i], p->mother[i]) / 2;
    }
    return value;
}
int oracle_phi_pair(const oracle_ped *p, int64_t id1, int64_t id2, double *out) {
    int a = oracle_rank_of(p, id1), b = oracle_rank_of(p, id2);
    if (a < 0 || b < 0) return ORACLE_EKEY;
    *out = pair_phi(p, a, b);
    return ORACLE_OK;
}

/* ------------------------------------------- compute.jl: the square matrix */

typedef struct { int32_t *v; int n, cap; } ivec;
static void iv_push(ivec *a, int32_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = realloc(a->v, (size_t)a->cap * sizeof *a->v); }
    a->v[a->n++] = x;
}

/* compute.jl:193-207 -- father then mother per member, unique! keeps the first
 * occurrence.  `stamp`/`tick` implement unique!. */
static ivec previous_generation(const int32_t *father, const int32_t *mother, const ivec *next,
                                int32_t *stamp, int32_t tick) {
    ivec prev = {0};
    for (int t = 0; t < next->n; t++) {
        int x = next->v[t];
        int f = father[x], m = mother[x];
        if (f >= 0 && stamp[f] != tick) { stamp[f] = tick; iv_push(&prev, f); }
        if (m >= 0 && stamp[m] != tick) { stamp[m] = tick; iv_push(&prev, m); }
    }
    return prev;
}

typedef struct {
    int S;        /* number of levels (length(cut_vertices)) */
    ivec *cut;    /* cut[k], k = 0..S-1, members are 0-based ranks */
} cuts_t;

static void cuts_free(cuts_t *c) { for (int k = 0; k < c->S; k++) free(c->cut[k].v); free(c->cut); }

/* compute.jl:236-251.  top_down[k] = union(levels[k], top_down[k-1]) and
 * bottom_up[k] = union(levels[k], bottom_up[k+1]) are realised with
 * first/last level marks: x is in top_down[k] iff first[x] <= k, in
 * bottom_up[k] iff last[x] >= k.  The ORDER of cut[k] is that of
 * intersect(top_down[k], bottom_up[k]) = top_down[k] filtered, i.e. levels[k]
 * (deduplicated) followed by the surviving members of top_down[k-1] in their
 * previous order. */
static cuts_t build_cuts(int n, const int32_t *father, const int32_t *mother, const ivec *probands) {
    int32_t *stamp = malloc((size_t)(n + 1) * sizeof *stamp);
    for (int i = 0; i < n; i++) stamp[i] = -1;
    /* raw levels, bottom first, then reversed (pushfirst!) */
    int cap = 16, S = 0; ivec *rev = malloc((size_t)cap * sizeof *rev);
    ivec cur = {0};
    for (int t = 0; t < probands->n; t++) iv_push(&cur, probands->v[t]);
    rev[S++] = cur;
    for (;;) {
        ivec prev = previous_generation(father, mother, &rev[S - 1], stamp, S);
        if (prev.n == 0) { free(prev.v); break; }
        if (S == cap) { cap *= 2; rev = realloc(rev, (size_t)cap * sizeof *rev); }
        rev[S++] = prev;
    }
    ivec *levels = malloc((size_t)S * sizeof *levels);
    for (int k = 0; k < S; k++) levels[k] = rev[S - 1 - k];
    free(rev);
    int32_t *first = malloc((size_t)(n + 1) * sizeof *first), *last = malloc((size_t)(n + 1) * sizeof *last);
    for (int i = 0; i < n; i++) { first[i] = -1; last[i] = -1; }
    for (int k = 0; k < S; k++)
        for (int t = 0; t < levels[k].n; t++) {
            int x = levels[k].v[t];
            if (first[x] < 0) first[x] = k;
            last[x] = k;
        }
    cuts_t c; c.S = S; c.cut = calloc((size_t)S, sizeof *c.cut);
    ivec top = {0};
    for (int i = 0; i < n; i++) stamp[i] = -1;
    for (int k = 0; k < S; k++) {
        /* top_down[k] = union(levels[k], top_down[k-1]) */
        ivec nt = {0};
        for (int t = 0; t < levels[k].n; t++) {
            int x = levels[k].v[t];
            if (stamp[x] != k) { stamp[x] = k; iv_push(&nt, x); }
        }
        for (int t = 0; t < top.n; t++) {
            int x = top.v[t];
            /* members whose last level is behind us can never re-enter a cut:
             * dropping them here keeps later orders unchanged */
            if (stamp[x] != k && last[x] >= k) { stamp[x] = k; iv_push(&nt, x); }
        }
        free(top.v); top = nt;
        for (int t = 0; t < top.n; t++) if (last[top.v[t]] >= k) iv_push(&c.cut[k], top.v[t]);
    }
    free(top.v);
    for (int k = 0; k < S; k++) free(levels[k].v);
    free(levels); free(first); free(last); free(stamp);
    return c;
}

typedef struct {
    const int32_t *father, *mother;
    const int32_t *founder_index;   /* 0 = not indexed, else 1-based position (never reset) */
    const float *Psi;
    int ldpsi;
} phi_ctx;

/* compute.jl:105-158, branch for branch.  `i`, `j` are 0-based ranks, so the
 * reference's `rank` comparisons are index comparisons. */
static double cut_phi(const phi_ctx *c, int i, int j) {
    double value = 0.;
    int fi = c->founder_index[i], fj = c->founder_index[j];
    if (fi != 0 && fj != 0) {
        value += c->Psi[(size_t)(fi - 1) * c->ldpsi + (fj - 1)];
    } else if (fi != 0) {
        if (c->father[j] >= 0) value += cut_phi(c, i, c->father[j]) / 2;
        if (c->mother[j] >= 0) value += cut_phi(c, i, c->mother[j]) / 2;
    } else if (fj != 0) {
        if (c->father[i] >= 0) value += cut_phi(c, j, c->father[i]) / 2;
        if (c->mother[i] >= 0) value += cut_phi(c, j, c->mother[i]) / 2;
    } else {
        if (i > j) {
            if (c->father[i] >= 0) value += cut_phi(c, c->father[i], j) / 2;
            if (c->mother[i] >= 0) value += cut_phi(c, c->mother[i], j) / 2;
        } else if (j > i) {
            if (c->father[j] >= 0) value += cut_phi(c, c->father[j], i) / 2;
            if (c->mother[j] >= 0) value += cut_phi(c, c->mother[j], i) / 2;
        } else {
            value += 0.5;
            if (c->father[i] >= 0 && c->mother[i] >= 0)
                value += cut_phi(c, c->father[i], c->mother[i]) / 2;
        }
    }
    return value;
}

/* compute.jl:293-299 -- Threads.@threads over i (and j): rows are handed out
 * dynamically to `nthreads` POSIX threads; writes are disjoint (i <= j guard). */
typedef struct {
    const phi_ctx *ctx; const int32_t *mem; int nn; float *phi; int next_row; pthread_mutex_t mu;
} pair_job;
static void *pair_worker(void *arg) {
    pair_job *job = arg;
    const int chunk = 8;
    for (;;) {
        pthread_mutex_lock(&job->mu);
        int i0 = job->next_row; job->next_row += chunk;
        pthread_mutex_unlock(&job->mu);
        if (i0 >= job->nn) break;
        int i1 = i0 + chunk < job->nn ? i0 + chunk : job->nn;
        for (int i = i0; i < i1; i++)
            for (int j = i; j < job->nn; j++) {
                float v = (float)cut_phi(job->ctx, job->mem[i], job->mem[j]);
                job->phi[(size_t)i * job->nn + j] = v;
                job->phi[(size_t)j * job->nn + i] = v;
            }
    }
    return NULL;
}
static void pair_loop(const phi_ctx *ctx, const int32_t *mem, int nn, float *phi, int nthreads) {
    pair_job job = { ctx, mem, nn, phi, 0, PTHREAD_MUTEX_INITIALIZER };
    if (nthreads > 256) nthreads = 256;
    if (nthreads <= 1 || nn < 64) { pair_worker(&job); return; }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads - 1; t++) if (pthread_create(&th[started], NULL, pair_worker, &job) == 0) started++;
    pair_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

int oracle_num_threads(void);

/* Bounded CPU baseline: stop after the step that crosses this many seconds (<= 0: off). */
static double g_time_budget = 0.0;
void oracle_set_time_budget(double seconds) { g_time_budget = seconds; }

/*
 * compute.jl:233-304 on flat rank-indexed arrays.
 *   out        n_unique^2 floats (row-major; the matrix is symmetric), may be NULL
 *   n_unique   number of distinct probands (duplicates collapse, compute.jl:251)
 *   max_steps  <0: all steps; otherwise stop after that many steps (bounded CPU
 *              baseline); `out` is then not written
 *   step_info  optional, 6 doubles per step: founders, probands, both (the
 *              verbose line, compute.jl:257-260), pair evaluations, new members,
 *              seconds
 * Returns the number of steps S-1 (>= 0) or a negative error.
 */
int oracle_phi_ranks(int n, const int32_t *father, const int32_t *mother, int n_pro,
                     const int32_t *pro_rank, float *out, int *n_unique, int nthreads,
                     int max_steps, double *step_info, int step_info_cap) {
    ivec pro = {0};
    for (int t = 0; t < n_pro; t++) {
        if (pro_rank[t] < 0 || pro_rank[t] >= n) { free(pro.v); return -ORACLE_EKEY; }
        iv_push(&pro, pro_rank[t]);
    }
    cuts_t c = build_cuts(n, father, mother, &pro);
    free(pro.v);
    int S = c.S;
    int nu = c.cut[S - 1].n;
    if (n_unique) *n_unique = nu;
    int32_t *founder_index = calloc((size_t)n + 1, sizeof *founder_index);
    int32_t *stamp = malloc((size_t)(n + 1) * sizeof *stamp);
    for (int i = 0; i < n; i++) stamp[i] = -1;
    /* compute.jl:271-274 */
    int n0 = c.cut[0].n;
    float *Psi = calloc((size_t)n0 * n0 + 1, sizeof *Psi);
    for (int i = 0; i < n0; i++) Psi[(size_t)i * n0 + i] = 0.5f;
    int nPsi = n0;
    if (nthreads <= 0) nthreads = oracle_num_threads();
    int steps_done = 0;
    const double t_begin = now_s();
    for (int k = 0; k + 1 < S; k++) {                     /* compute.jl:276 */
        if (max_steps >= 0 && k >= max_steps) break;
        if (g_time_budget > 0 && now_s() - t_begin > g_time_budget) break;
        const ivec *prev = &c.cut[k], *next = &c.cut[k + 1];
        double t0 = now_s();
        for (int t = 0; t < prev->n; t++) { founder_index[prev->v[t]] = t + 1; stamp[prev->v[t]] = k; }
        int both = 0, nn = next->n;
        for (int t = 0; t < nn; t++) if (stamp[next->v[t]] == k) both++;
        float *phi = malloc(((size_t)nn * nn + 1) * sizeof *phi);   /* compute.jl:291 */
        phi_ctx ctx = { father, mother, founder_index, Psi, nPsi };
        const int32_t *mem = next->v;
        /* compute.jl:293-299: all pairs i <= j, both triangles stored */
        pair_loop(&ctx, mem, nn, phi, nthreads);
        free(Psi); Psi = phi; nPsi = nn;                  /* compute.jl:301 */
        double t1 = now_s();
        if (step_info && k < step_info_cap) {
            double *s = step_info + 6 * (size_t)k;
            s[0] = prev->n; s[1] = nn; s[2] = both;
            s[3] = (double)nn * (nn + 1) / 2; s[4] = nn - both; s[5] = t1 - t0;
        }
        steps_done++;
    }
    if (out && steps_done == S - 1) memcpy(out, Psi, (size_t)nu * nu * sizeof *out);
    free(Psi); free(founder_index); free(stamp);
    cuts_free(&c);
    return (max_steps >= 0 || g_time_budget > 0) ? steps_done : S - 1;
}

/* Row-updates of a complete gen.phi call = individuals that enter a cut after the first one
 * (compute.jl:243-251), without computing any kinship. */
int64_t oracle_row_updates(int n, const int32_t *father, const int32_t *mother, int n_pro, const int32_t *pro_rank) {
    ivec pro = {0};
    for (int t = 0; t < n_pro; t++) {
        if (pro_rank[t] < 0 || pro_rank[t] >= n) { free(pro.v); return -ORACLE_EKEY; }
        iv_push(&pro, pro_rank[t]);
    }
    cuts_t c = build_cuts(n, father, mother, &pro);
    free(pro.v);
    uint8_t *seen = calloc((size_t)n + 1, 1);
    int64_t rows = 0;
    for (int k = 0; k < c.S; k++)
        for (int t = 0; t < c.cut[k].n; t++) {
            const int x = c.cut[k].v[t];
            if (!seen[x]) { seen[x] = 1; if (k > 0) rows++; }
        }
    free(seen);
    cuts_free(&c);
    return rows;
}

/* gen.phi(ped, probandIDs): IDs -> ranks (KeyError on unknown ID, create.jl:70
 * reached from compute.jl:196), then the core above. */
int oracle_phi(const oracle_ped *p, int n_pro, const int64_t *proband_ids, float *out,
               int *n_unique, int nthreads, int max_steps, double *step_info, int step_info_cap) {
    int32_t *r = malloc((size_t)(n_pro + 1) * sizeof *r);
    for (int t = 0; t < n_pro; t++) {
        r[t] = oracle_rank_of(p, proband_ids[t]);
        if (r[t] < 0) { free(r); return -ORACLE_EKEY; }
    }
    int rc = oracle_phi_ranks(p->n, p->father, p->mother, n_pro, r, out, n_unique, nthreads,
                              max_steps, step_info, step_info_cap);
    free(r);
    return rc;
}

/* compute.jl:454-459 -- Float32 sums in Julia's order are pairwise; we return
 * the Float64 value of (sum - trace) / (n^2 - n) and let the caller compare
 * with a tolerance-free check only where the sums are exact (geneaJi). */
double oracle_phi_mean(const float *phi, int n) {
    double total = 0., diag = 0.;
    for (size_t i = 0; i < (size_t)n * n; i++) total += phi[i];
    for (int i = 0; i < n; i++) diag += phi[(size_t)i * n + i];
    return (total - diag) / ((double)n * n - n);
}

/* Threads the baseline uses by default: the cores this process may run on. */
int oracle_num_threads(void) {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) { int c = CPU_COUNT(&set); if (c > 0) return c; }
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

/* ---------------------------------------------------------------------------------------------
 * gen.sparse_phi(pedigree, probandIDs) -- src/compute.jl:321-447, transliterated on rank arrays.
 *
 * The reference keeps phi = Dict{rank, Dict{rank, Float32}}.  Two properties of that code decide
 * the values and are kept literally here (`directed` != 0):
 *   - the queue is seeded with founder(isolated_pedigree), i.e. the founders sorted by ID
 *     (src/identify.jl:15-19 via :335-339) -- `ids` gives the ID of every rank;
 *   - a kinship is STORED under phi[rank_j][rank_i] with j the individual processed earlier (:393),
 *     but every LOOK-UP (:350-358, :367-389) and getindex (:36-40) reads phi[lower rank][higher
 *     rank].  When the queue order of two individuals of the same depth inverts their rank order the
 *     stored value is never found again: the reference silently treats that kinship as absent (0).
 *     Such entries still count for the `show` line (:42-46) and the sums of phiMean (:466-472), and
 *     the clean-up (:407-411, :421-425) only deletes keys filed under a LOWER rank, so the inner
 *     dictionaries of the probands can keep keys of individuals that are long gone ("orphans").
 * `directed` == 0 stores symmetrically instead: the mathematically consistent variant of the same
 * schedule (engine: GENLIB_SCHEDULE_SPARSE_PHI_SYMMETRIC); no kinship is lost.
 *
 * phi[a][b] lives in val/has[slot a][slot b] over the LIVE individuals (a slot is cleared when it is
 * handed out again); keys whose individual has been dropped are only counted (orphans).  Other lines:
 *   :323      branching(pedigree, pro = probandIDs): only ancestors of the probands (ranks keep
 *             their order, src/extract.jl:136-159)
 *   :176-183  children lists in rank order (_index_pedigree)
 *   :349-361  self kinship 0.5 + phi[father, mother] / 2
 *   :363-395  kinship with every individual still to visit: `x / 2` is a FLOAT32 division, the
 *             accumulator is Float64, the store rounds to Float32; zeros are not stored
 *   :397-430  children_to_process, eviction of non-proband parents
 *   :431-439  a child enters the queue when both known parents are processed
 * out: n_unique x n_unique by getindex (:36-40), probands in first-occurrence order.
 * counts (nullable, 4 entries): [0] stored entries of the `show` line, [1] of those, entries a
 * look-up can find (diagonal + lower->higher keys between probands), [2] misfiled entries between
 * live individuals, [3] orphan keys.  sums (nullable, 2 doubles): sum of all stored values and of
 * the diagonal, Float64 accumulation (phiMean :466-472 sums the same multiset in Dict order).
 * Returns 0 or a negative ORACLE_* code.
 * --------------------------------------------------------------------------------------------- */
typedef struct { int64_t id; int32_t rank; } founder_key;
static int cmp_founder(const void *a, const void *b) {
    const founder_key *x = a, *y = b;
    if (x->id != y->id) return x->id < y->id ? -1 : 1;
    return (x->rank > y->rank) - (x->rank < y->rank);
}

int oracle_sparse_phi_ranks(int n, const int32_t *father, const int32_t *mother, const int64_t *ids,
                            int n_pro, const int32_t *pro_rank, int directed, float *out, int *n_unique,
                            int64_t *counts, double *sums) {
    uint8_t *is_pro = calloc((size_t)n + 1, 1), *keep = calloc((size_t)n + 1, 1);
    int32_t *uniq = malloc((size_t)(n_pro + 1) * sizeof *uniq);
    int nu = 0;
    for (int t = 0; t < n_pro; t++) {
        if (pro_rank[t] < 0 || pro_rank[t] >= n) { free(is_pro); free(keep); free(uniq); return -ORACLE_EKEY; }
        if (!is_pro[pro_rank[t]]) { is_pro[pro_rank[t]] = 1; uniq[nu++] = pro_rank[t]; }
    }
    if (n_unique) *n_unique = nu;
    /* :323 branching: ancestors of the probands (parents have lower ranks) */
    for (int t = 0; t < nu; t++) keep[uniq[t]] = 1;
    for (int x = n - 1; x >= 0; x--) if (keep[x]) {
        if (father[x] >= 0) keep[father[x]] = 1;
        if (mother[x] >= 0) keep[mother[x]] = 1;
    }
    /* :176-183 children in rank order */
    int32_t *cstart = calloc((size_t)n + 2, sizeof *cstart);
    for (int x = 0; x < n; x++) if (keep[x]) {
        if (father[x] >= 0) cstart[father[x] + 1]++;
        if (mother[x] >= 0) cstart[mother[x] + 1]++;
    }
    for (int x = 0; x < n; x++) cstart[x + 1] += cstart[x];
    int32_t *clist = malloc(((size_t)cstart[n] + 1) * sizeof *clist), *cfill = malloc(((size_t)n + 1) * sizeof *cfill);
    memcpy(cfill, cstart, (size_t)n * sizeof *cfill);
    for (int x = 0; x < n; x++) if (keep[x]) {
        if (father[x] >= 0) clist[cfill[father[x]]++] = x;
        if (mother[x] >= 0) clist[cfill[mother[x]]++] = x;
    }
    /* the queue order does not depend on the values: run it once to size the live set */
    int32_t *queue = malloc(((size_t)n + 1) * sizeof *queue), *todo = malloc(((size_t)n + 1) * sizeof *todo);
    uint8_t *done = calloc((size_t)n + 1, 1);
    int qn = 0;
    {   /* :335-339 founder(isolated_pedigree) = sort(founder IDs), identify.jl:15-19 */
        founder_key *fk = malloc(((size_t)n + 1) * sizeof *fk);
        int nf = 0;
        for (int x = 0; x < n; x++) if (keep[x] && father[x] < 0 && mother[x] < 0) { fk[nf].id = ids ? ids[x] : x; fk[nf].rank = x; nf++; }
        qsort(fk, (size_t)nf, sizeof *fk, cmp_founder);
        for (int k = 0; k < nf; k++) queue[qn++] = fk[k].rank;
        free(fk);
    }
    int live = 0, max_live = 0;
    for (int h = 0; h < qn; h++) {
        int i = queue[h];
        live++; if (live > max_live) max_live = live;
        done[i] = 1;
        todo[i] = cstart[i + 1] - cstart[i];
        int par[2] = { father[i], mother[i] };
        for (int s = 0; s < 2; s++) if (par[s] >= 0 && !is_pro[par[s]] && --todo[par[s]] == 0) live--;
        for (int k = cstart[i]; k < cstart[i + 1]; k++) {       /* :431-439 */
            int c = clist[k];
            if (father[c] >= 0 && mother[c] >= 0) { if (done[father[c]] && done[mother[c]]) queue[qn++] = c; }
            else queue[qn++] = c;
        }
    }
    const size_t W = (size_t)max_live + 1;
    float *val = calloc(W * W, sizeof *val);                     /* phi[a][b] = val[slot a][slot b] if has[..] */
    uint8_t *has = calloc(W * W, 1);
    int32_t *slot = malloc(((size_t)n + 1) * sizeof *slot), *free_slots = malloc(W * sizeof *free_slots);
    int32_t *visit = malloc(W * sizeof *visit), *vpos = malloc(((size_t)n + 1) * sizeof *vpos);   /* ranks_to_visit */
    int64_t *orphans = calloc((size_t)n + 1, sizeof *orphans);   /* keys of dropped individuals left in phi[x] */
    double *orphan_sum = calloc((size_t)n + 1, sizeof *orphan_sum);
    if (!val || !has || !slot || !free_slots || !visit || !vpos || !orphans || !orphan_sum) return -ORACLE_ENOMEM;
    int nfree = 0, nvisit = 0;
    for (size_t k = 0; k < W; k++) free_slots[nfree++] = (int32_t)(W - 1 - k);
    memset(done, 0, (size_t)n + 1);
#define HAS(a, b) has[(size_t)slot[a] * W + (size_t)slot[b]]
#define VAL(a, b) val[(size_t)slot[a] * W + (size_t)slot[b]]
    /* the key the reference looks up for the pair {a, b}: phi[lower rank][higher rank] */
#define LOOK(a, b, acc) do { const int lo_ = (a) < (b) ? (a) : (b), hi_ = (a) < (b) ? (b) : (a);       \
        if (HAS(lo_, hi_)) (acc) += (double)(float)(VAL(lo_, hi_) / 2.0f); } while (0)
    for (int h = 0; h < qn; h++) {
        const int i = queue[h], f = father[i], m = mother[i];
        const int si = free_slots[--nfree];
        slot[i] = si;
        for (size_t k = 0; k < W; k++) { has[(size_t)si * W + k] = 0; has[k * W + (size_t)si] = 0; }   /* phi[rank_i] = Dict() */
        orphans[i] = 0; orphan_sum[i] = 0.;
        double coefficient = 0.5;                                /* :349-361 */
        if (f >= 0 && m >= 0) {
            if (directed) { if (f < m) { if (HAS(f, m)) coefficient += (double)(float)(VAL(f, m) / 2.0f); }
                            else { if (HAS(m, f)) coefficient += (double)(float)(VAL(m, f) / 2.0f); } }
            else LOOK(f, m, coefficient);
        }
        VAL(i, i) = (float)coefficient; HAS(i, i) = 1;
        for (int v = 0; v < nvisit; v++) {                       /* :363-395 */
            const int j = visit[v];
            coefficient = 0.;
            if (f >= 0) {
                if (j < f) { if (HAS(j, f)) coefficient += (double)(float)(VAL(j, f) / 2.0f); }
                else { if (HAS(f, j)) coefficient += (double)(float)(VAL(f, j) / 2.0f); }
            }
            if (m >= 0) {
                if (j < m) { if (HAS(j, m)) coefficient += (double)(float)(VAL(j, m) / 2.0f); }
                else { if (HAS(m, j)) coefficient += (double)(float)(VAL(m, j) / 2.0f); }
            }
            if (coefficient > 0.) {
                if (directed) { VAL(j, i) = (float)coefficient; HAS(j, i) = 1; }          /* :393, whatever the ranks */
                else if (j < i) { VAL(j, i) = (float)coefficient; HAS(j, i) = 1; }
                else { VAL(i, j) = (float)coefficient; HAS(i, j) = 1; }
            }
        }
        vpos[i] = nvisit; visit[nvisit++] = i;                   /* :397-399 */
        done[i] = 1;
        todo[i] = cstart[i + 1] - cstart[i];
        int par[2] = { f, m };
        for (int s = 0; s < 2; s++) {                            /* :400-430 */
            const int p = par[s];
            if (p < 0 || is_pro[p]) continue;
            if (--todo[p] == 0) {
                const int last = visit[--nvisit];                /* delete!(ranks_to_visit, p) */
                visit[vpos[p]] = last; vpos[last] = vpos[p];
                for (int v = 0; v < nvisit; v++) {               /* only keys under a LOWER rank are deleted (:407-411) */
                    const int j = visit[v];
                    if (j > p && HAS(j, p)) { orphans[j]++; orphan_sum[j] += (double)VAL(j, p); }
                }
                free_slots[nfree++] = slot[p];                   /* empty!(phi[p]); delete!(phi, p) */
                slot[p] = -1;
            }
        }
    }
    if (out) {
        for (int a = 0; a < nu; a++)
            for (int b = 0; b < nu; b++) {                       /* getindex, :36-40 */
                const int lo = uniq[a] < uniq[b] ? uniq[a] : uniq[b], hi = uniq[a] < uniq[b] ? uniq[b] : uniq[a];
                out[(size_t)a * nu + b] = HAS(lo, hi) ? VAL(lo, hi) : 0.f;
            }
    }
    if (counts || sums) {
        int64_t nz = 0, findable = 0, misfiled = 0, norph = 0;
        double total = 0., diag = 0.;
        for (int v = 0; v < nvisit; v++) {                       /* what is left in phi: the probands */
            const int a = visit[v];
            norph += orphans[a]; total += orphan_sum[a];
            for (int w = 0; w < nvisit; w++) {
                const int b = visit[w];
                if (!HAS(a, b)) continue;
                nz++; total += (double)VAL(a, b);
                if (a == b) diag += (double)VAL(a, b);
                if (a <= b) findable++; else misfiled++;
            }
        }
        if (counts) { counts[0] = nz + norph; counts[1] = findable; counts[2] = misfiled; counts[3] = norph; }
        if (sums) { sums[0] = total; sums[1] = diag; }
    }
#undef HAS
#undef VAL
#undef LOOK
    free(is_pro); free(keep); free(uniq); free(cstart); free(clist); free(cfill); free(queue); free(todo);
    free(done); free(val); free(has); free(slot); free(free_slots); free(visit); free(vpos); free(orphans); free(orphan_sum);
    return ORACLE_OK;
}
