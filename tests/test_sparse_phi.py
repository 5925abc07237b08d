"""`gen.sparse_phi`: oracle restatement of src/compute.jl:321-447 pinned to the
reference's known answers, the planner's sparse_phi schedule replayed on the CPU, the KinshipMatrix
mirror.  The GPU half is in test_gpu_parity.py."""
import numpy as np
import pytest

from plan_replay import replay
from util import ladder_pedigree, random_pedigree


def test_oracle_sparse_phi_geneaji_golden(gen, ob):
    ped = gen.genealogy(gen.geneaJi)
    ranks = ped.rank_of(gen.pro(ped))
    dense, stored = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids)
    _, info = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, full=True)
    assert info["misfiled"] == 0 and info["orphans"] == 0 and info["findable"] == 6
    k = gen.KinshipMatrix(gen.pro(ped), ranks, dense)
    assert gen.phiMean(k) == np.float32(0.171875)                                # test/runtests.jl:55
    assert repr(k) == "3×3 KinshipMatrix with 6 stored entries." and stored == 6   # :56
    assert k[1, 2] == np.float32(0.37109375) and k[2, 1] == k[1, 2]              # :57
    # keyed lower rank -> higher rank (compute.jl:36-40); 29 is a founder, so it ranks first
    assert k.to_dict() == {29: {29: np.float32(0.53515625), 1: np.float32(0.072265625), 2: np.float32(0.072265625)},
                           1: {1: np.float32(0.591796875), 2: np.float32(0.37109375)}, 2: {2: np.float32(0.591796875)}}
    with pytest.raises(KeyError):
        k[1, 3]


def test_oracle_sparse_phi_equals_phi_where_nothing_rounds(gen, ob):
    """Shallow pedigrees: every value is a short dyadic fraction, so both reference functions
    (and the exact recursion behind test_oracle.py) give the same matrix."""
    s = gen.synth.generate(900, 6, 25, alpha=0.1, demes=1, overlap=2, seed=4)      # 6 generations: <= 13 halvings
    ped = gen.genealogy(s.as_columns())
    ranks = ped.rank_of(s.probands)
    dense, _ = ob.phi_ranks(ped.father, ped.mother, ranks)
    sparse, stored = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, directed=False)
    assert np.array_equal(dense, sparse)
    assert stored == 25 + np.count_nonzero(np.triu(sparse, 1))
    # the reference itself (misfiled kinships are lost, compute.jl:393): never above the true values
    lossy, info = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, full=True)
    assert info["misfiled"] + info["orphans"] > 0 and not np.array_equal(lossy, dense) and (lossy <= dense).all()


@pytest.mark.parametrize("seed", range(10))
def test_replayed_sparse_schedule_equals_oracle_random(gen, ob, seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(60, 600))
    rec = random_pedigree(rng, n, int(rng.integers(3, 12)), window=int(rng.choice([0, 0, 30, 80])))
    ped = gen.genealogy(rec)
    pro = rng.permutation(ped.ids)[: int(rng.integers(2, 40))]
    pro = np.concatenate([pro, pro[:2]])                                         # duplicates collapse
    ranks = ped.rank_of(pro)
    want, _ = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids)
    plan = gen.Plan(ped.father, ped.mother, ranks, schedule="sparse_phi", ids=ped.ids)
    assert np.array_equal(replay(plan), want)


def test_sparse_schedule_differs_from_phi_and_keeps_subnormals(gen, ob):
    """Deep pedigrees: the two reference functions round at different points (per individual vs
    per step), and sparse_phi halves in Float32, which matters once values are subnormal."""
    s = gen.synth.generate(16 * 60, 60, 16, alpha=0.2, overlap=1, seed=11)
    ped = gen.genealogy(s.as_columns())
    ranks = ped.rank_of(s.probands)
    want, _ = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids)
    dense, _ = ob.phi_ranks(ped.father, ped.mother, ranks)
    assert not np.array_equal(want, dense)
    assert np.array_equal(replay(gen.Plan(ped.father, ped.mother, ranks, schedule="sparse_phi", ids=ped.ids)), want)
    assert not np.array_equal(replay(gen.Plan(ped.father, ped.mother, ranks)), want)      # phi's schedule is another one
    # lineages 73 generations apart: single-bit subnormals, halved in Float32 (compute.jl:350-389)
    cols, pro = ladder_pedigree(73)
    ped = gen.genealogy(cols)
    ranks = ped.rank_of(pro)
    want, _ = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids)
    dense, _ = ob.phi_ranks(ped.father, ped.mother, ranks)
    tiny = np.float32(np.finfo(np.float32).tiny)
    assert ((want > 0) & (want < tiny)).any()                                    # gradual underflow reached
    assert not np.array_equal(want, dense)                                       # 0 here, one ulp of a subnormal there
    assert np.array_equal(replay(gen.Plan(ped.father, ped.mother, ranks, schedule="sparse_phi", ids=ped.ids)), want)
    assert np.array_equal(replay(gen.Plan(ped.father, ped.mother, ranks)), dense)


def test_sparse_schedule_layers_are_depths(gen):
    s = gen.synth.generate(3000, 9, 150, alpha=0.05, demes=2, migration=0.1, overlap=3, seed=8)
    ped = gen.genealogy(s.as_columns())
    ranks = ped.rank_of(s.probands)
    plan = gen.Plan(ped.father, ped.mother, ranks, schedule="sparse_phi", ids=ped.ids)
    dense_plan = gen.Plan(ped.father, ped.mother, ranks)
    assert plan.row_updates == dense_plan.row_updates                            # same ancestors (branching, compute.jl:323)
    last = -1
    seen = 0
    for t in range(plan.n_layers):
        seq = np.sort(plan.layer_arrays(t)["member_ind"])
        if len(seq):
            assert seq[0] == last + 1 and np.array_equal(seq, np.arange(seq[0], seq[0] + len(seq)))   # queue order, by depth
            last = int(seq[-1]); seen += len(seq)
    assert seen == plan.row_updates
    first = plan.layer_arrays(0)
    assert np.all(first["fam_father_slot"] == -1) and np.all(first["fam_mother_slot"] == -1)   # layer 0 = founders
    with pytest.raises(Exception):
        gen.Plan(ped.father, ped.mother, ranks, schedule="nope")


@pytest.mark.parametrize("schedule", ["phi", "sparse_phi"])
def test_edge_cases_both_schedules(gen, ob, schedule):
    """Empty / founder-only / duplicate proband lists, a single individual, a 3000-generation chain."""
    want_of = (lambda f, m, p: ob.sparse_phi_ranks(f, m, p)[0]) if schedule == "sparse_phi" \
        else (lambda f, m, p: ob.phi_ranks(f, m, p)[0])
    f, m = np.array([-1, -1, 0, 0, 2], np.int32), np.array([-1, -1, 1, 1, 3], np.int32)
    for pro in ([], [0], [0, 1], [4], [4, 4, 2], [0, 1, 2, 3, 4]):
        p = np.array(pro, np.int32)
        plan = gen.Plan(f, m, p, schedule=schedule)
        assert plan.n_unique == len(set(pro))
        if plan.n_unique:
            assert np.array_equal(replay(plan), want_of(f, m, p))
    one = np.array([-1], np.int32)
    assert np.array_equal(replay(gen.Plan(one, one, np.array([0], np.int32), schedule=schedule)), np.array([[0.5]], np.float32))
    n = 3000
    f, m, p = np.arange(-1, n - 1, dtype=np.int32), np.full(n, -1, np.int32), np.array([n - 1, n - 2, 5], np.int32)
    plan = gen.Plan(f, m, p, schedule=schedule)
    assert plan.n_layers == n and np.array_equal(replay(plan), want_of(f, m, p))


def six():
    """Founders 1-4, A = (3, 4) and B = (1, 3), A before B in the file: ranks 5 and 6.  The queue hands
    out B first (founder 3 completes it before founder 4 completes A), so the reference files
    phi[rank B][rank A] and never finds it again: phi[A, B] = 0, although the true kinship is 1/8."""
    return {"ind": np.array([1, 2, 3, 4, 5, 6]), "father": np.array([0, 0, 0, 0, 3, 1]),
            "mother": np.array([0, 0, 0, 0, 4, 3]), "sex": np.array([1, 2, 1, 2, 1, 2], np.int32)}


def test_misfiled_kinship_is_lost_like_in_the_reference(gen, ob):
    ped = gen.genealogy(six())
    ranks = ped.rank_of([5, 6])
    dense, info = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, full=True)
    assert np.array_equal(dense, np.array([[0.5, 0.0], [0.0, 0.5]], np.float32))
    assert info == {"stored": 3, "findable": 2, "misfiled": 1, "orphans": 0, "sum": 1.125, "diag": 1.0}
    sym, info = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, directed=False, full=True)
    assert np.array_equal(sym, np.array([[0.5, 0.125], [0.125, 0.5]], np.float32)) and info["misfiled"] == 0
    assert np.array_equal(sym, ob.phi_ranks(ped.father, ped.mother, ranks)[0])
    for schedule, want in (("sparse_phi", dense), ("sparse_phi_symmetric", sym)):
        assert np.array_equal(replay(gen.Plan(ped.father, ped.mother, ranks, schedule=schedule, ids=ped.ids)), want)
    # a child of A and B: its self-kinship misses the lost 1/8 in the reference
    cols = six()
    cols = {k: np.append(v, x) for (k, v), x in zip(cols.items(), (7, 5, 6, 1))}
    ped = gen.genealogy(cols)
    ranks = ped.rank_of([7])
    assert ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids)[0][0, 0] == np.float32(0.5)
    assert ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, directed=False)[0][0, 0] == np.float32(0.5625)
    assert replay(gen.Plan(ped.father, ped.mother, ranks, schedule="sparse_phi", ids=ped.ids))[0, 0] == np.float32(0.5)
    assert replay(gen.Plan(ped.father, ped.mother, ranks, schedule="sparse_phi_symmetric", ids=ped.ids))[0, 0] == np.float32(0.5625)


@pytest.mark.parametrize("seed,n,window,rounding_only", [(303, 400, 40, True), (305, 2000, 40, True), (302, 1000, 120, False)])
def test_founders_enter_the_queue_in_id_order(gen, ob, seed, n, window, rounding_only):
    """founder() sorts by ID (identify.jl:15-19): with permuted IDs the queue differs from the rank
    order.  That decides which kinships the reference misfiles, and on deep, narrow pedigrees it
    also moves the last bits of the consistent variant (who is climbed first)."""
    rng = np.random.default_rng(seed)
    rec = random_pedigree(rng, n, 8 if n < 2000 else 10, window=window)
    ped = gen.genealogy(rec)
    founders = np.nonzero((ped.father < 0) & (ped.mother < 0))[0]
    assert not np.array_equal(np.argsort(ped.ids[founders], kind="stable"), np.arange(len(founders)))
    ranks = ped.rank_of(rng.permutation(ped.ids)[:100])
    for schedule, directed in (("sparse_phi", True), ("sparse_phi_symmetric", False)):
        by_id = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, directed=directed)[0]
        by_rank = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=None, directed=directed)[0]
        if directed or rounding_only:
            assert not np.array_equal(by_id, by_rank)
        assert np.array_equal(replay(gen.Plan(ped.father, ped.mother, ranks, schedule=schedule, ids=ped.ids)), by_id)
        assert np.array_equal(replay(gen.Plan(ped.father, ped.mother, ranks, schedule=schedule)), by_rank)


def test_kinship_matrix_counts(gen, ob):
    """`stored` of the mirror = entries a look-up finds; the oracle also counts what the reference's
    `show` line counts (misfiled and orphaned keys included)."""
    rng = np.random.default_rng(5)
    ped = gen.genealogy(random_pedigree(rng, 300, 10))
    pro = rng.permutation(ped.ids)[:40]
    ranks = ped.rank_of(pro)
    dense, info = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, full=True)
    k = gen.KinshipMatrix(pro, ranks, dense)
    assert k.stored == info["findable"] and info["stored"] == info["findable"] + info["misfiled"] + info["orphans"]
    assert info["misfiled"] + info["orphans"] > 0


@pytest.mark.parametrize("seed,window", [(5, 0), (302, 40), (303, 120)])
def test_kinship_matrix_keeps_only_what_the_reference_stores(gen, ob, seed, window):
    """The KinshipMatrix holds compressed sparse rows -- per individual the diagonal and the non-zero kinships with
    higher-ranked individuals, what the reference's Dict of Dicts holds where a look-up finds it (compute.jl:391-394)
    -- and still answers every look-up, its dense form and its mean like the dense matrix it was made from."""
    rng = np.random.default_rng(seed)
    ped = gen.genealogy(random_pedigree(rng, 600, 12, window=window))
    pro = rng.permutation(ped.ids)[:80]
    ranks = ped.rank_of(pro)
    order = np.argsort(ranks, kind="stable")
    for directed in (True, False):
        dense = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, directed=directed)[0]
        assert np.array_equal(dense, dense.T)
        k = gen.KinshipMatrix(pro, ranks, dense)
        want = np.ascontiguousarray(dense[np.ix_(order, order)])
        assert np.array_equal(k.to_dense().view(np.uint32), want.view(np.uint32))
        assert k.stored == int(np.count_nonzero(np.triu(want, 1))) + len(pro) == len(k._data)
        assert k.stored < len(pro) * (len(pro) + 1) // 2                   # zeros are not kept
        for a in range(0, 80, 7):
            for b in range(0, 80, 5):
                assert k[int(pro[a]), int(pro[b])] == dense[a, b] == k[int(pro[b]), int(pro[a])]
        d = k.to_dict()
        assert sum(len(v) for v in d.values()) == k.stored
        assert all(d[int(i)][int(i)] == want[r, r] for r, i in enumerate(np.asarray(pro)[order]))
        # the mean over the stored entries == the mean over the upper triangle (exact in binary64 here)
        up = np.triu(want.astype(np.float64), 1)
        assert abs(float(gen.phiMean(k)) - up.sum() / (80 * 79 / 2)) < 1e-6
    with pytest.raises(KeyError):
        k[int(pro[0]), -12345]
    empty = gen.KinshipMatrix(np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros((0, 0), np.float32))
    assert len(empty) == 0 and empty.stored == 0 and empty.to_dense().shape == (0, 0)
