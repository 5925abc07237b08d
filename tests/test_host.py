"""Host-side mirror of the reference interface: genealogy / pro / phi front end, planner."""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest

from plan_replay import replay
from util import random_pedigree


def test_pedigree_matches_reference_show(gen):
    ped = gen.genealogy(gen.genea140)
    assert repr(ped) == ("A pedigree with:\n41523 individuals;\n68248 parent-child relations;\n20773 men;"
                         "\n20750 women;\n140 subjects;\n18 generations.")        # runtests.jl:18-20
    assert repr(ped[33724]) == "ind: 33724\nfather: 10086\nmother: 10087\nsex: 1"  # runtests.jl:24-25
    assert [c.ID for c in ped[33724].children] == [10033, 113470]                 # runtests.jl:26
    assert ped[33724].children[1].father.ID == 33724                              # runtests.jl:27
    assert len(ped) == 41523 and ped.depth() == 18


def test_custom_pedigree(gen):                                                    # runtests.jl:5-13
    df = {"ind": [1, 2, 3, 4, 5, 6, 7, 8, 9, 10], "father": [0, 0, 0, 1, 1, 0, 3, 3, 6, 6],
          "mother": [0, 0, 0, 2, 2, 0, 4, 4, 5, 5], "sex": [1, 2, 1, 2, 2, 1, 2, 1, 1, 2]}
    ped = gen.genealogy(df)
    assert ped[9].mother.sex == 2


def test_geneaji_selectors(gen):
    ped = gen.genealogy(gen.geneaJi)
    assert gen.pro(ped).tolist() == [1, 2, 29]                                    # runtests.jl:41
    assert gen.founder(ped).tolist() == [17, 19, 20, 23, 25, 26]                  # runtests.jl:42
    with pytest.raises(KeyError):
        ped[1000]
    assert 17 in ped and 1000 not in ped


def test_rank_order_equals_oracle(gen, ob):
    for path in (gen.geneaJi, gen.genea140):
        ped, o = gen.genealogy(path), ob.OraclePedigree.from_csv(path)
        assert np.array_equal(ped.ids, o.ids) and np.array_equal(ped.father, o.father)
        assert np.array_equal(ped.mother, o.mother) and np.array_equal(gen.pro(ped), o.pro())
    rng = np.random.default_rng(0)
    rec = random_pedigree(rng, 300, 20)
    ped = gen.genealogy(rec)
    o = ob.OraclePedigree.from_arrays(rec["ind"], rec["father"], rec["mother"], rec["sex"])
    assert np.array_equal(ped.ids, o.ids) and np.array_equal(ped.father, o.father)


def test_unsorted_file_needs_sort(gen):
    rec = {"ind": [1, 2, 3], "father": [2, 0, 0], "mother": [3, 0, 0], "sex": [1, 1, 2]}
    assert gen.genealogy(rec).ids.tolist() == [2, 3, 1]
    with pytest.raises(KeyError):
        gen.genealogy(rec, sort=False)                                            # create.jl:240


def test_verbose_lines_match_reference_cuts(gen, ob, genea140_oracle):
    """`Step i of S: a founders, b probands, c both.` (compute.jl:257-260) from the planner
    equal the oracle's cut sizes (which restate compute.jl:236-251)."""
    for path in (gen.geneaJi, gen.genea140):
        ped = gen.genealogy(path)
        plan = gen.Plan(ped.father, ped.mother, ped.rank_of(gen.pro(ped)))
        steps = genea140_oracle[2] if path == gen.genea140 else \
            ob.OraclePedigree.from_csv(path).phi(with_steps=True)[1]
        lines = list(plan.verbose_lines())
        assert len(lines) == len(steps)
        for k, (line, s) in enumerate(zip(lines, steps), 1):
            assert line == (f"Step {k} of {len(steps)}: {int(s[0])} founders, {int(s[1])} probands, "
                            f"{int(s[2])} both.")
    buf = io.StringIO()
    ped = gen.genealogy(gen.geneaJi)
    with redirect_stdout(buf):
        assert gen.phi(ped, compute=False) is None                               # compute.jl:264-266
    assert buf.getvalue().splitlines()[0] == "Step 1 of 7: 2 founders, 4 probands, 2 both."


def test_plan_errors(gen):
    with pytest.raises(KeyError):
        gen.Plan([-1, -1, 0], [-1, -1, 1], [5])
    with pytest.raises(gen.GenlibError):
        gen.Plan([1, -1], [-1, -1], [0])                    # parent after child
    ped = gen.genealogy(gen.geneaJi)
    with pytest.raises(KeyError):
        gen.phi(ped, [1, 12345], compute=False)
    assert gen.phi(ped, [], compute=True).shape == (0, 0)   # empty list needs no device


def test_plan_reports_the_first_broken_rank(gen):
    """The pedigree is checked inside the planner's reverse sweep (children before parents); what is reported is still
    the FIRST offending rank, a broken pedigree goes before a bad proband, an empty proband list does not hide it,
    and the streamed planner says the same."""
    n = 50
    fa, mo = np.full(n, -1, np.int32), np.full(n, -1, np.int32)
    fa[10:] = np.arange(0, n - 10); mo[10:] = np.arange(1, n - 9)
    ok = gen.Plan(fa, mo, [n - 1])
    assert ok.n_layers > 1
    for stream in (False, True):
        bad = fa.copy(); bad[40] = 45; bad[20] = 33                       # two parents after their children
        with pytest.raises(gen.GenlibError, match="parent does not precede child at rank 20"):
            gen.Plan(bad, mo, [n - 1], stream=stream)
        bad = fa.copy(); bad[30] = 30                                      # its own parent
        with pytest.raises(gen.GenlibError, match="at rank 30"):
            gen.Plan(bad, mo, [n - 1], stream=stream)
        bad = mo.copy(); bad[44] = n + 7; bad[12] = -2
        with pytest.raises(KeyError, match="out of range at rank 12"):
            gen.Plan(fa, bad, [n - 1], stream=stream)
        with pytest.raises(KeyError, match="out of range at rank 12"):    # ... before the proband that is not there
            gen.Plan(fa, bad, [n + 3], stream=stream)
        with pytest.raises(KeyError, match="out of range at rank 12"):
            gen.Plan(fa, bad, [], stream=stream)
        with pytest.raises(KeyError, match="proband index"):
            gen.Plan(fa, mo, [0, n + 3], stream=stream)
        # a broken row nobody descends from is found as well
        bad = fa.copy(); bad[n - 2] = n - 1
        with pytest.raises(gen.GenlibError, match=f"at rank {n - 2}"):
            gen.Plan(bad, mo, [n - 1], stream=stream)


def test_plan_invariants(gen):
    s = gen.synth.generate(4000, 10, 200, alpha=0.05, demes=2, migration=0.1, overlap=3, seed=5)
    ped = gen.genealogy(s.as_columns())
    plan = gen.Plan(ped.father, ped.mother, ped.rank_of(s.probands))
    seen = set()
    live = {}
    for t in range(plan.n_layers):
        info, arr = plan.layer_info(t), plan.layer_arrays(t)
        flags = arr["live_flags"]
        assert set(np.nonzero(flags & 1)[0]) == set(live.values())
        assert info["n_fam"] <= info["n_new"] and np.all(np.diff(arr["member_fam"]) >= 0)
        for ind, slot in zip(arr["member_ind"], arr["member_slot"]):
            assert ind not in seen and slot not in live.values()
            seen.add(int(ind))
        carried = set(np.nonzero(flags & 2)[0])
        live = {i: sl for i, sl in live.items() if sl in carried}
        live.update({int(i): int(sl) for i, sl in zip(arr["member_ind"], arr["member_slot"])})
        # parents of the layer are live before it
        for f in np.concatenate([arr["fam_father_slot"], arr["fam_mother_slot"]]):
            assert f == -1 or flags[f] & 1
    assert plan.row_updates == len(seen)
    assert set(plan.proband_slots()) <= set(live.values())


def _plan_dump(gen, ped, probands, world):
    plan = gen.Plan(ped.father, ped.mother, ped.rank_of(probands), world=world)
    out = [plan.capacity, plan.n_layers, plan.row_updates, plan.proband_slots().tolist()]
    for t in range(plan.n_layers):
        out.append(sorted((k, np.asarray(v).tolist()) for k, v in plan.layer_arrays(t).items()))
        out.append(sorted((k, np.asarray(v).tolist()) for k, v in plan.layer_shard(t).items()))
        out.append(sorted(plan.layer_info(t).items()))
    return out


@pytest.mark.parametrize("world", [1, 2])
def test_plan_does_not_depend_on_helper_threads(gen, world, monkeypatch):
    """Couples are grouped ahead by helper threads (GENLIB_PLAN_THREADS) on large pedigrees."""
    s = gen.synth.generate(400000, 10, 8000, alpha=0.02, demes=4, migration=0.05, overlap=2, seed=21)
    ped = gen.genealogy(s.as_columns())
    ranks = ped.rank_of(s.probands)
    plans = []
    for threads in ("1", "3"):
        monkeypatch.setenv("GENLIB_PLAN_THREADS", threads)
        plans.append(gen.Plan(ped.father, ped.mother, ranks, world=world))
    a, b = plans
    assert (a.capacity, a.n_layers, a.row_updates) == (b.capacity, b.n_layers, b.row_updates)
    assert a.row_updates >= 200000                       # large enough for the threaded path
    for t in range(a.n_layers):
        for getter in ("layer_arrays", "layer_shard"):
            x, y = getattr(a, getter)(t), getattr(b, getter)(t)
            assert x.keys() == y.keys()
            for k in x:
                assert np.array_equal(x[k], y[k]), (t, k)
        assert a.layer_info(t) == b.layer_info(t)


@pytest.mark.parametrize("world", [1, 3])
def test_recycled_plan_storage_changes_nothing(gen, world):
    """A destroyed plan's arrays and the planner's scratch are reused by the next
    genlib_plan_create; the schedule must not depend on what was planned before."""
    from genlib_jl_b200 import _lib
    big = gen.synth.generate(6000, 12, 300, alpha=0.05, demes=2, migration=0.1, overlap=2, seed=11)
    small = gen.synth.generate(900, 25, 40, alpha=0.1, demes=1, migration=0.0, overlap=3, seed=12)
    pb, ps = gen.genealogy(big.as_columns()), gen.genealogy(small.as_columns())
    _lib.lib().genlib_release_cache()
    fresh_small = _plan_dump(gen, ps, small.probands, world)
    _lib.lib().genlib_release_cache()
    fresh_big = _plan_dump(gen, pb, big.probands, world)        # retires the big plan's storage
    assert _plan_dump(gen, ps, small.probands, world) == fresh_small   # small after big
    assert _plan_dump(gen, pb, big.probands, world) == fresh_big       # big after small
    assert _plan_dump(gen, pb, big.probands, 1 if world > 1 else 2) is not None
    assert _plan_dump(gen, ps, small.probands, world) == fresh_small   # after a different world size


@pytest.mark.parametrize("numerics", ["reference", "fp64"])
def test_replayed_schedule_equals_oracle_geneaji(gen, ob, numerics):
    ped = gen.genealogy(gen.geneaJi)
    plan = gen.Plan(ped.father, ped.mother, ped.rank_of(gen.pro(ped)))
    want = ob.OraclePedigree.from_csv(gen.geneaJi).phi()
    assert np.array_equal(replay(plan, numerics).astype(np.float32), want)


@pytest.mark.parametrize("seed", range(8))
def test_replayed_schedule_equals_oracle_random(gen, ob, seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(40, 400))
    rec = random_pedigree(rng, n, int(rng.integers(2, 12)), p_single=0.15, p_none=0.03,
                          window=int(rng.choice([0, 0, 30, 80])))
    ped = gen.genealogy(rec)
    o = ob.OraclePedigree.from_arrays(rec["ind"], rec["father"], rec["mother"], rec["sex"])
    pro = rng.permutation(ped.ids)[: int(rng.integers(1, 40))]
    pro = np.concatenate([pro, pro[:2]])                     # duplicates
    plan = gen.Plan(ped.father, ped.mother, ped.rank_of(pro))
    assert np.array_equal(replay(plan), o.phi(pro))


def test_replayed_schedule_deep_pedigree_rounding(gen, ob):
    """40 generations x 16: Float32 stores are lossy and Float64 sums inexact, so the
    reference's rounding schedule and rank grouping both matter."""
    s = gen.synth.generate(16 * 40, 40, 16, alpha=0.2, overlap=1, seed=11)
    ped = gen.genealogy(s.as_columns())
    o = ob.OraclePedigree.from_arrays(s.ind, s.father, s.mother, s.sex)
    plan = gen.Plan(ped.father, ped.mother, ped.rank_of(s.probands))
    want = o.phi(s.probands)
    got = replay(plan)
    assert np.array_equal(got, want)
    assert not np.array_equal(replay(plan, "fp64").astype(np.float32), want)   # the modes do differ


def test_synth_is_deterministic(gen):
    a = gen.synth.generate(5000, 10, 100, alpha=0.02, demes=3, migration=0.05, overlap=2, seed=9)
    b = gen.synth.generate(5000, 10, 100, alpha=0.02, demes=3, migration=0.05, overlap=2, seed=9)
    assert np.array_equal(a.father, b.father) and np.array_equal(a.probands, b.probands)
    assert (a.father[a.generation > 0] > 0).all() and (a.father < a.ind).all()
    c = gen.synth.generate(5000, 10, 100, alpha=0.02, demes=3, migration=0.05, overlap=2, seed=10)
    assert not np.array_equal(a.father, c.father)


def test_loader_edge_cases(gen, tmp_path):
    """C++ loader (`genlib_genealogy_csv/arrays`): file format of create.jl:161-189 and its errors."""
    p = tmp_path / "ped.asc"
    p.write_text("ind\tfather\tmother\tsex\r\n3 1 2 1\r\n\n1\t0\t0\t1\n2   0 0   2\n4 3 0 2")       # CRLF, blanks, no final newline
    ped = gen.genealogy(str(p))
    assert ped.ids.tolist() == [1, 2, 3, 4] and ped.father.tolist() == [-1, -1, 0, 2] and ped.mother.tolist() == [-1, -1, 1, -1]
    assert gen.pro(ped).tolist() == [4] and ped[4].father.ID == 3 and ped[4].mother is None
    with pytest.raises(KeyError):
        gen.genealogy(str(p), sort=False)                       # child before its parents (create.jl:240)
    with pytest.raises(KeyError):
        gen.genealogy({"ind": [1, 2], "father": [0, 7], "mother": [0, 0], "sex": [1, 1]})   # unknown parent
    with pytest.raises(ValueError):
        gen.genealogy({"ind": [1, 1], "father": [0, 0], "mother": [0, 0], "sex": [1, 1]})   # duplicate ID
    with pytest.raises(gen.GenlibError):
        gen.genealogy({"ind": [1, 2], "father": [2, 1], "mother": [0, 0], "sex": [1, 1]})   # cycle
    with pytest.raises(FileNotFoundError):
        gen.genealogy(str(tmp_path / "missing.asc"))
    (tmp_path / "bad.asc").write_text("h\n1 0 0\n")
    with pytest.raises(ValueError):
        gen.genealogy(str(tmp_path / "bad.asc"))
    empty = gen.genealogy({"ind": [], "father": [], "mother": [], "sex": []})
    assert len(empty) == 0 and gen.pro(empty).tolist() == []


def test_loader_id_maps_agree(gen):
    """Small integer IDs go through a direct table, sparse or negative ones through the hash map:
    the same pedigree under a relabelling must give the same ranks, parents and errors."""
    s = gen.synth.generate(20000, 10, 300, alpha=0.05, demes=2, migration=0.1, overlap=2, seed=31)
    dense = gen.genealogy(s.as_columns())

    def relabel(x, f):
        x = np.asarray(x, np.int64)
        return np.where(x > 0, f(x), 0)

    for f in (lambda x: x * 1000003 + 17, lambda x: -x - 5, lambda x: x + (1 << 40)):
        cols = {"ind": relabel(s.ind, f), "father": relabel(s.father, f), "mother": relabel(s.mother, f), "sex": s.sex}
        other = gen.genealogy(cols)
        assert np.array_equal(other.father, dense.father) and np.array_equal(other.mother, dense.mother)
        assert np.array_equal(other.ids, relabel(dense.ids, f))
        pro = relabel(s.probands, f)
        assert np.array_equal(other.rank_of(pro), dense.rank_of(s.probands))
        with pytest.raises(KeyError):
            other.rank_of(np.array([3], np.int64))               # not an ID of the relabelled pedigree
    with pytest.raises(KeyError):
        dense.rank_of(np.array([10 ** 12], np.int64))
    with pytest.raises(KeyError):
        dense.rank_of(np.array([-1], np.int64))


def test_loader_file_equals_columns_and_oracle(gen, ob, tmp_path):
    s = gen.synth.generate(30000, 12, 500, alpha=0.05, demes=3, migration=0.1, overlap=3, seed=21)
    path = str(tmp_path / "synth.asc")
    s.to_csv(path)
    a, b = gen.genealogy(path), gen.genealogy(s.as_columns())
    o = ob.OraclePedigree.from_csv(path)
    for x in (a, b):
        assert np.array_equal(x.ids, o.ids) and np.array_equal(x.father, o.father) and np.array_equal(x.mother, o.mother)
        assert np.array_equal(x.sex, o.sex)
    assert np.array_equal(gen.pro(a), o.pro())
    # a deep chain must not overflow any stack (the reference recurses, create.jl:196-209)
    n = 200000
    chain = gen.genealogy({"ind": np.arange(1, n + 1), "father": np.arange(0, n), "mother": np.zeros(n, int), "sex": np.ones(n, int)})
    assert chain.ids[0] == 1 and chain.ids[-1] == n and chain.depth() == n


def _stream_selftest(gen, ped, ranks, world, slack):
    import ctypes as C
    from genlib_b200 import engine as ge
    fa, mo = np.ascontiguousarray(ped.father, np.int32), np.ascontiguousarray(ped.mother, np.int32)
    pr = np.ascontiguousarray(ranks, np.int32)
    n_l, ovf = C.c_int32(0), C.c_int32(0)
    rc = gen.lib().genlib_plan_stream_selftest(len(fa), ge.ptr(fa), ge.ptr(mo), len(pr), ge.ptr(pr), world, float(slack),
                                               C.byref(n_l), C.byref(ovf))
    assert rc == 0, gen.lib().genlib_last_error().decode()
    return n_l.value, bool(ovf.value)


@pytest.mark.parametrize("world", [1, 2, 8])
def test_streamed_plan_equals_the_plan_made_in_one_piece(gen, world):
    """genlib_phi hands the plan to the device layer by layer while it is still being made: every published
    slice must be final, and the finished plan must be the one planned in one piece (csrc/plan.hpp PlanStream)."""
    ped = gen.genealogy(gen.genea140)
    ranks = ped.rank_of(gen.pro(ped))
    n_layers = gen.Plan(ped.father, ped.mother, ranks, world=world).n_layers
    done, overflow = _stream_selftest(gen, ped, ranks, world, 6.25)
    assert done == n_layers and not overflow
    for seed in (3, 11):
        ped = gen.genealogy(random_pedigree(np.random.default_rng(seed), 3000, 40, window=60 if seed == 3 else 400))
        ranks = ped.rank_of(ped.ids[-200:])
        done, overflow = _stream_selftest(gen, ped, ranks, world, 6.25)
        assert not overflow and done == gen.Plan(ped.father, ped.mother, ranks, world=world).n_layers


def test_planner_threads_do_not_change_the_plan(gen, monkeypatch):
    """The planner runs the layers as a pipeline of two lanes (slot assignment on the planning thread; live flags and
    the couples' parents on a helper that also groups couples ahead, csrc/plan.cpp): with 1, 2, 3 or 5 threads, made
    in one piece or handed over layer by layer, on one rank or several, under either schedule, the plan is the same
    to the last index (genlib_plan_digest).  GENLIB_PLAN_THREADS forces helpers on plans that small."""
    cases = []
    ped = gen.genealogy(gen.genea140)
    cases.append((ped, ped.rank_of(gen.pro(ped))))
    for seed, win in ((3, 60), (11, 400), (5, 0)):
        ped = gen.genealogy(random_pedigree(np.random.default_rng(seed), 3000, 40, window=win))
        cases.append((ped, ped.rank_of(ped.ids[-300:])))
    s = gen.synth.generate(60000, 30, 800, alpha=0.05, demes=4, migration=0.05, overlap=3, seed=5)
    ped = gen.genealogy(s.as_columns())
    cases.append((ped, ped.rank_of(s.probands)))
    for ped, ranks in cases:
        for world, schedule in ((1, "phi"), (3, "phi"), (8, "phi"), (2, "sparse_phi")):
            ids = ped.ids if schedule != "phi" else None
            digests = set()
            for threads in ("1", "2", "3", "5"):
                monkeypatch.setenv("GENLIB_PLAN_THREADS", threads)
                for stream in (False, True):
                    plan = gen.Plan(ped.father, ped.mother, ranks, world=world, schedule=schedule, ids=ids, stream=stream)
                    digests.add(int(gen.lib().genlib_plan_digest(plan._h, 0)))
            assert len(digests) == 1 and 0 not in digests, (world, schedule, digests)


@pytest.mark.parametrize("world", [1, 4])
def test_streamed_plan_with_a_bound_that_does_not_hold(gen, world):
    """A frontier bound that is too small on purpose: the planner notices, waits for the consumer and finishes
    with exact sizes; what was published until then stays valid."""
    ped = gen.genealogy(gen.genea140)
    ranks = ped.rank_of(gen.pro(ped))
    n_layers = gen.Plan(ped.father, ped.mother, ranks, world=world).n_layers
    done, overflow = _stream_selftest(gen, ped, ranks, world, -40.0)
    assert overflow and 0 < done < n_layers


@pytest.mark.parametrize("name,worlds", [("C3", (1, 2, 8)), ("C5", (1, 2)), ("C4", (8,))])
def test_streamed_plan_bounds_hold_at_benchmark_sizes(gen, name, worlds):
    """The frontier / row bounds a streamed plan hands its engine must hold on the workloads that are
    benchmarked (an overflow is handled, but costs the overlap of planning and device)."""
    s = gen.synth.config(name)
    ped = gen.genealogy(s.as_columns())
    ranks = ped.rank_of(s.probands)
    for world in worlds:
        done, overflow = _stream_selftest(gen, ped, ranks, world, 6.25)
        assert not overflow and done > 0


@pytest.mark.parametrize("world", [1, 3])
def test_async_plan_is_the_same_plan(gen, world):
    """Plan(..., stream=True) = genlib_plan_create_async: made on a worker thread; queries wait for it.  Without
    an engine reading it, it ends as the plan made in one piece (layers, arrays; the frontier width is the bound
    its engine would have been sized with)."""
    ped = gen.genealogy(gen.genea140)
    ranks = ped.rank_of(gen.pro(ped))
    a = gen.Plan(ped.father, ped.mother, ranks, world=world, stream=True)
    assert a.n_unique == 140                                   # known before the layers are
    b = gen.Plan(ped.father, ped.mother, ranks, world=world)
    assert a.n_layers == b.n_layers and a.row_updates == b.row_updates and a.capacity >= b.capacity
    for t in range(b.n_layers):
        x, y = a.layer_arrays(t), b.layer_arrays(t)
        assert x.keys() == y.keys()
        for k in x:
            if k == "live_flags":                              # (as wide as the frontier: the bound vs the exact width)
                assert np.array_equal(x[k][: len(y[k])], y[k]) and not x[k][len(y[k]):].any()
            else:
                assert np.array_equal(x[k], y[k]), (t, k)
        assert a.layer_info(t) == b.layer_info(t)
    with pytest.raises(KeyError):                              # validation errors surface at creation, as usual
        gen.Plan(ped.father, ped.mother, np.array([len(ped.father) + 5], np.int32), stream=True)
    del a, b


def test_planner_self_check_on_many_pedigrees(gen, monkeypatch):
    """GENLIB_PLAN_VERIFY: the index ranges the layer kernel relies on, and the sole-reader marks (a consumer
    drops such a strip-buffer row from L2 after staging it: nobody else may stage it), on real, synthetic and
    random pedigrees, one and several ranks, both schedules."""
    monkeypatch.setenv("GENLIB_PLAN_VERIFY", "1")
    peds = [gen.genealogy(gen.geneaJi), gen.genealogy(gen.genea140)]
    for name, scale in (("C3", 0.05), ("C4", 0.01), ("C5", 0.05)):
        s = gen.synth.config(name, scale)
        peds.append(gen.genealogy(s.as_columns()))
    rng = np.random.default_rng(9)
    for k in range(12):
        peds.append(gen.genealogy(random_pedigree(rng, int(rng.integers(50, 4000)), int(rng.integers(2, 60)),
                                                  p_single=0.15, p_none=0.03, window=int(rng.choice([0, 30, 300])))))
    # polygamy and big sibships: parents with several couples, couples split at 32 members and across tiles
    n = 3000
    ind = np.arange(1, n + 1)
    father = np.where(ind > 40, 1 + (ind % 7), 0)
    mother = np.where(ind > 40, 8 + (ind % 11), 0)
    peds.append(gen.genealogy({"ind": ind, "father": father, "mother": mother, "sex": np.where(ind <= 7, 1, 2)}))
    for ped in peds:
        ranks = ped.rank_of(gen.pro(ped))
        for world in (1, 3):
            for schedule in ("phi", "sparse_phi"):
                plan = gen.Plan(ped.father, ped.mother, ranks, world=world, schedule=schedule, ids=ped.ids)
                assert plan.n_layers > 0
