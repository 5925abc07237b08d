"""Property test (hypothesis): arbitrary small pedigrees -- any mix of unknown/single/both parents,
polygamy, overlapping generations, probands anywhere in the pedigree -- replayed from the planner's
schedule must equal the oracle, for both reference schedules and 1-3 ranks."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from plan_replay import replay, replay_sharded


@st.composite
def pedigrees(draw):
    n = draw(st.integers(2, 72))
    n_founders = draw(st.integers(1, min(n, 8)))
    father, mother = [-1] * n, [-1] * n
    sex = [draw(st.integers(1, 2)) for _ in range(n)]
    sex[0] = 1
    if n > 1:
        sex[1] = 2
    window = draw(st.sampled_from([0, 0, 3, 5, 8]))         # > 0: deep, narrow pedigrees (lossy Float32 stores)
    for i in range(n_founders, n):
        lo = max(0, i - window) if window else 0
        males = [j for j in range(lo, i) if sex[j] == 1]
        females = [j for j in range(lo, i) if sex[j] == 2]
        kind = draw(st.integers(0, 9))                      # 0: no parent, 1: father only, 2: mother only, else both
        if kind != 0 and kind != 2 and males:
            father[i] = draw(st.sampled_from(males))
        if kind != 0 and kind != 1 and females:
            mother[i] = draw(st.sampled_from(females))
    k = draw(st.integers(1, min(n, 12)))
    probands = draw(st.lists(st.integers(0, n - 1), min_size=k, max_size=k))
    return np.array(father, np.int32), np.array(mother, np.int32), np.array(probands, np.int32)


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(ped=pedigrees(), world=st.integers(1, 3), threads=st.sampled_from([None, "2", "3", "4"]))
def test_any_small_pedigree_both_schedules(gen, ob, ped, world, threads):
    """`threads`: the planner's helper threads forced on these small plans too (its lanes must not change the plan)."""
    import os
    if threads is None:
        os.environ.pop("GENLIB_PLAN_THREADS", None)
    else:
        os.environ["GENLIB_PLAN_THREADS"] = threads
    try:
        _check(gen, ob, ped, world)
    finally:
        os.environ.pop("GENLIB_PLAN_THREADS", None)


def _check(gen, ob, ped, world):
    father, mother, probands = ped
    # parents precede children by construction, but ranks must also follow the depth order of
    # gen.genealogy (create.jl:217-227) for sparse_phi's queue; re-rank through the loader
    n = len(father)
    ids = np.arange(1, n + 1, dtype=np.int64)
    cols = {"ind": ids, "father": np.where(father >= 0, father + 1, 0), "mother": np.where(mother >= 0, mother + 1, 0),
            "sex": np.ones(n, np.int32)}
    pedg = gen.genealogy(cols)
    ranks = pedg.rank_of(ids[probands])
    want = {"phi": ob.phi_ranks(pedg.father, pedg.mother, ranks)[0],
            "sparse_phi": ob.sparse_phi_ranks(pedg.father, pedg.mother, ranks, ids=pedg.ids)[0]}
    for schedule in ("phi", "sparse_phi"):
        plan = gen.Plan(pedg.father, pedg.mother, ranks, world=world, schedule=schedule, ids=pedg.ids)
        got = replay(plan) if world == 1 else replay_sharded(plan)
        assert np.array_equal(got, want[schedule]), (schedule, world)
