"""Parity of the CUDA path (through the C ABI) with the oracle: bit-exact, every case.

Everything here runs on a B200 (`-m gpu`).  /root/reference is not read.
"""
import hashlib

import numpy as np
import pytest

from plan_replay import replay
from util import exact_kinship, ladder_pedigree, random_pedigree

pytestmark = pytest.mark.gpu

GENEAJI_PHI = np.array([[0.591796875, 0.37109375, 0.072265625],
                        [0.37109375, 0.591796875, 0.072265625],
                        [0.072265625, 0.072265625, 0.53515625]], np.float32)   # runtests.jl:51-52


def assert_bit_equal(got, want):
    assert got.dtype == want.dtype and got.shape == want.shape
    if not np.array_equal(got.view(np.uint32 if got.dtype == np.float32 else np.uint64),
                          want.view(np.uint32 if want.dtype == np.float32 else np.uint64)):
        bad = np.argwhere(got != want)
        i, j = bad[0]
        raise AssertionError(f"{len(bad)} of {got.size} entries differ; first at ({i},{j}): "
                             f"{got[i, j]!r} vs {want[i, j]!r}")


def test_native_library_is_loaded(gen):
    assert gen.lib().genlib_device_count() >= 1
    maps = open("/proc/self/maps").read()
    assert "libgenlib_cuda.so" in maps


def test_geneaji_golden(gen):
    ped = gen.genealogy(gen.geneaJi)
    phi = gen.phi(ped)
    assert_bit_equal(phi, GENEAJI_PHI)                         # runtests.jl:50-52
    assert gen.phiMean(phi) == np.float32(0.171875)            # runtests.jl:53
    out, stats = gen.phi_arrays(ped.father, ped.mother, ped.rank_of([1, 2, 29]))   # one-shot C ABI
    assert_bit_equal(out, GENEAJI_PHI)
    assert stats["n_unique"] == 3 and stats["kernel_launches"] > 0
    assert_bit_equal(gen.phi(ped, numerics="fp64", dtype=np.float64), GENEAJI_PHI.astype(np.float64))


def test_geneaji_every_layer_equals_replay(gen):
    """Frontier after each layer == NumPy replay of the same schedule (localises a kernel bug)."""
    ped = gen.genealogy(gen.geneaJi)
    plan = gen.Plan(ped.father, ped.mother, ped.rank_of(gen.pro(ped)))
    states = {}
    replay(plan, on_layer=lambda t, A, live: states.__setitem__(t, (A[np.ix_(live, live)].copy(), live)))
    eng = gen.Engine(plan)
    for t in range(plan.n_layers):
        eng.run(layer_limit=t + 1)
        want, live = states[t]
        got = eng.read_block(live).astype(np.float32)
        assert_bit_equal(got, want)
    eng.close()


def test_genea140_equals_oracle(gen, genea140_oracle):
    _, want, _ = genea140_oracle
    ped = gen.genealogy(gen.genea140)
    got, stats = gen.phi(ped, return_stats=True)
    assert_bit_equal(got, want)
    assert hashlib.sha256(got.tobytes()).hexdigest() == \
        "fe0313bf6871185b7c7f4ac42edaa5f50ef781d567722f27bc373b1bd013ddea"
    assert stats["row_updates"] == 41523 and stats["n_layers"] == 18
    # Float64 storage rounds once at the end; on genea140 that is the same matrix
    assert_bit_equal(gen.phi(ped, numerics="fp64"), want)


def test_genea140_subsets_and_order(gen, ob):
    ped = gen.genealogy(gen.genea140)
    o = ob.OraclePedigree.from_csv(gen.genea140)
    pro = gen.pro(ped)
    rng = np.random.default_rng(1)
    sel = rng.permutation(pro)[:17]
    sel = np.concatenate([sel, sel[:3], [10086, 33724]])          # duplicates, a founder, an ancestor
    assert_bit_equal(gen.phi(ped, sel), o.phi(sel))


@pytest.mark.parametrize("seed", range(12))
def test_random_pedigrees(gen, ob, seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(30, 1500))
    rec = random_pedigree(rng, n, int(rng.integers(2, 40)), p_single=0.15, p_none=0.03,
                          window=int(rng.choice([0, 0, 40, 200])))
    ped = gen.genealogy(rec)
    o = ob.OraclePedigree.from_arrays(rec["ind"], rec["father"], rec["mother"], rec["sex"])
    k = int(rng.integers(1, min(n, 300)))
    pro = rng.permutation(ped.ids)[:k]
    assert_bit_equal(gen.phi(ped, pro), o.phi(pro))
    if seed % 3 == 0:
        assert_bit_equal(gen.phi(ped), o.phi())                    # default probands = gen.pro


@pytest.mark.parametrize("name,scale", [("C3", 0.02), ("C4", 0.004), ("C5", 0.02)])
def test_named_configs_scaled(gen, ob, name, scale):
    s = gen.synth.config(name, scale)
    ped = gen.genealogy(s.as_columns())
    o = ob.OraclePedigree.from_arrays(s.ind, s.father, s.mother, s.sex)
    got, stats = gen.phi(ped, s.probands, return_stats=True)
    assert_bit_equal(got, o.phi(s.probands))
    assert stats["n_unique"] == len(s.probands)


@pytest.mark.parametrize("knobs", [{"GENLIB_PROD_WARPS": "4"}, {"GENLIB_PROD_WARPS": "8"}, {"GENLIB_GANGS": "1"},
                                   {"GENLIB_GANGS": "8", "GENLIB_MAX_SW": "16"}, {"GENLIB_DISCARD": "0"},
                                   {"GENLIB_MAX_NBUF": "2", "GENLIB_CTAS_PER_ROLE": "37"}])
def test_launch_shapes_do_not_change_the_bits(gen, ob, monkeypatch, knobs):
    """Every launch shape of the layer kernel the engine can choose (producer / consumer warp split, gangs,
    strip width, buffers, L2 discard) gives the oracle's bits, with and without carried individuals."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    for name, scale in (("C3", 0.03), ("C5", 0.03)):
        s = gen.synth.config(name, scale)
        ped = gen.genealogy(s.as_columns())
        o = ob.OraclePedigree.from_arrays(s.ind, s.father, s.mother, s.sex)
        assert_bit_equal(gen.phi(ped, s.probands), o.phi(s.probands))
    rng = np.random.default_rng(41)
    ped_cols = random_pedigree(rng, 1500, 20, p_single=0.1, p_none=0.02, window=200)
    ped = gen.genealogy(ped_cols)
    pro = rng.permutation(ped.ids)[:200]
    assert_bit_equal(gen.phi(ped, pro), ob.OraclePedigree.from_arrays(ped_cols["ind"], ped_cols["father"], ped_cols["mother"],
                                                                     ped_cols["sex"]).phi(pro))


def test_deep_pedigree_rounding_schedule(gen, ob):
    """Float32 stores are lossy and Float64 sums inexact here: only the
    reference's rounding points and rank grouping reproduce the matrix."""
    s = gen.synth.generate(16 * 60, 60, 16, alpha=0.2, overlap=1, seed=11)
    ped = gen.genealogy(s.as_columns())
    o = ob.OraclePedigree.from_arrays(s.ind, s.father, s.mother, s.sex)
    want = o.phi(s.probands)
    assert_bit_equal(gen.phi(ped, s.probands), want)
    assert not np.array_equal(gen.phi(ped, s.probands, numerics="fp64"), want)
    # 300 generations: values reach Float32 subnormals (gradual underflow must be kept)
    s = gen.synth.generate(12 * 300, 300, 12, alpha=0.0, overlap=1, seed=3)
    ped = gen.genealogy(s.as_columns())
    o = ob.OraclePedigree.from_arrays(s.ind, s.father, s.mother, s.sex)
    assert_bit_equal(gen.phi(ped, s.probands), o.phi(s.probands))


def test_fp64_mode_is_exact_on_shallow_pedigrees(gen):
    rng = np.random.default_rng(5)
    rec = random_pedigree(rng, 400, 30, p_single=0.1, window=60)
    ped = gen.genealogy(rec)
    pro = rng.permutation(ped.ids)[:40]
    got = gen.phi(ped, pro, numerics="fp64", dtype=np.float64)
    ex = exact_kinship(ped.father, ped.mother)
    rk = ped.rank_of(pro)
    for a in range(0, 40, 3):
        for b in range(40):
            v = ex(int(rk[a]), int(rk[b]))
            if v.denominator.bit_length() <= 50:
                assert got[a, b] == float(v)


def test_edge_cases(gen, ob):
    ped = gen.genealogy(gen.geneaJi)
    assert gen.phi(ped, []).shape == (0, 0)
    assert gen.phi(ped, [17]).tolist() == [[0.5]]
    assert_bit_equal(gen.phi(ped, [1, 2, 1, 29, 2]), GENEAJI_PHI)
    assert_bit_equal(gen.phi(ped, [29, 1]), GENEAJI_PHI[np.ix_([2, 0], [2, 0])])
    o = ob.OraclePedigree.from_csv(gen.geneaJi)
    for sel in ([1, 4, 17], [17, 19, 20, 23, 25, 26], list(range(1, 30)), [9, 11]):
        assert_bit_equal(gen.phi(ped, sel), o.phi(np.array(sel)))
    with pytest.raises(KeyError):
        gen.phi(ped, [1, 999])
    # a 1000-child sibship (split into couples of <= 32 members) and 3000 unrelated founders
    n = 3002 + 1000
    father = np.full(n, -1, np.int32); mother = np.full(n, -1, np.int32)
    father[3002:] = 0; mother[3002:] = 1
    pro = np.arange(2, n, dtype=np.int32)
    got, _ = gen.phi_arrays(father, mother, pro)
    want, _ = ob.phi_ranks(father, mother, pro)
    assert_bit_equal(got, want)


def test_out_dtype_and_mean(gen):
    ped = gen.genealogy(gen.genea140)
    plan = gen.Plan(ped.father, ped.mother, ped.rank_of(gen.pro(ped)))
    eng = gen.Engine(plan)
    eng.run(time_layers=True)
    f32 = eng.fetch()
    f64 = eng.fetch(dtype=np.float64)
    assert_bit_equal(f64, f32.astype(np.float64))
    mean = eng.phi_mean()
    want = (f64.sum() - np.trace(f64)) / (140 * 139)
    assert abs(mean - want) < 1e-15
    assert abs(mean - 0.0011437357710109676) < 1e-15
    infos = [eng.layer_info(t) for t in range(plan.n_layers)]
    assert all(i["ms_layer"] > 0 and i["dram_write_bytes"] > 0 and i["strip_width"] > 0 for i in infos)
    assert all(i["dram_read_bytes"] > 0 and i["l2_bytes"] > 0 for i in infos[1:]) and infos[0]["dram_read_bytes"] == 0
    eng.close()


@pytest.mark.parametrize("tag,name,scale", [("c3_full", "C3", 1.0), ("c4_x0.1", "C4", 0.1), ("c5_x0.1", "C5", 0.1)])
def test_benchmark_size_output_equals_the_oracles_golden_hash(gen, tag, name, scale):
    """Parity AT THE SIZES bench.py measures: the sha256 of the proband matrix equals the one the oracle
    alone produced (tests/golden/make_golden.py: full C3 takes the oracle minutes, the GPU 0.1 s)."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", tag + ".sha256")
    if not os.path.exists(path):
        pytest.skip(f"{tag}: golden hash not generated")
    g = json.load(open(path))
    s = gen.synth.config(name, scale)
    ped = gen.genealogy(s.as_columns())
    got, stats = gen.phi(ped, s.probands, return_stats=True)
    # (the golden record counts the individuals born after the top level, like the oracle's steps)
    top = gen.Plan(ped.father, ped.mother, ped.rank_of(s.probands)).layer_info(0)["n_new"]
    assert stats["row_updates"] - top == g["row_updates"] and got.shape == (g["n"], g["n"])
    assert hashlib.sha256(got.tobytes()).hexdigest() == g["sha256"]
    assert abs(float(got.astype(np.float64).sum()) - g["sum"]) < 1e-6 * g["sum"]
    ngpu = gen.lib().genlib_device_count()
    if ngpu >= 2 and name != "C5":                                  # the same bits from several devices
        multi = gen.phi(ped, s.probands, devices=list(range(min(ngpu, 8))))
        assert hashlib.sha256(multi.tobytes()).hexdigest() == g["sha256"]


def test_row_sums_and_device_mean_are_deterministic(gen):
    """phiMean on the device: binary64 in a fixed order, per-row sums that add up the same on any number
    of ranks; the host mirror of gen.phiMean(::Matrix{Float32}) for comparison."""
    s = gen.synth.config("C3", 0.05)
    ped = gen.genealogy(s.as_columns())
    ranks = ped.rank_of(s.probands)
    plan = gen.Plan(ped.father, ped.mother, ranks)
    eng = gen.Engine(plan)
    eng.run()
    m = eng.fetch()
    rs = eng.row_sums()
    assert rs.shape == (plan.n_unique, 2) and np.array_equal(rs[:, 1], np.diag(m).astype(np.float64))
    assert np.allclose(rs[:, 0], m.astype(np.float64).sum(1), rtol=1e-14, atol=0)
    mean = eng.phi_mean()
    n = plan.n_unique
    assert mean == (float(np.sum(rs[:, 0])) - float(np.sum(rs[:, 1]))) / (n * n - n) or \
        abs(mean - (m.astype(np.float64).sum() - np.trace(m.astype(np.float64))) / (n * n - n)) < 1e-15
    assert eng.phi_mean() == mean                                   # reproducible
    assert abs(float(gen.phiMean(m)) - mean) < 1e-6 * mean          # Float32 pairwise sum vs binary64
    eng.close()


def test_full_size_properties(gen):
    """C3 at 1/4 scale (10k individuals per generation): too big for the oracle in a test,
    checked through size-independent properties."""
    s = gen.synth.config("C3", 0.25)
    ped = gen.genealogy(s.as_columns())
    a, stats = gen.phi(ped, s.probands, return_stats=True)
    b = gen.phi(ped, s.probands)
    assert_bit_equal(a, b)                                         # deterministic
    assert np.array_equal(a, a.T)                                  # symmetric
    d = np.diag(a)
    assert (d >= 0.5).all() and (d < 1).all() and (a >= 0).all() and (a <= d[:, None] + 1e-7).all()
    # full siblings among the probands have identical kinship with everybody else
    rk = ped.rank_of(s.probands)
    key = ped.father[rk].astype(np.int64) * (len(ped) + 1) + ped.mother[rk]
    order = np.argsort(key, kind="stable")
    same = np.nonzero(key[order][1:] == key[order][:-1])[0]
    assert len(same) > 10
    for q in same[:200]:
        i, j = order[q], order[q + 1]
        mask = np.ones(len(rk), bool); mask[[i, j]] = False
        assert np.array_equal(a[i, mask], a[j, mask])
    # the diagonal is 1/2 (1 + kinship of the parents): check through a second run that
    # asks for the parents themselves (fp64 storage so both runs are exact to ~2^-40)
    sub = s.probands[:50]
    parents = np.unique(np.concatenate([s.father[sub - 1], s.mother[sub - 1]]))
    pp = gen.phi(ped, parents, numerics="fp64", dtype=np.float64)
    dd = gen.phi(ped, sub, numerics="fp64", dtype=np.float64)
    pos = {int(x): k for k, x in enumerate(parents)}
    for k, x in enumerate(sub):
        f, m = pos[int(s.father[x - 1])], pos[int(s.mother[x - 1])]
        assert abs(dd[k, k] - 0.5 * (1 + pp[f, m])) < 1e-12


def test_multi_gpu_equals_single_gpu(gen):
    """Row-sharded engine on every visible GPU (one rank per GPU under torchrun) == one GPU ==
    oracle, bit for bit (tests/dist_check.py).  Needs >= 2 GPUs; the 1-GPU driver run skips it."""
    import os
    import subprocess
    import sys
    n = gen.lib().genlib_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert res.returncode == 0 and "dist_check: all equal" in res.stdout, (res.stdout + res.stderr)[-3000:]


def test_one_process_several_devices(gen, ob):
    """genlib_phi_multi (one process, one host thread per device, NVLink peer access) == one device,
    bit for bit; on a 1-GPU box the call degenerates to genlib_phi."""
    ngpu = gen.lib().genlib_device_count()
    devs = list(range(min(ngpu, 4)))
    ped = gen.genealogy(gen.genea140)
    want = gen.phi(ped)
    assert_bit_equal(gen.phi(ped, devices=devs), want)
    assert_bit_equal(gen.phi(ped, devices=devs[::-1], numerics="fp64"), want)
    s = gen.synth.config("C3", 0.05)
    ped = gen.genealogy(s.as_columns())
    want = ob.OraclePedigree.from_arrays(s.ind, s.father, s.mother, s.sex).phi(s.probands)
    got, stats = gen.phi(ped, s.probands, devices=devs, return_stats=True)
    assert_bit_equal(got, want)
    assert stats["d2h_bytes"] == got.nbytes
    rng = np.random.default_rng(77)
    ped = gen.genealogy(random_pedigree(rng, 900, 12, p_single=0.15, p_none=0.03, window=60))
    pro = rng.permutation(ped.ids)[:150]
    assert_bit_equal(gen.phi(ped, pro, devices=devs), gen.phi(ped, pro))
    with pytest.raises(Exception):
        gen.phi(ped, pro, devices=[0, 0])


def test_one_call_streams_the_plan_to_the_device(gen, ob, monkeypatch):
    """genlib_phi plans on a worker thread and runs every layer as soon as it is planned: the same bits as
    the engine on the finished plan -- also when the frontier bound of the streamed plan does not hold
    (GENLIB_STREAM_SLACK_PCT < 0: the call starts over on the finished plan) and with streaming off."""
    cases = []
    ped = gen.genealogy(gen.genea140)
    cases.append((ped, ped.rank_of(gen.pro(ped))))
    for name, scale in (("C3", 0.05), ("C5", 0.05)):
        s = gen.synth.config(name, scale)
        ped = gen.genealogy(s.as_columns())
        cases.append((ped, ped.rank_of(s.probands)))
    rng = np.random.default_rng(5)
    ped = gen.genealogy(random_pedigree(rng, 2500, 30, p_single=0.1, p_none=0.02, window=300))
    cases.append((ped, ped.rank_of(rng.permutation(ped.ids)[:300])))
    for ped, ranks in cases:
        plan = gen.Plan(ped.father, ped.mother, ranks)
        eng = gen.Engine(plan)
        eng.run()
        want = eng.fetch()
        eng.close()
        for env in ({}, {"GENLIB_STREAM_SLACK_PCT": "-40"}, {"GENLIB_STREAM": "0"}):
            for k in ("GENLIB_STREAM_SLACK_PCT", "GENLIB_STREAM"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got, stats = gen.phi_arrays(ped.father, ped.mother, ranks)
            assert_bit_equal(got, want)
            assert stats["row_updates"] == plan.row_updates and stats["kernel_launches"] > 0
            got64, _ = gen.phi_arrays(ped.father, ped.mother, ranks, numerics="fp64", dtype=np.float64)
            assert np.abs(got64 - want).max() < 1e-6
    for k in ("GENLIB_STREAM_SLACK_PCT", "GENLIB_STREAM"):
        monkeypatch.delenv(k, raising=False)


def test_inbreeding_f_genea140(gen, ob):
    """gen.f on genea140 against the reference's own definition, phi(father, mother) by the pairwise
    Karigl recursion (src/compute.jl:66-95, restated in the oracle), for a sample of individuals."""
    ped = gen.genealogy(gen.genea140)
    o = ob.OraclePedigree.from_csv(gen.genea140)
    rng = np.random.default_rng(5)
    deep = np.nonzero((ped.father >= 0) & (ped.mother >= 0))[0]
    ids = ped.ids[rng.choice(deep[-4000:], 40, replace=False)]
    got = gen.f(ped, ids)
    for v, ID in zip(got, ids):
        x = ped[int(ID)]
        assert v == np.float32(o.phi_pair(x.father.ID, x.mother.ID))
    assert got.dtype == np.float32 and (got > 0).any()


def test_inbreeding_f(gen, ob):
    """gen.f (src/compute.jl:500-511) from one engine sweep; known answers of test/runtests.jl:47-48."""
    ped = gen.genealogy(gen.geneaJi)
    assert gen.f(ped, [1]).tolist() == [0.18359375] and gen.f(ped, [17]).tolist() == [0.0]
    got = gen.f(ped, ped.ids)
    o = ob.OraclePedigree.from_csv(gen.geneaJi)
    for ID, v in zip(ped.ids, got):
        x = ped[int(ID)]
        want = 0.0 if x.father is None or x.mother is None else o.phi_pair(x.father.ID, x.mother.ID)
        assert v == np.float32(want)


# ---- gen.sparse_phi: the same engine on sparse_phi's own schedule ----------

def test_sparse_phi_geneaji_golden(gen):
    ped = gen.genealogy(gen.geneaJi)
    k = gen.sparse_phi(ped)
    assert gen.phiMean(k) == np.float32(0.171875)                                # test/runtests.jl:55
    assert repr(k) == "3×3 KinshipMatrix with 6 stored entries."                # :56
    assert k[1, 2] == np.float32(0.37109375)                                     # :57
    assert_bit_equal(np.array([[k[a, b] for b in (1, 2, 29)] for a in (1, 2, 29)], np.float32), GENEAJI_PHI)


@pytest.mark.parametrize("seed", range(6))
def test_sparse_phi_random_pedigrees(gen, ob, seed):
    rng = np.random.default_rng(300 + seed)
    n = int(rng.integers(100, 1500))
    rec = random_pedigree(rng, n, int(rng.integers(3, 30)), window=int(rng.choice([0, 0, 40, 120])))
    ped = gen.genealogy(rec)
    pro = rng.permutation(ped.ids)[: int(rng.integers(2, min(n, 200)))]
    ranks = ped.rank_of(pro)
    want, info = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, full=True)
    k = gen.sparse_phi(ped, pro)
    order = np.argsort(ranks, kind="stable")                                     # the KinshipMatrix is kept in rank order
    assert_bit_equal(k._dense, np.ascontiguousarray(want[np.ix_(order, order)]))
    assert k.stored == info["findable"] and len(k) == len(pro)
    assert k[int(pro[0]), int(pro[-1])] == want[0, -1]
    # the consistent variant of the schedule: nothing misfiled (compute.jl:393), founders still by ID
    sym = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, directed=False)[0]
    k = gen.sparse_phi(ped, pro, symmetric=True)
    assert_bit_equal(k._dense, np.ascontiguousarray(sym[np.ix_(order, order)]))


def test_sparse_phi_misfiled_kinship_and_founder_id_order(gen, ob):
    """The two properties of the reference's sparse_phi that the first round missed: kinships filed
    under phi[earlier][later] but looked up under phi[lower rank][higher rank] are lost
    (compute.jl:393), and the founders enter the queue sorted by ID (identify.jl:15-19)."""
    cols = {"ind": np.array([1, 2, 3, 4, 5, 6, 7]), "father": np.array([0, 0, 0, 0, 3, 1, 5]),
            "mother": np.array([0, 0, 0, 0, 4, 3, 6]), "sex": np.array([1, 2, 1, 2, 1, 2, 1], np.int32)}
    ped = gen.genealogy(cols)
    k = gen.sparse_phi(ped, [5, 6, 7])
    assert k[5, 6] == 0 and k[7, 7] == np.float32(0.5)               # the reference loses phi[5, 6] = 1/8
    k = gen.sparse_phi(ped, [5, 6, 7], symmetric=True)
    assert k[5, 6] == np.float32(0.125) and k[7, 7] == np.float32(0.5625)
    for seed, n, window in ((303, 400, 40), (305, 2000, 40), (302, 1000, 120)):
        rng = np.random.default_rng(seed)
        ped = gen.genealogy(random_pedigree(rng, n, 8 if n < 2000 else 10, window=window))
        pro = rng.permutation(ped.ids)[:100]
        ranks = ped.rank_of(pro)
        order = np.argsort(ranks, kind="stable")
        for symmetric in (False, True):
            want = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, directed=not symmetric)[0]
            k = gen.sparse_phi(ped, pro, symmetric=symmetric)
            assert_bit_equal(k._dense, np.ascontiguousarray(want[np.ix_(order, order)]))


def test_sparse_phi_deep_pedigrees_and_subnormals(gen, ob):
    s = gen.synth.generate(16 * 60, 60, 16, alpha=0.2, overlap=1, seed=11)
    ped = gen.genealogy(s.as_columns())
    ranks = ped.rank_of(s.probands)
    want, _ = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids)
    plan = gen.Plan(ped.father, ped.mother, ranks, schedule="sparse_phi", ids=ped.ids)
    eng = gen.Engine(plan)
    eng.run()
    assert_bit_equal(eng.fetch(), want)
    eng.close()
    assert not np.array_equal(gen.phi(ped, s.probands), want)                    # phi rounds elsewhere
    cols, pro = ladder_pedigree(73)                                              # Float32 halving of subnormals
    ped = gen.genealogy(cols)
    ranks = ped.rank_of(pro)
    want, _ = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids)
    plan = gen.Plan(ped.father, ped.mother, ranks, schedule="sparse_phi", ids=ped.ids)
    eng = gen.Engine(plan)
    eng.run()
    assert_bit_equal(eng.fetch(), want)
    eng.close()
    dense, _ = ob.phi_ranks(ped.father, ped.mother, ranks)
    assert_bit_equal(gen.phi(ped, pro), dense)
    assert not np.array_equal(dense, want)
    with pytest.raises(Exception):
        gen.Engine(plan, "fp64")                                                 # sparse_phi stores Float32


def test_sparse_phi_genea140_and_scaled_configs(gen, ob):
    ped = gen.genealogy(gen.genea140)
    pro = gen.pro(ped)
    ranks = ped.rank_of(pro)
    want, info = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids, full=True)
    k = gen.sparse_phi(ped)
    order = np.argsort(ranks, kind="stable")
    assert_bit_equal(k._dense, np.ascontiguousarray(want[np.ix_(order, order)]))
    assert k.stored == info["findable"]
    for name, scale in (("C3", 0.02), ("C5", 0.02)):
        s = gen.synth.config(name, scale)
        ped = gen.genealogy(s.as_columns())
        ranks = ped.rank_of(s.probands)
        want, _ = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids)
        plan = gen.Plan(ped.father, ped.mother, ranks, schedule="sparse_phi", ids=ped.ids)
        eng = gen.Engine(plan)
        eng.run()
        assert_bit_equal(eng.fetch(), want)
        eng.close()
