"""The oracle against every golden vector the reference's own tests hold for this path
(test/runtests.jl:41-60), the survey's genea140 checksum, and exact rationals."""
import hashlib

import numpy as np
import pytest

from util import exact_kinship, random_pedigree

GENEAJI_PHI = np.array([[0.591796875, 0.37109375, 0.072265625],
                        [0.37109375, 0.591796875, 0.072265625],
                        [0.072265625, 0.072265625, 0.53515625]], np.float32)   # runtests.jl:51-52


def test_geneaji_known_answers(gen, ob):
    p = ob.OraclePedigree.from_csv(gen.geneaJi)
    assert p.n == 29
    assert p.pro().tolist() == [1, 2, 29]                          # runtests.jl:41
    phi, steps = p.phi(with_steps=True)
    assert phi.dtype == np.float32 and np.array_equal(phi, GENEAJI_PHI)   # runtests.jl:50-52
    assert ob.phi_mean(phi) == 0.171875                            # runtests.jl:53
    assert p.phi_pair(1, 2) == 0.37109375                          # runtests.jl:49
    assert p.phi_pair(17, 19) == 0.0                               # runtests.jl:58-60
    # gen.f(ped, [1]) == 0.18359375 (runtests.jl:47): f = phi(father, mother) = 2 phi_ii - 1
    assert 2 * float(phi[0, 0]) - 1 == 0.18359375
    # (prev, next, carried) per step
    assert steps[:, :3].astype(int).tolist() == [[2, 4, 2], [4, 6, 4], [6, 7, 4], [7, 9, 4],
                                                 [9, 8, 0], [8, 4, 0], [4, 3, 0]]
    assert steps[:, 3].sum() == 156


def test_genea140_checksums(gen, ob, genea140_oracle):
    """No genea140 kinship value is pinned by the reference's tests; this pins the oracle to
    the survey's independent NumPy emulation."""
    p, phi, steps = genea140_oracle
    assert p.n == 41523
    pro = p.pro()
    assert len(pro) == 140 and pro[:3].tolist() == [217891, 218089, 219947] and pro[-1] == 868572
    assert hashlib.sha256(phi.tobytes()).hexdigest() == \
        "fe0313bf6871185b7c7f4ac42edaa5f50ef781d567722f27bc373b1bd013ddea"
    assert phi.astype(np.float64).sum() == 92.808104778639972
    assert np.trace(phi.astype(np.float64)) == 70.551006674766541
    assert phi[0, 0].view(np.uint32) == 0x3f00548c and phi[0, 1].view(np.uint32) == 0x3976b800
    assert phi[1, 1].view(np.uint32) == 0x3f0007bb                 # a lossy Float32 store
    assert int(steps[:, 3].sum()) == 405511361 and int(steps[:, 4].sum()) == 41513
    assert steps[:, 1].astype(int).tolist() == [88, 501, 1934, 4357, 7217, 10338, 12989, 13654, 12032,
                                                8795, 5779, 3557, 2030, 1071, 560, 280, 140]
    # siblings 10033 & 113470 (docs/src/tutorials.md:123-125) have kinship exactly 1/4
    sib = p.phi(np.array([10033, 113470]))
    assert sib[0, 1] == 0.25


def test_exact_rationals_on_genea140_pairs(gen, ob, genea140_oracle):
    p, phi, _ = genea140_oracle
    pro = p.pro()
    ranks = [int(np.nonzero(p.ids == i)[0][0]) for i in pro]
    ex = exact_kinship(p.father, p.mother)
    rng = np.random.default_rng(3)
    for _ in range(24):
        a, b = rng.integers(0, 140, 2)
        v = ex(ranks[a], ranks[b])
        assert np.float32(float(v)) == phi[a, b]       # fp64-exact here, one RN32


@pytest.mark.parametrize("seed", range(6))
def test_random_pedigrees_vs_exact(ob, seed):
    rng = np.random.default_rng(seed)
    rec = random_pedigree(rng, 60, 6, p_single=0.2, p_none=0.05)
    p = ob.OraclePedigree.from_arrays(rec["ind"], rec["father"], rec["mother"], rec["sex"])
    pro = rng.permutation(p.ids)[:12]
    phi = p.phi(pro)
    ex = exact_kinship(p.father, p.mother)
    rk = {int(i): r for r, i in enumerate(p.ids)}
    for a in range(12):
        for b in range(12):
            # shallow pedigrees: every value fits Float32, so the oracle must be exact
            assert float(phi[a, b]) == float(ex(rk[int(pro[a])], rk[int(pro[b])]))
            assert p.phi_pair(int(pro[a]), int(pro[b])) == float(phi[a, b])


def test_edge_cases(gen, ob):
    p = ob.OraclePedigree.from_csv(gen.geneaJi)
    assert p.phi(np.array([], np.int64)).shape == (0, 0)                   # empty proband list
    assert p.phi(np.array([17])).tolist() == [[0.5]]                       # single founder
    d = p.phi(np.array([1, 2, 1, 29, 2]))                                  # duplicates collapse
    assert np.array_equal(d, GENEAJI_PHI)
    r = p.phi(np.array([29, 1]))                                           # order = probandIDs
    assert np.array_equal(r, GENEAJI_PHI[np.ix_([2, 0], [2, 0])])
    anc = p.phi(np.array([1, 4, 17]))                                      # probands that are ancestors
    assert anc[0, 0] == GENEAJI_PHI[0, 0] and anc[2, 2] == 0.5
    with pytest.raises(KeyError):
        p.phi(np.array([1, 999]))
    with pytest.raises(KeyError):                                          # unsorted file, sort=false
        ob.OraclePedigree.from_arrays([1, 2, 3], [2, 0, 0], [3, 0, 0], [1, 1, 2], sort=False)


def test_threads_do_not_change_bits(gen, ob):
    p = ob.OraclePedigree.from_csv(gen.genea140)
    pro = p.pro()[:40]
    assert np.array_equal(p.phi(pro, nthreads=1), p.phi(pro, nthreads=4))
