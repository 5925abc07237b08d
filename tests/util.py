"""Shared helpers for the tests: random pedigrees and an exact-rational Karigl recursion."""
from __future__ import annotations

from fractions import Fraction
from functools import lru_cache

import numpy as np


def random_pedigree(rng: np.random.Generator, n: int, n_founders: int, p_single: float = 0.1,
                    p_none: float = 0.02, window: int = 0):
    """Random pedigree in file order with parents before children.

    IDs are a random permutation of 1..3n (so ID order != file order), single-parent
    and parentless late individuals occur, generations overlap freely; `window` > 0
    restricts parents to the previous `window` individuals (deep, narrow pedigrees).
    Returns dict(ind, father, mother, sex) with parents as IDs (0 = unknown).
    """
    ids = rng.permutation(np.arange(1, 3 * n + 1))[:n].astype(np.int64)
    sex = rng.integers(1, 3, n).astype(np.int32)
    father = np.zeros(n, np.int64)
    mother = np.zeros(n, np.int64)
    for i in range(n_founders, n):
        lo = max(0, i - window) if window else 0
        males = [j for j in range(lo, i) if sex[j] == 1]
        females = [j for j in range(lo, i) if sex[j] == 2]
        u = rng.random()
        if u < p_none:
            continue
        if males and not (p_none <= u < p_none + p_single / 2):
            father[i] = ids[males[rng.integers(len(males))]]
        if females and not (p_none + p_single / 2 <= u < p_none + p_single):
            mother[i] = ids[females[rng.integers(len(females))]]
    return {"ind": ids, "father": father, "mother": mother, "sex": sex}


def exact_kinship(father: np.ndarray, mother: np.ndarray):
    """Memoised Karigl recursion (src/compute.jl:66-95) in exact rationals on rank arrays."""
    father = [int(x) for x in father]
    mother = [int(x) for x in mother]

    @lru_cache(maxsize=None)
    def phi(i: int, j: int) -> Fraction:
        if i < j:
            i, j = j, i
        if i == j:
            v = Fraction(1, 2)
            if father[i] >= 0 and mother[i] >= 0:
                v += phi(father[i], mother[i]) / 2
            return v
        v = Fraction(0)
        if father[i] >= 0:
            v += phi(father[i], j) / 2
        if mother[i] >= 0:
            v += phi(mother[i], j) / 2
        return v

    return phi


def ladder_pedigree(K: int):
    """Three lineages from one founder couple / a half-sib, K generations each, always married to
    fresh founders, and nine children of the last members of the first two lineages.

    Kinships between the lineages halve every generation (2**-(2K+3) after K), so around K = 73 they
    are Float32 subnormals with a single low bit, and a child of two lineages sums TWO such halves:
    `x / 2` in Float32 (what sparse_phi does) and in Float64 (what phi does) then differ.
    Returns (columns dict, probands)."""
    ind, fa, mo, sex = [], [], [], []

    def add(f, m, s):
        ind.append(len(ind) + 1); fa.append(f); mo.append(m); sex.append(s)
        return ind[-1]

    p1, p2, q = add(0, 0, 1), add(0, 0, 2), add(0, 0, 2)
    A, B, Cc = [add(p1, p2, 1)], [add(p1, p2, 2)], [add(p1, q, 2)]
    for _ in range(K):
        A.append(add(A[-1], add(0, 0, 2), 1))
        B.append(add(add(0, 0, 1), B[-1], 2))
        Cc.append(add(add(0, 0, 1), Cc[-1], 2))
    D = [add(A[-1 - i], B[-1 - j], 1) for i in range(3) for j in range(3)]
    cols = {"ind": np.array(ind, np.int64), "father": np.array(fa, np.int64), "mother": np.array(mo, np.int64),
            "sex": np.array(sex, np.int32)}
    return cols, np.array(D + [Cc[-1], Cc[-2], Cc[-3], A[-1], B[-1]], np.int64)
