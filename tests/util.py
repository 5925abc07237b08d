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
