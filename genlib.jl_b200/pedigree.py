"""Host-side mirror of the reference's pedigree type, on flat arrays.

Mirrors (reference, relative to /root/reference):
  src/create.jl:39-46,60-74   Individual / Pedigree (ordered, indexable by ID)
  src/create.jl:131-146       genealogy(::DataFrame; sort)
  src/create.jl:161-189       genealogy(::String; sort)   (4-column whitespace file)
  src/create.jl:196-227       rank = stable sort by maximum ancestral depth
  src/create.jl:234-254       _finalize_pedigree (parents must precede children)
  src/identify.jl:35-39       pro
The hot path only needs the rank-ordered parent arrays; objects are created
lazily on indexing so that 5 M-individual pedigrees stay cheap.
"""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np


class Individual:
    """View of one individual (src/create.jl:39-46).  `rank` is 1-based as in Julia."""

    __slots__ = ("_ped", "_r")

    def __init__(self, ped: "Pedigree", r: int):
        self._ped, self._r = ped, r

    ID = property(lambda s: int(s._ped.ids[s._r]))
    sex = property(lambda s: int(s._ped.sex[s._r]))
    rank = property(lambda s: s._r + 1)

    @property
    def father(self) -> Optional["Individual"]:
        f = self._ped.father[self._r]
        return None if f < 0 else Individual(self._ped, int(f))

    @property
    def mother(self) -> Optional["Individual"]:
        m = self._ped.mother[self._r]
        return None if m < 0 else Individual(self._ped, int(m))

    @property
    def children(self):
        return [Individual(self._ped, int(c)) for c in self._ped.children_of(self._r)]

    def __eq__(self, other):
        return isinstance(other, Individual) and other._ped is self._ped and other._r == self._r

    def __hash__(self):
        return hash((id(self._ped), self._r))

    def __repr__(self):  # src/create.jl:48-53
        f, m = self.father, self.mother
        return (f"ind: {self.ID}\nfather: {f.ID if f else 0}\n"
                f"mother: {m.ID if m else 0}\nsex: {self.sex}")


class Pedigree:
    """Rank-ordered pedigree (src/create.jl:60-74).  Iteration order is rank order."""

    def __init__(self, ids, father, mother, sex):
        self.ids = np.ascontiguousarray(ids, np.int64)
        self.father = np.ascontiguousarray(father, np.int32)   # 0-based rank or -1
        self.mother = np.ascontiguousarray(mother, np.int32)
        self.sex = np.ascontiguousarray(sex, np.int32)
        self._sorted_ids = None
        self._sorted_rank = None
        self._child_ptr = None
        self._child_idx = None

    # ---- Dict-like surface ----
    def __len__(self):
        return len(self.ids)

    def _index(self):
        if self._sorted_ids is None:
            order = np.argsort(self.ids, kind="stable")
            self._sorted_ids, self._sorted_rank = self.ids[order], order.astype(np.int64)
        return self._sorted_ids, self._sorted_rank

    def rank_of(self, IDs) -> np.ndarray:
        """0-based ranks of IDs; KeyError on an unknown ID (src/create.jl:70)."""
        IDs = np.atleast_1d(np.asarray(IDs, np.int64))
        sid, srank = self._index()
        if len(sid) == 0:
            if len(IDs):
                raise KeyError(int(IDs[0]))
            return np.zeros(0, np.int32)
        pos = np.minimum(np.searchsorted(sid, IDs), len(sid) - 1)
        bad = sid[pos] != IDs
        if bad.any():
            raise KeyError(int(IDs[np.argmax(bad)]))
        return srank[pos].astype(np.int32)

    def __contains__(self, ID):
        try:
            self.rank_of([ID])
            return True
        except KeyError:
            return False

    def __getitem__(self, ID) -> Individual:
        return Individual(self, int(self.rank_of([ID])[0]))

    def keys(self):
        return (int(i) for i in self.ids)

    def values(self):
        return (Individual(self, r) for r in range(len(self.ids)))

    __iter__ = keys

    # ---- children (src/create.jl:246-251) ----
    def _children(self):
        if self._child_ptr is None:
            n = len(self.ids)
            child = np.concatenate([np.nonzero(self.father >= 0)[0], np.nonzero(self.mother >= 0)[0]])
            parent = np.concatenate([self.father[self.father >= 0], self.mother[self.mother >= 0]])
            order = np.lexsort((child, parent))      # per parent, children in rank order
            self._child_idx = child[order].astype(np.int32)
            self._child_ptr = np.zeros(n + 1, np.int64)
            np.cumsum(np.bincount(parent, minlength=n), out=self._child_ptr[1:])
        return self._child_ptr, self._child_idx

    def children_of(self, r: int) -> np.ndarray:
        ptr, idx = self._children()
        return idx[ptr[r]:ptr[r + 1]]

    def n_children(self) -> np.ndarray:
        n = len(self.ids)
        return (np.bincount(self.father[self.father >= 0], minlength=n)
                + np.bincount(self.mother[self.mother >= 0], minlength=n))

    def depth(self) -> int:
        if getattr(self, "_depth", None) is None:
            self._depth = int(_max_depth(self.father, self.mother).max()) if len(self.ids) else 0
        return self._depth

    def __repr__(self):  # src/create.jl:76-111 (counts only)
        n = len(self.ids)
        rel = int((self.father >= 0).sum() + (self.mother >= 0).sum())
        men, women = int((self.sex == 1).sum()), int((self.sex == 2).sum())
        subjects = int((self.n_children() == 0).sum())
        d = self.depth()
        pl = lambda k: "" if k == 1 else "s"
        return (f"A pedigree with:\n{n} individual{pl(n)};\n{rel} parent-child relation{pl(rel)};\n"
                f"{men} m{'a' if men == 1 else 'e'}n;\n{women} wom{'a' if women == 1 else 'e'}n;\n"
                f"{subjects} subject{pl(subjects)};\n{d} generation{pl(d)}.")


def _max_depth(father: np.ndarray, mother: np.ndarray) -> np.ndarray:
    """src/create.jl:196-209 without recursion: founders 1, child = max(parents) + 1."""
    n = len(father)
    depth = np.ones(n, np.int32)
    hf, hm = father >= 0, mother >= 0
    for _ in range(n + 1):
        df = np.where(hf, depth[np.where(hf, father, 0)], 0)
        dm = np.where(hm, depth[np.where(hm, mother, 0)], 0)
        new = np.maximum(df, dm) + 1
        if np.array_equal(new, depth):
            return depth
        depth = new.astype(np.int32)
    raise ValueError("pedigree contains a cycle")


def _from_handle(h) -> Pedigree:
    from ._lib import check, lib, ptr
    try:
        n = lib().genlib_pedigree_n(h)
        ids = np.zeros(n, np.int64)
        father, mother, sex = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        check(lib().genlib_pedigree_arrays(h, ptr(ids), ptr(father), ptr(mother), ptr(sex)))
        depth = lib().genlib_pedigree_depth(h)
    finally:
        lib().genlib_pedigree_destroy(h)
    ped = Pedigree(ids, father, mother, sex)
    ped._depth = int(depth)
    return ped


def genealogy(source, sort: bool = True) -> Pedigree:
    """gen.genealogy: a path to a 4-column file (create.jl:161-189), a pandas DataFrame with
    columns ind/father/mother/sex (create.jl:131-146), or a mapping of such arrays.  Parsing,
    depth ordering and ranking run in the library's C++ loader (`genlib_genealogy_*`)."""
    import ctypes as C
    from ._lib import check, lib, ptr
    h = C.c_void_p()
    if isinstance(source, (str, bytes)) or hasattr(source, "__fspath__"):
        path = source if isinstance(source, bytes) else str(source).encode()
        check(lib().genlib_genealogy_csv(path, int(sort), C.byref(h)))
        return _from_handle(h)
    ind = np.ascontiguousarray(np.asarray(source["ind"]), np.int64)
    fid = np.ascontiguousarray(np.asarray(source["father"]), np.int64)
    mid = np.ascontiguousarray(np.asarray(source["mother"]), np.int64)
    sex = np.ascontiguousarray(np.asarray(source["sex"]), np.int32)
    if not (len(ind) == len(fid) == len(mid) == len(sex)):
        raise ValueError("ind, father, mother and sex must have the same length")
    check(lib().genlib_genealogy_arrays(len(ind), ptr(ind), ptr(fid), ptr(mid), ptr(sex), int(sort), C.byref(h)))
    return _from_handle(h)


def pro(pedigree: Pedigree) -> np.ndarray:
    """gen.pro: IDs of individuals without children, sorted (src/identify.jl:35-39)."""
    return np.sort(pedigree.ids[pedigree.n_children() == 0])


def founder(pedigree: Pedigree) -> np.ndarray:
    """gen.founder: IDs without any known parent, sorted (src/identify.jl:15-19)."""
    return np.sort(pedigree.ids[(pedigree.father < 0) & (pedigree.mother < 0)])
