"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this module (see oracle/genlib_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "genlib_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_genealogy_csv.restype = C.c_void_p
        L.oracle_genealogy_csv.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.oracle_genealogy_arrays.restype = C.c_void_p
        L.oracle_genealogy_arrays.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_ped_n.argtypes = [C.c_void_p]
        L.oracle_ped_arrays.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.oracle_rank_of.argtypes = [C.c_void_p, C.c_int64]
        L.oracle_pro.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_phi_pair.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_double)]
        L.oracle_phi.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                 C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.oracle_phi_ranks.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int,
                                       C.c_void_p, C.c_int]
        L.oracle_sparse_phi_ranks.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                              C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        L.oracle_row_updates.restype = C.c_int64
        L.oracle_row_updates.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_set_time_budget.argtypes = [C.c_double]
        L.oracle_phi_mean.restype = C.c_double
        L.oracle_phi_mean.argtypes = [C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OraclePedigree:
    """gen.genealogy as the oracle sees it (rank-ordered flat arrays)."""

    def __init__(self, handle):
        self._h = handle
        L = lib()
        n = L.oracle_ped_n(handle)
        self.n = n
        self.ids = np.zeros(n, np.int64)
        self.father = np.zeros(n, np.int32)
        self.mother = np.zeros(n, np.int32)
        self.sex = np.zeros(n, np.int32)
        L.oracle_ped_arrays(handle, _p(self.ids), _p(self.father), _p(self.mother), _p(self.sex))

    def __del__(self):
        try:
            lib().oracle_free(self._h)
        except Exception:
            pass

    @classmethod
    def from_csv(cls, path: str, sort: bool = True):
        st = C.c_int(0)
        h = lib().oracle_genealogy_csv(path.encode(), int(sort), C.byref(st))
        if not h:
            raise (KeyError if st.value in (1, 3) else OSError)(f"oracle_genealogy_csv status {st.value}")
        return cls(h)

    @classmethod
    def from_arrays(cls, ind, father, mother, sex=None, sort: bool = True):
        ind = np.ascontiguousarray(ind, np.int64)
        father = np.ascontiguousarray(father, np.int64)
        mother = np.ascontiguousarray(mother, np.int64)
        sex = None if sex is None else np.ascontiguousarray(sex, np.int32)
        st = C.c_int(0)
        h = lib().oracle_genealogy_arrays(len(ind), _p(ind), _p(father), _p(mother), _p(sex),
                                          int(sort), C.byref(st))
        if not h:
            raise KeyError(f"oracle_genealogy_arrays status {st.value}")
        return cls(h)

    def pro(self) -> np.ndarray:
        k = lib().oracle_pro(self._h, None)
        out = np.zeros(k, np.int64)
        lib().oracle_pro(self._h, _p(out))
        return out

    def phi_pair(self, a: int, b: int) -> float:
        v = C.c_double(0)
        if lib().oracle_phi_pair(self._h, a, b, C.byref(v)):
            raise KeyError((a, b))
        return v.value

    def phi(self, probands=None, nthreads: int = 0, max_steps: int = -1, with_steps: bool = False):
        """gen.phi(ped, probands) -> float32 matrix (oracle). `max_steps` bounds the run."""
        pro = self.pro() if probands is None else np.ascontiguousarray(probands, np.int64)
        cap = 4096
        steps = np.zeros((cap, 6), np.float64)
        nu = C.c_int(0)
        npro = len(pro)
        out = np.zeros((npro, npro), np.float32) if max_steps < 0 else None
        rc = lib().oracle_phi(self._h, npro, _p(pro), _p(out), C.byref(nu), nthreads, max_steps,
                              _p(steps), cap)
        if rc < 0:
            raise KeyError(f"oracle_phi status {rc}")
        res = None
        if out is not None:
            u = nu.value
            res = out.reshape(-1)[: u * u].reshape(u, u).copy()
        if with_steps:
            return res, steps[:rc].copy()
        return res


def phi_ranks(father, mother, pro_ranks, nthreads: int = 0, max_steps: int = -1):
    """Core on flat rank-indexed arrays (what the product's C ABI also takes)."""
    father = np.ascontiguousarray(father, np.int32)
    mother = np.ascontiguousarray(mother, np.int32)
    pro = np.ascontiguousarray(pro_ranks, np.int32)
    cap = 4096
    steps = np.zeros((cap, 6), np.float64)
    nu = C.c_int(0)
    out = np.zeros((len(pro), len(pro)), np.float32) if max_steps < 0 else None
    rc = lib().oracle_phi_ranks(len(father), _p(father), _p(mother), len(pro), _p(pro), _p(out),
                                C.byref(nu), nthreads, max_steps, _p(steps), cap)
    if rc < 0:
        raise KeyError(f"oracle_phi_ranks status {rc}")
    res = None
    if out is not None:
        u = nu.value
        res = out.reshape(-1)[: u * u].reshape(u, u).copy()
    return res, steps[:rc].copy()


def sparse_phi_ranks(father, mother, pro_ranks, ids=None, directed: bool = True, full: bool = False):
    """gen.sparse_phi on flat rank arrays, transliterated (src/compute.jl:321-447): the dense
    n_unique x n_unique Float32 matrix `getindex` would return and the number of stored entries of
    the `show` line.  `ids` (by rank) orders the founders in the queue like founder() does
    (identify.jl:15-19; None: rank order).  directed=False is the consistent variant that files every
    kinship where it is looked up.  full=True returns a dict instead of the count: stored, findable,
    misfiled, orphans, sum (all stored values, float64) and diag."""
    father = np.ascontiguousarray(father, np.int32)
    mother = np.ascontiguousarray(mother, np.int32)
    pro = np.ascontiguousarray(pro_ranks, np.int32)
    ids = None if ids is None else np.ascontiguousarray(ids, np.int64)
    nu = C.c_int(0)
    counts, sums = np.zeros(4, np.int64), np.zeros(2, np.float64)
    out = np.zeros((len(pro), len(pro)), np.float32)
    rc = lib().oracle_sparse_phi_ranks(len(father), _p(father), _p(mother), _p(ids), len(pro), _p(pro),
                                       int(directed), _p(out), C.byref(nu), _p(counts), _p(sums))
    if rc < 0:
        raise KeyError(f"oracle_sparse_phi_ranks status {rc}")
    u = nu.value
    dense = out.reshape(-1)[: u * u].reshape(u, u).copy()
    if full:
        return dense, {"stored": int(counts[0]), "findable": int(counts[1]), "misfiled": int(counts[2]),
                       "orphans": int(counts[3]), "sum": float(sums[0]), "diag": float(sums[1])}
    return dense, int(counts[0])


def bounded_steps(father, mother, pro_ranks, seconds: float, nthreads: int = 0):
    """Run the reference algorithm step by step until `seconds` have elapsed (the step in
    flight finishes); returns the per-step info of the steps that ran.  CPU baseline only."""
    father = np.ascontiguousarray(father, np.int32)
    mother = np.ascontiguousarray(mother, np.int32)
    pro = np.ascontiguousarray(pro_ranks, np.int32)
    cap = 4096
    steps = np.zeros((cap, 6), np.float64)
    nu = C.c_int(0)
    lib().oracle_set_time_budget(float(seconds))
    try:
        rc = lib().oracle_phi_ranks(len(father), _p(father), _p(mother), len(pro), _p(pro), None,
                                    C.byref(nu), nthreads, -1, _p(steps), cap)
    finally:
        lib().oracle_set_time_budget(0.0)
    if rc < 0:
        raise KeyError(f"oracle_phi_ranks status {rc}")
    return steps[:rc].copy()


def row_updates(father, mother, pro_ranks) -> int:
    """Individuals born over a complete gen.phi call (the benchmark's unit of work)."""
    father = np.ascontiguousarray(father, np.int32)
    mother = np.ascontiguousarray(mother, np.int32)
    pro = np.ascontiguousarray(pro_ranks, np.int32)
    return int(lib().oracle_row_updates(len(father), _p(father), _p(mother), len(pro), _p(pro)))


def phi_mean(phi: np.ndarray) -> float:
    phi = np.ascontiguousarray(phi, np.float32)
    return lib().oracle_phi_mean(_p(phi), phi.shape[0])


def num_threads() -> int:
    return lib().oracle_num_threads()


# ---- benchmark workloads without the product library ---------------------------------------------
def _synth():
    """genlib.jl_b200/synth.py loaded by path: pure NumPy, does not import the package and so never
    maps libgenlib_cuda.so into a process that only runs the oracle (bench.py --impl reference)."""
    import importlib.util
    import sys
    name = "_genlib_synth_for_oracle"
    if name not in sys.modules:
        path = os.path.join(os.path.dirname(_HERE), "genlib.jl_b200", "synth.py")
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return sys.modules[name]


def workload(name: str, scale: float = 1.0):
    """The named benchmark pedigree through the ORACLE's loader (create.jl:131-254 restated):
    (father ranks, mother ranks, proband ranks in output order, description)."""
    if name == "genea140":
        ped = OraclePedigree.from_csv(os.path.join(os.path.dirname(_HERE), "tests", "data", "genea140.csv"))
        pro = ped.pro()
        desc = "genea140 (41523 individuals, 140 probands)"
    else:
        s = _synth().config(name, scale)
        ped = OraclePedigree.from_arrays(s.ind, s.father, s.mother, s.sex)
        pro = s.probands
        p = s.params
        desc = (f"{name} synthetic: {p['n_individuals']} individuals, {p['generations']} generations, "
                f"{p['n_probands']} probands, alpha={p['alpha']}, demes={p['demes']}, migration={p['migration']}, "
                f"overlap={p['overlap']}, seed={p['seed']}" + (f", scale={scale}" if scale != 1 else ""))
    ranks = np.array([lib().oracle_rank_of(ped._h, int(i)) for i in pro], np.int32)
    if (ranks < 0).any():
        raise KeyError("unknown proband ID")
    return ped.father, ped.mother, ranks, desc


def matrix_sha256(m: np.ndarray) -> str:
    """sha256 of the raw Float32 buffer of a (symmetric) kinship matrix."""
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(m, np.float32).tobytes()).hexdigest()
