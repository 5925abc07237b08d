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
        L.oracle_sparse_phi_ranks.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                              C.POINTER(C.c_int), C.POINTER(C.c_int64)]
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


def sparse_phi_ranks(father, mother, pro_ranks):
    """gen.sparse_phi on flat rank arrays: (dense n_unique x n_unique Float32 values of the
    KinshipMatrix, number of stored entries).  src/compute.jl:321-447."""
    father = np.ascontiguousarray(father, np.int32)
    mother = np.ascontiguousarray(mother, np.int32)
    pro = np.ascontiguousarray(pro_ranks, np.int32)
    nu, stored = C.c_int(0), C.c_int64(0)
    out = np.zeros((len(pro), len(pro)), np.float32)
    rc = lib().oracle_sparse_phi_ranks(len(father), _p(father), _p(mother), len(pro), _p(pro), _p(out),
                                       C.byref(nu), C.byref(stored))
    if rc < 0:
        raise KeyError(f"oracle_sparse_phi_ranks status {rc}")
    u = nu.value
    return out.reshape(-1)[: u * u].reshape(u, u).copy(), int(stored.value)


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


def phi_mean(phi: np.ndarray) -> float:
    phi = np.ascontiguousarray(phi, np.float32)
    return lib().oracle_phi_mean(_p(phi), phi.shape[0])


def num_threads() -> int:
    return lib().oracle_num_threads()
