"""CPU replay of the engine's schedule with NumPy -- TEST INFRASTRUCTURE ONLY.

It executes the plan exported by `genlib_plan_layer_arrays` with exactly the
semantics the CUDA kernels implement (cross block, transposed fp64 scratch,
couple-compressed intra block, in-place slot recycling), on a matrix that is
NaN-poisoned wherever nothing has been written.  Agreement with the oracle
validates the planner and the kernel DESIGN without a GPU; the kernels
themselves are checked by the `-m gpu` tests.
"""
from __future__ import annotations

import numpy as np


def replay(plan, numerics: str = "reference", check_poison: bool = True, on_layer=None) -> np.ndarray:
    T = np.float32 if numerics == "reference" else np.float64
    W = int(plan.capacity)
    A = np.full((W, W), np.nan, T)
    for t in range(plan.n_layers):
        info = plan.layer_info(t)
        arr = plan.layer_arrays(t)
        n, nf = info["n_new"], info["n_fam"]
        if n == 0:
            continue
        slot, fam, ind = arr["member_slot"], arr["member_fam"], arr["member_ind"]
        pf, pm = arr["fam_father_slot"], arr["fam_mother_slot"]
        live = np.nonzero(arr["live_flags"] & 1)[0]
        carried = np.nonzero(arr["live_flags"] & 2)[0]
        assert len(live) == info["live_before"] and len(carried) == info["carried"]
        assert not np.intersect1d(slot, live).size, "new slots must not alias live rows"
        # ---- cross: R[F, p] over live columns (fp64, one rounding) ----
        Al = A[:, live].astype(np.float64)                 # columns restricted to live
        zero = np.zeros((1, len(live)))
        rows_f = np.where(pf[:, None] >= 0, Al[np.maximum(pf, 0)], zero)
        rows_m = np.where(pm[:, None] >= 0, Al[np.maximum(pm, 0)], zero)
        R = 0.5 * rows_f + 0.5 * rows_m                    # (nf, |live|)
        if check_poison and len(live):
            assert not np.isnan(R).any(), f"layer {t}: cross block read an unwritten entry"
        # Rt indexed by slot for the intra gather
        Rt = np.full((W, nf), np.nan)
        Rt[live] = R.T
        # rows/columns new x carried, rounded once
        if len(carried):
            pos = np.searchsorted(live, carried)
            blk = R[fam][:, pos].astype(T)                 # (n, |carried|)
            A[np.ix_(slot, carried)] = blk
            A[np.ix_(carried, slot)] = blk.T
        # ---- intra: V[F, G] = 1/2 Rt[pf F, G] + 1/2 Rt[pm F, G] ----
        zf = np.zeros((1, nf))
        V = 0.5 * np.where(pf[:, None] >= 0, Rt[np.maximum(pf, 0)], zf) \
            + 0.5 * np.where(pm[:, None] >= 0, Rt[np.maximum(pm, 0)], zf)   # (nf, nf)
        if check_poison:
            assert not np.isnan(V).any(), f"layer {t}: intra block read an unwritten entry"
        Vab = V[np.ix_(fam, fam)]                          # row member climbed first
        hi = ind[:, None] > ind[None, :]
        blk = np.where(hi, Vab, Vab.T)
        both = (pf >= 0) & (pm >= 0)
        dv = np.full(nf, 0.5)
        if both.any():
            dv[both] = 0.5 + 0.5 * A[pf[both], pm[both]].astype(np.float64)
        blk[np.arange(n), np.arange(n)] = dv[fam]
        A[np.ix_(slot, slot)] = blk.astype(T)
        if on_layer is not None:
            on_layer(t, A, np.union1d(carried, slot))      # frontier after the layer, live slots
    ps = plan.proband_slots()
    out = A[np.ix_(ps, ps)]
    if check_poison:
        assert not np.isnan(out).any()
    return out
