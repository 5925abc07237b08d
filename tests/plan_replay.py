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


def replay_sharded(plan, numerics: str = "reference", exchange=None) -> np.ndarray:
    """Replay of the ROW-SHARDED schedule: every rank holds only the full-width rows of the
    individuals it owns (A[g] is rows_cap[g] x W, NaN-poisoned).  Mirrors what the multi-GPU
    kernels do: parent rows are read wherever they live, the couple matrix is computed in column
    blocks (own couples) and pushed to the owner of each row couple, every rank expands its own
    rows.  All ranks are simulated in this process; `exchange`, if given, is called as
    exchange(layer, rank_from, rank_to, nbytes) for every remote access (traffic accounting).
    """
    T = np.float32 if numerics == "reference" else np.float64
    W, G = int(plan.capacity), plan.world
    A = [np.full((max(plan.rank_rows(g), 1), W), np.nan, T) for g in range(G)]
    es = np.dtype(T).itemsize

    def note(layer, src, dst, nbytes):
        if exchange is not None and src != dst and nbytes:
            exchange(layer, int(src), int(dst), int(nbytes))

    for t in range(plan.n_layers):
        info, arr, sh = plan.layer_info(t), plan.layer_arrays(t), plan.layer_shard(t)
        n, nf = info["n_new"], info["n_fam"]
        if n == 0:
            continue
        slot, fam, ind, lrow = arr["member_slot"], arr["member_fam"], arr["member_ind"], sh["member_lrow"]
        pf, pm = arr["fam_father_slot"], arr["fam_mother_slot"]
        live = np.nonzero(arr["live_flags"] & 1)[0]
        carried = np.nonzero(arr["live_flags"] & 2)[0]
        fb, mb = sh["fam_base"], sh["mem_base"]
        assert fb[0] == 0 and fb[-1] == nf and mb[-1] == n and np.all(np.diff(fam) >= 0)
        pos_in_live = np.full(W, -1)
        pos_in_live[live] = np.arange(len(live))

        def row_of(owner, lr, cols, t=t, reader=None):
            """Row (owner, lr) restricted to `cols` (fp64)."""
            if reader is not None:
                note(t, owner, reader, len(cols) * es)
            return A[owner][lr, cols].astype(np.float64)

        # ---- cross: each rank, own couples ----
        Rt = [None] * G
        for g in range(G):
            F0, F1 = fb[g], fb[g + 1]
            R = np.zeros((F1 - F0, len(live)))
            for F in range(F0, F1):
                acc = np.zeros(len(live))
                for own, lr in ((sh["fam_father_owner"][F], sh["fam_father_lrow"][F]),
                                (sh["fam_mother_owner"][F], sh["fam_mother_lrow"][F])):
                    if own >= 0:
                        acc = acc + 0.5 * row_of(own, lr, live, reader=g) if False else acc
                # grouping must match: 0.5*father + 0.5*mother with one rounding
                f_ok, m_ok = sh["fam_father_owner"][F] >= 0, sh["fam_mother_owner"][F] >= 0
                x = row_of(sh["fam_father_owner"][F], sh["fam_father_lrow"][F], live, reader=g) if f_ok else 0.0
                y = row_of(sh["fam_mother_owner"][F], sh["fam_mother_lrow"][F], live, reader=g) if m_ok else 0.0
                R[F - F0] = 0.5 * x + 0.5 * y
            if len(live):
                assert not np.isnan(R).any(), f"layer {t} rank {g}: cross block read an unwritten entry"
            Rt[g] = R                                        # [own couple][live position]
        # rows/columns new x carried (all cross blocks are complete before anything is written)
        writes = []
        for g in range(G):
            M0, M1 = mb[g], mb[g + 1]
            if len(carried) and M1 > M0:
                cpos = pos_in_live[carried]
                blk = Rt[g][fam[M0:M1] - fb[g]][:, cpos].astype(T)          # (own members, carried)
                writes.append((g, lrow[M0:M1], carried, blk))
                for k, c in enumerate(carried):                              # mirror into the carried rows
                    co, cl = sh["live_owner"][c], sh["live_lrow"][c]
                    writes.append((co, np.array([cl]), slot[M0:M1], blk[:, k][None, :]))
                    note(t, g, co, (M1 - M0) * es)
        for g, rows_, cols_, blk in writes:
            A[g][np.ix_(rows_, cols_)] = blk
        # ---- couple: rank g computes V[F, G] for all F and its own G; pushes rows to owner(F) ----
        Vrow = [np.full((fb[g + 1] - fb[g], nf), np.nan) for g in range(G)]   # V[F own, G]
        Vt = [np.full((fb[g + 1] - fb[g], nf), np.nan) for g in range(G)]     # V[G, F own] stored [F own][G]
        owner_of_fam = np.repeat(np.arange(G), np.diff(fb))
        for g in range(G):
            G0, G1 = fb[g], fb[g + 1]
            if G1 == G0:
                continue
            Rg = Rt[g]                                       # [own couple G][live pos]
            zero = np.zeros(G1 - G0)
            for F in range(nf):
                a = Rg[:, pos_in_live[pf[F]]] if pf[F] >= 0 else zero
                b = Rg[:, pos_in_live[pm[F]]] if pm[F] >= 0 else zero
                v = 0.5 * a + 0.5 * b                        # V[F, G0:G1]
                o = owner_of_fam[F]
                Vrow[o][F - fb[o], G0:G1] = v
                Vt[g][:, F] = v
                note(t, g, o, (G1 - G0) * es)
        # ---- expand: every rank its own rows ----
        for g in range(G):
            M0, M1 = mb[g], mb[g + 1]
            if M1 == M0:
                continue
            fl = fam[M0:M1] - fb[g]
            a = Vrow[g][fl][:, fam]                          # V[F_i, G_j]
            b = Vt[g][fl][:, fam]                            # V[G_j, F_i]
            hi = ind[M0:M1, None] > ind[None, :]
            blk = np.where(hi, a, b)
            for q in range(M0, M1):                          # diagonal: 1/2 + 1/2 Psi[father, mother]
                F = fam[q]
                d = 0.5
                if pf[F] >= 0 and pm[F] >= 0:
                    d = 0.5 + 0.5 * float(row_of(sh["fam_father_owner"][F], sh["fam_father_lrow"][F], [pm[F]], reader=g)[0])
                blk[q - M0, q] = d
            assert not np.isnan(blk).any(), f"layer {t} rank {g}: expand read an unwritten couple entry"
            A[g][np.ix_(lrow[M0:M1], slot)] = blk.astype(T)
    ps = plan.proband_slots()
    po, pl = plan.proband_rows()
    out = np.stack([A[o][l, ps] for o, l in zip(po, pl)]) if len(ps) else np.zeros((0, 0), T)
    assert not np.isnan(out).any()
    return out
