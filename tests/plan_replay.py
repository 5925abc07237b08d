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
        sparse = getattr(plan, "schedule", "phi").startswith("sparse_phi")
        if sparse:      # `x / 2` is a Float32 division of a stored Float32 (compute.jl:350-389), the sum Float64
            R = (rows_f.astype(np.float32) * np.float32(0.5)).astype(np.float64) \
                + (rows_m.astype(np.float32) * np.float32(0.5)).astype(np.float64)
        else:
            R = 0.5 * rows_f + 0.5 * rows_m                # (nf, |live|)
        if check_poison and len(live):
            assert not np.isnan(R).any(), f"layer {t}: cross block read an unwritten entry"
        # Rt indexed by slot for the intra gather; the sparse_phi schedule reads the STORED (Float32)
        # cross values there (compute.jl:331, 363-395), phi the unrounded ones (compute.jl:107, 296)
        Rt = np.full((W, nf), np.nan)
        Rt[live] = R.T.astype(np.float32).astype(np.float64) if sparse else R.T
        # rows/columns new x carried, rounded once
        if len(carried):
            pos = np.searchsorted(live, carried)
            blk = R[fam][:, pos].astype(T)                 # (n, |carried|)
            A[np.ix_(slot, carried)] = blk
            A[np.ix_(carried, slot)] = blk.T
        # ---- intra: V[F, G] = 1/2 Rt[pf F, G] + 1/2 Rt[pm F, G] ----
        zf = np.zeros((1, nf))
        Vf, Vm = np.where(pf[:, None] >= 0, Rt[np.maximum(pf, 0)], zf), np.where(pm[:, None] >= 0, Rt[np.maximum(pm, 0)], zf)
        if sparse:
            V = (Vf.astype(np.float32) * np.float32(0.5)).astype(np.float64) \
                + (Vm.astype(np.float32) * np.float32(0.5)).astype(np.float64)
        else:
            V = 0.5 * Vf + 0.5 * Vm                         # (nf, nf)
        if check_poison:
            assert not np.isnan(V).any(), f"layer {t}: intra block read an unwritten entry"
        Vab = V[np.ix_(fam, fam)]                          # row member climbed first
        hi = ind[:, None] > ind[None, :]
        blk = np.where(hi, Vab, Vab.T)
        both = (pf >= 0) & (pm >= 0)
        dv = np.full(nf, 0.5)
        if both.any():
            dv[both] = 0.5 + ((A[pf[both], pm[both]].astype(np.float32) * np.float32(0.5)).astype(np.float64) if sparse
                              else 0.5 * A[pf[both], pm[both]].astype(np.float64))
        if getattr(plan, "schedule", "phi") == "sparse_phi":
            # the reference files phi[earlier][later] and reads phi[lower rank][higher rank]
            # (compute.jl:393 vs :367-389): a pair whose queue order inverts its rank order reads as 0
            rk = arr["member_rank"]
            blk = np.where(hi == (rk[:, None] > rk[None, :]), blk, 0.0)
        blk[np.arange(n), np.arange(n)] = dv[fam]
        A[np.ix_(slot, slot)] = blk.astype(T)
        if on_layer is not None:
            on_layer(t, A, np.union1d(carried, slot))      # frontier after the layer, live slots
    ps = plan.proband_slots()
    out = A[np.ix_(ps, ps)]
    if check_poison:
        assert not np.isnan(out).any()
    return out


class ShardedReplay:
    """The row-sharded schedule as the layer kernel (csrc/layer_kernel.cuh) executes it, rank by rank.

    `A[g]` is rank g's rows_cap[g] x W frontier block, NaN where nothing was written.  One layer on
    rank g (`step`): the parent rows of g's own couples are read over the live columns (a remote
    row is a peer read) -- that is the strip buffer Q[p][F] = (Psi[f_F, p], Psi[m_F, p]); from it
    come the members' rows against the carried columns, the members' columns in the carried rows
    (a peer store when the carried row lives elsewhere: `writes`, applied by `apply_writes`) and,
    for EVERY couple G of the layer, both groupings of the four entries of (F, G), of which the
    ranks pick one per member pair.  Every rank writes only rows it owns, except for the mirror."""

    def __init__(self, plan, numerics="reference", exchange=None):
        self.plan, self.T = plan, (np.float32 if numerics == "reference" else np.float64)
        self.W, self.G = int(plan.capacity), plan.world
        self.A = [np.full((max(plan.rank_rows(g), 1), self.W), np.nan, self.T) for g in range(self.G)]
        self.es = np.dtype(self.T).itemsize
        self.exchange = exchange
        self.sparse = getattr(plan, "schedule", "phi").startswith("sparse_phi")

    def half_sum(self, x, y):
        """phi: 1/2 x + 1/2 y in Float64.  sparse_phi: the halves are Float32 divisions of STORED
        Float32 values (compute.jl:350-389), the sum is Float64."""
        if self.sparse:
            h = np.float32(0.5)
            return (np.asarray(x, np.float64).astype(np.float32) * h).astype(np.float64) \
                + (np.asarray(y, np.float64).astype(np.float32) * h).astype(np.float64)
        return 0.5 * np.asarray(x, np.float64) + 0.5 * np.asarray(y, np.float64)

    def inner(self, x, y):
        """An intermediate kinship: unrounded Float64 in phi, a stored Float32 in sparse_phi."""
        v = self.half_sum(x, y)
        return v.astype(np.float32).astype(np.float64) if self.sparse else v

    def note(self, src, dst, nbytes):
        if self.exchange is not None and src != dst and nbytes:
            self.exchange(self.t, int(src), int(dst), int(nbytes))

    def begin(self, t):
        plan = self.plan
        self.t = t
        self.info, self.arr, self.sh = plan.layer_info(t), plan.layer_arrays(t), plan.layer_shard(t)
        a, sh = self.arr, self.sh
        self.n, self.nf = self.info["n_new"], self.info["n_fam"]
        self.live = np.nonzero(a["live_flags"] & 1)[0]
        self.carried = np.nonzero(a["live_flags"] & 2)[0]
        self.pos_in_live = np.full(self.W + 1, len(self.live))          # slot -> column of the strip buffer (-1 -> zero column)
        self.pos_in_live[self.live] = np.arange(len(self.live))
        fb, mb = sh["fam_base"], sh["mem_base"]
        assert fb[0] == 0 and fb[-1] == self.nf and mb[-1] == self.n and np.all(np.diff(a["member_fam"]) >= 0)
        self.writes = []
        return self.n > 0

    def row(self, owner, lr, cols, reader):
        self.note(owner, reader, len(cols) * self.es)
        return self.A[owner][lr, cols].astype(np.float64)

    def step(self, g):
        a, sh, live = self.arr, self.sh, self.live
        F0, F1 = sh["fam_base"][g], sh["fam_base"][g + 1]
        M0, M1 = sh["mem_base"][g], sh["mem_base"][g + 1]
        if F1 == F0:
            return
        # ---- producer: Q[p][F] = (father row, mother row) of the own couples, raw stored values ----
        X, Y = np.zeros((F1 - F0, len(live) + 1)), np.zeros((F1 - F0, len(live) + 1))   # last column: "no such parent"
        for F in range(F0, F1):
            fo, mo = sh["fam_father_owner"][F], sh["fam_mother_owner"][F]
            if fo >= 0:
                X[F - F0, :-1] = self.row(fo, sh["fam_father_lrow"][F], live, g)
            if mo >= 0:
                Y[F - F0, :-1] = self.row(mo, sh["fam_mother_lrow"][F], live, g)
        if len(live):
            assert not np.isnan(X).any() and not np.isnan(Y).any(), f"layer {self.t} rank {g}: a parent row has an unwritten live entry"
        fam, ind = a["member_fam"], a["member_ind"]
        fl = fam[M0:M1] - F0
        if len(self.carried) and M1 > M0:                       # new x carried, rounded once; and its mirror
            cpos = self.pos_in_live[self.carried]
            blk = self.half_sum(X[:, cpos], Y[:, cpos])[fl].astype(self.T)
            self.A[g][np.ix_(sh["member_lrow"][M0:M1], self.carried)] = blk
            for k, c in enumerate(self.carried):                # a peer store when the carried row lives elsewhere
                co, cl = sh["live_owner"][c], sh["live_lrow"][c]
                self.writes.append((co, np.array([cl]), a["member_slot"][M0:M1], blk[:, k][None, :]))
                self.note(g, co, (M1 - M0) * self.es)
        # ---- consumer: for every couple G of the layer, both groupings of the four entries of (F, G) ----
        pf, pm = a["fam_father_slot"], a["fam_mother_slot"]
        gf, gm = self.pos_in_live[pf], self.pos_in_live[pm]     # slot -1 -> the zero column
        ax, ay, cx, cy = X[:, gf], Y[:, gf], X[:, gm], Y[:, gm]
        Vf = self.half_sum(self.inner(ax, cx), self.inner(ay, cy)).astype(self.T)   # the strip couple's member is climbed first
        Vg = self.half_sum(self.inner(ax, ay), self.inner(cx, cy)).astype(self.T)   # the tile couple's member is climbed first
        if M1 == M0:
            return
        hi = ind[M0:M1, None] > ind[None, :]
        blk = np.where(hi, Vf[fl][:, fam], Vg[fl][:, fam])
        if getattr(self.plan, "schedule", "phi") == "sparse_phi":       # misfiled kinships read as 0 (compute.jl:393)
            rk = a["member_rank"]
            blk = np.where(hi == (rk[M0:M1, None] > rk[None, :]), blk, 0)
        for q in range(M0, M1):                                 # diagonal: 1/2 + 1/2 Psi[father, mother]
            F, d = fam[q], 0.5
            if pf[F] >= 0 and pm[F] >= 0:
                d = float(self.half_sum(self.row(sh["fam_father_owner"][F], sh["fam_father_lrow"][F], [pm[F]], g)[0], 1.0))
            blk[q - M0, q] = self.T(d)
        self.A[g][np.ix_(sh["member_lrow"][M0:M1], a["member_slot"])] = blk

    def apply_writes(self, only=None):
        for g, rows_, cols_, blk in self.writes:
            if only is None or g in only:
                self.A[g][np.ix_(rows_, cols_)] = blk

    def result_rows(self, g):
        ps = self.plan.proband_slots()
        po, pl = self.plan.proband_rows()
        idx = np.nonzero(po == g)[0]
        return idx, (np.stack([self.A[g][pl[u], ps] for u in idx]) if len(idx) else np.zeros((0, len(ps)), self.T))


def replay_sharded(plan, numerics: str = "reference", exchange=None) -> np.ndarray:
    """All ranks of the row-sharded schedule simulated in this process; `exchange(layer, src,
    dst, nbytes)` is called for every remote access (traffic accounting)."""
    R = ShardedReplay(plan, numerics, exchange)
    for t in range(plan.n_layers):
        if not R.begin(t):
            continue
        for g in range(R.G):
            R.step(g)
        R.apply_writes()
    n = plan.n_unique
    out = np.zeros((n, n), R.T)
    for g in range(R.G):
        idx, rows = R.result_rows(g)
        out[idx] = rows
    assert not np.isnan(out).any()
    return out
