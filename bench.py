#!/usr/bin/env python
"""bench.py -- gen.phi on the named synthetic pedigree: kinship row-updates/s.

    python bench.py --gpus N --steps K --warmup W            (N = 1: this process;
    N > 1: one rank per GPU under torchrun, RANK/LOCAL_RANK/WORLD_SIZE from the env)
    python bench.py --impl reference ...                     (CPU restatement of the
    reference algorithm on the host cores, same metric and config)

A "step" is one complete pass of the hot path -- every generation layer of
gen.phi(ped, probands) -- over one synthetic pedigree.  `value` counts the pass
with the schedule already resident in HBM (CUDA events on the engine's
stream); `e2e` is the public call with HOST buffers: planning, H2D of the
schedule, all layers, proband gather and D2H into pinned host memory.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gen.phi kinship row-updates/s"
UNIT = "row-updates/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_workload(args):
    import genlib_b200 as gen
    if args.workload == "genea140":
        ped = gen.genealogy(gen.genea140)
        pro = gen.pro(ped)
        desc = "genea140 (41523 individuals, 140 probands)"
    else:
        s = gen.synth.config(args.workload, args.scale)
        ped = gen.genealogy(s.as_columns())
        pro = s.probands
        p = s.params
        desc = (f"{args.workload} synthetic: {p['n_individuals']} individuals, {p['generations']} generations, "
                f"{p['n_probands']} probands, alpha={p['alpha']}, demes={p['demes']}, migration={p['migration']}, "
                f"overlap={p['overlap']}, seed={p['seed']}" + (f", scale={args.scale}" if args.scale != 1 else ""))
    return gen, ped, ped.rank_of(pro), desc


def cpu_baseline(ped, ranks, seconds):
    """The oracle (C restatement of src/compute.jl:233-304, all host threads) on a bounded
    sample: the first generation steps of the SAME workload until `seconds` have elapsed."""
    from oracle import binding as ob
    cores = ob.num_threads()
    t0 = time.time()
    steps = ob.bounded_steps(ped.father, ped.mother, ranks, seconds, nthreads=cores)
    wall = time.time() - t0
    secs = float(steps[:, 5].sum()) if len(steps) else 0.0
    rows = float(steps[:, 4].sum()) if len(steps) else 0.0
    pairs = float(steps[:, 3].sum()) if len(steps) else 0.0
    return {"value": rows / secs if secs > 0 else 0.0, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"first {len(steps)} generation steps ({int(rows)} row-updates, {pairs:.3g} pair "
                       f"evaluations) of the same pedigree, {secs:.1f} s in the pair loops, {wall:.1f} s wall; "
                       "C restatement of the reference algorithm (Julia is not installable here)"),
            "seconds": secs}


def run_reference(args, rank, world):
    if rank != 0:
        return
    gen, ped, ranks, desc = build_workload(args)
    # each "step" of this arm is one bounded sample; W warm-up + K timed samples
    per = max(1.0, args.cpu_seconds / max(1, args.steps))
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(ped, ranks, min(per, 2.0))
    vals, last = [], None
    t0 = time.time()
    for _ in range(args.steps):
        last = cpu_baseline(ped, ranks, per)
        vals.append(last["value"])
    ms = (time.time() - t0) * 1e3 / max(1, args.steps)
    v = float(np.mean(vals))
    last["value"] = v
    emit(({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": desc}, "cpu_baseline": last,
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL, torchrun banners) print to fd 1; the contract is ONE JSON line on stdout.
    Everything written to fd 1 from here on goes to stderr; emit() writes the line to the real one."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C3", "C4", "C5", "genea140"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--numerics", default="reference", choices=["reference", "fp64"])
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU baseline budget (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--layers-json", default="", help="write per-layer timings here")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world > 1:
        from bench_dist import main_dist          # sharded path (one rank per GPU)
        return main_dist(args, rank, world, local, sys.modules[__name__])

    gen, ped, ranks, desc = build_workload(args)
    if gen.lib().genlib_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    esize = 4 if args.numerics == "reference" else 8
    t0 = time.time()
    plan = gen.Plan(ped.father, ped.mother, ranks)
    eng = gen.Engine(plan, numerics=args.numerics, device=local)
    setup_s = time.time() - t0
    rows = plan.row_updates
    for _ in range(max(args.warmup, 3)):
        eng.run()
    K = args.steps
    cross_ms, couple_ms, expand_ms, step_ms = [], [], [], []
    with ClockSampler(local) as clocks:
        time.sleep(0.6)                      # let nvidia-smi start sampling
        n0 = len(clocks.rows)
        while len(clocks.rows) == n0 and clocks.proc and time.time() < t0 + 60:
            eng.run()                        # keep the GPU under load until the first sample lands
        clocks.rows.clear()
        t0 = time.time()
        for _ in range(K):
            step_ms.append(eng.run(time_layers=True))
            infos = [eng.layer_info(t) for t in range(plan.n_layers)]
            cross_ms.append([i["ms_cross"] for i in infos])
            couple_ms.append([i["ms_couple"] for i in infos])
            expand_ms.append([i["ms_expand"] for i in infos])
        wall_ms = (time.time() - t0) * 1e3
    stats = eng.stats()
    total_ms = float(np.sum(step_ms))
    value = rows * K / (total_ms * 1e-3)
    infos = plan.layers()
    cross_ms, couple_ms, expand_ms = np.array(cross_ms), np.array(couple_ms), np.array(expand_ms)
    intra_ms = couple_ms + expand_ms
    # dominant kernel: cross_kernel.  Algorithmic bytes per launch = s * 4 n L.
    cross_bytes = np.array([esize * 4.0 * i["n_new"] * i["live_before"] for i in infos])
    intra_bytes = np.array([esize * 3.0 * i["n_new"] ** 2 for i in infos])
    launched = cross_bytes > 0
    peak, peak_src = measured_peaks()
    c_t = cross_ms[:, launched].sum() * 1e-3
    i_t = intra_ms.sum() * 1e-3
    achieved = cross_bytes[launched].sum() * K / c_t / 1e9 if c_t > 0 else 0.0
    whole = (cross_bytes.sum() + intra_bytes.sum()) * K / (total_ms * 1e-3) / 1e9
    # DRAM traffic per launch from `ncu --set full` (profiles/r01/ncu_full_c3_final_summary.json): a
    # 39208 x 39114 layer moved 6.166 GB read + 6.111 GB written for 24.54 GB of algorithmic bytes
    # (couples halve the rows that are read and written); scaled to the average launch.
    traffic_ratio = (6.165890e9 + 6.110764e9) / (4 * 4.0 * 39208 * 39114) if args.numerics == "reference" else None
    roofline = {"bound": "hbm", "kernel": "cross_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": (traffic_ratio * float(cross_bytes[launched].mean())) if (traffic_ratio and launched.any()) else None,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch, scaled by algorithmic bytes",
                "achieved_dram": (traffic_ratio * achieved) if traffic_ratio else None,      # GB/s of measured DRAM bytes
                "dram_frac": (traffic_ratio * achieved / peak) if traffic_ratio else None,   # measured DRAM bytes / time / peak
                "peak_source": peak_src,
                "launches": int(launched.sum()) * K,
                "avg_launch_ms": float(cross_ms[:, launched].mean()) if launched.any() else 0.0,
                "alg_bytes_per_launch": float(cross_bytes[launched].mean()) if launched.any() else 0.0,
                "share_of_step": c_t / (total_ms * 1e-3),
                "intra_kernels": {"achieved": intra_bytes.sum() * K / i_t / 1e9 if i_t > 0 else 0.0,
                                  "share_of_step": i_t / (total_ms * 1e-3),
                                  "couple_share": float(couple_ms.sum() / total_ms),
                                  "expand_share": float(expand_ms.sum() / total_ms)},
                "whole_step": {"achieved": whole, "frac": whole / peak,
                               "frac_of_8TBs_nominal": whole / 8000.0}}
    if args.layers_json:
        with open(args.layers_json, "w") as fh:
            json.dump([{**i, "ms_cross": float(cross_ms[:, t].mean()), "ms_couple": float(couple_ms[:, t].mean()),
                        "ms_expand": float(expand_ms[:, t].mean())} for t, i in enumerate(infos)], fh, indent=1)
    eng.close()

    # ---- end to end: the public call with host buffers, pinned output ----
    n = plan.n_unique
    pinned = gen.PinnedMatrix(n, np.float32)
    e2e_t, h2d, d2h = [], 0, 0
    for it in range(args.e2e_steps + 1):
        t0 = time.time()
        _, st = gen.phi_arrays(ped.father, ped.mother, ranks, numerics=args.numerics, device=local,
                               out=pinned.array)
        dt = time.time() - t0
        if it > 0 or args.e2e_steps == 0:
            e2e_t.append(dt)
        h2d = st["h2d_bytes"] + ped.father.nbytes + ped.mother.nbytes + ranks.nbytes
        d2h = st["d2h_bytes"]
    e2e = {"value": rows / float(np.mean(e2e_t)), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_call": float(np.mean(e2e_t)) * 1e3,
           "breakdown_ms": {k: st[k] for k in ("ms_plan", "ms_upload", "ms_kernels", "ms_fetch")}}
    checksum = float(pinned.array.astype(np.float64).sum())
    pinned.free()

    base = cpu_baseline(ped, ranks, args.cpu_seconds) if args.cpu_seconds > 0 else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic" if args.workload != "genea140" else "genea140.csv",
            "config": {"workload": desc, "numerics": args.numerics, "storage_bytes": esize,
                       "row_updates_per_step": int(rows), "layers": plan.n_layers, "capacity_slots": int(plan.capacity),
                       "device_bytes": int(stats["device_bytes"]), "l2": "working set >> L2 (no flush needed)",
                       "alg_bytes_per_step": float(stats["alg_bytes"]), "setup_s": setup_s,
                       "host_wall_ms_per_step": wall_ms / K, "output_checksum": checksum},
            "roofline": roofline, "cpu_baseline": base, "e2e": e2e,
            "gpu_launches": int(stats["kernel_launches"]) * K, "clocks": clocks.summary()}
    emit(line)


if __name__ == "__main__":
    main()
