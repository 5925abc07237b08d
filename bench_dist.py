"""bench.py's N > 1 arm: one rank per GPU (torchrun), rows of the frontier sharded over the ranks.

No data-path collective library call: kernels read parent rows and push couple-matrix rows
through NVLink peer mappings (CUDA IPC) and meet at in-stream barriers.  torch.distributed is
the control plane only (handle exchange, timing reduction)."""
from __future__ import annotations

import json
import time

import numpy as np


def main_dist(args, rank, world, local, B):
    import torch
    import torch.distributed as dist
    METRIC, UNIT, ClockSampler, build_workload = B.METRIC, B.UNIT, B.ClockSampler, B.build_workload
    cpu_baseline, emit, measured_peaks = B.cpu_baseline, B.emit, B.measured_peaks

    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    gen, ped, ranks, desc = build_workload(args)          # deterministic: identical on every rank
    esize = 4 if args.numerics == "reference" else 8

    def make_engine():
        plan = gen.Plan(ped.father, ped.mother, ranks, world=world)
        eng = gen.Engine(plan, numerics=args.numerics, device=local, rank=rank)
        handles = [None] * world
        dist.all_gather_object(handles, eng.ipc_handle())
        eng.attach(handles)
        dist.barrier()
        return plan, eng

    def reduce_max(values):
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()

    t0 = time.time()
    plan, eng = make_engine()
    setup_s = time.time() - t0
    rows = plan.row_updates
    W = max(args.warmup, 3)
    for _ in range(W):
        eng.run()
    K = args.steps
    step_ms, cross_ms, couple_ms, expand_ms, wait_ms = [], [], [], [], []
    with ClockSampler(local) as clocks:
        time.sleep(0.6)
        clocks.rows.clear()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(K):
            step_ms.append(eng.run(time_layers=True))
            infos = [eng.layer_info(t) for t in range(plan.n_layers)]
            cross_ms.append([i["ms_cross"] for i in infos])
            couple_ms.append([i["ms_couple"] for i in infos])
            expand_ms.append([i["ms_expand"] for i in infos])
            wait_ms.append([i["ms_wait"] for i in infos])
        torch.cuda.synchronize(); dist.barrier()
        wall_ms = (time.time() - t0) * 1e3
    step_ms = reduce_max(step_ms)                          # per step: slowest rank, device time
    cross_ms = reduce_max(np.array(cross_ms).ravel()).reshape(K, -1)
    couple_ms = reduce_max(np.array(couple_ms).ravel()).reshape(K, -1)
    expand_ms = reduce_max(np.array(expand_ms).ravel()).reshape(K, -1)
    wait_ms = -reduce_max(-np.array(wait_ms).ravel()).reshape(K, -1)      # the rank that waited least
    stats = eng.stats()
    dev_bytes = reduce_max([float(stats["device_bytes"])])[0]
    total_ms = float(step_ms.sum())
    value = rows * K / (total_ms * 1e-3)
    infos = plan.layers()
    cross_bytes = np.array([esize * 4.0 * i["n_new"] * i["live_before"] for i in infos])
    intra_bytes = np.array([esize * 3.0 * i["n_new"] ** 2 for i in infos])
    launched = cross_bytes > 0
    peak1, peak_src = measured_peaks()
    peak = peak1 * world
    c_t = cross_ms[:, launched].sum() * 1e-3
    i_t = (couple_ms + expand_ms).sum() * 1e-3
    achieved = cross_bytes[launched].sum() * K / c_t / 1e9 if c_t > 0 else 0.0
    whole = (cross_bytes.sum() + intra_bytes.sum()) * K / (total_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "cross_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None,
                "peak_source": f"{world} x {peak_src}", "launches": int(launched.sum()) * K,
                "avg_launch_ms": float(cross_ms[:, launched].mean()) if launched.any() else 0.0,
                "alg_bytes_per_launch": float(cross_bytes[launched].mean()) if launched.any() else 0.0,
                "share_of_step": c_t / (total_ms * 1e-3),
                "intra_kernels": {"achieved": intra_bytes.sum() * K / i_t / 1e9 if i_t > 0 else 0.0,
                                  "share_of_step": i_t / (total_ms * 1e-3),
                                  "couple_share": float(couple_ms.sum() / total_ms),
                                  "expand_share": float(expand_ms.sum() / total_ms),
                                  "min_barrier_wait_share": float(wait_ms.sum() / total_ms)},
                "whole_step": {"achieved": whole, "frac": whole / peak, "frac_of_8TBs_nominal": whole / (8000.0 * world)},
                "note": "aggregate over ranks; per-kernel times are the slowest rank's for that kernel"}
    if args.layers_json and rank == 0:
        with open(args.layers_json, "w") as fh:
            json.dump([{**i, "ms_cross": float(cross_ms[:, t].mean()), "ms_couple": float(couple_ms[:, t].mean()),
                        "ms_expand": float(expand_ms[:, t].mean()), "ms_wait_min": float(wait_ms[:, t].mean())}
                       for t, i in enumerate(infos)], fh, indent=1)
    launches = int(reduce_max([float(stats["kernel_launches"])])[0])
    dist.barrier()
    eng.close()

    # ---- end to end: host arrays in, this rank's proband rows out into pinned host memory ----
    e2e_t, st, own_n = [], None, 0
    pinned = None
    for it in range(args.e2e_steps + 1):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.time()
        p2, e2 = make_engine()
        e2.run()
        own_n = len(e2.own_probands())
        if pinned is None:
            pinned = torch.empty((max(own_n, 1), p2.n_unique), dtype=torch.float32).pin_memory()
        out = pinned.numpy()[:own_n]
        e2.fetch(out=out)
        st = e2.stats()
        dist.barrier()
        dt = time.time() - t0
        e2.close()
        if it > 0 or args.e2e_steps == 0:
            e2e_t.append(dt)
    e2e_ms = reduce_max([float(np.mean(e2e_t)) * 1e3])[0]
    h2d = int(reduce_max([float(st["h2d_bytes"] + ped.father.nbytes + ped.mother.nbytes + ranks.nbytes)])[0]) * world
    d2h_t = torch.tensor([float(st["d2h_bytes"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(d2h_t)
    csum = torch.tensor([float(pinned.numpy()[:own_n].astype(np.float64).sum())], dtype=torch.float64, device="cuda")
    dist.all_reduce(csum)
    e2e = {"value": rows / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": int(d2h_t.item()), "ms_per_call": e2e_ms,
           "breakdown_ms": {k: st[k] for k in ("ms_plan", "ms_upload", "ms_kernels", "ms_fetch")},
           "note": "every rank plans the whole pedigree, owns 1/N of the rows and streams its own proband rows to its pinned buffer"}
    base = cpu_baseline(ped, ranks, args.cpu_seconds) if (rank == 0 and args.cpu_seconds > 0) else None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic" if args.workload != "genea140" else "genea140.csv",
                "config": {"workload": desc, "numerics": args.numerics, "storage_bytes": esize,
                           "row_updates_per_step": int(rows), "layers": plan.n_layers,
                           "capacity_slots": int(plan.capacity), "device_bytes_per_rank_max": int(dev_bytes),
                           "parallelism": f"rows sharded over {world} ranks, NVLink peer reads/stores, 2 in-stream barriers per layer",
                           "l2": "working set >> L2 (no flush needed)", "alg_bytes_per_step": float(stats["alg_bytes"]),
                           "setup_s": setup_s, "host_wall_ms_per_step": wall_ms / K, "output_checksum": float(csum.item())},
                "roofline": roofline, "cpu_baseline": base, "e2e": e2e, "gpu_launches": launches * K * world,
                "clocks": clocks.summary()}
        emit(line)
    dist.barrier()
    dist.destroy_process_group()
