"""bench.py's N > 1 arm: one rank per GPU (torchrun), rows of the frontier sharded over the ranks.

No data-path collective library call: the layer kernel reads parent rows from its peers through
NVLink mappings (CUDA IPC) and ranks meet at one in-stream barrier per layer.  torch.distributed is
the control plane only (handle exchange, timing reduction, gathering row digests)."""
from __future__ import annotations

import hashlib
import json
import time

import numpy as np


def main_dist(args, rank, world, local, B):
    import torch
    import torch.distributed as dist
    METRIC, UNIT, ClockSampler, build_workload = B.METRIC, B.UNIT, B.ClockSampler, B.build_workload

    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    gen, ped, ranks = build_workload(args)          # deterministic: identical on every rank
    esize = 4 if args.numerics == "reference" else 8

    def make_engine():
        """Plan (on a worker thread, handed over layer by layer), engine, peers attached, ONE run."""
        plan = gen.Plan(ped.father, ped.mother, ranks, world=world, stream=True)
        eng = gen.run_distributed(plan, numerics=args.numerics, device=local, rank=rank)
        return plan, eng

    def reduce(values, op=dist.ReduceOp.MAX):
        t = torch.tensor(np.asarray(values, np.float64), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return t.cpu().numpy()

    t0 = time.time()
    plan, eng = make_engine()
    setup_s = time.time() - t0
    rows = plan.metric_row_updates
    W = max(args.warmup, 3)
    for _ in range(W - 1):                          # (make_engine ran once)
        eng.run()
    K = args.steps
    step_ms, layer_ms, wait_ms = [], [], []
    with ClockSampler(local) as clocks:
        time.sleep(0.6)
        clocks.rows.clear()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(K):
            step_ms.append(eng.run(time_layers=True))
            infos = [eng.layer_info(t) for t in range(plan.n_layers)]
            layer_ms.append([i["ms_layer"] for i in infos])
            wait_ms.append([i["ms_wait"] for i in infos])
        torch.cuda.synchronize(); dist.barrier()
        wall_ms = (time.time() - t0) * 1e3
    step_ms = reduce(step_ms)                               # per step: slowest rank, device time
    layer_ms = reduce(np.array(layer_ms).ravel()).reshape(K, -1)
    wait_ms = -reduce(-np.array(wait_ms).ravel()).reshape(K, -1)          # the rank that waited least
    stats = eng.stats()
    dev_bytes = reduce([float(stats["device_bytes"])])[0]
    total_ms = float(step_ms.sum())
    value = rows * K / (total_ms * 1e-3)
    infos = [eng.layer_info(t) for t in range(plan.n_layers)]
    # required bytes: summed over the ranks
    keys = ("dram_read_bytes", "dram_write_bytes", "l2_bytes", "nvlink_bytes")
    summed = reduce(np.array([[i[k] for k in keys] for i in infos]).ravel(), dist.ReduceOp.SUM).reshape(len(infos), len(keys))
    for i, row in zip(infos, summed):
        i.update({k: float(v) for k, v in zip(keys, row)})
    peak1, peak_src = B.measured_peaks()
    roofline = B.roofline_record(infos, layer_ms, K, esize, peak1, peak_src, world, args)
    roofline["share_of_step"] = float(layer_ms.sum() / total_ms)
    roofline["min_barrier_wait_share"] = float(wait_ms.sum() / total_ms)
    roofline["note"] = "aggregate over ranks; per-layer times are the slowest rank's"
    if args.layers_json and rank == 0:
        with open(args.layers_json, "w") as fh:
            json.dump([{**i, "ms_layer": float(layer_ms[:, t].mean()), "ms_wait_min": float(wait_ms[:, t].mean())}
                       for t, i in enumerate(infos)], fh, indent=1)
    launches = int(reduce([float(stats["kernel_launches"])])[0])
    dist.barrier()
    eng.close()

    # ---- end to end: host arrays in, this rank's proband rows out into pinned host memory ----
    e2e_t, st, own_n, own_idx = [], None, 0, None
    pinned = None
    for it in range(args.e2e_steps + 1):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.time()
        p2, e2 = make_engine()
        own_idx = e2.own_probands()
        own_n = len(own_idx)
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
    e2e_ms = reduce([float(np.mean(e2e_t)) * 1e3])[0]
    h2d = int(reduce([float(st["h2d_bytes"] + ped.father.nbytes + ped.mother.nbytes + ranks.nbytes)])[0]) * world
    d2h = int(reduce([float(st["d2h_bytes"])], dist.ReduceOp.SUM)[0])
    own = pinned.numpy()[:own_n]
    csum = float(reduce([float(own.astype(np.float64).sum())], dist.ReduceOp.SUM)[0])
    # parity: every rank hashes its own rows; rank 0 orders the digests by proband and hashes them.
    # The whole matrix is gathered (and hashed) only when it is small enough to be worth it.
    n = plan.n_unique
    parts = [None] * world if rank == 0 else None
    dist.gather_object((own_idx, B.rows_digest(own)), parts, dst=0)
    full_sha = None
    if n * n * 4 <= (1 << 30):
        blocks = [None] * world if rank == 0 else None
        dist.gather_object(own.copy(), blocks, dst=0)
        if rank == 0:
            full = np.empty((n, n), np.float32)
            for (idx, _), blk in zip(parts, blocks):
                full[idx] = blk
            full_sha = hashlib.sha256(full.tobytes()).hexdigest()
    e2e = {"value": rows / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_call": e2e_ms,
           "breakdown_ms": {k: st[k] for k in ("ms_plan", "ms_upload", "ms_kernels", "ms_fetch")},
           "note": "every rank plans the whole pedigree on a worker thread and runs each layer as soon as it is planned; it owns 1/N of the rows and streams its own proband rows to its pinned buffer"}
    base = B.cpu_baseline(ped, ranks, args.cpu_seconds) if (rank == 0 and args.cpu_seconds > 0) else None
    if rank == 0:
        digests = [b""] * n
        for idx, dig in parts:
            for k, u in enumerate(idx):
                digests[int(u)] = dig[32 * k: 32 * k + 32]
        parity = B.parity_fields(args, full_sha, hashlib.sha256(b"".join(digests)).hexdigest())
        cfg = B.shared_config(args, rows)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": B.data_tag(args), "config": cfg,
                "engine": {"storage_bytes": esize, "layers": plan.n_layers, "capacity_slots": int(plan.capacity),
                           "device_bytes_per_rank_max": int(dev_bytes),
                           "parallelism": f"rows sharded over {world} ranks, NVLink peer reads of parent rows, 1 in-stream barrier per layer",
                           "alg_bytes_per_step": float(stats["alg_bytes"]), "setup_s": setup_s,
                           "host_wall_ms_per_step": wall_ms / K, "output_checksum": csum},
                "roofline": roofline, "cpu_baseline": base, "e2e": e2e, "parity": parity,
                "gpu_launches": launches * K * world, "clocks": clocks.summary()}
        B.emit(line)
    dist.barrier()
    dist.destroy_process_group()
