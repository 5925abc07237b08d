"""One rank of the world_size-2 CPU test (gloo): each process simulates ONLY its own rank of the
row-sharded schedule with NumPy and exchanges what the GPU kernels move through NVLink peer
memory (parent rows read from peers, mirror stores into carried rows) through torch.distributed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(rank: int, world: int, port: int, out_path: str):
    import hashlib
    import torch.distributed as dist
    import genlib_b200 as gen
    from oracle import binding as ob
    from plan_replay import ShardedReplay
    from util import random_pedigree

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    cases = [("geneaJi", gen.genealogy(gen.geneaJi), None)]
    rng = np.random.default_rng(42)
    for k in range(3):
        rec = random_pedigree(rng, int(rng.integers(80, 260)), int(rng.integers(3, 12)), p_single=0.15,
                              p_none=0.03, window=int(rng.choice([0, 50])))
        ped = gen.genealogy(rec)
        cases.append((f"random{k}", ped, rng.permutation(ped.ids)[: int(rng.integers(4, 40))], rec))
    for case, schedule in [(c, sch) for c in cases for sch in ("phi", "sparse_phi")]:
        name, ped, pro = case[0], case[1], case[2]
        IDs = gen.pro(ped) if pro is None else pro
        plan = gen.Plan(ped.father, ped.mother, ped.rank_of(IDs), world=world, schedule=schedule, ids=ped.ids)
        # every rank must have built the same schedule
        sig = hashlib.sha256()
        for t in range(plan.n_layers):
            for v in list(plan.layer_arrays(t).values()) + list(plan.layer_shard(t).values()):
                sig.update(np.ascontiguousarray(v).tobytes())
        sigs = [None] * world
        dist.all_gather_object(sigs, sig.hexdigest())
        ok &= len(set(sigs)) == 1
        R = ShardedReplay(plan)

        def sync_rows():                       # peers' frontier rows as this rank would read them
            blocks = [None] * world
            dist.all_gather_object(blocks, R.A[rank])
            for g in range(world):
                if g != rank:
                    R.A[g] = blocks[g]

        for t in range(plan.n_layers):
            if not R.begin(t):
                continue
            sync_rows()                        # parent rows (and the diagonal's entry) may live on a peer
            R.step(rank)
            writes = [None] * world            # mirror stores into carried rows (maybe remote)
            dist.all_gather_object(writes, R.writes)
            R.writes = [w for ws in writes for w in ws]
            R.apply_writes(only={rank})
            dist.barrier()
        parts = [None] * world
        dist.all_gather_object(parts, R.result_rows(rank))
        n = plan.n_unique
        got = np.zeros((n, n), np.float32)
        for idx, rows in parts:
            got[idx] = rows
        if rank == 0:
            if schedule == "sparse_phi":           # gen.sparse_phi's own values (src/compute.jl:321-447)
                want, _ = ob.sparse_phi_ranks(ped.father, ped.mother, ped.rank_of(IDs), ids=ped.ids)
            elif name == "geneaJi":
                want = ob.OraclePedigree.from_csv(gen.geneaJi).phi()
            else:
                rec = case[3]
                want = ob.OraclePedigree.from_arrays(rec["ind"], rec["father"], rec["mother"], rec["sex"]).phi(IDs)
            ok &= bool(np.array_equal(got, want))
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    dist.destroy_process_group()
    if rank == 0:
        with open(out_path, "w") as fh:
            fh.write("ok" if all(flags) else "fail")


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
