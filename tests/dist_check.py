"""Multi-GPU parity check, run under torchrun (one rank per GPU of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py

Every rank computes gen.phi with the row-sharded engine; rank 0 compares the gathered matrix
bit for bit with the single-GPU engine and with the oracle.  Exit code 0 = all equal.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import genlib_b200 as gen
    from oracle import binding as ob
    from util import random_pedigree

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    failures = []

    def check(name, ped, probands, want=None, numerics="reference"):
        got = gen.phi_distributed(ped, probands, numerics=numerics, device=local)
        if rank == 0:
            single = gen.phi(ped, probands, numerics=numerics, device=local)
            ok = got.shape == single.shape and np.array_equal(got.view(np.uint32), single.view(np.uint32))
            if want is not None:
                ok = ok and np.array_equal(got.view(np.uint32), want.view(np.uint32))
            print(f"[dist x{world}] {name}: {'ok' if ok else 'MISMATCH'} ({got.shape[0]}x{got.shape[0]})", flush=True)
            if not ok:
                failures.append(name)
                if got.shape[0] <= 29:
                    np.set_printoptions(linewidth=250, precision=4)
                    print("got\n", got, "\nwant\n", single, flush=True)
                else:
                    bad = np.argwhere(got != single)
                    print("  mismatches", len(bad), "of", got.size, "first", bad[:5].tolist(), "rows affected", len(set(bad[:, 0])), "cols", len(set(bad[:, 1])), flush=True)

    ped = gen.genealogy(gen.geneaJi)
    check("geneaJi", ped, None, ob.OraclePedigree.from_csv(gen.geneaJi).phi() if rank == 0 else None)
    check("geneaJi all individuals", ped, np.arange(1, 30))
    ped = gen.genealogy(gen.genea140)
    check("genea140", ped, None, ob.OraclePedigree.from_csv(gen.genea140).phi() if rank == 0 else None)
    check("genea140 fp64", ped, None, numerics="fp64")
    for seed in range(4):
        rng = np.random.default_rng(3000 + seed)
        rec = random_pedigree(rng, int(rng.integers(200, 1500)), int(rng.integers(4, 40)), p_single=0.15,
                              p_none=0.03, window=int(rng.choice([0, 60, 300])))
        ped = gen.genealogy(rec)
        pro = rng.permutation(ped.ids)[: int(rng.integers(5, 200))]
        want = None
        if rank == 0:
            want = ob.OraclePedigree.from_arrays(rec["ind"], rec["father"], rec["mother"], rec["sex"]).phi(pro)
        check(f"random {seed}", ped, pro, want)
    for name, scale in (("C3", 0.05), ("C4", 0.01), ("C5", 0.05)):
        s = gen.synth.config(name, scale)
        ped = gen.genealogy(s.as_columns())
        check(f"{name} x{scale}", ped, s.probands)
    # gen.sparse_phi's schedule on the sharded engine: against the oracle's restatement of sparse_phi
    for seed in range(3):
        rng = np.random.default_rng(4000 + seed)
        rec = random_pedigree(rng, int(rng.integers(300, 1500)), int(rng.integers(4, 40)), p_single=0.15,
                              p_none=0.03, window=int(rng.choice([0, 60, 300])))
        ped = gen.genealogy(rec)
        pro = rng.permutation(ped.ids)[: int(rng.integers(5, 200))]
        got = gen.phi_distributed(ped, pro, device=local, schedule="sparse_phi")
        if rank == 0:
            want, _ = ob.sparse_phi_ranks(ped.father, ped.mother, ped.rank_of(pro), ids=ped.ids)
            ok = got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
            print(f"[dist x{world}] sparse_phi schedule, random {seed}: {'ok' if ok else 'MISMATCH'} ({got.shape[0]}x{got.shape[0]})", flush=True)
            if not ok:
                failures.append(f"sparse {seed}")
    dist.barrier()
    flag = torch.tensor([len(failures)], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if flag.item():
        sys.exit(1)
    if rank == 0:
        print("dist_check: all equal", flush=True)


if __name__ == "__main__":
    main()
