"""The N > 1 path without GPUs: the row-sharded schedule (owners, local rows, rank-major couples)
replayed on the CPU, in one process and as a world_size-2 gloo job."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from plan_replay import replay_sharded
from util import random_pedigree

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharded_replay_geneaji(gen, ob, world):
    ped = gen.genealogy(gen.geneaJi)
    plan = gen.Plan(ped.father, ped.mother, ped.rank_of(gen.pro(ped)), world=world)
    assert np.array_equal(replay_sharded(plan), ob.OraclePedigree.from_csv(gen.geneaJi).phi())


@pytest.mark.parametrize("seed", range(5))
def test_sharded_replay_random(gen, ob, seed):
    rng = np.random.default_rng(500 + seed)
    rec = random_pedigree(rng, int(rng.integers(60, 300)), int(rng.integers(2, 12)), p_single=0.15, p_none=0.03,
                          window=int(rng.choice([0, 40])))
    ped = gen.genealogy(rec)
    o = ob.OraclePedigree.from_arrays(rec["ind"], rec["father"], rec["mother"], rec["sex"])
    pro = rng.permutation(ped.ids)[: int(rng.integers(2, 40))]
    want = o.phi(pro)
    for world in (2, 4):
        plan = gen.Plan(ped.father, ped.mother, ped.rank_of(pro), world=world)
        traffic = []
        got = replay_sharded(plan, exchange=lambda t, a, b, n: traffic.append(n))
        assert np.array_equal(got, want)
        assert sum(traffic) > 0                         # the shards do talk to each other


@pytest.mark.parametrize("seed", range(4))
def test_sharded_replay_sparse_phi_schedule(gen, ob, seed):
    """gen.sparse_phi's schedule (queue order, Float32 stores and halves) on 2 and 3 ranks."""
    rng = np.random.default_rng(700 + seed)
    rec = random_pedigree(rng, int(rng.integers(80, 500)), int(rng.integers(3, 12)), window=int(rng.choice([0, 40, 90])))
    ped = gen.genealogy(rec)
    pro = rng.permutation(ped.ids)[: int(rng.integers(2, 40))]
    ranks = ped.rank_of(pro)
    want, _ = ob.sparse_phi_ranks(ped.father, ped.mother, ranks, ids=ped.ids)
    for world in (2, 3):
        plan = gen.Plan(ped.father, ped.mother, ranks, world=world, schedule="sparse_phi", ids=ped.ids)
        assert np.array_equal(replay_sharded(plan), want)


def test_shard_invariants(gen):
    s = gen.synth.generate(6000, 10, 300, alpha=0.05, demes=2, migration=0.1, overlap=2, seed=3)
    ped = gen.genealogy(s.as_columns())
    one = gen.Plan(ped.father, ped.mother, ped.rank_of(s.probands), world=1)
    for world in (2, 8):
        plan = gen.Plan(ped.father, ped.mother, ped.rank_of(s.probands), world=world)
        assert plan.n_layers == one.n_layers and plan.row_updates == one.row_updates
        # column slots are global; their number depends on the rank-major couple order only through
        # which individuals share a recycled line of slots
        assert abs(plan.capacity - one.capacity) <= 0.15 * one.capacity + 256
        rows = [plan.rank_rows(g) for g in range(world)]
        assert max(rows) <= 1.4 * sum(rows) / world + 64          # balanced row storage
        used = [set() for _ in range(world)]
        for t in range(plan.n_layers):
            info, arr, sh = plan.layer_info(t), plan.layer_arrays(t), plan.layer_shard(t)
            fb, mb = sh["fam_base"], sh["mem_base"]
            assert np.all(fb % 4 == 0) and fb[-1] == info["n_fam"] and mb[-1] == info["n_new"]
            assert np.array_equal(arr["member_owner"], np.repeat(np.arange(world), np.diff(mb)))
            sizes = np.diff(mb)
            if info["n_new"] > 40 * world:
                assert sizes.max() <= 1.15 * info["n_new"] / world + 40   # balanced work per layer
            # a local row is never handed out twice while it is live
            live_rows = {(int(o), int(r)) for o, r in zip(sh["live_owner"], sh["live_lrow"]) if o >= 0}
            for o, r in zip(arr["member_owner"], sh["member_lrow"]):
                assert (int(o), int(r)) not in live_rows
                live_rows.add((int(o), int(r)))
                used[o].add(int(r))
        assert all((max(u) + 1 if u else 1) <= r for u, r in zip(used, rows))
        owner, lrow = plan.proband_rows()
        assert len(owner) == plan.n_unique and (owner >= 0).all() and (owner < world).all()


@pytest.mark.parametrize("worker", ["gloo_worker.py", "gloo_protocol_worker.py"])
def test_world2_gloo(tmp_path, worker):
    """Two processes, gloo backend.  gloo_worker: each simulates its own rank and exchanges peer memory;
    gloo_protocol_worker: gen.run_distributed with a stub engine (all fine / all start over / one rank fails)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "result.txt"
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, worker), str(r), "2", str(port), str(out)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            p.kill()
            logs.append("timeout")
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)[-3000:]
    assert out.read_text() == "ok", "\n".join(logs)[-3000:]
