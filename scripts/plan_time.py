"""Planner timing on this host: python scripts/plan_time.py [C3|C4|C5] (threads via GENLIB_PLAN_THREADS)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genlib_b200 as gen

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
s = gen.synth.config(name)
ped = gen.genealogy(s.as_columns())
ranks = ped.rank_of(s.probands)
for threads in ("1", "2", "4", ""):
    if threads:
        os.environ["GENLIB_PLAN_THREADS"] = threads
    else:
        os.environ.pop("GENLIB_PLAN_THREADS", None)
    ts = []
    for _ in range(5):
        t = time.time(); plan = gen.Plan(ped.father, ped.mother, ranks, world=world); ts.append((time.time() - t) * 1e3); del plan
    print(f"{name} world {world} threads {threads or 'default'}: min {min(ts):.1f} ms  median {sorted(ts)[2]:.1f} ms", flush=True)
