"""Writes the planner harness's inputs: python scripts/planner/dump_cases.py OUTDIR [C3 C5 C4 genea140 geneaJi rand...].
One binary file per pedigree: n, n_pro (int64), father, mother (int32), ids (int64), proband ranks (int32)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import genlib_b200 as gen  # noqa: E402
from util import random_pedigree  # noqa: E402


def dump(out, name, father, mother, ids, pro):
    with open(os.path.join(out, f"{name}.bin"), "wb") as fh:
        np.array([len(father), len(pro)], np.int64).tofile(fh)
        np.ascontiguousarray(father, np.int32).tofile(fh)
        np.ascontiguousarray(mother, np.int32).tofile(fh)
        np.ascontiguousarray(ids, np.int64).tofile(fh)
        np.ascontiguousarray(pro, np.int32).tofile(fh)
    print(name, len(father), len(pro))


def main():
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    names = sys.argv[2:] or ["C3", "C5", "genea140", "geneaJi", "rand302", "rand303", "rand305"]
    for name in names:
        if name in ("C3", "C4", "C5"):
            s = gen.synth.config(name)
            ped = gen.genealogy(s.as_columns())
            dump(out, name, ped.father, ped.mother, ped.ids, ped.rank_of(s.probands))
        elif name in ("genea140", "geneaJi"):
            ped = gen.genealogy(os.path.join(ROOT, "tests", "data", f"{name}.csv"))
            dump(out, name, ped.father, ped.mother, ped.ids, ped.rank_of(gen.pro(ped)))
        else:                                               # randSEED: permuted IDs, overlapping generations
            seed = int(name[4:])
            rng = np.random.default_rng(seed)
            ped = gen.genealogy(random_pedigree(rng, 3000 + seed % 3 * 1000, 30 + seed % 40, window=(0, 40, 120)[seed % 3]))
            dump(out, name, ped.father, ped.mother, ped.ids, ped.rank_of(gen.pro(ped)))


if __name__ == "__main__":
    main()
