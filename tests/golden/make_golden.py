#!/usr/bin/env python
"""Golden hashes of gen.phi at the sizes bench.py measures, from the ORACLE alone.

    python tests/golden/make_golden.py C3 [C5 C4x0.1 genea140 ...]

Each workload is built by the oracle's own loader (no product library in the process), run
through the full reference algorithm (oracle/genlib_oracle.c = src/compute.jl:233-304 restated),
and the sha256 of the raw n x n Float32 matrix goes to tests/golden/<name>_full.sha256 together
with `rows_sha256` (sha256 of the concatenated per-row sha256 digests: what ranks that hold row
shards can compute without gathering the matrix), sum / trace (float64 accumulation) and the run time.  bench.py prints the sha256 of what the
GPU fetched (`output_sha256`) next to the golden one; tests/test_gpu_parity.py compares them."""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

from oracle import binding as ob  # noqa: E402


def parse(tag: str):
    name, _, scale = tag.partition("x")
    return name, float(scale) if scale else 1.0


def golden_name(name: str, scale: float) -> str:
    return f"{name.lower()}_full.sha256" if scale == 1.0 else f"{name.lower()}_x{scale:g}.sha256"


def main():
    for tag in sys.argv[1:]:
        name, scale = parse(tag)
        father, mother, ranks, desc = ob.workload(name, scale)
        t0 = time.time()
        phi, steps = ob.phi_ranks(father, mother, ranks)
        dt = time.time() - t0
        rows_sha = hashlib.sha256(b"".join(hashlib.sha256(np.ascontiguousarray(r).tobytes()).digest() for r in phi)).hexdigest()
        rec = {"sha256": ob.matrix_sha256(phi), "rows_sha256": rows_sha, "workload": desc, "n": int(phi.shape[0]),
               "sum": float(phi.astype(np.float64).sum()), "trace": float(np.trace(phi.astype(np.float64))),
               "row_updates": int(steps[:, 4].sum()), "oracle_seconds": dt, "oracle_threads": ob.num_threads(),
               "source": "oracle/genlib_oracle.c oracle_phi_ranks (C restatement of src/compute.jl:233-304)"}
        path = os.path.join(HERE, golden_name(name, scale))
        with open(path, "w") as fh:
            json.dump(rec, fh, indent=1)
            fh.write("\n")
        print(tag, rec["sha256"], f"{dt:.1f}s", flush=True)


if __name__ == "__main__":
    main()
