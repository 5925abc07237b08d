"""Why is planning slower inside e2e than back to back?  (a) back to back, (b) 100 ms idle between
plans, (c) 400 MB of other memory touched between plans."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genlib_b200 as gen

s = gen.synth.config("C3")
ped = gen.genealogy(s.as_columns())
ranks = ped.rank_of(s.probands)
junk = np.zeros(100 << 20, np.float32)
for label, between in (("back to back", lambda: None), ("100 ms idle", lambda: time.sleep(0.1)),
                       ("400 MB touched", lambda: junk.fill(1.0)), ("idle + touched", lambda: (time.sleep(0.1), junk.fill(2.0)))):
    ts = []
    for _ in range(6):
        between()
        t = time.time(); plan = gen.Plan(ped.father, ped.mother, ranks); ts.append((time.time() - t) * 1e3); del plan
    print(f"{label}: min {min(ts[1:]):.1f} ms  median {sorted(ts[1:])[2]:.1f} ms", flush=True)
