"""Times solver:init (uploads, entry-stream building, smoother preprocess, base factorisation) of the configs[1] solver:
the first call (cold context) and three re-inits with a re-assembled matrix of the same pattern (Solver.set_matrix — what
every Newton / time step pays).  python scripts/time_init.py [refs]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from ugcore_b200 import problems as pr, solver as S  # noqa: E402

refs = int(sys.argv[1]) if len(sys.argv) > 1 else 7
S.host_init(0, None)
t = time.perf_counter()
prob = pr.Problem(dim=3, num_refs=refs)
levels = {l: (prob.matrix(l), prob.prolongation(l) if l else None, prob.restriction(l) if l else None) for l in range(refs + 1)}
t_gen = time.perf_counter() - t
s = S.Solver(bench.solver_desc(refs), levels[refs][0], levels)
times = []
for i in range(4):
    if i:
        s.set_matrix(levels[refs][0], levels)
    t = time.perf_counter()
    s.init()
    times.append(time.perf_counter() - t)
x, ok, h = s.apply(np.array(prob.rhs()))
print(json.dumps({"refs": refs, "generator_s": round(t_gen, 3), "init_s": [round(v, 3) for v in times], "steps": len(h) - 1, "ok": bool(ok),
                  "env": {k: v for k, v in os.environ.items() if k.startswith("UG4B200_")}}))
