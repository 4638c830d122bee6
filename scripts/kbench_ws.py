#!/usr/bin/env python
"""Dominant kernel of a non-default workload timed alone (bench.workload_roofline) — the target of the ncu captures of
gs_color_kernel (convdiff) and spmvB_kernel<3> (elasticity) at the top level:  kbench_ws.py convdiff|elasticity [refs] [base_mult]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ugcore_b200 import problems as pr
from ugcore_b200.solver import host_init

w = sys.argv[1] if len(sys.argv) > 1 else "convdiff"
refs = int(sys.argv[2]) if len(sys.argv) > 2 else (7 if w == "convdiff" else 6)
bm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
host_init(0, None)
spec = bench.workload_spec(w)
prob = pr.Problem(dim=3, num_refs=refs, problem=spec["problem"], base=(bm,) * 3, **spec["kw"])
peak, src = bench.measured_peak_gbs()
print(json.dumps(bench.workload_roofline(w, prob, refs, peak, src)))
