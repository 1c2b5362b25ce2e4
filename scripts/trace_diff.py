"""Diagnostic (run under gpurun): per-iteration comparison of the GPU solve with a golden reference trace."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chase_b200
from oracle import chase_oracle as co
from tests.golden_util import DT, load, parse_trace

name = sys.argv[1] if len(sys.argv) > 1 else "c1_clement_z_N1001"
g = load(name); p = g["problems"][0]
H = co.clement(g["N"], DT[g["type"]]) if g["matrix"] == "clement" else co.uniform_diag(g["N"], DT[g["type"]])
with chase_b200.ChASE(H, g["nev"], g["nex"]) as s:
    res = s.solve(deg=g["deg"], tol=g["tol"], trace=True, opt="S" if g["opt"] else "N")
ref = parse_trace(p["trace"]); got = parse_trace(res.trace)
print("iters", res.iterations, p["iterations"], "vecs", res.filtered_vecs, p["filtered_vecs"], "qr", res.qr_log)
print("lanczos", got["lanczos"], ref["lanczos"])
for it, (a, b) in enumerate(zip(got["ritzv"], ref["ritzv"])):
    n = min(len(a), len(b))
    ra, rb = got["resid"][it], ref["resid"][it]
    m = min(len(ra), len(rb))
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(ra[:m] - rb[:m]) / np.abs(rb[:m])
    print(f"it {it}: block {len(a)}/{len(b)} max|dritz| {np.max(np.abs(a[:n]-b[:n])):.3e}  max rel dresid {np.nanmax(rel):.3e} "
          f"median {np.nanmedian(rel):.3e} locks {got['locks'][it] if it < len(got['locks']) else None}/{ref['locks'][it]}")
# first differing HEMM
for i, (x, y) in enumerate(zip(got["hemm"], ref["hemm"])):
    if x[:2] != y[:2]:
        print("first HEMM difference at call", i, x, y)
        break
