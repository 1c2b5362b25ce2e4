"""FP32 problems through ?chase_ on the three FP32 routes (tcgen05 kind::tf32 with 3 / 4 partial products, the round-1
FP64 copy, the generic widening kernel): iterations, filtered vectors, time, eigenvalue error against the reference CPU
FP32 trace / the known spectrum.  One subprocess per route (the route is read at construction)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, time
import numpy as np
sys.path.insert(0, %r)
import chase_b200
from oracle import chase_oracle as co
from tests.golden_util import DT, load
out = []
for name in ("serial_clement_s_N256", "serial_clement_c_N256"):
    g = load(name); p = g["problems"][0]
    H = co.clement(g["N"], DT[g["type"]])
    with chase_b200.ChASE(H, g["nev"], g["nex"]) as s:
        res = s.solve(deg=g["deg"], tol=g["tol"])
    refv = np.array(p["ritzv"][:g["nev"]])
    out.append(dict(case=name, iters=res.iterations, ref_iters=p["iterations"], vecs=res.filtered_vecs, ref_vecs=p["filtered_vecs"],
                    eig_err=float(np.max(np.abs(res.ritzv[:g["nev"]] - refv) / np.abs(refv))), max_resid=float(res.resid[:g["nev"]].max())))
for t, N, nev, nex in (("s", 6000, 300, 120), ("c", 4000, 200, 80)):
    dt = np.float32 if t == "s" else np.complex64
    lam = co.uniform_spectrum(N)
    H = co.dense_from_spectrum(lam, np.float64 if t == "s" else np.complex128).astype(dt)
    with chase_b200.ChASE(np.asfortranarray(H), nev, nex) as s:
        s.solve(); t0 = time.time(); res = s.solve(); dtm = time.time() - t0
    out.append(dict(case=f"uniform_{t}_N{N}", iters=res.iterations, vecs=res.filtered_vecs, secs=dtm, t_filter=res.stats["t_filter"],
                    filter_tflops=res.stats["gflop_filter"] / res.stats["t_filter"] / 1e3,
                    eig_err=float(np.max(np.abs(res.ritzv[:nev] - lam[:nev]) / lam[:nev])), max_resid=float(res.resid[:nev].max())))
print(json.dumps(out))
''' % ROOT
res = {}
for label, env in (("tf32_3", {"CHASE_B200_FP32_PATH": "tf32"}), ("tf32_all4", {"CHASE_B200_FP32_PATH": "tf32", "CHASE_B200_TF32_TERMS": "4"}),
                   ("fp64copy", {"CHASE_B200_FP32_PATH": "fp64copy"}), ("generic", {"CHASE_B200_FP32_PATH": "generic"})):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=600)
    try:
        res[label] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        res[label] = {"error": (r.stdout + r.stderr)[-600:]}
    print(label, json.dumps(res[label]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "fp32_paths.json"), "w"), indent=1)
