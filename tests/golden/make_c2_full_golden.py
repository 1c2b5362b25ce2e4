"""tests/golden/c2_uniform_d_N20000.json from the reference's own full-size run of BASELINE config C2.

Source: baseline/oracle_logs/c2_uniform_N20000_nev1000_nex400_real_double_cpu.log = raw stdout of the UNMODIFIED
reference CPU backend (ChASECPU, -DCHASE_OUTPUT) on the reference's `isMatGen` diagonal uniform-spectrum matrix,
N=20000, nev=1000, nex=400, tol 1e-10, deg 20, opt, mt19937(1337) start vectors (2 267 s on 8 cores; see the README
next to the log).  Per iteration the reference prints `iteration: k lambda lowerb upperb unconverged`, the QR variant
and the first 20 rows of (degree, resid, residLast, ritzv).

    python tests/golden/make_c2_full_golden.py
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LOG = os.path.join(HERE, "..", "..", "baseline", "oracle_logs", "c2_uniform_N20000_nev1000_nex400_real_double_cpu.log")

its, qr, cur = [], [], None
lines = open(LOG).read().splitlines()
for i, ln in enumerate(lines):
    m = re.match(r"cond\(V\): (\S+)", ln)
    if m:
        nxt = lines[i + 1]
        deg = int(re.match(r"choldegree: (\d)", nxt).group(1))
        variant = "shifted2" if "shift =" in nxt else ("chol1" if deg == 1 else "chol2")
        qr.append({"cond": float(m.group(1)), "variant": variant})
    m = re.match(r"iteration: (\d+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\d+)", ln)
    if m:
        cur = {"iteration": int(m.group(1)), "lambda": float(m.group(2)), "lowerb": float(m.group(3)),
               "upperb": float(m.group(4)), "unconverged": int(m.group(5)), "first20": []}
        its.append(cur)
    m = re.search(r"unconverged = \s*(\d+)\s+new_converged\s+(\d+)", ln)
    if m and cur is not None:
        cur["new_converged_printed"] = int(m.group(2))  # iteration 0 prints an uninitialised variable
    m = re.match(r"(\d+)\t(\d+)\t(\S+)\t(\S+)\t(\S+)$", ln)
    if m and cur is not None:
        cur["first20"].append({"degree": int(m.group(2)), "resid": float(m.group(3)), "ritzv": float(m.group(5))})
    if ln.startswith("|         1 |"):
        f = [x.strip() for x in ln.split("|")[1:-1]]
        totals = {"iterations": int(f[1]), "filtered_vecs": int(f[2]), "t_all_s_8cores": float(f[3]),
                  "t_filter_s_8cores": float(f[6])}
out = {"type": "d", "N": 20000, "nev": 1000, "nex": 400, "matrix": "uniform", "tol": 1e-10, "deg": 20, "opt": 1,
       "source": "baseline/oracle_logs/c2_uniform_N20000_nev1000_nex400_real_double_cpu.log", **totals,
       "qr": qr, "iterations_log": its}
json.dump(out, open(os.path.join(HERE, "c2_uniform_d_N20000.json"), "w"), indent=1)
print(totals, [q["variant"] for q in qr], [x["unconverged"] for x in its])
