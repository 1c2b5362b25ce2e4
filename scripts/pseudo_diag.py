"""Side-by-side summary of a pseudo-Hermitian solve on the GPU vs the golden trace of the reference CPU solver
(debug aid for tests/test_pseudo_gpu.py): python scripts/pseudo_diag.py <golden name> [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.golden_util import load, parse_trace  # noqa: E402
from tests.test_pseudo_gpu import _golden_spectrum, _matrix, _solve  # noqa: E402


def main():
    name = sys.argv[1]
    g = load(name)
    H, _ = _matrix(g)
    out = {"name": name}
    try:
        res = _solve(H, g)
    except Exception as e:  # noqa: BLE001
        print("SOLVE FAILED:", e)
        out["error"] = str(e)
        res = None
    p = g["problems"][0]
    ref = parse_trace(p["trace"])
    if res is not None:
        got = parse_trace(res.trace)
        nev = g["nev"]
        exact = _golden_spectrum(g)[:nev]
        out.update(
            iterations=[res.iterations, p["iterations"]], filtered=[res.filtered_vecs, p["filtered_vecs"]],
            locks=[got["locks"], ref["locks"]], applyk=[got["applyk"], ref["applyk"]], dos=[got["dos"], ref["dos"]],
            lanczos=[got["lanczos"], ref["lanczos"]],
            qr=[got["qr"], ref["qr"]], swaps=[int(res.stats["swaps"]), p["swaps"]],
            h2_sched_equal=[h[:2] for h in got["hemm_h2"]] == [h[:2] for h in ref["hemm_h2"]],
            n_h2=[len(got["hemm_h2"]), len(ref["hemm_h2"])],
            eig_err_vs_ref=float(np.max(np.abs(res.ritzv[:nev] - np.array(p["ritzv"][:nev])) / exact)),
            eig_err_vs_exact=float(np.max(np.abs(res.ritzv[:nev] - exact) / exact)),
            max_resid=float(res.resid[:nev].max()), qr_log=res.qr_log,
            theta_err=float(np.max(np.abs(np.sort(got["theta"]) - np.sort(ref["theta"])))) if got["theta"] is not None else None,
            stats=res.stats,
        )
        for it, (a, b) in enumerate(zip(got["ritzv"], ref["ritzv"])):
            n = min(len(a), len(b), g["nev"] + g["nex"])
            out[f"ritz_err_it{it}"] = float(np.max(np.abs(a[:n] - b[:n]) / np.maximum(np.abs(b[:n]), 1e-300)))
        for it, (a, b) in enumerate(zip(got["resid"], ref["resid"])):
            n = min(len(a), len(b))
            out[f"resid_ratio_it{it}"] = [float(np.min(a[:n] / b[:n])), float(np.max(a[:n] / b[:n]))]
    print(json.dumps(out, indent=1, default=str))
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], "w"), indent=1, default=str)


if __name__ == "__main__":
    main()
