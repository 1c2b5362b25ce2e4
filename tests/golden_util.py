"""Helpers to read the golden traces written by oracle/ref_driver.cpp."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return json.load(open(os.path.join(GOLDEN, name + ".json")))


def parse_trace(trace):
    """-> dict with hemm [(block, off_left)], qr [(locked, cond)], locks [n], ritzv [arrays], resid [arrays], lanczos"""
    out = dict(hemm=[], qr=[], locks=[], ritzv=[], resid=[], shift=[], lanczos=None, dos=None, theta=None, tau=None,
               hemm_h2=[], applyk=[])
    for t in trace:
        f = t.split()
        if f[0] == "HEMM":
            out["hemm"].append((int(f[1]), int(f[4]), float(f[2]), float(f[3])))
        elif f[0] == "HEMM_H2":  # (block, off_left, alpha, beta, gamma)
            out["hemm_h2"].append((int(f[1]), int(f[5]), float(f[2]), float(f[3]), float(f[4])))
        elif f[0] == "ApplyK":
            out["applyk"].append(int(f[1]))
        elif f[0] == "QR":
            out["qr"].append((int(f[1]), float(f[2])))
        elif f[0] == "Lock":
            out["locks"].append(int(f[1]))
        elif f[0] == "RITZV":
            out["ritzv"].append(np.array([float(x) for x in f[1:]]))
        elif f[0] == "RESID":
            out["resid"].append(np.array([float(x) for x in f[2:]]))
        elif f[0] == "Shift":
            out["shift"].append(float(f[1]))
        elif f[0] == "Lanczos":
            out["lanczos"] = (int(f[1]), int(f[2]), float(f[3]))
        elif f[0] == "Lanczos1":
            out["lanczos"] = (int(f[1]), 1, float(f[2]))
        elif f[0] == "LanczosDos":
            out["dos"] = (int(f[1]), int(f[2]))
        elif f[0] == "THETA":
            out["theta"] = np.array([float(x) for x in f[1:]])
        elif f[0] == "TAU":
            out["tau"] = np.array([float(x) for x in f[1:]])
    return out


DT = {"d": np.float64, "z": np.complex128, "s": np.float32, "c": np.complex64}
