"""Regenerates tests/golden/*.json by running the UNMODIFIED reference CPU solver
(oracle/_ref/chase_ref_cpu_<type>, built from /root/reference by oracle/Makefile).

Run in the build container only (needs /root/reference to build oracle/_ref):
    make -C oracle all && python tests/golden/make_golden.py

Each fixture records, for one problem (or a short sequence), the reference's
iteration count, filtered-vector count, final Ritz values / residuals and the
full ChaseBase call trace (HEMM schedule, QR condition estimates, Ritz values
and residuals per iteration, lock counts).
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CASES = {
    # BASELINE.json configs[0]: Clement N=1001, nev=100, nex=40, real double
    "c1_clement_d_N1001": dict(type="d", N=1001, nev=100, nex=40, matrix="clement", tol=1e-10, deg=20),
    "c1_clement_z_N1001": dict(type="z", N=1001, nev=100, nex=40, matrix="clement", tol=1e-10, deg=20),
    # /root/reference/tests/chase_serial_solve.cpp sizes (N=256, nev=24, nex=16, deg 16)
    "serial_clement_d_N256": dict(type="d", N=256, nev=24, nex=16, matrix="clement", tol=1e-10, deg=16),
    "serial_clement_z_N256": dict(type="z", N=256, nev=24, nex=16, matrix="clement", tol=1e-10, deg=16),
    "serial_clement_s_N256": dict(type="s", N=256, nev=24, nex=16, matrix="clement", tol=1e-5, deg=16),
    "serial_clement_c_N256": dict(type="c", N=256, nev=24, nex=16, matrix="clement", tol=1e-5, deg=16),
    # scaled BASELINE.json configs[1]: uniform spectrum (reference --isMatGen generator)
    "c2s_uniform_d_N2000": dict(type="d", N=2000, nev=100, nex=40, matrix="uniform", tol=1e-10, deg=20),
    # tests/noinput.cpp-style correlated sequence (approximate start vectors)
    "seq_clement_z_N400": dict(type="z", N=400, nev=40, nex=20, matrix="clement", tol=1e-10, deg=20, seq=3, perturb=1e-4),
    "seq_clement_d_N400": dict(type="d", N=400, nev=40, nex=20, matrix="clement", tol=1e-10, deg=20, seq=3, perturb=1e-4),
    # no degree optimisation
    "noopt_clement_d_N300": dict(type="d", N=300, nev=30, nex=10, matrix="clement", tol=1e-10, deg=20, opt=0),
    # Householder QR in every iteration (reference: CHASE_DISABLE_CHOLQR=1 -> houseHoulderQR, chase_cpu.hpp:670-690)
    "hhqr_clement_d_N300": dict(type="d", N=300, nev=30, nex=10, matrix="clement", tol=1e-10, deg=20,
                                env={"CHASE_DISABLE_CHOLQR": "1"}),
    "hhqr_clement_z_N256": dict(type="z", N=256, nev=24, nex=16, matrix="clement", tol=1e-10, deg=16,
                                env={"CHASE_DISABLE_CHOLQR": "1"}),
    # pseudo-Hermitian (BSE) solves, reference binary chase_ref_cpu_p<z|c> (Solve_pseudo).  First case = the
    # configuration of /root/reference/tests/chase_distributed_solve_pseudo_bse_test.cpp:131-250 on the reference's
    # own fixture; "bse_fixture:<file>" is read from tests/golden/bse_fixtures/, "bse_synth:<seed>" is
    # oracle.chase_oracle.bse_matrix(N, seed=<seed>) (exactly known spectrum).
    "pseudo_bse_z_N200": dict(type="pz", N=200, nev=20, nex=20, matrix="bse_fixture:cdouble_random_BSE.bin", tol=1e-10,
                              deg=20, numlanczos=10, lanczositer=40),
    "pseudo_bse_z_N200_dflt": dict(type="pz", N=200, nev=20, nex=10, matrix="bse_fixture:cdouble_random_BSE.bin",
                                   tol=1e-10, deg=20),
    "pseudo_bse_c_N200": dict(type="pc", N=200, nev=20, nex=20, matrix="bse_fixture:cfloat_random_BSE.bin", tol=1e-5,
                              deg=10),
    "pseudo_synth_z_N600": dict(type="pz", N=600, nev=40, nex=20, matrix="bse_synth:11", tol=1e-10, deg=20),
    "pseudo_synth_z_N600_noopt": dict(type="pz", N=600, nev=40, nex=20, matrix="bse_synth:11", tol=1e-10, deg=20,
                                      opt=0),
}


def run_case(name, c):
    exe = os.path.join(ROOT, "oracle", "_ref", f"chase_ref_cpu_{c['type']}")
    out = os.path.join(HERE, name + ".json")
    cmd = [exe, "--out", out]
    tmp = None
    for k, v in c.items():
        if k in ("type", "env"):
            continue
        if k == "matrix" and str(v).startswith("bse_fixture:"):
            v = "file:" + os.path.join(HERE, "bse_fixtures", v.split(":", 1)[1])
        elif k == "matrix" and str(v).startswith("bse_synth:"):
            sys.path.insert(0, ROOT)
            import numpy as np
            from oracle import chase_oracle as co

            Hm, _ = co.bse_matrix(c["N"], np.complex128 if c["type"] == "pz" else np.complex64,
                                  seed=int(v.split(":", 1)[1]))
            tmp = os.path.join(HERE, "_tmp_matrix.bin")
            Hm.T.tofile(tmp)  # column-major on disk
            v = "file:" + tmp
        cmd += [f"--{k}", str(v)]
    env = dict(os.environ, OPENBLAS_NUM_THREADS="8", OMP_NUM_THREADS="8", **c.get("env", {}))
    subprocess.check_call(cmd, env=env, stdout=subprocess.DEVNULL)
    j = json.load(open(out))
    j["matrix"] = c["matrix"]
    if "env" in c:
        j["env"] = c["env"]
    if tmp:
        os.remove(tmp)
    # keep fixtures small: round-trip through json with no extra whitespace
    json.dump(j, open(out, "w"), separators=(",", ":"))
    p = j["problems"]
    print(name, [(q["iterations"], q["filtered_vecs"]) for q in p], os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    only = sys.argv[1:]
    for n, c in CASES.items():
        if only and n not in only:
            continue
        run_case(n, c)
