"""Multi-rank parity check of the distributed backend (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_check.py [--grid 2x2] [--layout block|cyclic] [--case c1_clement_d_N1001] ...

Every rank builds the (small) global matrix, keeps its local block, solves through p?chase_init_[blockcyclic_] /
p?chase_, and the gathered result is compared with the golden trace of the UNMODIFIED reference CPU solver
(tests/golden/*.json): because the start block is the reference's global mt19937 stream regardless of the layout, the
distributed solve must take the same decisions (iterations, filtered vectors, HEMM schedule, locks) and deliver the
same eigenvalues (1e-10 relative) as the serial reference."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from chase_b200 import dist as cd  # noqa: E402
from oracle import chase_oracle as co  # noqa: E402  (checker only)
from tests.golden_util import DT, load, parse_trace  # noqa: E402


def run_case(world, name, grid, layout, major):
    import torch.distributed as dist

    g = load(name)
    p = g["problems"][0]
    dt = DT[g["type"]]
    N, nev, nex = g["N"], g["nev"], g["nex"]
    H = co.clement(N, dt) if g["matrix"] == "clement" else co.uniform_diag(N, dt)
    nb = 0 if layout == "block" else 32
    r, c = grid
    i, j = cd.grid_coords(r, c, major, world.rank)
    gr, gc = cd.global_indices(N, r, nb, i), cd.global_indices(N, c, nb, j)
    Hloc = np.asfortranarray(H[np.ix_(gr, gc)])
    os.environ.update(g.get("env", {}))  # e.g. CHASE_DISABLE_CHOLQR=1: Householder QR in every iteration
    try:
        with cd.PChASE(world, N, nev, nex, Hloc, grid=grid, major=major, mb=nb, nb=nb) as s:
            res = s.solve(deg=g["deg"], tol=g["tol"], opt="S" if g["opt"] else "N", trace=True)
    finally:
        for k in g.get("env", {}):
            os.environ.pop(k, None)
    if g.get("env", {}).get("CHASE_DISABLE_CHOLQR") == "1" and "householder" not in res.qr_log:
        return dict(case=name, grid=f"{grid[0]}x{grid[1]}", layout=layout, major=major, fails=["Householder QR not used"])
    ref, got = parse_trace(p["trace"]), parse_trace(res.trace)
    fails = []
    if res.iterations != p["iterations"]:
        fails.append(f"iterations {res.iterations} != {p['iterations']}")
    if res.filtered_vecs != p["filtered_vecs"]:
        fails.append(f"filtered_vecs {res.filtered_vecs} != {p['filtered_vecs']}")
    if [(b, o) for (b, o, _, _) in got["hemm"]] != [(b, o) for (b, o, _, _) in ref["hemm"]]:
        fails.append("HEMM schedule differs")
    if got["locks"] != ref["locks"]:
        fails.append(f"locks {got['locks']} != {ref['locks']}")
    refv = np.array(p["ritzv"][:nev])
    rel = float(np.max(np.abs(res.ritzv[:nev] - refv) / np.abs(refv)))
    if rel > 1e-10:
        fails.append(f"eigenvalues off by {rel:.2e}")
    if not np.all(res.resid[:nev] < 100 * g["tol"]):
        fails.append("residuals above tolerance")
    # assemble the eigenvectors (column layout: rows split over grid rows, replicated over grid columns)
    parts = [None] * world.size
    if world.size > 1:
        dist.all_gather_object(parts, (i, j, gr, res.V[:len(gr), :nev]))
    else:
        parts = [(i, j, gr, res.V[:len(gr), :nev])]
    V = np.zeros((N, nev), dtype=dt)
    for (pi, pj, rows, blk) in parts:
        if pj == 0:
            V[rows, :] = blk
    rr = np.linalg.norm(H @ V - V * res.ritzv[:nev], axis=0)
    if not (np.all(rr < 1e-8) and np.all(rr > 0)):
        fails.append(f"recomputed residual max {rr.max():.2e}")
    orth = float(np.linalg.norm(V.conj().T @ V - np.eye(nev)))
    if orth > 1e-9:
        fails.append(f"orthogonality {orth:.2e}")
    # replicas over grid columns must agree
    for (pi, pj, rows, blk) in parts:
        if pj != 0 and np.max(np.abs(V[rows, :] - blk)) > 1e-12:
            fails.append("column-layout replicas differ")
    return dict(case=name, grid=f"{r}x{c}", layout=layout, major=major, iterations=res.iterations,
                filtered_vecs=res.filtered_vecs, max_rel_eig=rel, max_resid=float(rr.max()), orth=orth, fails=fails)


def run_pseudo_case(world, name, grid, layout, major):
    """Pseudo-Hermitian (BSE) solve on a grid (p?chase_init_pseudo_[blockcyclic_]) against the golden trace of the
    serial reference solver (chase_ref_cpu_pz, Solve_pseudo) and the known spectrum."""
    import torch.distributed as dist

    from tests.test_pseudo_gpu import _golden_spectrum, _matrix

    g = load(name)
    p = g["problems"][0]
    N, nev, nex = g["N"], g["nev"], g["nex"]
    H, _ = _matrix(g)
    nb = 0 if layout == "block" else 32
    r, c = grid
    i, j = cd.grid_coords(r, c, major, world.rank)
    gr, gc = cd.global_indices(N, r, nb, i), cd.global_indices(N, c, nb, j)
    with cd.PChASE(world, N, nev, nex, np.asfortranarray(H[np.ix_(gr, gc)]), grid=grid, major=major, mb=nb, nb=nb,
                   pseudo=True) as s:
        if "numlanczos" in g:
            L = s._lib
            import ctypes

            L.chase_set_num_lanczos_(ctypes.byref(ctypes.c_int(g["numlanczos"])))
            L.chase_set_lanczos_iter_(ctypes.byref(ctypes.c_int(g["lanczositer"])))
        res = s.solve(deg=g["deg"], tol=g["tol"], opt="S" if g["opt"] else "N", trace=True)
    ref, got = parse_trace(p["trace"]), parse_trace(res.trace)
    fails = []
    if res.iterations != p["iterations"]:
        fails.append(f"iterations {res.iterations} != {p['iterations']}")
    if res.filtered_vecs != p["filtered_vecs"]:
        fails.append(f"filtered_vecs {res.filtered_vecs} != {p['filtered_vecs']}")
    if [h[:2] for h in got["hemm_h2"]] != [h[:2] for h in ref["hemm_h2"]]:
        fails.append("HEMM_H2 schedule differs")
    if got["locks"] != ref["locks"]:
        fails.append(f"locks {got['locks']} != {ref['locks']}")
    if got["applyk"] != ref["applyk"]:
        fails.append("ApplyKconjugate sequence differs")
    exact = _golden_spectrum(g)[:nev]
    rel = float(np.max(np.abs(res.ritzv[:nev] - exact) / exact))
    relref = float(np.max(np.abs(res.ritzv[:nev] - np.array(p["ritzv"][:nev])) / exact))
    if max(rel, relref) > 1e-10:
        fails.append(f"eigenvalues off by {rel:.2e} (exact) / {relref:.2e} (reference)")
    if not np.all(res.resid[:nev] < 1000 * g["tol"]):
        fails.append("residuals above tolerance")
    parts = [None] * world.size
    if world.size > 1:
        dist.all_gather_object(parts, (i, j, gr, res.V[:len(gr), :nev]))
    else:
        parts = [(i, j, gr, res.V[:len(gr), :nev])]
    V = np.zeros((N, nev), dtype=H.dtype)
    for (pi, pj, rows, blk) in parts:
        if pj == 0:
            V[rows, :] = blk
    rr = np.linalg.norm(H @ V - V * res.ritzv[:nev], axis=0)
    if not (np.all(rr < 1e-7 * np.abs(exact).max()) and np.all(rr > 0)):
        fails.append(f"recomputed residual max {rr.max():.2e}")
    for (pi, pj, rows, blk) in parts:
        if pj != 0 and np.max(np.abs(V[rows, :] - blk)) > 1e-12:
            fails.append("column-layout replicas differ")
    return dict(case=name + " (pseudo-Hermitian)", grid=f"{r}x{c}", layout=layout, major=major,
                iterations=res.iterations, filtered_vecs=res.filtered_vecs, max_rel_eig=rel, max_resid=float(rr.max()),
                fails=fails)


def run_io_case(world, grid, layout, major):
    """p?chase_readHam_ / p?chase_wrtHam_ on a grid: every rank reads the pieces of its local block from one global
    column-major file, solves, and writes them back into a second file, which must equal the first byte for byte."""
    import tempfile

    import torch.distributed as dist

    N, nev, nex = 256, 24, 16
    H = co.clement(N, np.float64)
    H[3, 9] = H[9, 3] = 0.5
    tmp = tempfile.gettempdir()
    src, dst = os.path.join(tmp, "chase_b200_io_in.bin"), os.path.join(tmp, "chase_b200_io_out.bin")
    if world.rank == 0:
        np.asfortranarray(H).T.tofile(src)
        if os.path.exists(dst):
            os.remove(dst)
    world.barrier()
    nb = 0 if layout == "block" else 32
    r, c = grid
    i, j = cd.grid_coords(r, c, major, world.rank)
    gr, gc = cd.global_indices(N, r, nb, i), cd.global_indices(N, c, nb, j)
    fails = []
    with cd.PChASE(world, N, nev, nex, np.zeros((len(gr), len(gc)), order="F"), grid=grid, major=major, mb=nb, nb=nb) as s:
        s._lib.pdchase_readHam_(src.encode())
        if not np.array_equal(s.H, H[np.ix_(gr, gc)]):
            fails.append("local block read from file differs")
        res = s.solve(deg=16, tol=1e-10)
        s._lib.pdchase_wrtHam_(dst.encode())
    world.barrier()
    w = np.linalg.eigvalsh(H)[:nev]
    rel = float(np.max(np.abs(res.ritzv[:nev] - w) / np.abs(w)))
    if rel > 1e-10:
        fails.append(f"eigenvalues off by {rel:.2e}")
    if world.rank == 0 and open(src, "rb").read() != open(dst, "rb").read():
        fails.append("file written by wrtHam differs from the file read")
    if world.size > 1:
        flags = [None] * world.size
        dist.all_gather_object(flags, fails)
        fails = [f for fl in flags for f in fl]
    return dict(case="readHam/wrtHam", grid=f"{r}x{c}", layout=layout, major=major, max_rel_eig=rel, fails=fails)


def run_sequence(world, name, grid, layout, major):
    """tests/noinput.cpp-style sequence on a grid: problem 0 random start, then perturbed matrices re-using the
    distributed V / ritzv (mode 'A'); the local host blocks are re-read at every solve."""
    g = load(name)
    dt = DT[g["type"]]
    N, nev, nex = g["N"], g["nev"], g["nex"]
    H = co.clement(N, dt)
    nb = 0 if layout == "block" else 32
    r, c = grid
    i, j = cd.grid_coords(r, c, major, world.rank)
    gr, gc = cd.global_indices(N, r, nb, i), cd.global_indices(N, c, nb, j)
    nseq = len(g["problems"])
    stream = co.mt_normal(1337, (2 if g["type"] == "z" else 1) * nseq * N * N)
    pos, fails, its = 0, [], []
    with cd.PChASE(world, N, nev, nex, np.asfortranarray(H[np.ix_(gr, gc)]), grid=grid, major=major, mb=nb, nb=nb) as s:
        for idx, p in enumerate(g["problems"]):
            res = s.solve(deg=g["deg"], tol=g["tol"], mode="R" if idx == 0 else "A", trace=True)
            its.append(res.iterations)
            ref, got = parse_trace(p["trace"]), parse_trace(res.trace)
            if res.iterations != p["iterations"] or res.filtered_vecs != p["filtered_vecs"]:
                fails.append(f"problem {idx}: iterations/filtered {res.iterations}/{res.filtered_vecs} != "
                             f"{p['iterations']}/{p['filtered_vecs']}")
            if [(b, o) for (b, o, _, _) in got["hemm"]] != [(b, o) for (b, o, _, _) in ref["hemm"]]:
                fails.append(f"problem {idx}: HEMM schedule differs")
            refv = np.array(p["ritzv"][:nev])
            rel = float(np.max(np.abs(res.ritzv[:nev] - refv) / np.abs(refv)))
            if rel > 1e-10:
                fails.append(f"problem {idx}: eigenvalues off by {rel:.2e}")
            if idx + 1 < nseq:
                pos += co.perturb_hermitian(H, stream[pos:], 1e-4)
                s.H[...] = H[np.ix_(gr, gc)]
    return dict(case=name + " (sequence)", grid=f"{r}x{c}", layout=layout, major=major, iterations=its, fails=fails)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="serial_clement_d_N256,serial_clement_z_N256,c1_clement_d_N1001,c2s_uniform_d_N2000,"
                                       "hhqr_clement_z_N256")
    ap.add_argument("--grid", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--no-seq", action="store_true")
    ap.add_argument("--pseudo-cases", default="pseudo_bse_z_N200,pseudo_synth_z_N600")
    a = ap.parse_args()
    world = cd.World()
    grids = [tuple(int(x) for x in a.grid.split("x"))] if a.grid else [cd.grid_dims(world.size)]
    if not a.grid and world.size in (4, 8):
        grids.append((world.size, 1))
    results, bad = [], 0
    for name in a.cases.split(","):
        for grid in grids:
            for layout, major in (("block", "R"), ("cyclic", "C")):
                r = run_case(world, name, grid, layout, major)
                results.append(r)
                bad += len(r["fails"])
                if world.rank == 0:
                    print(("FAIL " if r["fails"] else "ok   ") + json.dumps(r), flush=True)
    for name in [x for x in a.pseudo_cases.split(",") if x]:
        for grid in grids:
            for layout, major in (("block", "R"), ("cyclic", "C")):
                r = run_pseudo_case(world, name, grid, layout, major)
                results.append(r)
                bad += len(r["fails"])
                if world.rank == 0:
                    print(("FAIL " if r["fails"] else "ok   ") + json.dumps(r), flush=True)
    for grid in grids[:1]:
        for layout, major in (("block", "R"), ("cyclic", "C")):
            r = run_io_case(world, grid, layout, major)
            results.append(r)
            bad += len(r["fails"])
            if world.rank == 0:
                print(("FAIL " if r["fails"] else "ok   ") + json.dumps(r), flush=True)
    if not a.no_seq:
        for grid in grids[:1]:
            for layout, major in (("block", "R"), ("cyclic", "R")):
                for name in ("seq_clement_d_N400", "seq_clement_z_N400"):
                    r = run_sequence(world, name, grid, layout, major)
                    results.append(r)
                    bad += len(r["fails"])
                    if world.rank == 0:
                        print(("FAIL " if r["fails"] else "ok   ") + json.dumps(r), flush=True)
    if world.rank == 0 and a.out:
        json.dump(results, open(a.out, "w"), indent=1)
    world.close()
    import torch.distributed as dist

    if dist.is_initialized():
        dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
