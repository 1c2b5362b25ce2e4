"""Distributed solve of a BASELINE-shaped problem with a known spectrum, matrix built on the GPUs (never on the host).

    torchrun --nproc-per-node 8 scripts/run_dist.py --type z --N 120000 --nev 1000 --nex 400   # BASELINE config C4

A = Q diag(lambda) Q^H with the reference generator's uniform spectrum (lambda_k = 100 (1e-4 + k (1 - 1e-4) / N),
examples/2_input_output/2_input_output.cpp:250-262) and Q = 3 Householder reflectors; every rank forms its
block-cyclic block from 9 rank-one terms on its own GPU and hands it to the solver on the device.

    torchrun --nproc-per-node 8 scripts/run_dist.py --pseudo --type z --N 40000 --nev 500 --nex 200   # config C5

    torchrun --nproc-per-node 4 scripts/run_dist.py --type z --N 60000 --nev 1000 --nex 300 --layout block --seq 5  # C3

--seq k: a sequence of k correlated problems (same eigenvectors, spectrum perturbed by 1e-4 relative per step,
chase_b200.bench_dist.sequence_spectrum): the first is solved from random vectors, the others re-use the previous
eigenvectors and Ritz values (mode 'A'), the new local block being rebuilt and handed over on the device each time.
--layout block: the reference's block distribution instead of block-cyclic.

--pseudo: the pseudo-Hermitian (BSE) problem class through p?chase_init_pseudo_blockcyclic_ on the synthetic BSE matrix
of chase_b200.bench_dist.bse_terms (spectrum +-lam known exactly; type z or c)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chase_b200  # noqa: E402
from chase_b200 import bench_dist as bd  # noqa: E402
from chase_b200 import dist as cd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--type", default="z")
ap.add_argument("--N", type=int, default=120000)
ap.add_argument("--nev", type=int, default=1000)
ap.add_argument("--nex", type=int, default=400)
ap.add_argument("--nb", type=int, default=64)
ap.add_argument("--solves", type=int, default=1)
ap.add_argument("--out", default="")
ap.add_argument("--pseudo", action="store_true")
ap.add_argument("--seq", type=int, default=1)
ap.add_argument("--perturb", type=float, default=1e-4)
ap.add_argument("--layout", default="cyclic", choices=["cyclic", "block"])
ap.add_argument("--tol", type=float, default=0.0)
ap.add_argument("--deg", type=int, default=20)
ap.add_argument("--lam-max", type=float, default=100.0, help="--pseudo: positive spectrum is uniform in [1, lam-max]")
ap.add_argument("--upperb-scale", type=float, default=1.0,
                help="chase_set_upperb_scale_rate_: safety factor on the Lanczos estimate of max(lambda^2)")
ap.add_argument("--max-iter", type=int, default=25)
ap.add_argument("--in-place", action="store_true",
                help="generate the local block directly into the solver's device buffer (no second copy)")
a = ap.parse_args()

L = chase_b200.lib()
world = cd.World()
G = world.size
r, c = cd.grid_dims(G)
i, j = cd.grid_coords(r, c, "R", world.rank)
if a.layout == "block":
    a.nb = 0
gr, gc = cd.global_indices(a.N, r, a.nb, i), cd.global_indices(a.N, c, a.nb, j)
cplx = a.type in ("z", "c")
dt = {"z": np.complex128, "c": np.complex64, "d": np.float64}[a.type]
tol = a.tol or (1e-10 if a.type in ("z", "d") else 1e-5)
solver = cd.PChASE(world, a.N, a.nev, a.nex, dt, grid=(r, c), major="R", mb=a.nb, nb=a.nb, pseudo=a.pseudo)
# row-major (n_loc, m_loc) == column-major m_loc x n_loc with ld = m_loc
if a.in_place and not a.pseudo:
    # the block is generated straight into the solver's buffer (it exists once in HBM): C4 on 2 GPUs = 115 GB per GPU
    ptr, ld = solver.device_matrix()
    lam = bd.fill_local_block(ptr, ld, a.N, gr, gc, cplx, f"cuda:{world.device}")
    solver.mark_device_matrix()
    At = None
elif a.pseudo:
    At, lam = bd.bse_local_block(a.N, gr, gc, f"cuda:{world.device}",
                                 dtype=torch.complex128 if a.type == "z" else torch.complex64, transposed=True,
                                 lam_max=a.lam_max)
else:
    At, lam = bd.local_block(a.N, gr, gc, cplx, f"cuda:{world.device}", transposed=True)
    if a.type == "c":  # the generator works in double precision
        At = At.to(torch.complex64)
if At is not None:
    solver.load_device_matrix(At.data_ptr(), len(gr))
del At
torch.cuda.empty_cache()
if a.seq > 1 and a.pseudo:
    raise SystemExit("--seq is implemented for Hermitian problems")
import ctypes  # noqa: E402

L.chase_b200_set_device_rng_(ctypes.byref(ctypes.c_int(1)))
L.chase_set_upperb_scale_rate_(ctypes.byref(ctypes.c_float(a.upperb_scale)))
L.chase_set_max_iter_(ctypes.byref(ctypes.c_int(a.max_iter)))
out = []
for s in range(a.solves * a.seq):
    step = s % a.seq
    if a.seq > 1 and s > 0:
        # next problem of the sequence: same Q, perturbed spectrum; previous eigenpairs stay in solver.V / ritzv
        lam_s = bd.sequence_spectrum(a.N, step, a.perturb)
        At, lam = bd.local_block(a.N, gr, gc, cplx, f"cuda:{world.device}", transposed=True, lam=lam_s)
        if a.type == "c":
            At = At.to(torch.complex64)
        lam = np.sort(lam)
        solver.load_device_matrix(At.data_ptr(), len(gr))
        del At
        torch.cuda.empty_cache()
    torch.cuda.synchronize()
    world.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = solver.solve(deg=a.deg, tol=tol, copy=False, trace=bool(os.environ.get("CHASE_B200_TRACE")),
                       mode="A" if (a.seq > 1 and step > 0) else "R")
    e1.record()
    torch.cuda.synchronize()
    secs = world.max(e0.elapsed_time(e1) * 1e-3)
    rel = float(np.max(np.abs(res.ritzv[:a.nev] - lam[:a.nev]) / lam[:a.nev]))
    st = res.stats
    es = np.dtype(dt).itemsize
    rec = dict(type=a.type, pseudo_hermitian=bool(a.pseudo), tol=tol, lam_max=a.lam_max if a.pseudo else None,
               sequence_step=step if a.seq > 1 else None, mode="A" if (a.seq > 1 and step > 0) else "R",
               upperb_scale=a.upperb_scale, N=a.N, nev=a.nev, nex=a.nex, gpus=G,
               grid=f"{r}x{c}", layout=f"block-cyclic {a.nb}" if a.nb else "block",
               time_to_solution_s=secs, iterations=res.iterations, filtered_vecs=res.filtered_vecs,
               filter_tflops_whole_job=st["gflop_filter"] / st["t_filter"] / 1e3,
               filter_tflops_per_gpu=st["gflop_filter"] / st["t_filter"] / 1e3 / G,
               phases_s={k[2:]: st[k] for k in st if k.startswith("t_")}, max_rel_eig_err=rel,
               max_resid=float(res.resid[:a.nev].max()), local_matrix_gb=len(gr) * len(gc) * es / 1e9,
               qr=res.qr_log, heev_sweeps=st["heev_sweeps"])
    out.append(rec)
    if world.rank == 0:
        print(json.dumps(rec), flush=True)
if world.rank == 0 and a.out:
    json.dump(out, open(a.out, "w"), indent=1)
solver.finalize()
world.close()
if torch.distributed.is_initialized():
    torch.distributed.destroy_process_group()
