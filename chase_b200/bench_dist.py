"""Multi-GPU arm of bench.py (one rank per GPU under torchrun).

Headline: STRONG scaling of BASELINE config C2 (real double N = 20000, nev = 1000, nex = 400: the problem of the 1-GPU
line, unchanged) over an r x c grid (2x1, 2x2, 4x2), block-cyclic layout with 64 x 64 blocks as in the reference's
examples (examples/1_hello_world/1_hello_world.cpp:102).  Extra keys: `weak` = the round-1 arm (N grows like sqrt(#GPUs)
so that every GPU keeps a 3.2 GB block: N = 28288, 40000, 56576) and `complex_fixed` (z N = 24000, same at every G).  The matrix is the same dense uniform-spectrum generator as the
single-GPU arm, A = Q diag(lambda) Q^T with 3 Householder reflectors, written as diag + 9 rank-one terms so that every
rank builds its own block on its own GPU without ever holding N^2 numbers.
"""
from __future__ import annotations

import ctypes
import json
import os

import numpy as np

BASE = ("d", 20000, 1000, 400)


def weak_n(gpus: int) -> int:
    return int(round(BASE[1] * np.sqrt(gpus) / 64.0)) * 64


def sequence_spectrum(N: int, step: int, delta: float = 1e-4):
    """Spectrum of problem `step` of a correlated sequence (BASELINE config C3): the uniform spectrum
    with lambda_k <- lambda_k (1 + delta g_k) applied once per step (g ~ N(0,1), seeded per step); same eigenvectors,
    spectrum known exactly.  step 0 = the unperturbed problem."""
    lam = 100.0 * (1e-4 + np.arange(N) * (1.0 - 1e-4) / N)
    for t in range(1, step + 1):
        lam = lam * (1.0 + delta * np.random.default_rng(1000 + t).standard_normal(N))
    return lam


def lowrank_terms(N: int, cplx: bool, lam=None):
    """A = diag(lam) + sum_t coef[t] * X[:, t] Y[:, t]^H, identical to bench.make_matrix / oracle.dense_from_spectrum
    before its final symmetrisation (lam: the reference generator's uniform spectrum unless given)."""
    if lam is None:
        lam = 100.0 * (1e-4 + np.arange(N) * (1.0 - 1e-4) / N)
    rng = np.random.default_rng(7)
    X, Y, coef = [], [], []

    def apply(v):
        w = lam * v
        for x, y, c in zip(X, Y, coef):
            w = w + c * x * np.vdot(y, v)
        return w

    for _ in range(3):
        v = rng.standard_normal(N)
        if cplx:
            v = v + 1j * rng.standard_normal(N)
        v = v / np.linalg.norm(v)
        w = apply(v)
        s = np.vdot(v, w)
        X += [v, w, v]
        Y += [w, v, v]
        coef += [-2.0, -2.0, 4.0 * s]
    return lam, np.stack(X, 1), np.stack(Y, 1), np.array(coef)


def local_block(N, gr, gc, cplx, device, transposed=False, lam=None):
    """This rank's block A[gr, gc] as a torch tensor (m_loc, n_loc); transposed=True returns it as (n_loc, m_loc)
    row-major, i.e. the column-major m_loc x n_loc array the solver wants, without an extra copy."""
    import torch

    lam, X, Y, coef = lowrank_terms(N, cplx, lam)
    dt = torch.complex128 if cplx else torch.float64
    Xl = torch.from_numpy(X[gr, :] * coef[None, :]).to(device).to(dt)
    Yl = torch.from_numpy(Y[gc, :]).to(device).to(dt)
    # global diagonal entries inside this block
    pos = {int(g): k for k, g in enumerate(gc)}
    rows = [k for k, g in enumerate(gr) if int(g) in pos]
    cols = [pos[int(gr[k])] for k in rows]
    dvals = torch.from_numpy(lam[gr[rows]]).to(device).to(dt) if rows else None
    if transposed:
        A = Yl.conj() @ Xl.T  # n_loc x m_loc
        if rows:
            A[cols, rows] += dvals
    else:
        A = Xl @ Yl.conj().T  # m_loc x n_loc
        if rows:
            A[rows, cols] += dvals
    return A, lam


def fill_local_block(ptr, ld, N, gr, gc, cplx, device, lam=None, chunk=4096):
    """Writes this rank's block A[gr, gc] straight into the column-major device buffer at `ptr` (leading dimension `ld`
    elements), `chunk` columns at a time: the same matrix as local_block(), without ever holding a second copy."""
    import torch

    lam, X, Y, coef = lowrank_terms(N, cplx, lam)
    dt = torch.complex128 if cplx else torch.float64

    class _View:  # zero-copy torch view of the solver's buffer: (n_loc, ld) row-major == column-major m_loc x n_loc
        __cuda_array_interface__ = {"shape": (len(gc), int(ld)), "typestr": "<c16" if cplx else "<f8",
                                    "data": (int(ptr), False), "version": 3}

    H = torch.as_tensor(_View(), device=device)
    Xl = torch.from_numpy(X[gr, :] * coef[None, :]).to(device).to(dt)  # m_loc x 9
    Yl = torch.from_numpy(Y[gc, :]).to(device).to(dt)  # n_loc x 9
    pos = {int(g): k for k, g in enumerate(gr)}
    for c0 in range(0, len(gc), chunk):
        c1 = min(c0 + chunk, len(gc))
        blk = Yl[c0:c1].conj() @ Xl.T  # (c1 - c0) x m_loc: columns c0..c1 of the block, transposed
        cols = [c for c in range(c0, c1) if int(gc[c]) in pos]
        if cols:
            rows = [pos[int(gc[c])] for c in cols]
            blk[[c - c0 for c in cols], rows] += torch.from_numpy(lam[gc[cols]]).to(device).to(dt)
        H[c0:c1, :len(gr)] = blk
        del blk
    torch.cuda.synchronize()
    return lam


def _flag(L, name, v):
    getattr(L, name)(ctypes.byref(ctypes.c_int(v)))


def extra_arm(world, t, N, nev, nex, nb, solves, label):
    """One fixed problem solved `solves` times after one warm-up, matrix built per block on the GPUs and handed over on
    the device, device RNG start vectors: an extra (non-headline) measurement reported under its own key."""
    import torch

    import chase_b200
    from chase_b200 import dist as cd

    L = chase_b200.lib()
    G = world.size
    r, c = cd.grid_dims(G)
    i, j = cd.grid_coords(r, c, "R", world.rank)
    gr, gc = cd.global_indices(N, r, nb, i), cd.global_indices(N, c, nb, j)
    cplx = t == "z"
    dt = np.complex128 if cplx else np.float64
    solver = cd.PChASE(world, N, nev, nex, dt, grid=(r, c), major="R", mb=nb, nb=nb)
    At, lam = local_block(N, gr, gc, cplx, f"cuda:{world.device}", transposed=True)
    solver.load_device_matrix(At.data_ptr(), len(gr))
    del At
    torch.cuda.empty_cache()
    _flag(L, "chase_b200_set_device_rng_", 1)
    _flag(L, "chase_b200_set_matrix_resident_", 1)
    solver.solve(deg=20, tol=1e-10, copy=False)  # warm-up
    torch.cuda.synchronize()
    L.chase_b200_device_sync()
    world.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rs = [solver.solve(deg=20, tol=1e-10, copy=False) for _ in range(solves)]
    e1.record()
    torch.cuda.synchronize()
    L.chase_b200_device_sync()
    world.barrier()
    secs = world.max(e0.elapsed_time(e1) * 1e-3)
    rel = max(float(np.max(np.abs(x.ritzv[:nev] - lam[:nev]) / lam[:nev])) for x in rs)
    assert rel < 1e-10, f"{label}: eigenvalues off: {rel}"
    st = rs[-1].stats
    solver.finalize()
    _flag(L, "chase_b200_set_matrix_resident_", 0)
    return {"workload": f"{label}: {t} N={N} nev={nev} nex={nex} uniform spectrum, dense Q diag Q^H, {r}x{c} grid, "
                        f"block-cyclic {nb}", "solves": solves, "warmup": 1,
            "value": sum(x.stats["gflop_filter"] for x in rs) * 1e9 / secs / 1e12, "unit": "TFLOP/s",
            "time_to_solution_s": secs / solves, "iterations": rs[-1].iterations,
            "filtered_vecs": rs[-1].filtered_vecs,
            "filter_phase_tflops": st["gflop_filter"] / st["t_filter"] / 1e3 if st["t_filter"] > 0 else None,
            "phases_s": {k[2:]: st[k] for k in ("t_all", "t_initvecs", "t_lanczos", "t_filter", "t_qr", "t_rr", "t_resid")},
            "max_rel_eig_err": rel}


def run(a):
    """STRONG scaling of BASELINE config C2 (the workload of the 1-GPU line, N = 20000 fixed) over the r x c grid;
    extra keys: `weak` (N = 20000 sqrt(G), the round-1 arm) and `complex_fixed` (z N = 24000, nev = 1000, nex = 400,
    the same problem at every G)."""
    import torch
    import torch.distributed as tdist

    import chase_b200
    from chase_b200 import dist as cd

    if "WORLD_SIZE" not in os.environ or int(os.environ["WORLD_SIZE"]) != a.gpus:
        # started without the launcher: re-launch ourselves the way the driver does
        import subprocess
        import sys

        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533"] + sys.argv
        raise SystemExit(subprocess.call(cmd))

    from bench import Clocks, hemm_traffic  # noqa: E402 (bench.py is on sys.path)

    L = chase_b200.lib()
    world = cd.World()
    rank, G = world.rank, world.size
    t, N, nev, nex = BASE
    m = nev + nex
    r, c = cd.grid_dims(G)
    nb = 64
    i, j = cd.grid_coords(r, c, "R", rank)
    gr, gc = cd.global_indices(N, r, nb, i), cd.global_indices(N, c, nb, j)
    dev = f"cuda:{world.device}"
    A, lam = local_block(N, gr, gc, False, dev)
    Hh = torch.empty((len(gc), len(gr)), dtype=A.dtype, pin_memory=True)  # column-major m_loc x n_loc
    Hh.copy_(A.T.contiguous())
    del A
    torch.cuda.empty_cache()
    Vh = torch.zeros((m, max(len(gr), 1)), dtype=Hh.dtype, pin_memory=True)
    H, V = Hh.numpy().T, Vh.numpy().T

    peak = max(L.chase_b200_dmma_peak(40000, None) for _ in range(3)) / 1e12

    solver = cd.PChASE(world, N, nev, nex, H, grid=(r, c), major="R", mb=nb, nb=nb, V_loc=V)
    tol = 1e-10

    def check(res):
        rel = float(np.max(np.abs(res.ritzv[:nev] - lam[:nev]) / lam[:nev]))
        assert rel < 1e-10, f"eigenvalues off: {rel}"
        assert float(res.resid[:nev].max()) < 100 * tol
        return rel

    def sync():
        torch.cuda.synchronize()
        L.chase_b200_device_sync()
        world.barrier()

    def timed(nsteps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.chase_b200_launch_count()
        e0.record()
        rs = [solver.solve(deg=20, tol=tol, copy=False) for _ in range(nsteps)]
        e1.record()
        sync()
        return rs, world.max(e0.elapsed_time(e1) * 1e-3), L.chase_b200_launch_count() - l0

    # ---- value: local blocks resident in HBM, device RNG ----------------------------------------------------
    _flag(L, "chase_b200_set_device_rng_", 1)
    _flag(L, "chase_b200_set_matrix_resident_", 0)
    solver.solve(deg=20, tol=tol, copy=False)
    _flag(L, "chase_b200_set_matrix_resident_", 1)
    for _ in range(max(a.warmup - 1, 0)):
        solver.solve(deg=20, tol=tol, copy=False)
    L.chase_b200_hemm_profile_enable(1)
    ck = Clocks(world.device)
    if rank == 0:
        ck.start()
    rs, secs, launches = timed(a.steps)
    clocks = ck.stop() if rank == 0 else None
    hp = (ctypes.c_double * 4)()
    L.chase_b200_hemm_profile_read(hp)
    L.chase_b200_hemm_profile_enable(0)
    rel = max(check(x) for x in rs)
    flop_filter = sum(x.stats["gflop_filter"] for x in rs) * 1e9  # whole job (global N)
    value = flop_filter / secs / 1e12
    st = rs[-1].stats
    kern_tf = hp[2] / (hp[1] * 1e-3) / 1e12 if hp[1] > 0 else 0.0
    kern_tf_min = -world.max(-kern_tf)
    # per-phase seconds: max over ranks (the serial fraction of the strong-scaling curve is read from these)
    phases = {k[2:]: world.max(st[k]) for k in ("t_all", "t_initvecs", "t_lanczos", "t_filter", "t_qr", "t_rr", "t_resid")}

    # ---- e2e: host buffers through p?chase_ ---------------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        _flag(L, "chase_b200_set_device_rng_", 1)
        _flag(L, "chase_b200_set_matrix_resident_", 0)
        solver.solve(deg=20, tol=tol, copy=False)
        rs2, secs2, _ = timed(a.steps)
        rel = max(rel, max(check(x) for x in rs2))
        es = H.itemsize
        e2e = {"value": sum(x.stats["gflop_filter"] for x in rs2) * 1e9 / secs2 / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": N * N * es, "d2h_bytes_per_step": G * (len(gr) * m * es + 2 * m * 8),
               "time_to_solution_s": secs2 / a.steps, "iterations": rs2[-1].iterations,
               "filtered_vecs": rs2[-1].filtered_vecs,
               "start_vectors": "device Philox RNG, regenerated inside every timed step (the reference GPU backend "
                                "regenerates with cuRAND every solve); nothing is cached between steps"}
    solver.finalize()
    del Hh, Vh
    _flag(L, "chase_b200_set_matrix_resident_", 0)

    # ---- extra arms (not the headline): the round-1 weak-scaled C2 and a fixed complex problem ------------------------
    extras = {}
    if not getattr(a, "no_extras", False):
        if G > 1:
            extras["weak"] = extra_arm(world, "d", weak_n(G), nev, nex, nb, 2, "c2 weak-scaled (N = 20000 sqrt(G))")
        extras["complex_fixed"] = extra_arm(world, "z", 24000, 1000, 400, nb, 2, "fixed complex problem")

    if rank == 0:
        roofline = {"bound": "tensor", "kernel": "hemm_tma_kernel (FP64 DMMA, TMA-fed; local A and A^H blocks)",
                    "achieved": kern_tf, "peak": peak, "unit": "TFLOP/s", "frac": kern_tf / peak,
                    "traffic": hemm_traffic(f"c2_g{G}"),
                    "peak_source": "measured live per GPU: register-resident DMMA.8x8x4 loop (chase_b200_dmma_peak)",
                    "launches": int(hp[0]), "kernel_ms_total": hp[1], "kernel_share_of_step": hp[1] * 1e-3 / secs,
                    "achieved_min_over_ranks": kern_tf_min, "per": "GPU (rank 0)"}
        line = {
            "metric": "filter_hemm_tflops_per_time_to_solution", "value": value, "unit": "TFLOP/s", "n_gpus": G,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * secs / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"c2: d N={N} nev={nev} nex={nex} uniform spectrum, dense Q diag Q^H, "
                                   f"tol 1e-10 deg 20 opt (the same problem at every GPU count); {r}x{c} grid, "
                                   f"block-cyclic {nb}x{nb}, NCCL",
                       "l2": "inputs larger than L2 (local A block is %.2f GB)" % (len(gr) * len(gc) * H.itemsize / 1e9),
                       "start_vectors": "device Philox RNG, regenerated every solve (value and e2e)"},
            "time_to_solution_s": secs / a.steps, "iterations": rs[-1].iterations, "filtered_vecs": rs[-1].filtered_vecs,
            "filter_phase_tflops": st["gflop_filter"] / st["t_filter"] / 1e3 if st["t_filter"] > 0 else None,
            "phases_s": phases, "phases_s_note": "last timed solve, max over ranks",
            "max_rel_eig_err": rel, "gpu_launches": int(launches), "gpu_launches_per": "rank 0", "clocks": clocks,
            "roofline": roofline, "e2e": e2e,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    world.close()
    if tdist.is_initialized():
        tdist.destroy_process_group()


# ---- pseudo-Hermitian (BSE) benchmark matrix (BASELINE config C5) ---------------------------------------------------
def bse_terms(N: int, seed: int = 11, lam_min: float = 1.0, lam_max: float = 100.0, coupling: float = 0.3,
              nrefl: int = 3):  # same defaults as oracle.chase_oracle.bse_matrix
    """Low-rank description of the synthetic BSE matrix H = [[A, B], [-conj(B), -conj(A)]] with exactly known
    spectrum +-lam (same matrix as the test generator oracle.chase_oracle.bse_matrix, checked in
    tests/test_pseudo_cpu.py):  A = Q diag(a) Q^H, B = Q diag(b) Q^T, Q = I + X Y^H (nrefl Householder reflectors),
        A = diag(a) + X Ga + Pa X^H,   Ga = Y^H diag(a) + (Y^H diag(a) Y) X^H,        Pa = diag(a) Y
        B = diag(b) + X Gb + Pb X^T,   Gb = Y^H diag(b) + (Y^H diag(b) conj(Y)) X^T,  Pb = diag(b) conj(Y)
    so that every rank can form its own block on its own GPU from O(N) data."""
    assert N % 2 == 0
    k = N // 2
    rng = np.random.default_rng(seed)
    lam = lam_min + (lam_max - lam_min) * (np.arange(k) / max(k - 1, 1))
    phase = np.exp(2j * np.pi * rng.random(k))
    b = coupling * lam * phase
    a = np.sqrt(lam**2 + np.abs(b) ** 2)
    X = np.zeros((k, nrefl), dtype=np.complex128)
    Y = np.zeros((k, nrefl), dtype=np.complex128)
    for p in range(nrefl):
        v = rng.standard_normal(k) + 1j * rng.standard_normal(k)
        v /= np.linalg.norm(v)
        Qv = v + X[:, :p] @ (Y[:, :p].conj().T @ v)
        X[:, p] = -2.0 * Qv
        Y[:, p] = v
    Ya = Y.conj().T * a
    Ga = Ya + (Ya @ Y) @ X.conj().T
    Pa = a[:, None] * Y
    Yb = Y.conj().T * b
    Gb = Yb + (Yb @ Y.conj()) @ X.T
    Pb = b[:, None] * Y.conj()
    return dict(k=k, lam=lam, a=a, b=b, X=X, Ga=Ga, Pa=Pa, Gb=Gb, Pb=Pb)


def bse_local_block(N, gr, gc, device, dtype=None, transposed=False, terms=None, **kw):
    """This rank's block H[gr, gc] of the synthetic BSE matrix as a torch tensor (m_loc, n_loc), or with
    transposed=True as (n_loc, m_loc) row-major = the column-major m_loc x n_loc array the solver wants."""
    import torch

    t = terms or bse_terms(N, **kw)
    k = t["k"]
    dt = dtype or torch.complex128
    gr, gc = np.asarray(gr), np.asarray(gc)
    Hb = torch.zeros((len(gr), len(gc)), dtype=torch.complex128, device=device)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(device)  # noqa: E731
    rsel = [np.nonzero(gr < k)[0], np.nonzero(gr >= k)[0]]
    csel = [np.nonzero(gc < k)[0], np.nonzero(gc >= k)[0]]
    for bi in (0, 1):
        ri = rsel[bi]
        if len(ri) == 0:
            continue
        rowblk = torch.zeros((len(ri), len(gc)), dtype=torch.complex128, device=device)
        for bj in (0, 1):
            cj = csel[bj]
            if len(cj) == 0:
                continue
            lr, lc = gr[ri] - bi * k, gc[cj] - bj * k  # indices inside the k x k sub-block
            if bi == bj:  # A (top-left) or -conj(A) (bottom-right)
                blk = dev(t["X"][lr]) @ dev(t["Ga"][:, lc]) + dev(t["Pa"][lr]) @ dev(t["X"][lc].conj().T)
                dvec = t["a"]
            else:  # B (top-right) or -conj(B) (bottom-left)
                blk = dev(t["X"][lr]) @ dev(t["Gb"][:, lc]) + dev(t["Pb"][lr]) @ dev(t["X"][lc].T)
                dvec = t["b"]
            pos = {int(g): q for q, g in enumerate(lc)}
            rows = [q for q, g in enumerate(lr) if int(g) in pos]
            if rows:
                cols = [pos[int(lr[q])] for q in rows]
                blk[rows, cols] += dev(dvec[lr[rows]].astype(np.complex128))
            if bi == 1:
                blk = -blk.conj()
            rowblk.index_copy_(1, torch.from_numpy(cj).to(device), blk.resolve_conj())
            del blk
        Hb.index_copy_(0, torch.from_numpy(ri).to(device), rowblk)
        del rowblk
    Hb = Hb.to(dt)
    if transposed:
        Hb = Hb.T.contiguous()
    return Hb, t["lam"]
