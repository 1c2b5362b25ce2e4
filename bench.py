#!/usr/bin/env python
"""bench.py — the reference's headline measurement on B200.

Workload (BASELINE.json configs[1], "C2"): real symmetric double, N=20000, nev=1000, nex=400, uniform synthetic
spectrum lambda_k = 100 (1e-4 + k (1-1e-4)/N) (the reference's --isMatGen generator,
examples/2_input_output/2_input_output.cpp:250-262) carried by a DENSE matrix A = Q diag(lambda) Q^T, Q = product of
3 Householder reflectors.  tol 1e-10, deg 20, opt 'S', all other ChaseConfig defaults.

One "step" = one full chase::Solve of that problem.  Metric (BASELINE.json: "time-to-solution (s) + filter HEMM
TFLOP/s"): `value` = filter FLOPs (2 f N^2 x filtered vectors, the reference's own model,
algorithm/performance.hpp:248-260) divided by time-to-solution, in TFLOP/s; `ms_per_step` = time-to-solution.
  * value  : matrix already in HBM, start vectors from the device RNG — no host<->device traffic but the results.
  * e2e    : the same solve through the reference C interface (dchase_init_/dchase_) on HOST buffers: every step
             uploads H (3.2 GB) and the start block and downloads V, ritzv, resid.
  * roofline: the filter HEMM kernel itself (CUDA events around every launch on its stream) against the FP64
             tensor (DMMA) peak measured live on this GPU (MEASURED_PEAKS.json has no FP64 figure).
  * cpu_baseline / --impl reference: the UNMODIFIED reference CPU solver (oracle/_ref, ChASECPU + OpenBLAS) on all
             host cores on a bounded, scaled-down sample of the same workload (same generator, same nev/N, nex/nev).
  * gpu_reference (extra key, informational): the reference's OWN single-GPU backend (ChASEGPU: cuBLAS / cuSOLVER /
             cuRAND, oracle/_ref/chase_ref_gpu_<t>, `make -C oracle refgpu`) solving the FULL workload on this same
             GPU from host buffers — the like-for-like comparison for `e2e` (SURVEY.md 8c/8d "same-box").
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (type, N, nev, nex)
    "c2": ("d", 20000, 1000, 400),
    "c2z": ("z", 12000, 600, 240),
    "smoke": ("d", 3000, 150, 60),
}
OB = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the non-headline arms (weak, complex_fixed)")
    ap.add_argument("--gpu-ref-repeats", type=int, default=3, help="timed solves of the same-box reference GPU backend")
    ap.add_argument("--ref-n", type=int, default=0, help="override the bounded-sample N of the CPU reference arm")
    ap.add_argument("--profile", action="store_true",
                    help="one warm solve + one profiled solve only (for `ncu`: no e2e, no CPU baseline)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# --------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.p, self.index = [], None, index

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1])), pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------------------
# reference CPU arm (oracle/_ref: the unmodified reference ChASECPU built from /root/reference by oracle/Makefile)
# --------------------------------------------------------------------------------------------------------------
def ref_sample_shape(workload, cores, override=0, budget_s=30.0):
    """Bounded sample of the workload for the CPU solver: same generator, same nev/N and nex/N, N chosen so that one
    solve takes about `budget_s` seconds (measured: real double N=6000 -> 12.5 s on 16 cores; time grows like N^3 at
    fixed nev/N).  The full-size C2 run of the same binary is recorded in BASELINE.md / baseline/oracle_logs."""
    t, N, nev, nex = WORKLOADS[workload]
    if override:
        n = override
    else:
        t6000 = 12.5 * 16.0 / max(min(cores, 16), 1) * (4.0 if t == "z" else 1.0)
        n = int(6000.0 * (budget_s / t6000) ** (1.0 / 3.0)) // 500 * 500
        n = max(n, 2000)
    n = min(n, N)
    return t, n, max(nev * n // N, 4), max(nex * n // N, 4)


def run_reference_cpu(workload, override=0, budget_s=30.0):
    """One solve of the bounded sample with the reference CPU solver on all host cores -> dict."""
    cores = os.cpu_count() or 1
    t, n, nev, nex = ref_sample_shape(workload, cores, override, budget_s)
    exe = os.path.join(ROOT, "oracle", "_ref", f"chase_ref_cpu_{t}")
    if not os.path.exists(exe):
        return {"error": f"{exe} not built (oracle/Makefile ref needs /root/reference)"}
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = OB + ":" + env.get("LD_LIBRARY_PATH", "")
    threads = min(cores, 128)  # the wheel's OpenBLAS is built with MAX_THREADS=128
    env["OPENBLAS_NUM_THREADS"] = env["OMP_NUM_THREADS"] = str(threads)
    out = f"/tmp/chase_ref_{os.getpid()}.json"
    t0 = time.time()
    subprocess.run([exe, "--N", str(n), "--nev", str(nev), "--nex", str(nex), "--matrix", "uniform_dense", "--out", out],
                   env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    wall = time.time() - t0
    r = json.load(open(out))["problems"][0]
    os.unlink(out)
    t_all = r["timings"]["All"]
    return {"N": n, "nev": nev, "nex": nex, "type": t, "threads": threads, "t_all": t_all,
            "t_filter": r["timings"]["Filter"], "gflop_filter": r["gflop_filter"], "iterations": r["iterations"],
            "filtered_vecs": r["filtered_vecs"], "wall_incl_matrix_gen": wall,
            "tflops_per_solve": r["gflop_filter"] / t_all / 1e3, "tflops_filter_phase": r["gflop_filter"] / r["timings"]["Filter"] / 1e3}


def run_reference_gpu(workload, repeats=3):
    """The reference's own single-GPU backend (separate process, same GPU) on the FULL workload: one warm-up solve and
    `repeats` timed solves of the same problem in ONE process (fresh cuRAND start vectors each), median and spread
    reported.  BLAS threads pinned so the host-side phases (stemr, the driver's matrix generation) are reproducible."""
    t, N, nev, nex = WORKLOADS[workload]
    exe = os.path.join(ROOT, "oracle", "_ref", f"chase_ref_gpu_{t}")
    if not os.path.exists(exe):
        return {"unavailable": f"{exe} not built (make -C oracle refgpu needs /root/reference)"}
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = OB + ":/usr/local/cuda/lib64:" + env.get("LD_LIBRARY_PATH", "")
    threads = min(os.cpu_count() or 1, 16)
    env["OPENBLAS_NUM_THREADS"] = env["OMP_NUM_THREADS"] = str(threads)
    out = f"/tmp/chase_refgpu_{os.getpid()}.json"
    try:
        subprocess.run([exe, "--N", str(N), "--nev", str(nev), "--nex", str(nex), "--matrix", "uniform_dense", "--repeat",
                        str(repeats + 1), "--out", out], env=env, check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE, timeout=600)
        ps = json.load(open(out))["problems"]
        os.unlink(out)
    except Exception as e:  # noqa: BLE001
        err = getattr(e, "stderr", b"") or b""
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]} {err[-300:].decode(errors='replace')}"}
    warm, timed = ps[0], ps[1:]
    alls = sorted(r["timings"]["All"] for r in timed)
    med = timed[[r["timings"]["All"] for r in timed].index(alls[len(alls) // 2])]
    tm = med["timings"]
    return {"impl": "reference ChASEGPU (cuBLAS / cuSOLVER / cuRAND), unmodified, same GPU, host buffers",
            "workload": f"{t} N={N} nev={nev} nex={nex} uniform spectrum, dense (driver's own reflectors)",
            "protocol": f"one process: 1 warm-up solve + {len(timed)} timed solves, median reported; "
                        f"OPENBLAS/OMP threads = {threads}",
            "time_to_solution_s": tm["All"], "time_to_solution_all_s": [r["timings"]["All"] for r in timed],
            "time_to_solution_min_s": alls[0], "time_to_solution_max_s": alls[-1],
            "warmup_solve_s": warm["timings"]["All"],
            "iterations": med["iterations"], "filtered_vecs": med["filtered_vecs"],
            "value": med["gflop_filter"] / tm["All"] / 1e3, "unit": "TFLOP/s",
            "filter_phase_tflops": med["gflop_filter"] / tm["Filter"] / 1e3,
            "phases_s": {k.lower(): tm[k] for k in ("InitVecs", "Lanczos", "Filter", "QR", "RR", "Resid")},
            "max_resid": max(med["resid"][:nev]),
            "note": "start vectors from cuRAND, so the iteration count may differ from the parity-mode runs"}


def hemm_traffic(key):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum of an ncu capture,
    committed under profiles/ and indexed by profiles/hemm_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "hemm_traffic.json"))).get(key)
    except Exception:  # noqa: BLE001
        return None


def sample_text(r):
    return (f"one full solve of the same generator scaled to N={r['N']}, nev={r['nev']}, nex={r['nex']} "
            f"({r['iterations']} iterations, {r['filtered_vecs']} filtered vectors, {r['t_all']:.1f} s) by the unmodified "
            f"reference ChASECPU + OpenBLAS 0.3.15 on {r['threads']} threads")


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the whole run (1 warm-up + K steps) is sized to ~5 minutes: every step solves the same bounded sample
    budget = min(max(300.0 / (max(a.steps, 1) + 1), 5.0), 120.0)
    for _ in range(min(a.warmup, 1)):  # one untimed warm-up is enough for a CPU solver (page-in, thread pool)
        run_reference_cpu(a.workload, a.ref_n, budget)
    rs = [run_reference_cpu(a.workload, a.ref_n, budget) for _ in range(max(a.steps, 1))]
    if "error" in rs[0]:
        print(json.dumps({"impl": "reference", "unavailable": rs[0]["error"]}))
        return
    t_all = sum(r["t_all"] for r in rs)
    v = sum(r["gflop_filter"] for r in rs) / t_all / 1e3
    t, N, nev, nex = WORKLOADS[a.workload]
    line = {
        "impl": "reference", "metric": "filter_hemm_tflops_per_time_to_solution", "value": v, "unit": "TFLOP/s",
        "n_gpus": a.gpus, "steps": len(rs), "warmup": min(a.warmup, 1), "ms_per_step": 1e3 * t_all / len(rs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if t == "d" else "c128",
        "data": "synthetic",
        "config": {"workload": f"{a.workload}: {t} N={N} nev={nev} nex={nex} uniform spectrum, dense Q diag Q^H",
                   "sample": sample_text(rs[0])},
        "cpu_baseline": {"value": v, "unit": "TFLOP/s", "cores": rs[0]["threads"], "kind": "reference",
                         "sample": sample_text(rs[0])},
        "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "filter_phase_tflops": sum(r["gflop_filter"] for r in rs) / sum(r["t_filter"] for r in rs) / 1e3,
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def make_matrix(t, N, device):
    """Dense A = Q diag(lambda) Q^H on the device (torch is plumbing: input synthesis only).  Same construction as
    oracle.chase_oracle.dense_from_spectrum (seed 7, 3 reflectors)."""
    import numpy as np
    import torch

    dt = torch.float64 if t == "d" else torch.complex128
    lam = 100.0 * (1e-4 + np.arange(N) * (1.0 - 1e-4) / N)
    rng = np.random.default_rng(7)
    A = torch.zeros((N, N), dtype=dt, device=device)
    A.diagonal().copy_(torch.from_numpy(lam).to(device).to(dt))
    for _ in range(3):
        v = rng.standard_normal(N)
        if t == "z":
            v = v + 1j * rng.standard_normal(N)
        v = torch.from_numpy(v / np.linalg.norm(v)).to(device).to(dt)
        w = A @ v
        s = torch.vdot(v, w)
        A.sub_(2 * torch.outer(v, w.conj()))
        A.sub_(2 * torch.outer(w, v.conj()))
        A.add_(4 * s * torch.outer(v, v.conj()))
    A = 0.5 * (A + A.conj().T)
    return A, lam


def ours(a):
    import numpy as np
    import torch

    if a.gpus != 1 or int(os.environ.get("WORLD_SIZE", "1")) != 1:
        from chase_b200 import bench_dist  # distributed arm lives with the distributed backend

        return bench_dist.run(a)

    import chase_b200

    L = chase_b200.lib()
    torch.cuda.set_device(0)
    t, N, nev, nex = WORKLOADS[a.workload]
    m = nev + nex
    f = 1 if t == "d" else 4
    ndt = np.float64 if t == "d" else np.complex128

    A, lam = make_matrix(t, N, "cuda")
    # host copies in pinned memory: H column-major (A^T row-major == A column-major), V, as a C caller would own them
    Hh = torch.empty((N, N), dtype=A.dtype, pin_memory=True)
    Hh.copy_(A.T.contiguous())
    del A
    torch.cuda.empty_cache()
    Vh = torch.zeros((m, N), dtype=Hh.dtype, pin_memory=True)
    H = Hh.numpy().T  # F-contiguous view
    V = Vh.numpy().T
    assert H.flags.f_contiguous and V.flags.f_contiguous

    peak = max(L.chase_b200_dmma_peak(40000, None) for _ in range(3)) / 1e12  # TFLOP/s, measured live

    def flag(name, v):
        getattr(L, name)(ctypes.byref(ctypes.c_int(v)))

    solver = chase_b200.ChASE(H, nev, nex, V=V)
    tol = 1e-10

    def check(res):
        rel = float(np.max(np.abs(res.ritzv[:nev] - lam[:nev]) / lam[:nev]))
        assert rel < 1e-10, f"eigenvalues off: {rel}"
        assert float(res.resid[:nev].max()) < 100 * tol
        return rel

    def sync():
        torch.cuda.synchronize()
        L.chase_b200_device_sync()

    def timed(nsteps):
        """nsteps solves bracketed by synchronize on both sides, CUDA events on the current stream."""
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.chase_b200_launch_count()
        e0.record()
        rs = [solver.solve(deg=20, tol=tol, copy=False) for _ in range(nsteps)]
        e1.record()
        sync()
        return rs, e0.elapsed_time(e1) * 1e-3, L.chase_b200_launch_count() - l0

    if a.profile:
        flag("chase_b200_set_device_rng_", 1)
        solver.solve(deg=20, tol=tol, copy=False)
        flag("chase_b200_set_matrix_resident_", 1)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        r = solver.solve(deg=20, tol=tol, copy=False)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        print(json.dumps({"profile_run": True, "iterations": r.iterations, "phases_s": {k: r.stats[k] for k in r.stats if k.startswith("t_")}}))
        solver.finalize()
        return

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------
    flag("chase_b200_set_device_rng_", 1)
    flag("chase_b200_set_matrix_resident_", 0)
    solver.solve(deg=20, tol=tol, copy=False)  # first warm-up also places H in HBM
    flag("chase_b200_set_matrix_resident_", 1)
    for _ in range(max(a.warmup - 1, 0)):
        solver.solve(deg=20, tol=tol, copy=False)
    L.chase_b200_hemm_profile_enable(1)
    ck = Clocks(0)
    ck.start()
    rs, secs, launches = timed(a.steps)
    clocks = ck.stop()
    hp = (ctypes.c_double * 4)()
    L.chase_b200_hemm_profile_read(hp)
    L.chase_b200_hemm_profile_enable(0)
    rel = max(check(r) for r in rs)
    flop_filter = sum(r.stats["gflop_filter"] for r in rs) * 1e9
    value = flop_filter / secs / 1e12
    st = rs[-1].stats

    # ---- e2e: the reference C interface on host buffers ----------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        flag("chase_b200_set_device_rng_", 1)
        flag("chase_b200_set_matrix_resident_", 0)
        solver.solve(deg=20, tol=tol, copy=False)
        rs2, secs2, _ = timed(a.steps)
        rel = max(rel, max(check(r) for r in rs2))
        es = H.itemsize
        e2e = {"value": sum(r.stats["gflop_filter"] for r in rs2) * 1e9 / secs2 / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": N * N * es, "d2h_bytes_per_step": N * m * es + 2 * m * 8,
               "time_to_solution_s": secs2 / a.steps, "iterations": rs2[-1].iterations,
               "filtered_vecs": rs2[-1].filtered_vecs,
               "start_vectors": "device Philox RNG, regenerated inside every timed step (the reference GPU backend "
                                "regenerates with cuRAND every solve); nothing is cached between steps"}
    # ---- extra: the mixed-precision filter (reference option ENABLE_MIXED_PRECISION, off by default there and here) ----
    mixed = None
    if not a.no_extras:
        flag("chase_b200_set_device_rng_", 1)
        flag("chase_b200_set_matrix_resident_", 1)
        flag("chase_b200_set_mixed_precision_", 1)
        try:
            solver.solve(deg=20, tol=tol, copy=False)
            rs3, secs3, _ = timed(min(a.steps, 3))
            L.chase_b200_last_sp_filter_cols_.restype = ctypes.c_double
            mixed = {"what": "same workload with chase_b200_set_mixed_precision_(1): filters run in single precision on the "
                             "tcgen05 kind::tf32 kernel while min residual > 1e-3 (reference: ENABLE_MIXED_PRECISION)",
                     "time_to_solution_s": secs3 / len(rs3), "iterations": rs3[-1].iterations,
                     "filtered_vecs": rs3[-1].filtered_vecs, "sp_filter_cols": L.chase_b200_last_sp_filter_cols_(),
                     "value": sum(r.stats["gflop_filter"] for r in rs3) * 1e9 / secs3 / 1e12, "unit": "TFLOP/s",
                     "phases_s": {k[2:]: rs3[-1].stats[k] for k in ("t_all", "t_filter", "t_qr", "t_rr")},
                     "max_rel_eig_err": max(check(r) for r in rs3)}
        finally:
            flag("chase_b200_set_mixed_precision_", 0)
    solver.finalize()

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------
    traffic = hemm_traffic(a.workload)
    achieved = hp[2] / (hp[1] * 1e-3) / 1e12 if hp[1] > 0 else None
    roofline = {"bound": "tensor", "kernel": "hemm_tma_kernel (FP64 DMMA, TMA-fed)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak if achieved else None, "traffic": traffic,
                "peak_source": "measured live: register-resident DMMA.8x8x4 loop on all SMs (chase_b200_dmma_peak); "
                               "MEASURED_PEAKS.json carries no FP64 figure",
                "launches": int(hp[0]), "kernel_ms_total": hp[1], "kernel_share_of_step": hp[1] * 1e-3 / secs}

    line = {
        "metric": "filter_hemm_tflops_per_time_to_solution", "value": value, "unit": "TFLOP/s", "n_gpus": 1,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * secs / a.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64" if t == "d" else "c128", "data": "synthetic",
        "config": {"workload": f"{a.workload}: {t} N={N} nev={nev} nex={nex} uniform spectrum, dense Q diag Q^H, tol 1e-10 deg 20 opt",
                   "l2": "inputs larger than L2 (A is %.1f GB)" % (N * N * H.itemsize / 1e9),
                   "start_vectors": "device Philox RNG, regenerated every solve (value and e2e)"},
        "time_to_solution_s": secs / a.steps, "iterations": rs[-1].iterations, "filtered_vecs": rs[-1].filtered_vecs,
        "filter_phase_tflops": st["gflop_filter"] / st["t_filter"] / 1e3 if st["t_filter"] > 0 else None,
        "phases_s": {k[2:]: st[k] for k in ("t_all", "t_initvecs", "t_lanczos", "t_filter", "t_qr", "t_rr", "t_resid")},
        "max_rel_eig_err": rel, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "e2e": e2e,
    }
    if not a.no_cpu_baseline:
        r = run_reference_cpu(a.workload, a.ref_n)
        if "error" in r:
            line["cpu_baseline"] = {"value": None, "unit": "TFLOP/s", "cores": 0, "kind": "reference", "sample": r["error"]}
        else:
            line["cpu_baseline"] = {"value": r["tflops_per_solve"], "unit": "TFLOP/s", "cores": r["threads"],
                                    "kind": "reference", "sample": sample_text(r),
                                    "filter_phase_tflops": r["tflops_filter_phase"], "host_cores": os.cpu_count()}
            try:  # the one full-size run of the same binary on this class of box (recorded, not re-run: 460 s)
                fs = json.load(open(os.path.join(ROOT, "profiles", "r2_ref_cpu_c2_full.json")))
                if a.workload == "c2":
                    line["cpu_baseline"]["full_size_recorded"] = {
                        "time_to_solution_s": fs["timings_s"]["All"], "filter_s": fs["timings_s"]["Filter"],
                        "iterations": fs["iterations"], "filtered_vecs": fs["filtered_vecs"],
                        "value": fs["tflops_per_time_to_solution"], "unit": "TFLOP/s", "cores": 16,
                        "source": "profiles/r2_ref_cpu_c2_full.json"}
            except Exception:  # noqa: BLE001
                pass
    if not a.no_gpu_reference:
        g = run_reference_gpu(a.workload, a.gpu_ref_repeats)
        line["gpu_reference"] = g
        if e2e and "time_to_solution_s" in g:
            # same box, same workload, host buffers on both sides: the like-for-like speed-up of the drop-in
            line["e2e_vs_gpu_reference"] = g["time_to_solution_s"] / e2e["time_to_solution_s"]
            line["filter_vs_gpu_reference"] = g["phases_s"]["filter"] / st["t_filter"]
    if mixed:
        line["mixed_precision"] = mixed
    if not a.no_extras and a.workload == "c2":
        # the fixed complex problem of the multi-GPU arm, on the 1x1 grid of the distributed backend
        from chase_b200 import bench_dist
        from chase_b200 import dist as cd

        world = cd.World(0, 1, 0)
        line["complex_fixed"] = bench_dist.extra_arm(world, "z", 24000, 1000, 400, 64, 2, "fixed complex problem")
        world.close()
    print(json.dumps(line))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)
