"""A/B of the Jacobi round kernels (register-tiled FMA vs DMMA) at Rayleigh-Ritz sizes: one subprocess per variant."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for dm in ("0", "1"):
    for t, sizes in (("d", ["419", "609", "1400"]), ("z", ["300", "700", "1400"])):
        e = dict(os.environ, CHASE_B200_OSJ_DMMA=dm)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "heev_bench.py"), t] + sizes, env=e,
                           capture_output=True, text=True, timeout=600)
        for ln in (r.stdout + r.stderr[-500:]).splitlines():
            print(f"dmma={dm} {ln}", flush=True)
