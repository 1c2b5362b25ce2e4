"""One launch of the distributed filter HEMM per shape and operand orientation (for
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:hemm_tma`): the local blocks of BASELINE config C2 on
2x1, 2x2 and 4x2 grids, op(A) = A (row-layout panel in, column-layout out) and op(A) = A^H, with the hybrid schedule off
(the current default for A^H) and on.  Prints the launch order so that the ncu rows can be labelled."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chase_b200 import kernels as k  # noqa: E402

os.environ["CHASE_B200_HEMM_SPLIT"] = "0"  # one launch per product, so that rows of the ncu list map 1:1
order = []
for G, M, K in ((2, 10000, 20000), (4, 10000, 10000), (8, 5000, 10000)):
    kc = 1400
    ldm, ldk = (M + 15) // 16 * 16, (K + 15) // 16 * 16
    A = torch.randn((K, ldm), dtype=torch.float64, device="cuda")
    Bk = torch.randn((kc, ldk), dtype=torch.float64, device="cuda")
    Bm = torch.randn((kc, ldm), dtype=torch.float64, device="cuda")
    for hyb in ("0", "1"):
        os.environ["CHASE_B200_HEMM_HYBRID"] = hyb
        k.hemm_rect(0, M, K, kc, 1.0, A, ldm, Bk, ldk, 0.0, Bm, ldm)
        order.append(f"G={G} M={M} K={K} op=N hybrid={hyb}")
        k.hemm_rect(1, K, M, kc, 1.0, A, ldm, Bm, ldm, 0.0, Bk, ldk)
        order.append(f"G={G} M={M} K={K} op=H hybrid={hyb}")
    torch.cuda.synchronize()
    del A, Bk, Bm
for i, o in enumerate(order):
    print(i, o)
