import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tnl_lbm_b200 import binding as B
N = 512
e = B.Engine(lattice=B.D3Q27, coll=B.CUM, eq=B.EQ_INV_CUM, streaming=B.AA, precision=B.F64, X=N, Y=N, Z=N)
n = N ** 3
pin = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(4)]
for t in pin: t.fill_(0.0)
pin[0].fill_(1.0)
fields_pin = [t.numpy().reshape(N, N, N) for t in pin]
fields_page = [a.copy() for a in fields_pin]
for name, fl in (("pinned", fields_pin), ("pageable", fields_page)):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e.set_equilibrium_field(*fl)
        e.sync(); t1 = time.perf_counter()
        print(f"LBMX_EQ_CHUNK={os.environ.get('LBMX_EQ_CHUNK','default')} {name} rep {rep}: {t1 - t0:.3f} s  ({4 * n * 8 / (t1 - t0) / 1e9:.1f} GB/s)", flush=True)
e.close()
