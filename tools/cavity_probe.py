"""Small-lattice probe (BASELINE.json configs[1]: D2Q9 lid-driven cavity 1024^2 fp64): where do the microseconds of a step go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tnl_lbm_b200 import binding as B

def run(N, streaming, coll=B.SRT, steps=2000, macro=B.MACRO_DEFAULT):
    e = B.Engine(lattice=B.D2Q9, coll=coll, eq=B.EQ_STD, streaming=streaming, macro=macro, inflow=B.INFLOW_CONST, precision=B.F64, X=N, Y=N, Z=1)
    m = np.zeros((N, 1, N), dtype=np.int16)
    m[0], m[N - 1] = 1, 1
    m[:, :, 0] = 1
    m[:, :, N - 1] = 2
    e.map_upload(m)
    e.set_equilibrium(1.0, 0.0, 0.0, 0.0)
    e.set_params(lbmViscosity=0.05, inflow_vx=0.1)
    e.step(20)
    e.sync()
    ms = e.step_timed(steps)
    assert not e.has_nan()
    st = e.stats()
    e.close()
    return N * N * steps / (ms * 1e-3) / 1e9, ms / steps * 1e3, st

for N in (512, 1024, 2048, 4096):
    for st in (B.AB, B.AA):
        g, us, s = run(N, st, steps=4000 if N <= 1024 else 1000)
        print(f"cavity {N}^2 {'A-A' if st == B.AA else 'A-B'} SRT fp64: {g:7.2f} GLUPS  {us:7.2f} us/step  ({g * 144 / 1e3:.2f} TB/s algorithmic) bulk cells {s.bulk_cells} boundary cells {s.boundary_cells} launches {s.kernel_launches}", flush=True)
