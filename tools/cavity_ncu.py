import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tnl_lbm_b200 import binding as B
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for st in (B.AB, B.AA):
    e = B.Engine(lattice=B.D2Q9, coll=B.SRT, eq=B.EQ_STD, streaming=st, macro=B.MACRO_DEFAULT, inflow=B.INFLOW_CONST, precision=B.F64, X=N, Y=N, Z=1)
    m = np.zeros((N, 1, N), dtype=np.int16)
    m[0], m[N - 1] = 1, 1
    m[:, :, 0] = 1
    m[:, :, N - 1] = 2
    e.map_upload(m)
    e.set_equilibrium(1.0, 0.0, 0.0, 0.0)
    e.set_params(lbmViscosity=0.05, inflow_vx=0.1)
    e.step(12)
    e.sync()
    e.close()
