import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, golden_cases as gc, lbm_cases as lc
from oracle import oracle as O
from tnl_lbm_b200 import binding as B
from engine_runner import run_case_engine
for coll, eq in ((O.CUM, O.EQ_INV_CUM), (O.MRT_LES, O.EQ_STD)):
  for st in (O.AB, O.AA):
    for n in (1, 2, 10, 100):
        d = O.Desc(coll=coll, eq=eq, streaming=st, precision=O.F32, X=8, Y=8, Z=8)
        case = gc.Case("dbg", d, O.Params(lbmViscosity=1e-3, fx=1e-6), lc.map_periodic, n, "smooth")
        df, mac, _ = run_case_engine(case, flags=B.FLAG_STRICT_ARITH)
        rdf, rmac = gc.run_case(case, "port")
        diff = np.abs(df.astype(np.float64) - rdf.astype(np.float64))
        i = np.unravel_index(np.argmax(diff), diff.shape)
        print("coll", coll, "st", st, "steps", n, "max abs %.3e" % diff.max(), "at", i, "ndiff", int((diff > 0).sum()), "of", diff.size,
              "macro diff", np.abs(mac.astype(np.float64) - rmac.astype(np.float64)).max(axis=(1, 2, 3)), flush=True)
