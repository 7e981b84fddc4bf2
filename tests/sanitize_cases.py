#!/usr/bin/env python
"""Every golden case once through the C ABI, in default and in parity arithmetic, plus a two-batch graph-replay run -- the workload to put
under compute-sanitizer on a GPU box:

    compute-sanitizer --tool memcheck  python tests/sanitize_cases.py [case-name substrings]
    compute-sanitizer --tool initcheck python tests/sanitize_cases.py cum_f64

Results are still compared with the CPU restatement, so a run that the sanitizer slowed down is also a run that was checked.
(Not collected by pytest: no test_ prefix.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_cases as gc  # noqa: E402
import lbm_cases as lc  # noqa: E402
from engine_runner import run_case_engine  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tnl_lbm_b200 import binding as B  # noqa: E402


def main():
    only = sys.argv[1:]
    n = 0
    for case in gc.CASES:
        if only and not any(s in case.name for s in only):
            continue
        ref_df, ref_mac = gc.run_case(case, "port", nthreads=4)
        for flags in (0, B.FLAG_STRICT_ARITH):
            if case.desc.lattice == O.D3Q19 and flags:
                continue
            df, mac, stats = run_case_engine(case, flags=flags)
            if flags:
                assert np.array_equal(df, ref_df), case.name
            else:
                assert lc.rel_err_df(df, ref_df, case.desc) <= (1e-10 if case.desc.precision == O.F64 else 1e-3), case.name
            n += 1
        if case.nsteps >= 20 and case.desc.macro == O.MACRO_DEFAULT:  # graph replay of step pairs + odd batch boundaries
            df, mac, stats = run_case_engine(case, chunk=9)
            n += 1
    print(f"sanitize_cases: {n} engine runs completed")


if __name__ == "__main__":
    main()
