#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own per-cell code (oracle/_ref/libref_*.so, built by
`make -C oracle ref` from /root/reference/include).  Run in a container that has /root/reference:

    make -C oracle ref && python tests/golden/make_golden.py

Each fixture holds a strided sample of the final distributions and macroscopic fields of one case of
tests/golden_cases.py; manifest.json holds the SHA-256 of the full arrays, so the restatement (bit for bit) and the CUDA engine (within tolerance) can be checked
where the reference is not available."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import golden_cases as gc  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    assert O.available("reference", O.AB) and O.available("reference", O.AA), "build oracle/_ref first (make -C oracle ref)"
    manifest = {}
    for case in gc.CASES:
        df, mac = gc.run_case(case, "reference")
        assert np.isfinite(df).all() and np.isfinite(mac).all(), case.name
        # full arrays are pinned by SHA-256; the stored sample (every `stride`-th cell of every population) keeps the
        # fixture small and is what the CUDA engine is compared with, within tolerance, on the GPU box
        stride = gc.sample_stride(case)
        np.savez_compressed(os.path.join(HERE, case.name + ".npz"), df_sample=gc.sample(df, stride), macro_sample=gc.sample(mac, stride), stride=stride)
        manifest[case.name] = {
            "df_sha256": hashlib.sha256(df.tobytes()).hexdigest(),
            "macro_sha256": hashlib.sha256(mac.tobytes()).hexdigest(),
            "shape": list(df.shape),
            "dtype": str(df.dtype),
            "nsteps": case.nsteps,
        }
        print(f"{case.name:32s} df{df.shape} rho in [{mac[0].min():.6f}, {mac[0].max():.6f}]")
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
