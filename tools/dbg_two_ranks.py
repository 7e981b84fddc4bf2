import os, socket, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "examples", "bin", sys.argv[1] if len(sys.argv) > 1 else "channel3d")
X, Y, Z = 48, 16, 16
for steps in (1, 2, 3, 60):
    with tempfile.TemporaryDirectory() as t1, tempfile.TemporaryDirectory() as t2:
        r = subprocess.run([exe, str(X), str(Y), str(Z), str(steps), os.path.join(t1, "one")], capture_output=True, text=True, cwd=t1)
        assert r.returncode == 0, r.stdout + r.stderr
        one = np.fromfile(os.path.join(t1, "one.macro"), dtype=np.float64).reshape(4, X, Z, Y)
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
        procs = []
        for rank in range(2):
            env = dict(os.environ, LBMX_RANK=str(rank), LBMX_WORLD_SIZE="2", LBMX_LOCAL_RANK=str(rank), LBMX_MASTER_ADDR="127.0.0.1", LBMX_MASTER_PORT=str(port))
            procs.append(subprocess.Popen([exe, str(X), str(Y), str(Z), str(steps), os.path.join(t2, "two")], env=env, cwd=t2, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=600)[0] for p in procs]
        two = np.concatenate([np.fromfile(os.path.join(t2, f"two.rank{r}.macro"), dtype=np.float64).reshape(4, X // 2, Z, Y) for r in range(2)], axis=1)
        cmap = np.concatenate([np.fromfile(os.path.join(t2, f"two.rank{r}.map"), dtype=np.int16).reshape(X // 2, Z, Y) for r in range(2)], axis=0)
        d = np.abs(one - two)
        print("steps", steps, "nan one/two", np.isnan(one).sum(), np.isnan(two).sum(), "max diff", np.nanmax(d))
        planes = np.where(d.max(axis=(0, 2, 3)) > 0)[0]
        print("  x-planes with differences:", planes.tolist())
        if len(planes):
            i = np.unravel_index(np.nanargmax(d), d.shape)
            print("  worst at (k,x,z,y)=", i, "one", one[i], "two", two[i], "cell type", cmap[i[1:]], "types among differing cells:", np.unique(cmap[d.max(axis=0) > 0]).tolist())
