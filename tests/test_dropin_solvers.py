"""Drop-in check of the host mirror (tnl_lbm_b200/host/lbm3d/*.h): solvers written against TNL-LBM's solver-facing interface
-- LBM_CONFIG, State<NSE>, nse.setBoundaryX/..., block.data.*, execute(state) -- compile against it and run on the engine.

* examples/channel3d.cpp : this repository's solver in the reference's style; its result is compared with the CPU oracle.
* sim_NSE/sim_1.cu, sim_2D/sim2d_1.cu : the reference's UNMODIFIED sources (compiled where /root/reference exists; the
  binaries travel to the GPU box): they run to completion and their cuts land in the raw-dump writers.
* checkpoint / restart through State::saveState / loadState (reference variable names, local-storage shapes) is bit-identical
  to an uninterrupted run; the 3-D / cut writers keep the reference's (z, y, x) float ordering."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import golden_cases as gc
import lbm_cases as lc
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "bin")
HAVE_REF = os.path.isdir("/root/reference/sim_NSE")


def test_clients_compile_against_the_host_mirror():
    from tnl_lbm_b200 import build_examples

    built = build_examples.build()
    names = {os.path.basename(b) for b in built}
    assert "channel3d" in names and "abi_minimal" in names  # the latter is compiled as C99 with -pedantic -Werror
    if HAVE_REF:
        assert {"ref_sim_1_ab", "ref_sim_1_aa", "ref_sim_2_ab", "ref_sim_2_aa", "ref_sim_3_ab", "ref_sim_3_aa", "ref_sim2d_1_ab", "ref_sim2d_1_aa", "ref_sim2d_2_ab", "ref_sim2d_2_aa", "ref_sim2d_3_ab"} <= names


def test_custom_device_traits_are_rejected_at_compile_time():
    """A user-defined MACRO is device-side code in the reference and cannot cross the C ABI: unless it is one the engine has built in
    (D2Q9_MACRO_WithMean of sim_2D/sim2d_2.cu:53-104, recognised by its channel list), the mirror must refuse it when the solver is
    compiled, not at run time."""
    src = r'''
#include "lbm3d/core.h"
template <typename TRAITS> struct MyMacro { enum { e_rho, N }; };
using T = TraitsDP;
using COLL = D3Q27_CUM<T>;
using NSE = LBM_CONFIG<T, D3Q27_KernelStruct, NSE_Data_ConstInflow<T>, COLL, typename COLL::EQ, D3Q27_STREAMING<T>, D3Q27_BC_All, MyMacro<T>>;
int main() { return NSE::lbmx_macro; }
'''
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, "bad.cpp")
        open(f, "w").write(src)
        r = subprocess.run(["g++", "-std=c++17", f"-I{ROOT}/tnl_lbm_b200/host", f"-I{ROOT}/include", "-fsyntax-only", f], capture_output=True, text=True)
    assert r.returncode != 0 and "lbmx_macro" in r.stderr


@pytest.mark.parametrize("switch,expr,const", [
    ("", "D3Q27_CUM<TraitsDP>", "LBMX_COLL_CUM"),
    ("-DUSE_GEIER_CUM_2017", "D3Q27_CUM<TraitsDP>", "LBMX_COLL_CUM_2017"),
    ("-DUSE_GEIER_CUM_ANTIALIAS", "D3Q27_CUM<TraitsDP>", "LBMX_COLL_CUM_ANTIALIAS"),
    ("-DUSE_GEIER_CUM_2017 -DUSE_GEIER_CUM_ANTIALIAS", "D3Q27_CUM<TraitsSP>", "LBMX_COLL_CUM_2017_ANTIALIAS"),
    ("-DUSE_HIGH_PRECISION_RHO", "D3Q27_CUM<TraitsDP>", "LBMX_COLL_CUM_HP_RHO"),
    ("-DUSE_GALILEAN_CORRECTION", "D3Q27_BGK<TraitsDP>", "LBMX_COLL_BGK_GALILEAN"),
    ("", "D3Q27_BGK<TraitsSP>", "LBMX_COLL_BGK"),
])
def test_the_reference_build_switches_select_kernel_families(switch, expr, const):
    """defs.h:252-255: a solver compiled with one of the reference's arithmetic switches gets the kernel family that is that build of the
    operator (the engine has no #ifdef of its own: the switch becomes the collision kind of the lbmx_desc)."""
    src = f'#include "lbm3d/core.h"\nstatic_assert({expr}::lbmx_coll == {const}, "wrong kernel family");\nint main() {{ return 0; }}\n'
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, "tag.cpp")
        open(f, "w").write(src)
        r = subprocess.run(["g++", "-std=c++17", *switch.split(), f"-I{ROOT}/tnl_lbm_b200/host", f"-I{ROOT}/include", "-fsyntax-only", f], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_channel3d_matches_oracle(prec):
    exe = os.path.join(BIN, "channel3d")
    if not os.path.exists(exe):
        pytest.skip("examples/bin/channel3d not built")
    X, Y, Z, steps = 48, 20, 20, 60
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "run")
        r = subprocess.run([exe, str(X), str(Y), str(Z), str(steps), out] + (["f32"] if prec == "f32" else []), capture_output=True, text=True, timeout=300, cwd=tmp)
        assert r.returncode == 0, r.stdout + r.stderr
        m = re.search(r"iterations=(\d+) mass=(\S+) lbmViscosity=(\S+) inflow_vx=(\S+)", r.stdout)
        assert m and int(m.group(1)) == steps, r.stdout
        nu, vin = float(m.group(3)), float(m.group(4))
        dt = np.float64 if prec == "f64" else np.float32
        cmap = np.fromfile(out + ".map", dtype=np.int16).reshape(X, Z, Y)
        mac = np.fromfile(out + ".macro", dtype=dt).reshape(4, X, Z, Y)
    assert "GLUPS=" in r.stdout and "physFinalTime reached" in r.stdout
    # the map painted through nse.setBoundary*/setMap: bit-exact against the same painting order in numpy
    g = lc.G3
    ref_map = np.full((X, Z, Y), g["FLUID"], dtype=np.int16)
    ref_map[0], ref_map[X - 1] = g["INFLOW_LEFT"], g["OUTFLOW_RIGHT"]
    ref_map[:, 1, :] = ref_map[:, Z - 2, :] = g["WALL"]
    ref_map[:, :, 1] = ref_map[:, :, Y - 2] = g["WALL"]
    ref_map[:, 0, :] = ref_map[:, Z - 1, :] = g["NOTHING"]
    ref_map[:, :, 0] = ref_map[:, :, Y - 1] = g["NOTHING"]
    cx, width = X // 5, Z // 10
    for px in range(cx, cx + width + 1):
        for pz in range(1, Z - 1):
            for py in range(1, Y - 1):
                if not (Z * 4 // 10 <= pz <= Z * 6 // 10 and Y * 4 // 10 <= py <= Y * 6 // 10):
                    ref_map[px, pz, py] = g["WALL"]
    assert np.array_equal(cmap, ref_map)
    # the flow field against the CPU oracle driven the way State::SimInit / SimUpdate drive the reference
    d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, precision=O.F64 if prec == "f64" else O.F32, X=X, Y=Y, Z=Z)
    orc = O.Oracle(d, "port")
    a = d.new_df()
    orc.set_equilibrium(a, 1.0, 0.0, 0.0, 0.0)
    b = a.copy()
    ref = d.new_macro()
    p = O.Params(lbmViscosity=nu, inflow_vx=vin)
    orc.step(p, a, b, ref, cmap, 0, steps, os.cpu_count() or 4)
    tol = 1e-12 if prec == "f64" else 1e-5
    for lo, hi, label in lc.macro_groups(d):
        e = lc.rel_err(mac[lo:hi], ref[lo:hi])
        assert e <= tol, f"{label}: {e:.3e}"
    assert abs(float(m.group(2)) - float(mac[0].astype(np.float64).sum())) < 1e-3 * mac[0].size  # probe1() ran on a host copy of rho


@pytest.mark.gpu
@pytest.mark.parametrize("exe_name,arg", [("ref_sim_1_ab", "1"), ("ref_sim_1_aa", "1"), ("ref_sim2d_1_ab", "1"), ("ref_sim_3_ab", None)])  # sim_3 paints GEO_OUTFLOW_RIGHT_INTERP, which exists for A-B only (streaming_AA.h has no streamingInterpRight)
def test_unmodified_reference_solvers_run(exe_name, arg):
    exe = os.path.join(BIN, exe_name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe_name} not built (needs /root/reference at build time)")
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([exe] + ([arg] if arg else []), capture_output=True, text=True, timeout=600, cwd=tmp)
        dumps = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp) for f in fs if f.endswith(".txt")]
        flags = [f for dp, _, fs in os.walk(tmp) for f in fs if f.startswith("flag.")]
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert len(dumps) > 10, "the solver's cuts did not reach the raw-dump writers"
    assert "flag.finished" in flags or "flag.terminated" in flags
    assert "GLUPS=" in r.stdout
    assert "physFinalTime reached" in r.stdout or "terminate flag triggered" in r.stdout
    iters = [int(x) for x in re.findall(r"iter=(\d+)", r.stdout)]
    assert iters and iters[-1] > 1000
    print(exe_name, "last lines:", r.stdout.strip().splitlines()[-3:])


def _read_dump(base):
    """RawWriter output: <base>.txt = index (name dtype count dofs byte_offset), <base>.bin = payloads."""
    raw = open(base + ".bin", "rb").read()
    out, header = {}, None
    for line in open(base + ".txt"):
        t = line.split()
        if line.startswith("#"):
            header = {t[i]: [int(v) for v in t[i + 1:i + 4]] for i in (1, 5, 9)}
        elif t[1] == "scalar":
            out[t[0]] = float(t[2])
        else:
            out[t[0]] = np.frombuffer(raw, dtype=np.dtype(t[1]), count=int(t[2]), offset=int(t[4]))
    return header, out


@pytest.mark.gpu
@pytest.mark.parametrize("exe_name,halt", [("channel3d", 30), ("channel3d", 31), ("channel3d_aa", 31)])
def test_checkpoint_restart_is_bit_identical(exe_name, halt):
    """Stop after `halt` steps, saveState(), start again in the same results directory: loadState() restores map, DFs (ghost planes
    and A-A parity included), macro, iteration and counters, and the run ends bit-identical to an uninterrupted one."""
    exe = os.path.join(BIN, exe_name)
    if not os.path.exists(exe):
        pytest.skip(f"examples/bin/{exe_name} not built")
    X, Y, Z, steps = 40, 16, 16, 60
    args = [exe, str(X), str(Y), str(Z), str(steps)]
    with tempfile.TemporaryDirectory() as t1, tempfile.TemporaryDirectory() as t2:
        r = subprocess.run(args + [os.path.join(t1, "full")], capture_output=True, text=True, timeout=300, cwd=t1)
        assert r.returncode == 0, r.stdout + r.stderr
        full = np.fromfile(os.path.join(t1, "full.macro"), dtype=np.float64)
        r = subprocess.run(args + [os.path.join(t2, "part"), f"halt={halt}"], capture_output=True, text=True, timeout=300, cwd=t2)
        assert r.returncode == 0 and f"iterations={halt} " in r.stdout, r.stdout + r.stderr
        ck = os.path.join(t2, "results_channel3d", "checkpoint.bp")
        assert os.path.exists(os.path.join(t2, "results_channel3d", "flag.loadstate"))
        attrs = dict(line.rstrip("\n").split("\t") for line in open(os.path.join(ck, "attributes.txt")))
        assert int(attrs["LBM_iterations"]) == halt and int(attrs["LBM_total_blocks"]) == 1
        variables = {t[0]: (t[1], int(t[2])) for t in (line.split() for line in open(os.path.join(ck, "variables.txt")))}
        n = X * Y * Z
        dfmax = 1 if exe_name.endswith("_aa") else 2
        expect = {"LBM_map_block_0": ("int16", n), "LBM_macro_block_0": ("float64", 4 * n)}
        expect.update({f"LBM_df_{k}_block_0": ("float64", 27 * n) for k in range(dfmax)})
        assert variables == expect
        assert os.path.getsize(os.path.join(ck, "LBM_df_0_block_0.bin")) == 27 * n * 8
        r = subprocess.run(args + [os.path.join(t2, "resumed")], capture_output=True, text=True, timeout=300, cwd=t2)
        assert r.returncode == 0 and "Loading data from checkpoint" in r.stdout and f"iterations={steps} " in r.stdout, r.stdout + r.stderr
        resumed = np.fromfile(os.path.join(t2, "resumed.macro"), dtype=np.float64)
        assert np.array_equal(np.fromfile(os.path.join(t2, "resumed.map"), dtype=np.int16), np.fromfile(os.path.join(t1, "full.map"), dtype=np.int16))
    assert np.array_equal(full, resumed)


@pytest.mark.gpu
def test_raw_dump_writers_keep_reference_ordering():
    """writeVTKs_3D / _3Dcut / _2D feed the solver's outputData() hook in the reference's (z, y, x) order with its variable names
    (lbm_block.hpp:800-1110): "wall", scalar "<id>", vector "<id>X/Y/Z", "TIME"."""
    exe = os.path.join(BIN, "channel3d")
    if not os.path.exists(exe):
        pytest.skip("examples/bin/channel3d not built")
    X, Y, Z, steps = 40, 16, 16, 60
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([exe, str(X), str(Y), str(Z), str(steps), os.path.join(tmp, "run"), "dump"], capture_output=True, text=True, timeout=300, cwd=tmp)
        assert r.returncode == 0, r.stdout + r.stderr
        mac = np.fromfile(os.path.join(tmp, "run.macro"), dtype=np.float64).reshape(4, X, Z, Y)
        cmap = np.fromfile(os.path.join(tmp, "run.map"), dtype=np.int16).reshape(X, Z, Y)
        res = os.path.join(tmp, "results_channel3d")
        h3, d3 = _read_dump(os.path.join(res, "output_3D.2"))  # cycles 0, 1, 2 at steps 0, 30, 60
        hb, db = _read_dump(os.path.join(res, "output_3Dcut_box.2"))
        hx, dx = _read_dump(os.path.join(res, "output_2D_cutsX", "cut_X.3"))  # steps 0, 20, 40, 60
        hz, dz = _read_dump(os.path.join(res, "output_2D_cut_Z.3"))
        line_x = np.loadtxt(os.path.join(res, "probes1D", "centre_line_rank000_000002"))  # cycles at steps 0, 30, 60
        line_z = np.loadtxt(os.path.join(res, "probes1D", "profile_z_rank000_000002"))
        header = open(os.path.join(res, "probes1D", "centre_line_rank000_000002")).read().splitlines()[:2]
    zyx = lambda a: a.transpose(1, 2, 0)  # [x, z, y] -> [z, y, x]
    assert h3["global"] == [X, Y, Z] and sorted(d3) == ["TIME", "lbm_density", "velocityX", "velocityY", "velocityZ", "wall"]
    assert np.array_equal(d3["wall"].reshape(Z, Y, X), zyx(cmap).astype(np.int32))
    assert np.array_equal(d3["lbm_density"].reshape(Z, Y, X), zyx(mac[0]).astype(np.float32))
    for k, c in enumerate("XYZ"):
        assert np.array_equal(d3["velocity" + c].reshape(Z, Y, X), zyx(mac[1 + k]).astype(np.float32))
    assert d3["TIME"] > 0
    # box [X/4, X/4 + X/2) x ... with stride 2
    sl = lambda n: slice(n // 4, n // 4 + n // 2, 2)
    assert hb["global"] == [X // 4, Y // 4, Z // 4]
    assert np.array_equal(db["lbm_density"].reshape(Z // 4, Y // 4, X // 4), zyx(mac[0][sl(X), sl(Z), sl(Y)]).astype(np.float32))
    assert hx["global"] == [1, Y, Z] and np.array_equal(dx["velocityX"].reshape(Z, Y), mac[1][X // 2].astype(np.float32))
    assert hz["global"] == [X, Y, 1] and np.array_equal(dz["lbm_density"].reshape(Y, X), mac[0][:, Z // 2, :].T.astype(np.float32))
    # 1-D cuts: text tables "#time", "#1:x  2:lbm_density  3:velocity[0] ..." (state.hpp:225-238), %e formatting
    assert header[0].startswith("#time ") and header[1].split("\t") == ["#1:x", "2:lbm_density", "3:velocity[0]", "4:velocity[1]", "5:velocity[2]"]
    assert line_x.shape == (X, 5) and line_z.shape == (Z, 5)
    assert np.allclose(line_x[:, 1], mac[0][:, Z // 2, Y // 2], rtol=1e-6) and np.allclose(line_x[:, 2], mac[1][:, Z // 2, Y // 2], rtol=1e-6, atol=1e-12)
    assert np.allclose(line_z[:, 1], mac[0][X // 2, :, Y // 2], rtol=1e-6) and np.allclose(line_z[:, 4], mac[3][X // 2, :, Y // 2], rtol=1e-6, atol=1e-12)


def test_checkpoint_manager_round_trip_on_cpu():
    """CheckpointManager alone (no GPU): attributes and raw variables written under the reference's naming scheme read back equal;
    a wrong shape is refused."""
    src = r'''
#include "lbm3d/core.h"
struct Blk { int id = 3; };
int main(int, char** argv)
{
	const std::string dir = argv[1];
	std::vector<double> a = {1.5, -2.25, 3.0e-300}, b(3);
	std::vector<short> m = {0, 7, -1}, m2(3);
	int it = 12345, it2 = 0;
	float period = 0.125f, period2 = 0;
	Blk blk;
	CheckpointManager ck;
	ck.start(dir, adios2::Mode::Write);
	ck.saveLoadAttribute("LBM_iterations", it);
	ck.saveLoadAttribute("State_counter_0_period", period);
	ck.saveLoadVariable("LBM_df_0", blk, a);
	ck.saveLoadLocalArray("LBM_map", 0, m);
	ck.finalize();
	ck.start(dir, adios2::Mode::Read);
	ck.saveLoadAttribute("LBM_iterations", it2);
	ck.saveLoadAttribute("State_counter_0_period", period2);
	ck.saveLoadVariable("LBM_df_0", blk, b);
	ck.saveLoadLocalArray("LBM_map", 0, m2);
	std::vector<double> wrong(4);
	bool refused = false;
	try { ck.saveLoadVariable("LBM_df_0", blk, wrong); } catch (const std::runtime_error&) { refused = true; }
	ck.finalize();
	return (it2 == it && period2 == period && a == b && m == m2 && refused) ? 0 : 1;
}
'''
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, "ck.cpp")
        open(f, "w").write(src)
        exe = os.path.join(tmp, "ck")
        r = subprocess.run(["g++", "-std=c++17", f"-I{ROOT}/tnl_lbm_b200/host", f"-I{ROOT}/include", f"-I{ROOT}/tests/solver_shims", f, "-o", exe,
                            f"-L{ROOT}/tnl_lbm_b200", "-llbmx", f"-Wl,-rpath,{ROOT}/tnl_lbm_b200"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        d = os.path.join(tmp, "results_x", "checkpoint.bp")
        r = subprocess.run([exe, d], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert sorted(os.listdir(d)) == ["LBM_df_0_block_3.bin", "LBM_map_rank_0.bin", "attributes.txt", "variables.txt"]
        assert np.array_equal(np.fromfile(os.path.join(d, "LBM_df_0_block_3.bin")), [1.5, -2.25, 3.0e-300])


def test_process_group_collectives_without_mpi():
    """lbmx_host::World (the stand-in for the reference's MPI ranks): 3 processes found through LBMX_RANK / LBMX_WORLD_SIZE, TCP
    rendezvous on rank 0, and the TNL::MPI calls the reference's solvers and State use -- reduce (SUM, LOR, MAX), Bcast, Barrier."""
    src = r'''
#include "lbm3d/core.h"
int main(int argc, char** argv)
{
	TNLMPI_INIT mpi(argc, argv);
	const int r = TNL::MPI::GetRank(MPI_COMM_WORLD), n = TNL::MPI::GetSize(MPI_COMM_WORLD);
	const double sum = TNL::MPI::reduce(1.5 * (r + 1), MPI_SUM, MPI_COMM_WORLD);
	const bool any = TNL::MPI::reduce(r == 2, MPI_LOR, MPI_COMM_WORLD), all = TNL::MPI::reduce(r == 2, MPI_LAND, MPI_COMM_WORLD);
	const float mx = TNL::MPI::reduce((float) r, MPI_MAX, MPI_COMM_WORLD);
	int token[2] = {r == 1 ? 4711 : 0, r == 1 ? -3 : 0};
	TNL::MPI::Bcast(token, 2, 1, MPI_COMM_WORLD);
	TNL::MPI::Barrier(MPI_COMM_WORLD);
	const bool ok = n == 3 && sum == 1.5 * 6 && any && ! all && mx == 2.0f && token[0] == 4711 && token[1] == -3;
	std::printf("rank %d of %d ok=%d\\n", r, n, (int) ok);
	return ok ? 0 : 1;
}
'''
    import socket

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, "pg.cpp")
        open(f, "w").write(src)
        exe = os.path.join(tmp, "pg")
        r = subprocess.run(["g++", "-std=c++17", f"-I{ROOT}/tnl_lbm_b200/host", f"-I{ROOT}/include", f"-I{ROOT}/tests/solver_shims", f, "-o", exe,
                            f"-L{ROOT}/tnl_lbm_b200", "-llbmx", f"-Wl,-rpath,{ROOT}/tnl_lbm_b200"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        procs = []
        for rank in (2, 0, 1):  # start order must not matter
            env = dict(os.environ, LBMX_RANK=str(rank), LBMX_WORLD_SIZE="3", LBMX_MASTER_ADDR="127.0.0.1", LBMX_MASTER_PORT=str(port))
            procs.append(subprocess.Popen([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=120)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), outs


def _gpu_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("exe_name", ["channel3d", "channel3d_aa"])
def test_two_ranks_through_the_host_mirror_equal_one(exe_name):
    """The same solver binary started as two processes (one per GPU; ranks from the environment, NCCL id over the process group):
    the two x-slabs assembled equal the single-process result bit for bit (A-B), and a stop / saveState / restart on two ranks
    reproduces the uninterrupted two-rank run."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    exe = os.path.join(BIN, exe_name)
    if not os.path.exists(exe):
        pytest.skip(f"examples/bin/{exe_name} not built")
    import socket

    X, Y, Z, steps = 48, 16, 16, 60

    def launch(tmp, out, extra):
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        procs = []
        for rank in range(2):
            env = dict(os.environ, LBMX_RANK=str(rank), LBMX_WORLD_SIZE="2", LBMX_LOCAL_RANK=str(rank), LBMX_MASTER_ADDR="127.0.0.1", LBMX_MASTER_PORT=str(port))
            procs.append(subprocess.Popen([exe, str(X), str(Y), str(Z), str(steps), os.path.join(tmp, out)] + extra, env=env, cwd=tmp,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(outs)
        return outs

    def assemble(tmp, out):
        parts = [np.fromfile(os.path.join(tmp, f"{out}.rank{r}.macro"), dtype=np.float64).reshape(4, X // 2, Z, Y) for r in range(2)]
        return np.concatenate(parts, axis=1)

    with tempfile.TemporaryDirectory() as t1, tempfile.TemporaryDirectory() as t2, tempfile.TemporaryDirectory() as t3:
        outs = launch(t2, "two", [])
        assert any("ranks=2" in o for o in outs), outs
        two = assemble(t2, "two")
        if not exe_name.endswith("_aa"):
            r = subprocess.run([exe, str(X), str(Y), str(Z), str(steps), os.path.join(t1, "one")], capture_output=True, text=True, timeout=300, cwd=t1)
            assert r.returncode == 0, r.stdout + r.stderr
            one = np.fromfile(os.path.join(t1, "one.macro"), dtype=np.float64).reshape(4, X, Z, Y)
            assert np.array_equal(two, one), f"max |two - one| = {np.abs(two - one).max():.3e} in x-planes {np.where(np.abs(two - one).max(axis=(0, 2, 3)) > 0)[0].tolist()}"
        # (A-A on non-periodic faces: a single process has no ghost planes for the x +- 1 accesses of its face cells -- out of bounds in the
        # reference, kernels.h:30-38 -- so only runs with the same decomposition are comparable there)
        assert np.isfinite(two).all() and abs(two[0].mean() - 1.0) < 0.05
        launch(t3, "part", ["halt=31"])
        ck = os.path.join(t3, "results_channel3d", "checkpoint.bp")
        assert sorted(f for f in os.listdir(ck) if f.endswith(".txt")) == ["attributes.txt", "variables_rank_0.txt", "variables_rank_1.txt"]
        outs = launch(t3, "resumed", [])
        assert any("Loading data from checkpoint" in o for o in outs) and any(f"iterations={steps} " in o for o in outs), outs
        resumed = assemble(t3, "resumed")
        assert np.array_equal(resumed, two), f"max |resumed - two| = {np.abs(resumed - two).max():.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("exe_name,args", [("ref_sim_2_ab", ["--use-forcing"]), ("ref_sim_2_aa", ["--use-forcing"]), ("ref_sim_2_ab", []), ("ref_sim_2_ab", ["--precision", "double"])])
def test_reference_duct_verification_solver(exe_name, args):
    """sim_NSE/sim_2.cu, unmodified: the reference's own verification case (square duct against the analytical series solution,
    sim_2.cu:63-88).  With --use-forcing the flow spins up from rest (periodic, body force); without it the analytical profile is the
    inflow (NSE_Data_XProfileInflow, a DATA class defined in the solver itself: its host-side vx_profile array is uploaded by the
    mirror).  The solver prints its L1 / L2 error against the analytical solution at every probe: it must be finite and shrink."""
    exe = os.path.join(BIN, exe_name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe_name} not built (needs /root/reference at build time)")
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([exe, "--min-resolution", "1", "--max-resolution", "1", "--final-time", "40"] + args, capture_output=True, text=True, timeout=900, cwd=tmp)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    errs = [(float(a), float(b)) for a, b in re.findall(r"l1error_phys=(\S+) l2error_phys=(\S+)", r.stdout)]
    assert len(errs) >= 10, r.stdout[-2000:]
    l1 = np.array([e[0] for e in errs])
    assert np.isfinite(l1).all() and (l1 > 0).all()
    if "--use-forcing" in args:
        # spin-up from rest towards the analytical profile: the viscous time H^2 / nu is ~4000 s, so 40 s only start the approach --
        # the error must fall monotonically
        assert np.all(np.diff(l1[1:]) <= 0) and l1[-1] < 0.97 * l1[1], f"error against the analytical profile does not shrink: {l1[1]:.3e} -> {l1[-1]:.3e}"
    else:
        assert l1[-1] < 0.5 * l1[1] or l1[-1] < 1e-9, f"{l1[1]:.3e} -> {l1[-1]:.3e}"  # the inflow already carries the analytical profile
    print(exe_name, args, "l1 error:", l1[1], "->", l1[-1], "| last line:", r.stdout.strip().splitlines()[-1][:120])


@pytest.mark.gpu
@pytest.mark.parametrize("exe_name", ["channel3d", "channel3d_aa"])
def test_deferred_batches_are_invisible_in_the_mirror(exe_name):
    """execute() records the steps between two host-observable points and hands them to the engine as one lbmx_step(n) batch
    (LBM_BLOCK::pending in lbmx_host.h; the reference copies to the host only on the cadence of its counters, state.hpp:1134-1142).
    Everything the solver writes must be bit-identical to LBMX_HOST_BATCH=0 (one launch per SimUpdate, as the reference does)."""
    exe = os.path.join(BIN, exe_name)
    if not os.path.exists(exe):
        pytest.skip(f"examples/bin/{exe_name} not built")
    X, Y, Z, steps = 40, 16, 16, 61
    out = {}
    for mode in ("0", "1"):
        with tempfile.TemporaryDirectory() as tmp:
            r = subprocess.run([exe, str(X), str(Y), str(Z), str(steps), os.path.join(tmp, "run"), "dump"], capture_output=True, text=True, timeout=300, cwd=tmp,
                               env=dict(os.environ, LBMX_HOST_BATCH=mode))
            assert r.returncode == 0, r.stdout + r.stderr
            m = re.search(r"lbmx: (\d+) steps enqueued in (\d+) batches", r.stdout)
            assert m, r.stdout[-2000:]
            files = {}
            for dp, _, fs in os.walk(tmp):
                for f in fs:
                    if f.endswith((".bin", ".macro", ".map", ".df")):
                        files[os.path.relpath(os.path.join(dp, f), tmp)] = open(os.path.join(dp, f), "rb").read()
            out[mode] = (int(m.group(1)), int(m.group(2)), files, [ln for ln in r.stdout.splitlines() if "mass=" in ln])
    assert out["0"][0] == out["1"][0] == steps
    assert out["0"][1] == steps and out["1"][1] < steps / 2, f"batches: {out['0'][1]} without, {out['1'][1]} with deferral"
    assert out["0"][2].keys() == out["1"][2].keys() and len(out["0"][2]) >= 3
    for k in out["0"][2]:
        assert out["0"][2][k] == out["1"][2][k], f"{k} differs between per-step launches and deferred batches"
    assert out["0"][3] == out["1"][3]


@pytest.mark.gpu
@pytest.mark.parametrize("exe_name,arg", [("ref_sim_2_aa", ["--min-resolution", "1", "--max-resolution", "1", "--final-time", "20", "--use-forcing"]), ("ref_sim_1_ab", ["1"]), ("ref_sim2d_1_ab", ["1"])])
def test_unmodified_solvers_print_the_same_with_deferred_batches(exe_name, arg):
    """The reference's own solvers: every number they print or dump (probe errors of sim_2.cu:193-262, cuts of sim_1 / sim2d_1, which
    override updateKernelVelocities with a constant inflow) is the same with and without deferral."""
    exe = os.path.join(BIN, exe_name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe_name} not built (needs /root/reference at build time)")
    got = {}
    for mode in ("0", "1"):
        with tempfile.TemporaryDirectory() as tmp:
            r = subprocess.run([exe] + arg, capture_output=True, text=True, timeout=900, cwd=tmp, env=dict(os.environ, LBMX_HOST_BATCH=mode))
            assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
            m = re.findall(r"lbmx: (\d+) steps enqueued in (\d+) batches", r.stdout)
            assert m, r.stdout[-2000:]
            files = {}
            for dp, _, fs in os.walk(tmp):
                for f in fs:
                    if f.endswith(".bin"):
                        files[os.path.relpath(os.path.join(dp, f), tmp)] = open(os.path.join(dp, f), "rb").read()
            keep = [ln.split("] ", 1)[-1] for ln in r.stdout.splitlines() if "error_phys" in ln]
            got[mode] = ([(int(a), int(b)) for a, b in m], files, keep)
    assert [a for a, _ in got["0"][0]] == [a for a, _ in got["1"][0]]
    assert all(a == b for a, b in got["0"][0]) and all(b < a / 4 for a, b in got["1"][0]), (got["0"][0], got["1"][0])
    assert got["0"][1].keys() == got["1"][1].keys()
    for k in got["0"][1]:
        assert got["0"][1][k] == got["1"][1][k], k
    assert got["0"][2] == got["1"][2]


def test_obstacle_painters_on_cpu():
    """lbm3d/obstacles_lbm.h (lbmDrawSphere / Cylinder / Cube / BoundingBox) paints exactly the cells the reference's formulas select
    (obstacles_lbm.h:3-87), checked against numpy; host-only, no engine."""
    src = r'''
#include "lbm3d/core.h"
#include "lbm3d/obstacles_lbm.h"
#include <fstream>
using T = TraitsDP;
using COLL = D3Q27_CUM<T, D3Q27_EQ_INV_CUM<T>>;
using NSE = LBM_CONFIG<T, D3Q27_KernelStruct, NSE_Data_ConstInflow<T>, COLL, typename COLL::EQ, D3Q27_STREAMING<T>, D3Q27_BC_All, D3Q27_MACRO_Default<T>>;
int main(int argc, char** argv)
{
	TNLMPI_INIT mpi(argc, argv);
	using lat_t = Lattice<3, double, long>;
	lat_t lat;
	lat.global = lat_t::CoordinatesType(40, 24, 20);
	lat.physOrigin = {0., 0., 0.};
	lat.physDl = 0.05;
	lat.physDt = 1e-3;
	lat.physViscosity = 1e-5;
	LBM<NSE> nse(MPI_COMM_WORLD, lat);
	nse.allocateHostData();
	nse.resetMap(0);
	lbmDrawSphere(nse, 1, {0.52, 0.61, 0.49}, 0.21);
	lbmDrawCylinder(nse, 2, {1.31, 0.3, 0.52}, 0.16);
	lbmDrawCube(nse, 3, {1.71, 0.92, 0.71}, 0.11);
	lbmDrawBoundingBox(nse, 4, {0.2, 0.1, 0.1}, {0.45, 0.3, 0.25});
	auto& m = nse.blocks.front().hmap.v;
	std::ofstream(argv[1], std::ios::binary).write((const char*) m.data(), m.size() * sizeof(short));
	return 0;
}
'''
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, "paint.cpp")
        open(f, "w").write(src)
        exe = os.path.join(tmp, "paint")
        r = subprocess.run(["g++", "-std=c++17", f"-I{ROOT}/tnl_lbm_b200/host", f"-I{ROOT}/include", f"-I{ROOT}/tests/solver_shims", f, "-o", exe,
                            f"-L{ROOT}/tnl_lbm_b200", "-llbmx", f"-Wl,-rpath,{ROOT}/tnl_lbm_b200"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        out = os.path.join(tmp, "map.bin")
        r = subprocess.run([exe, out], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        got = np.fromfile(out, dtype=np.int16).reshape(40, 20, 24)  # [x, z, y]
    X, Y, Z, dl = 40, 24, 20, 0.05
    cell = lambda p: tuple(int(c / dl + 0.5) for c in p)  # Lattice::phys2lbm: (x - origin) / dl + 1/2, truncated
    x, z, y = np.meshgrid(np.arange(X), np.arange(Z), np.arange(Y), indexing="ij")
    ref = np.zeros((X, Z, Y), dtype=np.int16)
    c, r_ = cell((0.52, 0.61, 0.49)), 0.21 / dl
    rng = int(np.ceil(r_)) + 1
    box = lambda c, rng: (abs(x - c[0]) <= rng) & (abs(y - c[1]) <= rng) & (abs(z - c[2]) <= rng)
    ref[box(c, rng) & (np.sqrt((x - c[0]) ** 2.0 + (y - c[1]) ** 2.0 + (z - c[2]) ** 2.0) < r_)] = 1
    c, r_ = cell((1.31, 0.3, 0.52)), 0.16 / dl
    rng = int(np.ceil(r_)) + 1
    ref[(abs(x - c[0]) <= rng) & (abs(z - c[2]) <= rng) & (np.sqrt((x - c[0]) ** 2.0 + (z - c[2]) ** 2.0) < r_)] = 2
    c, r_ = cell((1.71, 0.92, 0.71)), 0.11 / dl
    rng = int(np.ceil(r_)) + 1
    ref[box(c, rng) & (x - c[0] < r_) & (y - c[1] < r_) & (z - c[2] < r_)] = 3
    p1 = [v / dl + 0.5 + 0.5 for v in (0.2, 0.1, 0.1)]
    p2 = [v / dl + 0.5 - 0.5 for v in (0.45, 0.3, 0.25)]
    ext = [int(round(abs(a - b))) for a, b in zip(p1, p2)]
    for i in range(ext[0] + 1):
        for j in range(ext[1] + 1):
            for k in range(ext[2] + 1):
                ref[int(p1[0] + i), int(p1[2] + k), int(p1[1] + j)] = 4
    assert np.array_equal(got, ref), f"{(got != ref).sum()} cells differ; painted counts {[(int(t), int((got == t).sum())) for t in range(5)]}"
    assert all((got == t).sum() > 0 for t in (1, 2, 3, 4))


@pytest.mark.gpu
def test_plain_c_client_of_the_abi():
    """examples/abi_minimal.c: include/lbmx.h used from C99 -- create, upload, step, download; mass and momentum budget on the host."""
    exe = os.path.join(BIN, "abi_minimal")
    if not os.path.exists(exe):
        pytest.skip("examples/bin/abi_minimal not built")
    r = subprocess.run([exe, "48", "101"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "nan=0" in r.stdout and "48^3 cells, 101 steps" in r.stdout


def _cylinder_geometry(X, Y, cx, cy, rad):
    """Per-cell types (0 fluid, 1 fluid next to the body, 2 solid) and the eight link fractions to a cylinder, in the file format of
    sim_2D/sim2d_2.cu:214-330 / sim2d_3.cu: `x y type theta_E theta_N theta_W theta_S theta_NE theta_NW theta_SW theta_SE`."""
    xs, ys = np.meshgrid(np.arange(X), np.arange(Y), indexing="ij")
    solid = (xs - cx) ** 2 + (ys - cy) ** 2 <= rad ** 2
    links = [(1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]  # lbm_data.h:69-83
    theta = np.full((8, X, Y), -1.0)
    for k, (dx, dy) in enumerate(links):
        for x in range(1, X - 1):
            for y in range(1, Y - 1):
                if solid[x, y] or not solid[x + dx, y + dy]:
                    continue
                a, b, c = dx * dx + dy * dy, 2 * ((x - cx) * dx + (y - cy) * dy), (x - cx) ** 2 + (y - cy) ** 2 - rad ** 2
                t = (-b - np.sqrt(b * b - 4 * a * c)) / (2 * a)
                theta[k, x, y] = min(max(t, 1e-3), 1.0)
    near = (theta > 0).any(axis=0)
    return np.where(solid, 2, np.where(near, 1, 0)), theta


def _write_geometry(path, ctype, theta):
    X, Y = ctype.shape
    with open(path, "w") as f:
        for x in range(X):
            for y in range(Y):
                f.write(f"{x} {y} {ctype[x, y]} " + " ".join(repr(float(theta[k, x, y])) for k in range(8)) + "\n")


def _sim2d_2_on_oracle(ctype, theta):
    """sim_2D/sim2d_2.cu at resolution 1 on the CPU oracle, with the solver's host logic (sim2d_2.cu:391-436, 471-505, 741-801) restated
    step by step; returns (TKE integral, [(tag, time, mean_samples, fluc_samples)], steps)."""
    X, Y = ctype.shape
    g = lc.G2
    d = O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=O.AB, macro=O.MACRO_WITH_MEAN_2D, inflow=O.INFLOW_PARABOLIC_Y, X=X, Y=Y, Z=1)
    m = d.new_map(g["FLUID"])
    m[:, 0, :] = np.where(ctype == 2, g["WALL"], np.where(ctype == 1, 12, g["FLUID"]))  # 12 = GEO_FLUID_NEAR_WALL
    m[0], m[X - 1] = g["INFLOW"], g["OUTFLOW_RIGHT"]
    m[:, :, 1] = m[:, :, Y - 2] = g["WALL"]
    m[:, :, 0] = m[:, :, Y - 1] = g["NOTHING"]
    dl = 0.50 / (Y - 2)
    dt = 1.0e-3 / 1.0e-3 * dl * dl
    vscale = dl / dt
    p = O.Params(lbmViscosity=dt / dl / dl * 1.0e-3, inflow_vx=1.5 * dt / dl, inflow_vy=1.0, inflow_vz=1.0 / 29.0, bouzidi=np.ascontiguousarray(theta[:, :, None, :]))
    orc = O.Oracle(d, "port")
    a = d.new_df()
    orc.set_equilibrium(a, 1.0, 0.0, 0.0, 0.0)
    b = a.copy()
    mac = d.new_macro()
    orc.initial_macro(p, a, mac)
    fluid = (m[:, 0, :] == g["FLUID"]) | (m[:, 0, :] == 12)
    roi = np.zeros((X, Y), dtype=bool)
    roi[64:96, 3:29] = True
    roi &= fluid
    ch = lambda k: mac[k][:, 0, :]  # noqa: E731
    mean_samples = fluc_samples = 0
    means_frozen = flucs_frozen = False
    mean_freeze_time, next_mean_check, prev_mean, mean_hits = -1.0, 1.5 + 0.05, -1.0, 0
    next_fluc_check, prev_fluc, fluc_hits = -1.0, -1.0, 0
    expect, events, it = None, [], 0

    def roi_mean_speed():
        if mean_samples <= 0:
            return 0.0
        return float(np.sqrt((vscale * (ch(3) / mean_samples)) ** 2 + (vscale * (ch(4) / mean_samples)) ** 2)[roi].mean())

    def roi_rms_fluc():
        if fluc_samples <= 0:
            return 0.0
        return float(np.sqrt(max(0.0, (ch(8) / fluc_samples + ch(9) / fluc_samples)[roi].sum() / roi.sum() * vscale * vscale)))

    def freeze():
        gc.freeze_means(mac, mean_samples)
        events.append(("mean_frozen", dt * it, mean_samples, fluc_samples))

    while True:
        t = dt * it
        acc_means = (not means_frozen) and 1.5 <= t < 5.5
        if acc_means:
            mean_samples += 1
        acc_flucs = False
        if not means_frozen:
            if t >= 1.5 + 1.0 and not t + 1e-12 < next_mean_check:
                curr = roi_mean_speed()
                mean_hits = 0 if prev_mean < 0 else (mean_hits + 1 if abs(curr - prev_mean) <= max(1e-3, 1e-3 * max(curr, 1e-6)) else 0)
                prev_mean, next_mean_check = curr, next_mean_check + 0.05
                if mean_hits >= 10:
                    means_frozen, mean_freeze_time = True, t
            if not means_frozen and t >= 5.5:
                means_frozen, mean_freeze_time = True, 5.5
            if means_frozen:
                acc_means = False
                freeze()
                fluc_samples, flucs_frozen, prev_fluc, next_fluc_check = 0, False, -1.0, mean_freeze_time + 0.05
        if means_frozen:
            acc_flucs = (not flucs_frozen) and t >= mean_freeze_time + 1.0
            if acc_flucs:
                fluc_samples += 1
            if not flucs_frozen and fluc_samples > 0 and not (t + 1e-12 < next_fluc_check or t < mean_freeze_time + 1.0):
                curr = roi_rms_fluc()
                fluc_hits = 0 if prev_fluc < 0 else (fluc_hits + 1 if abs(curr - prev_fluc) <= max(1e-3, 1e-3 * max(curr, 1e-6)) else 0)
                prev_fluc, next_fluc_check = curr, next_fluc_check + 0.05
                if fluc_hits >= 10:
                    flucs_frozen, acc_flucs = True, False
                    events.append(("fluc_frozen", t, mean_samples, fluc_samples))
            if flucs_frozen:
                events.append(("export", t, mean_samples, fluc_samples))
                expect = float((0.5 * (ch(8) / fluc_samples + ch(9) / fluc_samples) * vscale * vscale)[roi].sum() * dl * dl)
                break
        p.macro_gates = (O.GATE_MEANS if acc_means else 0) | (O.GATE_FLUCS if acc_flucs else 0)
        orc.step(p, a, b, mac, m, it, 1, 1)
        it += 1
        assert dt * it <= 10.0, "the restated solver logic never froze the fluctuation statistics"
    return expect, events, it


@pytest.mark.gpu
def test_reference_sim2d_2_turbulence_statistics_match_the_oracle():
    """sim_2D/sim2d_2.cu, unmodified: its solver-defined macro class (velocity sums, a mean the host freezes once it has settled, then
    fluctuation sums about it), the per-step gates in block.data, the macro round trip through the host at the freeze, and the
    turbulent kinetic energy of the wake integrated over a window as its one output.  Run in parity arithmetic (LBMX_STRICT_ARITH=1)
    the device state is the oracle's bit for bit, so every threshold decision of the solver falls on the same step here and there."""
    exe = os.path.join(BIN, "ref_sim2d_2_ab")
    if not os.path.exists(exe):
        pytest.skip("ref_sim2d_2_ab not built (needs /root/reference at build time)")
    X, Y = 128, 32
    ctype, theta = _cylinder_geometry(X, Y, 30.3, 15.2, 4.2)
    with tempfile.TemporaryDirectory() as tmp:
        geo = os.path.join(tmp, "cyl.txt")
        _write_geometry(geo, ctype, theta)
        r = subprocess.run([exe, "1", geo], capture_output=True, text=True, timeout=900, cwd=tmp, env=dict(os.environ, LBMX_STRICT_ARITH="1"))
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        value = float(open(os.path.join(tmp, "sim_2D", "values", "value_cyl.txt")).read())
        stats = [line.strip().split(",") for line in open(os.path.join(tmp, "sim_2D", "stats", "stats_cyl.csv"))][1:]
    assert "terminate flag triggered" in r.stdout
    expect, events, it = _sim2d_2_on_oracle(ctype, theta)
    # the solver's own log of its decisions: same tags, same times, same sample counts
    assert [s[0] for s in stats] == [e[0] for e in events], (stats, events)
    for s, e in zip(stats, events):
        assert abs(float(s[1]) - e[1]) <= 1e-7 and int(s[6]) == e[2] and int(s[7]) == e[3], (s, e)  # the CSV holds %.9g; one step is 2.8e-4 s
    assert expect > 0 and abs(value - expect) <= 1e-9 * expect, (value, expect)
    print(f"sim2d_2 wake TKE integral: solver {value:.12e}, oracle {expect:.12e}; mean frozen at t={events[0][1]:.4f} s, exported at t={events[-1][1]:.4f} s ({it} steps)")


@pytest.mark.gpu
def test_reference_sim2d_3_with_geometry_file_matches_the_oracle():
    """sim_2D/sim2d_3.cu, unmodified: D2Q9_CLBM with a solver-defined parabolic inflow (NSE2D_Data_ParabolicInflow), a geometry file with
    per-cell types and Bouzidi thetas, and one number as output -- the kinetic energy integrated over the third quarter of the channel.
    The geometry (a cylinder) is synthesised here; the same run is repeated on the CPU oracle and must give the same number."""
    exe = os.path.join(BIN, "ref_sim2d_3_ab")
    if not os.path.exists(exe):
        pytest.skip("ref_sim2d_3_ab not built (needs /root/reference at build time)")
    X, Y = 128, 32
    cx, cy, rad = 40.3, 15.6, 4.2
    xs, ys = np.meshgrid(np.arange(X), np.arange(Y), indexing="ij")
    solid = (xs - cx) ** 2 + (ys - cy) ** 2 <= rad ** 2
    links = [(1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]  # E N W S NE NW SW SE (lbm_data.h:69-83)
    theta = np.full((8, X, Y), -1.0)
    for k, (dx, dy) in enumerate(links):
        for x in range(1, X - 1):
            for y in range(1, Y - 1):
                if solid[x, y] or not solid[x + dx, y + dy]:
                    continue
                # first crossing of the link p + t (dx, dy), t in (0, 1], with the circle
                a, b, c = dx * dx + dy * dy, 2 * ((x - cx) * dx + (y - cy) * dy), (x - cx) ** 2 + (y - cy) ** 2 - rad ** 2
                t = (-b - np.sqrt(b * b - 4 * a * c)) / (2 * a)
                theta[k, x, y] = min(max(t, 1e-3), 1.0)
    near = (theta > 0).any(axis=0)
    ctype = np.where(solid, 2, np.where(near, 1, 0))
    with tempfile.TemporaryDirectory() as tmp:
        geo = os.path.join(tmp, "cyl.txt")
        with open(geo, "w") as f:
            for x in range(X):
                for y in range(Y):
                    f.write(f"{x} {y} {ctype[x, y]} " + " ".join(repr(float(theta[k, x, y])) for k in range(8)) + "\n")
        r = subprocess.run([exe, "1", geo], capture_output=True, text=True, timeout=600, cwd=tmp)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        value = float(open(os.path.join(tmp, "sim_2D", "values", "value_cyl.txt")).read())
    assert "physFinalTime reached" in r.stdout
    # the same run on the CPU oracle, set up the way the solver sets it up (sim2d_3.cu:82-99, 268-318)
    g = lc.G2
    d = O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=O.AB, inflow=O.INFLOW_PARABOLIC_Y, X=X, Y=Y, Z=1)
    m = d.new_map(g["FLUID"])
    m[:, 0, :] = np.where(ctype == 2, g["WALL"], np.where(ctype == 1, 12, g["FLUID"]))  # 12 = GEO_FLUID_NEAR_WALL
    m[0], m[X - 1] = g["INFLOW"], g["OUTFLOW_RIGHT"]
    m[:, :, 1] = m[:, :, Y - 2] = g["WALL"]
    m[:, :, 0] = m[:, :, Y - 1] = g["NOTHING"]
    dl = 0.50 / (Y - 2)
    dt = 1.0e-3 / 1.0e-3 * dl * dl
    nu = dt / dl / dl * 1.0e-3  # Lattice::phys2lbmViscosity
    steps = 0
    while dt * steps <= 4.0:  # core.h: quit once physTime() > physFinalTime
        steps += 1
    p = O.Params(lbmViscosity=nu, inflow_vx=1.5 * dt / dl, inflow_vy=1.0, inflow_vz=1.0 / 29.0, bouzidi=np.ascontiguousarray(theta[:, :, None, :]))
    orc = O.Oracle(d, "port")
    a = d.new_df()
    orc.set_equilibrium(a, 1.0, 0.0, 0.0, 0.0)
    b = a.copy()
    mac = d.new_macro()
    orc.step(p, a, b, mac, m, 0, steps, os.cpu_count() or 4)
    fluid = (m[:, 0, :] == g["FLUID"]) | (m[:, 0, :] == 12)
    roi = np.zeros((X, Y), dtype=bool)
    roi[64:96, 1:Y - 1] = True
    ux, uy = mac[1][:, 0, :] / dt * dl, mac[2][:, 0, :] / dt * dl
    expect = float((0.5 * (ux * ux + uy * uy))[roi & fluid].sum() * dl * dl)
    assert value > 0 and abs(value - expect) <= 1e-9 * expect, (value, expect)
    print(f"sim2d_3 kinetic-energy integral: solver {value:.12e}, oracle {expect:.12e}, {steps} steps")
