// channel3d.cpp -- a solver written against the host mirror of TNL-LBM's interface (tnl_lbm_b200/host/lbm3d/core.h), in the
// style of the reference's sim_NSE/sim_1.cu: compose LBM_CONFIG, derive StateLocal from State<NSE>, paint the map, execute().
// Orifice channel: moment inflow on the left (GEO_INFLOW_LEFT), GEO_OUTFLOW_RIGHT, walls behind a GEO_NOTHING shell.
//
//   g++ -std=c++17 -Itnl_lbm_b200/host -Iinclude examples/channel3d.cpp -Ltnl_lbm_b200 -llbmx -o channel3d
//   ./channel3d X Y Z steps out_prefix [f32] [halt=N] [dump]
//        -> out_prefix.map (int16) and out_prefix.macro (dreal), reference layout
//   halt=N : stop after N steps and save a checkpoint (results_channel3d/checkpoint.bp + flag.loadstate); the next invocation
//            in the same directory resumes from it, exactly like a reference solver after a wall-time stop
//   dump   : register cuts and write outputData() fields through the raw-dump writers (results_channel3d/output_*)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "lbm3d/core.h"

template <typename NSE>
struct StateLocal : State<NSE>
{
	using TRAITS = typename NSE::TRAITS;
	using BC = typename NSE::BC;
	using MACRO = typename NSE::MACRO;
	using State<NSE>::nse;
	using real = typename TRAITS::real;
	using idx = typename TRAITS::idx;
	using lat_t = Lattice<3, real, idx>;

	real lbm_inflow_vx = 0;
	double probed_mass = 0;
	int halt_at = 0;  // stop the time loop after this many steps (stands in for the wall-time limit), 0 = never


	void setupBoundaries() override
	{
		const idx X = nse.lat.global.x(), Y = nse.lat.global.y(), Z = nse.lat.global.z();
		nse.setBoundaryX(0, BC::GEO_INFLOW_LEFT);
		nse.setBoundaryX(X - 1, BC::GEO_OUTFLOW_RIGHT);
		nse.setBoundaryZ(1, BC::GEO_WALL);
		nse.setBoundaryZ(Z - 2, BC::GEO_WALL);
		nse.setBoundaryY(1, BC::GEO_WALL);
		nse.setBoundaryY(Y - 2, BC::GEO_WALL);
		nse.setBoundaryZ(0, BC::GEO_NOTHING);
		nse.setBoundaryZ(Z - 1, BC::GEO_NOTHING);
		nse.setBoundaryY(0, BC::GEO_NOTHING);
		nse.setBoundaryY(Y - 1, BC::GEO_NOTHING);
		const idx cx = X / 5, width = Z / 10;
		for (idx px = cx; px <= cx + width; px++)
			for (idx pz = 1; pz <= Z - 2; pz++)
				for (idx py = 1; py <= Y - 2; py++)
					if (! (pz >= Z * 4 / 10 && pz <= Z * 6 / 10 && py >= Y * 4 / 10 && py <= Y * 6 / 10))
						nse.setMap(px, py, pz, BC::GEO_WALL);
	}
	void updateKernelVelocities() override
	{
		if (halt_at > 0 && nse.iterations + 1 >= halt_at)  // the step about to run is the last one (the hook runs before SimUpdate, core.h:44-46)
			nse.terminate = true;
		for (auto& block : nse.blocks) {
			block.data.inflow_vx = lbm_inflow_vx;
			block.data.inflow_vy = 0;
			block.data.inflow_vz = 0;
		}
	}
	// the fields a writer asks for, by index (same protocol as the reference's solvers, e.g. sim_NSE/sim_1.cu:57-73)
	bool outputData(const LBM_BLOCK<NSE>& block, int index, int dof, char* desc, idx x, idx y, idx z, real& value, int& dofs) override
	{
		int k = 0;
		if (index == k++)
			return this->vtk_helper("lbm_density", block.hmacro(MACRO::e_rho, x, y, z), 1, desc, value, dofs);
		if (index == k++) {
			const int c = dof == 0 ? MACRO::e_vx : (dof == 1 ? MACRO::e_vy : MACRO::e_vz);
			return this->vtk_helper("velocity", block.hmacro(c, x, y, z), 3, desc, value, dofs);
		}
		return false;
	}
	void probe1() override
	{
		// total mass from the host copy of the density field (copied on this cadence by SimUpdate)
		probed_mass = 0;
		nse.forLocalLatticeSites([&](auto& block, idx x, idx y, idx z) { probed_mass += block.hmacro(MACRO::e_rho, x, y, z); });
	}
	StateLocal(const std::string& id, const TNL::MPI::Comm& communicator, lat_t lat) : State<NSE>(id, communicator, std::move(lat)) {}
};

template <typename TRAITS>
int run(int X, int Y, int Z, int steps, const char* prefix, int halt, bool dump)
{
	using COLL = D3Q27_CUM<TRAITS, D3Q27_EQ_INV_CUM<TRAITS>>;
	using NSE = LBM_CONFIG<TRAITS, D3Q27_KernelStruct, NSE_Data_ConstInflow<TRAITS>, COLL, typename COLL::EQ, D3Q27_STREAMING<TRAITS>, D3Q27_BC_All,
						   D3Q27_MACRO_Default<TRAITS>>;
	using real = typename TRAITS::real;
	using lat_t = Lattice<3, real, typename TRAITS::idx>;

	const real LBM_VISCOSITY = 1e-3, PHYS_VISCOSITY = 1.5e-5, PHYS_DL = 0.41 / ((real) Y - 2);
	lat_t lat;
	lat.global = typename lat_t::CoordinatesType(X, Y, Z);
	lat.physDl = PHYS_DL;
	lat.physDt = LBM_VISCOSITY / PHYS_VISCOSITY * PHYS_DL * PHYS_DL;
	lat.physViscosity = PHYS_VISCOSITY;

	StateLocal<NSE> state("channel3d", MPI_COMM_WORLD, lat);
	if (! state.canCompute())
		return 0;
	state.lbm_inflow_vx = 0.04;
	state.halt_at = halt;
	state.nse.physFinalTime = (steps - 0.5) * lat.physDt;
	state.cnt[PRINT].period = 10 * lat.physDt;
	state.cnt[PROBE1].period = 5 * lat.physDt;
	if (dump) {
		state.cnt[VTK2D].period = 20 * lat.physDt;
		state.cnt[VTK3D].period = 30 * lat.physDt;
		state.cnt[VTK3DCUT].period = 30 * lat.physDt;
		state.add2Dcut_X(X / 2, "cutsX/cut_X");
		state.add2Dcut_Z(Z / 2, "cut_Z");
		state.add3Dcut(X / 4, Y / 4, Z / 4, X / 2, Y / 2, Z / 2, 2, "box");
		state.cnt[VTK1D].period = 30 * lat.physDt;
		state.add1Dcut_X(lat.lbm2physY(Y / 2), lat.lbm2physZ(Z / 2), "centre_line");
		state.add1Dcut_Z(lat.lbm2physX(X / 2), lat.lbm2physY(Y / 2), "profile_z");
	}
	execute(state);
	if (halt > 0) {	 // what core.h does when the wall time runs out (core.h:57-63)
		state.copyAllToHost();
		state.saveState();
	}

	auto& block = state.nse.blocks.front();
	state.nse.copyMacroToHost();
	state.nse.copyMapToHost();
	// one piece per rank (x-slab [offset.x, offset.x + local.x) of the global lattice), named like the single-process output when alone
	const std::string piece = std::string(prefix) + (state.nse.nproc > 1 ? ".rank" + std::to_string(state.nse.rank) : "");
	std::ofstream(piece + ".map", std::ios::binary).write((const char*) block.hmap.v.data(), block.hmap.v.size() * sizeof(short));
	std::ofstream(piece + ".macro", std::ios::binary).write((const char*) block.hmacro.v.data(), block.hmacro.v.size() * sizeof(typename TRAITS::dreal));
	const double mass = TNL::MPI::reduce(state.probed_mass, MPI_SUM, MPI_COMM_WORLD);  // as sim_NSE/sim_2.cu:259-260 reduces its error norms
	if (state.nse.rank == 0)
		std::printf("iterations=%d mass=%.12f lbmViscosity=%.17g inflow_vx=%.17g ranks=%d\n", state.nse.iterations, mass, (double) block.data.lbmViscosity,
					(double) block.data.inflow_vx, state.nse.nproc);
	std::printf("rank %d: x-slab offset %ld size %ld\n", state.nse.rank, (long) block.offset.x(), (long) block.local.x());
	return 0;
}

int main(int argc, char** argv)
{
	TNLMPI_INIT mpi(argc, argv);
	if (argc < 6) {
		std::fprintf(stderr, "usage: %s X Y Z steps out_prefix [f32] [halt=N] [dump]\n", argv[0]);
		return 1;
	}
	bool f32 = false, dump = false;
	int halt = 0;
	for (int i = 6; i < argc; i++) {
		if (std::strcmp(argv[i], "f32") == 0)
			f32 = true;
		else if (std::strcmp(argv[i], "dump") == 0)
			dump = true;
		else if (std::strncmp(argv[i], "halt=", 5) == 0)
			halt = atoi(argv[i] + 5);
	}
	try {
		return f32 ? run<TraitsSP>(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argv[5], halt, dump)
				   : run<TraitsDP>(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argv[5], halt, dump);
	}
	catch (const std::exception& e) {
		std::fprintf(stderr, "error: %s\n", e.what());
		return 2;
	}
}
