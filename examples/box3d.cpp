// box3d.cpp -- the headline workload of bench.py (BASELINE.json configs[2]: D3Q27 cumulant, EQ_INV_CUM, fp64, periodic box, body force)
// written as an ordinary solver against the host mirror of TNL-LBM's interface, in the style of sim_NSE/sim_2.cu --use-forcing:
// compose LBM_CONFIG, derive StateLocal from State<NSE>, paint the map, execute().  Nothing here knows about batches or the C ABI;
// the GLUPS= lines are the mirror's own (State::AfterSimUpdate, state.hpp:1244-1262).  bench.py runs it as a child process and
// reports the last GLUPS= line next to its own figure ("dropin_mirror").
//
//   g++ -std=c++17 [-DAA_PATTERN] -Itnl_lbm_b200/host -Iinclude examples/box3d.cpp -Ltnl_lbm_b200 -llbmx -o box3d
//   ./box3d X Y Z steps print_period [out_prefix]
//        print_period : steps between two GLUPS= lines (= between two copies of rho,u to the host, state.hpp:1134-1142)
//        out_prefix   : write out_prefix.macro (dreal, reference layout) at the end
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
	using State<NSE>::nse;
	using real = typename TRAITS::real;
	using idx = typename TRAITS::idx;
	using lat_t = Lattice<3, real, idx>;

	void setupBoundaries() override { nse.resetMap(BC::GEO_PERIODIC); }
	StateLocal(const std::string& id, const TNL::MPI::Comm& communicator, lat_t lat) : State<NSE>(id, communicator, std::move(lat), true) {}
};

int main(int argc, char** argv)
{
	TNLMPI_INIT mpi(argc, argv);
	if (argc < 6) {
		std::fprintf(stderr, "usage: %s X Y Z steps print_period [out_prefix]\n", argv[0]);
		return 1;
	}
	using TRAITS = TraitsDP;
	using COLL = D3Q27_CUM<TRAITS, D3Q27_EQ_INV_CUM<TRAITS>>;
	using NSE = LBM_CONFIG<TRAITS, D3Q27_KernelStruct, NSE_Data_ConstInflow<TRAITS>, COLL, typename COLL::EQ, D3Q27_STREAMING<TRAITS>, D3Q27_BC_All,
						   D3Q27_MACRO_Default<TRAITS>>;
	using real = typename TRAITS::real;
	using lat_t = Lattice<3, real, typename TRAITS::idx>;
	try {
		const int X = atoi(argv[1]), Y = atoi(argv[2]), Z = atoi(argv[3]), steps = atoi(argv[4]), period = atoi(argv[5]);
		const real LBM_VISCOSITY = 1e-3, PHYS_VISCOSITY = 1.5e-5, PHYS_DL = 1.0 / (real) Y;
		lat_t lat;
		lat.global = typename lat_t::CoordinatesType(X, Y, Z);
		lat.physDl = PHYS_DL;
		lat.physDt = LBM_VISCOSITY / PHYS_VISCOSITY * PHYS_DL * PHYS_DL;
		lat.physViscosity = PHYS_VISCOSITY;

		StateLocal<NSE> state("box3d", MPI_COMM_WORLD, lat);
		if (! state.canCompute())
			return 0;
		for (auto& block : state.nse.blocks) {
			block.data.fx = 1e-6;
			block.data.fy = 0;
			block.data.fz = 0;
		}
		state.nse.physFinalTime = (steps - 0.5) * lat.physDt;
		state.cnt[PRINT].period = period * lat.physDt;
		execute(state);
		state.nse.copyMacroToHost();
		auto& block = state.nse.blocks.front();
		double mass = 0, mom = 0;
		const size_t n = block.hmacro.n;
		for (size_t i = 0; i < n; i++) {
			mass += block.hmacro.v[i];
			mom += block.hmacro.v[n + i];
		}
		if (argc > 6)
			std::ofstream(std::string(argv[6]) + ".macro", std::ios::binary).write((const char*) block.hmacro.v.data(), block.hmacro.v.size() * sizeof(typename TRAITS::dreal));
		std::printf("iterations=%d mean_rho=%.15f mean_vx=%.15e\n", state.nse.iterations, mass / (double) n, mom / (double) n);
		return 0;
	}
	catch (const std::exception& e) {
		std::fprintf(stderr, "error: %s\n", e.what());
		return 2;
	}
}
