// obstacles_lbm.h -- map-painting helpers a solver calls from setupBoundaries() (same names and arguments as the reference's
// include/lbm3d/obstacles_lbm.h:3-102; host code on top of LBM::setMap, nothing device-side).
#pragma once
#include <cmath>

#include "lbmx_host.h"

namespace lbmx_host {
// lattice cell that contains a physical point, and a physical length in lattice units (obstacles_lbm.h:6-8)
template <typename LBM>
typename LBM::idx3d cell_of(const LBM& lbm, typename LBM::point_t phys)
{
	const typename LBM::point_t p = lbm.lat.phys2lbmPoint(phys);
	return typename LBM::idx3d((typename LBM::idx) p.x(), (typename LBM::idx) p.y(), (typename LBM::idx) p.z());
}
}  // namespace lbmx_host

// cells whose offset from the centre cell is below the radius in every direction (obstacles_lbm.h:3-16: the lower sides are not
// bounded by the radius, only by the scan range -- reproduced)
template <typename LBM>
void lbmDrawCube(LBM& lbm, typename LBM::map_t wall_tag, typename LBM::point_t phys_center, typename LBM::real phys_radius)
{
	using idx = typename LBM::idx;
	const typename LBM::idx3d c = lbmx_host::cell_of(lbm, phys_center);
	const typename LBM::real r = phys_radius / lbm.lat.physDl;
	const idx range = (idx) std::ceil(r) + 1;
	for (idx y = c.y() - range; y <= c.y() + range; y++)
		for (idx z = c.z() - range; z <= c.z() + range; z++)
			for (idx x = c.x() - range; x <= c.x() + range; x++)
				if (x - c.x() < r && y - c.y() < r && z - c.z() < r)
					lbm.setMap(x, y, z, wall_tag);
}

// cells whose centre-to-centre distance (in cells) is below the radius (obstacles_lbm.h:18-34)
template <typename LBM>
void lbmDrawSphere(LBM& lbm, typename LBM::map_t wall_tag, typename LBM::point_t phys_center, typename LBM::real phys_radius)
{
	using idx = typename LBM::idx;
	using real = typename LBM::real;
	const typename LBM::idx3d c = lbmx_host::cell_of(lbm, phys_center);
	const real r = phys_radius / lbm.lat.physDl;
	const idx range = (idx) std::ceil(r) + 1;
	for (idx y = c.y() - range; y <= c.y() + range; y++)
		for (idx z = c.z() - range; z <= c.z() + range; z++)
			for (idx x = c.x() - range; x <= c.x() + range; x++) {
				const real dx = (real) (x - c.x()), dy = (real) (y - c.y()), dz = (real) (z - c.z());
				if (std::sqrt(dx * dx + dy * dy + dz * dz) < r)
					lbm.setMap(x, y, z, wall_tag);
			}
}

// a cylinder along y through the whole lattice (obstacles_lbm.h:36-52)
template <typename LBM>
void lbmDrawCylinder(LBM& lbm, typename LBM::map_t wall_tag, typename LBM::point_t phys_center, typename LBM::real phys_radius)
{
	using idx = typename LBM::idx;
	using real = typename LBM::real;
	const typename LBM::idx3d c = lbmx_host::cell_of(lbm, phys_center);
	const real r = phys_radius / lbm.lat.physDl;
	const idx range = (idx) std::ceil(r) + 1;
	for (idx y = 0; y <= lbm.lat.global.y() - 1; y++)
		for (idx z = c.z() - range; z <= c.z() + range; z++)
			for (idx x = c.x() - range; x <= c.x() + range; x++) {
				const real dx = (real) (x - c.x()), dz = (real) (z - c.z());
				if (std::sqrt(dx * dx + dz * dz) < r)
					lbm.setMap(x, y, z, wall_tag);
			}
}

// box between two wall coordinates; walls sit half-way between lattice sites, hence the half-cell shifts (obstacles_lbm.h:54-87)
template <typename LBM>
void lbmDrawBoundingBox(LBM& lbm, typename LBM::map_t wall_tag, typename LBM::point_t phys_point1, typename LBM::point_t phys_point2)
{
	using idx = typename LBM::idx;
	typename LBM::point_t p1 = lbm.lat.phys2lbmPoint(phys_point1), p2 = lbm.lat.phys2lbmPoint(phys_point2);
	for (int a = 0; a < 3; a++) {
		if (p1[a] < p2[a]) {
			p1[a] += 0.5f;
			p2[a] -= 0.5f;
		}
		else {
			p1[a] -= 0.5f;
			p2[a] += 0.5f;
		}
	}
	for (idx y = 0; y <= std::round(std::abs(p1.y() - p2.y())); y++)
		for (idx z = 0; z <= std::round(std::abs(p1.z() - p2.z())); z++)
			for (idx x = 0; x <= std::round(std::abs(p1.x() - p2.x())); x++)
				lbm.setMap((idx) (p1.x() + x), (idx) (p1.y() + y), (idx) (p1.z() + z), wall_tag);
}

// two cubes side by side with a third on top of the +x one (obstacles_lbm.h:89-102)
template <typename LBM>
void lbmDrawCUBI(LBM& lbm, typename LBM::map_t wall_tag, typename LBM::point_t phys_center, typename LBM::real phys_edge_length)
{
	using point_t = typename LBM::point_t;
	const typename LBM::real e = phys_edge_length;
	lbmDrawBoundingBox(lbm, wall_tag, point_t(phys_center.x() - e, phys_center.y() - e / 2, phys_center.z() - e), point_t(phys_center.x() + e, phys_center.y() + e / 2, phys_center.z()));
	lbmDrawBoundingBox(lbm, wall_tag, point_t(phys_center.x(), phys_center.y() - e / 2, phys_center.z()), point_t(phys_center.x() + e, phys_center.y() + e / 2, phys_center.z() + e));
}
