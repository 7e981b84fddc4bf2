// obstacles_lbm.h -- map-painting helpers a solver calls from setupBoundaries().  Same function names, argument meaning and painted
// cells as the reference's include/lbm3d/obstacles_lbm.h:3-102; host code on top of LBM::setMap, nothing device-side.  All shapes
// go through one scan routine: a box of lattice cells and a predicate on the offset from the shape's centre cell.
#pragma once
#include <cmath>

#include "lbmx_host.h"

namespace lbmx_host {
template <typename LBM>
struct Painter
{
	using idx = typename LBM::idx;
	using real = typename LBM::real;
	using point_t = typename LBM::point_t;
	LBM& lbm;
	typename LBM::map_t tag;

	// lattice cell that contains a physical point (truncation, as the reference's conversion of phys2lbmPoint to idx3d does)
	void centre_cell(point_t phys, idx (&c)[3]) const
	{
		const point_t p = lbm.lat.phys2lbmPoint(phys);
		for (int a = 0; a < 3; a++)
			c[a] = (idx) p[a];
	}
	// tag every cell of [lo, hi] (inclusive, per axis) whose offset (dx, dy, dz) from `c` satisfies `inside`
	template <typename Inside>
	void scan(const idx (&c)[3], const idx (&lo)[3], const idx (&hi)[3], Inside inside) const
	{
		for (idx y = lo[1]; y <= hi[1]; y++)
			for (idx z = lo[2]; z <= hi[2]; z++)
				for (idx x = lo[0]; x <= hi[0]; x++)
					if (inside(x - c[0], y - c[1], z - c[2]))
						lbm.setMap(x, y, z, tag);
	}
	// scan box = centre cell +- (ceil(radius) + 1), optionally the whole lattice along y
	template <typename Inside>
	void around(point_t phys_center, real radius_cells, bool whole_y, Inside inside) const
	{
		idx c[3], lo[3], hi[3];
		centre_cell(phys_center, c);
		const idx range = (idx) std::ceil(radius_cells) + 1;
		for (int a = 0; a < 3; a++) {
			lo[a] = c[a] - range;
			hi[a] = c[a] + range;
		}
		if (whole_y) {
			lo[1] = 0;
			hi[1] = lbm.lat.global.y() - 1;
		}
		scan(c, lo, hi, inside);
	}
};
}  // namespace lbmx_host

// obstacles_lbm.h:3-16.  Only the upper sides are bounded by the radius (offset < r); the lower sides end with the scan box, i.e.
// the cube reaches ceil(r) + 1 cells below the centre -- the reference's behaviour, kept.
template <typename LBM>
void lbmDrawCube(LBM& lbm, typename LBM::map_t wall_tag, typename LBM::point_t phys_center, typename LBM::real phys_radius)
{
	using idx = typename LBM::idx;
	const typename LBM::real r = phys_radius / lbm.lat.physDl;
	lbmx_host::Painter<LBM>{lbm, wall_tag}.around(phys_center, r, false, [r](idx dx, idx dy, idx dz) { return dx < r && dy < r && dz < r; });
}

// obstacles_lbm.h:18-34: centre-to-centre distance in cells below the radius
template <typename LBM>
void lbmDrawSphere(LBM& lbm, typename LBM::map_t wall_tag, typename LBM::point_t phys_center, typename LBM::real phys_radius)
{
	using idx = typename LBM::idx;
	using real = typename LBM::real;
	const real r = phys_radius / lbm.lat.physDl;
	lbmx_host::Painter<LBM>{lbm, wall_tag}.around(phys_center, r, false, [r](idx dx, idx dy, idx dz) {
		return std::sqrt((real) dx * (real) dx + (real) dy * (real) dy + (real) dz * (real) dz) < r;
	});
}

// obstacles_lbm.h:36-52: a cylinder along y through the whole lattice
template <typename LBM>
void lbmDrawCylinder(LBM& lbm, typename LBM::map_t wall_tag, typename LBM::point_t phys_center, typename LBM::real phys_radius)
{
	using idx = typename LBM::idx;
	using real = typename LBM::real;
	const real r = phys_radius / lbm.lat.physDl;
	lbmx_host::Painter<LBM>{lbm, wall_tag}.around(phys_center, r, true, [r](idx dx, idx, idx dz) { return std::sqrt((real) dx * (real) dx + (real) dz * (real) dz) < r; });
}

// obstacles_lbm.h:54-87: box between two wall coordinates.  Walls sit half-way between lattice sites, so both corners move half a
// cell inwards; the painted extent per axis is round(|p1 - p2|) + 1 cells starting at the first corner.
template <typename LBM>
void lbmDrawBoundingBox(LBM& lbm, typename LBM::map_t wall_tag, typename LBM::point_t phys_point1, typename LBM::point_t phys_point2)
{
	using idx = typename LBM::idx;
	typename LBM::point_t p1 = lbm.lat.phys2lbmPoint(phys_point1), p2 = lbm.lat.phys2lbmPoint(phys_point2);
	idx extent[3];
	for (int a = 0; a < 3; a++) {
		const float inwards = p1[a] < p2[a] ? 0.5f : -0.5f;
		p1[a] += inwards;
		p2[a] -= inwards;
		extent[a] = (idx) std::round(std::abs(p1[a] - p2[a]));
	}
	for (idx j = 0; j <= extent[1]; j++)
		for (idx k = 0; k <= extent[2]; k++)
			for (idx i = 0; i <= extent[0]; i++)
				lbm.setMap((idx) (p1.x() + i), (idx) (p1.y() + j), (idx) (p1.z() + k), wall_tag);
}

// obstacles_lbm.h:89-102: two cubes side by side along x with a third one on top of the +x cube
template <typename LBM>
void lbmDrawCUBI(LBM& lbm, typename LBM::map_t wall_tag, typename LBM::point_t phys_center, typename LBM::real phys_edge_length)
{
	using point_t = typename LBM::point_t;
	const typename LBM::real e = phys_edge_length, cx = phys_center.x(), cy = phys_center.y(), cz = phys_center.z();
	lbmDrawBoundingBox(lbm, wall_tag, point_t(cx - e, cy - e / 2, cz - e), point_t(cx + e, cy + e / 2, cz));
	lbmDrawBoundingBox(lbm, wall_tag, point_t(cx, cy - e / 2, cz), point_t(cx + e, cy + e / 2, cz + e));
}
