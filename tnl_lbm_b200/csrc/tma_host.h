// tma_host.h -- host side of k_bulk_tma: the tile geometry.
#pragma once

namespace lbmx {

// cells of one row per CTA: the largest power of two that divides Y, at most 128; 0 if k_bulk_tma cannot be used on this lattice.
// The copy engine moves multiples of 16 bytes between 16-byte aligned addresses: a row shifted by one element goes as its aligned middle
// part (tile_y - E elements, E = 16 / sizeof(real)), so a tile row has to span at least a few granules.
inline int tma_tile_y(long long Y, int sizeof_real)
{
	int t = 1;
	while (t < 128 && Y % (2 * t) == 0)
		t *= 2;
	const int E = 16 / sizeof_real;
	return t >= 4 * E ? t : 0;
}

}  // namespace lbmx
