#pragma once
#include "../Backend/Macros.h"
namespace TNL::Containers {

template <typename I, std::size_t... sizes>
struct SizesHolder {};
template <typename I, std::size_t dim, std::size_t value>
struct ConstStaticSizesHolder {};

// Storage index of the reference's 3-D arrays: permutation (x, z, y) with y fastest,
// symmetric overlaps o[] per dimension (call sites: lbm_data.h:49-57, defs.h:85-86).
template <typename I>
struct Indexer3
{
	I s[3] = {0, 0, 0};	 // sizes x, y, z (without overlaps)
	I o[3] = {0, 0, 0};	 // overlaps x, y, z
	template <int i>
	__cuda_callable__ I getSize() const
	{
		return s[i];
	}
	template <int i>
	__cuda_callable__ const I& getOverlap() const
	{
		return o[i];
	}
	__cuda_callable__ I getStorageSize() const
	{
		return (s[0] + 2 * o[0]) * (s[1] + 2 * o[1]) * (s[2] + 2 * o[2]);
	}
	__cuda_callable__ I getStorageIndex(I x, I y, I z) const
	{
		return ((x + o[0]) * (s[2] + 2 * o[2]) + (z + o[2])) * (s[1] + 2 * o[1]) + (y + o[1]);
	}
};

template <typename Value, typename Sizes, typename Perm, typename Device, typename I, typename Overlaps>
struct NDArray
{
	using IndexerType = Indexer3<I>;
	using ViewType = NDArray;
	using ConstViewType = NDArray;
};

template <typename Array>
struct DistributedNDArray
{
	using ViewType = DistributedNDArray;
	using ConstViewType = DistributedNDArray;
};

// only the enumerator names used by defs.h:309-340 matter; the values are arbitrary
enum class SyncDirection : std::uint8_t
{
	None = 0,
	Right = 1,
	Left = 2,
	Top = 4,
	Bottom = 8,
	Front = 16,
	Back = 32,
	TopRight = Top | Right,
	BottomLeft = Bottom | Left,
	BottomRight = Bottom | Right,
	TopLeft = Top | Left,
	FrontRight = Front | Right,
	BackLeft = Back | Left,
	BackRight = Back | Right,
	FrontLeft = Front | Left,
	FrontTop = Front | Top,
	BackBottom = Back | Bottom,
	BackTop = Back | Top,
	FrontBottom = Front | Bottom,
	FrontTopRight = Front | Top | Right,
	BackBottomLeft = Back | Bottom | Left,
	BackTopRight = Back | Top | Right,
	FrontBottomLeft = Front | Bottom | Left,
	FrontBottomRight = Front | Bottom | Right,
	BackTopLeft = Back | Top | Left,
	BackBottomRight = Back | Bottom | Right,
	FrontTopLeft = Front | Top | Left,
	All = 63
};
}  // namespace TNL::Containers
