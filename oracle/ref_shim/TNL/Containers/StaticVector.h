#pragma once
#include "../Backend/Macros.h"
namespace TNL::Containers {
template <int N, typename T>
struct StaticVector
{
	T d[N]{};
	StaticVector() = default;
	StaticVector(T fill)
	{
		for (int i = 0; i < N; i++)
			d[i] = fill;
	}
	StaticVector(T a, T b, T c) : d{a, b, c} {}
	T& x() { return d[0]; }
	T& y() { return d[1]; }
	T& z() { return d[2]; }
	const T& x() const { return d[0]; }
	const T& y() const { return d[1]; }
	const T& z() const { return d[2]; }
	T& operator[](int i) { return d[i]; }
	const T& operator[](int i) const { return d[i]; }
};
}  // namespace TNL::Containers
