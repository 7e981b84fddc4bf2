#pragma once
#include "../Backend/Macros.h"
namespace TNL::Containers {
template <int N, typename T>
struct StaticVector
{
	T d[N]{};
	__cuda_callable__ StaticVector() {}
	__cuda_callable__ StaticVector(T fill)
	{
		for (int i = 0; i < N; i++)
			d[i] = fill;
	}
	__cuda_callable__ StaticVector(T a, T b, T c) : d{a, b, c} {}
	__cuda_callable__ T& x() { return d[0]; }
	__cuda_callable__ T& y() { return d[1]; }
	__cuda_callable__ T& z() { return d[2]; }
	__cuda_callable__ const T& x() const { return d[0]; }
	__cuda_callable__ const T& y() const { return d[1]; }
	__cuda_callable__ const T& z() const { return d[2]; }
	__cuda_callable__ T& operator[](int i) { return d[i]; }
	__cuda_callable__ const T& operator[](int i) const { return d[i]; }
};
}  // namespace TNL::Containers
