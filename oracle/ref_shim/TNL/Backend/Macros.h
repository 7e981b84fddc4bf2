// Minimal stand-in for the un-vendored TNL library (pinned by the reference at
// commit 9e7b0f44, CMakeLists.txt:168-175), written fresh for the oracle build.
// It provides only what the reference's hot-path headers name.
// TEST INFRASTRUCTURE ONLY: used by oracle/Makefile to compile the reference's own
// per-cell code from /root/reference/include into oracle/_ref/.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
	#define __cuda_callable__ __host__ __device__
#else
	#define __cuda_callable__
#endif

namespace TNL {
namespace Devices {
struct Host {};
struct Cuda {};
}  // namespace Devices

namespace Backend {
template <typename T>
__cuda_callable__ inline T ldg(const T& value)
{
#ifdef __CUDA_ARCH__
	return __ldg(&value);  // what TNL::Backend::ldg does in device code
#else
	return value;
#endif
}
}  // namespace Backend

template <typename T>
__cuda_callable__ inline void swap(T& a, T& b)
{
	T t = a;
	a = b;
	b = t;
}
template <typename A, typename B>
__cuda_callable__ inline auto min(const A& a, const B& b) -> decltype(a + b)
{
	using R = decltype(a + b);
	return (R) a < (R) b ? (R) a : (R) b;
}
template <typename A, typename B>
__cuda_callable__ inline auto max(const A& a, const B& b) -> decltype(a + b)
{
	using R = decltype(a + b);
	return (R) a > (R) b ? (R) a : (R) b;
}

#ifndef __CUDACC__
struct dim3
{
	unsigned x = 1, y = 1, z = 1;
};
#endif
}  // namespace TNL
