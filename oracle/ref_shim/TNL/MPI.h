#pragma once
namespace TNL::MPI {
struct ScopedInitializer
{
	ScopedInitializer(int&, char**&) {}
};
}  // namespace TNL::MPI
