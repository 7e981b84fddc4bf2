// Minimal stand-in for magic_enum (test scaffolding only; the reference fetches the real library with CMake):
// enum_name(E) and enum_cast<E>(string_view) for enums whose values lie in [0, 32), via __PRETTY_FUNCTION__ (gcc / clang).
#pragma once
#include <array>
#include <optional>
#include <string_view>
#include <utility>

namespace magic_enum {
namespace detail {
template <typename E, E V>
constexpr std::string_view raw_name()
{
	// "... [with E = Scaling; E V = Scaling::strong]" (gcc) -- an invalid value prints as "(Scaling)7"
	constexpr std::string_view s = __PRETTY_FUNCTION__;
	constexpr auto end = s.find_last_of(";]");
	constexpr auto start = s.find_last_of(" :=)", end - 1) + 1;
	return s.substr(start, end - start);
}
template <typename E, E V>
constexpr std::string_view name_or_empty()
{
	constexpr std::string_view n = raw_name<E, V>();
	return (! n.empty() && ((n[0] >= 'a' && n[0] <= 'z') || (n[0] >= 'A' && n[0] <= 'Z') || n[0] == '_')) ? n : std::string_view{};
}
template <typename E, std::size_t... I>
constexpr std::array<std::string_view, sizeof...(I)> names(std::index_sequence<I...>)
{
	return {name_or_empty<E, static_cast<E>(I)>()...};
}
}  // namespace detail

template <typename E>
constexpr std::string_view enum_name(E value)
{
	constexpr auto table = detail::names<E>(std::make_index_sequence<32>{});
	const auto i = static_cast<std::size_t>(value);
	return i < table.size() ? table[i] : std::string_view{};
}
template <typename E>
constexpr std::optional<E> enum_cast(std::string_view name)
{
	constexpr auto table = detail::names<E>(std::make_index_sequence<32>{});
	for (std::size_t i = 0; i < table.size(); i++)
		if (! table[i].empty() && table[i] == name)
			return static_cast<E>(i);
	return std::nullopt;
}
}  // namespace magic_enum
