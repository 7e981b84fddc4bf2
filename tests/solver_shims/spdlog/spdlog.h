// Minimal stand-in for spdlog (fetched by the reference's CMake, absent here).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <fmt/core.h>
namespace spdlog {
template <typename... A> void info(const char* s, const A&... a) { fmt::println(stdout, s, a...); }
template <typename... A> void warn(const char* s, const A&... a) { fmt::println(stderr, s, a...); }
template <typename... A> void error(const char* s, const A&... a) { fmt::println(stderr, s, a...); }
template <typename... A> void trace(const char*, const A&...) {}
template <typename... A> void debug(const char*, const A&...) {}
}  // namespace spdlog
