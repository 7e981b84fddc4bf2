// Minimal stand-in for {fmt} (the reference fetches fmt 12.0.0 at configure time, CMakeLists.txt:120-166; absent here).
// TEST INFRASTRUCTURE ONLY: lets tests/test_dropin_solvers.py compile the reference's unmodified solver sources.
// Supports "{}" and "{:0Nd}"-style fields, which is all sim_1.cu / sim2d_1.cu use.
#pragma once
#include <cstdio>
#include <sstream>
#include <string>
#include <iomanip>
namespace fmt {
namespace detail {
inline void emit(std::ostringstream& os, const std::string& spec, ...) { (void) spec; (void) os; }
template <typename T>
void put(std::ostringstream& os, const std::string& spec, const T& v)
{
	std::ios_base::fmtflags f = os.flags();
	char fill = os.fill();
	size_t i = 0;
	if (i < spec.size() && spec[i] == '0') { os << std::setfill('0'); i++; }
	size_t w = 0;
	while (i < spec.size() && spec[i] >= '0' && spec[i] <= '9') w = w * 10 + (spec[i++] - '0');
	if (w) os << std::setw((int) w);
	os << v;
	os.flags(f);
	os.fill(fill);
}
inline void format_to(std::ostringstream& os, const char* s)
{
	os << s;
}
template <typename T, typename... R>
void format_to(std::ostringstream& os, const char* s, const T& v, const R&... rest)
{
	for (; *s; s++) {
		if (*s == '{') {
			const char* e = s;
			while (*e && *e != '}') e++;
			std::string spec(s + 1, e);
			if (! spec.empty() && spec[0] == ':') spec = spec.substr(1);
			put(os, spec, v);
			format_to(os, *e ? e + 1 : e, rest...);
			return;
		}
		os << *s;
	}
}
}  // namespace detail
template <typename... A>
std::string format(const char* s, const A&... a)
{
	std::ostringstream os;
	detail::format_to(os, s, a...);
	return os.str();
}
template <typename... A>
void print(const char* s, const A&... a) { std::fputs(format(s, a...).c_str(), stdout); }
template <typename... A>
void println(std::FILE* f, const char* s, const A&... a) { std::fputs((format(s, a...) + "\n").c_str(), f); }
template <typename... A>
void println(const char* s, const A&... a) { println(stdout, s, a...); }
}  // namespace fmt
