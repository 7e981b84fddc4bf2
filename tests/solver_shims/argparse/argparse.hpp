// Minimal stand-in for p-ranav/argparse v3.2 (fetched by the reference's CMake, absent here).  TEST INFRASTRUCTURE ONLY:
// positional integer arguments with defaults, as the reference's sim_1.cu / sim2d_1.cu use them.
#pragma once
#include <cstdlib>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
namespace argparse {
class Argument
{
public:
	std::string name, help_;
	int value = 0;
	bool has_default = false, is_flag = false, set = false;
	Argument& help(const std::string& h) { help_ = h; return *this; }
	template <char, typename T> Argument& scan() { return *this; }
	Argument& default_value(int v) { value = v; has_default = true; return *this; }
	Argument& flag() { is_flag = true; return *this; }
};
class ArgumentParser
{
	std::string prog, desc;
	std::vector<Argument> args;
public:
	explicit ArgumentParser(std::string p) : prog(std::move(p)) {}
	void add_description(const std::string& d) { desc = d; }
	Argument& add_argument(const std::string& n) { args.emplace_back(); args.back().name = n; return args.back(); }
	void parse_args(int argc, char** argv)
	{
		size_t pos = 0;
		for (int i = 1; i < argc; i++) {
			std::string a = argv[i];
			while (pos < args.size() && args[pos].name.rfind("--", 0) == 0) pos++;
			if (pos >= args.size()) throw std::runtime_error("unexpected argument " + a);
			args[pos].value = std::atoi(a.c_str());
			args[pos].set = true;
			pos++;
		}
		for (auto& a : args)
			if (! a.set && ! a.has_default && a.name.rfind("--", 0) != 0) throw std::runtime_error("missing argument " + a.name);
	}
	template <typename T> T get(const std::string& n) const
	{
		for (auto& a : args) if (a.name == n) return (T) a.value;
		throw std::logic_error("no such argument " + n);
	}
	friend std::ostream& operator<<(std::ostream& os, const ArgumentParser& p) { return os << "usage: " << p.prog << " ...\n" << p.desc << "\n"; }
};
}  // namespace argparse
