// Minimal stand-in for p-ranav/argparse v3.2 (fetched by the reference's CMake, absent here).  TEST INFRASTRUCTURE ONLY: what the
// reference's sim_1.cu / sim_2.cu / sim2d_1.cu use -- positional and "--option value" arguments with scan<>, default_value, nargs(1),
// choices(...), flag(), get<T>().
#pragma once
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>
namespace argparse {
class Argument
{
public:
	std::string name, help_, text;	// text: the value as given on the command line or by default_value
	std::vector<std::string> allowed;
	bool has_default = false, is_flag = false, set = false;
	Argument& help(const std::string& h) { help_ = h; return *this; }
	template <char, typename T> Argument& scan() { return *this; }
	Argument& default_value(int v) { text = std::to_string(v); has_default = true; return *this; }
	Argument& default_value(double v) { text = std::to_string(v); has_default = true; return *this; }
	Argument& default_value(bool v) { text = v ? "1" : "0"; has_default = true; return *this; }
	Argument& default_value(const char* v) { text = v; has_default = true; return *this; }
	Argument& default_value(const std::string& v) { text = v; has_default = true; return *this; }
	Argument& nargs(int) { return *this; }
	Argument& required() { return *this; }
	template <typename... S> Argument& choices(S... s) { allowed = {std::string(s)...}; return *this; }
	Argument& flag() { is_flag = true; text = "0"; has_default = true; return *this; }
	Argument& implicit_value(bool) { is_flag = true; return *this; }
};
class ArgumentParser
{
	std::string prog, desc;
	std::vector<Argument> args;
public:
	explicit ArgumentParser(std::string p) : prog(std::move(p)) {}
	void add_description(const std::string& d) { desc = d; }
	Argument& add_argument(const std::string& n) { args.reserve(64); args.emplace_back(); args.back().name = n; return args.back(); }
	void parse_args(int argc, char** argv)
	{
		size_t pos = 0;
		for (int i = 1; i < argc; i++) {
			const std::string a = argv[i];
			if (a.rfind("--", 0) == 0) {
				Argument* opt = nullptr;
				for (auto& x : args) if (x.name == a) opt = &x;
				if (! opt) throw std::runtime_error("unknown option " + a);
				if (opt->is_flag) { opt->text = "1"; opt->set = true; continue; }
				if (i + 1 >= argc) throw std::runtime_error("option " + a + " needs a value");
				opt->text = argv[++i];
				opt->set = true;
				if (! opt->allowed.empty()) {
					bool ok = false;
					for (auto& c : opt->allowed) ok |= c == opt->text;
					if (! ok) throw std::runtime_error("invalid value for " + a + ": " + opt->text);
				}
				continue;
			}
			while (pos < args.size() && args[pos].name.rfind("--", 0) == 0) pos++;
			if (pos >= args.size()) throw std::runtime_error("unexpected argument " + a);
			args[pos].text = a;
			args[pos].set = true;
			pos++;
		}
		for (auto& a : args)
			if (! a.set && ! a.has_default && a.name.rfind("--", 0) != 0) throw std::runtime_error("missing argument " + a.name);
	}
	template <typename T> T get(const std::string& n) const
	{
		for (auto& a : args)
			if (a.name == n || a.name == "-" + n || a.name == "--" + n) {  // like argparse: a name without its dashes finds the option
				if constexpr (std::is_same<T, std::string>::value) return a.text;
				else if constexpr (std::is_same<T, bool>::value) return a.text == "1" || a.text == "true";
				else if constexpr (std::is_floating_point<T>::value) return (T) std::atof(a.text.c_str());
				else return (T) std::atoll(a.text.c_str());
			}
		throw std::logic_error("no such argument " + n);
	}
	friend std::ostream& operator<<(std::ostream& os, const ArgumentParser& p) { return os << "usage: " << p.prog << " ...\n" << p.desc << "\n"; }
};
}  // namespace argparse
