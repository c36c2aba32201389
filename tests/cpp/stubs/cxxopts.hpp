// tests/cpp/stubs/cxxopts.hpp -- TEST INFRASTRUCTURE. cxxopts is not installed (and cannot be
// fetched); the reference's examples/spmv/helpers.hxx needs only this much of it: Options with a
// chained add_options()("s,long", "help"[, value<T>()]), parse(argc, argv) -> result with
// count("long") / ["long"].as<T>(), and help({""}). Lets the reference's example mains compile
// UNCHANGED against this repo's include/ tree.
#pragma once
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cxxopts {
struct value_base { virtual ~value_base() = default; };
template <typename T> struct typed_value : value_base {};
template <typename T> std::shared_ptr<value_base> value() { return std::make_shared<typed_value<T>>(); }

struct option_value {
  std::string text;
  template <typename T> T as() const {
    std::istringstream in(text);
    T v{};
    if constexpr (std::is_same_v<T, std::string>) return text;
    else { in >> v; return v; }
  }
};

struct parse_result {
  std::map<std::string, std::vector<std::string>> seen;
  std::size_t count(const std::string& name) const { auto it = seen.find(name); return it == seen.end() ? 0 : it->second.size(); }
  option_value operator[](const std::string& name) const {
    auto it = seen.find(name);
    if (it == seen.end() || it->second.empty()) throw std::runtime_error("option not given: " + name);
    return option_value{it->second.back()};
  }
};

class Options {
  struct spec { std::string short_name, long_name, help; bool takes_value; };
  std::vector<spec> specs_;
  std::string program_, description_;

 public:
  Options(std::string program, std::string description = "") : program_(std::move(program)), description_(std::move(description)) {}
  struct adder {
    Options& o;
    adder& operator()(const std::string& names, const std::string& help, std::shared_ptr<value_base> v = nullptr) {
      spec s;
      const std::size_t comma = names.find(',');
      if (comma == std::string::npos) s.long_name = names;
      else { s.short_name = names.substr(0, comma); s.long_name = names.substr(comma + 1); }
      s.help = help;
      s.takes_value = v != nullptr;
      o.specs_.push_back(s);
      return *this;
    }
  };
  adder add_options() { return adder{*this}; }
  parse_result parse(int argc, char** argv) const {
    parse_result r;
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i], name, inline_value;
      bool has_inline = false;
      if (a.rfind("--", 0) == 0) {
        name = a.substr(2);
        const std::size_t eq = name.find('=');
        if (eq != std::string::npos) { inline_value = name.substr(eq + 1); name = name.substr(0, eq); has_inline = true; }
      } else if (a.rfind("-", 0) == 0 && a.size() >= 2) {
        name = a.substr(1);
      } else continue;
      const spec* hit = nullptr;
      for (const spec& s : specs_) if (s.long_name == name || s.short_name == name) hit = &s;
      if (!hit) throw std::runtime_error("unknown option: " + a);
      std::string v;
      if (hit->takes_value) {
        if (has_inline) v = inline_value;
        else if (i + 1 < argc) v = argv[++i];
        else throw std::runtime_error("option needs a value: " + a);
      }
      r.seen[hit->long_name].push_back(v);
    }
    return r;
  }
  std::string help(const std::vector<std::string>& = {}) const {
    std::ostringstream out;
    out << description_ << "\nUsage:\n  " << program_ << " [OPTION...]\n\n";
    for (const spec& s : specs_)
      out << "  " << (s.short_name.empty() ? "    " : "-" + s.short_name + ", ") << "--" << s.long_name
          << (s.takes_value ? " arg" : "") << "  " << s.help << "\n";
    return out.str();
  }
};
}  // namespace cxxopts
