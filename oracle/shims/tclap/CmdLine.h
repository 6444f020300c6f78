// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the TCLAP command line library.
// The reference (pawelswoboda/LP_MP) includes "tclap/CmdLine.h" (include/config.hxx:13,
// include/LP_MP.h:28, include/solver.hxx:14, include/visitors/standard_visitor.hxx:7) but the
// submodule is not vendored. This shim provides exactly the surface those headers use so that
// the reference's own scheduler/solver/visitor code compiles unmodified for the oracle.
// It is our own code, not a copy of TCLAP.
#pragma once
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace TCLAP {

class ArgException : public std::exception {
public:
  explicit ArgException(const std::string& text = "", const std::string& id = "") : text_(text), id_(id) {}
  std::string error() const { return text_; }
  std::string argId() const { return id_; }
  const char* what() const noexcept override { return text_.c_str(); }
private:
  std::string text_, id_;
};

template <typename T>
class Constraint {
public:
  virtual ~Constraint() {}
  virtual std::string description() const = 0;
  virtual std::string shortID() const = 0;
  virtual bool check(const T& value) const = 0;
};

class Arg {
public:
  Arg(const std::string& flag, const std::string& name) : flag_(flag), name_(name), set_(false) {}
  virtual ~Arg() {}
  virtual bool takes_value() const = 0;
  virtual void assign(const std::string& text) = 0;
  bool matches(const std::string& token) const {
    return (!name_.empty() && token == "--" + name_) || (!flag_.empty() && token == "-" + flag_);
  }
  bool isSet() const { return set_; }
  const std::string& getName() const { return name_; }
protected:
  std::string flag_, name_;
  bool set_;
};

class CmdLine {
public:
  CmdLine(const std::string& message, char delimiter = ' ', const std::string& version = "none", bool help = true) {
    (void)message; (void)delimiter; (void)version; (void)help;
  }
  void add(Arg& a) { args_.push_back(&a); }
  void add(Arg* a) { args_.push_back(a); }
  void parse(int argc, const char* const* argv) {
    std::vector<std::string> v;
    for (int i = 0; i < argc; ++i) v.emplace_back(argv[i]);
    parse(v);
  }
  // first entry is the program name, as with argv
  void parse(std::vector<std::string>& tokens) {
    for (std::size_t i = 1; i < tokens.size(); ++i) {
      Arg* hit = nullptr;
      for (Arg* a : args_) if (a->matches(tokens[i])) { hit = a; break; }
      if (hit == nullptr) throw ArgException("unknown argument " + tokens[i], tokens[i]);
      if (hit->takes_value()) {
        if (i + 1 >= tokens.size()) throw ArgException("missing value", tokens[i]);
        hit->assign(tokens[++i]);
      } else {
        hit->assign("");
      }
    }
  }
private:
  std::vector<Arg*> args_;
};

namespace detail {
template <typename T> inline T convert(const std::string& s) {
  std::istringstream is(s);
  T v{};
  is >> v;
  if (is.fail()) throw ArgException("cannot parse value '" + s + "'");
  return v;
}
template <> inline std::string convert<std::string>(const std::string& s) { return s; }
}  // namespace detail

template <typename T>
class ValueArg : public Arg {
public:
  ValueArg(const std::string& flag, const std::string& name, const std::string& desc, bool req, T value,
           const std::string& typeDesc)
      : Arg(flag, name), value_(value), constraint_(nullptr) { (void)desc; (void)req; (void)typeDesc; }
  ValueArg(const std::string& flag, const std::string& name, const std::string& desc, bool req, T value,
           const std::string& typeDesc, CmdLine& cmd)
      : ValueArg(flag, name, desc, req, value, typeDesc) { cmd.add(*this); }
  ValueArg(const std::string& flag, const std::string& name, const std::string& desc, bool req, T value,
           Constraint<T>* constraint)
      : Arg(flag, name), value_(value), constraint_(constraint) { (void)desc; (void)req; }
  ValueArg(const std::string& flag, const std::string& name, const std::string& desc, bool req, T value,
           Constraint<T>* constraint, CmdLine& cmd)
      : ValueArg(flag, name, desc, req, value, constraint) { cmd.add(*this); }
  bool takes_value() const override { return true; }
  void assign(const std::string& text) override {
    T v = detail::convert<T>(text);
    if (constraint_ != nullptr && !constraint_->check(v))
      throw ArgException("value violates " + constraint_->description() + " (value text '" + text + "', arg " + name_ + ")", name_);
    value_ = v;
    set_ = true;
  }
  T& getValue() { return value_; }
  const T& getValue() const { return value_; }
private:
  T value_;
  Constraint<T>* constraint_;
};

class SwitchArg : public Arg {
public:
  SwitchArg(const std::string& flag, const std::string& name, const std::string& desc, bool def = false)
      : Arg(flag, name), value_(def) { (void)desc; }
  SwitchArg(const std::string& flag, const std::string& name, const std::string& desc, CmdLine& cmd, bool def = false)
      : Arg(flag, name), value_(def) { (void)desc; cmd.add(*this); }
  bool takes_value() const override { return false; }
  void assign(const std::string&) override { value_ = !value_; set_ = true; }
  bool getValue() const { return value_; }
private:
  bool value_;
};

}  // namespace TCLAP
