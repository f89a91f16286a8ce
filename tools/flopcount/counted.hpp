/*
 * counted.hpp — MEASUREMENT TOOL (tools/count_flops.py), never part of the product or of the test oracle.
 *
 * SURVEY.md section 8(d) asks for the algorithmic flops of the model stages "counted from the oracle built with an
 * instrumented scalar type". This header is force-included (g++ -include) in front of every oracle/ source: it pulls in the
 * standard headers first, then re-reads the word `double` as `Counted`, a one-double struct whose operators tally every
 * addition, multiplication, division, square root, transcendental call and comparison into the bucket the driver has
 * selected. The oracle sources are compiled unchanged; the arithmetic (and so every result) is the oracle's own.
 * CasADi functions of the reference are separately compiled C (oracle/_ref): their calls are tallied per function and
 * priced with the instruction counts of the generated code (count_flops.py).
 */
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace flopcount {
typedef double real_t;
enum Kind { ADD = 0, MUL, DIV, SQRT, TRANS, CMP, NKIND };
constexpr int NBUCKET = 64;
inline long long g_cnt[NBUCKET][NKIND];
inline int g_bucket = 0;
inline void tick(Kind k) { ++g_cnt[g_bucket][k]; }

struct Counted {
  real_t v;
  Counted() = default;
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  Counted(T x) : v((real_t)x) {}
  explicit operator real_t() const { return v; }
  explicit operator float() const { return (float)v; }
  explicit operator int() const { return (int)v; }
  explicit operator long() const { return (long)v; }
  explicit operator bool() const { return v != 0; }
  Counted& operator+=(const Counted& o) { tick(ADD); v += o.v; return *this; }
  Counted& operator-=(const Counted& o) { tick(ADD); v -= o.v; return *this; }
  Counted& operator*=(const Counted& o) { tick(MUL); v *= o.v; return *this; }
  Counted& operator/=(const Counted& o) { tick(DIV); v /= o.v; return *this; }
};
static_assert(sizeof(Counted) == sizeof(real_t) && std::is_trivially_copyable<Counted>::value, "Counted must alias a double");

inline Counted operator-(const Counted& a) { Counted r; r.v = -a.v; return r; }   // sign flips are free
inline Counted operator+(const Counted& a) { return a; }
#define FC_ARITH(T) typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0
#define FC_BIN(op, kind)                                                                                                          \
  inline Counted operator op(const Counted& a, const Counted& b) { tick(kind); Counted r; r.v = a.v op b.v; return r; }           \
  template <class T, FC_ARITH(T)> inline Counted operator op(const Counted& a, T b) { tick(kind); Counted r; r.v = a.v op (real_t)b; return r; } \
  template <class T, FC_ARITH(T)> inline Counted operator op(T a, const Counted& b) { tick(kind); Counted r; r.v = (real_t)a op b.v; return r; }
FC_BIN(+, ADD)
FC_BIN(-, ADD)
FC_BIN(*, MUL)
FC_BIN(/, DIV)
#define FC_CMP(op)                                                                                              \
  inline bool operator op(const Counted& a, const Counted& b) { tick(CMP); return a.v op b.v; }                 \
  template <class T, FC_ARITH(T)> inline bool operator op(const Counted& a, T b) { tick(CMP); return a.v op (real_t)b; } \
  template <class T, FC_ARITH(T)> inline bool operator op(T a, const Counted& b) { tick(CMP); return (real_t)a op b.v; }
FC_CMP(<)
FC_CMP(>)
FC_CMP(<=)
FC_CMP(>=)
FC_CMP(==)
FC_CMP(!=)
}  // namespace flopcount
// math functions live outside Counted's namespace: the oracle has its own unqualified sin / cos overloads, ADL must not see these
namespace flopcount_fn {
using namespace flopcount;
#define FC_FN1(name, kind) inline Counted name(const Counted& a) { tick(kind); Counted r; r.v = std::name(a.v); return r; }
FC_FN1(sqrt, SQRT)
FC_FN1(sin, TRANS)
FC_FN1(cos, TRANS)
FC_FN1(tan, TRANS)
FC_FN1(exp, TRANS)
FC_FN1(log, TRANS)
FC_FN1(atan, TRANS)
FC_FN1(asin, TRANS)
FC_FN1(acos, TRANS)
inline Counted fabs(const Counted& a) { Counted r; r.v = std::fabs(a.v); return r; }
inline Counted abs(const Counted& a) { Counted r; r.v = std::fabs(a.v); return r; }
inline Counted pow(const Counted& a, const Counted& b) { tick(TRANS); Counted r; r.v = std::pow(a.v, b.v); return r; }
inline Counted atan2(const Counted& a, const Counted& b) { tick(TRANS); Counted r; r.v = std::atan2(a.v, b.v); return r; }
inline Counted fmax(const Counted& a, const Counted& b) { tick(CMP); Counted r; r.v = std::fmax(a.v, b.v); return r; }
inline Counted fmin(const Counted& a, const Counted& b) { tick(CMP); Counted r; r.v = std::fmin(a.v, b.v); return r; }
inline bool isfinite(const Counted& a) { return std::isfinite(a.v); }
inline bool isnan(const Counted& a) { return std::isnan(a.v); }
template <class T, FC_ARITH(T)> inline Counted max(const Counted& a, T b) { tick(CMP); return a.v < (real_t)b ? Counted(b) : a; }
template <class T, FC_ARITH(T)> inline Counted max(T a, const Counted& b) { tick(CMP); return (real_t)a < b.v ? b : Counted(a); }
template <class T, FC_ARITH(T)> inline Counted min(const Counted& a, T b) { tick(CMP); return (real_t)b < a.v ? Counted(b) : a; }
template <class T, FC_ARITH(T)> inline Counted min(T a, const Counted& b) { tick(CMP); return b.v < (real_t)a ? b : Counted(a); }
}  // namespace flopcount_fn

// the oracle calls these with the std:: qualifier
namespace std {
using flopcount_fn::sqrt; using flopcount_fn::sin; using flopcount_fn::cos; using flopcount_fn::tan; using flopcount_fn::exp; using flopcount_fn::log;
using flopcount_fn::atan; using flopcount_fn::asin; using flopcount_fn::acos; using flopcount_fn::fabs; using flopcount_fn::abs; using flopcount_fn::pow;
using flopcount_fn::atan2; using flopcount_fn::fmax; using flopcount_fn::fmin; using flopcount_fn::isfinite; using flopcount_fn::isnan;
using flopcount_fn::max; using flopcount_fn::min;
}  // namespace std

using flopcount::Counted;
#define double Counted
