// TEST INFRASTRUCTURE — stand-in for the third-party `leap` math library
// (github.com/pniekamp/leap, un-pinned sibling checkout `../leap`, see
// reference CMakeLists.txt:35-52 and README.md:31).  leap is not vendored in
// /root/reference, so this header restates the small part of its published
// API that src/math/{vec,color,transform}.h and tools/{ibl,hdr}.cpp use.
// It exists only so that the UNMODIFIED reference sources compile into
// oracle/_ref/ as a parity checker.  Semantics assumed:
//   lerp(a,b,t) = (1-t)*a + t*b        normalise(v) = v / norm(v)
//   fmod2(a,b)  = positive modulo      clamp(v,lo,hi) = min(max(v,lo),hi)
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <algorithm>
#include <iosfwd>
#include <ostream>

namespace leap { namespace lml
{
  template<typename T> constexpr T pi() { return T(3.14159265358979323846264338327950288L); }

  template<typename T> constexpr T clamp(T value, T lower, T upper) { return std::max(lower, std::min(value, upper)); }

  template<typename T, typename std::enable_if<std::is_arithmetic<T>::value>::type* = nullptr>
  constexpr T lerp(T lower, T upper, T alpha) { return (1 - alpha)*lower + alpha*upper; }

  template<typename T> inline T fmod2(T numer, T denom)
  {
    T r = std::fmod(numer, denom);
    return (r < 0) ? r + denom : r;
  }

  //|-------------------- VectorView ----------------------------------------
  // CRTP view over `sizeof...(Indices)` consecutive members of `Vector`
  template<typename Vector, typename T, size_t... Indices>
  class VectorView
  {
    public:
      typedef T value_type;
      static constexpr size_t size() { return sizeof...(Indices); }

      constexpr T const &operator[](size_t i) const { return reinterpret_cast<T const*>(static_cast<Vector const*>(this))[i]; }
      T &operator[](size_t i) { return reinterpret_cast<T*>(static_cast<Vector*>(this))[i]; }

      constexpr Vector const &self() const { return *static_cast<Vector const*>(this); }

    protected:
      VectorView() = default;
  };

  namespace detail
  {
    template<typename Vector, typename T, size_t... I, typename F>
    constexpr Vector map1(VectorView<Vector, T, I...> const &u, F f) { return Vector(f(u[I])...); }

    template<typename Vector, typename T, size_t... I, typename F>
    constexpr Vector map2(VectorView<Vector, T, I...> const &u, VectorView<Vector, T, I...> const &v, F f) { return Vector(f(u[I], v[I])...); }
  }

  template<typename Vector, typename T, size_t... I>
  constexpr Vector operator -(VectorView<Vector, T, I...> const &u) { return Vector((-u[I])...); }

  template<typename Vector, typename T, size_t... I>
  constexpr Vector operator +(VectorView<Vector, T, I...> const &u, VectorView<Vector, T, I...> const &v) { return Vector((u[I] + v[I])...); }

  template<typename Vector, typename T, size_t... I>
  constexpr Vector operator -(VectorView<Vector, T, I...> const &u, VectorView<Vector, T, I...> const &v) { return Vector((u[I] - v[I])...); }

  template<typename Vector, typename T, size_t... I, typename S, typename std::enable_if<std::is_arithmetic<S>::value>::type* = nullptr>
  constexpr Vector operator *(S s, VectorView<Vector, T, I...> const &v) { return Vector((T(s) * v[I])...); }

  template<typename Vector, typename T, size_t... I, typename S, typename std::enable_if<std::is_arithmetic<S>::value>::type* = nullptr>
  constexpr Vector operator *(VectorView<Vector, T, I...> const &v, S s) { return Vector((v[I] * T(s))...); }

  template<typename Vector, typename T, size_t... I, typename S, typename std::enable_if<std::is_arithmetic<S>::value>::type* = nullptr>
  constexpr Vector operator /(VectorView<Vector, T, I...> const &v, S s) { return Vector((v[I] / T(s))...); }

  template<typename Vector, typename T, size_t... I>
  Vector &operator +=(VectorView<Vector, T, I...> &u, VectorView<Vector, T, I...> const &v)
  {
    int unused[] = { ((u[I] += v[I]), 0)... }; (void)unused;
    return static_cast<Vector&>(u);
  }

  template<typename Vector, typename T, size_t... I>
  Vector &operator -=(VectorView<Vector, T, I...> &u, VectorView<Vector, T, I...> const &v)
  {
    int unused[] = { ((u[I] -= v[I]), 0)... }; (void)unused;
    return static_cast<Vector&>(u);
  }

  template<typename Vector, typename T, size_t... I, typename S, typename std::enable_if<std::is_arithmetic<S>::value>::type* = nullptr>
  Vector &operator *=(VectorView<Vector, T, I...> &u, S s)
  {
    int unused[] = { ((u[I] *= T(s)), 0)... }; (void)unused;
    return static_cast<Vector&>(u);
  }

  template<typename Vector, typename T, size_t... I, typename S, typename std::enable_if<std::is_arithmetic<S>::value>::type* = nullptr>
  Vector &operator /=(VectorView<Vector, T, I...> &u, S s)
  {
    int unused[] = { ((u[I] /= T(s)), 0)... }; (void)unused;
    return static_cast<Vector&>(u);
  }

  template<typename Vector, typename T, size_t... I>
  constexpr bool operator ==(VectorView<Vector, T, I...> const &u, VectorView<Vector, T, I...> const &v)
  {
    bool result = true;
    bool unused[] = { (result = result && (u[I] == v[I]))... }; (void)unused;
    return result;
  }

  template<typename Vector, typename T, size_t... I>
  constexpr bool operator !=(VectorView<Vector, T, I...> const &u, VectorView<Vector, T, I...> const &v) { return !(u == v); }

  // hadamard (component-wise) product
  template<typename Vector, typename T, size_t... I>
  constexpr Vector hada(VectorView<Vector, T, I...> const &u, VectorView<Vector, T, I...> const &v) { return Vector((u[I] * v[I])...); }

  template<typename Vector, typename T, size_t... I>
  constexpr T dot(VectorView<Vector, T, I...> const &u, VectorView<Vector, T, I...> const &v)
  {
    T result = 0;
    int unused[] = { ((result += u[I] * v[I]), 0)... }; (void)unused;
    return result;
  }

  template<typename Vector, typename T, size_t... I>
  constexpr T normsqr(VectorView<Vector, T, I...> const &v) { return dot(v, v); }

  template<typename Vector, typename T, size_t... I>
  constexpr T norm(VectorView<Vector, T, I...> const &v) { return std::sqrt(dot(v, v)); }

  template<typename Vector, typename T, size_t... I>
  constexpr Vector normalise(VectorView<Vector, T, I...> const &v) { return v / norm(v); }

  template<typename Vector, typename T, size_t... I>
  constexpr Vector abs(VectorView<Vector, T, I...> const &v) { return Vector(std::abs(v[I])...); }

  template<typename Vector, typename T, size_t... I>
  constexpr Vector frac(VectorView<Vector, T, I...> const &v) { return Vector((v[I] - std::floor(v[I]))...); }

  template<typename Vector, typename T, size_t... I>
  constexpr Vector lerp(VectorView<Vector, T, I...> const &lower, VectorView<Vector, T, I...> const &upper, T alpha) { return (1 - alpha)*lower + alpha*upper; }

  // 3d cross product
  template<typename Vector, typename T, size_t I0, size_t I1, size_t I2>
  constexpr Vector cross(VectorView<Vector, T, I0, I1, I2> const &u, VectorView<Vector, T, I0, I1, I2> const &v)
  {
    return Vector(u[1]*v[2] - u[2]*v[1], u[2]*v[0] - u[0]*v[2], u[0]*v[1] - u[1]*v[0]);
  }

  // component of u orthogonal to v (used by Transform::lookat only)
  template<typename Vector, typename T, size_t... I>
  constexpr Vector orthogonal(VectorView<Vector, T, I...> const &u, VectorView<Vector, T, I...> const &v) { return cross(u, v); }

  template<typename Vector, typename T, size_t... I>
  std::ostream &operator <<(std::ostream &os, VectorView<Vector, T, I...> const &v)
  {
    os << "(";
    int unused[] = { ((os << (I ? "," : "") << v[I]), 0)... }; (void)unused;
    os << ")";
    return os;
  }

  inline float frac(float v) { return v - std::floor(v); }
} }
