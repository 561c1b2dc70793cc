// TEST INFRASTRUCTURE — stand-in for leap/lml/quaternion.h (see vector.h in this
// directory).  Quaternion(axis, angle) = (cos(angle/2), axis*sin(angle/2)); the
// convention is confirmed in-repo by the literal face quaternions of
// data/project.comp:27-32 and the algebra of data/transform.inc:13-37.
#pragma once
#include "vector.h"

namespace leap { namespace lml
{
  template<typename T, typename V>
  class Quaternion
  {
    public:
      Quaternion() = default;
      constexpr Quaternion(T w, T x, T y, T z) : w(w), x(x), y(y), z(z) { }
      constexpr Quaternion(T scalar, V const &vector) : w(scalar), x(vector.x), y(vector.y), z(vector.z) { }
      constexpr Quaternion(V const &axis, T angle) : w(std::cos(angle/2)), x(axis.x * std::sin(angle/2)), y(axis.y * std::sin(angle/2)), z(axis.z * std::sin(angle/2)) { }
      Quaternion(V const &xaxis, V const &yaxis, V const &zaxis); // basis form: not used on the IBL path

    union
    {
      struct
      {
        T w;
        T x;
        T y;
        T z;
      };

      struct
      {
        T scalar;
        V vector;
      };

      struct
      {
        T pad_;
        V xyz;
      };
    };
  };

  template<typename T, typename V>
  constexpr Quaternion<T, V> conjugate(Quaternion<T, V> const &q) { return Quaternion<T, V>(q.w, -q.x, -q.y, -q.z); }

  template<typename T, typename V>
  constexpr Quaternion<T, V> operator +(Quaternion<T, V> const &q1, Quaternion<T, V> const &q2) { return Quaternion<T, V>(q1.w + q2.w, q1.x + q2.x, q1.y + q2.y, q1.z + q2.z); }

  template<typename T, typename V>
  constexpr Quaternion<T, V> operator -(Quaternion<T, V> const &q1, Quaternion<T, V> const &q2) { return Quaternion<T, V>(q1.w - q2.w, q1.x - q2.x, q1.y - q2.y, q1.z - q2.z); }

  template<typename T, typename V>
  constexpr Quaternion<T, V> operator -(Quaternion<T, V> const &q) { return Quaternion<T, V>(-q.w, -q.x, -q.y, -q.z); }

  // hamilton product, same term order as data/transform.inc:19-27
  template<typename T, typename V>
  constexpr Quaternion<T, V> operator *(Quaternion<T, V> const &q1, Quaternion<T, V> const &q2)
  {
    return Quaternion<T, V>(
      q1.w*q2.w - q1.x*q2.x - q1.y*q2.y - q1.z*q2.z,
      q1.w*q2.x + q1.x*q2.w + q1.y*q2.z - q1.z*q2.y,
      q1.w*q2.y + q1.y*q2.w + q1.z*q2.x - q1.x*q2.z,
      q1.w*q2.z + q1.z*q2.w + q1.x*q2.y - q1.y*q2.x);
  }

  template<typename T, typename V, typename S, typename std::enable_if<std::is_arithmetic<S>::value>::type* = nullptr>
  constexpr Quaternion<T, V> operator *(S s, Quaternion<T, V> const &q) { return Quaternion<T, V>(T(s)*q.w, T(s)*q.x, T(s)*q.y, T(s)*q.z); }

  template<typename T, typename V, typename S, typename std::enable_if<std::is_arithmetic<S>::value>::type* = nullptr>
  constexpr Quaternion<T, V> operator *(Quaternion<T, V> const &q, S s) { return Quaternion<T, V>(q.w*T(s), q.x*T(s), q.y*T(s), q.z*T(s)); }

  template<typename T, typename V, typename S, typename std::enable_if<std::is_arithmetic<S>::value>::type* = nullptr>
  constexpr Quaternion<T, V> operator /(Quaternion<T, V> const &q, S s) { return Quaternion<T, V>(q.w/T(s), q.x/T(s), q.y/T(s), q.z/T(s)); }

  template<typename T, typename V>
  constexpr bool operator ==(Quaternion<T, V> const &q1, Quaternion<T, V> const &q2) { return q1.w == q2.w && q1.x == q2.x && q1.y == q2.y && q1.z == q2.z; }

  template<typename T, typename V>
  constexpr T dot(Quaternion<T, V> const &q1, Quaternion<T, V> const &q2) { return q1.w*q2.w + q1.x*q2.x + q1.y*q2.y + q1.z*q2.z; }

  template<typename T, typename V>
  constexpr T norm(Quaternion<T, V> const &q) { return std::sqrt(dot(q, q)); }

  template<typename T, typename V>
  constexpr Quaternion<T, V> normalise(Quaternion<T, V> const &q) { return q / norm(q); }

  template<typename T, typename V>
  constexpr Quaternion<T, V> lerp(Quaternion<T, V> const &q1, Quaternion<T, V> const &q2, T alpha) { return (1 - alpha)*q1 + alpha*q2; }

  template<typename T, typename V>
  Quaternion<T, V> slerp(Quaternion<T, V> const &q1, Quaternion<T, V> const &q2, T alpha); // not used on the IBL path

  template<typename T, typename V>
  std::ostream &operator <<(std::ostream &os, Quaternion<T, V> const &q) { os << "(" << q.w << "," << q.x << "," << q.y << "," << q.z << ")"; return os; }

  // 4x4 matrix: only Transform::matrix() (not on the IBL path) touches it
  template<typename T, size_t M, size_t N>
  class Matrix
  {
    public:
      T &operator()(size_t i, size_t j) { return data[i][j]; }
      T const &operator()(size_t i, size_t j) const { return data[i][j]; }
      T data[M][N];
  };

  typedef Matrix<float, 4, 4> Matrix4f;
} }
