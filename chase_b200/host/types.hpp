// chase_b200 host layer — scalar traits (API mirror of the reference's
// algorithm/types.hpp:32-304: chase::Base<T>, conjugate, getRandomT).
#pragma once
#include <complex>
#include <cstddef>
#include <random>
#include <type_traits>

namespace chase
{

template <class Q>
struct BaseOf
{
    using type = Q;
};
template <class Q>
struct BaseOf<std::complex<Q>>
{
    using type = Q;
};
template <class Q>
using Base = typename BaseOf<Q>::type;

template <class T>
struct is_complex_t : std::false_type
{
};
template <class T>
struct is_complex_t<std::complex<T>> : std::true_type
{
};

template <class T>
inline T conjugate(const T& x)
{
    if constexpr (is_complex_t<T>::value)
        return std::conj(x);
    else
        return x;
}

// Draws one scalar from `f`.  For complex types the reference writes
// `std::complex<T>(f(), f())` (types.hpp:262-271) which g++ evaluates right to
// left: the FIRST draw becomes the imaginary part.  Reproduced explicitly here
// so the start vectors match the reference CPU backend bit for bit.
template <class T, class F>
inline T getRandomT(F&& f)
{
    if constexpr (is_complex_t<T>::value)
    {
        const auto im = f();
        const auto re = f();
        return T(static_cast<Base<T>>(re), static_cast<Base<T>>(im));
    }
    else
    {
        return static_cast<T>(f());
    }
}

// Matrix-type tags selecting the problem class of a backend, named like the reference's containers
// (linalg/matrix/matrix.hpp:1202 Matrix<T, GPU>, :1597 PseudoHermitianMatrix<T, GPU>).  Here they carry no storage: the backends own
// their device buffers and take the caller's raw host pointers.
namespace platform
{
struct CPU
{
};
struct GPU
{
};
} // namespace platform
namespace matrix
{
template <class T, class Platform = chase::platform::GPU>
struct Matrix
{
    using value_type = T;
    using platform_type = Platform;
};
template <class T, class Platform = chase::platform::GPU>
struct PseudoHermitianMatrix
{
    using value_type = T;
    using platform_type = Platform;
};
} // namespace matrix

// type code used by the C-ABI kernel layer: 0 s, 1 d, 2 c, 3 z
template <class T>
constexpr int type_code()
{
    if constexpr (std::is_same<T, float>::value)
        return 0;
    else if constexpr (std::is_same<T, double>::value)
        return 1;
    else if constexpr (std::is_same<T, std::complex<float>>::value)
        return 2;
    else
        return 3;
}

} // namespace chase
