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

// Matrix types selecting the problem class of a backend, named like the reference's containers
// (linalg/matrix/matrix.hpp:1202 Matrix<T, GPU>, :1597 PseudoHermitianMatrix<T, GPU>).  Here they are views of the
// caller's host buffer (rows, cols, leading dimension, pointer): the backends own their device storage.  They serve
// as the MatrixType template argument and as the argument of the reference's second backend constructor
// (Impl/chase_gpu/chase_gpu.hpp:195-267).
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
class Matrix
{
public:
    using value_type = T;
    using platform_type = Platform;
    Matrix() = default;
    Matrix(std::size_t rows, std::size_t cols, std::size_t ld, T* host) : rows_(rows), cols_(cols), ld_(ld), data_(host) {}
    std::size_t rows() const { return rows_; }
    std::size_t cols() const { return cols_; }
    std::size_t cpu_ld() const { return ld_; }
    T* cpu_data() const { return data_; }

private:
    std::size_t rows_ = 0, cols_ = 0, ld_ = 0;
    T* data_ = nullptr;
};
template <class T, class Platform = chase::platform::GPU>
class PseudoHermitianMatrix : public Matrix<T, Platform>
{
public:
    using Matrix<T, Platform>::Matrix;
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
