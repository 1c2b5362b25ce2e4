// chase_b200 host layer — typed C++ view of the kernel C ABI
// (include/chase_b200_kernels.h): K<T>::gemm(...) -> chase_b200_gemm_<s|d|c|z>.
#pragma once
#include "../../include/chase_b200_kernels.h"
#include "types.hpp"

#include <complex>

namespace chase
{
namespace b200
{

template <class T>
struct K;

#define CB2_HOST_K(X, TT)                                                                                              \
    template <>                                                                                                        \
    struct K<TT>                                                                                                       \
    {                                                                                                                  \
        static constexpr auto gemm = chase_b200_gemm_##X;                                                              \
        static constexpr auto hemm = chase_b200_hemm_##X;                                                              \
        static constexpr auto hemm_rect = chase_b200_hemm_rect_##X;                                                    \
        static constexpr auto gather_rows = chase_b200_gather_rows_##X;                                                \
        static constexpr auto axpy_cols = chase_b200_axpy_cols_##X;                                                    \
        static constexpr auto shift_diag_list = chase_b200_shift_diag_list_##X;                                        \
        static constexpr auto rng_normal_rows = chase_b200_rng_normal_rows_##X;                                        \
        static constexpr auto potrf = chase_b200_potrf_##X;                                                            \
        static constexpr auto trsm = chase_b200_trsm_##X;                                                              \
        static constexpr auto shift_abstrace = chase_b200_shift_abstrace_##X;                                          \
        static constexpr auto heev = chase_b200_heev_##X;                                                              \
        static constexpr auto colnorms = chase_b200_colnorms_##X;                                                      \
        static constexpr auto lacpy = chase_b200_lacpy_##X;                                                            \
        static constexpr auto tri_pack = chase_b200_tri_pack_##X;                                                      \
        static constexpr auto tri_unpack = chase_b200_tri_unpack_##X;                                                  \
        static constexpr auto gather_cols = chase_b200_gather_cols_##X;                                                \
        static constexpr auto gemv_conjt = chase_b200_gemv_conjt_##X;                                                  \
        static constexpr auto lanczos_step = chase_b200_lanczos_step_##X;                                              \
        static constexpr auto normalize_cols = chase_b200_normalize_cols_##X;                                          \
        static constexpr auto rng_normal = chase_b200_rng_normal_##X;                                                  \
        static constexpr auto herm_check = chase_b200_herm_check_##X;                                                  \
        static constexpr auto shift_diag = chase_b200_shift_diag_##X;                                                  \
        static constexpr auto herm_mirror = chase_b200_herm_mirror_##X;                                                \
        static constexpr auto hhqr = chase_b200_hhqr_##X;                                                              \
        static constexpr auto scale_rows = chase_b200_scale_rows_##X;                                                  \
        static constexpr auto scale_rows_map = chase_b200_scale_rows_map_##X;                                          \
        static constexpr auto kconj = chase_b200_kconj_##X;                                                            \
        static constexpr auto lanczos_pseudo_norm = chase_b200_lanczos_pseudo_norm_##X;                                \
        static constexpr auto lanczos_pseudo_step = chase_b200_lanczos_pseudo_step_##X;                                \
    };

CB2_HOST_K(s, float)
CB2_HOST_K(d, double)
CB2_HOST_K(c, std::complex<float>)
CB2_HOST_K(z, std::complex<double>)
#undef CB2_HOST_K

template <class T>
inline double re_of(const T& x)
{
    return (double)std::real(x);
}
template <class T>
inline double im_of(const T& x)
{
    return (double)std::imag(x);
}

} // namespace b200
} // namespace chase
