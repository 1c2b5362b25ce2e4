// chase_b200 — TMA-fed DMMA pipeline for the Chebyshev-filter HEMM (placeholder
// until the pipeline lands: every shape is routed to the generic DMMA tiles).
#pragma once
#include "common.cuh"

namespace cb2
{

template <class T>
inline bool hemm_tma_supported(int64_t, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t)
{
    return false;
}

template <class T>
inline int hemm_tma_launch(int64_t, int64_t, typename Traits<T>::comp, const T*, int64_t, const T*, int64_t,
                           typename Traits<T>::comp, T*, int64_t, double, const double*, cudaStream_t)
{
    return -9;
}

} // namespace cb2
