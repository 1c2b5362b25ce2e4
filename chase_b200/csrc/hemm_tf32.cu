// chase_b200 — launchers and registry of the tcgen05 kind::tf32 filter product (hemm_tf32.cuh)
#include "../../include/chase_b200_kernels.h"
#include "hemm_tf32.cuh"
#include "tf32_api.hpp"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>

using namespace cb2;

namespace
{
std::mutex g_mu;
std::map<const void*, Tf32Reg>& registry()
{
    static std::map<const void*, Tf32Reg> r;
    return r;
}
thread_local int g_terms = 3; // per host thread: one host thread drives one solver instance
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
} // namespace

namespace cb2
{
bool tf32_lookup(const void* A, Tf32Reg* out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = registry().find(A);
    if (it == registry().end())
        return false;
    *out = it->second;
    return true;
}
int tf32_terms()
{
    // CHASE_B200_TF32_TERMS=3|4 forces the number of partial products of every call (diagnostics)
    static int forced = -1;
    if (forced < 0)
    {
        const char* e = std::getenv("CHASE_B200_TF32_TERMS");
        forced = e ? std::atoi(e) : 0;
    }
    return forced >= 3 ? (forced >= 4 ? 4 : 3) : g_terms;
}
} // namespace cb2

extern "C" int chase_b200_tf32_register(const void* A, void* Alo, int64_t ld, int64_t rows, int64_t cols, int kind,
                                        void* scratch, size_t scratch_bytes)
{
    if (!A || !Alo || ld < rows || rows <= 0 || cols <= 0 || kind < 0 || kind > 2)
        return -2;
    std::lock_guard<std::mutex> lk(g_mu);
    registry()[A] = Tf32Reg{Alo, ld, rows, cols, kind, scratch, scratch_bytes};
    return 0;
}
extern "C" int chase_b200_tf32_unregister(const void* A)
{
    std::lock_guard<std::mutex> lk(g_mu);
    registry().erase(A);
    return 0;
}
// refresh the lo part after the matrix changed (upload): one pass over ld x cols elements
extern "C" int chase_b200_tf32_sync(char type, const void* A, void* stream)
{
    Tf32Reg r;
    if (!tf32_lookup(A, &r))
        return -2;
    const long long floats = (long long)r.ld * r.cols * ((type == 'c' || type == 'C') ? 2 : 1);
    if (floats % 4)
        return -2;
    const long long n4 = floats / 4;
    const int blocks = (int)std::min<long long>((n4 + 255) / 256, 148 * 16);
    tf32_split_lo_kernel<<<blocks, 256, 0, kcount(S(stream))>>>(n4, (const float4*)A, (float4*)r.lo);
    CB2_CUDA_OK(cudaGetLastError());
    return 0;
}
// refresh the lo parts of `cnt` elements given by linear element indices (device array) after an in-place update
extern "C" int chase_b200_tf32_sync_list(char type, const void* A, int64_t cnt, const int64_t* lin_dev, void* stream)
{
    Tf32Reg r;
    if (!tf32_lookup(A, &r))
        return -2;
    if (cnt <= 0)
        return 0;
    const int fpe = (type == 'c' || type == 'C') ? 2 : 1;
    const int blocks = (int)std::min<long long>((cnt + 255) / 256, 148 * 4);
    tf32_split_lo_list_kernel<<<blocks, 256, 0, kcount(S(stream))>>>(cnt, (const long long*)lin_dev, fpe, (const float*)A, (float*)r.lo);
    CB2_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" void chase_b200_tf32_set_terms(int terms) { g_terms = terms >= 4 ? 4 : 3; }
extern "C" size_t chase_b200_hemm_tf32_scratch_bytes(int64_t K, int64_t k, int elem_bytes)
{
    return hemm_tf32_scratch_bytes(K, k, elem_bytes);
}

#define CB2_TF32_API(X, TT)                                                                                            \
    extern "C" int chase_b200_hemm_tf32_##X(int64_t M, int64_t K, int64_t k, double are, double aim, const void* A,   \
                                            const void* Alo, int64_t lda, const void* B, int64_t ldb, double bre,     \
                                            double bim, void* C, int64_t ldc, double shift, const double* theta,      \
                                            int64_t sflip, int terms, void* scratch, size_t scratch_bytes, void* st)  \
    {                                                                                                                  \
        if (!hemm_tf32_supported<TT>(M, K, k, A, Alo, lda, B, ldb, C, ldc))                                           \
            return -5;                                                                                                 \
        return hemm_tf32_launch<TT>(M, K, k, are, aim, (const TT*)A, (const TT*)Alo, lda, (const TT*)B, ldb, bre,     \
                                    bim, (TT*)C, ldc, shift, theta, sflip, terms, scratch, scratch_bytes, S(st));     \
    }
CB2_TF32_API(s, float)
CB2_TF32_API(c, cxf)
