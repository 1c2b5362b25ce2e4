// chase_b200 — extern "C" launchers for include/chase_b200_kernels.h
#include "../../include/chase_b200_kernels.h"
#include "aux.cuh"
#include "common.cuh"
#include "factor.cuh"
#include "gemm_generic.cuh"
#include "hemm_tma.cuh"
#include "householder.cuh"
#include "jacobi.cuh"
#include "osj.cuh"
#include "tf32_api.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <vector>

using namespace cb2;

namespace
{

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Optional per-launch timing of the filter HEMM (CUDA events on the launching stream); bench.py turns it on to
// report the kernel's own TFLOP/s next to the whole-solve numbers.
struct HemmProfile
{
    bool on = false;
    std::vector<cudaEvent_t> ev; // pairs
    std::vector<double> flops;
    size_t used = 0;
    cudaEvent_t get()
    {
        if (used == ev.size())
        {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev.push_back(e);
        }
        return ev[used++];
    }
};
HemmProfile g_hprof;

template <class T>
int hemm_rect_impl(int ta, int64_t M, int64_t K, int64_t k, double are, double aim, const void* A, int64_t lda,
                   const void* B, int64_t ldb, double bre, double bim, void* C, int64_t ldc, double shift,
                   const double* theta, void* stream);

// FP32-storage matrices with a registered FP64 copy (chase_b200_widen_register): the filter product then runs on the
// TMA + DMMA kernel of the wide type, with the panels widened into / narrowed out of a scratch area per call
// (O(N k) traffic against O(N^2 k) flop).  Keyed by the base pointer of the narrow matrix.
struct WideCopy
{
    void* wide = nullptr;
    int64_t ld = 0, rows = 0, cols = 0;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
};
inline std::map<const void*, WideCopy>& wide_registry()
{
    static std::map<const void*, WideCopy> r;
    return r;
}
inline std::mutex& wide_mutex()
{
    static std::mutex m;
    return m;
}
// copy of the entry under the lock (solvers on different host threads register / unregister concurrently)
inline bool wide_lookup(const void* A, WideCopy* out)
{
    std::lock_guard<std::mutex> lk(wide_mutex());
    auto it = wide_registry().find(A);
    if (it == wide_registry().end())
        return false;
    *out = it->second;
    return true;
}
template <class T>
struct WideOf;
template <>
struct WideOf<float>
{
    using type = double;
};
template <>
struct WideOf<cxf>
{
    using type = cxd;
};
inline dim3 grid2d(int64_t rows, int64_t cols);

template <class T>
int gemm_impl(int ta, int tb, int64_t M, int64_t N, int64_t K, double are, double aim, const void* A, int64_t lda,
              const void* B, int64_t ldb, double bre, double bim, void* C, int64_t ldc, int uplo, void* ws,
              size_t ws_bytes, void* stream)
{
    using C_ = typename Traits<T>::comp;
    if (M < 0 || N < 0 || K < 0)
        return -2;
    // Tall products with a K-major right operand (Q Z, (A Q) Z, the TRSM block updates, W^H Q when it has enough
    // tiles) run on the TMA + DMMA pipeline of the filter kernel; narrow / deep Gram products stay on the split-K
    // path below.
    if constexpr (std::is_same<T, double>::value || std::is_same<T, cxd>::value)
    {
        constexpr int BN = HemmCfg<Traits<T>::cplx>::BN;
        const int64_t tiles = ((M + 127) / 128) * ((N + BN - 1) / BN);
        if (tb == 0 && uplo == 0 && M >= 128 && K >= 64 && N >= 8 && tiles >= 96 &&
            hemm_tma_supported<T>(M, K, N, A, lda, B, ldb, C, ldc, 128, 64))
            return hemm_tma_launch<T>(ta != 0, M, K, N, make_comp<C_>(are, aim), (const T*)A, lda, (const T*)B, ldb,
                                      make_comp<C_>(bre, bim), (T*)C, ldc, 0.0, nullptr, S(stream));
    }
    GemmArgs<T> p{};
    p.M = M;
    p.N = N;
    p.K = K;
    p.A = (const T*)A;
    p.lda = lda;
    p.B = (const T*)B;
    p.ldb = ldb;
    p.C = (T*)C;
    p.ldc = ldc;
    p.alpha = make_comp<C_>(are, aim);
    p.beta = make_comp<C_>(bre, bim);
    p.E = nullptr;
    p.lde = 0;
    p.gscale = czero<C_>();
    p.gvec = nullptr;
    p.uplo = uplo;
    return gemm_launch<T>(ta != 0, tb != 0, p, ws, ws_bytes, S(stream));
}

template <class T>
int hemm_rect_impl(int ta, int64_t M, int64_t K, int64_t k, double are, double aim, const void* A, int64_t lda,
                   const void* B, int64_t ldb, double bre, double bim, void* C, int64_t ldc, double shift,
                   const double* theta, void* stream)
{
    using C_ = typename Traits<T>::comp;
    if (M < 0 || K < 0 || k < 0)
        return -2;
    if ((theta || shift != 0.0) && (ta || M != K))
        return -2; // the folded diagonal shift needs B and C to share row coordinates
    if (k == 0 || M == 0)
        return 0;
    const C_ alpha = make_comp<C_>(are, aim), beta = make_comp<C_>(bre, bim);
    struct Timed
    {
        cudaStream_t st;
        cudaEvent_t e1 = nullptr;
        Timed(cudaStream_t s, double fl) : st(s)
        {
            if (g_hprof.on)
            {
                cudaEvent_t e0 = g_hprof.get();
                e1 = g_hprof.get();
                g_hprof.flops.push_back(fl);
                cudaEventRecord(e0, st);
            }
        }
        ~Timed()
        {
            if (e1)
                cudaEventRecord(e1, st);
        }
    } timed(S(stream), 2.0 * (Traits<T>::cplx ? 4.0 : 1.0) * (double)M * (double)K * (double)k);
    if constexpr (std::is_same<T, double>::value || std::is_same<T, cxd>::value)
    {
        if (K > 0 && hemm_tma_supported<T>(M, K, k, A, lda, B, ldb, C, ldc))
            return hemm_tma_launch<T>(ta != 0, M, K, k, alpha, (const T*)A, lda, (const T*)B, ldb, beta, (T*)C, ldc,
                                      shift, theta, S(stream));
    }
    else
    {
        // registered TF32 split of A: tcgen05 kind::tf32 kernel (hemm_tf32.cuh)
        Tf32Reg tr;
        if (K > 0 && tf32_lookup(A, &tr) && tr.ld == lda)
        {
            int64_t sflip = 0;
            bool ok = ta != 0; // C = A^H B is the kernel's native orientation
            if (!ta && M == K && tr.kind == 0)
                ok = true; // Hermitian: A B = A^H B on the same storage
            if (!ta && M == K && tr.kind == 1)
            {
                ok = true; // pseudo-Hermitian: H B = S H^H (S B)
                sflip = M / 2;
            }
            if (ok)
            {
                auto fn = Traits<T>::cplx ? chase_b200_hemm_tf32_c : chase_b200_hemm_tf32_s;
                const int rc = fn(M, K, k, are, aim, A, tr.lo, lda, B, ldb, bre, bim, C, ldc, shift, theta, sflip,
                                  tf32_terms(), tr.scratch, tr.scratch_bytes, stream);
                if (rc != -5 && rc != -3)
                    return rc;
            }
        }
        using TW = typename WideOf<T>::type;
        WideCopy w;
        if (K > 0 && wide_lookup(A, &w))
        {
            const int64_t ldbw = (K + 15) / 16 * 16, ldcw = (M + 15) / 16 * 16;
            TW* Bw = (TW*)w.scratch;
            TW* Cw = Bw + ldbw * k;
            const bool fits = (size_t)(ldbw + ldcw) * (size_t)k * sizeof(TW) <= w.scratch_bytes && w.ld == lda;
            if (fits && hemm_tma_supported<TW>(M, K, k, w.wide, lda, Bw, ldbw, Cw, ldcw))
            {
                cudaStream_t st = S(stream);
                convert_kernel<T, TW><<<grid2d(K, k), 256, 0, kcount(st)>>>(K, k, (const T*)B, ldb, Bw, ldbw);
                if (cnonzero(beta))
                    convert_kernel<T, TW><<<grid2d(M, k), 256, 0, kcount(st)>>>(M, k, (const T*)C, ldc, Cw, ldcw);
                int rc = hemm_tma_launch<TW>(ta != 0, M, K, k, alpha, (const TW*)w.wide, lda, Bw, ldbw, beta, Cw, ldcw,
                                             shift, theta, st);
                if (rc)
                    return rc;
                convert_kernel<TW, T><<<grid2d(M, k), 256, 0, kcount(st)>>>(M, k, Cw, ldcw, (T*)C, ldc);
                CB2_CUDA_OK(cudaGetLastError());
                return 0;
            }
        }
    }
    GemmArgs<T> p{};
    p.M = M;
    p.N = k;
    p.K = K;
    p.A = (const T*)A;
    p.lda = lda;
    p.B = (const T*)B;
    p.ldb = ldb;
    p.C = (T*)C;
    p.ldc = ldc;
    p.alpha = alpha;
    p.beta = beta;
    p.uplo = 0;
    if (theta)
    {
        p.E = (const T*)B;
        p.lde = ldb;
        p.gvec = theta;
        p.gscale = cmul(-1.0, alpha);
    }
    else if (shift != 0.0)
    {
        p.E = (const T*)B;
        p.lde = ldb;
        p.gvec = nullptr;
        p.gscale = cmul(-shift, alpha);
    }
    return gemm_launch<T>(ta != 0, false, p, nullptr, 0, S(stream));
}

template <class T>
int hemm_impl(int64_t n, int64_t k, double are, double aim, const void* A, int64_t lda, const void* B, int64_t ldb,
              double bre, double bim, void* C, int64_t ldc, double shift, const double* theta, void* stream)
{
    return hemm_rect_impl<T>(0, n, n, k, are, aim, A, lda, B, ldb, bre, bim, C, ldc, shift, theta, stream);
}

template <class T>
int potrf_impl(int64_t n, void* Gv, int64_t ldg, int* info_dev, void* stream)
{
    using C_ = typename Traits<T>::comp;
    T* G = (T*)Gv;
    cudaStream_t st = S(stream);
    for (int64_t j0 = 0; j0 < n; j0 += POTRF_NB)
    {
        const int nb = (int)std::min<int64_t>(POTRF_NB, n - j0);
        potrf_diag_kernel<T><<<1, 256, 0, kcount(st)>>>(nb, G, ldg, (int)j0, info_dev);
        const int64_t rem = n - j0 - nb;
        if (rem > 0)
        {
            potrf_panel_kernel<T><<<(unsigned)((rem + 127) / 128), 128, 0, kcount(st)>>>((int)n, nb, G, ldg, (int)j0, info_dev);
            // G[j0+nb:, j0+nb:] -= R[j0:j0+nb, j0+nb:]^H R[j0:j0+nb, j0+nb:]   (upper tiles)
            GemmArgs<T> p{};
            p.M = rem;
            p.N = rem;
            p.K = nb;
            p.A = G + j0 + (j0 + nb) * ldg;
            p.lda = ldg;
            p.B = G + j0 + (j0 + nb) * ldg;
            p.ldb = ldg;
            p.C = G + (j0 + nb) + (j0 + nb) * ldg;
            p.ldc = ldg;
            p.alpha = from_real<C_>(-1.0);
            p.beta = from_real<C_>(1.0);
            p.uplo = 1;
            int rc = gemm_launch<T>(true, false, p, nullptr, 0, st);
            if (rc)
                return rc;
        }
    }
    CB2_CUDA_OK(cudaGetLastError());
    return 0;
}

constexpr int TRSM_NB = 128;

template <class T>
int trsm_impl(int64_t rows, int64_t n, const void* Rv, int64_t ldr, void* Vv, int64_t ldv, void* Xv, int64_t ldx,
              void* ws, size_t ws_bytes, void* stream)
{
    using C_ = typename Traits<T>::comp;
    if (n == 0 || rows == 0)
        return 0;
    const int64_t nblk = (n + TRSM_NB - 1) / TRSM_NB;
    if (ws_bytes < (size_t)nblk * TRSM_NB * TRSM_NB * sizeof(T))
        return -3;
    const T* Rm = (const T*)Rv;
    T* V = (T*)Vv;
    T* X = (T*)Xv;
    T* Rinv = (T*)ws;
    cudaStream_t st = S(stream);
    trinv_kernel<T><<<(unsigned)nblk, TRSM_NB, 0, kcount(st)>>>((int)n, TRSM_NB, Rm, ldr, Rinv);
    CB2_CUDA_OK(cudaGetLastError());
    for (int64_t b = 0; b < nblk; ++b)
    {
        const int64_t j0 = b * TRSM_NB;
        const int64_t nb = std::min<int64_t>(TRSM_NB, n - j0);
        if (b > 0)
        {
            // V_b -= X[:, :j0] R[:j0, j0:j0+nb]
            int rc = gemm_impl<T>(0, 0, rows, nb, j0, -1.0, 0.0, X, ldx, Rm + j0 * ldr, ldr, 1.0, 0.0, V + j0 * ldv, ldv,
                                  0, nullptr, 0, stream);
            if (rc)
                return rc;
        }
        // X_b = V_b inv(R_bb)
        int rc = gemm_impl<T>(0, 0, rows, nb, nb, 1.0, 0.0, V + j0 * ldv, ldv, Rinv + b * TRSM_NB * TRSM_NB, TRSM_NB, 0.0,
                              0.0, X + j0 * ldx, ldx, 0, nullptr, 0, stream);
        if (rc)
            return rc;
    }
    return 0;
}

inline dim3 grid2d(int64_t rows, int64_t cols);

// Householder QR of the rows x n matrix A (rows >= n): on exit Q holds the orthonormal factor (LAPACK ?geqrf +
// ?orgqr/?ungqr conventions), A holds R in its upper triangle and the reflectors below.  Workspace layout (elements of
// T unless noted): Vp rows x NB | W NB x n | W2 NB x n | Tall NB x n | G NB x NB | tau n (compute type) | split-K
// scratch for the deep, narrow V^H C products.
constexpr size_t HH_SPLITK_BYTES = size_t(32) << 20;
inline size_t hhqr_ws_bytes(int64_t rows, int64_t n, int elem_bytes)
{
    const size_t nb = HH_NB;
    return ((size_t)rows * nb + 3 * nb * (size_t)n + nb * nb) * (size_t)elem_bytes + (size_t)n * 16 + 512 +
           HH_SPLITK_BYTES;
}

template <class T>
int hhqr_impl(int64_t rows, int64_t n, void* Av, int64_t lda, void* Qv, int64_t ldq, void* ws, size_t ws_bytes,
              void* stream)
{
    using C_ = typename Traits<T>::comp;
    if (n <= 0 || rows <= 0)
        return 0;
    if (rows < n)
        return -2;
    if (ws_bytes < hhqr_ws_bytes(rows, n, (int)sizeof(T)))
        return -3;
    cudaStream_t st = S(stream);
    T* A = (T*)Av;
    T* Q = (T*)Qv;
    const int64_t NB = HH_NB;
    T* Vp = (T*)ws;
    T* W = Vp + (size_t)rows * NB;
    T* W2 = W + (size_t)NB * n;
    T* Tall = W2 + (size_t)NB * n;
    T* G = Tall + (size_t)NB * n;
    C_* tau = (C_*)(((uintptr_t)(G + NB * NB) + 63) & ~(uintptr_t)63);
    void* sk = (void*)(((uintptr_t)(tau + n) + 255) & ~(uintptr_t)255);
    auto panel_v = [&](int64_t j0, int nb) -> int
    {
        hh_copy_v_kernel<T><<<grid2d(rows - j0, nb), 256, 0, kcount(st)>>>(rows, j0, nb, A, lda, Vp, rows);
        CB2_CUDA_OK(cudaGetLastError());
        return 0;
    };
    // ---- factorisation ---------------------------------------------------------------------------------------
    for (int64_t j0 = 0; j0 < n; j0 += NB)
    {
        const int nb = (int)std::min<int64_t>(NB, n - j0);
        for (int64_t j = j0; j < j0 + nb; ++j)
        {
            hh_reflector_kernel<T><<<1, 1024, 0, kcount(st)>>>(rows, j, A, lda, tau);
            if (j + 1 < j0 + nb)
                hh_apply_kernel<T><<<(unsigned)(j0 + nb - j - 1), 1024, 0, kcount(st)>>>(rows, j, A, lda, tau);
        }
        CB2_CUDA_OK(cudaGetLastError());
        int rc = panel_v(j0, nb);
        if (rc)
            return rc;
        const int64_t len = rows - j0;
        // T factor of the panel
        rc = gemm_impl<T>(1, 0, nb, nb, len, 1.0, 0.0, Vp, rows, Vp, rows, 0.0, 0.0, G, NB, 0, sk, HH_SPLITK_BYTES,
                          stream);
        if (rc)
            return rc;
        T* Tp = Tall + (size_t)j0 * NB;
        hh_larft_kernel<T><<<1, 64, 0, kcount(st)>>>(nb, G, NB, tau + j0, Tp, NB);
        CB2_CUDA_OK(cudaGetLastError());
        const int64_t rest = n - j0 - nb;
        if (rest > 0)
        {
            // C <- (I - V T^H V^H) C,  C = A[j0:, j0+nb:]
            T* Cm = A + j0 + (j0 + nb) * lda;
            rc = gemm_impl<T>(1, 0, nb, rest, len, 1.0, 0.0, Vp, rows, Cm, lda, 0.0, 0.0, W, NB, 0, sk, HH_SPLITK_BYTES,
                              stream);
            if (rc)
                return rc;
            rc = gemm_impl<T>(1, 0, nb, rest, nb, 1.0, 0.0, Tp, NB, W, NB, 0.0, 0.0, W2, NB, 0, nullptr, 0, stream);
            if (rc)
                return rc;
            rc = gemm_impl<T>(0, 0, len, rest, nb, -1.0, 0.0, Vp, rows, W2, NB, 1.0, 0.0, Cm, lda, 0, nullptr, 0, stream);
            if (rc)
                return rc;
        }
    }
    // ---- Q = H_1 ... H_n [I; 0], panels applied last to first ----------------------------------------------------
    hh_eye_kernel<T><<<grid2d(rows, n), 256, 0, kcount(st)>>>(rows, n, Q, ldq);
    CB2_CUDA_OK(cudaGetLastError());
    const int64_t npan = (n + NB - 1) / NB;
    for (int64_t p = npan - 1; p >= 0; --p)
    {
        const int64_t j0 = p * NB;
        const int nb = (int)std::min<int64_t>(NB, n - j0);
        int rc = panel_v(j0, nb);
        if (rc)
            return rc;
        const int64_t len = rows - j0, cols = n - j0;
        T* Tp = Tall + (size_t)j0 * NB;
        T* Qm = Q + j0 + j0 * ldq;
        rc = gemm_impl<T>(1, 0, nb, cols, len, 1.0, 0.0, Vp, rows, Qm, ldq, 0.0, 0.0, W, NB, 0, sk, HH_SPLITK_BYTES,
                          stream);
        if (rc)
            return rc;
        rc = gemm_impl<T>(0, 0, nb, cols, nb, 1.0, 0.0, Tp, NB, W, NB, 0.0, 0.0, W2, NB, 0, nullptr, 0, stream);
        if (rc)
            return rc;
        rc = gemm_impl<T>(0, 0, len, cols, nb, -1.0, 0.0, Vp, rows, W2, NB, 1.0, 0.0, Qm, ldq, 0, nullptr, 0, stream);
        if (rc)
            return rc;
    }
    return 0;
}

template <class T>
int heev_jacobi_impl(int64_t n64, const void* Gv, int64_t ldg, void* Zv, int64_t ldz, double* w_host, void* ws,
                     size_t ws_bytes, int* sweeps_out, void* stream)
{
    using C_ = typename Traits<T>::comp;
    const int n = (int)n64;
    if (n <= 0)
        return 0;
    cudaStream_t st = S(stream);
    const size_t need = chase_b200_heev_ws_bytes(n, Traits<T>::cplx);
    if (ws_bytes < need)
        return -3;
    // workspace layout: Gw | Zw | rots | w | perm | fro2 | nrot
    unsigned char* base = (unsigned char*)ws;
    C_* Gw = (C_*)base;
    C_* Zw = Gw + (size_t)n * n;
    size_t off = 2 * (size_t)n * n * sizeof(C_);
    const int np = (n + 1) & ~1, h = np / 2;
    JRot* rots = (JRot*)(base + off);
    off += (size_t)h * sizeof(JRot);
    off = (off + 15) & ~(size_t)15;
    double* w_dev = (double*)(base + off);
    off += (size_t)n * sizeof(double);
    double* fro2 = (double*)(base + off);
    off += 16;
    int* nrot = (int*)(base + off);
    off += 16;
    int* perm_dev = (int*)(base + off);

    CB2_CUDA_OK(cudaMemsetAsync(fro2, 0, 32, st));
    {
        const long long total = (long long)n * n;
        const int blocks = (int)std::min<long long>((total + 255) / 256, 2368);
        jacobi_init_kernel<T><<<blocks, 256, 0, kcount(st)>>>(n, (const T*)Gv, ldg, Gw, Zw, fro2);
    }
    int sweep = 0, converged = 0;
    const int max_sweeps = 40;
    const dim3 agrid((unsigned)((std::max(n, h) + 127) / 128), (unsigned)h, 2);
    for (; sweep < max_sweeps; ++sweep)
    {
        CB2_CUDA_OK(cudaMemsetAsync(nrot, 0, sizeof(int), st));
        for (int r = 0; r < np - 1; ++r)
        {
            jacobi_rot_kernel<C_><<<(h + 127) / 128, 128, 0, kcount(st)>>>(n, np, r, Gw, rots, fro2, nrot);
            jacobi_apply_kernel<C_><<<agrid, 128, 0, kcount(st)>>>(n, np, r, Gw, Zw, rots);
        }
        int nr = 0;
        CB2_CUDA_OK(cudaMemcpyAsync(&nr, nrot, sizeof(int), cudaMemcpyDeviceToHost, st));
        CB2_CUDA_OK(cudaStreamSynchronize(st));
        if (nr == 0)
        {
            converged = 1;
            break;
        }
    }
    jacobi_diag_kernel<C_><<<(n + 255) / 256, 256, 0, kcount(st)>>>(n, Gw, w_dev);
    std::vector<double> w(n);
    CB2_CUDA_OK(cudaMemcpyAsync(w.data(), w_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CB2_CUDA_OK(cudaStreamSynchronize(st));
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return w[a] < w[b]; });
    for (int i = 0; i < n; ++i)
        w_host[i] = w[perm[i]];
    CB2_CUDA_OK(cudaMemcpyAsync(perm_dev, perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    jacobi_gather_kernel<T><<<dim3((unsigned)((n + 255) / 256), (unsigned)n), 256, 0, kcount(st)>>>(n, Zw, n, perm_dev, (T*)Zv, ldz);
    CB2_CUDA_OK(cudaGetLastError());
    CB2_CUDA_OK(cudaStreamSynchronize(st)); // perm (host vector) must outlive the copy
    if (sweeps_out)
        *sweeps_out = sweep + (converged ? 1 : 0);
    return converged ? 0 : 1;
}

// blocked one-sided Jacobi (osj.cuh).  workspace: B | Gs | Tm | glo | ghi | w | sigma | maxoff | nrot | perm
template <class T>
int heev_osj_impl(int64_t n64, const void* Gv, int64_t ldg, void* Zv, int64_t ldz, double* w_host, void* ws,
                  size_t ws_bytes, int* sweeps_out, void* stream)
{
    using C_ = typename Traits<T>::comp;
    const int n = (int)n64;
    cudaStream_t st = S(stream);
    if (ws_bytes < chase_b200_heev_ws_bytes(n, Traits<T>::cplx))
        return -3;
    if (((uintptr_t)ws) & 15)
        return -3;
    // column stride: a multiple of 4 (16-byte aligned TMA bulk copies; whole 4-row k-steps for the DMMA round kernel),
    // rows n .. ldb-1 are zero
    const int ldb = (n + 3) & ~3;
    unsigned char* base = (unsigned char*)ws;
    const size_t nb = (size_t)ldb * n;
    C_* Bm = (C_*)base;
    C_* Tm = Bm + nb;
    C_* Gs = Tm + nb;
    size_t off = (2 * nb + (size_t)n * n) * sizeof(C_);
    double* glo = (double*)(base + off);
    off += (size_t)n * sizeof(double);
    double* ghi = (double*)(base + off);
    off += (size_t)n * sizeof(double);
    double* w_dev = (double*)(base + off);
    off += (size_t)n * sizeof(double);
    double* sigma = (double*)(base + off);
    off += 16;
    unsigned long long* maxoff = (unsigned long long*)(base + off);
    off += 16;
    int* nrot = (int*)(base + off);
    off += 16;
    int* perm_dev = (int*)(base + off);

    osj_prepare_kernel<T><<<n, 256, 0, kcount(st)>>>(n, (const T*)Gv, ldg, Gs, glo, ghi);
    osj_shift_kernel<C_><<<n, 256, 0, kcount(st)>>>(n, ldb, Gs, Bm, glo, ghi, sigma);
    int nblk = (n + OSJ_B - 1) / OSJ_B;
    if (nblk & 1)
        ++nblk;
    if (nblk < 2)
        nblk = 2;
    const double tol = std::sqrt((double)n) * 2.220446049250313e-16;
    const int max_sweeps = 60;
    // DMMA round kernel (default) or the register-tiled FMA kernel (CHASE_B200_OSJ_DMMA=0)
    static const bool use_dmma = []
    {
        const char* e = getenv("CHASE_B200_OSJ_DMMA");
        return !(e && atoi(e) == 0);
    }();
    int rpc, cs = 0;
    size_t smem;
    if (use_dmma)
    {
        // 16 columns of cs elements + the per-warp partial Gram blocks in what the kernel's static shared memory leaves
        // of the 227 KB a CTA may use; cs = rpc + 4 gives the conflict-free stride (32 / 64 bytes mod 128)
        cudaFuncAttributes fa;
        CB2_CUDA_OK(cudaFuncGetAttributes(&fa, osj_round_dmma_kernel<C_>));
        const size_t dyn_max = (size_t)227 * 1024 - fa.sharedSizeBytes - 256;
        const int cs_max = (int)((dyn_max / sizeof(C_) - OsjDmma<C_>::NRED) / OSJ_K);
        const int rmax = (cs_max - 4) / 32 * 32;
        rpc = std::min((ldb + 31) / 32 * 32, rmax);
        cs = rpc + 4;
        smem = ((size_t)OSJ_K * cs + OsjDmma<C_>::NRED) * sizeof(C_);
        CB2_CUDA_OK(cudaFuncSetAttribute(osj_round_dmma_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    else
    {
        // rows per chunk: all of them if 16 columns fit into 208 KB of shared memory
        const int rmax = (int)(208 * 1024 / (OSJ_K * sizeof(C_))) / 32 * 32;
        rpc = std::min((n + 31) / 32 * 32, rmax);
        smem = (size_t)OSJ_K * rpc * sizeof(C_);
        CB2_CUDA_OK(cudaFuncSetAttribute(osj_round_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int sweep = 0, converged = 0;
    for (; sweep < max_sweeps; ++sweep)
    {
        CB2_CUDA_OK(cudaMemsetAsync(maxoff, 0, 32, st)); // maxoff and nrot
        for (int r = 0; r < nblk - 1; ++r)
        {
            if (use_dmma)
                osj_round_dmma_kernel<C_><<<nblk / 2, OSJ_THREADS, smem, kcount(st)>>>(n, ldb, rpc, cs, nblk, r, Bm, tol,
                                                                                         nrot, maxoff);
            else
                osj_round_kernel<C_><<<nblk / 2, OSJ_THREADS, smem, kcount(st)>>>(n, ldb, rpc, nblk, r, Bm, tol, nrot,
                                                                                    maxoff);
        }
        struct
        {
            unsigned long long bits;
            unsigned long long pad;
            int nr;
        } h;
        CB2_CUDA_OK(cudaMemcpyAsync(&h, maxoff, 20, cudaMemcpyDeviceToHost, st));
        CB2_CUDA_OK(cudaStreamSynchronize(st));
        double mo;
        std::memcpy(&mo, &h.bits, sizeof mo);
        // mo = largest scaled off-diagonal seen BEFORE its rotation during this sweep.  (Stopping as soon as
        // mo < 3e-8 "because the next sweep squares it" costs orthogonality inside clusters of eigenvalues: measured
        // 1e-12 instead of 1e-14 on a 50-fold eigenvalue, so the verification sweep stays.)
        if (h.nr == 0 || mo < 64.0 * tol)
        {
            converged = 1;
            break;
        }
    }
    osj_normalize_kernel<C_><<<n, 256, 0, kcount(st)>>>(n, ldb, Bm);
    {
        GemmArgs<C_> p{};
        p.M = n;
        p.N = n;
        p.K = n;
        p.A = Gs;
        p.lda = n;
        p.B = Bm;
        p.ldb = ldb;
        p.C = Tm;
        p.ldc = ldb;
        p.alpha = from_real<C_>(1.0);
        p.beta = czero<C_>();
        int rc = gemm_launch<C_>(false, false, p, nullptr, 0, st);
        if (rc)
            return rc;
    }
    osj_rayleigh_kernel<C_><<<n, 256, 0, kcount(st)>>>(n, ldb, Bm, Tm, w_dev);
    std::vector<double> w(n);
    CB2_CUDA_OK(cudaMemcpyAsync(w.data(), w_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CB2_CUDA_OK(cudaStreamSynchronize(st));
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return w[a] < w[b]; });
    for (int i = 0; i < n; ++i)
        w_host[i] = w[perm[i]];
    CB2_CUDA_OK(cudaMemcpyAsync(perm_dev, perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    jacobi_gather_kernel<T><<<dim3((unsigned)((n + 255) / 256), (unsigned)n), 256, 0, kcount(st)>>>(
        n, Bm, ldb, perm_dev, (T*)Zv, ldz);
    CB2_CUDA_OK(cudaGetLastError());
    CB2_CUDA_OK(cudaStreamSynchronize(st));
    if (sweeps_out)
        *sweeps_out = sweep + (converged ? 1 : 0);
    return converged ? 0 : 1;
}

// 0 = automatic (one-sided blocked Jacobi for n >= 48), 1 = two-sided cyclic Jacobi, 2 = one-sided always
inline int heev_method()
{
    static int v = -1;
    if (v < 0)
    {
        const char* e = getenv("CHASE_B200_HEEV");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <class T>
int heev_impl(int64_t n, const void* Gv, int64_t ldg, void* Zv, int64_t ldz, double* w_host, void* ws, size_t ws_bytes,
              int* sweeps_out, void* stream)
{
    if (n <= 0)
        return 0;
    const int m = heev_method();
    if (m == 2 || (m == 0 && n >= 48))
        return heev_osj_impl<T>(n, Gv, ldg, Zv, ldz, w_host, ws, ws_bytes, sweeps_out, stream);
    return heev_jacobi_impl<T>(n, Gv, ldg, Zv, ldz, w_host, ws, ws_bytes, sweeps_out, stream);
}

template <class T>
int gemv_impl(int64_t rows, int64_t cols, const void* A, int64_t lda, const void* X, int64_t ldx, int nv, void* Y,
              int64_t ldy, void* stream)
{
    cudaStream_t st = S(stream);
    if (rows <= 0 || cols <= 0 || nv <= 0)
        return 0;
    // one warp per GEMV_CJ columns, 8 warps per block (scalar-load kernel)
    const int blocks = (int)std::min<int64_t>((cols + 8 * GEMV_CJ - 1) / (8 * GEMV_CJ), 148 * 8);
    // CTA-per-column-group kernel: persistent grid, 2 CTAs per SM
    const int blk_blocks = (int)std::min<int64_t>((cols + GEMV_CJ - 1) / GEMV_CJ, 148 * 2);
    // 16-byte loads when every column of A and X starts on a 16-byte boundary (CHASE_B200_GEMV_VEC=0: scalar loads)
    constexpr int64_t VEC = 16 / (int64_t)sizeof(T) > 0 ? 16 / (int64_t)sizeof(T) : 1;
    static const bool vec_on = []
    {
        const char* e = getenv("CHASE_B200_GEMV_VEC");
        return !(e && atoi(e) == 0);
    }();
    // (complex<double> already moves 16 bytes per element and measured slower on the CTA-per-group kernel)
    const bool vec = vec_on && sizeof(T) < 16 && lda % VEC == 0 && ldx % VEC == 0 && (((uintptr_t)A | (uintptr_t)X) & 15) == 0;
    int v = 0;
    while (v < nv)
    {
        const T* x = (const T*)X + (int64_t)v * ldx;
        T* y = (T*)Y + (int64_t)v * ldy;
        if (nv - v >= 4)
        {
            if (vec)
                gemv_conjT_blk_kernel<T, 4><<<blk_blocks, 256, 0, kcount(st)>>>(rows, cols, (const T*)A, lda, x, ldx, y, ldy);
            else
                gemv_conjT_kernel<T, 4><<<blocks, 256, 0, kcount(st)>>>(rows, cols, (const T*)A, lda, x, ldx, y, ldy);
            v += 4;
        }
        else if (nv - v >= 2)
        {
            gemv_conjT_kernel<T, 2><<<blocks, 256, 0, kcount(st)>>>(rows, cols, (const T*)A, lda, x, ldx, y, ldy);
            v += 2;
        }
        else
        {
            gemv_conjT_kernel<T, 1><<<blocks, 256, 0, kcount(st)>>>(rows, cols, (const T*)A, lda, x, ldx, y, ldy);
            v += 1;
        }
    }
    CB2_CUDA_OK(cudaGetLastError());
    return 0;
}

inline dim3 grid2d(int64_t rows, int64_t cols)
{
    unsigned gx = (unsigned)std::min<int64_t>(std::max<int64_t>((rows + 255) / 256, 1), 64);
    unsigned gy = (unsigned)std::min<int64_t>(std::max<int64_t>(cols, 1), 65535);
    return dim3(gx, gy);
}

// register-resident DMMA loop: 8 independent accumulator pairs per warp
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* sink)
{
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        acc[i][0] = acc[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        s += acc[i][0] + acc[i][1];
    if (s == 12345.678)
        sink[0] = s;
}

} // namespace

#define CB2_DEFINE_API(X, TT)                                                                                          \
    extern "C" int chase_b200_gemm_##X(int ta, int tb, int64_t M, int64_t N, int64_t K, double are, double aim,       \
                                       const void* A, int64_t lda, const void* B, int64_t ldb, double bre,            \
                                       double bim, void* C, int64_t ldc, int uplo, void* ws, size_t wsb, void* st)    \
    {                                                                                                                  \
        return gemm_impl<TT>(ta, tb, M, N, K, are, aim, A, lda, B, ldb, bre, bim, C, ldc, uplo, ws, wsb, st);         \
    }                                                                                                                  \
    extern "C" int chase_b200_hemm_##X(int64_t n, int64_t k, double are, double aim, const void* A, int64_t lda,      \
                                       const void* B, int64_t ldb, double bre, double bim, void* C, int64_t ldc,      \
                                       double shift, const double* theta, void* st)                                   \
    {                                                                                                                  \
        return hemm_impl<TT>(n, k, are, aim, A, lda, B, ldb, bre, bim, C, ldc, shift, theta, st);                     \
    }                                                                                                                  \
    extern "C" int chase_b200_hemm_rect_##X(int ta, int64_t M, int64_t K, int64_t k, double are, double aim,          \
                                            const void* A, int64_t lda, const void* B, int64_t ldb, double bre,       \
                                            double bim, void* C, int64_t ldc, void* st)                               \
    {                                                                                                                  \
        return hemm_rect_impl<TT>(ta, M, K, k, are, aim, A, lda, B, ldb, bre, bim, C, ldc, 0.0, nullptr, st);         \
    }                                                                                                                  \
    extern "C" int chase_b200_potrf_##X(int64_t n, void* G, int64_t ldg, int* info, void* st)                         \
    {                                                                                                                  \
        return potrf_impl<TT>(n, G, ldg, info, st);                                                                    \
    }                                                                                                                  \
    extern "C" int chase_b200_trsm_##X(int64_t rows, int64_t n, const void* R, int64_t ldr, void* V, int64_t ldv,     \
                                       void* Xo, int64_t ldx, void* ws, size_t wsb, void* st)                         \
    {                                                                                                                  \
        return trsm_impl<TT>(rows, n, R, ldr, V, ldv, Xo, ldx, ws, wsb, st);                                          \
    }                                                                                                                  \
    extern "C" int chase_b200_shift_abstrace_##X(int64_t n, void* G, int64_t ldg, double scale, double* so, void* st) \
    {                                                                                                                  \
        if (n <= 0)                                                                                                    \
            return 0;                                                                                                  \
        shift_by_abstrace_kernel<TT><<<1, 256, 0, kcount(S(st))>>>((int)n, (TT*)G, ldg, scale, so);                           \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_heev_##X(int64_t n, const void* G, int64_t ldg, void* Z, int64_t ldz, double* w,        \
                                       void* ws, size_t wsb, int* sweeps, void* st)                                   \
    {                                                                                                                  \
        return heev_impl<TT>(n, G, ldg, Z, ldz, w, ws, wsb, sweeps, st);                                              \
    }                                                                                                                  \
    extern "C" int chase_b200_colnorms_##X(int64_t rows, int64_t cols, const void* Xm, int64_t ldx, double* out,      \
                                           int take_sqrt, void* st)                                                   \
    {                                                                                                                  \
        if (cols <= 0)                                                                                                 \
            return 0;                                                                                                  \
        colnorm_kernel<TT><<<(unsigned)cols, 256, 0, kcount(S(st))>>>(rows, (const TT*)Xm, ldx, out, take_sqrt);              \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_lacpy_##X(int64_t rows, int64_t cols, const void* src, int64_t lds, void* dst,          \
                                        int64_t ldd, void* st)                                                        \
    {                                                                                                                  \
        if (rows <= 0 || cols <= 0)                                                                                    \
            return 0;                                                                                                  \
        lacpy_kernel<TT><<<grid2d(rows, cols), 256, 0, kcount(S(st))>>>(rows, cols, (const TT*)src, lds, (TT*)dst, ldd);      \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_tri_pack_##X(int64_t n, const void* G, int64_t ldg, void* P, int lower, void* st)       \
    {                                                                                                                  \
        if (n <= 0)                                                                                                    \
            return 0;                                                                                                  \
        tri_pack_kernel<TT, true><<<grid2d(n, n), 256, 0, kcount(S(st))>>>(n, (TT*)const_cast<void*>(G), ldg, (TT*)P, lower); \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_tri_unpack_##X(int64_t n, const void* P, void* G, int64_t ldg, int lower, void* st)     \
    {                                                                                                                  \
        if (n <= 0)                                                                                                    \
            return 0;                                                                                                  \
        tri_pack_kernel<TT, false><<<grid2d(n, n), 256, 0, kcount(S(st))>>>(n, (TT*)G, ldg, (TT*)const_cast<void*>(P), lower); \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_gather_cols_##X(int64_t rows, int cnt, const int* sc, const int* dc, const void* src,   \
                                              int64_t lds, void* dst, int64_t ldd, void* st)                          \
    {                                                                                                                  \
        if (rows <= 0 || cnt <= 0)                                                                                     \
            return 0;                                                                                                  \
        gather_cols_kernel<TT><<<grid2d(rows, cnt), 256, 0, kcount(S(st))>>>(rows, cnt, sc, dc, (const TT*)src, lds,          \
                                                                      (TT*)dst, ldd);                                  \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_gather_rows_##X(int64_t rows, int64_t cols, const int64_t* src_row, const void* src,    \
                                              int64_t lds, int64_t piece_stride, void* dst, int64_t ldd, void* st)    \
    {                                                                                                                  \
        if (rows <= 0 || cols <= 0)                                                                                    \
            return 0;                                                                                                  \
        gather_rows_kernel<TT><<<grid2d(rows, cols), 256, 0, kcount(S(st))>>>(                                        \
            rows, cols, (const long long*)src_row, (const TT*)src, lds, piece_stride, (TT*)dst, ldd);                  \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_axpy_cols_##X(int64_t rows, int64_t cols, const double* gvec, double gre, double gim,   \
                                            const void* E, int64_t lde, void* Cm, int64_t ldc, void* st)              \
    {                                                                                                                  \
        if (rows <= 0 || cols <= 0)                                                                                    \
            return 0;                                                                                                  \
        axpy_cols_kernel<TT><<<grid2d(rows, cols), 256, 0, kcount(S(st))>>>(                                          \
            rows, cols, gvec, make_comp<typename Traits<TT>::comp>(gre, gim), (const TT*)E, lde, (TT*)Cm, ldc);        \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_shift_diag_list_##X(int64_t cnt, const int64_t* lin, void* A, double c, void* st)       \
    {                                                                                                                  \
        if (cnt <= 0)                                                                                                  \
            return 0;                                                                                                  \
        shift_diag_list_kernel<TT><<<(unsigned)std::min<int64_t>((cnt + 255) / 256, 1184), 256, 0, kcount(S(st))>>>(  \
            cnt, (const long long*)lin, (TT*)A, c);                                                                    \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_rng_normal_rows_##X(int64_t rows, int64_t cols, const int64_t* grow, int64_t nglobal,   \
                                                  void* Xm, int64_t ldx, uint64_t seed, void* st)                     \
    {                                                                                                                  \
        if (rows <= 0 || cols <= 0)                                                                                    \
            return 0;                                                                                                  \
        rng_normal_rows_kernel<TT><<<1184, 256, 0, kcount(S(st))>>>(rows, cols, (const long long*)grow, nglobal,      \
                                                                      (TT*)Xm, ldx, seed);                             \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_gemv_conjt_##X(int64_t rows, int64_t cols, const void* A, int64_t lda, const void* Xm,  \
                                             int64_t ldx, int nv, void* Y, int64_t ldy, void* st)                     \
    {                                                                                                                  \
        return gemv_impl<TT>(rows, cols, A, lda, Xm, ldx, nv, Y, ldy, st);                                            \
    }                                                                                                                  \
    extern "C" int chase_b200_lanczos_step_##X(int64_t rows, int nv, int k, int M, const void* v0, const void* v1,    \
                                               void* v2, int64_t ld, double* d, double* e, double* rb, void* st)      \
    {                                                                                                                  \
        if (nv <= 0)                                                                                                   \
            return 0;                                                                                                  \
        lanczos_step_kernel<TT><<<nv, 1024, 0, kcount(S(st))>>>(rows, k, M, (const TT*)v0, (const TT*)v1, (TT*)v2, ld, d, e,  \
                                                        rb);                                                           \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_normalize_cols_##X(int64_t rows, int64_t cols, void* Xm, int64_t ldx, void* st)         \
    {                                                                                                                  \
        if (cols <= 0)                                                                                                 \
            return 0;                                                                                                  \
        normalize_cols_kernel<TT><<<(unsigned)cols, 1024, 0, kcount(S(st))>>>(rows, (TT*)Xm, ldx);                            \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_rng_normal_##X(int64_t rows, int64_t cols, void* Xm, int64_t ldx, uint64_t seed,        \
                                             void* st)                                                                \
    {                                                                                                                  \
        if (rows <= 0 || cols <= 0)                                                                                    \
            return 0;                                                                                                  \
        rng_normal_kernel<TT><<<1184, 256, 0, kcount(S(st))>>>(rows, cols, (TT*)Xm, ldx, seed);                               \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_herm_check_##X(int64_t n, const void* A, int64_t lda, double tol,                       \
                                             unsigned long long* bad, void* st)                                       \
    {                                                                                                                  \
        if (n <= 0)                                                                                                    \
            return 0;                                                                                                  \
        herm_check_kernel<TT><<<1184, 256, 0, kcount(S(st))>>>(n, (const TT*)A, lda, tol, bad);                               \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_shift_diag_##X(int64_t n, void* A, int64_t lda, double c, void* st)                     \
    {                                                                                                                  \
        if (n <= 0)                                                                                                    \
            return 0;                                                                                                  \
        shift_diag_kernel<TT><<<(unsigned)((n + 255) / 256), 256, 0, kcount(S(st))>>>(n, (TT*)A, lda, c);                     \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_herm_mirror_##X(int64_t n, void* A, int64_t lda, int from_upper, void* st)              \
    {                                                                                                                  \
        if (n <= 0)                                                                                                    \
            return 0;                                                                                                  \
        herm_mirror_kernel<TT><<<1184, 256, 0, kcount(S(st))>>>(n, (TT*)A, lda, from_upper);                                  \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_hhqr_##X(int64_t rows, int64_t n, void* A, int64_t lda, void* Q, int64_t ldq, void* ws, \
                                       size_t wsb, void* st)                                                          \
    {                                                                                                                  \
        return hhqr_impl<TT>(rows, n, A, lda, Q, ldq, ws, wsb, st);                                                   \
    }                                                                                                                  \
    extern "C" int chase_b200_scale_rows_##X(int64_t nrows, int64_t cols, void* Xm, int64_t ldx, double a, void* st)  \
    {                                                                                                                  \
        if (nrows <= 0 || cols <= 0)                                                                                   \
            return 0;                                                                                                  \
        scale_rows_kernel<TT><<<grid2d(nrows, cols), 256, 0, kcount(S(st))>>>(nrows, cols, (TT*)Xm, ldx, a);          \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_scale_rows_map_##X(int64_t rows, int64_t cols, const int64_t* grow, int64_t g0,         \
                                                 void* Xm, int64_t ldx, double a, void* st)                           \
    {                                                                                                                  \
        if (rows <= 0 || cols <= 0)                                                                                    \
            return 0;                                                                                                  \
        scale_rows_map_kernel<TT><<<grid2d(rows, cols), 256, 0, kcount(S(st))>>>(rows, cols, (const long long*)grow,  \
                                                                                 g0, (TT*)Xm, ldx, a);                 \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_kconj_##X(int64_t rows, int64_t cols, const void* src, int64_t lds, void* dst,          \
                                        int64_t ldd, void* st)                                                        \
    {                                                                                                                  \
        if (rows % 2 != 0)                                                                                             \
            return -2;                                                                                                 \
        if (rows <= 0 || cols <= 0)                                                                                    \
            return 0;                                                                                                  \
        kconj_kernel<TT><<<grid2d(rows / 2, cols), 256, 0, kcount(S(st))>>>(rows / 2, cols, (const TT*)src, lds,      \
                                                                            (TT*)dst, ldd);                            \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_lanczos_pseudo_norm_##X(int64_t rows, int nv, int ke, int M, void* v1, void* v2,        \
                                                      int64_t ld, double* e, double* bnorm, void* st)                 \
    {                                                                                                                  \
        if (nv <= 0)                                                                                                   \
            return 0;                                                                                                  \
        lanczos_pseudo_norm_kernel<TT><<<nv, 1024, 0, kcount(S(st))>>>(rows, rows / 2, ke, M, (TT*)v1, (TT*)v2, ld,   \
                                                                       e, bnorm);                                      \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int chase_b200_lanczos_pseudo_step_##X(int64_t rows, int nv, int k, int M, const void* v0,             \
                                                      const void* v1, void* v2, int64_t ld, double* d,                \
                                                      const double* bnorm, void* st)                                  \
    {                                                                                                                  \
        if (nv <= 0)                                                                                                   \
            return 0;                                                                                                  \
        lanczos_pseudo_step_kernel<TT><<<nv, 1024, 0, kcount(S(st))>>>(rows, rows / 2, k, M, (const TT*)v0,           \
                                                                       (const TT*)v1, (TT*)v2, ld, d, bnorm);          \
        CB2_CUDA_OK(cudaGetLastError());                                                                               \
        return 0;                                                                                                      \
    }

CB2_DEFINE_API(s, float)
CB2_DEFINE_API(d, double)
CB2_DEFINE_API(c, cxf)
CB2_DEFINE_API(z, cxd)

extern "C" int chase_b200_tridiag_eig(int n, int batch, const double* d, const double* e, int ldde, double* w,
                                      double* Z, void* st)
{
    if (n <= 0 || batch <= 0)
        return 0;
    if (n > JSMALL_MAX)
        return -2;
    jacobi_small_tridiag_kernel<<<batch, 256, 0, kcount(S(st))>>>(n, d, e, ldde, w, Z, nullptr);
    CB2_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int chase_b200_widen_register(const void* A, void* A_wide, int64_t ld, int64_t rows, int64_t cols,
                                         void* scratch, size_t scratch_bytes)
{
    if (!A || !A_wide)
        return -2;
    WideCopy w;
    w.wide = A_wide;
    w.ld = ld;
    w.rows = rows;
    w.cols = cols;
    w.scratch = scratch;
    w.scratch_bytes = scratch_bytes;
    {
        std::lock_guard<std::mutex> lk(wide_mutex());
        wide_registry()[A] = w;
    }
    return 0;
}
extern "C" int chase_b200_widen_unregister(const void* A)
{
    {
        std::lock_guard<std::mutex> lk(wide_mutex());
        wide_registry().erase(A);
    }
    return 0;
}
// refresh the FP64 copy after the narrow matrix changed (upload); type: 's' or 'c'
extern "C" int chase_b200_widen_sync(char type, const void* A, void* stream)
{
    WideCopy w;
    if (!wide_lookup(A, &w))
        return -2;
    if (w.rows <= 0 || w.cols <= 0)
        return 0;
    cudaStream_t st = S(stream);
    if (type == 's')
        convert_kernel<float, double><<<grid2d(w.rows, w.cols), 256, 0, kcount(st)>>>(w.rows, w.cols, (const float*)A, w.ld,
                                                                                      (double*)w.wide, w.ld);
    else if (type == 'c')
        convert_kernel<cxf, cxd><<<grid2d(w.rows, w.cols), 256, 0, kcount(st)>>>(w.rows, w.cols, (const cxf*)A, w.ld,
                                                                                 (cxd*)w.wide, w.ld);
    else
        return -2;
    CB2_CUDA_OK(cudaGetLastError());
    return 0;
}

// precision change of a rows x cols column-major array: (from, to) in {d->s, s->d, z->c, c->z}
extern "C" int chase_b200_convert(char from, char to, int64_t rows, int64_t cols, const void* src, int64_t lds, void* dst,
                                  int64_t ldd, void* stream)
{
    if (rows <= 0 || cols <= 0)
        return 0;
    cudaStream_t st = S(stream);
    if (from == 'd' && to == 's')
        convert_kernel<double, float><<<grid2d(rows, cols), 256, 0, kcount(st)>>>(rows, cols, (const double*)src, lds, (float*)dst, ldd);
    else if (from == 's' && to == 'd')
        convert_kernel<float, double><<<grid2d(rows, cols), 256, 0, kcount(st)>>>(rows, cols, (const float*)src, lds, (double*)dst, ldd);
    else if (from == 'z' && to == 'c')
        convert_kernel<cxd, cxf><<<grid2d(rows, cols), 256, 0, kcount(st)>>>(rows, cols, (const cxd*)src, lds, (cxf*)dst, ldd);
    else if (from == 'c' && to == 'z')
        convert_kernel<cxf, cxd><<<grid2d(rows, cols), 256, 0, kcount(st)>>>(rows, cols, (const cxf*)src, lds, (cxd*)dst, ldd);
    else
        return -2;
    CB2_CUDA_OK(cudaGetLastError());
    return 0;
}

// frees the per-stream scratch of the filter HEMM; call before destroying a stream that ran it
extern "C" int chase_b200_stream_release(void* stream)
{
    hemm_scratch_release(S(stream));
    return 0;
}

extern "C" size_t chase_b200_hhqr_ws_bytes(int64_t rows, int64_t n, int elem_bytes)
{
    return hhqr_ws_bytes(rows, n, elem_bytes);
}

extern "C" size_t chase_b200_trsm_ws_bytes(int64_t n, int elem_bytes)
{
    const int64_t nblk = (n + TRSM_NB - 1) / TRSM_NB;
    return (size_t)nblk * TRSM_NB * TRSM_NB * (size_t)elem_bytes;
}

extern "C" size_t chase_b200_heev_ws_bytes(int64_t n, int is_complex)
{
    const size_t ce = is_complex ? 16 : 8;
    const size_t np = (size_t)((n + 1) & ~(int64_t)1);
    size_t b = 3 * (size_t)(n + 4) * n * ce; // one-sided path: B | T | Gs (the two-sided path uses the first two)
    b += (np / 2) * sizeof(JRot) + 16;
    b += 3 * (size_t)n * sizeof(double) + 64;
    b += (size_t)n * sizeof(int) + 64;
    return b;
}

extern "C" int chase_b200_hemm_path(int code, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc)
{
    // pointers are assumed 128-byte aligned here (the solver's own buffers are)
    const void* al = reinterpret_cast<const void*>(uintptr_t(1024));
    switch (code)
    {
        case 0: return hemm_tma_supported<float>(n, n, k, al, lda, al, ldb, al, ldc) ? 1 : 0;
        case 1: return hemm_tma_supported<double>(n, n, k, al, lda, al, ldb, al, ldc) ? 1 : 0;
        case 2: return hemm_tma_supported<cxf>(n, n, k, al, lda, al, ldb, al, ldc) ? 1 : 0;
        case 3: return hemm_tma_supported<cxd>(n, n, k, al, lda, al, ldb, al, ldc) ? 1 : 0;
    }
    return 0;
}

extern "C" double chase_b200_dmma_peak(int iters, void* st)
{
    cudaStream_t s = S(st);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess)
        return -1.0;
    const int blocks = sms * 4;
    dmma_peak_kernel<<<blocks, 256, 0, s>>>(iters / 4 + 1, sink); // warm-up
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    dmma_peak_kernel<<<blocks, 256, 0, s>>>(iters, sink);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    const double flops = 2.0 * 8 * 8 * 4 * 8.0 * (double)iters * 8.0 * blocks; // per DMMA 512 flop, 8/iter/warp, 8 warps
    return flops / (ms * 1e-3);
}

extern "C" unsigned long long chase_b200_launch_count(void) { return launch_counter(); }

extern "C" long long chase_b200_hemm_tile_remap(long long v, long long ntiles, long long nctas)
{
    return hemm_tile_remap(v, ntiles, nctas);
}

// Host replay of the tile-part sequence CTA `cta` of the filter HEMM executes for ntiles x nkt k-blocks on `sms` SMs
// (same HemmWalk / hemm_schedule code as the kernel): out[3 i .. 3 i + 2] = raster tile, kt_begin, kt_end.
// Returns the number of parts (may exceed cap: only cap are written), or -1 if cta is outside the grid.
extern "C" long long chase_b200_hemm_walk(long long ntiles, long long nkt, int sms, int cta, long long* out, long long cap)
{
    int grid = 0, dp_waves = 0, remap = 0;
    long long span = 0, sk_tiles = 0;
    hemm_schedule(ntiles, nkt, sms, false, grid, span, sk_tiles, dp_waves, remap);
    if (cta < 0 || cta >= grid)
        return -1;
    HemmWalk walk(cta, grid, span, sk_tiles, (int)nkt, dp_waves, remap);
    HemmSpan sp;
    long long n = 0;
    while (walk.next(sp))
    {
        if (n < cap)
        {
            out[3 * n] = sp.tile;
            out[3 * n + 1] = sp.kt_begin;
            out[3 * n + 2] = sp.kt_end;
        }
        ++n;
    }
    return n;
}

extern "C" int chase_b200_hemm_profile_enable(int on)
{
    g_hprof.on = (on != 0);
    g_hprof.used = 0;
    g_hprof.flops.clear();
    return 0;
}

extern "C" int chase_b200_hemm_profile_read(double* out4)
{
    if (cudaDeviceSynchronize() != cudaSuccess)
        return -1;
    double ms = 0, fl = 0, mx = 0;
    const size_t n = g_hprof.flops.size();
    for (size_t i = 0; i < n; ++i)
    {
        float t = 0;
        if (cudaEventElapsedTime(&t, g_hprof.ev[2 * i], g_hprof.ev[2 * i + 1]) != cudaSuccess)
            return -1;
        ms += t;
        fl += g_hprof.flops[i];
        if (t > mx)
            mx = t;
    }
    out4[0] = (double)n;
    out4[1] = ms;
    out4[2] = fl;
    out4[3] = mx;
    return 0;
}

extern "C" const char* chase_b200_version(void) { return "chase_b200 0.1 (sm_100a)"; }
