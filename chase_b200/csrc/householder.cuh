// chase_b200 — Householder QR (the reference's fallback when CholQR breaks down or qr == 'H').
//
// Replaces cusolverDnTgeqrf + cusolverDnTorgqr/ungqr of cuda::houseHoulderQR
// (/root/reference/linalg/internal/cuda/cholqr.hpp:524-556; CPU statement
// cpu/cholqr1.hpp:199-215 = LAPACK ?geqrf + ?orgqr).  Same conventions as LAPACK ?larfg / ?larft / ?larfb:
//   H_j = I - tau_j v_j v_j^H,  v_j(j) = 1,  R_jj = beta_j = -sign(Re alpha_j) ||x_j||  (real),
//   panel of NB reflectors H_1..H_NB = I - V T V^H  (forward, column-wise),
// so the orthonormal factor is the one LAPACK returns (up to rounding).
//
// Blocked right-looking factorisation: the NB columns of a panel are reduced one at a time by two small kernels
// (reflector: one CTA, block-wide norm; application to the rest of the panel: one CTA per column), everything else
// — T factor Gram, trailing update, formation of Q — is GEMMs on the DMMA kernels (driver: hhqr_impl in
// kernels_capi.cu).  The fallback is rare; it is built for robustness, launch-bound in the panel (2 launches per
// column) and tensor-bound elsewhere.
#pragma once
#include "common.cuh"

namespace cb2
{

constexpr int HH_NB = 32;

__device__ __forceinline__ cxd cdiv(cxd a, cxd b)
{
    const double d = b.re * b.re + b.im * b.im;
    return cxd{(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
__device__ __forceinline__ double cdiv(double a, double b) { return a / b; }
__device__ __forceinline__ double cimag(double) { return 0.0; }
__device__ __forceinline__ double cimag(cxd a) { return a.im; }

// Reflector for column j of A (rows x n): x = A[j:, j].  On exit A[j, j] = beta, A[j+1:, j] = v[1:], tau[j] = tau.
template <class T>
__global__ void __launch_bounds__(1024) hh_reflector_kernel(long long rows, long long j, T* A, long long lda,
                                                             typename Traits<T>::comp* tau)
{
    using C = typename Traits<T>::comp;
    __shared__ double sh[32];
    T* x = A + j + j * lda;
    const long long len = rows - j;
    double ss = 0.0;
    for (long long i = 1 + threadIdx.x; i < len; i += blockDim.x)
        ss += cabs2(widen(x[i]));
    ss = block_sum(ss, sh);
    const C alpha = (C)widen(x[0]);
    if (ss == 0.0 && cimag(alpha) == 0.0)
    {
        if (threadIdx.x == 0)
            tau[j] = czero<C>(); // H = I
        return;
    }
    const double nrm = sqrt(cabs2(alpha) + ss);
    const double beta = (creal(alpha) >= 0.0) ? -nrm : nrm;
    const C bmc = csub(from_real<C>(beta), alpha);          // beta - alpha
    const C t = cmul(1.0 / beta, bmc);                      // tau = (beta - alpha) / beta
    const C scale = cdiv(from_real<C>(1.0), csub(alpha, from_real<C>(beta))); // 1 / (alpha - beta)
    for (long long i = 1 + threadIdx.x; i < len; i += blockDim.x)
        x[i] = narrow<T>(cmul(scale, (C)widen(x[i])));
    if (threadIdx.x == 0)
    {
        x[0] = narrow<T>(from_real<C>(beta));
        tau[j] = t;
    }
}

// y <- H_j^H y = y - conj(tau_j) v (v^H y) for the panel columns c = j+1+blockIdx.x (one CTA each)
template <class T>
__global__ void __launch_bounds__(1024) hh_apply_kernel(long long rows, long long j, T* A, long long lda,
                                                         const typename Traits<T>::comp* tau)
{
    using C = typename Traits<T>::comp;
    __shared__ double sh[32];
    const long long c = j + 1 + blockIdx.x;
    const T* v = A + j + j * lda;
    T* y = A + j + c * lda;
    const long long len = rows - j;
    const C tj = tau[j];
    if (!cnonzero(tj))
        return;
    C acc = czero<C>();
    for (long long i = 1 + threadIdx.x; i < len; i += blockDim.x)
        acc = cadd(acc, cmul(cconj((C)widen(v[i])), (C)widen(y[i])));
    C w;
    if constexpr (Traits<T>::cplx)
    {
        const double re = block_sum(acc.re, sh);
        const double im = block_sum(acc.im, sh);
        w = cxd{re, im};
    }
    else
        w = block_sum(acc, sh);
    w = cadd(w, (C)widen(y[0])); // v[0] = 1
    const C f = cmul(cconj(tj), w);
    for (long long i = 1 + threadIdx.x; i < len; i += blockDim.x)
        y[i] = narrow<T>(csub((C)widen(y[i]), cmul(f, (C)widen(v[i]))));
    if (threadIdx.x == 0)
        y[0] = narrow<T>(csub((C)widen(y[0]), f));
}

// Clean copy of a panel's reflectors: Vp (rows - j0) x nb, unit diagonal, zeros above it
template <class T>
__global__ void hh_copy_v_kernel(long long rows, long long j0, int nb, const T* A, long long lda, T* Vp, long long ldvp)
{
    using C = typename Traits<T>::comp;
    const long long len = rows - j0;
    for (int c = blockIdx.y; c < nb; c += gridDim.y)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len;
             i += (long long)gridDim.x * blockDim.x)
        {
            T val;
            if (i < c)
                val = narrow<T>(czero<C>());
            else if (i == c)
                val = narrow<T>(from_real<C>(1.0));
            else
                val = A[(j0 + i) + (j0 + c) * lda];
            Vp[i + c * ldvp] = val;
        }
}

// LAPACK ?larft (forward, column-wise) from the Gram matrix G = Vp^H Vp:
//   T[j, j] = tau_j;  T[0:j, j] = -tau_j T[0:j, 0:j] G[0:j, j];  strictly lower part = 0
template <class T>
__global__ void __launch_bounds__(64) hh_larft_kernel(int nb, const T* G, long long ldgm,
                                                       const typename Traits<T>::comp* tau, T* Tm, long long ldt)
{
    using C = typename Traits<T>::comp;
    __shared__ C Ts[HH_NB][HH_NB + 1];
    const int i = threadIdx.x;
    for (int c = 0; c < nb; ++c)
        if (i < nb)
            Ts[i][c] = czero<C>();
    __syncthreads();
    for (int j = 0; j < nb; ++j)
    {
        const C tj = tau[j];
        C s = czero<C>();
        if (i < j)
        {
            for (int l = i; l < j; ++l)
                s = cadd(s, cmul(Ts[i][l], (C)widen(G[l + (long long)j * ldgm])));
            s = cmul(csub(czero<C>(), tj), s);
        }
        __syncthreads();
        if (i < j)
            Ts[i][j] = s;
        if (i == j)
            Ts[j][j] = tj;
        __syncthreads();
    }
    if (i < nb)
        for (int c = 0; c < nb; ++c)
            Tm[i + (long long)c * ldt] = narrow<T>(Ts[i][c]);
}

// Q <- [I; 0] (rows x n)
template <class T>
__global__ void hh_eye_kernel(long long rows, long long n, T* Q, long long ldq)
{
    using C = typename Traits<T>::comp;
    for (long long c = blockIdx.y; c < n; c += gridDim.y)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
             i += (long long)gridDim.x * blockDim.x)
            Q[i + c * ldq] = narrow<T>(i == c ? from_real<C>(1.0) : czero<C>());
}

} // namespace cb2
