// chase_b200 — small dense factorisation kernels behind CholQR
// (replaces cusolverDn?potrf + cublas?trsm of the reference:
//  /root/reference/linalg/internal/cuda/cholqr.hpp:110-132, 238-278, 387-474;
//  shift: cuda/absTrace.cu:113-, cuda/shiftDiagonal.cu:52-98).
//
// Cholesky: right-looking blocked, upper (G = R^H R), NB = 32:
//   potrf_diag_kernel   one CTA factors the NBxNB diagonal block in smem,
//                       records the LAPACK-style info (first bad pivot, 1-based)
//   potrf_panel_kernel  R[j, j+1:] = R_jj^-H G[j, j+1:]  (one thread per column)
//   trailing update     generic DMMA GEMM (TA=C), upper tiles only
// TRSM (V <- V R^-1): diagonal blocks of R are inverted explicitly
// (trinv_kernel, one thread per column, back substitution) and the panel is
// updated with GEMMs (see capi: chase_b200_trsm_*).
#pragma once
#include "common.cuh"

namespace cb2
{

constexpr int POTRF_NB = 32;

template <class T>
__global__ void __launch_bounds__(256) potrf_diag_kernel(int nb, T* G, long long ldg, int j0, int* info)
{
    using C = typename Traits<T>::comp;
    __shared__ C S[POTRF_NB][POTRF_NB + 1];
    __shared__ int bad;
    const int tid = threadIdx.x;
    if (tid == 0)
        bad = 0;
    for (int idx = tid; idx < nb * nb; idx += blockDim.x)
    {
        const int i = idx % nb, j = idx / nb;
        S[i][j] = (i <= j) ? widen(G[(j0 + i) + (long long)(j0 + j) * ldg]) : czero<C>();
    }
    __syncthreads();
    if (*info != 0)
        return; // an earlier block already failed
    for (int k = 0; k < nb; ++k)
    {
        const double piv = creal(S[k][k]);
        if (!(piv > 0.0) || !isfinite(piv))
        {
            if (tid == 0)
            {
                bad = 1;
                *info = j0 + k + 1;
            }
            break;
        }
        const double rkk = sqrt(piv);
        __syncthreads();
        // scale row k
        for (int j = k + tid; j < nb; j += blockDim.x)
            S[k][j] = (j == k) ? from_real<C>(rkk) : cmul(1.0 / rkk, S[k][j]);
        __syncthreads();
        // trailing update S[i][j] -= conj(S[k][i]) * S[k][j], k < i <= j
        const int rem = nb - k - 1;
        for (int idx = tid; idx < rem * rem; idx += blockDim.x)
        {
            const int i = k + 1 + idx % rem, j = k + 1 + idx / rem;
            if (i <= j)
                S[i][j] = csub(S[i][j], cmul(cconj(S[k][i]), S[k][j]));
        }
        __syncthreads();
    }
    __syncthreads();
    if (bad)
        return;
    for (int idx = tid; idx < nb * nb; idx += blockDim.x)
    {
        const int i = idx % nb, j = idx / nb;
        if (i <= j)
            G[(j0 + i) + (long long)(j0 + j) * ldg] = narrow<T>(S[i][j]);
    }
}

// columns c in [j0+nb, n): x = R_jj^-H g  (forward substitution with L = R_jj^H)
template <class T>
__global__ void __launch_bounds__(128) potrf_panel_kernel(int n, int nb, T* G, long long ldg, int j0, const int* info)
{
    using C = typename Traits<T>::comp;
    __shared__ C S[POTRF_NB][POTRF_NB + 1];
    if (*info != 0)
        return;
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x)
    {
        const int i = idx % nb, j = idx / nb;
        S[i][j] = (i <= j) ? widen(G[(j0 + i) + (long long)(j0 + j) * ldg]) : czero<C>();
    }
    __syncthreads();
    const long long c = (long long)j0 + nb + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= n)
        return;
    C x[POTRF_NB];
#pragma unroll
    for (int i = 0; i < POTRF_NB; ++i)
        x[i] = (i < nb) ? widen(G[(j0 + i) + c * ldg]) : czero<C>();
#pragma unroll
    for (int i = 0; i < POTRF_NB; ++i)
    {
        if (i < nb)
        {
            C acc = x[i];
#pragma unroll
            for (int l = 0; l < i; ++l)
                acc = csub(acc, cmul(cconj(S[l][i]), x[l]));
            x[i] = cmul(1.0 / creal(S[i][i]), acc);
        }
    }
#pragma unroll
    for (int i = 0; i < POTRF_NB; ++i)
        if (i < nb)
            G[(j0 + i) + c * ldg] = narrow<T>(x[i]);
}

// Inverse of the nb x nb upper-triangular diagonal block starting at j0.
// One thread per column of the inverse (back substitution); output block is
// dense nb x nb column-major with zeros below the diagonal.
template <class T>
__global__ void trinv_kernel(int n, int nbmax, const T* Rm, long long ldr, T* Rinv)
{
    using C = typename Traits<T>::comp;
    const int blk = blockIdx.x;
    const int j0 = blk * nbmax;
    const int nb = (n - j0 < nbmax) ? n - j0 : nbmax;
    T* X = Rinv + (long long)blk * nbmax * nbmax;
    const int j = threadIdx.x;
    if (j >= nbmax)
        return;
    if (j >= nb)
    {
        for (int i = 0; i < nbmax; ++i)
            X[i + (long long)j * nbmax] = narrow<T>(czero<C>());
        return;
    }
    // x_j = 1/r_jj ; x_i = -(sum_{l=i+1..j} r_il x_l)/r_ii
    X[j + (long long)j * nbmax] =
        narrow<T>(from_real<C>(1.0 / creal(widen(Rm[(j0 + j) + (long long)(j0 + j) * ldr]))));
    for (int i = j - 1; i >= 0; --i)
    {
        C acc = czero<C>();
        for (int l = i + 1; l <= j; ++l)
            acc = cadd(acc, cmul(widen(Rm[(j0 + i) + (long long)(j0 + l) * ldr]), widen(X[l + (long long)j * nbmax])));
        const double rii = creal(widen(Rm[(j0 + i) + (long long)(j0 + i) * ldr]));
        X[i + (long long)j * nbmax] = narrow<T>(cmul(-1.0 / rii, acc));
    }
    for (int i = j + 1; i < nbmax; ++i)
        X[i + (long long)j * nbmax] = narrow<T>(czero<C>());
}

// s = sum_i |G_ii| * scale ; G_ii += s   (absTrace + shiftDiagonalFromDeviceShift)
template <class T>
__global__ void __launch_bounds__(256) shift_by_abstrace_kernel(int n, T* G, long long ldg, double scale,
                                                                 double* shift_out)
{
    using C = typename Traits<T>::comp;
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        acc += sqrt(cabs2(widen(G[i + (long long)i * ldg])));
    acc = block_sum(acc, sh);
    const double s = acc * scale;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
    {
        C v = widen(G[i + (long long)i * ldg]);
        v = cadd(v, from_real<C>(s));
        G[i + (long long)i * ldg] = narrow<T>(v);
    }
    if (threadIdx.x == 0 && shift_out)
        *shift_out = s;
}

} // namespace cb2
