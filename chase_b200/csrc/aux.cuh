// chase_b200 — HBM-bound helper kernels: copies, column permutation (batched
// Swap), column norms (residuals), the Lanczos A^H*[v1..v4] product and its
// fused vector updates, start-vector RNG, Hermitian check.
//
// Reference counterparts: /root/reference/linalg/internal/cuda/lacpy.cu:62-496,
// residuals.cu:113-296, lanczos_kernels.cu:40-1188 (one 256-thread block per
// Lanczos vector), random_normal_distribution.cu:21-93, shiftDiagonal.cu:23-50.
#pragma once
#include "common.cuh"

namespace cb2
{

template <class T>
__global__ void lacpy_kernel(long long rows, long long cols, const T* src, long long lds, T* dst, long long ldd)
{
    const long long j = blockIdx.y;
    for (long long jj = j; jj < cols; jj += gridDim.y)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
             i += (long long)gridDim.x * blockDim.x)
            dst[i + jj * ldd] = src[i + jj * lds];
}

// Packed triangle <-> square (column-major packing, LAPACK 'U' / 'L' order): what travels in the Gram / projected
// matrix allreduces of the distributed backend (n (n+1) / 2 elements instead of ldg n).  Reference:
// cuda::extractUpperTriangular / unpackUpperTriangular (linalg/internal/cuda/lacpy.cu:837-, 956-), used at
// nccl/cholqr.hpp:152-157 and nccl/rayleighRitz.hpp:124-132.
//   upper: column j holds rows 0..j      at P[j (j+1) / 2 + i]
//   lower: column j holds rows j..n-1    at P[j n - j (j-1) / 2 + (i - j)]
template <class T, bool PACK>
__global__ void __launch_bounds__(256) tri_pack_kernel(long long n, T* G, long long ldg, T* P, int lower)
{
    const long long j = blockIdx.y;
    for (long long jj = j; jj < n; jj += gridDim.y)
    {
        const long long i0 = lower ? jj : 0, i1 = lower ? n : jj + 1;
        const long long base = lower ? jj * n - jj * (jj - 1) / 2 - jj : jj * (jj + 1) / 2;
        for (long long i = i0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < i1;
             i += (long long)gridDim.x * blockDim.x)
        {
            if (PACK)
                P[base + i] = G[i + jj * ldg];
            else
                G[i + jj * ldg] = P[base + i];
        }
    }
}

// dst[:, dcols[t]] = src[:, scols[t]] for t < cnt  (src and dst must not alias)
template <class T>
__global__ void gather_cols_kernel(long long rows, int cnt, const int* scols, const int* dcols, const T* src,
                                   long long lds, T* dst, long long ldd)
{
    for (int t = blockIdx.y; t < cnt; t += gridDim.y)
    {
        const long long s = scols[t], d = dcols[t];
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
             i += (long long)gridDim.x * blockDim.x)
            dst[i + d * ldd] = src[i + s * lds];
    }
}

// dst[i, j] = src[src_row[i], j] for every i with src_row[i] >= 0.  One primitive for all layout changes of the
// distributed backend (pick local rows out of an all-gathered panel, un-permute block-cyclic pieces into global
// order); replaces the per-block broadcasts of the reference's redistributeImpl
// (linalg/distMatrix/distMultiVector.hpp:2817-2909).
// piece_stride > 0: the source is an all-gathered stack of pieces, each a column-major (lds x cols) panel stored
// piece_stride elements apart; src_row = piece * lds + row inside the piece.
template <class T>
__global__ void gather_rows_kernel(long long rows, long long cols, const long long* src_row, const T* src,
                                   long long lds, long long piece_stride, T* dst, long long ldd)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
         i += (long long)gridDim.x * blockDim.x)
    {
        long long s = src_row[i];
        if (s < 0)
            continue;
        if (piece_stride > 0)
        {
            const long long piece = s / lds;
            s = piece * piece_stride + (s - piece * lds);
        }
        for (long long j = blockIdx.y; j < cols; j += gridDim.y)
            dst[i + j * ldd] = src[s + j * lds];
    }
}

// C[:, j] += g * gvec[j] * E[:, j]   (residual block R = A V - V diag(theta) when A V and V are already at hand)
template <class T>
__global__ void axpy_cols_kernel(long long rows, long long cols, const double* gvec, typename Traits<T>::comp g,
                                 const T* E, long long lde, T* Cm, long long ldc)
{
    using C = typename Traits<T>::comp;
    for (long long j = blockIdx.y; j < cols; j += gridDim.y)
    {
        const C f = cmul(gvec[j], g);
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
             i += (long long)gridDim.x * blockDim.x)
            Cm[i + j * ldc] = narrow<T>(cadd((C)widen(Cm[i + j * ldc]), cmul(f, (C)widen(E[i + j * lde]))));
    }
}

// A[lin[t]] += c for the local copies of global diagonal entries (reference chase_shift_mgpu_matrix,
// cuda/shiftDiagonal.cu:100-150; index lists built like Impl/pchase_gpu/pchase_gpu.hpp:340-409)
template <class T>
__global__ void shift_diag_list_kernel(long long cnt, const long long* lin, T* A, double c)
{
    using C = typename Traits<T>::comp;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < cnt; t += (long long)gridDim.x * blockDim.x)
        A[lin[t]] = narrow<T>(cadd((C)widen(A[lin[t]]), from_real<C>(c)));
}

// out[j] = ||X[:, j]||_2 (take_sqrt) or its square
template <class T>
__global__ void __launch_bounds__(256) colnorm_kernel(long long rows, const T* X, long long ldx, double* out,
                                                       int take_sqrt)
{
    __shared__ double sh[32];
    const long long j = blockIdx.x;
    const T* x = X + j * ldx;
    double acc = 0.0;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x)
        acc += cabs2(widen(x[i]));
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0)
        out[j] = take_sqrt ? sqrt(acc) : acc;
}

// X[:, j] *= 1/||X[:, j]||
template <class T>
__global__ void __launch_bounds__(1024) normalize_cols_kernel(long long rows, T* X, long long ldx)
{
    using C = typename Traits<T>::comp;
    __shared__ double sh[32];
    T* x = X + (long long)blockIdx.x * ldx;
    double acc = 0.0;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x)
        acc += cabs2(widen(x[i]));
    acc = block_sum(acc, sh);
    // the reference rounds the norm and its reciprocal to Base<T> (cpu/lanczos.hpp:69-82)
    using R = typename Traits<T>::real;
    const R nrm = (R)sqrt(acc);
    const double inv = (double)(R)(1 / nrm);
    for (long long i = threadIdx.x; i < rows; i += blockDim.x)
        x[i] = narrow<T>(cmul(inv, (C)widen(x[i])));
}

// Y[j, v] = sum_i conj(A[i, j]) * X[i, v].  One warp per group of GEMV_CJ columns of A: every X value loaded is used
// for GEMV_CJ columns, so the L2 traffic for X (which every warp has to stream in full) is 1/GEMV_CJ of A's instead
// of NV times A's; A itself is read exactly once from HBM with GEMV_CJ * 2 independent 256-byte requests in flight
// per warp.
constexpr int GEMV_CJ = 4;
template <class T, int NV>
__global__ void __launch_bounds__(256) gemv_conjT_kernel(long long rows, long long cols, const T* A, long long lda,
                                                          const T* X, long long ldx, T* Y, long long ldy)
{
    using C = typename Traits<T>::comp;
    constexpr int CJ = GEMV_CJ;
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long ngroups = (cols + CJ - 1) / CJ;
    for (long long g = warp; g < ngroups; g += nwarps)
    {
        const long long j0 = g * CJ;
        const T* a[CJ];
#pragma unroll
        for (int c = 0; c < CJ; ++c)
            a[c] = A + (j0 + c < cols ? j0 + c : cols - 1) * lda; // clamped: duplicates are not stored
        C acc[CJ][NV];
#pragma unroll
        for (int c = 0; c < CJ; ++c)
#pragma unroll
            for (int v = 0; v < NV; ++v)
                acc[c][v] = czero<C>();
        long long i = lane;
        for (; i + 32 < rows; i += 64)
        {
            C av[2][CJ], xv[2][NV];
#pragma unroll
            for (int u = 0; u < 2; ++u)
            {
#pragma unroll
                for (int c = 0; c < CJ; ++c)
                    av[u][c] = cconj(widen(a[c][i + 32 * u]));
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    xv[u][v] = widen(X[i + 32 * u + v * ldx]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int c = 0; c < CJ; ++c)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                        acc[c][v] = cadd(acc[c][v], cmul(av[u][c], xv[u][v]));
        }
        for (; i < rows; i += 32)
        {
            C xv[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v)
                xv[v] = widen(X[i + v * ldx]);
#pragma unroll
            for (int c = 0; c < CJ; ++c)
            {
                const C av = cconj(widen(a[c][i]));
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    acc[c][v] = cadd(acc[c][v], cmul(av, xv[v]));
            }
        }
#pragma unroll
        for (int c = 0; c < CJ; ++c)
#pragma unroll
            for (int v = 0; v < NV; ++v)
            {
                C r;
                if constexpr (Traits<T>::cplx)
                    r = cxd{warp_sum(acc[c][v].re), warp_sum(acc[c][v].im)};
                else
                    r = warp_sum(acc[c][v]);
                if (lane == 0 && j0 + c < cols)
                    Y[j0 + c + v * ldy] = narrow<T>(r);
            }
    }
}

// Same product, one CTA per group of GEMV_CJ columns: the 8 warps walk the SAME columns side by side with 16-byte
// loads, so every request wave is 4 KB contiguous per column and two waves are in flight (32 KB per CTA), and the
// partial sums are added across the warps at the end (fixed order).  Background: with one warp per column group the
// kernel ran ~6700 concurrent 256/512-byte streams and reached 0.68-0.77 of the copy peak at N = 20000 (DRAM page
// locality, 300+ dependent iterations per warp); fewer, fatter streams are what the copy benchmark itself does.
// Needs 16-byte aligned columns (lda, ldx multiples of VEC, aligned bases).
template <class T, int NV>
__global__ void __launch_bounds__(256) gemv_conjT_blk_kernel(long long rows, long long cols, const T* A, long long lda,
                                                              const T* X, long long ldx, T* Y, long long ldy)
{
    using C = typename Traits<T>::comp;
    constexpr int CJ = GEMV_CJ;
    constexpr int VEC = 16 / (int)sizeof(T) > 0 ? 16 / (int)sizeof(T) : 1;
    constexpr int U = 2;
    struct alignas(16) Pack
    {
        T v[VEC];
    };
    __shared__ C sred[8][CJ * NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ngroups = (cols + CJ - 1) / CJ;
    const long long wave = 8 * 32 * VEC; // rows covered by one request wave of the CTA
    for (long long g = blockIdx.x; g < ngroups; g += gridDim.x)
    {
        const long long j0 = g * CJ;
        const T* a[CJ];
#pragma unroll
        for (int c = 0; c < CJ; ++c)
            a[c] = A + (j0 + c < cols ? j0 + c : cols - 1) * lda;
        C acc[CJ][NV];
#pragma unroll
        for (int c = 0; c < CJ; ++c)
#pragma unroll
            for (int v = 0; v < NV; ++v)
                acc[c][v] = czero<C>();
        long long b = 0; // CTA-uniform row base
        for (; b + U * wave <= rows; b += U * wave)
        {
            const long long i = b + (long long)(warp * 32 + lane) * VEC;
            Pack av[U][CJ], xv[U][NV];
#pragma unroll
            for (int u = 0; u < U; ++u)
            {
#pragma unroll
                for (int c = 0; c < CJ; ++c)
                    av[u][c] = *reinterpret_cast<const Pack*>(a[c] + i + u * wave);
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    xv[u][v] = *reinterpret_cast<const Pack*>(X + i + u * wave + v * ldx);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int e = 0; e < VEC; ++e)
#pragma unroll
                    for (int c = 0; c < CJ; ++c)
                    {
                        const C ac = cconj(widen(av[u][c].v[e]));
#pragma unroll
                        for (int v = 0; v < NV; ++v)
                            acc[c][v] = cadd(acc[c][v], cmul(ac, widen(xv[u][v].v[e])));
                    }
        }
        // remaining rows (< U * wave), element by element
        for (long long r = b + threadIdx.x; r < rows; r += 256)
        {
            C xs[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v)
                xs[v] = widen(X[r + v * ldx]);
#pragma unroll
            for (int c = 0; c < CJ; ++c)
            {
                const C ac = cconj(widen(a[c][r]));
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    acc[c][v] = cadd(acc[c][v], cmul(ac, xs[v]));
            }
        }
#pragma unroll
        for (int c = 0; c < CJ; ++c)
#pragma unroll
            for (int v = 0; v < NV; ++v)
            {
                C r;
                if constexpr (Traits<T>::cplx)
                    r = cxd{warp_sum(acc[c][v].re), warp_sum(acc[c][v].im)};
                else
                    r = warp_sum(acc[c][v]);
                if (lane == 0)
                    sred[warp][c * NV + v] = r;
            }
        __syncthreads();
        if (threadIdx.x < CJ * NV)
        {
            const int c = threadIdx.x / NV, v = threadIdx.x % NV;
            C r = sred[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < 8; ++w)
                r = cadd(r, sred[w][threadIdx.x]);
            if (j0 + c < cols)
                Y[j0 + c + v * ldy] = narrow<T>(r);
        }
        __syncthreads();
    }
}

// One Lanczos step for vector blockIdx.x (reference: cpu/lanczos.hpp:93-150,
// cuda/lanczos.hpp:178-262):
//   alpha = <v1, v2>; v2 -= alpha v1; d[k] = Re alpha;
//   if k > 0: v2 -= beta_prev v0;  beta = ||v2||;
//   if k < M-1: v2 *= 1/beta; e[k] = beta
// scalars live on the device: d, e are M x numvec (column per vector), rbeta[numvec]
template <class T>
__global__ void __launch_bounds__(1024) lanczos_step_kernel(long long rows, int k, int M, const T* v0, const T* v1,
                                                             T* v2, long long ld, double* d, double* e, double* rbeta)
{
    using C = typename Traits<T>::comp;
    using R = typename Traits<T>::real;
    __shared__ double sh[32];
    const int vi = blockIdx.x;
    const T* x0 = v0 + (long long)vi * ld;
    const T* x1 = v1 + (long long)vi * ld;
    T* x2 = v2 + (long long)vi * ld;
    C acc = czero<C>();
    for (long long i = threadIdx.x; i < rows; i += blockDim.x)
        acc = cadd(acc, cmul(cconj((C)widen(x1[i])), (C)widen(x2[i])));
    C alpha;
    if constexpr (Traits<T>::cplx)
    {
        const double re = block_sum(acc.re, sh);
        const double im = block_sum(acc.im, sh);
        alpha = narrow_round<T>(cxd{re, im});
    }
    else
    {
        alpha = narrow_round<T>(block_sum(acc, sh));
    }
    const double bprev = (k > 0) ? (double)(R)rbeta[vi] : 0.0;
    double nrm2 = 0.0;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x)
    {
        C w = csub((C)widen(x2[i]), cmul(alpha, (C)widen(x1[i])));
        if (k > 0)
            w = csub(w, cmul(bprev, (C)widen(x0[i])));
        const T wt = narrow<T>(w);
        x2[i] = wt;
        nrm2 += cabs2(widen(wt));
    }
    nrm2 = block_sum(nrm2, sh);
    const R beta = (R)sqrt(nrm2);
    if (threadIdx.x == 0)
    {
        d[k + (long long)M * vi] = creal(alpha);
        rbeta[vi] = (double)beta;
        if (k < M - 1)
            e[k + (long long)M * vi] = (double)beta;
    }
    if (k < M - 1)
    {
        const double inv = (double)(R)(1 / beta);
        for (long long i = threadIdx.x; i < rows; i += blockDim.x)
            x2[i] = narrow<T>(cmul(inv, (C)widen(x2[i])));
    }
}

// ---- FP32 storage on the FP64 TMA pipeline ----------------------------------------------------------------------
// dst (wide: double / complex<double>) <- src (float / complex<float>) and back; rows x cols, column-major
template <class TS, class TD>
__global__ void convert_kernel(long long rows, long long cols, const TS* src, long long lds, TD* dst, long long ldd)
{
    for (long long j = blockIdx.y; j < cols; j += gridDim.y)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
             i += (long long)gridDim.x * blockDim.x)
            dst[i + j * ldd] = narrow<TD>(widen(src[i + j * lds]));
}

// ---- pseudo-Hermitian (BSE) helpers ------------------------------------------------------------------------------
// H = [[A, B], [-conj(B), -conj(A)]], S = diag(I, -I).  Reference counterparts: flipSign.cu (S X), conjugate.cu +
// lacpy (K-conjugation, Impl/chase_gpu/chase_gpu.hpp:718-742), pseudo_hermitian_lanczos_diag.cu and the S-inner-
// product Lanczos of linalg/internal/cuda/lanczos.hpp:547-785 (CPU statement: cpu/lanczos.hpp:332-516).

// X[i, j] *= a for i < nrows, j < cols (X already points at the first row to scale): a = -1 on the lower half is
// S X (flipLowerHalfMatrixSign), a = 1e-3 is the start-vector damping scaleLowerBlockRows (chase_gpu.hpp:518-529)
template <class T>
__global__ void scale_rows_kernel(long long nrows, long long cols, T* X, long long ldx, double a)
{
    using C = typename Traits<T>::comp;
    for (long long j = blockIdx.y; j < cols; j += gridDim.y)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nrows;
             i += (long long)gridDim.x * blockDim.x)
            X[i + j * ldx] = narrow<T>(cmul(a, (C)widen(X[i + j * ldx])));
}

// distributed panels: X[i, j] *= a for the local rows whose GLOBAL index grow[i] >= g0 (S X on a row-split panel,
// reference: flipSign on the local lower-half rows, linalg/distMatrix/distMultiVector.hpp:1879-2060 context)
template <class T>
__global__ void scale_rows_map_kernel(long long rows, long long cols, const long long* grow, long long g0, T* X,
                                      long long ldx, double a)
{
    using C = typename Traits<T>::comp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows;
         i += (long long)gridDim.x * blockDim.x)
    {
        if (grow[i] < g0)
            continue;
        for (long long j = blockIdx.y; j < cols; j += gridDim.y)
            X[i + j * ldx] = narrow<T>(cmul(a, (C)widen(X[i + j * ldx])));
    }
}

// K-conjugate partner vectors: dst[:, j] = conj([src[half:2 half, j]; src[0:half, j]])  (src and dst: disjoint columns)
template <class T>
__global__ void kconj_kernel(long long half, long long cols, const T* src, long long lds, T* dst, long long ldd)
{
    using C = typename Traits<T>::comp;
    for (long long j = blockIdx.y; j < cols; j += gridDim.y)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < half;
             i += (long long)gridDim.x * blockDim.x)
        {
            const C up = (C)widen(src[i + j * lds]), lo = (C)widen(src[i + half + j * lds]);
            dst[i + half + j * ldd] = narrow<T>(cconj(up));
            dst[i + j * ldd] = narrow<T>(cconj(lo));
        }
}

// S-inner-product normalisation of one Lanczos vector (blockIdx.x): with v2 = H v1,
//   beta = sqrt(Re <v1, S v2>);  v1 /= beta;  v2 /= beta;  e[ke] = beta (ke >= 0);  bnorm[vi] = beta
template <class T>
__global__ void __launch_bounds__(1024) lanczos_pseudo_norm_kernel(long long rows, long long half, int ke, int M,
                                                                    T* v1, T* v2, long long ld, double* e,
                                                                    double* bnorm)
{
    using C = typename Traits<T>::comp;
    using R = typename Traits<T>::real;
    __shared__ double sh[32];
    const int vi = blockIdx.x;
    T* x1 = v1 + (long long)vi * ld;
    T* x2 = v2 + (long long)vi * ld;
    double acc = 0.0;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x)
    {
        const double t = creal(cmul(cconj((C)widen(x1[i])), (C)widen(x2[i])));
        acc += (i < half) ? t : -t;
    }
    acc = block_sum(acc, sh);
    const R beta = (R)sqrt(acc);
    if (threadIdx.x == 0)
    {
        bnorm[vi] = (double)beta;
        if (ke >= 0)
            e[ke + (long long)M * vi] = (double)beta;
    }
    const double inv = (double)(R)(1 / beta);
    for (long long i = threadIdx.x; i < rows; i += blockDim.x)
    {
        x1[i] = narrow<T>(cmul(inv, (C)widen(x1[i])));
        x2[i] = narrow<T>(cmul(inv, (C)widen(x2[i])));
    }
}

// One pseudo-Hermitian Lanczos step for vector blockIdx.x (v1, v2 already S-normalised):
//   alpha = <v2, S v2> (real);  v2 -= alpha v1;  d[k] = alpha;  if 0 < k < M-1: v2 -= beta_k v0
template <class T>
__global__ void __launch_bounds__(1024) lanczos_pseudo_step_kernel(long long rows, long long half, int k, int M,
                                                                    const T* v0, const T* v1, T* v2, long long ld,
                                                                    double* d, const double* bnorm)
{
    using C = typename Traits<T>::comp;
    using R = typename Traits<T>::real;
    __shared__ double sh[32];
    const int vi = blockIdx.x;
    const T* x0 = v0 + (long long)vi * ld;
    const T* x1 = v1 + (long long)vi * ld;
    T* x2 = v2 + (long long)vi * ld;
    double acc = 0.0;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x)
    {
        const double t = cabs2(widen(x2[i]));
        acc += (i < half) ? t : -t;
    }
    acc = block_sum(acc, sh);
    const double alpha = (double)(R)acc;
    if (threadIdx.x == 0)
        d[k + (long long)M * vi] = alpha;
    const bool with_v0 = (k > 0 && k < M - 1);
    const double beta = with_v0 ? (double)(R)bnorm[vi] : 0.0;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x)
    {
        C w = csub((C)widen(x2[i]), cmul(alpha, (C)widen(x1[i])));
        if (with_v0)
            w = csub(w, cmul(beta, (C)widen(x0[i])));
        x2[i] = narrow<T>(w);
    }
}

// ---- Philox4x32-10 + Box-Muller: production-mode start vectors ----------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2])
{
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
}

__device__ __forceinline__ void philox_normal2(unsigned long long seed, unsigned long long idx, double& n0, double& n1)
{
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r)
        philox_round(c, k);
    const double u0 = ((double)(((unsigned long long)c[0] << 21) ^ (c[1] >> 11)) + 1.0) * (1.0 / 9007199254740993.0);
    const double u1 = ((double)(((unsigned long long)c[2] << 21) ^ (c[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u0));
    double sn, cs;
    sincospi(2.0 * u1, &sn, &cs);
    n0 = rad * cs;
    n1 = rad * sn;
}

template <class T>
__global__ void rng_normal_kernel(long long rows, long long cols, T* X, long long ldx, unsigned long long seed)
{
    const long long total = rows * cols;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
    {
        double a, b;
        philox_normal2(seed, (unsigned long long)idx, a, b);
        const long long i = idx % rows, j = idx / rows;
        if constexpr (Traits<T>::cplx)
            X[i + j * ldx] = narrow<T>(cxd{a, b});
        else
            X[i + j * ldx] = narrow<T>(a);
    }
}

// Same generator addressed by GLOBAL element index grow[i] + j * nglobal: the start block does not depend on how the
// rows are distributed (1 GPU and r x c grids draw the same matrix).
template <class T>
__global__ void rng_normal_rows_kernel(long long rows, long long cols, const long long* grow, long long nglobal, T* X,
                                       long long ldx, unsigned long long seed)
{
    const long long total = rows * cols;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
    {
        const long long i = idx % rows, j = idx / rows;
        double a, b;
        philox_normal2(seed, (unsigned long long)(grow[i] + j * nglobal), a, b);
        if constexpr (Traits<T>::cplx)
            X[i + j * ldx] = narrow<T>(cxd{a, b});
        else
            X[i + j * ldx] = narrow<T>(a);
    }
}

// count entries with |A_ij - conj(A_ji)| > tol * (|A_ij| + |A_ji|) ; also used by isSym
template <class T>
__global__ void herm_check_kernel(long long n, const T* A, long long lda, double tol, unsigned long long* bad)
{
    const long long total = n * n;
    unsigned long long cnt = 0;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
    {
        const long long i = idx % n, j = idx / n;
        if (i < j)
            continue;
        const auto a = widen(A[i + j * lda]);
        const auto b = cconj(widen(A[j + i * lda]));
        const double df = sqrt(cabs2(csub(a, b)));
        if (df > tol * (sqrt(cabs2(a)) + sqrt(cabs2(b))))
            ++cnt;
    }
    if (cnt)
        atomicAdd(bad, cnt);
}

template <class T>
__global__ void shift_diag_kernel(long long n, T* A, long long lda, double c)
{
    using C = typename Traits<T>::comp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        A[i + i * lda] = narrow<T>(cadd((C)widen(A[i + i * lda]), from_real<C>(c)));
}

// A <- upper or lower triangle mirrored to the other one (symOrHermMatrix)
template <class T>
__global__ void herm_mirror_kernel(long long n, T* A, long long lda, int from_upper)
{
    const long long total = n * n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
    {
        const long long i = idx % n, j = idx / n;
        if (i <= j)
            continue; // (i, j) strictly lower
        if (from_upper)
            A[i + j * lda] = narrow<T>(cconj(widen(A[j + i * lda])));
        else
            A[j + i * lda] = narrow<T>(cconj(widen(A[i + j * lda])));
    }
}

} // namespace cb2
