// chase_b200 — general FP64-tensor-core GEMM for every "small/odd" GEMM-shaped
// step of the solver (Gram matrices, V*Z, Cholesky trailing updates, TRSM
// blocks, LanczosDos) and the bring-up / fallback path of the filter HEMM.
//
//   C <- alpha * op(A) * op(B) + beta * C  [+ gscale * gvec[j] * E[:, j]]
//
// op(X) = X or X^H.  Column-major everywhere (as the reference's BLAS calls:
// /root/reference/Impl/chase_gpu/chase_gpu.hpp:656-678 for HEMM,
// linalg/internal/cuda/rayleighRitz.hpp:125-214, cuda/cholqr.hpp:110-132).
//
// Inner product: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) on the *transposed*
// product so that every lane owns two consecutive rows of one C column
// (16-byte contiguous stores into column-major C):
//     D^T(n,m) += B^T(n,k) * A^T(k,m)
// Complex types use four real DMMAs per complex tile product with the sign of
// the imaginary part folded into the operand fragments.
#pragma once
#include "common.cuh"

namespace cb2
{

template <class T>
struct GemmArgs
{
    using C_ = typename Traits<T>::comp;
    long long M, N, K;
    const T* A;
    long long lda;
    const T* B;
    long long ldb;
    T* C;
    long long ldc;
    C_ alpha, beta;
    const T* E; // optional extra term (same shape as C)
    long long lde;
    C_ gscale;
    const double* gvec; // optional per-column real factor for the extra term
    int uplo;           // 0: all tiles, 1: only tiles that intersect the upper triangle, 2: lower
    int splitk;         // >1: partial sums to ws[z][N][M], reduced by gemm_splitk_reduce
    C_* ws;
    long long kchunk;
};

template <bool CPLX>
struct GemmTile
{
    // real: 128x128 CTA tile, 8 warps as 2(M) x 4(N), warp tile 64x32
    // cplx: 128x64  CTA tile, 8 warps as 4(M) x 2(N), warp tile 32x32
    static constexpr int BM = 128;
    static constexpr int BN = CPLX ? 64 : 128;
    static constexpr int BK = 16;
    static constexpr int WM = CPLX ? 32 : 64;
    static constexpr int WN = 32;
    static constexpr int WARPS_M = BM / WM;
    static constexpr int PAD_MN = CPLX ? 2 : 4; // leading dim == 4 (mod 16) doubles / 2 (mod 8) cxd
    static constexpr int PAD_K = 4;
    static constexpr int A_ELEMS = (BK * (BM + PAD_MN) > BM * (BK + PAD_K)) ? BK * (BM + PAD_MN) : BM * (BK + PAD_K);
    static constexpr int B_ELEMS = (BK * (BN + PAD_MN) > BN * (BK + PAD_K)) ? BK * (BN + PAD_MN) : BN * (BK + PAD_K);
    static constexpr int ELEM_BYTES = CPLX ? 16 : 8;
    static constexpr int SMEM_BYTES = 2 * (A_ELEMS + B_ELEMS) * ELEM_BYTES;
};

// one warp-level k=4 step on fragments already in registers
template <int MI, int NJ>
__device__ __forceinline__ void mma_step(double (&acc)[NJ][MI][2], const double (&a)[MI], const double (&b)[NJ])
{
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int i = 0; i < MI; ++i)
            dmma884(acc[j][i][0], acc[j][i][1], b[j], a[i]);
}

template <class T, bool TA, bool TB>
__global__ void __launch_bounds__(256, 1) gemm_kernel(const GemmArgs<T> p)
{
    using TR = Traits<T>;
    using C_ = typename TR::comp;
    constexpr bool CPLX = TR::cplx;
    using TL = GemmTile<CPLX>;
    constexpr int BM = TL::BM, BN = TL::BN, BK = TL::BK, WM = TL::WM, WN = TL::WN;
    constexpr int MI = WM / 8, NJ = WN / 8;
    constexpr int LDA_S = TA ? (BK + TL::PAD_K) : (BM + TL::PAD_MN);
    constexpr int LDB_S = TB ? (BN + TL::PAD_MN) : (BK + TL::PAD_K);
    constexpr int A_PER_T = BM * BK / 256;
    constexpr int B_PER_T = BN * BK / 256;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    C_* sA = reinterpret_cast<C_*>(smem_raw);
    C_* sB = sA + 2 * TL::A_ELEMS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % TL::WARPS_M, wn = warp / TL::WARPS_M;
    const long long tiles_m = (p.M + BM - 1) / BM;
    // n fastest: co-resident CTAs share a row block of A through L2
    const long long tiles_n = (p.N + BN - 1) / BN;
    const long long tn = blockIdx.x % tiles_n, tm = blockIdx.x / tiles_n;
    if (tm >= tiles_m)
        return;
    const long long m0 = tm * BM, n0 = tn * BN;
    if (p.uplo == 1 && m0 > n0 + BN - 1)
        return;
    if (p.uplo == 2 && n0 > m0 + BM - 1)
        return;
    long long k_begin = 0, k_end = p.K;
    if (p.splitk > 1)
    {
        k_begin = (long long)blockIdx.z * p.kchunk;
        k_end = k_begin + p.kchunk < p.K ? k_begin + p.kchunk : p.K;
    }

    double accr[NJ][MI][2];
    double acci[CPLX ? NJ : 1][CPLX ? MI : 1][2];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int i = 0; i < MI; ++i)
        {
            accr[j][i][0] = accr[j][i][1] = 0.0;
            if constexpr (CPLX)
                acci[j][i][0] = acci[j][i][1] = 0.0;
        }

    C_ ra[A_PER_T], rb[B_PER_T];

    auto gload = [&](long long k0)
    {
#pragma unroll
        for (int r = 0; r < A_PER_T; ++r)
        {
            const int idx = tid + 256 * r;
            C_ v = czero<C_>();
            if constexpr (!TA)
            {
                const int m = idx % BM, k = idx / BM;
                if (m0 + m < p.M && k0 + k < k_end)
                    v = widen(p.A[(m0 + m) + (k0 + k) * p.lda]);
            }
            else
            {
                const int k = idx % BK, m = idx / BK;
                if (m0 + m < p.M && k0 + k < k_end)
                    v = cconj(widen(p.A[(k0 + k) + (m0 + m) * p.lda]));
            }
            ra[r] = v;
        }
#pragma unroll
        for (int r = 0; r < B_PER_T; ++r)
        {
            const int idx = tid + 256 * r;
            C_ v = czero<C_>();
            if constexpr (!TB)
            {
                const int k = idx % BK, n = idx / BK;
                if (n0 + n < p.N && k0 + k < k_end)
                    v = widen(p.B[(k0 + k) + (n0 + n) * p.ldb]);
            }
            else
            {
                const int n = idx % BN, k = idx / BN;
                if (n0 + n < p.N && k0 + k < k_end)
                    v = cconj(widen(p.B[(n0 + n) + (k0 + k) * p.ldb]));
            }
            rb[r] = v;
        }
    };
    auto sstore = [&](int buf)
    {
        C_* a = sA + buf * TL::A_ELEMS;
        C_* b = sB + buf * TL::B_ELEMS;
#pragma unroll
        for (int r = 0; r < A_PER_T; ++r)
        {
            const int idx = tid + 256 * r;
            if constexpr (!TA)
                a[(idx / BM) * LDA_S + (idx % BM)] = ra[r];
            else
                a[(idx / BK) * LDA_S + (idx % BK)] = ra[r];
        }
#pragma unroll
        for (int r = 0; r < B_PER_T; ++r)
        {
            const int idx = tid + 256 * r;
            if constexpr (!TB)
                b[(idx / BK) * LDB_S + (idx % BK)] = rb[r];
            else
                b[(idx / BN) * LDB_S + (idx % BN)] = rb[r];
        }
    };
    auto compute = [&](int buf)
    {
        const C_* a = sA + buf * TL::A_ELEMS;
        const C_* b = sB + buf * TL::B_ELEMS;
        const int fk = lane & 3, fq = lane >> 2;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4)
        {
            C_ fa[MI], fb[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i)
            {
                const int m = wm * WM + 8 * i + fq, k = kk + fk;
                fa[i] = TA ? a[m * LDA_S + k] : a[k * LDA_S + m];
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j)
            {
                const int n = wn * WN + 8 * j + fq, k = kk + fk;
                fb[j] = TB ? b[k * LDB_S + n] : b[n * LDB_S + k];
            }
            if constexpr (!CPLX)
            {
                mma_step<MI, NJ>(accr, fa, fb);
            }
            else
            {
                double are[MI], aim[MI], bre[NJ], bim[NJ], bimn[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i)
                {
                    are[i] = fa[i].re;
                    aim[i] = fa[i].im;
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                {
                    bre[j] = fb[j].re;
                    bim[j] = fb[j].im;
                    bimn[j] = -fb[j].im;
                }
                mma_step<MI, NJ>(accr, are, bre);
                mma_step<MI, NJ>(accr, aim, bimn);
                mma_step<MI, NJ>(acci, aim, bre);
                mma_step<MI, NJ>(acci, are, bim);
            }
        }
    };

    const long long nkt = (k_end > k_begin) ? (k_end - k_begin + BK - 1) / BK : 0;
    if (nkt > 0)
    {
        gload(k_begin);
        sstore(0);
    }
    __syncthreads();
    for (long long kt = 0; kt < nkt; ++kt)
    {
        const int buf = (int)(kt & 1);
        if (kt + 1 < nkt)
            gload(k_begin + (kt + 1) * BK);
        compute(buf);
        if (kt + 1 < nkt)
            sstore(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue ----------------------------------------------------------
    const int fk = lane & 3, fq = lane >> 2;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
    {
        const long long n = n0 + wn * WN + 8 * j + fq;
        if (n >= p.N)
            continue;
        C_ g = czero<C_>();
        if (p.E)
            g = p.gvec ? cmul(p.gvec[n], p.gscale) : p.gscale;
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                const long long m = m0 + wm * WM + 8 * i + 2 * fk + h;
                if (m >= p.M)
                    continue;
                C_ acc;
                if constexpr (CPLX)
                    acc = cxd{accr[j][i][h], acci[j][i][h]};
                else
                    acc = accr[j][i][h];
                if (p.splitk > 1)
                {
                    p.ws[((long long)blockIdx.z * p.N + n) * p.M + m] = acc;
                    continue;
                }
                C_ out = cmul(p.alpha, acc);
                if (cnonzero(p.beta))
                    out = cadd(out, cmul(p.beta, widen(p.C[m + n * p.ldc])));
                if (p.E)
                    out = cadd(out, cmul(g, widen(p.E[m + n * p.lde])));
                p.C[m + n * p.ldc] = narrow<T>(out);
            }
    }
}

template <class T>
__global__ void gemm_splitk_reduce(const GemmArgs<T> p)
{
    using C_ = typename Traits<T>::comp;
    const long long total = p.M * p.N;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
    {
        const long long m = idx % p.M, n = idx / p.M;
        if (p.uplo == 1 && m > n)
            continue;
        if (p.uplo == 2 && n > m)
            continue;
        C_ acc = czero<C_>();
        for (int z = 0; z < p.splitk; ++z)
            acc = cadd(acc, p.ws[((long long)z * p.N + n) * p.M + m]);
        C_ out = cmul(p.alpha, acc);
        if (cnonzero(p.beta))
            out = cadd(out, cmul(p.beta, widen(p.C[m + n * p.ldc])));
        if (p.E)
        {
            C_ g = p.gvec ? cmul(p.gvec[n], p.gscale) : p.gscale;
            out = cadd(out, cmul(g, widen(p.E[m + n * p.lde])));
        }
        p.C[m + n * p.ldc] = narrow<T>(out);
    }
}

// Host-side launcher.  ws/ws_bytes: optional split-K workspace (may be null).
template <class T>
int gemm_launch(bool ta, bool tb, GemmArgs<T> p, void* ws, size_t ws_bytes, cudaStream_t st)
{
    using TR = Traits<T>;
    using C_ = typename TR::comp;
    using TL = GemmTile<TR::cplx>;
    if (p.M <= 0 || p.N <= 0)
        return 0;
    const long long tiles = ((p.M + TL::BM - 1) / TL::BM) * ((p.N + TL::BN - 1) / TL::BN);
    // split-K only for deep, narrow products (Gram matrices of tall panels)
    int splitk = 1;
    if (ws && p.K >= 4096 && tiles < 96)
    {
        long long want = 296 / (tiles > 0 ? tiles : 1);
        long long maxk = p.K / 1024;
        splitk = (int)(want < maxk ? want : maxk);
        while (splitk > 1 && (size_t)splitk * p.M * p.N * sizeof(C_) > ws_bytes)
            --splitk;
        if (splitk < 1)
            splitk = 1;
    }
    p.splitk = splitk;
    p.ws = reinterpret_cast<C_*>(ws);
    p.kchunk = p.K;
    if (splitk > 1)
    {
        long long kc = (p.K + splitk - 1) / splitk;
        kc = (kc + TL::BK - 1) / TL::BK * TL::BK;
        p.kchunk = kc;
        p.splitk = (int)((p.K + kc - 1) / kc);
    }
    dim3 grid((unsigned)tiles, 1, (unsigned)p.splitk);
    auto launch = [&](auto kern) -> int
    {
        CB2_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TL::SMEM_BYTES));
        kern<<<grid, 256, TL::SMEM_BYTES, kcount(st)>>>(p);
        CB2_CUDA_OK(cudaGetLastError());
        return 0;
    };
    int rc;
    if (!ta && !tb)
        rc = launch(gemm_kernel<T, false, false>);
    else if (ta && !tb)
        rc = launch(gemm_kernel<T, true, false>);
    else if (!ta && tb)
        rc = launch(gemm_kernel<T, false, true>);
    else
        rc = launch(gemm_kernel<T, true, true>);
    if (rc)
        return rc;
    if (p.splitk > 1)
    {
        long long total = p.M * p.N;
        int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
        gemm_splitk_reduce<T><<<blocks, 256, 0, kcount(st)>>>(p);
        CB2_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

} // namespace cb2
