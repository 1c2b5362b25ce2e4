// chase_b200 — dense Hermitian eigensolver of the Rayleigh-Ritz step, large-n path:
// blocked one-sided (Hestenes) Jacobi.  Replaces cusolverDn?heevd / ?syevd called by the reference at
// /root/reference/linalg/internal/cuda/rayleighRitz.hpp:169-214.
//
// Why one-sided and blocked: the two-sided cyclic Jacobi of jacobi.cuh streams the whole of G and Z through L2 once
// per round (n-1 rounds per sweep, ~32 n^3 bytes per sweep), which at n = 1400 made the eigensolver 14 % of the
// time-to-solution of BASELINE config C2.  Here only ONE n x n matrix evolves and a round touches it once for 8
// rotation steps' worth of work:
//
//   B = G + sigma I                (sigma from Gershgorin bounds so that every lambda + sigma >= (hi - lo) / 2 > 0)
//   repeat sweeps: for every pair of 8-column blocks (round-robin, all pairs of a round in parallel, one CTA each):
//        M  = Bp^H Bp   (16 x 16 Gram matrix of the 16 columns, accumulated in registers)
//        J  = one cyclic sweep of Jacobi rotations on M (15 parallel steps of 8 disjoint pairs, in shared memory)
//        Bp = Bp J      (each thread owns whole rows: in place)
//   until no block needed a rotation (|m_pq| <= sqrt(n) eps sqrt(m_pp m_qq) everywhere).
//   Then the columns of B are mutually orthogonal: B = U diag(lambda + sigma), U = eigenvectors of G.
//   Eigenvalues are taken as Rayleigh quotients u^H G u with the UNSHIFTED G (no sigma cancellation).
//
// Arithmetic: 8 n^3 flop per sweep (22 GFLOP at n = 1400), register-tiled FP64 FMAs; bound by FP64 issue on the
// n/16 CTAs of a round plus one launch per round.
#pragma once
#include "common.cuh"
#include "jacobi.cuh"

namespace cb2
{

constexpr int OSJ_B = 8;       // columns per block
constexpr int OSJ_K = 2 * OSJ_B; // columns per CTA
constexpr int OSJ_THREADS = 256;

// Gs <- Hermitian copy of the LOWER triangle of G (compute type), per-row Gershgorin interval
template <class T>
__global__ void __launch_bounds__(256) osj_prepare_kernel(int n, const T* G, long long ldg,
                                                           typename Traits<T>::comp* Gs, double* glo, double* ghi)
{
    using C = typename Traits<T>::comp;
    __shared__ double sh[32];
    const int i = blockIdx.x; // row
    double r = 0.0, d = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x)
    {
        C v;
        if (i >= j)
            v = widen(G[i + (long long)j * ldg]);
        else
            v = cconj(widen(G[j + (long long)i * ldg]));
        if (i == j)
        {
            v = real_only(v);
            d = creal(v);
        }
        else
            r += sqrt(cabs2(v));
        Gs[i + (long long)j * n] = v;
    }
    r = block_sum(r, sh);
    d = block_sum(d, sh);
    if (threadIdx.x == 0)
    {
        glo[i] = d - r;
        ghi[i] = d + r;
    }
}

// B <- Gs + sigma I with sigma = -lo + (hi - lo)/2; sigma is also written to *sigma_out
template <class C>
__global__ void __launch_bounds__(256) osj_shift_kernel(int n, int ldb, const C* Gs, C* B, const double* glo,
                                                         const double* ghi, double* sigma_out)
{
    __shared__ double slo[256], shi[256];
    double lo = 1e300, hi = -1e300;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
    {
        lo = fmin(lo, glo[i]);
        hi = fmax(hi, ghi[i]);
    }
    slo[threadIdx.x] = lo;
    shi[threadIdx.x] = hi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
    {
        if ((int)threadIdx.x < o)
        {
            slo[threadIdx.x] = fmin(slo[threadIdx.x], slo[threadIdx.x + o]);
            shi[threadIdx.x] = fmax(shi[threadIdx.x], shi[threadIdx.x + o]);
        }
        __syncthreads();
    }
    lo = slo[0];
    hi = shi[0];
    double w = hi - lo;
    if (!(w > 0.0))
        w = fmax(fabs(hi), 1.0);
    const double sigma = -lo + 0.5 * w;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        *sigma_out = sigma;
    const int j = blockIdx.x; // column
    for (int i = threadIdx.x; i < n; i += blockDim.x)
    {
        C v = Gs[i + (long long)j * n];
        if (i == j)
            v = cadd(v, from_real<C>(sigma));
        B[i + (long long)j * ldb] = v;
    }
    if (threadIdx.x == 0)
        for (int i = n; i < ldb; ++i)
            B[i + (long long)j * ldb] = czero<C>();
}

template <class C>
__device__ __forceinline__ C shfl_xor_c(C v, int o);
template <>
__device__ __forceinline__ double shfl_xor_c<double>(double v, int o)
{
    return __shfl_xor_sync(0xffffffffu, v, o);
}
template <>
__device__ __forceinline__ cxd shfl_xor_c<cxd>(cxd v, int o)
{
    return cxd{__shfl_xor_sync(0xffffffffu, v.re, o), __shfl_xor_sync(0xffffffffu, v.im, o)};
}

// one round: CTA a handles the column blocks (p, q) = rr_pair(round, a, nblk).
// The 16 columns are staged in shared memory by 1-D TMA bulk copies (cp.async.bulk, one per column, completion on
// an mbarrier); when all n rows fit (n <= rows_per_chunk: real n <= 1664) the columns are read from HBM/L2 exactly
// once per round, otherwise the Gram pass and the update pass each stream row chunks.
__device__ __forceinline__ void osj_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void osj_mbar_init(uint32_t bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void osj_mbar_expect(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void osj_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "OSJ_WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra OSJ_DONE_%=;\n"
                 "bra OSJ_WAIT_%=;\n"
                 "OSJ_DONE_%=:\n"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}

// ldb: column stride of B in elements (even for real types so that every column is 16-byte aligned);
// rpc: rows per chunk (multiple of 32)
template <class C>
__global__ void __launch_bounds__(OSJ_THREADS) osj_round_kernel(int n, int ldb, int rpc, int nblk, int round, C* B,
                                                                 double tol, int* nrot,
                                                                 unsigned long long* maxoff_bits)
{
    constexpr int K = OSJ_K;
    extern __shared__ __align__(128) unsigned char osj_smem[];
    C* sB = reinterpret_cast<C*>(osj_smem); // [K][rpc]
    __shared__ C sM[K][K + 1];
    __shared__ C sJ[K][K];
    __shared__ JRot sR[K / 2];
    __shared__ int s_active;
    __shared__ int s_col[K];
    __shared__ __align__(8) unsigned long long s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    int bp, bq;
    rr_pair(round, blockIdx.x, nblk, bp, bq);
    if (bp > bq)
    {
        const int t = bp;
        bp = bq;
        bq = t;
    }
    if (tid < K)
    {
        const int c = (tid < OSJ_B) ? bp * OSJ_B + tid : bq * OSJ_B + (tid - OSJ_B);
        s_col[tid] = (c < n) ? c : -1;
    }
    if (tid == 0)
    {
        s_active = 0;
        osj_mbar_init(bar);
    }
    __syncthreads();
    const int nchunks = (n + rpc - 1) / rpc;
    uint32_t parity = 0;

    // rows [r0, r0 + rows) of the 16 columns -> sB   (thread 0 issues, everybody waits)
    auto load_chunk = [&](int r0, int rows)
    {
        if (tid == 0)
        {
            const int rows_al = (rows * (int)sizeof(C) + 15) / 16 * 16 / (int)sizeof(C); // 16-byte granules
            int ncols = 0;
            for (int a = 0; a < K; ++a)
                ncols += (s_col[a] >= 0);
            osj_mbar_expect(bar, (uint32_t)(ncols * rows_al * (int)sizeof(C)));
            for (int a = 0; a < K; ++a)
                if (s_col[a] >= 0)
                    osj_bulk_load((uint32_t)__cvta_generic_to_shared(sB + (size_t)a * rpc),
                                  B + (size_t)s_col[a] * ldb + r0, (uint32_t)(rows_al * (int)sizeof(C)), bar);
        }
        osj_mbar_wait(bar, parity);
        parity ^= 1;
    };

    // ---- phase 1: Gram matrix.  warp w: rows a0..a0+3 of M (a0 = 4 (w/2)), columns c0..c0+7 (c0 = 8 (w%2)) ----
    {
        const int a0 = 4 * (warp >> 1), c0 = 8 * (warp & 1);
        C acc[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 8; ++c)
                acc[a][c] = czero<C>();
        bool ua[4], uc[8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
            ua[a] = s_col[a0 + a] >= 0;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            uc[c] = s_col[c0 + c] >= 0;
        for (int ch = 0; ch < nchunks; ++ch)
        {
            const int r0 = ch * rpc, rows = min(rpc, n - r0);
            if (ch > 0)
                __syncthreads(); // everybody done with the previous chunk
            load_chunk(r0, rows);
            for (int r = lane; r < rows; r += 32)
            {
                C va[4], vc[8];
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    va[a] = ua[a] ? cconj(sB[(size_t)(a0 + a) * rpc + r]) : czero<C>();
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    vc[c] = uc[c] ? sB[(size_t)(c0 + c) * rpc + r] : czero<C>();
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        acc[a][c] = cadd(acc[a][c], cmul(va[a], vc[c]));
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 8; ++c)
            {
                C v = acc[a][c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    v = cadd(v, shfl_xor_c<C>(v, o));
                if (lane == 0)
                    sM[a0 + a][c0 + c] = v;
            }
    }
    // J = I
    {
        const int i = tid / K, j = tid % K;
        sJ[i][j] = (i == j) ? from_real<C>(1.0) : czero<C>();
    }
    __syncthreads();

    // ---- convergence test on the fresh Gram matrix --------------------------------------------------------
    {
        const int i = tid / K, j = tid % K;
        if (i < j)
        {
            const double a = creal(sM[i][i]), b = creal(sM[j][j]);
            const double g = sqrt(cabs2(sM[i][j]));
            if (a > 0.0 && b > 0.0)
            {
                const double ratio = g / sqrt(a * b);
                if (ratio > tol)
                {
                    s_active = 1; // benign race
                    atomicMax(maxoff_bits, (unsigned long long)__double_as_longlong(ratio));
                }
            }
        }
    }
    __syncthreads();
    if (!s_active)
        return;
    if (tid == 0)
        atomicAdd(nrot, 1);

    // ---- phase 2: one cyclic sweep on M in shared memory (15 steps of 8 disjoint pairs), J accumulated ----
    for (int step = 0; step < K - 1; ++step)
    {
        if (tid < K / 2)
        {
            int p, q;
            rr_pair(step, tid, K, p, q);
            if (p > q)
            {
                const int t = p;
                p = q;
                q = t;
            }
            const double a = creal(sM[p][p]), b = creal(sM[q][q]);
            const C g = sM[p][q];
            const double thresh = (a > 0.0 && b > 0.0) ? tol * sqrt(a * b) : 1e300;
            sR[tid] = make_rot<C>(a, b, g, thresh);
        }
        __syncthreads();
        if (tid < (K / 2) * (K / 2))
        {
            // 2x2 block (pair ia rows, pair ib columns) of M <- Ja^H M Jb
            const int ia = tid / (K / 2), ib = tid % (K / 2);
            int pa, qa, pb, qb;
            rr_pair(step, ia, K, pa, qa);
            rr_pair(step, ib, K, pb, qb);
            if (pa > qa)
            {
                const int t = pa;
                pa = qa;
                qa = t;
            }
            if (pb > qb)
            {
                const int t = pb;
                pb = qb;
                qb = t;
            }
            const JRot Ra = sR[ia], Rb = sR[ib];
            if (Ra.active || Rb.active)
            {
                C b00 = sM[pa][pb], b01 = sM[pa][qb], b10 = sM[qa][pb], b11 = sM[qa][qb];
                rot_block<C>(b00, b01, b10, b11, Ra, Rb);
                if (ia == ib)
                {
                    b00 = real_only(b00);
                    b11 = real_only(b11);
                    b01 = czero<C>();
                    b10 = czero<C>();
                }
                sM[pa][pb] = b00;
                sM[pa][qb] = b01;
                sM[qa][pb] = b10;
                sM[qa][qb] = b11;
            }
        }
        else if (tid >= 128 && tid < 128 + K * (K / 2))
        {
            // J <- J Jb : row i, pair ib
            const int t = tid - 128;
            const int i = t / (K / 2), ib = t % (K / 2);
            int pb, qb;
            rr_pair(step, ib, K, pb, qb);
            if (pb > qb)
            {
                const int tt = pb;
                pb = qb;
                qb = tt;
            }
            const JRot Rb = sR[ib];
            if (Rb.active)
            {
                const C ub = cconj(load_u<C>(Rb));
                const C x = sJ[i][pb], y = sJ[i][qb];
                sJ[i][pb] = csub(cmul(Rb.c, x), cmul(Rb.s, cmul(ub, y)));
                sJ[i][qb] = cadd(cmul(Rb.s, x), cmul(Rb.c, cmul(ub, y)));
            }
        }
        __syncthreads();
    }

    // ---- phase 3: Bp <- Bp J.  Each thread owns two rows (one J value feeds two FMAs) -----------------------
    for (int ch = 0; ch < nchunks; ++ch)
    {
        const int r0 = ch * rpc, rows = min(rpc, n - r0);
        if (nchunks > 1)
        {
            __syncthreads();
            load_chunk(r0, rows);
        }
        for (int r = tid; r < rows; r += 2 * OSJ_THREADS)
        {
            const int r1 = r + OSJ_THREADS;
            const bool two = r1 < rows;
            C y0[K], y1[K];
#pragma unroll
            for (int c = 0; c < K; ++c)
            {
                y0[c] = czero<C>();
                y1[c] = czero<C>();
            }
#pragma unroll 2
            for (int a = 0; a < K; ++a)
            {
                if (s_col[a] < 0)
                    continue;
                const C x0 = sB[(size_t)a * rpc + r];
                const C x1 = two ? sB[(size_t)a * rpc + r1] : czero<C>();
#pragma unroll
                for (int c = 0; c < K; ++c)
                {
                    const C j = sJ[a][c];
                    y0[c] = cadd(y0[c], cmul(x0, j));
                    y1[c] = cadd(y1[c], cmul(x1, j));
                }
            }
#pragma unroll
            for (int c = 0; c < K; ++c)
                if (s_col[c] >= 0)
                {
                    B[(size_t)s_col[c] * ldb + r0 + r] = y0[c];
                    if (two)
                        B[(size_t)s_col[c] * ldb + r0 + r1] = y1[c];
                }
        }
    }
}

// ---- the same round on the FP64 tensor pipe --------------------------------------------------------------------
// osj_round_kernel spends most of a round in shared-memory loads: its register-tiled FMAs read 0.4 (Gram) and 0.56
// (update) LDS per FMA.  Here both phases are DMMA.8x8x4 products whose fragments come straight from the column
// tiles (2 loads per 3 DMMAs in the Gram phase, 4 loads per 8 DMMAs in the update):
//   Gram    M[I][J] (8x8) += conj(Bp[r..r+3, 8I..8I+7])^T Bp[r..r+3, 8J..8J+7]   -- the A fragment of block I and the B
//           fragment of block J are the same register (lane l: row r + l%4, column 8I + l/4); warps split the rows,
//           the per-warp partial sums are added in a fixed order (bit-reproducible: the distributed backend runs this
//           solver redundantly on every GPU and relies on identical results).
//   update  Y[r..r+7, 8J..8J+7] = sum_a Bp[r..r+7, 4a..4a+3] J[4a..4a+3, 8J..8J+7]  -- J fragments live in registers
//           for the whole phase; a warp reads all 16 columns of its 8 rows before it overwrites them in place; the
//           finished tile goes back with TMA bulk stores.
// Column stride in shared memory: cs elements with cs * sizeof(C) = 32 (real) / 64 (complex) mod 128, which makes the
// Gram fragment loads bank-conflict free.  ldb (global column stride) is a multiple of 4 with zero rows behind n, so
// every 4-row k-step is fully defined.
template <class C>
struct OsjDmma
{
    static constexpr bool CPLX = sizeof(C) == 16;
    static constexpr int NRED = 8 * 3 * 64; // per-warp partial Gram blocks (elements of C)
};

__device__ __forceinline__ void osj_bulk_store(void* dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}

template <class C>
__global__ void __launch_bounds__(OSJ_THREADS) osj_round_dmma_kernel(int n, int ldb, int rpc, int cs, int nblk, int round,
                                                                      C* B, double tol, int* nrot,
                                                                      unsigned long long* maxoff_bits)
{
    constexpr int K = OSJ_K;
    constexpr bool CPLX = OsjDmma<C>::CPLX;
    extern __shared__ __align__(128) unsigned char osj_smem[];
    C* sB = reinterpret_cast<C*>(osj_smem);          // [K][cs]
    C* sRed = sB + (size_t)K * cs;                     // [8 warps][3 blocks][64]
    __shared__ C sM[K][K + 1];
    __shared__ C sJ[K][K];
    __shared__ JRot sR[K / 2];
    __shared__ int s_active;
    __shared__ int s_col[K];
    __shared__ __align__(8) unsigned long long s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lc = lane >> 2, lq = lane & 3;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    int bp, bq;
    rr_pair(round, blockIdx.x, nblk, bp, bq);
    if (bp > bq)
    {
        const int t = bp;
        bp = bq;
        bq = t;
    }
    if (tid < K)
    {
        const int c = (tid < OSJ_B) ? bp * OSJ_B + tid : bq * OSJ_B + (tid - OSJ_B);
        s_col[tid] = (c < n) ? c : -1;
    }
    if (tid == 0)
    {
        s_active = 0;
        osj_mbar_init(bar);
    }
    __syncthreads();
    const int nchunks = (ldb + rpc - 1) / rpc;
    uint32_t parity = 0;

    // rows [r0, r0 + rows) of the 16 columns -> sB (missing columns are zero-filled); rows is a multiple of 4
    auto load_chunk = [&](int r0, int rows)
    {
        for (int a = 0; a < K; ++a)
            if (s_col[a] < 0)
                for (int r = tid; r < rows; r += OSJ_THREADS)
                    sB[(size_t)a * cs + r] = czero<C>();
        if (tid == 0)
        {
            int ncols = 0;
            for (int a = 0; a < K; ++a)
                ncols += (s_col[a] >= 0);
            osj_mbar_expect(bar, (uint32_t)(ncols * rows * (int)sizeof(C)));
            for (int a = 0; a < K; ++a)
                if (s_col[a] >= 0)
                    osj_bulk_load((uint32_t)__cvta_generic_to_shared(sB + (size_t)a * cs),
                                  B + (size_t)s_col[a] * ldb + r0, (uint32_t)(rows * (int)sizeof(C)), bar);
        }
        osj_mbar_wait(bar, parity);
        parity ^= 1;
        __syncthreads(); // the zero fill of missing columns
    };

    // ---- phase 1: Gram blocks M00, M01, M11 on DMMA; warps split the 4-row k-steps --------------------------------
    {
        double g[3][2][CPLX ? 2 : 1]; // [block][d0,d1][re,im]
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int z = 0; z < (CPLX ? 2 : 1); ++z)
                    g[b][h][z] = 0.0;
        for (int ch = 0; ch < nchunks; ++ch)
        {
            const int r0 = ch * rpc, rows = min(rpc, ldb - r0);
            if (ch > 0)
                __syncthreads();
            load_chunk(r0, rows);
            const C* p0 = sB + (size_t)lc * cs + lq;
            const C* p1 = sB + (size_t)(8 + lc) * cs + lq;
            for (int r = 4 * warp; r < rows; r += 32)
            {
                const C a0 = p0[r], a1 = p1[r];
                if constexpr (!CPLX)
                {
                    dmma884(g[0][0][0], g[0][1][0], a0, a0);
                    dmma884(g[1][0][0], g[1][1][0], a0, a1);
                    dmma884(g[2][0][0], g[2][1][0], a1, a1);
                }
                else
                {
                    // conj(x) y = (xr yr + xi yi) + i (xr yi - xi yr)
                    auto blk = [&](int b, const C& x, const C& y)
                    {
                        dmma884(g[b][0][0], g[b][1][0], x.re, y.re);
                        dmma884(g[b][0][0], g[b][1][0], x.im, y.im);
                        dmma884(g[b][0][1], g[b][1][1], x.re, y.im);
                        dmma884(g[b][0][1], g[b][1][1], -x.im, y.re);
                    };
                    blk(0, a0, a0);
                    blk(1, a0, a1);
                    blk(2, a1, a1);
                }
            }
        }
        // per-warp partials -> sRed[warp][block][lc*8 + 2*lq + h]
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                C v;
                if constexpr (!CPLX)
                    v = g[b][h][0];
                else
                    v = cxd{g[b][h][0], g[b][h][1]};
                sRed[(warp * 3 + b) * 64 + lc * 8 + 2 * lq + h] = v;
            }
        __syncthreads();
        if (tid < 192)
        {
            const int b = tid / 64, e = tid % 64, i = e / 8, j = e % 8;
            C v = sRed[(0 * 3 + b) * 64 + e];
#pragma unroll
            for (int w = 1; w < 8; ++w)
                v = cadd(v, sRed[(w * 3 + b) * 64 + e]);
            if (b == 0)
                sM[i][j] = v;
            else if (b == 2)
                sM[8 + i][8 + j] = v;
            else
            {
                sM[i][8 + j] = v;
                sM[8 + j][i] = cconj(v);
            }
        }
    }
    // J = I
    {
        const int i = tid / K, j = tid % K;
        sJ[i][j] = (i == j) ? from_real<C>(1.0) : czero<C>();
    }
    __syncthreads();

    // ---- convergence test on the fresh Gram matrix --------------------------------------------------------
    {
        const int i = tid / K, j = tid % K;
        if (i < j)
        {
            const double a = creal(sM[i][i]), b = creal(sM[j][j]);
            const double gg = sqrt(cabs2(sM[i][j]));
            if (a > 0.0 && b > 0.0)
            {
                const double ratio = gg / sqrt(a * b);
                if (ratio > tol)
                {
                    s_active = 1; // benign race
                    atomicMax(maxoff_bits, (unsigned long long)__double_as_longlong(ratio));
                }
            }
        }
    }
    __syncthreads();
    if (!s_active)
        return;
    if (tid == 0)
        atomicAdd(nrot, 1);

    // ---- phase 2: one cyclic sweep on M in shared memory (15 steps of 8 disjoint pairs), J accumulated ----
    for (int step = 0; step < K - 1; ++step)
    {
        if (tid < K / 2)
        {
            int p, q;
            rr_pair(step, tid, K, p, q);
            if (p > q)
            {
                const int t = p;
                p = q;
                q = t;
            }
            const double a = creal(sM[p][p]), b = creal(sM[q][q]);
            const C gpq = sM[p][q];
            const double thresh = (a > 0.0 && b > 0.0) ? tol * sqrt(a * b) : 1e300;
            sR[tid] = make_rot<C>(a, b, gpq, thresh);
        }
        __syncthreads();
        if (tid < (K / 2) * (K / 2))
        {
            const int ia = tid / (K / 2), ib = tid % (K / 2);
            int pa, qa, pb, qb;
            rr_pair(step, ia, K, pa, qa);
            rr_pair(step, ib, K, pb, qb);
            if (pa > qa)
            {
                const int t = pa;
                pa = qa;
                qa = t;
            }
            if (pb > qb)
            {
                const int t = pb;
                pb = qb;
                qb = t;
            }
            const JRot Ra = sR[ia], Rb = sR[ib];
            if (Ra.active || Rb.active)
            {
                C b00 = sM[pa][pb], b01 = sM[pa][qb], b10 = sM[qa][pb], b11 = sM[qa][qb];
                rot_block<C>(b00, b01, b10, b11, Ra, Rb);
                if (ia == ib)
                {
                    b00 = real_only(b00);
                    b11 = real_only(b11);
                    b01 = czero<C>();
                    b10 = czero<C>();
                }
                sM[pa][pb] = b00;
                sM[pa][qb] = b01;
                sM[qa][pb] = b10;
                sM[qa][qb] = b11;
            }
        }
        else if (tid >= 128 && tid < 128 + K * (K / 2))
        {
            const int t = tid - 128;
            const int i = t / (K / 2), ib = t % (K / 2);
            int pb, qb;
            rr_pair(step, ib, K, pb, qb);
            if (pb > qb)
            {
                const int tt = pb;
                pb = qb;
                qb = tt;
            }
            const JRot Rb = sR[ib];
            if (Rb.active)
            {
                const C ub = cconj(load_u<C>(Rb));
                const C x = sJ[i][pb], y = sJ[i][qb];
                sJ[i][pb] = csub(cmul(Rb.c, x), cmul(Rb.s, cmul(ub, y)));
                sJ[i][qb] = cadd(cmul(Rb.s, x), cmul(Rb.c, cmul(ub, y)));
            }
        }
        __syncthreads();
    }

    // ---- phase 3: Bp <- Bp J on DMMA; warp w owns the 8-row blocks w, w + 8, ... ---------------------------------
    C jf[4][2]; // J fragments: B operand of (k-step a, column block jb): J[4a + lq][8jb + lc]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int jb = 0; jb < 2; ++jb)
            jf[a][jb] = sJ[4 * a + lq][8 * jb + lc];
    for (int ch = 0; ch < nchunks; ++ch)
    {
        const int r0 = ch * rpc, rows = min(rpc, ldb - r0);
        if (nchunks > 1)
        {
            __syncthreads();
            load_chunk(r0, rows);
        }
        for (int r = 8 * warp; r < rows; r += 64)
        {
            // A fragments: X[r + lc][4a + lq]; rows beyond the chunk (rows is a multiple of 4, not of 8) read as zero
            const bool in = r + lc < rows;
            C xf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
                xf[a] = in ? sB[(size_t)(4 * a + lq) * cs + r + lc] : czero<C>();
            double y[2][2][CPLX ? 2 : 1];
#pragma unroll
            for (int jb = 0; jb < 2; ++jb)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int z = 0; z < (CPLX ? 2 : 1); ++z)
                        y[jb][h][z] = 0.0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int jb = 0; jb < 2; ++jb)
                {
                    if constexpr (!CPLX)
                        dmma884(y[jb][0][0], y[jb][1][0], xf[a], jf[a][jb]);
                    else
                    {
                        dmma884(y[jb][0][0], y[jb][1][0], xf[a].re, jf[a][jb].re);
                        dmma884(y[jb][0][0], y[jb][1][0], -xf[a].im, jf[a][jb].im);
                        dmma884(y[jb][0][1], y[jb][1][1], xf[a].re, jf[a][jb].im);
                        dmma884(y[jb][0][1], y[jb][1][1], xf[a].im, jf[a][jb].re);
                    }
                }
            __syncwarp(); // every lane has read its fragments of these 8 rows
            if (in)
            {
#pragma unroll
                for (int jb = 0; jb < 2; ++jb)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                    {
                        C v;
                        if constexpr (!CPLX)
                            v = y[jb][h][0];
                        else
                            v = cxd{y[jb][h][0], y[jb][h][1]};
                        sB[(size_t)(8 * jb + 2 * lq + h) * cs + r + lc] = v;
                    }
            }
        }
        // shared memory -> global with TMA bulk stores (the generic-proxy writes above must be visible to it)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0)
        {
            for (int a = 0; a < K; ++a)
                if (s_col[a] >= 0)
                    osj_bulk_store(B + (size_t)s_col[a] * ldb + r0, (uint32_t)__cvta_generic_to_shared(sB + (size_t)a * cs),
                                   (uint32_t)(rows * (int)sizeof(C)));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the tile is re-used by the next chunk / released at exit: wait until the stores have READ it
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    // tid 0 has waited until the bulk stores have read the tile (wait_group.read above); the writes themselves are
    // ordered before the next kernel of the stream by the kernel boundary
}

// U[:, j] <- B[:, j] / ||B[:, j]||
template <class C>
__global__ void __launch_bounds__(256) osj_normalize_kernel(int n, int ldb, C* B)
{
    __shared__ double sh[32];
    C* x = B + (long long)blockIdx.x * ldb;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        acc += cabs2(x[i]);
    acc = block_sum(acc, sh);
    const double inv = acc > 0.0 ? 1.0 / sqrt(acc) : 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        x[i] = cmul(inv, x[i]);
}

// w[j] = Re <U[:, j], T[:, j]>   (Rayleigh quotients with T = G U)
template <class C>
__global__ void __launch_bounds__(256) osj_rayleigh_kernel(int n, int ldb, const C* U, const C* Tm, double* w)
{
    __shared__ double sh[32];
    const C* u = U + (long long)blockIdx.x * ldb;
    const C* t = Tm + (long long)blockIdx.x * ldb;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        acc += creal(cmul(cconj(u[i]), t[i]));
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0)
        w[blockIdx.x] = acc;
}

} // namespace cb2
