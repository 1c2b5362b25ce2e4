// chase_b200 — the Chebyshev-filter HEMM: TMA-fed, mbarrier-pipelined, persistent
// FP64 tensor-core (DMMA) kernel for sm_100a.
//
//   C <- alpha * (A - shift_j I) * B + beta * C          (A: n x n, B/C: n x k panels)
//
// replaces Shift + cublasTgemm of the reference's ChASEGPU::HEMM
// (/root/reference/Impl/chase_gpu/chase_gpu.hpp:599-603, 656-678) and, with a
// per-column shift theta_j, the GEMM + axpy of cuda::residuals
// (linalg/internal/cuda/residuals.hpp:92-110).
//
// Structure (one CTA per SM, persistent over output tiles, n-fastest raster so
// that co-resident CTAs share row blocks of A in L2):
//   warp 8      : producer.  One lane issues cp.async.bulk.tensor (TMA) loads of
//                 128-byte-wide boxes with SWIZZLE_128B into a STAGES-deep ring
//                 and arms the stage's "full" mbarrier with the byte count.
//   warps 0..7  : consumers.  Wait on "full", read DMMA fragments straight from
//                 the swizzled tiles (index maps below make every LDS
//                 bank-conflict free), issue mma.sync.m8n8k4.f64, release the
//                 stage through the "empty" mbarrier.  The epilogue (alpha, beta,
//                 folded diagonal shift) goes register -> global with 16-byte
//                 accesses while the producer already streams the next tile.
//
// The same kernel serves the distributed backend (pChASEGPU): A may be a
// rectangular local block (M x K) and TA selects C <- alpha * A^H * B + beta * C
// for the column-layout -> row-layout step of the reference's distributed HEMM
// (linalg/internal/nccl/hemm.hpp:325-332).  With TA the A tile is K-major like
// the B tile (one 128-row TMA box per stage) and uses the B-style fragment map.
//
// tcgen05/TMEM cannot be used here: tcgen05.mma has no f64 kind (ptxas rejects
// it), so the FP64 tensor path on Blackwell is warp-level DMMA.
//
// Shared-memory tiles (EPB = elements per 128 B: 16 double / 8 complex<double>;
// BK == EPB):
//   A stage : [BM/EPB boxes][BK rows k][EPB elements m]   row = 128 B
//   B stage : [BN rows n][EPB elements k]                  row = 128 B
// SWIZZLE_128B XORs the 16-byte chunk index with (row & 7).  Conflict-free
// fragment reads need, per mma k-step, k values {0,3,4,7} / {1,2,5,6} (+8h) and
// the column map n(c) = {0,1,4,5,2,3,6,7}[c] for lane groups c = lane/4.
#pragma once
#include "common.cuh"

#include <cuda.h>

#include <cstdlib>
#include <type_traits>

namespace cb2
{

template <bool CPLX>
struct HemmCfg
{
    static constexpr int ELEM = CPLX ? 16 : 8;
    static constexpr int EPB = 128 / ELEM; // elements per 128-byte row
    static constexpr int BM = 128;
    static constexpr int BN = CPLX ? 64 : 128;
    static constexpr int BK = EPB;
    static constexpr int WM = CPLX ? 32 : 64;
    static constexpr int WN = 32;
    static constexpr int WARPS_M = BM / WM;
    static constexpr int KSTEPS = BK / 4;
    static constexpr int A_BYTES = BM * BK * ELEM;
    static constexpr int B_BYTES = BN * BK * ELEM;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = CPLX ? 8 : 6;
    static constexpr int CONSUMER_WARPS = 8;
    static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 2 * STAGES * 8;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :
                 : "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}

template <class T>
struct HemmParams
{
    using C_ = typename Traits<T>::comp;
    long long M, K, k;  // op(A) is M x K, B is K x k, C is M x k
    const T* B;         // only for the folded shift term in the epilogue (square, non-transposed case)
    long long ldb;
    T* C;
    long long ldc;
    C_ alpha, beta;
    double shift;        // scalar shift (used when theta == nullptr)
    const double* theta; // per-column shift
    int tiles_m, tiles_n;
};

template <class T, bool TA>
__global__ void __launch_bounds__(HemmCfg<Traits<T>::cplx>::THREADS, 1)
    hemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const HemmParams<T> p)
{
    using TR = Traits<T>;
    using C_ = typename TR::comp;
    constexpr bool CPLX = TR::cplx;
    using CF = HemmCfg<CPLX>;
    constexpr int BM = CF::BM, BN = CF::BN, BK = CF::BK, WM = CF::WM, WN = CF::WN, EPB = CF::EPB, ELEM = CF::ELEM;
    constexpr int MI = WM / 8, NJ = WN / 8, STAGES = CF::STAGES;
    constexpr int IMUL = CPLX ? 2 : 1; // the tensor maps see complex<double> as two FLOAT64

    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * CF::STAGE_BYTES; // full[STAGES], empty[STAGES]
    unsigned char* gen_base = smem_raw + (base - smem_u32(smem_raw));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
    {
        for (int s = 0; s < STAGES; ++s)
        {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (STAGES + s), CF::CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long ntiles = (long long)p.tiles_m * p.tiles_n;
    const int nkt = (int)((p.K + BK - 1) / BK);

    if (warp == CF::CONSUMER_WARPS)
    {
        // ------------------------------ producer ------------------------------
        if (lane == 0)
        {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
            uint32_t it = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
            {
                const int tn = (int)(t % p.tiles_n), tm = (int)(t / p.tiles_n);
                const int m0 = tm * BM, n0 = tn * BN;
                for (int kt = 0; kt < nkt; ++kt, ++it)
                {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * (STAGES + s), ph ^ 1);
                    const uint32_t full = bars + 8 * s;
                    mbar_expect_tx(full, CF::STAGE_BYTES);
                    const uint32_t sa = base + s * CF::STAGE_BYTES;
                    const uint32_t sb = sa + CF::A_BYTES;
                    if constexpr (TA)
                        tma_load_2d(sa, &mapA, full, kt * BK * IMUL, m0);
                    else
                    {
#pragma unroll
                        for (int b = 0; b < BM / EPB; ++b)
                            tma_load_2d(sa + b * (BK * 128), &mapA, full, (m0 + b * EPB) * IMUL, kt * BK);
                    }
                    tma_load_2d(sb, &mapB, full, kt * BK * IMUL, n0);
                }
            }
        }
        return;
    }

    // -------------------------------- consumers --------------------------------
    const int wm = warp % CF::WARPS_M, wn = warp / CF::WARPS_M;
    const int q = lane & 3, c = lane >> 2;
    const int ncol = (c & 1) | ((c & 2) << 1) | ((c & 4) >> 1); // {0,1,4,5,2,3,6,7}
    // per-lane k offsets inside an 8-group for the two k-sets
    const int kq0 = 2 * q + (q & 1), kq1 = 2 * q + ((q & 1) ^ 1);

    uint32_t it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
    {
        const int tn = (int)(t % p.tiles_n), tm = (int)(t / p.tiles_n);
        const long long m0 = (long long)tm * BM, n0 = (long long)tn * BN;

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

        for (int kt = 0; kt < nkt; ++kt, ++it)
        {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(bars + 8 * s, ph);
            const unsigned char* sa = gen_base + s * CF::STAGE_BYTES;
            const unsigned char* sb = sa + CF::A_BYTES;
#pragma unroll
            for (int ks = 0; ks < CF::KSTEPS; ++ks)
            {
                const int kin = 8 * (ks >> 1) + ((ks & 1) ? kq1 : kq0); // k inside the stage
                const int k7 = kin & 7;
                C_ fa[MI], fb[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i)
                {
                    if constexpr (TA)
                    {
                        const int m = wm * WM + 8 * i + ncol;
                        const int byte_in = kin * ELEM;
                        const int off = m * 128 + ((((byte_in >> 4) ^ (m & 7)) << 4) | (byte_in & 15));
                        fa[i] = *reinterpret_cast<const C_*>(sa + off);
                    }
                    else
                    {
                        const int m = wm * WM + 8 * i + c;
                        const int mblk = m / EPB, min_ = m % EPB;
                        const int byte_in = min_ * ELEM;
                        const int off = (mblk * BK + kin) * 128 + ((((byte_in >> 4) ^ k7) << 4) | (byte_in & 15));
                        fa[i] = *reinterpret_cast<const C_*>(sa + off);
                    }
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                {
                    const int n = wn * WN + 8 * j + ncol;
                    const int byte_in = kin * ELEM;
                    const int off = n * 128 + ((((byte_in >> 4) ^ (n & 7)) << 4) | (byte_in & 15));
                    fb[j] = *reinterpret_cast<const C_*>(sb + off);
                }
                if constexpr (!CPLX)
                {
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
#pragma unroll
                        for (int i = 0; i < MI; ++i)
                            dmma884(accr[j][i][0], accr[j][i][1], fb[j], fa[i]);
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                    {
                        // TA: the A operand enters conjugated: (ar - i ai)(br + i bi)
                        const double bre = fb[j].re, bim = fb[j].im;
                        const double s_ri = TA ? fb[j].im : -fb[j].im; // multiplies a.im into the real part
                        const double s_ir = TA ? -fb[j].re : fb[j].re; // multiplies a.im into the imaginary part
#pragma unroll
                        for (int i = 0; i < MI; ++i)
                        {
                            dmma884(accr[j][i][0], accr[j][i][1], bre, fa[i].re);
                            dmma884(accr[j][i][0], accr[j][i][1], s_ri, fa[i].im);
                            dmma884(acci[j][i][0], acci[j][i][1], s_ir, fa[i].im);
                            dmma884(acci[j][i][0], acci[j][i][1], bim, fa[i].re);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(bars + 8 * (STAGES + s));
        }

        // ------------------------------ epilogue ------------------------------
        const bool has_beta = cnonzero(p.beta);
        const bool has_shift = !TA && ((p.theta != nullptr) || (p.shift != 0.0));
#pragma unroll
        for (int j = 0; j < NJ; ++j)
        {
            const long long n = n0 + wn * WN + 8 * j + ncol;
            if (n >= p.k)
                continue;
            // g = -alpha * shift_j
            C_ g = czero<C_>();
            if (has_shift)
                g = cmul(-(p.theta ? p.theta[n] : p.shift), p.alpha);
            T* cc = p.C + n * p.ldc;
            const T* bb = p.B + n * p.ldb;
#pragma unroll
            for (int i = 0; i < MI; ++i)
            {
                // TA: operand column c' of the A fragment holds row ncol(c') (see the fragment map above)
                const int mq = TA ? ((q & 1) << 2 | (q & 2)) : 2 * q; // ncol(2q) = {0,4,2,6}
                const long long m = m0 + wm * WM + 8 * i + mq;
                if (m >= p.M)
                    continue;
                C_ o0, o1;
                if constexpr (CPLX)
                {
                    o0 = cmul(p.alpha, cxd{accr[j][i][0], acci[j][i][0]});
                    o1 = cmul(p.alpha, cxd{accr[j][i][1], acci[j][i][1]});
                }
                else
                {
                    o0 = p.alpha * accr[j][i][0];
                    o1 = p.alpha * accr[j][i][1];
                }
                const bool pair = (m + 1 < p.M);
                if constexpr (!CPLX)
                {
                    if (pair)
                    {
                        if (has_beta)
                        {
                            const double2 cv = *reinterpret_cast<const double2*>(cc + m);
                            o0 += p.beta * cv.x;
                            o1 += p.beta * cv.y;
                        }
                        if (has_shift)
                        {
                            const double2 bv = *reinterpret_cast<const double2*>(bb + m);
                            o0 += g * bv.x;
                            o1 += g * bv.y;
                        }
                        *reinterpret_cast<double2*>(cc + m) = make_double2(o0, o1);
                    }
                    else
                    {
                        if (has_beta)
                            o0 += p.beta * cc[m];
                        if (has_shift)
                            o0 += g * bb[m];
                        cc[m] = o0;
                    }
                }
                else
                {
                    if (has_beta)
                        o0 = cadd(o0, cmul(p.beta, cc[m]));
                    if (has_shift)
                        o0 = cadd(o0, cmul(g, bb[m]));
                    cc[m] = o0;
                    if (pair)
                    {
                        if (has_beta)
                            o1 = cadd(o1, cmul(p.beta, cc[m + 1]));
                        if (has_shift)
                            o1 = cadd(o1, cmul(g, bb[m + 1]));
                        cc[m + 1] = o1;
                    }
                }
            }
        }
    }
}

// ---- host side -----------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn()
{
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried)
    {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

inline bool hemm_tma_disabled()
{
    static int v = -1;
    if (v < 0)
    {
        const char* e = getenv("CHASE_B200_NO_TMA");
        v = (e && atoi(e) != 0) ? 1 : 0;
    }
    return v == 1;
}

// FP64 storage only (TMA moves raw bytes; FP32 storage is widened by the generic kernel).
template <class T>
inline bool hemm_tma_supported(int64_t M, int64_t K, int64_t k, const void* A, int64_t lda, const void* B,
                               int64_t ldb, const void* C, int64_t ldc)
{
    if (!(std::is_same<T, double>::value || std::is_same<T, cxd>::value))
        return false;
    if (hemm_tma_disabled() || get_encode_fn() == nullptr)
        return false;
    const int64_t per16 = 16 / (int64_t)sizeof(T) > 0 ? 16 / (int64_t)sizeof(T) : 1; // elements per 16 B
    if (M < 256 || K < 256 || k < 8)
        return false; // tiny problems: launch-bound anyway
    if (lda % per16 || ldb % per16 || ldc % per16)
        return false;
    if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15)
        return false;
    return true;
}

template <class T>
inline int hemm_tma_launch(bool ta, int64_t M, int64_t K, int64_t k, typename Traits<T>::comp alpha, const T* A,
                           int64_t lda, const T* B, int64_t ldb, typename Traits<T>::comp beta, T* C, int64_t ldc,
                           double shift, const double* theta, cudaStream_t st)
{
    constexpr bool CPLX = Traits<T>::cplx;
    using CF = HemmCfg<CPLX>;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc)
        return -4;
    // complex<double> travels as 2 x FLOAT64 along the innermost dimension
    const cuuint64_t inner_mul = CPLX ? 2 : 1;
    CUtensorMap mapA, mapB;
    {
        // stored A is (M x K) column-major, or (K x M) when op(A) = A^H
        cuuint64_t dims[2] = {(cuuint64_t)(ta ? K : M) * inner_mul, (cuuint64_t)(ta ? M : K)};
        cuuint64_t strides[1] = {(cuuint64_t)lda * sizeof(T)};
        cuuint32_t box[2] = {16, (cuuint32_t)(ta ? CF::BM : CF::BK)}; // 16 doubles = 128 B
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)A, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
        {
            std::fprintf(stderr, "chase_b200: cuTensorMapEncodeTiled(A) failed: %d\n", (int)r);
            return -4;
        }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)K * inner_mul, (cuuint64_t)k};
        cuuint64_t strides[1] = {(cuuint64_t)ldb * sizeof(T)};
        cuuint32_t box[2] = {16, (cuuint32_t)CF::BN};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)B, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
        {
            std::fprintf(stderr, "chase_b200: cuTensorMapEncodeTiled(B) failed: %d\n", (int)r);
            return -4;
        }
    }
    HemmParams<T> p;
    p.M = M;
    p.K = K;
    p.k = k;
    p.B = B;
    p.ldb = ldb;
    p.C = C;
    p.ldc = ldc;
    p.alpha = alpha;
    p.beta = beta;
    p.shift = shift;
    p.theta = theta;
    p.tiles_m = (int)((M + CF::BM - 1) / CF::BM);
    p.tiles_n = (int)((k + CF::BN - 1) / CF::BN);
    int dev = 0, sms = 0;
    CB2_CUDA_OK(cudaGetDevice(&dev));
    CB2_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long ntiles = (long long)p.tiles_m * p.tiles_n;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    if (ta)
    {
        CB2_CUDA_OK(cudaFuncSetAttribute(hemm_tma_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         CF::SMEM_BYTES));
        hemm_tma_kernel<T, true><<<grid, CF::THREADS, CF::SMEM_BYTES, kcount(st)>>>(mapA, mapB, p);
    }
    else
    {
        CB2_CUDA_OK(cudaFuncSetAttribute(hemm_tma_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         CF::SMEM_BYTES));
        hemm_tma_kernel<T, false><<<grid, CF::THREADS, CF::SMEM_BYTES, kcount(st)>>>(mapA, mapB, p);
    }
    CB2_CUDA_OK(cudaGetLastError());
    return 0;
}

} // namespace cb2
