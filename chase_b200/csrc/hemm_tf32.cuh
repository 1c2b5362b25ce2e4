// chase_b200 — the filter product for the single-precision value types (float, complex<float>) on the
// 5th-generation tensor cores of sm_100a:
//
//   C <- alpha * S (A^s)^H S' B + beta * C - alpha * shift_j * B        A^s: K x M as stored (column-major)
//
// TMA (cp.async.bulk.tensor, SWIZZLE_128B, K-major boxes) -> shared memory -> tcgen05.mma kind::tf32 issued by one
// thread, FP32 accumulators in TMEM, tcgen05.ld epilogue.  Replaces the cublasSgemm / cublasCgemm the reference
// reaches from ChASEGPU::HEMM (/root/reference/Impl/chase_gpu/chase_gpu.hpp:656-678 through
// external/cublaspp/cublaspp.hpp:563, 623) and the round-1 detour through an FP64 copy of the matrix.
//
// Precision: TF32 keeps 11 significand bits, the parity bar for FP32 problems is 1e-4 on eigenvalues with residuals
// below 1e-5, so every operand is split x = hi + lo with hi = the TF32 truncation the tensor core applies to a raw
// FP32 container (low 13 bits ignored) and lo = rna_tf32(x - hi) (exact difference, rounded to TF32 once):
//     A B ~= A_hi B_hi + A_hi B_lo + A_lo B_hi (+ A_lo B_lo with terms = 4)
// accumulated in FP32 in TMEM.  A_hi IS the matrix as stored, so an FP32 problem keeps one extra FP32 array (A_lo,
// refreshed after every upload) instead of the FP64 copy (2x instead of 3x the matrix memory); the panel's lo part
// and, for complex types, its second real view are O(K k) scratch written per call.
//
// Only the K-major ("A^H B") operand orientation is built: a K-major box of the stored matrix is the tile of
// (A^s)^H up to conjugation.  That covers
//   * Hermitian A (single GPU):        A B   = A^H B
//   * the distributed V -> W step:     A_loc^H V
//   * pseudo-Hermitian H (BSE):        H B   = S H^H (S B)   (S H Hermitian; S = diag(I, -I) folded into the panel
//                                                              views and the epilogue: `sflip`)
// Complex arithmetic on the real tensor core: the interleaved (re, im) storage of a K-major tile is a real tile with
// 2K columns [ar0 ai0 ar1 ai1 ...]; with a = A^s(kk, i):  conj(a) b = (ar br + ai bi) + i (ar bi - ai br), so
//     Re C = A~ . B1~,  B1~ = [br,  bi]  (the panel as stored)
//     Im C = A~ . B2~,  B2~ = [bi, -br]  (second view, scratch)
// i.e. two real GEMMs sharing the A tile, accumulated in two TMEM column ranges.  The real types use the same
// structure with the two accumulator halves covering 2 x 128 panel columns.
//
// CTA = 384 threads: warp 0 TMA producer (one lane), warp 1 MMA issuer (one lane; allocates TMEM), warps 4-11
// epilogue (TMEM lane quarter = warp % 4, column half = (warp - 4) / 4).  Two shared-memory stages of 6 operand tiles
// (A hi/lo, Bx hi/lo, By hi/lo; 128 rows x 128 B each), two accumulator stages of 256 TMEM columns: the MMA chain of a
// tile is cut into chunks of KCH k-blocks that alternate between the stages, the epilogue warps drain a finished
// chunk into FP32 register sums (round-to-nearest) while the next chunk accumulates (see Tf32Cfg::KCH).
// Persistent over output tiles, n-fastest raster (CTAs resident together share A row blocks in L2).
#pragma once
#include "common.cuh"
#include "hemm_tma.cuh" // mbarrier / TMA helpers, cuTensorMapEncodeTiled entry point

namespace cb2
{

struct Tf32Cfg
{
    static constexpr int BM = 128;  // output rows per tile = TMEM lanes
    static constexpr int BNH = 128; // columns per accumulator half
    static constexpr int BKF = 32;  // floats per k-block: 128-byte rows
    static constexpr int STAGES = 2;
    static constexpr int OP_BYTES = 128 * BKF * 4; // one operand tile: 128 rows x 128 B
    static constexpr int NOPS = 6;
    static constexpr int STAGE_BYTES = NOPS * OP_BYTES;
    static constexpr int ACC_COLS = 2 * BNH; // TMEM columns per accumulator stage
    static constexpr int TMEM_COLS = 512;
    // The tensor core adds each MMA's products into the FP32 accumulator with truncation, so a long chain of
    // accumulations drifts (measured: relative error ~ #MMAs x 2^-25, 1.8e-4 at K = 4096).  The chain is therefore
    // cut every KCH k-blocks: the partial sum of a chunk is drained from TMEM by the epilogue warps and added to
    // register accumulators with round-to-nearest FADDs while the next chunk runs in the other TMEM stage.
    static constexpr int KCH = 8;
    // warpgroup 0: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation); warpgroups 1, 2: epilogue
    // (setmaxnreg moves registers from warpgroup 0 to the 128 FP32 running sums of every epilogue thread)
    static constexpr int THREADS = 384;
    static constexpr int NBARS = 2 * STAGES + 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 8 * NBARS + 16;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6) = 1, A/B = TF32 [7,10), [10,13) = 2,
    // K-major A and B (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BNH >> 3) << 17) |
                                      ((uint32_t)(BM >> 4) << 24);
};

template <class T>
struct Tf32Params
{
    long long M, K, k;   // C is M x k, the stored matrix is K x M, the panel K x k
    int kf;              // K in floats (2K for complex)
    const T* B;          // the panel as stored (shift term)
    long long ldb;
    T* C;
    long long ldc;
    double are, aim, bre, bim;
    double shift;
    const double* theta;
    long long sflip;     // > 0: rows >= sflip of the product are negated (S of the pseudo-Hermitian identity)
    int tiles_m, tiles_n;
    int terms;           // 3 or 4 partial products
};

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row x 128-byte atoms 1024 B apart (cute::UMMA::SmemDescriptor:
// start >> 4 at [0,14), LBO [16,30) unused for this layout, SBO >> 4 at [32,46), version 1 at [46,48), layout 2 at [61,64))
__device__ __forceinline__ uint64_t tf32_smem_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tf32_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "setp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
                 "}" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tf32_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 32 consecutive TMEM columns of this thread's lane -> registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                   "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <class T>
__global__ void __launch_bounds__(Tf32Cfg::THREADS, 1)
    hemm_tf32_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                     const __grid_constant__ CUtensorMap mapBxh, const __grid_constant__ CUtensorMap mapBxl,
                     const __grid_constant__ CUtensorMap mapByh, const __grid_constant__ CUtensorMap mapByl,
                     const Tf32Params<T> p)
{
    using CF = Tf32Cfg;
    constexpr bool CPLX = Traits<T>::cplx;
    constexpr int STAGES = CF::STAGES;
    constexpr int TILE_N = CPLX ? CF::BNH : 2 * CF::BNH; // panel columns per output tile

    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * CF::STAGE_BYTES;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES, bar_tfull = bars + 16 * STAGES,
                   bar_tempty = bars + 16 * STAGES + 16, tmem_slot = bars + 8 * CF::NBARS;
    unsigned char* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < STAGES; ++s)
        {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a)
        {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 8); // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)CF::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

    const int nkb = (p.kf + CF::BKF - 1) / CF::BKF;
    const long long ntiles = (long long)p.tiles_m * p.tiles_n;

    if (warp < 4)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;"); // warpgroup 0 hands its registers to the epilogue
    if (warp == 0)
    {
        // ------------------------------------------ TMA producer ------------------------------------------------
        if (lane == 0)
        {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAh) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAl) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBxh) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBxl) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapByh) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapByl) : "memory");
            uint32_t it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
            {
                const int tm = (int)(tile / p.tiles_n), tn = (int)(tile % p.tiles_n);
                const int m0 = tm * CF::BM, nx = tn * TILE_N, ny = CPLX ? nx : nx + CF::BNH;
                const bool y_on = CPLX || ny < p.k;
                for (int kb = 0; kb < nkb; ++kb, ++it)
                {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t full = bar_full + 8 * s;
                    mbar_expect_tx(full, (y_on ? 6u : 4u) * CF::OP_BYTES);
                    const uint32_t sa = base + s * CF::STAGE_BYTES;
                    const int kc = kb * CF::BKF;
                    tma_load_2d(sa + 0 * CF::OP_BYTES, &mapAh, full, kc, m0);
                    tma_load_2d(sa + 1 * CF::OP_BYTES, &mapAl, full, kc, m0);
                    tma_load_2d(sa + 2 * CF::OP_BYTES, &mapBxh, full, kc, nx);
                    tma_load_2d(sa + 3 * CF::OP_BYTES, &mapBxl, full, kc, nx);
                    if (y_on)
                    {
                        tma_load_2d(sa + 4 * CF::OP_BYTES, &mapByh, full, kc, ny);
                        tma_load_2d(sa + 5 * CF::OP_BYTES, &mapByl, full, kc, ny);
                    }
                }
            }
        }
        __syncwarp();
    }
    else if (warp < 4)
    {
        // ------------------------------------------ MMA issuer --------------------------------------------------
        if (warp == 1 && lane == 0)
        {
            uint32_t it = 0, ci = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
            {
                const int tn = (int)(tile % p.tiles_n);
                const int nx = tn * TILE_N, ny = CPLX ? nx : nx + CF::BNH;
                const bool y_on = CPLX || ny < p.k;
                for (int kb0 = 0; kb0 < nkb; kb0 += CF::KCH, ++ci)
                {
                    const uint32_t a = ci & 1, aph = (ci >> 1) & 1;
                    mbar_wait(bar_tempty + 8 * a, aph ^ 1); // the epilogue has drained this accumulator stage
                    tc_fence_after();
                    const uint32_t dx = tmem_base + a * CF::ACC_COLS, dy = dx + CF::BNH;
                    const int kb1 = kb0 + CF::KCH < nkb ? kb0 + CF::KCH : nkb;
                    for (int kb = kb0; kb < kb1; ++kb, ++it)
                    {
                        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                        mbar_wait(bar_full + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t sa = base + s * CF::STAGE_BYTES;
                        const uint64_t dAh = tf32_smem_desc(sa), dAl = tf32_smem_desc(sa + CF::OP_BYTES);
                        const uint64_t dXh = tf32_smem_desc(sa + 2 * CF::OP_BYTES),
                                       dXl = tf32_smem_desc(sa + 3 * CF::OP_BYTES);
                        const uint64_t dYh = tf32_smem_desc(sa + 4 * CF::OP_BYTES),
                                       dYl = tf32_smem_desc(sa + 5 * CF::OP_BYTES);
#pragma unroll
                        for (int kk = 0; kk < CF::BKF / 8; ++kk)
                        {
                            // one MMA covers 8 floats of K = 32 B: advance the start address inside the swizzle atom
                            const uint64_t o = (uint64_t)(kk * 2);
                            const uint32_t first = (kb == kb0 && kk == 0) ? 0u : 1u;
                            tf32_mma(dx, dAl + o, dXh + o, CF::IDESC, first);
                            tf32_mma(dx, dAh + o, dXl + o, CF::IDESC, 1u);
                            if (p.terms >= 4)
                                tf32_mma(dx, dAl + o, dXl + o, CF::IDESC, 1u);
                            tf32_mma(dx, dAh + o, dXh + o, CF::IDESC, 1u);
                            if (y_on)
                            {
                                tf32_mma(dy, dAl + o, dYh + o, CF::IDESC, first);
                                tf32_mma(dy, dAh + o, dYl + o, CF::IDESC, 1u);
                                if (p.terms >= 4)
                                    tf32_mma(dy, dAl + o, dYl + o, CF::IDESC, 1u);
                                tf32_mma(dy, dAh + o, dYh + o, CF::IDESC, 1u);
                            }
                        }
                        tf32_commit(bar_empty + 8 * s); // the stage is free once these MMAs have read it
                    }
                    tf32_commit(bar_tfull + 8 * a); // this chunk's partial sums are complete
                }
            }
        }
        __syncwarp();
    }
    else
    {
        // ------------------------------------------ epilogue ----------------------------------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int q = warp & 3;        // this warp may touch TMEM lanes [32 q, 32 q + 32)
        const int g = (warp - 4) >> 2; // column group
        const double are = p.are, aim = p.aim, bre = p.bre, bim = p.bim;
        const bool has_beta = (bre != 0.0) || (bim != 0.0);
        const bool has_shift = (p.theta != nullptr) || (p.shift != 0.0);
        // four 32-column segments per thread.  real: segment s = accumulator columns 128 g + 32 s = panel columns
        // nx + 128 g + 32 s.  complex: s = 0, 1: Re part, accumulator columns 64 g + 32 s; s = 2, 3: Im part of the
        // same panel columns nx + 64 g + 32 (s & 1), accumulator columns 128 + 64 g + 32 (s - 2).
        uint32_t ci = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        {
            const int tm = (int)(tile / p.tiles_n), tn = (int)(tile % p.tiles_n);
            const long long m = (long long)tm * CF::BM + q * 32 + lane;
            const long long nx = (long long)tn * TILE_N;
            const long long nseg0 = nx + (CPLX ? 64 : 128) * g; // first panel column of this thread's segments
            float acc[4][32];
            for (int kb0 = 0; kb0 < nkb; kb0 += CF::KCH, ++ci)
            {
                const uint32_t a = ci & 1, aph = (ci >> 1) & 1;
                mbar_wait(bar_tfull + 8 * a, aph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + a * CF::ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int sg = 0; sg < 4; ++sg)
                {
                    const int pcol = CPLX ? 32 * (sg & 1) : 32 * sg;                       // panel column offset
                    const int tcol = CPLX ? (sg >> 1) * CF::BNH + 64 * g + 32 * (sg & 1) : 128 * g + 32 * sg;
                    if (nseg0 + pcol < p.k) // warp-uniform
                    {
                        uint32_t r[32];
                        tmem_ld32(taddr + tcol, r);
                        tmem_ld_wait();
                        if (kb0 == 0)
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                acc[sg][j] = __uint_as_float(r[j]);
                        }
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                acc[sg][j] += __uint_as_float(r[j]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(bar_tempty + 8 * a);
            }
            if (m >= p.M)
                continue;
            const double sgn = (p.sflip > 0 && m >= p.sflip) ? -1.0 : 1.0;
            if constexpr (!CPLX)
            {
#pragma unroll
                for (int sg = 0; sg < 4; ++sg)
                {
                    if (nseg0 + 32 * sg >= p.k)
                        continue;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                    {
                        const long long n = nseg0 + 32 * sg + j;
                        if (n < p.k)
                        {
                            double o = are * sgn * (double)acc[sg][j];
                            T* cc = p.C + n * p.ldc + m;
                            if (has_beta)
                                o += bre * (double)*cc;
                            if (has_shift)
                                o -= are * (p.theta ? p.theta[n] : p.shift) * (double)p.B[n * p.ldb + m];
                            *cc = (float)o;
                        }
                    }
                }
            }
            else
            {
#pragma unroll
                for (int sg = 0; sg < 2; ++sg)
                {
                    if (nseg0 + 32 * sg >= p.k)
                        continue;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                    {
                        const long long n = nseg0 + 32 * sg + j;
                        if (n < p.k)
                        {
                            const double pr = sgn * (double)acc[sg][j];
                            const double pi = sgn * (double)acc[sg + 2][j];
                            double o_re = are * pr - aim * pi, o_im = are * pi + aim * pr;
                            cxf* cc = reinterpret_cast<cxf*>(p.C) + n * p.ldc + m;
                            if (has_beta)
                            {
                                const cxf cv = *cc;
                                o_re += bre * (double)cv.re - bim * (double)cv.im;
                                o_im += bre * (double)cv.im + bim * (double)cv.re;
                            }
                            if (has_shift)
                            {
                                const double sh = p.theta ? p.theta[n] : p.shift;
                                const cxf bv = reinterpret_cast<const cxf*>(p.B)[n * p.ldb + m];
                                const double gr = -sh * are, gi = -sh * aim;
                                o_re += gr * (double)bv.re - gi * (double)bv.im;
                                o_im += gr * (double)bv.im + gi * (double)bv.re;
                            }
                            *cc = cxf{(float)o_re, (float)o_im};
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1)
    {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)CF::TMEM_COLS)
                     : "memory");
    }
}

// ---- operand preparation ----------------------------------------------------------------------------------------
// lo part of a TF32 split whose hi part is the truncation the tensor core applies to the raw container
__device__ __forceinline__ float tf32_lo(float x)
{
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x - hi));
    return __uint_as_float(r);
}

// dst[i] = lo(src[i]) over a contiguous array of floats (the matrix with its padded leading dimension)
__global__ void __launch_bounds__(256) tf32_split_lo_kernel(long long n4, const float4* __restrict__ src,
                                                             float4* __restrict__ dst)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    {
        const float4 v = src[i];
        dst[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
    }
}

// lo parts of selected elements (linear element indices into the matrix): the distributed backend shifts the diagonal
// entries of its local block in place (pChASEGPU::Shift) and refreshes their lo parts with this
__global__ void __launch_bounds__(256) tf32_split_lo_list_kernel(long long cnt, const long long* __restrict__ lin,
                                                                  int floats_per_elem, const float* __restrict__ src,
                                                                  float* __restrict__ dst)
{
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += (long long)gridDim.x * blockDim.x)
        for (int f = 0; f < floats_per_elem; ++f)
            dst[lin[t] * floats_per_elem + f] = tf32_lo(src[lin[t] * floats_per_elem + f]);
}

// Panel views.  Real:    v0 = lo(s B)                                (hi = B itself; s B copied to v1 when sflip)
//               Complex: v0 = lo(s B), v1 = [s bi, -s br], v2 = lo(v1), v3 = s B (only when sflip)
// s = -1 on rows >= sflip (sflip > 0), else 1.
template <class T>
__global__ void __launch_bounds__(256) tf32_panel_views_kernel(long long rows, long long cols, const T* __restrict__ B,
                                                                long long ldb, T* __restrict__ v0, T* __restrict__ v1,
                                                                T* __restrict__ v2, T* __restrict__ v3, long long ldv,
                                                                long long sflip)
{
    const long long j = blockIdx.y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x)
    {
        const float s = (sflip > 0 && i >= sflip) ? -1.0f : 1.0f;
        if constexpr (!Traits<T>::cplx)
        {
            const float b = s * B[i + j * ldb];
            v0[i + j * ldv] = tf32_lo(b);
            if (sflip > 0)
                v1[i + j * ldv] = b;
        }
        else
        {
            const cxf b0 = B[i + j * ldb];
            const cxf b = cxf{s * b0.re, s * b0.im};
            v0[i + j * ldv] = cxf{tf32_lo(b.re), tf32_lo(b.im)};
            v1[i + j * ldv] = cxf{b.im, -b.re};
            v2[i + j * ldv] = cxf{tf32_lo(b.im), -tf32_lo(b.re)};
            if (sflip > 0)
                v3[i + j * ldv] = b;
        }
    }
}

inline size_t hemm_tf32_scratch_bytes(int64_t K, int64_t k, int elem_bytes)
{
    const int64_t ldv = (K + 15) / 16 * 16;
    return (size_t)4 * (size_t)ldv * (size_t)k * (size_t)elem_bytes;
}

inline bool tf32_encode(CUtensorMap* map, const void* ptr, uint64_t inner_floats, uint64_t outer, uint64_t stride_bytes)
{
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc)
        return false;
    cuuint64_t dims[2] = {(cuuint64_t)inner_floats, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)stride_bytes};
    cuuint32_t box[2] = {(cuuint32_t)Tf32Cfg::BKF, 128u};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        std::fprintf(stderr, "chase_b200: cuTensorMapEncodeTiled (tf32 operand) failed: %d\n", (int)r);
        return false;
    }
    return true;
}

template <class T>
inline bool hemm_tf32_supported(int64_t M, int64_t K, int64_t k, const void* A, const void* Alo, int64_t lda,
                                const void* B, int64_t ldb, const void* C, int64_t ldc)
{
    if (get_encode_fn() == nullptr || Alo == nullptr)
        return false;
    if (M < 128 || K < 128 || k < 1)
        return false;
    const int64_t per16 = 16 / (int64_t)sizeof(T) > 0 ? 16 / (int64_t)sizeof(T) : 1;
    if (lda % per16 || ldb % per16 || ldc % per16)
        return false;
    if (((uintptr_t)A | (uintptr_t)Alo | (uintptr_t)B | (uintptr_t)C) & 15)
        return false;
    return true;
}

// C(M x k) <- alpha S (A^s)^H S B + beta C - alpha shift_j B;  A^s, Alo: K x M column-major (lda); scratch: see
// hemm_tf32_scratch_bytes.  sflip = 0: no S.
template <class T>
inline int hemm_tf32_launch(int64_t M, int64_t K, int64_t k, double are, double aim, const T* A, const T* Alo,
                            int64_t lda, const T* B, int64_t ldb, double bre, double bim, T* C, int64_t ldc, double shift,
                            const double* theta, int64_t sflip, int terms, void* scratch, size_t scratch_bytes,
                            cudaStream_t st)
{
    using CF = Tf32Cfg;
    constexpr bool CPLX = Traits<T>::cplx;
    if (M <= 0 || k <= 0)
        return 0;
    if (scratch_bytes < hemm_tf32_scratch_bytes(K, k, (int)sizeof(T)))
        return -3;
    if ((theta || shift != 0.0) && M != K)
        return -2;
    const int64_t ldv = (K + 15) / 16 * 16;
    T* v0 = (T*)scratch;
    T* v1 = v0 + ldv * k;
    T* v2 = v1 + ldv * k;
    T* v3 = v2 + ldv * k;
    {
        dim3 grid((unsigned)std::min<int64_t>((K + 255) / 256, 64), (unsigned)k);
        tf32_panel_views_kernel<T><<<grid, 256, 0, kcount(st)>>>(K, k, B, ldb, v0, v1, v2, v3, ldv, sflip);
        CB2_CUDA_OK(cudaGetLastError());
    }
    const uint64_t fm = CPLX ? 2 : 1;
    CUtensorMap mAh, mAl, mXh, mXl, mYh, mYl;
    bool ok = tf32_encode(&mAh, A, (uint64_t)K * fm, (uint64_t)M, (uint64_t)lda * sizeof(T)) &&
              tf32_encode(&mAl, Alo, (uint64_t)K * fm, (uint64_t)M, (uint64_t)lda * sizeof(T));
    if constexpr (!CPLX)
    {
        const T* hi = sflip > 0 ? v1 : B;
        const int64_t ldh = sflip > 0 ? ldv : ldb;
        ok = ok && tf32_encode(&mXh, hi, (uint64_t)K, (uint64_t)k, (uint64_t)ldh * sizeof(T)) &&
             tf32_encode(&mXl, v0, (uint64_t)K, (uint64_t)k, (uint64_t)ldv * sizeof(T));
        mYh = mXh;
        mYl = mXl;
    }
    else
    {
        const T* hi = sflip > 0 ? v3 : B;
        const int64_t ldh = sflip > 0 ? ldv : ldb;
        ok = ok && tf32_encode(&mXh, hi, (uint64_t)K * 2, (uint64_t)k, (uint64_t)ldh * sizeof(T)) &&
             tf32_encode(&mXl, v0, (uint64_t)K * 2, (uint64_t)k, (uint64_t)ldv * sizeof(T)) &&
             tf32_encode(&mYh, v1, (uint64_t)K * 2, (uint64_t)k, (uint64_t)ldv * sizeof(T)) &&
             tf32_encode(&mYl, v2, (uint64_t)K * 2, (uint64_t)k, (uint64_t)ldv * sizeof(T));
    }
    if (!ok)
        return -4;
    Tf32Params<T> p;
    p.M = M;
    p.K = K;
    p.k = k;
    p.kf = (int)(K * (int64_t)fm);
    p.B = B;
    p.ldb = ldb;
    p.C = C;
    p.ldc = ldc;
    p.are = are;
    p.aim = aim;
    p.bre = bre;
    p.bim = bim;
    p.shift = shift;
    p.theta = theta;
    p.sflip = sflip;
    p.terms = terms >= 4 ? 4 : 3;
    p.tiles_m = (int)((M + CF::BM - 1) / CF::BM);
    const int tile_n = CPLX ? CF::BNH : 2 * CF::BNH;
    p.tiles_n = (int)((k + tile_n - 1) / tile_n);
    int dev = 0, sms = 0;
    CB2_CUDA_OK(cudaGetDevice(&dev));
    CB2_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long ntiles = (long long)p.tiles_m * p.tiles_n;
    const int grid = (int)std::min<long long>(ntiles, sms);
    CB2_CUDA_OK(cudaFuncSetAttribute(hemm_tf32_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
    hemm_tf32_kernel<T><<<grid, CF::THREADS, CF::SMEM_BYTES, kcount(st)>>>(mAh, mAl, mXh, mXl, mYh, mYl, p);
    CB2_CUDA_OK(cudaGetLastError());
    return 0;
}

} // namespace cb2
