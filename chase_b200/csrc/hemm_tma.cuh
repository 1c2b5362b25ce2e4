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
#include <map>
#include <mutex>
#include <type_traits>
#include <utility>

namespace cb2
{

// NARROW selects the tile width: 0 = the default tile (128 columns real, 64 complex), 1 = half of it, 2 = a quarter
// (real only).  The 8 consumer warps are re-arranged so that the warp tile stays 32 columns wide: 2 x 4 warps of 64 x 32
// (real default), 4 x 2 of 32 x 32, 8 x 1 of 16 x 32.  The narrow tiles serve the ragged last columns of a panel
// (hemm_tma_launch splits it off), where the default tile would leave whole warps idle.
template <bool CPLX, int NARROW = 0>
struct HemmCfg
{
    static_assert(NARROW >= 0 && NARROW <= (CPLX ? 1 : 2), "tile variant");
    static constexpr int ELEM = CPLX ? 16 : 8;
    static constexpr int EPB = 128 / ELEM; // elements per 128-byte row
    static constexpr int BM = 128;
    static constexpr int BN = (CPLX ? 64 : 128) >> NARROW;
    static constexpr int BK = EPB;
    static constexpr int WM = (CPLX ? 32 : 64) >> NARROW;
    static constexpr int WN = 32;
    static constexpr int WARPS_M = BM / WM;
    static constexpr int KSTEPS = BK / 4;
    static constexpr int A_BYTES = BM * BK * ELEM;
    static constexpr int B_BYTES = BN * BK * ELEM;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = CPLX ? 8 : 6;
    static constexpr int CONSUMER_WARPS = 8;
    // 8 consumer warps (2 warpgroups) + 1 producer warpgroup of which one lane works.  Register budget is per SM
    // sub-partition: setmaxnreg moves registers from the producer warpgroup (40) to the consumers (232), so the 64
    // FP64 accumulators + fragments + hoisted offsets of a consumer thread never spill.
    static constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;
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
    long long span;      // k-blocks per CTA (stream-K); a multiple of nkt means whole tiles only
    int remap;           // != 0: virtual -> raster tile index through hemm_tile_remap (stream-K launches)
    long long sk_tiles;  // tiles covered by the stream-K phase (all of them unless the hybrid schedule is on)
    int dp_waves;        // whole-tile waves after the stream-K phase (tile = sk_tiles + w * gridDim.x + blockIdx.x)
    double* scratch;     // gridDim.x slots of BM*BN accumulators for incomplete tiles
    unsigned* flags;     // one per CTA: epoch of the launch whose head part is parked in the slot
    unsigned epoch;
};

// Stream-K work distribution.  The iteration space (tiles x k-blocks) is cut into gridDim.x equal contiguous spans,
// so every CTA executes the same number of k-blocks (no wave quantisation: at N=20000, k=1400 the 1727 tiles were
// 11.67 waves on 148 CTAs).  A span covers at most one incomplete tile at each end (host guarantees span >= one tile):
// the CTA walks its span from the TOP, so the head part of a shared tile (k-blocks [0, x)) is produced first and
// parked in a per-CTA scratch slot; the CTA that owns the tail part [x, nkt) reaches it LAST, adds the parked
// partial sums and runs the epilogue.  Summation order is fixed (head + tail), so results are bit-reproducible.
struct HemmSpan
{
    long long tile;
    int kt_begin, kt_end;
};

// L2 locality of the stream-K walk.  Spans are contiguous in the tile index, so with the plain n-fastest raster every
// CTA streams "its own" ~T/G consecutive tiles = one row block of A after the other, the G co-resident CTAs touch G
// different row blocks at any time and nothing is shared in L2 (measured at N=20000, k=1400: 62 GB of DRAM reads for
// 3.9 GB of operands).  The spans therefore walk a VIRTUAL tile index v, mapped to the raster index r by a bijection
// that puts the s-th tile of every CTA next to each other: with start(c) = floor(c T / G) (first virtual tile of CTA
// c), q = floor(T / G), v = start(c) + s:
//     r = s G + c                      for s < q      (all CTAs have a tile at position s)
//     r = q G + (start(c) - c q)       for s = q      (only the CTAs with q + 1 tiles; rank among those)
// so the CTAs that run position s at the same time cover G consecutive raster tiles (13 row blocks x 11 column tiles
// at C2) and share each A row block tiles_n-fold, like a wave of the data-parallel schedule.  Depends on v only, so
// the two CTAs that share a split tile agree on it.
__host__ __device__ inline long long hemm_tile_remap(long long v, long long T, long long G)
{
    if (G <= 1 || T <= G)
        return v;
    long long c = (v * G) / T;
    while (c + 1 < G && ((c + 1) * T) / G <= v)
        ++c;
    while (c > 0 && (c * T) / G > v)
        --c;
    const long long start = (c * T) / G, q = T / G, s_ = v - start;
    return (s_ < q) ? s_ * G + c : q * G + (start - c * q);
}

// The sequence of tile parts one CTA executes; producer and consumer walk it in lock-step, and the host test
// (chase_b200_hemm_walk) replays it on the CPU.
//   phase 1 (stream-K): the first sk_tiles tiles x nkt k-blocks are cut into G equal spans; CTA b walks its span
//            [b span, (b+1) span) from the TOP, so a head part (kt_end < nkt, parked for CTA b+1) comes first and a
//            tail part (kt_begin > 0, waits for CTA b-1) last.  span >= nkt => at most two parts per tile.
//   phase 2 (data-parallel, dp_waves > 0): whole tiles sk_tiles + w G + b, w = 0 .. dp_waves-1.  All CTAs enter this
//            phase after the same number of k-blocks, so CTAs that share a row block of A read the same k range at
//            the same time (L2 sharing like a plain persistent schedule); phase 1 only absorbs the ragged last wave.
struct HemmWalk
{
    long long it_begin, hi, sk_tiles, G, b;
    int nkt, dp_waves, w, remap;
    __host__ __device__ HemmWalk(long long b_, long long G_, long long span, long long sk_tiles_, int nkt_, int dp_waves_,
                                 int remap_)
        : sk_tiles(sk_tiles_), G(G_), b(b_), nkt(nkt_), dp_waves(dp_waves_), w(0), remap(remap_)
    {
        const long long total = sk_tiles * nkt;
        it_begin = b * span < total ? b * span : total;
        hi = it_begin + span < total ? it_begin + span : total;
    }
    // next part: raster tile index and k-block range; false when the CTA is done
    __host__ __device__ bool next(HemmSpan& sp)
    {
        if (hi > it_begin)
        {
            const long long vt = (hi - 1) / nkt, first = vt * nkt;
            const long long lo = it_begin > first ? it_begin : first;
            sp.kt_begin = (int)(lo - first);
            sp.kt_end = (int)(hi - first);
            sp.tile = remap ? hemm_tile_remap(vt, sk_tiles, G) : vt;
            hi = lo;
            return true;
        }
        if (w < dp_waves)
        {
            sp.tile = sk_tiles + (long long)w * G + b;
            sp.kt_begin = 0;
            sp.kt_end = nkt;
            ++w;
            return true;
        }
        return false;
    }
};

template <class T, bool TA, int NARROW = 0>
__global__ void __launch_bounds__(HemmCfg<Traits<T>::cplx, NARROW>::THREADS, 1)
    hemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const HemmParams<T> p)
{
    using TR = Traits<T>;
    using C_ = typename TR::comp;
    constexpr bool CPLX = TR::cplx;
    using CF = HemmCfg<CPLX, NARROW>;
    constexpr int BM = CF::BM, BN = CF::BN, BK = CF::BK, WM = CF::WM, WN = CF::WN, EPB = CF::EPB, ELEM = CF::ELEM;
    constexpr int MI = WM / 8, NJ = WN / 8, STAGES = CF::STAGES, KSTEPS = CF::KSTEPS;
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

    const int nkt = (int)((p.K + BK - 1) / BK);
    HemmWalk walk((long long)blockIdx.x, (long long)gridDim.x, p.span, p.sk_tiles, nkt, p.dp_waves, p.remap);

    if (warp >= CF::CONSUMER_WARPS)
    {
        // ------------------------------ producer ------------------------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == CF::CONSUMER_WARPS && lane == 0)
        {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
            uint32_t it = 0;
            HemmSpan sp;
            while (walk.next(sp))
            {
                const int tn = (int)(sp.tile % p.tiles_n), tm = (int)(sp.tile / p.tiles_n);
                const int m0 = tm * BM, n0 = tn * BN;
                for (int kt = sp.kt_begin; kt < sp.kt_end; ++kt, ++it)
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
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int wm = warp % CF::WARPS_M, wn = warp / CF::WARPS_M;
    const int q = lane & 3, c = lane >> 2;
    const int ncol = (c & 1) | ((c & 2) << 1) | ((c & 4) >> 1); // {0,1,4,5,2,3,6,7}
    // per-lane k offsets inside an 8-group for the two k-sets
    const int kq0 = 2 * q + (q & 1), kq1 = 2 * q + ((q & 1) ^ 1);

    // Byte offsets of this lane's fragments inside a stage, hoisted out of the k loop.  The swizzle XOR makes them
    // non-affine in (i, ks) only through (i & AH) and ks; everything else is a compile-time stride that ends up as
    // an immediate of the LDS.
    //   M-major A tile (op N): m = wm*WM + 8i + c ; box = m / EPB, row = k, 16-byte chunk = (m % EPB * ELEM / 16) ^ (k & 7)
    //   K-major tiles (B, and A for op C): row = n (or m), chunk = (k * ELEM / 16) ^ (row & 7)
    constexpr int AH = (!TA && !CPLX) ? 2 : 1;                    // distinct i-classes of the M-major real tile
    constexpr int A_ISTRIDE = TA ? 8 * 128 : (CPLX ? BK * 128 : BK * 128); // per AH-group of i (see below)
    int a_off[AH][KSTEPS], b_off[KSTEPS];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
    {
        const int kin = 8 * (ks >> 1) + ((ks & 1) ? kq1 : kq0);
        const int k7 = kin & 7;
#pragma unroll
        for (int h = 0; h < AH; ++h)
        {
            if constexpr (TA)
            {
                const int m = wm * WM + ncol;
                const int byte_in = kin * ELEM;
                a_off[h][ks] = m * 128 + ((((byte_in >> 4) ^ (m & 7)) << 4) | (byte_in & 15));
            }
            else
            {
                const int m = wm * WM + 8 * h + c;
                const int mblk = m / EPB, min_ = m % EPB;
                const int byte_in = min_ * ELEM;
                a_off[h][ks] = (mblk * BK + kin) * 128 + ((((byte_in >> 4) ^ k7) << 4) | (byte_in & 15));
            }
        }
        {
            const int n = wn * WN + ncol;
            const int byte_in = kin * ELEM;
            b_off[ks] = CF::A_BYTES + n * 128 + ((((byte_in >> 4) ^ (n & 7)) << 4) | (byte_in & 15));
        }
    }

    uint32_t it = 0;
    HemmSpan sp;
    while (walk.next(sp))
    {
        const int tn = (int)(sp.tile % p.tiles_n), tm = (int)(sp.tile / p.tiles_n);
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

        // Ragged last N tile: 8-column groups beyond k are skipped (warp-uniform), so a tile with <= BN/2 valid columns
        // keeps only one warp per SM sub-partition busy and costs about half a tile.
        const int nvalid = (int)((p.k - n0) < (long long)BN ? (p.k - n0) : (long long)BN);
        int jmax = (nvalid - wn * WN + 7) / 8;
        jmax = jmax < 0 ? 0 : (jmax > NJ ? NJ : jmax);

        auto k_block = [&](const unsigned char* st, auto full_tag)
        {
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
            {
                C_ fa[MI], fb[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i)
                {
                    // real M-major: i = 2g + h -> box g (+BK*128 bytes each); complex M-major: box i; K-major: row 8i
                    const int h = (AH == 2) ? (i & 1) : 0;
                    const int g = (AH == 2) ? (i >> 1) : i;
                    fa[i] = *reinterpret_cast<const C_*>(st + a_off[h][ks] + g * A_ISTRIDE);
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    if (FULL || j < jmax)
                        fb[j] = *reinterpret_cast<const C_*>(st + b_off[ks] + j * (8 * 128));
                if constexpr (!CPLX)
                {
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                        if (FULL || j < jmax)
                        {
#pragma unroll
                            for (int i = 0; i < MI; ++i)
                                dmma884(accr[j][i][0], accr[j][i][1], fb[j], fa[i]);
                        }
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                        if (FULL || j < jmax)
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
        };

        for (int kt = sp.kt_begin; kt < sp.kt_end; ++kt, ++it)
        {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(bars + 8 * s, ph);
            const unsigned char* st = gen_base + s * CF::STAGE_BYTES;
            if (jmax == NJ)
                k_block(st, std::true_type{});
            else if (jmax > 0)
                k_block(st, std::false_type{});
            __syncwarp();
            if (lane == 0)
                mbar_arrive(bars + 8 * (STAGES + s));
        }

        // ---------------- stream-K hand-over of incomplete tiles (consumer threads only: barrier 1) -----------
        constexpr int NACC = NJ * MI * 2 * (CPLX ? 2 : 1);
        const int ctid = tid; // consumers are threads 0..255
        if (sp.kt_end < nkt)
        {
            // head part: park the raw accumulators for the CTA that owns the tail (blockIdx.x + 1)
            double* slot = p.scratch + (size_t)blockIdx.x * (NACC * CF::CONSUMER_WARPS * 32);
            int r = 0;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                    {
                        __stcg(slot + (size_t)(r++) * 256 + ctid, accr[j][i][h]);
                        if constexpr (CPLX)
                            __stcg(slot + (size_t)(r++) * 256 + ctid, acci[j][i][h]);
                    }
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (ctid == 0)
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.flags + blockIdx.x), "r"(p.epoch) : "memory");
            continue;
        }
        if (sp.kt_begin > 0)
        {
            // tail part: wait for the head parked by CTA blockIdx.x - 1 and add it
            if (ctid == 0)
            {
                unsigned v;
                do
                {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.flags + blockIdx.x - 1) : "memory");
                } while (v != p.epoch);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const double* slot = p.scratch + (size_t)(blockIdx.x - 1) * (NACC * CF::CONSUMER_WARPS * 32);
            int r = 0;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                    {
                        accr[j][i][h] += __ldcg(slot + (size_t)(r++) * 256 + ctid);
                        if constexpr (CPLX)
                            acci[j][i][h] += __ldcg(slot + (size_t)(r++) * 256 + ctid);
                    }
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
                               int64_t ldb, const void* C, int64_t ldc, int64_t min_m = 256, int64_t min_k = 256)
{
    if (!(std::is_same<T, double>::value || std::is_same<T, cxd>::value))
        return false;
    if (hemm_tma_disabled() || get_encode_fn() == nullptr)
        return false;
    const int64_t per16 = 16 / (int64_t)sizeof(T) > 0 ? 16 / (int64_t)sizeof(T) : 1; // elements per 16 B
    if (M < min_m || K < min_k || k < 8)
        return false; // tiny problems: launch-bound anyway
    if (lda % per16 || ldb % per16 || ldc % per16)
        return false;
    if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15)
        return false;
    return true;
}

// Scratch for the stream-K hand-over: one slot of BM*BN accumulators and one flag per CTA, allocated once per
// (device, stream) so that launches on different streams never share slots.
struct HemmScratch
{
    double* slots = nullptr;
    unsigned* flags = nullptr;
    unsigned epoch = 0;
};
inline std::mutex& hemm_scratch_mutex()
{
    static std::mutex m;
    return m;
}
inline std::map<std::pair<int, cudaStream_t>, HemmScratch>& hemm_scratch_pool()
{
    static std::map<std::pair<int, cudaStream_t>, HemmScratch> pool;
    return pool;
}
inline HemmScratch* hemm_scratch(int dev, cudaStream_t st, int sms)
{
    std::lock_guard<std::mutex> lk(hemm_scratch_mutex());
    auto& pool = hemm_scratch_pool();
    auto key = std::make_pair(dev, st);
    auto it = pool.find(key);
    if (it != pool.end())
        return &it->second; // std::map nodes are stable: the pointer stays valid while other streams are added
    HemmScratch sc;
    const size_t slot_doubles = (size_t)HemmCfg<false>::BM * HemmCfg<false>::BN; // == complex BM*BN*2
    if (cudaMalloc(&sc.slots, (size_t)sms * slot_doubles * sizeof(double)) != cudaSuccess)
        return nullptr;
    if (cudaMalloc(&sc.flags, (size_t)sms * sizeof(unsigned)) != cudaSuccess)
        return nullptr;
    if (cudaMemset(sc.flags, 0, (size_t)sms * sizeof(unsigned)) != cudaSuccess)
        return nullptr;
    return &(pool[key] = sc);
}
// called by the owner of a stream before it destroys it: frees the stream's slots (a later stream whose handle value
// happens to be re-used must not inherit them)
inline void hemm_scratch_release(cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(hemm_scratch_mutex());
    auto& pool = hemm_scratch_pool();
    for (auto it = pool.begin(); it != pool.end();)
    {
        if (it->first.second == st)
        {
            cudaFree(it->second.slots);
            cudaFree(it->second.flags);
            it = pool.erase(it);
        }
        else
            ++it;
    }
}
// CHASE_B200_HEMM_REMAP=0 walks the plain raster (diagnostics); read at every launch
inline bool hemm_remap_disabled()
{
    const char* e = getenv("CHASE_B200_HEMM_REMAP");
    return e && atoi(e) == 0;
}
// Hybrid schedule: stream-K only for the ragged end, k-aligned whole-tile waves before it.  Default: on for both operand
// orientations.  Measured DRAM reads per launch at unchanged time (ncu, profiles/r2_hemm_rect_traffic.csv): op(A) = A,
// N=20000, k=1400: 8.4 GB instead of 42 GB; local block 10000 x 20000 of the 2-GPU grid: op(A) = A 7.9 instead of 30.3 GB,
// op(A) = A^H 3.7 instead of 15.3 GB; 10000 x 10000 (4 GPUs): 3.3 instead of 12.6-12.8 GB.  CHASE_B200_HEMM_HYBRID=0/1
// forces it either way.
inline bool hemm_hybrid_enabled(bool /*ta*/)
{
    const char* e = getenv("CHASE_B200_HEMM_HYBRID");
    return e ? atoi(e) != 0 : true;
}
inline bool hemm_streamk_disabled()
{
    static int v = -1;
    if (v < 0)
    {
        const char* e = getenv("CHASE_B200_NO_STREAMK");
        v = (e && atoi(e) != 0) ? 1 : 0;
    }
    return v == 1;
}

// Launch geometry of the stream-K / hybrid schedule (shared with the host replay used by the tests)
inline void hemm_schedule(long long ntiles, long long nkt, int sms, bool ta, int& grid, long long& span,
                          long long& sk_tiles, int& dp_waves, int& remap)
{
    remap = 0;
    dp_waves = 0;
    sk_tiles = ntiles;
    if (ntiles >= sms && !hemm_streamk_disabled())
    {
        grid = sms;
        remap = hemm_remap_disabled() ? 0 : 1;
        // hybrid: stream-K over the last full wave + the ragged rest (between G and 2 G tiles, so span >= nkt still
        // holds), whole-tile waves before it
        if (hemm_hybrid_enabled(ta) && ntiles >= 2 * (long long)sms)
        {
            dp_waves = (int)(ntiles / sms) - 1;
            sk_tiles = ntiles - (long long)dp_waves * sms;
        }
        // equal spans of k-blocks, at most one incomplete tile at each end of a span
        span = (sk_tiles * nkt + grid - 1) / grid;
    }
    else
    {
        // whole tiles per CTA (round-robin over tiles would need more than one span per CTA: use contiguous tiles)
        grid = (int)(ntiles < sms ? ntiles : sms);
        const long long tiles_per_cta = (ntiles + grid - 1) / grid;
        span = tiles_per_cta * nkt;
    }
}

template <class T, int NARROW>
inline int hemm_tma_launch_cfg(bool ta, int64_t M, int64_t K, int64_t k, typename Traits<T>::comp alpha, const T* A,
                               int64_t lda, const T* B, int64_t ldb, typename Traits<T>::comp beta, T* C, int64_t ldc,
                               double shift, const double* theta, cudaStream_t st)
{
    constexpr bool CPLX = Traits<T>::cplx;
    using CF = HemmCfg<CPLX, NARROW>;
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
    const long long nkt = (K + CF::BK - 1) / CF::BK;
    int grid;
    hemm_schedule(ntiles, nkt, sms, ta, grid, p.span, p.sk_tiles, p.dp_waves, p.remap);
    HemmScratch* sc = hemm_scratch(dev, st, sms);
    if (!sc)
        return -1;
    p.scratch = sc->slots;
    p.flags = sc->flags;
    p.epoch = ++sc->epoch;
    if (ta)
    {
        CB2_CUDA_OK(cudaFuncSetAttribute(hemm_tma_kernel<T, true, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         CF::SMEM_BYTES));
        hemm_tma_kernel<T, true, NARROW><<<grid, CF::THREADS, CF::SMEM_BYTES, kcount(st)>>>(mapA, mapB, p);
    }
    else
    {
        CB2_CUDA_OK(cudaFuncSetAttribute(hemm_tma_kernel<T, false, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         CF::SMEM_BYTES));
        hemm_tma_kernel<T, false, NARROW><<<grid, CF::THREADS, CF::SMEM_BYTES, kcount(st)>>>(mapA, mapB, p);
    }
    CB2_CUDA_OK(cudaGetLastError());
    return 0;
}

// CHASE_B200_HEMM_SPLIT=0: one launch with the default tile for the whole panel (diagnostics); read at every launch
inline bool hemm_split_enabled()
{
    const char* e = getenv("CHASE_B200_HEMM_SPLIT");
    return e ? atoi(e) != 0 : true;
}

// The panel is cut into the columns that fill whole default tiles and the ragged rest.  Inside one launch every tile
// costs the same, so the stream-K / hybrid schedule balances exactly; the rest runs as a second launch on the
// narrowest tile that covers it (all 8 warps busy).  In a single launch the ragged tiles were booked as full tiles:
// k = 1342 took as long as k = 1408 and k = 419 ran at 29 TFLOP/s where cuBLAS reaches 33.
template <class T>
inline int hemm_tma_launch(bool ta, int64_t M, int64_t K, int64_t k, typename Traits<T>::comp alpha, const T* A,
                           int64_t lda, const T* B, int64_t ldb, typename Traits<T>::comp beta, T* C, int64_t ldc,
                           double shift, const double* theta, cudaStream_t st)
{
    constexpr bool CPLX = Traits<T>::cplx;
    constexpr int BN = HemmCfg<CPLX>::BN;
    if (!hemm_split_enabled() || k % BN == 0 || k % BN > (3 * BN) / 4)
        return hemm_tma_launch_cfg<T, 0>(ta, M, K, k, alpha, A, lda, B, ldb, beta, C, ldc, shift, theta, st);
    // greedy strips: whole default tiles, then a half-width strip, then the narrowest tile for what is left
    int64_t c0 = 0;
    auto strip = [&](auto narrow_tag, int64_t cols) -> int
    {
        constexpr int NARROW = decltype(narrow_tag)::value;
        const int rc = hemm_tma_launch_cfg<T, NARROW>(ta, M, K, cols, alpha, A, lda, B + c0 * ldb, ldb, beta, C + c0 * ldc,
                                                      ldc, shift, theta ? theta + c0 : nullptr, st);
        c0 += cols;
        return rc;
    };
    int rc = 0;
    if (k >= BN)
        rc = strip(std::integral_constant<int, 0>{}, (k / BN) * BN);
    if constexpr (!CPLX)
    {
        if (!rc && k - c0 >= BN / 2)
            rc = strip(std::integral_constant<int, 1>{}, BN / 2);
        if (!rc && k - c0 > BN / 4)
            rc = strip(std::integral_constant<int, 1>{}, k - c0);
        else if (!rc && k - c0 > 0)
            rc = strip(std::integral_constant<int, 2>{}, k - c0);
    }
    else
    {
        if (!rc && k - c0 > 0)
            rc = strip(std::integral_constant<int, 1>{}, k - c0); // one or two 32-column tiles per row block
    }
    return rc;
}

} // namespace cb2
