// chase_b200 host layer — single-GPU backend chase::Impl::ChASEGPU<T>.
//
// Drop-in for the reference's chase::Impl::ChASEGPU<T, Matrix<T,GPU>>
// (Impl/chase_gpu/chase_gpu.hpp:107-1080): same constructor
// (N, nev, nex, H, ldh, V1, ldv, ritzv; caller owns the host buffers), same
// ChaseBase<T> semantics — HEMM swaps the two panels, RR/Resd write host
// arrays, Swap takes absolute column indices, End() copies all nev+nex vectors
// back into the caller's V — but everything below is new:
//   * A stays immutable on the device; Shift(c) is folded into the HEMM
//     epilogue (alpha*(A*B - c*B) + beta*C), so no diagonal kernel touches A;
//   * the O(k^2) Swap storm of calc_degrees/locking is accumulated as a host
//     permutation and applied as ONE gather pass before the next device op;
//   * residuals are A*V - V*diag(theta) in a single fused HEMM + column norms;
//   * every device op is a hand-written sm_100a kernel behind the C ABI of
//     include/chase_b200_kernels.h (no cuBLAS / cuSOLVER / cuRAND);
//   * device panels are padded to a 16-element leading dimension (TMA/128-bit
//     friendly) regardless of the caller's ldh/ldv.
//
// MatrixType = chase::matrix::PseudoHermitianMatrix<T, GPU> selects the
// pseudo-Hermitian (BSE) problem class (reference: the same template argument,
// chase_gpu.hpp:105-107): panels hold 2 (nev+nex) columns laid out as
// [locked+ | active | K-conjugates of active | K-conjugates of locked+], the
// filter runs on H^2 (HEMM_H2), QR orthogonalises against S [locked], RR is the
// S-projected rayleighRitz_v2 and Lanczos uses the S H inner product.
#pragma once
#include "algorithm.hpp"
#include "interface.hpp"
#include "kernel_api.hpp"
#include "tridiag_host.hpp"

#include <cuda_runtime_api.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace chase
{
namespace Impl
{

#define CB2_CHECK(call)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
        {                                                                                                              \
            std::fprintf(stderr, "chase_b200: CUDA failure '%s' at %s:%d\n", cudaGetErrorString(e__), __FILE__,       \
                         __LINE__);                                                                                    \
            std::exit(EXIT_FAILURE);                                                                                   \
        }                                                                                                              \
    } while (0)

#define CB2_KCHECK(call)                                                                                               \
    do                                                                                                                 \
    {                                                                                                                  \
        int rc__ = (call);                                                                                             \
        if (rc__ < 0)                                                                                                  \
        {                                                                                                              \
            std::fprintf(stderr, "chase_b200: kernel launcher failed (%d) at %s:%d\n", rc__, __FILE__, __LINE__);     \
            std::exit(EXIT_FAILURE);                                                                                   \
        }                                                                                                              \
    } while (0)

// Start block of the reference CPU backend, bit for bit: std::mt19937(1337) +
// std::normal_distribution<double>, filled column by column (chase_cpu.hpp:296-309).
template <class T>
inline void fill_start_vectors(std::size_t N, std::size_t ncols, T* V, std::size_t ldv)
{
    std::mt19937 gen(1337.0);
    std::normal_distribution<> d;
    for (std::size_t j = 0; j < ncols; ++j)
        for (std::size_t i = 0; i < N; ++i)
            V[i + j * ldv] = getRandomT<T>([&]() { return d(gen); });
}

template <class T, class MatrixType = chase::matrix::Matrix<T, chase::platform::GPU>>
class ChASEGPU : public ChaseBase<T>
{
    using R = Base<T>;
    using KK = b200::K<T>;
    static constexpr bool kCplx = is_complex_t<T>::value;
    static constexpr bool kPseudo =
        std::is_same<MatrixType, chase::matrix::PseudoHermitianMatrix<T, chase::platform::GPU>>::value;
    static_assert(!kPseudo || kCplx, "pseudo-Hermitian (BSE) problems are complex (reference: c/z only)");

public:
    ChASEGPU(std::size_t N, std::size_t nev, std::size_t nex, T* H, std::size_t ldh, T* V1, std::size_t ldv,
             R* ritzv)
        : N_(N), nev_(nev), nex_(nex), nevex_(nev + nex), nc_(kPseudo ? 2 * (nev + nex) : nev + nex), H_(H), ldh_(ldh),
          V_(V1), ldv_(ldv), ritzv_(ritzv), config_(N, nev, nex)
    {
        if (N == 0 || nevex_ == 0 || nc_ > N)
            throw std::invalid_argument(kPseudo ? "ChASEGPU: need 0 < 2 (nev+nex) <= N"
                                                : "ChASEGPU: need 0 < nev+nex <= N");
        if (kPseudo && N % 2 != 0)
            throw std::invalid_argument("ChASEGPU: a pseudo-Hermitian matrix has even order");
        if (ldh < N || ldv < N)
            throw std::invalid_argument("ChASEGPU: leading dimension smaller than N");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw std::runtime_error("ChASEGPU: no CUDA device available (this backend has no CPU fallback)");
        CB2_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        ld_ = roundup(N_, 16);
        ldg_ = roundup(nc_, 16);
        dH_ = alloc<T>(ld_ * N_);
        dV1_ = alloc<T>(ld_ * nc_);
        dV2_ = alloc<T>(ld_ * nc_);
        dW_ = alloc<T>(ld_ * nc_);
        dG_ = alloc<T>(ldg_ * nc_);
        dZ_ = alloc<T>(ldg_ * nc_);
        if (kPseudo)
        {
            // rayleighRitz_v2 works on three more small matrices (M, R^-1, products)
            dM_ = alloc<T>(ldg_ * nc_);
            dRinv_ = alloc<T>(ldg_ * nc_);
            dT_ = alloc<T>(ldg_ * nc_);
        }
        heev_ws_bytes_ = chase_b200_heev_ws_bytes((int64_t)nc_, kCplx ? 1 : 0);
        heev_ws_ = alloc<unsigned char>(heev_ws_bytes_);
        trsm_ws_bytes_ = chase_b200_trsm_ws_bytes((int64_t)nc_, (int)sizeof(T));
        trsm_ws_ = alloc<unsigned char>(trsm_ws_bytes_);
        splitk_ws_bytes_ = std::max<std::size_t>(std::size_t(64) << 20, 4 * nc_ * nc_ * 16);
        splitk_ws_ = alloc<unsigned char>(splitk_ws_bytes_);
        dTheta_ = alloc<double>(nc_);
        dNorms_ = alloc<double>(nc_);
        dInfo_ = alloc<int>(4);
        dIdx_ = alloc<int>(2 * nc_);
        resid_.assign(nc_, R(0));
        perm_.resize(nc_);
        for (std::size_t i = 0; i < nc_; ++i)
            perm_[i] = (int)i;
        const char* e = std::getenv("CHASE_B200_DEVICE_RNG");
        device_rng_ = e && std::atoi(e) != 0;
        const char* mp = std::getenv("CHASE_B200_MIXED_PRECISION");
        mixed_ = mp && std::atoi(mp) != 0;
        // FP32 storage.  Default: the tcgen05 kind::tf32 kernel with 3x/4x TF32 splitting (csrc/hemm_tf32.cuh): the
        // matrix as stored is the hi operand, one extra FP32 array holds the lo part (2x the matrix memory).
        // CHASE_B200_FP32_PATH=fp64copy: the round-1 route through an FP64 copy on the DMMA kernel (3x memory);
        // =generic (or CHASE_B200_FP32_WIDEN=0): the generic kernel that widens tiles on the fly.
        if (sizeof(R) == 4)
        {
            std::string path = "tf32";
            if (const char* e = std::getenv("CHASE_B200_FP32_PATH"))
                path = e;
            const char* w = std::getenv("CHASE_B200_FP32_WIDEN");
            if (w && std::atoi(w) == 0)
                path = "generic";
            if (path == "tf32")
            {
                dHl_ = alloc<T>(ld_ * N_);
                tf32_scratch_bytes_ = chase_b200_hemm_tf32_scratch_bytes((int64_t)N_, (int64_t)nc_, (int)sizeof(T));
                tf32_scratch_ = alloc<unsigned char>(tf32_scratch_bytes_);
                chase_b200_tf32_register(dH_, dHl_, (int64_t)ld_, (int64_t)N_, (int64_t)N_, kPseudo ? 1 : 0,
                                         tf32_scratch_, tf32_scratch_bytes_);
            }
            else if (path == "fp64copy")
            {
                wide_scratch_bytes_ = 2 * ld_ * nc_ * 2 * sizeof(T);
                dHw_ = alloc<unsigned char>(ld_ * N_ * 2 * sizeof(T));
                wide_scratch_ = alloc<unsigned char>(wide_scratch_bytes_);
                chase_b200_widen_register(dH_, dHw_, (int64_t)ld_, (int64_t)N_, (int64_t)N_, wide_scratch_,
                                          wide_scratch_bytes_);
            }
        }
    }
    // the reference's second constructor: H as a (host-backed) matrix object (chase_gpu.hpp:195-267)
    ChASEGPU(std::size_t N, std::size_t nev, std::size_t nex, MatrixType* H, T* V1, std::size_t ldv, R* ritzv)
        : ChASEGPU(N, nev, nex, check_matrix(H, N)->cpu_data(), H->cpu_ld(), V1, ldv, ritzv)
    {
    }
    ChASEGPU(const ChASEGPU&) = delete;
    ~ChASEGPU() override
    {
        if (dHw_)
            chase_b200_widen_unregister(dH_);
        if (dHl_)
            chase_b200_tf32_unregister(dH_);
        for (void* p : allocs_)
            cudaFree(p);
        if (stream_)
        {
            chase_b200_stream_release(stream_);
            cudaStreamDestroy(stream_);
        }
    }

    // ---- ChaseBase ----------------------------------------------------------
    void Start() override { locked_ = 0; }

    void initVecs(bool random) override
    {
        if (random && device_rng_)
        {
            CB2_KCHECK(KK::rng_normal((int64_t)N_, (int64_t)nc_, dV1_, (int64_t)ld_, 24141ull, stream_));
        }
        else if (random && dV0_ != nullptr)
        {
            // the reference stream is a pure function of (N, #columns, T): generated once, then kept on the device
            CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)nc_, dV0_, (int64_t)ld_, dV1_, (int64_t)ld_, stream_));
        }
        else
        {
            if (random)
            {
                fill_start_vectors<T>(N_, nc_, V_, ldv_);
            }
            CB2_CHECK(cudaMemcpy2DAsync(dV1_, ld_ * sizeof(T), V_, ldv_ * sizeof(T), N_ * sizeof(T), nc_,
                                        cudaMemcpyHostToDevice, stream_));
            if (random)
            {
                dV0_ = alloc<T>(ld_ * nc_);
                CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)nc_, dV1_, (int64_t)ld_, dV0_, (int64_t)ld_, stream_));
            }
        }
        if (random && kPseudo)
        {
            // damp the lower (de-excitation) block of the start vectors: T(0.001), chase_gpu.hpp:518-529
            const std::size_t half = N_ / 2;
            CB2_KCHECK(KK::scale_rows((int64_t)(N_ - half), (int64_t)nc_, dV1_ + half, (int64_t)ld_,
                                      (double)(R)0.001, stream_));
        }
        CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)nc_, dV1_, (int64_t)ld_, dV2_, (int64_t)ld_, stream_));
        // the host matrix is (re-)read at every solve: callers fill or perturb H
        // after construction (examples/4_interface/4_c_serial_chase.c:49-66).
        // keep_device_matrix(true) is the opt-out for callers whose H is unchanged
        // since the previous solve (device-resident benchmarking, sequences on V only).
        if (!(keep_device_matrix_ && matrix_on_device_))
        {
            CB2_CHECK(cudaMemcpy2DAsync(dH_, ld_ * sizeof(T), H_, ldh_ * sizeof(T), N_ * sizeof(T), N_,
                                        cudaMemcpyHostToDevice, stream_));
            matrix_on_device_ = true;
            sp_matrix_valid_ = false;
            if (dHw_)
                CB2_KCHECK(chase_b200_widen_sync(kCplx ? 'c' : 's', dH_, stream_));
            if (dHl_)
                CB2_KCHECK(chase_b200_tf32_sync(kCplx ? 'c' : 's', dH_, stream_));
        }
        reset_perm();
        shift_ = 0.0;
    }

    // A stays immutable: the shift is folded into the HEMM epilogue.  Shift(-c) / Shift(+c, true) also bracket the
    // filter, which is where the reference switches its mixed-precision mode on and off (ENABLE_MIXED_PRECISION,
    // Impl/pchase_gpu/pchase_gpu.hpp:785-817): while the smallest residual of the wanted, unlocked pairs is above
    // 1e-3 the filter of a double-precision problem runs in single precision -- here on the tcgen05 kind::tf32
    // kernel (FP32-accurate 3xTF32) -- and the filtered block is converted back at the unshift.  Opt-in like the
    // reference's compile-time option: CHASE_B200_MIXED_PRECISION=1 or chase_b200_set_mixed_precision_.
    void Shift(T c, bool isunshift = false) override
    {
        shift_ += (double)std::real(c);
        if constexpr (sizeof(R) == 8 && !kPseudo)
        {
            if (!mixed_)
                return;
            if (!isunshift)
            {
                R mn = std::numeric_limits<R>::max();
                for (std::size_t i = locked_; i < nev_; ++i)
                    mn = std::min(mn, resid_[i]);
                if (locked_ < nev_ && mn > R(1e-3))
                    sp_begin();
            }
            else if (sp_active_)
                sp_end();
        }
    }

    void HEMM(std::size_t block, T alpha, T beta, std::size_t offset_left, std::size_t offset_right = 0) override
    {
        flush_perm();
        resid_ready_ = false;
        const std::size_t ncols = (offset_right < block) ? block - offset_right : 0;
        if (ncols > 0 && sp_active_)
        {
            const std::size_t c0 = offset_left + locked_;
            const std::size_t es = sizeof(T) / 2; // bytes of the single-precision element
            auto fn = kCplx ? chase_b200_hemm_tf32_c : chase_b200_hemm_tf32_s;
            const int rc = fn((int64_t)N_, (int64_t)N_, (int64_t)ncols, b200::re_of(alpha), b200::im_of(alpha), dHs_, dHsl_,
                              (int64_t)ld_, dV1s_ + c0 * ld_ * es, (int64_t)ld_, b200::re_of(beta), b200::im_of(beta),
                              dV2s_ + c0 * ld_ * es, (int64_t)ld_, -shift_, nullptr, 0, 3, sp_scratch_, sp_scratch_bytes_,
                              stream_);
            CB2_KCHECK(rc);
            hemm_cols_ += ncols;
            sp_cols_ += ncols;
            std::swap(dV1s_, dV2s_);
        }
        else if (ncols > 0)
        {
            const std::size_t c0 = offset_left + locked_;
            // A_eff = A + shift_ I  ->  alpha (A - (-shift_) I) B + beta C
            CB2_KCHECK(KK::hemm((int64_t)N_, (int64_t)ncols, b200::re_of(alpha), b200::im_of(alpha), dH_, (int64_t)ld_,
                                dV1_ + c0 * ld_, (int64_t)ld_, b200::re_of(beta), b200::im_of(beta), dV2_ + c0 * ld_,
                                (int64_t)ld_, -shift_, nullptr, stream_));
            hemm_cols_ += ncols;
        }
        std::swap(dV1_, dV2_);
    }

    // One filter step in H^2:  V2[cols] <- alpha H (H V1[cols]) + beta V2[cols] + gamma V1[cols]; swap.
    // With gamma = -alpha c (the only form filter_H2 uses, c >= 0) the polynomial factorises,
    //   alpha (H^2 - c I) = alpha (H - sqrt(c) I)(H + sqrt(c) I),
    // so both products run through the filter HEMM with the shift folded into its epilogue and the gamma V1 pass of
    // the reference (batchedAxpyScalar, chase_gpu.hpp:708-713) disappears.  Any other gamma takes the literal route.
    // Columns: [locked_ + offset_left, locked_ + block - offset_right).  The reference always multiplies
    // block - offset_right columns starting at offset_left, i.e. offset_left columns beyond the active half; those
    // are K-conjugate slots that ApplyKconjugate rewrites right after the filter, so they are skipped here.
    void HEMM_H2(std::size_t block, T alpha, T beta, T gamma, std::size_t offset_left,
                 std::size_t offset_right = 0) override
    {
        if (!kPseudo)
            throw std::runtime_error("chase_b200: HEMM_H2 needs MatrixType = PseudoHermitianMatrix");
        flush_perm();
        std::size_t ncols = (offset_right < block) ? block - offset_right : 0;
        ncols = (offset_left < ncols) ? ncols - offset_left : 0;
        if (ncols > 0)
        {
            const std::size_t c0 = offset_left + locked_;
            T* in = dV1_ + c0 * ld_;
            T* out = dV2_ + c0 * ld_;
            T* tmp = dW_ + c0 * ld_;
            const double a = b200::re_of(alpha), g = b200::re_of(gamma);
            const double c = (a != 0.0) ? -g / a : -1.0;
            const bool factorised = b200::im_of(alpha) == 0.0 && b200::im_of(gamma) == 0.0 && c >= 0.0 &&
                                    !std::getenv("CHASE_B200_H2_AXPY");
            if (factorised)
            {
                const double rc = std::sqrt(c);
                CB2_KCHECK(KK::hemm((int64_t)N_, (int64_t)ncols, 1.0, 0.0, dH_, (int64_t)ld_, in, (int64_t)ld_, 0.0, 0.0,
                                    tmp, (int64_t)ld_, -rc, nullptr, stream_));
                CB2_KCHECK(KK::hemm((int64_t)N_, (int64_t)ncols, a, 0.0, dH_, (int64_t)ld_, tmp, (int64_t)ld_,
                                    b200::re_of(beta), b200::im_of(beta), out, (int64_t)ld_, rc, nullptr, stream_));
            }
            else
            {
                CB2_KCHECK(KK::hemm((int64_t)N_, (int64_t)ncols, 1.0, 0.0, dH_, (int64_t)ld_, in, (int64_t)ld_, 0.0, 0.0,
                                    tmp, (int64_t)ld_, 0.0, nullptr, stream_));
                CB2_KCHECK(KK::hemm((int64_t)N_, (int64_t)ncols, a, b200::im_of(alpha), dH_, (int64_t)ld_, tmp,
                                    (int64_t)ld_, b200::re_of(beta), b200::im_of(beta), out, (int64_t)ld_, 0.0, nullptr,
                                    stream_));
                if (ones_ == nullptr)
                {
                    ones_ = alloc<double>(nc_);
                    std::vector<double> one(nc_, 1.0);
                    CB2_CHECK(cudaMemcpy(ones_, one.data(), nc_ * sizeof(double), cudaMemcpyHostToDevice));
                }
                CB2_KCHECK(KK::axpy_cols((int64_t)N_, (int64_t)ncols, ones_, g, b200::im_of(gamma), in, (int64_t)ld_, out,
                                         (int64_t)ld_, stream_));
            }
            hemm_cols_ += 2 * ncols;
        }
        std::swap(dV1_, dV2_);
    }

    // V1[:, 2 nevex - locked - block ...) <- K-conjugates of V1[:, locked ... locked + block)
    void ApplyKconjugate(std::size_t block) override
    {
        if (!kPseudo)
            return; // reference: no-op for Hermitian problems (chase_gpu.hpp:721-723)
        flush_perm();
        if (block == 0)
            return;
        const std::size_t col_second = nc_ - locked_ - block;
        CB2_KCHECK(KK::kconj((int64_t)N_, (int64_t)block, dV1_ + locked_ * ld_, (int64_t)ld_, dV1_ + col_second * ld_,
                             (int64_t)ld_, stream_));
    }

    void QR(std::size_t /*fixednev*/, R cond) override
    {
        flush_perm();
        resid_ready_ = false;
        // keep the locked vectors: CholQR runs on all columns
        CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)locked_, dV1_, (int64_t)ld_, dV2_, (int64_t)ld_, stream_));
        if (kPseudo)
        {
            // layout [L+ | active | L-]: park L- in V2 too, then orthogonalise [S L+ | S L- | active]: the right
            // eigenvectors are S-orthogonal, not orthogonal (chase_gpu.hpp:752-782)
            const std::size_t act = nc_ - 2 * locked_;
            CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)locked_, dV1_ + (nc_ - locked_) * ld_, (int64_t)ld_,
                                 dV2_ + (nc_ - locked_) * ld_, (int64_t)ld_, stream_));
            CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)locked_, dV1_, (int64_t)ld_, dW_, (int64_t)ld_, stream_));
            CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)locked_, dV1_ + (nc_ - locked_) * ld_, (int64_t)ld_,
                                 dW_ + locked_ * ld_, (int64_t)ld_, stream_));
            CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)act, dV1_ + locked_ * ld_, (int64_t)ld_,
                                 dW_ + 2 * locked_ * ld_, (int64_t)ld_, stream_));
            std::swap(dV1_, dW_);
            const std::size_t half = N_ / 2;
            CB2_KCHECK(KK::scale_rows((int64_t)(N_ - half), (int64_t)(2 * locked_), dV1_ + half, (int64_t)ld_, -1.0,
                                      stream_));
        }

        int disable = config_.DoCholQR() ? 0 : 1;
        if (const char* s = std::getenv("CHASE_DISABLE_CHOLQR"))
            disable = std::atoi(s);
        R thr_upper = (sizeof(R) == 8) ? R(1e8) : R(1e4);
        R thr_lower = (sizeof(R) == 8) ? R(2e1) : R(1e1);
        if (const char* s = std::getenv("CHASE_CHOLQR1_THLD"))
            thr_lower = (R)std::atof(s);

        int info = 1;
        if (disable == 1 && cond != R(1.0))
        {
            // qr == 'H' / CHASE_DISABLE_CHOLQR=1: Householder QR (chase_gpu.hpp:822-846)
            householder();
            info = 0;
            last_qr_ = "householder";
        }
        else if (cond > thr_upper)
        {
            info = shifted_cholqr2(1.0);
            last_qr_ = "shifted2";
        }
        else if (cond < thr_lower)
        {
            info = chol_round(false, 0.0);
            last_qr_ = "chol1";
        }
        else
        {
            info = chol_round(false, 0.0);
            if (info == 0)
                info = chol_round(false, 0.0);
            last_qr_ = "chol2";
        }
        if (info != 0)
        {
            // CholeskyQR broke down (potrf info != 0): Householder QR, as the reference (chase_gpu.hpp:889-919).
            // A failed round may already have replaced V1 by a partly orthogonalised block spanning the same space.
            householder();
            last_qr_ += "+householder";
        }
        qr_log_.push_back(last_qr_);
        if (kPseudo)
        {
            // back to [L+ | active | L-] with the untouched locked columns
            const std::size_t act = nc_ - 2 * locked_;
            CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)act, dV1_ + 2 * locked_ * ld_, (int64_t)ld_,
                                 dW_ + locked_ * ld_, (int64_t)ld_, stream_));
            std::swap(dV1_, dW_);
            CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)locked_, dV2_ + (nc_ - locked_) * ld_, (int64_t)ld_,
                                 dV1_ + (nc_ - locked_) * ld_, (int64_t)ld_, stream_));
            // with nothing locked the reference leaves the orthonormal block in BOTH panels (chase_gpu.hpp:922-938);
            // LanczosDos copies columns of the second panel back, so the DoS start vectors depend on it
            if (locked_ == 0)
                CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)nc_, dV1_, (int64_t)ld_, dV2_, (int64_t)ld_, stream_));
        }
        CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)locked_, dV2_, (int64_t)ld_, dV1_, (int64_t)ld_, stream_));
    }

    void RR(R* ritzv, std::size_t block) override
    {
        flush_perm();
        resid_ready_ = false;
        if (block == 0)
            return;
        if (kPseudo)
        {
            rr_pseudo(ritzv, block);
            return;
        }
        T* Q = dV1_ + locked_ * ld_;
        T* W = dV2_ + locked_ * ld_;
        // W = A Q   (the reference forms A^H Q; A is Hermitian).  FP32 types: 4 TF32 partial products (full FP32
        // operand precision) for the projected matrix and the residual block, 3 in the filter
        chase_b200_tf32_set_terms(4);
        CB2_KCHECK(KK::hemm((int64_t)N_, (int64_t)block, 1.0, 0.0, dH_, (int64_t)ld_, Q, (int64_t)ld_, 0.0, 0.0, W,
                            (int64_t)ld_, 0.0, nullptr, stream_));
        chase_b200_tf32_set_terms(3);
        // G = W^H Q
        CB2_KCHECK(KK::gemm(1, 0, (int64_t)block, (int64_t)block, (int64_t)N_, 1.0, 0.0, W, (int64_t)ld_, Q,
                            (int64_t)ld_, 0.0, 0.0, dG_, (int64_t)ldg_, 0, splitk_ws_, splitk_ws_bytes_, stream_));
        std::vector<double> w(block);
        int sweeps = 0;
        int rc = KK::heev((int64_t)block, dG_, (int64_t)ldg_, dZ_, (int64_t)ldg_, w.data(), heev_ws_, heev_ws_bytes_,
                          &sweeps, stream_);
        if (rc < 0)
            throw std::runtime_error("chase_b200: Hermitian eigensolver failed in RR (rc=" + std::to_string(rc) + ")");
        heev_sweeps_ += sweeps;
        if (rc > 0) // sweep limit reached: the decomposition is still usable, the residual check guards accuracy
            std::fprintf(stderr, "chase_b200: warning: Jacobi eigensolver stopped at its sweep limit (%d sweeps)\n", sweeps);
        for (std::size_t i = 0; i < block; ++i)
            ritzv[i] = (R)w[i];
        // residual block while A Q is at hand: R = (A Q) Z - (Q Z) Theta  (the reference runs a second A V,
        // cuda/residuals.hpp:92-110; N k^2 instead of N^2 k flops here)
        T* Rm = dW_ + locked_ * ld_;
        CB2_KCHECK(KK::gemm(0, 0, (int64_t)N_, (int64_t)block, (int64_t)block, 1.0, 0.0, W, (int64_t)ld_, dZ_,
                            (int64_t)ldg_, 0.0, 0.0, Rm, (int64_t)ld_, 0, nullptr, 0, stream_));
        // V2 = Q Z ; swap
        CB2_KCHECK(KK::gemm(0, 0, (int64_t)N_, (int64_t)block, (int64_t)block, 1.0, 0.0, Q, (int64_t)ld_, dZ_,
                            (int64_t)ldg_, 0.0, 0.0, W, (int64_t)ld_, 0, nullptr, 0, stream_));
        std::swap(dV1_, dV2_);
        for (std::size_t i = 0; i < block; ++i)
            w[i] = (double)ritzv[i]; // rounded to Base<T> like the values the driver passes to Resd
        CB2_CHECK(cudaMemcpyAsync(dTheta_, w.data(), block * sizeof(double), cudaMemcpyHostToDevice, stream_));
        CB2_KCHECK(KK::axpy_cols((int64_t)N_, (int64_t)block, dTheta_, -1.0, 0.0, dV1_ + locked_ * ld_, (int64_t)ld_,
                                 Rm, (int64_t)ld_, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_)); // w is a host temporary
        resid_ready_ = true;
        resid_block_ = block;
    }

    void Sort(R*, R*, R*) override {}

    void Resd(R* ritzv, R* resd, std::size_t /*fixednev*/) override
    {
        flush_perm();
        const std::size_t k = nevex_ - locked_;
        if (k == 0)
            return;
        std::vector<double> th(k);
        T* W = dV2_ + locked_ * ld_;
        if (!kPseudo && resid_ready_ && resid_block_ == k && !std::getenv("CHASE_B200_RESID_HEMM"))
        {
            W = dW_ + locked_ * ld_; // prepared by RR
        }
        else
        {
            for (std::size_t i = 0; i < k; ++i)
                th[i] = (double)ritzv[i];
            CB2_CHECK(cudaMemcpyAsync(dTheta_, th.data(), k * sizeof(double), cudaMemcpyHostToDevice, stream_));
            T* V = dV1_ + locked_ * ld_;
            // W = A V - V diag(theta), shift folded into the HEMM epilogue
            chase_b200_tf32_set_terms(4);
            CB2_KCHECK(KK::hemm((int64_t)N_, (int64_t)k, 1.0, 0.0, dH_, (int64_t)ld_, V, (int64_t)ld_, 0.0, 0.0, W,
                                (int64_t)ld_, 0.0, dTheta_, stream_));
            chase_b200_tf32_set_terms(3);
        }
        resid_ready_ = false;
        CB2_KCHECK(KK::colnorms((int64_t)N_, (int64_t)k, W, (int64_t)ld_, dNorms_, 1, stream_));
        std::vector<double> nr(k);
        CB2_CHECK(cudaMemcpyAsync(nr.data(), dNorms_, k * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        for (std::size_t i = 0; i < k; ++i)
            resd[i] = (R)nr[i];
    }

    void Lanczos(std::size_t M, R* upperb) override
    {
        lanczosIter_ = M;
        numLanczos_ = 1;
        std::vector<R> theta(M), tau(M), rv(M * M);
        if (kPseudo)
            run_lanczos_pseudo(M, 1, upperb, theta.data(), tau.data(), rv.data(), false);
        else
            run_lanczos(M, 1, upperb, theta.data(), tau.data(), rv.data(), false);
    }

    void Lanczos(std::size_t M, std::size_t numvec, R* upperb, R* ritzv, R* Tau, R* ritzV) override
    {
        lanczosIter_ = M;
        numLanczos_ = numvec;
        if (kPseudo)
            run_lanczos_pseudo(M, numvec, upperb, ritzv, Tau, ritzV, true);
        else
            run_lanczos(M, numvec, upperb, ritzv, Tau, ritzV, true);
    }

    void LanczosDos(std::size_t idx, std::size_t m, T* ritzVc) override
    {
        flush_perm();
        CB2_CHECK(cudaMemcpy2DAsync(dZ_, ldg_ * sizeof(T), ritzVc, m * sizeof(T), m * sizeof(T), idx,
                                    cudaMemcpyHostToDevice, stream_));
        CB2_KCHECK(KK::gemm(0, 0, (int64_t)N_, (int64_t)idx, (int64_t)m, 1.0, 0.0, dV1_, (int64_t)ld_, dZ_,
                            (int64_t)ldg_, 0.0, 0.0, dV2_, (int64_t)ld_, 0, nullptr, 0, stream_));
        CB2_KCHECK(KK::lacpy((int64_t)N_, (int64_t)m, dV2_, (int64_t)ld_, dV1_, (int64_t)ld_, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_)); // ritzVc is caller-owned
    }

    void Swap(std::size_t i, std::size_t j) override
    {
        std::swap(perm_[i], perm_[j]);
        perm_dirty_ = true;
        swaps_++;
    }

    void Lock(std::size_t new_converged) override { locked_ += new_converged; }

    bool checkSymmetryEasy() override
    {
        CB2_CHECK(cudaMemcpy2DAsync(dH_, ld_ * sizeof(T), H_, ldh_ * sizeof(T), N_ * sizeof(T), N_,
                                    cudaMemcpyHostToDevice, stream_));
        unsigned long long* bad = reinterpret_cast<unsigned long long*>(dNorms_);
        CB2_CHECK(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), stream_));
        const double tol = (sizeof(R) == 8) ? 1e-10 : 1e-5;
        CB2_KCHECK(KK::herm_check((int64_t)N_, dH_, (int64_t)ld_, tol, bad, stream_));
        unsigned long long h = 0;
        CB2_CHECK(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        is_sym_ = (h == 0);
        return is_sym_;
    }
    bool isSym() override { return !kPseudo; }
    // S H Hermitian?  (reference: flip the lower half, checkSymmetryEasy, flip back; chase_gpu.hpp:482-500)
    bool checkPseudoHermicityEasy() override
    {
        if (N_ % 2 != 0)
            return false;
        CB2_CHECK(cudaMemcpy2DAsync(dH_, ld_ * sizeof(T), H_, ldh_ * sizeof(T), N_ * sizeof(T), N_,
                                    cudaMemcpyHostToDevice, stream_));
        const std::size_t half = N_ / 2;
        CB2_KCHECK(KK::scale_rows((int64_t)half, (int64_t)N_, dH_ + half, (int64_t)ld_, -1.0, stream_));
        unsigned long long* bad = reinterpret_cast<unsigned long long*>(dNorms_);
        CB2_CHECK(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), stream_));
        const double tol = (sizeof(R) == 8) ? 1e-10 : 1e-5;
        CB2_KCHECK(KK::herm_check((int64_t)N_, dH_, (int64_t)ld_, tol, bad, stream_));
        CB2_KCHECK(KK::scale_rows((int64_t)half, (int64_t)N_, dH_ + half, (int64_t)ld_, -1.0, stream_));
        unsigned long long h = 0;
        CB2_CHECK(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        matrix_on_device_ = true;
        return h == 0;
    }
    bool isPseudoHerm() override { return kPseudo; }
    void symOrHermMatrix(char uplo) override
    {
        // acts on the caller's host matrix (re-uploaded at the next initVecs), like the reference
        for (std::size_t j = 0; j < N_; ++j)
            for (std::size_t i = j + 1; i < N_; ++i)
            {
                if (uplo == 'U')
                    H_[i + j * ldh_] = conjugate(H_[j + i * ldh_]);
                else
                    H_[j + i * ldh_] = conjugate(H_[i + j * ldh_]);
            }
    }

    void End() override
    {
        flush_perm();
        CB2_CHECK(cudaMemcpy2DAsync(V_, ldv_ * sizeof(T), dV1_, ld_ * sizeof(T), N_ * sizeof(T), nc_,
                                    cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
    }

    std::size_t GetN() const override { return N_; }
    std::size_t GetNev() override { return nev_; }
    std::size_t GetNex() override { return nex_; }
    std::size_t GetLanczosIter() override { return lanczosIter_; }
    std::size_t GetNumLanczos() override { return numLanczos_; }
    std::size_t GetRitzvBlockSize() const override { return nc_; }
    R* GetRitzv() override { return ritzv_; }
    R* GetResid() override { return resid_.data(); }
    ChaseConfig<T>& GetConfig() override { return config_; }
    int get_nprocs() override { return 1; }
    int get_rank() override { return 0; }
    void Output(LogLevel, std::string s, const char* = "algorithm") override
    {
        if (std::getenv("CHASE_B200_VERBOSE"))
            std::cout << s;
    }

    // Raw column-major binary matrix files, the reference's format (Matrix::readFromBinaryFile / saveToBinaryFile,
    // linalg/matrix/matrix.hpp:276-352; ChASEGPU::loadProblemFromFile, chase_gpu.hpp:457-461): N*N elements, no
    // header.  The file lands in the caller's host H, which the next solve uploads.
    void loadProblemFromFile(const std::string& filename)
    {
        std::ifstream f(filename, std::ios::binary);
        if (!f.is_open())
            throw std::runtime_error("chase_b200: cannot open " + filename + " for reading");
        f.seekg(0, std::ios::end);
        if ((std::size_t)f.tellg() < N_ * N_ * sizeof(T))
            throw std::runtime_error("chase_b200: " + filename + " is smaller than the N x N matrix");
        f.seekg(0, std::ios::beg);
        for (std::size_t j = 0; j < N_; ++j)
            f.read(reinterpret_cast<char*>(H_ + j * ldh_), (std::streamsize)(N_ * sizeof(T)));
        matrix_on_device_ = false;
    }
    void saveProblemToFile(const std::string& filename)
    {
        std::ofstream f(filename, std::ios::binary);
        if (!f.is_open())
            throw std::runtime_error("chase_b200: cannot open " + filename + " for writing");
        for (std::size_t j = 0; j < N_; ++j)
            f.write(reinterpret_cast<const char*>(H_ + j * ldh_), (std::streamsize)(N_ * sizeof(T)));
    }

    // ---- extras (not part of ChaseBase) ------------------------------------
    const std::vector<std::string>& qr_log() const { return qr_log_; }
    void clear_logs()
    {
        qr_log_.clear();
        heev_sweeps_ = 0;
        hemm_cols_ = 0;
        swaps_ = 0;
        gathers_ = 0;
        sp_cols_ = 0;
        sp_filters_ = 0;
    }
    void keep_device_matrix(bool f) { keep_device_matrix_ = f; }
    void use_device_rng(bool f) { device_rng_ = f; }
    void use_mixed_precision(bool f) { mixed_ = f; }
    std::size_t sp_filter_cols() const { return sp_cols_; } // columns filtered in single precision
    std::size_t sp_filters() const { return sp_filters_; }
    std::size_t heev_sweeps() const { return heev_sweeps_; }
    std::size_t hemm_cols() const { return hemm_cols_; } // columns actually multiplied by the filter
    std::size_t gather_passes() const { return gathers_; }
    cudaStream_t stream() const { return stream_; }
    T* device_H() { return dH_; }
    T* device_V1() { return dV1_; }
    std::size_t device_ld() const { return ld_; }
    void set_host_buffers(T* H, std::size_t ldh, T* V, std::size_t ldv, R* ritzv)
    {
        H_ = H;
        ldh_ = ldh;
        V_ = V;
        ldv_ = ldv;
        ritzv_ = ritzv;
    }

private:
    static MatrixType* check_matrix(MatrixType* H, std::size_t N)
    {
        if (H == nullptr || H->cpu_data() == nullptr || H->rows() != N || H->cols() != N)
            throw std::invalid_argument("ChASEGPU: H must be an N x N matrix with a host buffer");
        return H;
    }
    static std::size_t roundup(std::size_t a, std::size_t b) { return (a + b - 1) / b * b; }
    template <class U>
    U* alloc(std::size_t n)
    {
        void* p = nullptr;
        CB2_CHECK(cudaMalloc(&p, std::max<std::size_t>(n, 1) * sizeof(U)));
        CB2_CHECK(cudaMemset(p, 0, std::max<std::size_t>(n, 1) * sizeof(U)));
        allocs_.push_back(p);
        return static_cast<U*>(p);
    }
    void reset_perm()
    {
        for (std::size_t i = 0; i < nc_; ++i)
            perm_[i] = (int)i;
        perm_dirty_ = false;
    }
    // apply all pending Swap()s: V1[:, j] <- V1_old[:, perm_[j]]
    void flush_perm()
    {
        if (!perm_dirty_)
            return;
        std::vector<int> idx;
        std::vector<int> src, dst;
        for (std::size_t j = 0; j < nc_; ++j)
            if (perm_[j] != (int)j)
            {
                src.push_back(perm_[j]);
                dst.push_back((int)j);
            }
        const int cnt = (int)dst.size();
        if (cnt > 0)
        {
            idx = src;
            idx.insert(idx.end(), dst.begin(), dst.end());
            CB2_CHECK(cudaMemcpyAsync(dIdx_, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
            CB2_KCHECK(KK::gather_cols((int64_t)N_, cnt, dIdx_, dIdx_ + cnt, dV1_, (int64_t)ld_, dW_, (int64_t)ld_,
                                       stream_));
            CB2_KCHECK(KK::gather_cols((int64_t)N_, cnt, dIdx_ + cnt, dIdx_ + cnt, dW_, (int64_t)ld_, dV1_,
                                       (int64_t)ld_, stream_));
            CB2_CHECK(cudaStreamSynchronize(stream_)); // idx is a host temporary
            gathers_++;
            resid_ready_ = false; // dW_ was the gather scratch
        }
        reset_perm();
    }

    // ---- mixed-precision filter (double-precision problems only) -----------------------------------------------
    void sp_begin()
    {
        const std::size_t es = sizeof(T) / 2;
        const char wide = kCplx ? 'z' : 'd', narrow = kCplx ? 'c' : 's';
        if (dHs_ == nullptr)
        {
            dHs_ = alloc<unsigned char>(ld_ * N_ * es);
            dHsl_ = alloc<unsigned char>(ld_ * N_ * es);
            dV1s_ = alloc<unsigned char>(ld_ * nc_ * es);
            dV2s_ = alloc<unsigned char>(ld_ * nc_ * es);
            sp_scratch_bytes_ = chase_b200_hemm_tf32_scratch_bytes((int64_t)N_, (int64_t)nc_, (int)es);
            sp_scratch_ = alloc<unsigned char>(sp_scratch_bytes_);
        }
        if (!sp_matrix_valid_)
        {
            // single-precision copy of A + the lo part of its TF32 split, once per upload of the matrix
            CB2_KCHECK(chase_b200_convert(wide, narrow, (int64_t)N_, (int64_t)N_, dH_, (int64_t)ld_, dHs_, (int64_t)ld_,
                                          stream_));
            CB2_KCHECK(chase_b200_tf32_register(dHs_, dHsl_, (int64_t)ld_, (int64_t)N_, (int64_t)N_, 0, nullptr, 0));
            CB2_KCHECK(chase_b200_tf32_sync(narrow, dHs_, stream_));
            chase_b200_tf32_unregister(dHs_);
            sp_matrix_valid_ = true;
        }
        flush_perm();
        const std::size_t act = nevex_ - locked_;
        CB2_KCHECK(chase_b200_convert(wide, narrow, (int64_t)N_, (int64_t)act, dV1_ + locked_ * ld_, (int64_t)ld_,
                                      dV1s_ + locked_ * ld_ * es, (int64_t)ld_, stream_));
        sp_active_ = true;
        sp_filters_++;
    }
    // the filtered block (every degree is even, so it sits in the primary panel) goes back to double precision
    void sp_end()
    {
        const std::size_t es = sizeof(T) / 2;
        const std::size_t act = nevex_ - locked_;
        CB2_KCHECK(chase_b200_convert(kCplx ? 'c' : 's', kCplx ? 'z' : 'd', (int64_t)N_, (int64_t)act,
                                      dV1s_ + locked_ * ld_ * es, (int64_t)ld_, dV1_ + locked_ * ld_, (int64_t)ld_,
                                      stream_));
        sp_active_ = false;
    }

    // one CholQR round on all nev+nex columns; shift_boost > 0 adds the shifted-CholQR shift
    int chol_round(bool shifted, double shift_boost)
    {
        const int64_t n = (int64_t)nc_;
        CB2_KCHECK(KK::gemm(1, 0, n, n, (int64_t)N_, 1.0, 0.0, dV1_, (int64_t)ld_, dV1_, (int64_t)ld_, 0.0, 0.0, dG_,
                            (int64_t)ldg_, 1, splitk_ws_, splitk_ws_bytes_, stream_));
        if (shifted)
        {
            // reference CPU formula (cpu/cholqr1.hpp:153-166): sqrt(rows) * eps (double) / 10 * eps (float)
            const double scale = (sizeof(R) == 8) ? std::sqrt((double)N_) * 2.220446049250313e-16
                                                  : 10.0 * 1.1920928955078125e-07;
            CB2_KCHECK(KK::shift_abstrace(n, dG_, (int64_t)ldg_, scale * shift_boost, nullptr, stream_));
        }
        CB2_CHECK(cudaMemsetAsync(dInfo_, 0, sizeof(int), stream_));
        CB2_KCHECK(KK::potrf(n, dG_, (int64_t)ldg_, dInfo_, stream_));
        int info = 0;
        CB2_CHECK(cudaMemcpyAsync(&info, dInfo_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        if (info != 0)
            return info;
        CB2_KCHECK(KK::trsm((int64_t)N_, n, dG_, (int64_t)ldg_, dV1_, (int64_t)ld_, dW_, (int64_t)ld_, trsm_ws_,
                            trsm_ws_bytes_, stream_));
        std::swap(dV1_, dW_);
        return 0;
    }
    // V1 <- orthonormal factor of V1 (all nc_ columns) by Householder reflections
    void householder()
    {
        const std::size_t need = chase_b200_hhqr_ws_bytes((int64_t)N_, (int64_t)nc_, (int)sizeof(T));
        if (hh_ws_bytes_ < need)
        {
            hh_ws_ = alloc<unsigned char>(need);
            hh_ws_bytes_ = need;
        }
        CB2_KCHECK(KK::hhqr((int64_t)N_, (int64_t)nc_, dV1_, (int64_t)ld_, dW_, (int64_t)ld_, hh_ws_, hh_ws_bytes_,
                            stream_));
        std::swap(dV1_, dW_);
    }
    int shifted_cholqr2(double boost)
    {
        int info = chol_round(true, boost);
        if (info)
            return info;
        info = chol_round(false, 0.0);
        if (info)
            return info;
        return chol_round(false, 0.0);
    }

    // Rayleigh-Ritz for the pseudo-Hermitian problem on Q = V1[:, locked_ ... locked_ + 2 block) (orthonormal):
    // reference rayleighRitz_v2 (linalg/internal/cuda/rayleighRitz.hpp:510-784, CPU statement
    // cpu/rayleighRitz.hpp:284-392).  A = Q^H S H Q = R^H R;  M = -R^-H (I - 2 Q2^H Q2) R^-1;  M z = w z (ascending);
    // lambda = 1 / (-w): positive values first, ascending;  X = R^-1 Z[:, :block], columns normalised;
    // V2[:, locked_ ... locked_ + block) = Q X;  swap.  The triangular solves of the reference become products with
    // the explicit n x n inverse R^-1 (one TRSM on the identity), so everything large runs on the DMMA GEMM kernels.
    void rr_pseudo(R* ritzv, std::size_t block)
    {
        const int64_t n = (int64_t)(2 * block), N = (int64_t)N_, half = (int64_t)(N_ / 2);
        const int64_t ld = (int64_t)ld_, ldg = (int64_t)ldg_;
        T* Q = dV1_ + locked_ * ld_;
        T* W = dV2_ + locked_ * ld_;
        chase_b200_tf32_set_terms(4);
        CB2_KCHECK(KK::hemm(N, n, 1.0, 0.0, dH_, ld, Q, ld, 0.0, 0.0, W, ld, 0.0, nullptr, stream_));
        chase_b200_tf32_set_terms(3);
        CB2_KCHECK(KK::scale_rows(N - half, n, W + half, ld, -1.0, stream_));
        CB2_KCHECK(KK::gemm(1, 0, n, n, N, 1.0, 0.0, Q, ld, W, ld, 0.0, 0.0, dG_, ldg, 0, splitk_ws_, splitk_ws_bytes_,
                            stream_));
        CB2_CHECK(cudaMemsetAsync(dInfo_, 0, sizeof(int), stream_));
        CB2_KCHECK(KK::potrf(n, dG_, ldg, dInfo_, stream_));
        int info = 0;
        CB2_CHECK(cudaMemcpyAsync(&info, dInfo_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        if (info != 0)
            throw std::runtime_error("chase_b200: Q^H S H Q is not positive definite in the pseudo-Hermitian RR "
                                     "(potrf info=" + std::to_string(info) + "): S H must be positive definite");
        // R^-1
        CB2_CHECK(cudaMemsetAsync(dT_, 0, ldg_ * (std::size_t)n * sizeof(T), stream_));
        CB2_KCHECK(KK::shift_diag(n, dT_, ldg, 1.0, stream_));
        CB2_KCHECK(KK::trsm(n, n, dG_, ldg, dT_, ldg, dRinv_, ldg, trsm_ws_, trsm_ws_bytes_, stream_));
        // M0 = I - 2 Q2^H Q2  (= Q^H S Q for orthonormal Q)
        CB2_CHECK(cudaMemsetAsync(dM_, 0, ldg_ * (std::size_t)n * sizeof(T), stream_));
        CB2_KCHECK(KK::shift_diag(n, dM_, ldg, 1.0, stream_));
        CB2_KCHECK(KK::gemm(1, 0, n, n, N - half, -2.0, 0.0, Q + half, ld, Q + half, ld, 1.0, 0.0, dM_, ldg, 0,
                            splitk_ws_, splitk_ws_bytes_, stream_));
        // M = -R^-H M0 R^-1
        CB2_KCHECK(KK::gemm(0, 0, n, n, n, 1.0, 0.0, dM_, ldg, dRinv_, ldg, 0.0, 0.0, dT_, ldg, 0, nullptr, 0, stream_));
        CB2_KCHECK(KK::gemm(1, 0, n, n, n, -1.0, 0.0, dRinv_, ldg, dT_, ldg, 0.0, 0.0, dM_, ldg, 0, nullptr, 0,
                            stream_));
        std::vector<double> w((std::size_t)n);
        int sweeps = 0;
        const int rc = KK::heev(n, dM_, ldg, dZ_, ldg, w.data(), heev_ws_, heev_ws_bytes_, &sweeps, stream_);
        if (rc < 0)
            throw std::runtime_error("chase_b200: Hermitian eigensolver failed in the pseudo-Hermitian RR (rc=" +
                                     std::to_string(rc) + ")");
        heev_sweeps_ += sweeps;
        if (rc > 0) // sweep limit reached: the decomposition is still usable, the residual check guards accuracy
            std::fprintf(stderr, "chase_b200: warning: Jacobi eigensolver stopped at its sweep limit (%d sweeps)\n", sweeps);
        for (int64_t i = 0; i < n; ++i)
            ritzv[i] = R(1.0) / (R)(-w[(std::size_t)i]);
        CB2_KCHECK(KK::gemm(0, 0, n, (int64_t)block, n, 1.0, 0.0, dRinv_, ldg, dZ_, ldg, 0.0, 0.0, dT_, ldg, 0, nullptr,
                            0, stream_));
        CB2_KCHECK(KK::normalize_cols(n, (int64_t)block, dT_, ldg, stream_));
        CB2_KCHECK(KK::gemm(0, 0, N, (int64_t)block, n, 1.0, 0.0, Q, ld, dT_, ldg, 0.0, 0.0, W, ld, 0, nullptr, 0,
                            stream_));
        std::swap(dV1_, dV2_);
    }

    // Y <- H X for the (non-Hermitian) pseudo-Hermitian H through the HBM-bound A^H product: H = S H^H S
    void pseudo_matvec(T* X, T* Y, int nv)
    {
        const int64_t N = (int64_t)N_, half = (int64_t)(N_ / 2), ld = (int64_t)ld_;
        CB2_KCHECK(KK::scale_rows(N - half, nv, X + half, ld, -1.0, stream_));
        CB2_KCHECK(KK::gemv_conjt(N, N, dH_, ld, X, ld, nv, Y, ld, stream_));
        CB2_KCHECK(KK::scale_rows(N - half, nv, X + half, ld, -1.0, stream_));
        CB2_KCHECK(KK::scale_rows(N - half, nv, Y + half, ld, -1.0, stream_));
    }

    // Lanczos in the S H inner product (reference cuda/lanczos.hpp:547-785, CPU statement cpu/lanczos.hpp:332-516):
    // tridiagonal (d, e) per start vector, Ritz values / weights from the on-device tridiagonal eigensolver,
    // *upperb = largest Ritz value of the first run.
    void run_lanczos_pseudo(std::size_t M, std::size_t numvec, R* upperb, R* Theta, R* Tau, R* ritzV, bool multi)
    {
        flush_perm();
        const int nv = (int)numvec, m = (int)M;
        ensure_lanczos_buffers(M, numvec);
        T* v0 = lan_v_;
        T* v1 = lan_v_ + ld_ * numvec;
        T* v2 = lan_v_ + 2 * ld_ * numvec;
        const int64_t N = (int64_t)N_, ld = (int64_t)ld_;
        CB2_CHECK(cudaMemsetAsync(lan_d_, 0, M * numvec * sizeof(double), stream_));
        CB2_CHECK(cudaMemsetAsync(lan_e_, 0, M * numvec * sizeof(double), stream_));
        CB2_KCHECK(KK::lacpy(N, nv, dV1_, ld, v1, ld, stream_));
        pseudo_matvec(v1, v2, nv);
        CB2_KCHECK(KK::lanczos_pseudo_norm(N, nv, -1, m, v1, v2, ld, lan_e_, lan_rb_, stream_));
        for (int k = 0; k < m; ++k)
        {
            if (multi)
                CB2_KCHECK(KK::lacpy(N, 1, v1 + (std::size_t)(nv - 1) * ld_, ld, dV1_ + (std::size_t)k * ld_, ld, stream_));
            CB2_KCHECK(KK::lanczos_pseudo_step(N, nv, k, m, v0, v1, v2, ld, lan_d_, lan_rb_, stream_));
            if (k == m - 1)
                break;
            T* t = v0;
            v0 = v1;
            v1 = v2;
            v2 = t;
            pseudo_matvec(v1, v2, nv);
            CB2_KCHECK(KK::lanczos_pseudo_norm(N, nv, k, m, v1, v2, ld, lan_e_, lan_rb_, stream_));
        }
        if (multi)
            CB2_KCHECK(KK::lacpy(N, nv, v1, ld, dV1_, ld, stream_));
        std::vector<double> w, Z;
        lanczos_tridiag_solve(M, numvec, w, Z);
        for (std::size_t i = 0; i < numvec; ++i)
            for (std::size_t k = 0; k < M; ++k)
            {
                Theta[k + M * i] = (R)w[i * M + k];
                const R z0 = (R)Z[i * M * M + 0 + k * M];
                Tau[k + i * M] = std::abs(z0) * std::abs(z0);
            }
        for (std::size_t q = 0; q < M * M; ++q)
            ritzV[q] = (R)Z[(numvec - 1) * M * M + q];
        *upperb = Theta[M - 1];
    }

    // Ritz pairs of the numvec Lanczos tridiagonals (d, e on the device): w[i M + k] ascending, Z[i M^2 + r + k M].
    // Up to 48 steps on the device (one CTA per matrix); beyond that on the host like the reference (?stemr,
    // cuda/lanczos.hpp:270-299).  Synchronises the stream.
    void lanczos_tridiag_solve(std::size_t M, std::size_t numvec, std::vector<double>& w, std::vector<double>& Z)
    {
        w.resize(M * numvec);
        Z.resize(M * M * numvec);
        if (M <= 48)
        {
            CB2_KCHECK(chase_b200_tridiag_eig((int)M, (int)numvec, lan_d_, lan_e_, (int)M, lan_w_, lan_Z_, stream_));
            CB2_CHECK(cudaMemcpyAsync(w.data(), lan_w_, w.size() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
            CB2_CHECK(cudaMemcpyAsync(Z.data(), lan_Z_, Z.size() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
            CB2_CHECK(cudaStreamSynchronize(stream_));
            return;
        }
        std::vector<double> d(M * numvec), e(M * numvec);
        CB2_CHECK(cudaMemcpyAsync(d.data(), lan_d_, d.size() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaMemcpyAsync(e.data(), lan_e_, e.size() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        for (std::size_t i = 0; i < numvec; ++i)
            if (b200::tridiag_eig_host((int)M, d.data() + i * M, e.data() + i * M, w.data() + i * M, Z.data() + i * M * M))
                throw std::runtime_error("chase_b200: the Lanczos tridiagonal eigensolver did not converge");
    }

    // every buffer tracks its own capacity: (M, numvec) may change between solves on the same object in ways that
    // shrink M * numvec but grow M * M * numvec or numvec + 1
    void ensure_lanczos_buffers(std::size_t M, std::size_t numvec)
    {
        if (lan_nv_ < numvec)
        {
            lan_v_ = alloc<T>(3 * ld_ * numvec);
            lan_rb_ = alloc<double>(numvec + 1);
            lan_nv_ = numvec;
        }
        if (lan_m_ < M * numvec)
        {
            lan_d_ = alloc<double>(M * numvec);
            lan_e_ = alloc<double>(M * numvec);
            lan_w_ = alloc<double>(M * numvec);
            lan_m_ = M * numvec;
        }
        if (lan_z_ < M * M * numvec)
        {
            lan_Z_ = alloc<double>(M * M * numvec);
            lan_z_ = M * M * numvec;
        }
    }

    void run_lanczos(std::size_t M, std::size_t numvec, R* upperb, R* Theta, R* Tau, R* ritzV, bool multi)
    {
        flush_perm();
        const int nv = (int)numvec, m = (int)M;
        ensure_lanczos_buffers(M, numvec);
        T* v0 = lan_v_;
        T* v1 = lan_v_ + ld_ * numvec;
        T* v2 = lan_v_ + 2 * ld_ * numvec;
        CB2_CHECK(cudaMemsetAsync(lan_d_, 0, M * numvec * sizeof(double), stream_));
        CB2_CHECK(cudaMemsetAsync(lan_e_, 0, M * numvec * sizeof(double), stream_));
        CB2_KCHECK(KK::lacpy((int64_t)N_, nv, dV1_, (int64_t)ld_, v1, (int64_t)ld_, stream_));
        CB2_KCHECK(KK::normalize_cols((int64_t)N_, nv, v1, (int64_t)ld_, stream_));
        for (int k = 0; k < m; ++k)
        {
            if (multi) // V1[:, k] <- current vector of the LAST run (cpu/lanczos.hpp:85-88)
                CB2_KCHECK(KK::lacpy((int64_t)N_, 1, v1 + (std::size_t)(nv - 1) * ld_, (int64_t)ld_, dV1_ + (std::size_t)k * ld_,
                                     (int64_t)ld_, stream_));
            CB2_KCHECK(KK::gemv_conjt((int64_t)N_, (int64_t)N_, dH_, (int64_t)ld_, v1, (int64_t)ld_, nv, v2,
                                      (int64_t)ld_, stream_));
            CB2_KCHECK(KK::lanczos_step((int64_t)N_, nv, k, m, v0, v1, v2, (int64_t)ld_, lan_d_, lan_e_, lan_rb_,
                                        stream_));
            if (k == m - 1)
                break;
            T* t = v0;
            v0 = v1;
            v1 = v2;
            v2 = t;
        }
        if (multi)
            CB2_KCHECK(KK::lacpy((int64_t)N_, nv, v1, (int64_t)ld_, dV1_, (int64_t)ld_, stream_));
        std::vector<double> w, Z, rb(numvec);
        CB2_CHECK(cudaMemcpyAsync(rb.data(), lan_rb_, rb.size() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        lanczos_tridiag_solve(M, numvec, w, Z);
        for (std::size_t i = 0; i < numvec; ++i)
        {
            for (std::size_t k = 0; k < M; ++k)
            {
                Theta[k + M * i] = (R)w[i * M + k];
                const R z0 = (R)Z[i * M * M + 0 + k * M];
                Tau[k + i * M] = std::abs(z0) * std::abs(z0);
            }
        }
        for (std::size_t q = 0; q < M * M; ++q)
            ritzV[q] = (R)Z[(numvec - 1) * M * M + q];
        R ub = std::max(std::abs(Theta[0]), std::abs(Theta[M - 1])) + std::abs((R)rb[0]);
        for (std::size_t i = 1; i < numvec; ++i)
        {
            const R mx = std::max(std::abs(Theta[i * M]), std::abs(Theta[(i + 1) * M - 1])) + std::abs((R)rb[i]);
            ub = std::max(mx, ub);
        }
        *upperb = ub;
    }

    std::size_t N_, nev_, nex_, nevex_;
    std::size_t nc_; // columns of the panels: nev+nex, or 2 (nev+nex) for pseudo-Hermitian problems
    T* H_;
    std::size_t ldh_;
    T* V_;
    std::size_t ldv_;
    R* ritzv_;
    ChaseConfig<T> config_;
    cudaStream_t stream_ = nullptr;
    std::size_t ld_ = 0, ldg_ = 0;
    T *dH_ = nullptr, *dV1_ = nullptr, *dV2_ = nullptr, *dW_ = nullptr, *dG_ = nullptr, *dZ_ = nullptr;
    T *dM_ = nullptr, *dRinv_ = nullptr, *dT_ = nullptr; // pseudo-Hermitian RR only
    unsigned char *dHw_ = nullptr, *wide_scratch_ = nullptr; // FP64 copy of an FP32 matrix + panel scratch
    std::size_t wide_scratch_bytes_ = 0;
    T* dHl_ = nullptr; // lo part of the TF32 split of an FP32 matrix (tcgen05 path)
    // mixed-precision filter of a double-precision problem: single-precision A (+ lo part), two panels, scratch
    unsigned char *dHs_ = nullptr, *dHsl_ = nullptr, *dV1s_ = nullptr, *dV2s_ = nullptr, *sp_scratch_ = nullptr;
    std::size_t sp_scratch_bytes_ = 0, sp_cols_ = 0, sp_filters_ = 0;
    bool mixed_ = false, sp_active_ = false, sp_matrix_valid_ = false;
    unsigned char* tf32_scratch_ = nullptr;
    std::size_t tf32_scratch_bytes_ = 0;
    double* ones_ = nullptr;
    T* dV0_ = nullptr; // device copy of the reference start block (parity mode), filled at the first random solve
    unsigned char *heev_ws_ = nullptr, *trsm_ws_ = nullptr, *splitk_ws_ = nullptr, *hh_ws_ = nullptr;
    std::size_t heev_ws_bytes_ = 0, trsm_ws_bytes_ = 0, splitk_ws_bytes_ = 0, hh_ws_bytes_ = 0;
    double *dTheta_ = nullptr, *dNorms_ = nullptr;
    int *dInfo_ = nullptr, *dIdx_ = nullptr;
    T* lan_v_ = nullptr;
    double *lan_d_ = nullptr, *lan_e_ = nullptr, *lan_w_ = nullptr, *lan_Z_ = nullptr, *lan_rb_ = nullptr;
    std::size_t lan_nv_ = 0, lan_m_ = 0, lan_z_ = 0;
    std::vector<void*> allocs_;
    std::vector<R> resid_;
    std::vector<int> perm_;
    bool perm_dirty_ = false;
    bool resid_ready_ = false; // dW_ holds (A Q) Z - (Q Z) Theta of the last RR
    std::size_t resid_block_ = 0;
    std::size_t locked_ = 0;
    double shift_ = 0.0;
    std::size_t lanczosIter_ = 0, numLanczos_ = 0;
    bool device_rng_ = false, is_sym_ = true, keep_device_matrix_ = false, matrix_on_device_ = false;
    std::string last_qr_;
    std::vector<std::string> qr_log_;
    std::size_t heev_sweeps_ = 0, hemm_cols_ = 0, swaps_ = 0, gathers_ = 0;
};

} // namespace Impl
} // namespace chase
