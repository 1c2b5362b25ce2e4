// chase_b200 host layer — distributed multi-GPU backend chase::Impl::pChASEGPU<T>.
//
// Drop-in for the reference's chase::Impl::pChASEGPU<MatrixType, InputMultiVectorType, NCCL>
// (Impl/pchase_gpu/pchase_gpu.hpp:121-1801) and the cuda_nccl kernels it calls
// (linalg/internal/nccl/{hemm,cholqr,rayleighRitz,residuals,lanczos,shiftDiagonal}.hpp): one process per GPU on an
// r x c grid, A block- or block-cyclic-distributed (Dist1D rows over grid rows, cols over grid columns), the
// "column-layout" panels V1/V2 split like A's rows and replicated over grid columns, the "row-layout" panels W*
// split like A's columns and replicated over grid rows (distMultiVector.hpp:1108-1120).
//
// What is the same as the reference: the layouts, the alternating HEMM (V1 -> W1 with A^H and an allreduce inside
// the grid column; W1 -> V1 with A and an allreduce inside the grid row; beta applied on one member of the reduce
// group only, nccl/hemm.hpp:282-289, 371-379), the in-place diagonal shift of A on precomputed local diagonal
// coordinates (pchase_gpu.hpp:340-409, 782-783), Gram allreduce inside the grid column for CholQR, the row-broadcast of
// V1 before RR (pchase_gpu.hpp:1631-1633), replicated small dense solves.
//
// What is new: every local product is the hand-written TMA + DMMA kernel (chase_b200_hemm_rect, both op(A) = A and
// A^H) or the generic DMMA GEMM; layout changes are ONE NCCL all-gather + one gather kernel instead of per-block
// broadcasts; the residual block re-uses A*Q from RR ((A Q) Z - (Q Z) Theta: no third distributed HEMM, no second
// redistribution per iteration); Lanczos keeps its numvec vectors replicated in full length, so its vector updates
// need no scalar allreduces (3 per step in nccl/lanczos.hpp:256-322); Gram / projected matrices are broadcast from
// one replica so that all ranks take bit-identical decisions; Swap storms become one gather pass.
//
// MatrixType = chase::matrix::PseudoHermitianMatrix<T, GPU> selects the pseudo-Hermitian (BSE) problem class on any
// of the layouts (the reference: PseudoHermitianBlockBlockMatrix / PseudoHermitianBlockCyclicMatrix, block-block
// multivectors only, pchase_gpu.hpp:903-1030): panels hold 2 (nev+nex) columns; H x is formed with the same
// two local products through H = S H^H S (S = sign flip of the rows whose global index is in the lower half);
// K-conjugation exchanges the two halves of full-length copies (all-gather inside the grid column) instead of the
// reference's pairwise send/recv (distMultiVector.hpp:1879-2060).
#pragma once
#include "algorithm.hpp"
#include "chase_gpu.hpp"
#include "comm.hpp"
#include "dist_layout.hpp"
#include "interface.hpp"
#include "kernel_api.hpp"
#include "tridiag_host.hpp"

#include <cuda_runtime_api.h>

#include <chrono>
#include <cstdio>
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

template <class T, class MatrixType = chase::matrix::Matrix<T, chase::platform::GPU>>
class pChASEGPU : public ChaseBase<T>
{
    using R = Base<T>;
    using KK = b200::K<T>;
    static constexpr bool kCplx = is_complex_t<T>::value;
    static constexpr bool kPseudo =
        std::is_same<MatrixType, chase::matrix::PseudoHermitianMatrix<T, chase::platform::GPU>>::value;
    static_assert(!kPseudo || kCplx, "pseudo-Hermitian (BSE) problems are complex (reference: c/z only)");

public:
    // H: this rank's local block (m_loc x n_loc, column-major, ldh); V: this rank's rows of the start / result
    // block (m_loc x (nev+nex), ldv); mb / nb: block-cyclic block sizes, 0 = the reference's block layout.
    pChASEGPU(std::size_t N, std::size_t nev, std::size_t nex, const b200::WorldComm& world, int dim0, int dim1,
              char grid_major, std::size_t mb, std::size_t nb, T* H, std::size_t ldh, T* V, std::size_t ldv, R* ritzv)
        : N_(N), nev_(nev), nex_(nex), nevex_(nev + nex), nc_(kPseudo ? 2 * (nev + nex) : nev + nex), H_(H), ldh_(ldh),
          V_(V), ldvh_(ldv), ritzv_(ritzv),
          grid_(b200::Grid2D::make(dim0, dim1, grid_major, world.rank, world.size)), comm_(world, grid_),
          Dr_((int64_t)N, dim0, (int64_t)mb), Dc_((int64_t)N, dim1, (int64_t)nb), config_(N, nev, nex)
    {
        if (N == 0 || nevex_ == 0 || nc_ > N)
            throw std::invalid_argument(kPseudo ? "pChASEGPU: need 0 < 2 (nev+nex) <= N"
                                                : "pChASEGPU: need 0 < nev+nex <= N");
        if (kPseudo && N % 2 != 0)
            throw std::invalid_argument("pChASEGPU: a pseudo-Hermitian matrix has even order");
        m_loc_ = (std::size_t)Dr_.local_size(grid_.i);
        n_loc_ = (std::size_t)Dc_.local_size(grid_.j);
        if (ldh < m_loc_ || ldv < m_loc_)
            throw std::invalid_argument("pChASEGPU: leading dimension smaller than the local row count");
        CB2_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        lda_ = roundup(std::max<std::size_t>(m_loc_, 1), 16);
        ldv_ = roundup(std::max<std::size_t>((std::size_t)Dr_.max_local_size(), 1), 16);
        ldw_ = roundup(std::max<std::size_t>((std::size_t)Dc_.max_local_size(), 1), 16);
        ldn_ = roundup(N_, 16);
        ldg_ = roundup(nc_, 16);
        dH_ = alloc<T>(lda_ * std::max<std::size_t>(n_loc_, 1));
        dV1_ = alloc<T>(ldv_ * nc_);
        dV2_ = alloc<T>(ldv_ * nc_);
        dVs_ = alloc<T>(ldv_ * nc_);
        for (auto& w : dW_)
            w = alloc<T>(ldw_ * nc_);
        gath_elems_ = std::max((std::size_t)grid_.r * ldv_, (std::size_t)grid_.c * ldw_) * nc_;
        dGath_ = alloc<T>(gath_elems_);
        dG_ = alloc<T>(ldg_ * nc_);
        dZ_ = alloc<T>(ldg_ * nc_);
        if (kPseudo)
        {
            dM_ = alloc<T>(ldg_ * nc_);
            dRinv_ = alloc<T>(ldg_ * nc_);
            dT_ = alloc<T>(ldg_ * nc_);
            dFull_ = alloc<T>(ldn_ * nevex_); // full-length copies of one half for the K-conjugation
            ones_ = alloc<double>(nc_);
            std::vector<double> one(nc_, 1.0);
            CB2_CHECK(cudaMemcpy(ones_, one.data(), nc_ * sizeof(double), cudaMemcpyHostToDevice));
        }
        heev_ws_bytes_ = chase_b200_heev_ws_bytes((int64_t)nc_, kCplx ? 1 : 0);
        heev_ws_ = alloc<unsigned char>(heev_ws_bytes_);
        trsm_ws_bytes_ = chase_b200_trsm_ws_bytes((int64_t)nc_, (int)sizeof(T));
        trsm_ws_ = alloc<unsigned char>(trsm_ws_bytes_);
        splitk_ws_bytes_ = std::max<std::size_t>(std::size_t(64) << 20, 4 * nc_ * nc_ * 16);
        splitk_ws_ = alloc<unsigned char>(splitk_ws_bytes_);
        dTheta_ = alloc<double>(std::max<std::size_t>(nc_, 64)); // >= communicator size (warm-up all-gather)
        dNorms_ = alloc<double>(std::max<std::size_t>(nc_, 64));
        dInfo_ = alloc<int>(4);
        dIdx_ = alloc<int>(2 * nc_);
        resid_.assign(nc_, R(0));
        perm_.resize(nc_);
        reset_perm();
        build_maps();
        // NCCL connects a communicator's peers lazily inside its first collective (0.1-0.3 s): do that here, like
        // the rest of the set-up, instead of inside the first QR of the first solve
        comm_.allreduce_sum(dNorms_, 1, comm_.world(), stream_);
        comm_.allreduce_sum(dNorms_, 1, comm_.row(), stream_);
        comm_.allreduce_sum(dNorms_, 1, comm_.col(), stream_);
        comm_.allgather(dNorms_, dTheta_, 1, comm_.row(), stream_);
        comm_.allgather(dNorms_, dTheta_, 1, comm_.col(), stream_);
        CB2_CHECK(cudaStreamSynchronize(stream_));
        CB2_CHECK(cudaMemsetAsync(dNorms_, 0, nc_ * sizeof(double), stream_));
        CB2_CHECK(cudaMemsetAsync(dTheta_, 0, nc_ * sizeof(double), stream_));
        const char* e = std::getenv("CHASE_B200_DEVICE_RNG");
        device_rng_ = e && std::atoi(e) != 0;
        // FP32 storage: FP64 copy of the local block for the TMA + DMMA filter kernel (see ChASEGPU)
        const char* w = std::getenv("CHASE_B200_FP32_WIDEN");
        if (sizeof(R) == 4 && !(w && std::atoi(w) == 0))
        {
            wide_scratch_bytes_ = (ldv_ + ldw_) * nc_ * 2 * sizeof(T);
            dHw_ = alloc<unsigned char>(lda_ * std::max<std::size_t>(n_loc_, 1) * 2 * sizeof(T));
            wide_scratch_ = alloc<unsigned char>(wide_scratch_bytes_);
            chase_b200_widen_register(dH_, dHw_, (int64_t)lda_, (int64_t)m_loc_, (int64_t)n_loc_, wide_scratch_,
                                      wide_scratch_bytes_);
            // The V -> W half of every distributed product, A_loc^H V, is the native orientation of the tcgen05
            // kind::tf32 kernel (hemm_tf32.cuh): with the lo part of the local block registered (kind 2 = only
            // op(A) = A^H) it runs there at ~5x the DMMA rate; the W -> V half (A_loc W, M-major operand) stays on the
            // FP64 copy.  CHASE_B200_FP32_PATH=fp64copy keeps both halves on the copy.
            const char* pth = std::getenv("CHASE_B200_FP32_PATH");
            if (!(pth && std::string(pth) != "tf32") && m_loc_ >= 128 && n_loc_ >= 128)
            {
                dHl_ = alloc<T>(lda_ * std::max<std::size_t>(n_loc_, 1));
                tf32_scratch_bytes_ = chase_b200_hemm_tf32_scratch_bytes((int64_t)m_loc_, (int64_t)nc_, (int)sizeof(T));
                tf32_scratch_ = alloc<unsigned char>(tf32_scratch_bytes_);
                chase_b200_tf32_register(dH_, dHl_, (int64_t)lda_, (int64_t)m_loc_, (int64_t)n_loc_, 2, tf32_scratch_,
                                         tf32_scratch_bytes_);
            }
        }
    }
    pChASEGPU(const pChASEGPU&) = delete;
    ~pChASEGPU() override
    {
        if (dHw_)
            chase_b200_widen_unregister(dH_);
        if (dHl_)
            chase_b200_tf32_unregister(dH_);
        cudaStreamSynchronize(stream_);
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
            // layout-independent Philox block (same matrix as the single-GPU backend draws)
            CB2_KCHECK(KK::rng_normal_rows((int64_t)m_loc_, (int64_t)nc_, map_full2v_, (int64_t)N_, dV1_,
                                           (int64_t)ldv_, 24141ull, stream_));
        }
        else if (random && dV0_ != nullptr)
        {
            // the reference stream is a pure function of (N, nev+nex, T): generated once, then kept on the device
            CB2_KCHECK(KK::lacpy((int64_t)m_loc_, (int64_t)nc_, dV0_, (int64_t)ldv_, dV1_, (int64_t)ldv_, stream_));
        }
        else
        {
            if (random)
            {
                // rows of the reference CPU stream (chase_cpu.hpp:296-309) owned by this grid row
                std::mt19937 gen(1337.0);
                std::normal_distribution<> d;
                const auto segs = Dr_.segments(grid_.i);
                for (std::size_t j = 0; j < nc_; ++j)
                {
                    std::size_t s = 0;
                    for (int64_t g = 0; g < (int64_t)N_; ++g)
                    {
                        const T x = getRandomT<T>([&]() { return d(gen); });
                        while (s < segs.size() && g >= segs[s].g0 + segs[s].len)
                            ++s;
                        if (s < segs.size() && g >= segs[s].g0)
                            V_[(std::size_t)(segs[s].l0 + (g - segs[s].g0)) + j * ldvh_] = x;
                    }
                }
            }
            if (m_loc_ > 0)
                CB2_CHECK(cudaMemcpy2DAsync(dV1_, ldv_ * sizeof(T), V_, ldvh_ * sizeof(T), m_loc_ * sizeof(T), nc_,
                                            cudaMemcpyHostToDevice, stream_));
            if (random)
            {
                dV0_ = alloc<T>(ldv_ * nc_);
                CB2_KCHECK(KK::lacpy((int64_t)m_loc_, (int64_t)nc_, dV1_, (int64_t)ldv_, dV0_, (int64_t)ldv_, stream_));
            }
        }
        if (random && kPseudo) // damp the lower (de-excitation) block: T(0.001), pchase_gpu.hpp:668-690
            CB2_KCHECK(KK::scale_rows_map((int64_t)m_loc_, (int64_t)nc_, map_full2v_, (int64_t)(N_ / 2), dV1_,
                                          (int64_t)ldv_, (double)(R)0.001, stream_));
        CB2_KCHECK(KK::lacpy((int64_t)ldv_, (int64_t)nc_, dV1_, (int64_t)ldv_, dV2_, (int64_t)ldv_, stream_));
        if (!(keep_device_matrix_ && matrix_on_device_) && m_loc_ > 0 && n_loc_ > 0)
        {
            CB2_CHECK(cudaMemcpy2DAsync(dH_, lda_ * sizeof(T), H_, ldh_ * sizeof(T), m_loc_ * sizeof(T), n_loc_,
                                        cudaMemcpyHostToDevice, stream_));
            matrix_on_device_ = true;
            wide_valid_ = false;
        }
        if (dHw_ && !wide_valid_)
        {
            CB2_KCHECK(chase_b200_widen_sync(kCplx ? 'c' : 's', dH_, stream_));
            if (dHl_)
                CB2_KCHECK(chase_b200_tf32_sync(kCplx ? 'c' : 's', dH_, stream_));
            wide_valid_ = true;
        }
        reset_perm();
        next_ = NextOp::bAc;
        resid_ready_ = false;
    }

    // in-place diagonal update on the local copies of global diagonal entries (pchase_gpu.hpp:773-783)
    void Shift(T c, bool isunshift = false) override
    {
        CB2_KCHECK(KK::shift_diag_list((int64_t)ndiag_, diag_lin_, dH_, (double)std::real(c), stream_));
        if (dHw_) // the FP64 copy of an FP32 block follows (same linear indices)
        {
            using TW = typename std::conditional<kCplx, std::complex<double>, double>::type;
            CB2_KCHECK(b200::K<TW>::shift_diag_list((int64_t)ndiag_, diag_lin_, dHw_, (double)std::real(c), stream_));
        }
        if (dHl_) // ... and the lo parts of the shifted entries
            CB2_KCHECK(chase_b200_tf32_sync_list(kCplx ? 'c' : 's', dH_, (int64_t)ndiag_, diag_lin_, stream_));
        if (isunshift)
            next_ = NextOp::bAc;
    }

    void HEMM(std::size_t block, T alpha, T beta, std::size_t offset_left, std::size_t offset_right = 0) override
    {
        flush_perm();
        resid_ready_ = false;
        const std::size_t ncols = (offset_right < block) ? block - offset_right : 0;
        const std::size_t c0 = offset_left + locked_;
        if (ncols > 0)
        {
            if (next_ == NextOp::bAc)
                hemm_v2w(alpha, dV1_ + c0 * ldv_, beta, dW_[0] + c0 * ldw_, ncols);
            else
                hemm_w2v(alpha, dW_[0] + c0 * ldw_, beta, dV1_ + c0 * ldv_, ncols);
            hemm_cols_ += ncols;
        }
        next_ = (next_ == NextOp::bAc) ? NextOp::cAb : NextOp::bAc;
    }

    // V2[cols] <- alpha H (H V1[cols]) + beta V2[cols] + gamma V1[cols]; swap (pchase_gpu.hpp:903-945).  One round
    // trip through the row layout: W = H V1 = S (A^H (S V1)) with the A^H product, V2 = alpha A W + beta V2 with the A
    // product.  Columns as in ChASEGPU::HEMM_H2: [locked_ + offset_left, locked_ + block - offset_right).
    void HEMM_H2(std::size_t block, T alpha, T beta, T gamma, std::size_t offset_left,
                 std::size_t offset_right = 0) override
    {
        if (!kPseudo)
            throw std::runtime_error("chase_b200: HEMM_H2 needs MatrixType = PseudoHermitianMatrix");
        flush_perm();
        resid_ready_ = false;
        std::size_t ncols = (offset_right < block) ? block - offset_right : 0;
        ncols = (offset_left < ncols) ? ncols - offset_left : 0;
        if (ncols > 0)
        {
            const std::size_t c0 = offset_left + locked_;
            T* in = dV1_ + c0 * ldv_;
            T* out = dV2_ + c0 * ldv_;
            T* tmp = dW_[0] + c0 * ldw_;
            times_H_v2w(in, tmp, ncols);
            hemm_w2v(alpha, tmp, beta, out, ncols);
            CB2_KCHECK(KK::axpy_cols((int64_t)m_loc_, (int64_t)ncols, ones_, b200::re_of(gamma), b200::im_of(gamma), in,
                                     (int64_t)ldv_, out, (int64_t)ldv_, stream_));
            hemm_cols_ += 2 * ncols;
        }
        std::swap(dV1_, dV2_);
    }

    // V1[:, 2 nevex - locked - block ...) <- K-conjugates of V1[:, locked ... locked + block): the partner of row g
    // is row g +- N/2, in general on another rank, so the halves are exchanged on full-length copies
    void ApplyKconjugate(std::size_t block) override
    {
        if (!kPseudo)
            return;
        flush_perm();
        if (block == 0)
            return;
        const std::size_t col_second = nc_ - locked_ - block;
        v_to_full(dV1_ + locked_ * ldv_, dFull_, block);
        CB2_KCHECK(KK::kconj((int64_t)N_, (int64_t)block, dFull_, (int64_t)ldn_, dFull_, (int64_t)ldn_, stream_));
        full_to_v(dFull_, dV1_ + col_second * ldv_, block);
    }

    void QR(std::size_t /*fixednev*/, R cond) override
    {
        flush_perm();
        resid_ready_ = false;
        CB2_KCHECK(KK::lacpy((int64_t)m_loc_, (int64_t)locked_, dV1_, (int64_t)ldv_, dV2_, (int64_t)ldv_, stream_));
        if (kPseudo)
        {
            // [L+ | active | L-] -> orthogonalise [S L+ | S L- | active] (pchase_gpu.hpp:961-1030)
            const std::size_t act = nc_ - 2 * locked_;
            const int64_t m = (int64_t)m_loc_, ld = (int64_t)ldv_;
            CB2_KCHECK(KK::lacpy(m, (int64_t)locked_, dV1_ + (nc_ - locked_) * ldv_, ld, dV2_ + (nc_ - locked_) * ldv_, ld,
                                 stream_));
            CB2_KCHECK(KK::lacpy(m, (int64_t)locked_, dV1_, ld, dVs_, ld, stream_));
            CB2_KCHECK(KK::lacpy(m, (int64_t)locked_, dV1_ + (nc_ - locked_) * ldv_, ld, dVs_ + locked_ * ldv_, ld,
                                 stream_));
            CB2_KCHECK(KK::lacpy(m, (int64_t)act, dV1_ + locked_ * ldv_, ld, dVs_ + 2 * locked_ * ldv_, ld, stream_));
            std::swap(dV1_, dVs_);
            flip_v(dV1_, 2 * locked_);
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
            householder(); // qr == 'H' / CHASE_DISABLE_CHOLQR=1 (pchase_gpu.hpp:1117-1192)
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
            householder(); // CholeskyQR broke down: Householder QR, as the reference (pchase_gpu.hpp:1455-1533)
            last_qr_ += "+householder";
        }
        qr_log_.push_back(last_qr_);
        if (kPseudo)
        {
            const std::size_t act = nc_ - 2 * locked_;
            const int64_t m = (int64_t)m_loc_, ld = (int64_t)ldv_;
            CB2_KCHECK(KK::lacpy(m, (int64_t)act, dV1_ + 2 * locked_ * ldv_, ld, dVs_ + locked_ * ldv_, ld, stream_));
            std::swap(dV1_, dVs_);
            CB2_KCHECK(KK::lacpy(m, (int64_t)locked_, dV2_ + (nc_ - locked_) * ldv_, ld, dV1_ + (nc_ - locked_) * ldv_, ld,
                                 stream_));
            if (locked_ == 0) // both panels hold the orthonormal block (LanczosDos reads the second one)
                CB2_KCHECK(KK::lacpy(m, (int64_t)nc_, dV1_, ld, dV2_, ld, stream_));
        }
        CB2_KCHECK(KK::lacpy((int64_t)m_loc_, (int64_t)locked_, dV2_, (int64_t)ldv_, dV1_, (int64_t)ldv_, stream_));
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
        T* Q = dV1_ + locked_ * ldv_;
        T* W1 = dW_[0] + locked_ * ldw_;
        T* W2 = dW_[1] + locked_ * ldw_;
        // replicas of V1 inside a grid row are re-synchronised first (pchase_gpu.hpp:1631-1633)
        if (grid_.c > 1)
            comm_.broadcast(Q, ldv_ * block, 0, comm_.row(), stream_);
        // W1 = A^H Q (row layout); FP32 types: all four TF32 partial products for the projected matrix
        chase_b200_tf32_set_terms(4);
        hemm_v2w(T(1), Q, T(0), W1, block);
        chase_b200_tf32_set_terms(3);
        // W2 = Q in row layout
        redistribute_v2w(Q, W2, block);
        // G = W2^H W1 summed over the grid row, then made bit-identical everywhere
        CB2_KCHECK(KK::gemm(1, 0, (int64_t)block, (int64_t)block, (int64_t)n_loc_, 1.0, 0.0, W2, (int64_t)ldw_, W1,
                            (int64_t)ldw_, 0.0, 0.0, dG_, (int64_t)ldg_, 0, splitk_ws_, splitk_ws_bytes_, stream_));
        sum_triangle(dG_, (int64_t)block, true, false); // the eigensolver reads the lower triangle
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
        // V2 = Q Z ; swap
        CB2_KCHECK(KK::gemm(0, 0, (int64_t)m_loc_, (int64_t)block, (int64_t)block, 1.0, 0.0, Q, (int64_t)ldv_, dZ_,
                            (int64_t)ldg_, 0.0, 0.0, dV2_ + locked_ * ldv_, (int64_t)ldv_, 0, nullptr, 0, stream_));
        std::swap(dV1_, dV2_);
        // residual block in row layout while A Q is at hand: (A Q) Z - (Q Z) Theta
        for (std::size_t i = 0; i < block; ++i)
            w[i] = (double)ritzv[i]; // rounded to Base<T> like the values the driver passes to Resd
        CB2_CHECK(cudaMemcpyAsync(dTheta_, w.data(), block * sizeof(double), cudaMemcpyHostToDevice, stream_));
        T* E = dW_[2] + locked_ * ldw_;
        T* Rm = dW_[3] + locked_ * ldw_;
        CB2_KCHECK(KK::gemm(0, 0, (int64_t)n_loc_, (int64_t)block, (int64_t)block, 1.0, 0.0, W2, (int64_t)ldw_, dZ_,
                            (int64_t)ldg_, 0.0, 0.0, E, (int64_t)ldw_, 0, nullptr, 0, stream_));
        CB2_KCHECK(KK::gemm(0, 0, (int64_t)n_loc_, (int64_t)block, (int64_t)block, 1.0, 0.0, W1, (int64_t)ldw_, dZ_,
                            (int64_t)ldg_, 0.0, 0.0, Rm, (int64_t)ldw_, 0, nullptr, 0, stream_));
        CB2_KCHECK(KK::axpy_cols((int64_t)n_loc_, (int64_t)block, dTheta_, -1.0, 0.0, E, (int64_t)ldw_, Rm,
                                 (int64_t)ldw_, stream_));
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
        T* Rm = dW_[3] + locked_ * ldw_;
        if (!(resid_ready_ && resid_block_ == k))
        {
            // general path (not taken by chase::Solve, which always calls RR first): A^H V, V in row layout
            std::vector<double> th(k);
            for (std::size_t i = 0; i < k; ++i)
                th[i] = (double)ritzv[i];
            CB2_CHECK(cudaMemcpyAsync(dTheta_, th.data(), k * sizeof(double), cudaMemcpyHostToDevice, stream_));
            T* Vb = dV1_ + locked_ * ldv_;
            T* E = dW_[2] + locked_ * ldw_;
            if (kPseudo)
                times_H_v2w(Vb, Rm, k); // H is not Hermitian: H V = S A^H S V
            else
                hemm_v2w(T(1), Vb, T(0), Rm, k);
            redistribute_v2w(Vb, E, k);
            CB2_KCHECK(KK::axpy_cols((int64_t)n_loc_, (int64_t)k, dTheta_, -1.0, 0.0, E, (int64_t)ldw_, Rm,
                                     (int64_t)ldw_, stream_));
            CB2_CHECK(cudaStreamSynchronize(stream_));
        }
        CB2_KCHECK(KK::colnorms((int64_t)n_loc_, (int64_t)k, Rm, (int64_t)ldw_, dNorms_, 0, stream_));
        if (grid_.c > 1)
            comm_.allreduce_sum(dNorms_, k, comm_.row(), stream_);
        if (grid_.r > 1)
            comm_.broadcast(dNorms_, k, 0, comm_.col(), stream_);
        std::vector<double> nr(k);
        CB2_CHECK(cudaMemcpyAsync(nr.data(), dNorms_, k * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        for (std::size_t i = 0; i < k; ++i)
            resd[i] = (R)std::sqrt(nr[i]);
        resid_ready_ = false;
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
        CB2_KCHECK(KK::gemm(0, 0, (int64_t)m_loc_, (int64_t)idx, (int64_t)m, 1.0, 0.0, dV1_, (int64_t)ldv_, dZ_,
                            (int64_t)ldg_, 0.0, 0.0, dV2_, (int64_t)ldv_, 0, nullptr, 0, stream_));
        CB2_KCHECK(KK::lacpy((int64_t)m_loc_, (int64_t)m, dV2_, (int64_t)ldv_, dV1_, (int64_t)ldv_, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
    }

    void Swap(std::size_t i, std::size_t j) override
    {
        std::swap(perm_[i], perm_[j]);
        perm_dirty_ = true;
        swaps_++;
    }

    void Lock(std::size_t new_converged) override { locked_ += new_converged; }

    // The reference checks A v == A^H v on a random vector through two distributed HEMMs
    // (nccl/symOrHerm.hpp); same idea here with the two local products.
    bool checkSymmetryEasy() override
    {
        flush_perm();
        if (!matrix_on_device_ && m_loc_ > 0 && n_loc_ > 0)
        {
            CB2_CHECK(cudaMemcpy2DAsync(dH_, lda_ * sizeof(T), H_, ldh_ * sizeof(T), m_loc_ * sizeof(T), n_loc_,
                                        cudaMemcpyHostToDevice, stream_));
            matrix_on_device_ = true;
        }
        // x (column layout, one Philox column) -> y1 = A^H x (row layout);  x in row layout -> y2 = A x (column layout)
        T* x = dVs_;
        CB2_KCHECK(KK::rng_normal_rows((int64_t)m_loc_, 1, map_full2v_, (int64_t)N_, x, (int64_t)ldv_, 777ull, stream_));
        T* y1 = dW_[2];
        T* xw = dW_[3];
        hemm_v2w(T(1), x, T(0), y1, 1);
        redistribute_v2w(x, xw, 1);
        T* y2 = dVs_ + ldv_;
        hemm_w2v(T(1), xw, T(0), y2, 1);
        // compare y1 (row layout) with y2 moved to the row layout
        T* y2w = dW_[3] + ldw_;
        redistribute_v2w(y2, y2w, 1);
        double two[2] = {-1.0, 0.0};
        CB2_CHECK(cudaMemcpyAsync(dTheta_, two, sizeof(double), cudaMemcpyHostToDevice, stream_));
        CB2_KCHECK(KK::colnorms((int64_t)n_loc_, 1, y1, (int64_t)ldw_, dNorms_ + 1, 0, stream_));
        CB2_KCHECK(KK::axpy_cols((int64_t)n_loc_, 1, dTheta_, 1.0, 0.0, y2w, (int64_t)ldw_, y1, (int64_t)ldw_, stream_));
        CB2_KCHECK(KK::colnorms((int64_t)n_loc_, 1, y1, (int64_t)ldw_, dNorms_, 0, stream_));
        if (grid_.c > 1)
            comm_.allreduce_sum(dNorms_, 2, comm_.row(), stream_);
        if (grid_.r > 1)
            comm_.broadcast(dNorms_, 2, 0, comm_.col(), stream_);
        double nr[2];
        CB2_CHECK(cudaMemcpyAsync(nr, dNorms_, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        const double tol = (sizeof(R) == 8) ? 1e-10 : 1e-4;
        is_sym_ = std::sqrt(nr[0]) <= tol * std::sqrt(nr[1]);
        return is_sym_;
    }
    bool isSym() override { return !kPseudo; }
    // S H x == H^H S x on a random vector (the distributed analogue of the single-GPU check)
    bool checkPseudoHermicityEasy() override
    {
        if (!kPseudo)
            return false;
        flush_perm();
        if (!matrix_on_device_ && m_loc_ > 0 && n_loc_ > 0)
        {
            CB2_CHECK(cudaMemcpy2DAsync(dH_, lda_ * sizeof(T), H_, ldh_ * sizeof(T), m_loc_ * sizeof(T), n_loc_,
                                        cudaMemcpyHostToDevice, stream_));
            matrix_on_device_ = true;
        }
        T* x = dVs_;
        CB2_KCHECK(KK::rng_normal_rows((int64_t)m_loc_, 1, map_full2v_, (int64_t)N_, x, (int64_t)ldv_, 777ull, stream_));
        T* y1 = dW_[2]; // A^H (S x)                (row layout)
        T* xw = dW_[3];
        flip_v(x, 1);
        hemm_v2w(T(1), x, T(0), y1, 1);
        flip_v(x, 1);
        redistribute_v2w(x, xw, 1);
        T* y2 = dVs_ + ldv_; // S (A x)             (column layout)
        hemm_w2v(T(1), xw, T(0), y2, 1);
        flip_v(y2, 1);
        T* y2w = dW_[3] + ldw_;
        redistribute_v2w(y2, y2w, 1);
        double m1[2] = {-1.0, 0.0};
        CB2_CHECK(cudaMemcpyAsync(dTheta_, m1, sizeof(double), cudaMemcpyHostToDevice, stream_));
        CB2_KCHECK(KK::colnorms((int64_t)n_loc_, 1, y1, (int64_t)ldw_, dNorms_ + 1, 0, stream_));
        CB2_KCHECK(KK::axpy_cols((int64_t)n_loc_, 1, dTheta_, 1.0, 0.0, y2w, (int64_t)ldw_, y1, (int64_t)ldw_, stream_));
        CB2_KCHECK(KK::colnorms((int64_t)n_loc_, 1, y1, (int64_t)ldw_, dNorms_, 0, stream_));
        if (grid_.c > 1)
            comm_.allreduce_sum(dNorms_, 2, comm_.row(), stream_);
        if (grid_.r > 1)
            comm_.broadcast(dNorms_, 2, 0, comm_.col(), stream_);
        double nr[2];
        CB2_CHECK(cudaMemcpyAsync(nr, dNorms_, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        const double tol = (sizeof(R) == 8) ? 1e-10 : 1e-4;
        return std::sqrt(nr[0]) <= tol * std::sqrt(nr[1]);
    }
    bool isPseudoHerm() override { return kPseudo; }
    // The reference completes a distributed matrix from one triangle with ScaLAPACK's p?tranc and throws in builds
    // without ScaLAPACK (linalg/internal/nccl/symOrHerm.hpp:96-158, 262-264).  This library is the ScaLAPACK-less
    // configuration (chase_has_scalapack_ reports 0), so it mirrors that error; single-GPU matrices are completed
    // (ChASEGPU::symOrHermMatrix).
    void symOrHermMatrix(char) override
    {
        throw std::runtime_error("For ChASE-MPI, symOrHermMatrix requires ScaLAPACK, which is not detected "
                                 "(chase_b200 is built without ScaLAPACK: pass the full distributed matrix)");
    }

    void End() override
    {
        flush_perm();
        if (qrt_.on && grid_.rank == 0)
            std::fprintf(stderr, "chase_b200 QR timing: %d rounds, gram %.3f s (max %.4f), allreduce+bcast %.3f s (max "
                                 "%.4f), potrf %.3f s (max %.4f), trsm %.3f s (max %.4f)\n", qrt_.rounds, qrt_.t[0],
                         qrt_.tmax[0], qrt_.t[1], qrt_.tmax[1], qrt_.t[2], qrt_.tmax[2], qrt_.t[3], qrt_.tmax[3]);
        if (m_loc_ > 0)
            CB2_CHECK(cudaMemcpy2DAsync(V_, ldvh_ * sizeof(T), dV1_, ldv_ * sizeof(T), m_loc_ * sizeof(T), nc_,
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
    int get_nprocs() override { return grid_.nranks; }
    int get_rank() override { return grid_.rank; }
    void Output(LogLevel, std::string s, const char* = "algorithm") override
    {
        if (grid_.rank == 0 && std::getenv("CHASE_B200_VERBOSE"))
            std::cout << s;
    }

    // Global matrix files (raw column-major N x N, the reference's format): every rank reads / writes the pieces of
    // its own local block at their global offsets (reference: MPI-IO darray views for block-cyclic matrices,
    // linalg/distMatrix/distMatrix.hpp:3117-3196; block matrices: MPI-IO subarray views, :2241-2330).  All ranks of one node share the
    // file; the caller synchronises the ranks between a write and a read.
    void loadProblemFromFile(const std::string& filename)
    {
        if (H_ == nullptr)
            throw std::runtime_error("chase_b200: no host matrix buffer (device-resident hand-over) to read into");
        std::ifstream f(filename, std::ios::binary);
        if (!f.is_open())
            throw std::runtime_error("chase_b200: cannot open " + filename + " for reading");
        f.seekg(0, std::ios::end);
        if ((std::size_t)f.tellg() < N_ * N_ * sizeof(T))
            throw std::runtime_error("chase_b200: " + filename + " is smaller than the N x N matrix");
        const auto segs = Dr_.segments(grid_.i);
        const auto gcols = Dc_.global_indices(grid_.j);
        for (std::size_t lc = 0; lc < gcols.size(); ++lc)
            for (const auto& sg : segs)
            {
                f.seekg((std::streamoff)(((std::size_t)gcols[lc] * N_ + (std::size_t)sg.g0) * sizeof(T)), std::ios::beg);
                f.read(reinterpret_cast<char*>(H_ + (std::size_t)sg.l0 + lc * ldh_), (std::streamsize)(sg.len * sizeof(T)));
            }
        matrix_on_device_ = false;
    }
    void saveProblemToFile(const std::string& filename)
    {
        if (H_ == nullptr)
            throw std::runtime_error("chase_b200: no host matrix buffer (device-resident hand-over) to write");
        // in | out without trunc: every rank patches its pieces into the shared file (created by whoever is first)
        {
            std::ofstream create(filename, std::ios::binary | std::ios::app);
        }
        std::fstream f(filename, std::ios::binary | std::ios::in | std::ios::out);
        if (!f.is_open())
            throw std::runtime_error("chase_b200: cannot open " + filename + " for writing");
        const auto segs = Dr_.segments(grid_.i);
        const auto gcols = Dc_.global_indices(grid_.j);
        for (std::size_t lc = 0; lc < gcols.size(); ++lc)
            for (const auto& sg : segs)
            {
                f.seekp((std::streamoff)(((std::size_t)gcols[lc] * N_ + (std::size_t)sg.g0) * sizeof(T)), std::ios::beg);
                f.write(reinterpret_cast<const char*>(H_ + (std::size_t)sg.l0 + lc * ldh_),
                        (std::streamsize)(sg.len * sizeof(T)));
            }
    }

    // ---- extras ----------------------------------------------------------------
    const std::vector<std::string>& qr_log() const { return qr_log_; }
    void clear_logs()
    {
        qr_log_.clear();
        heev_sweeps_ = 0;
        hemm_cols_ = 0;
        swaps_ = 0;
        gathers_ = 0;
    }
    void keep_device_matrix(bool f) { keep_device_matrix_ = f; }
    void use_device_rng(bool f) { device_rng_ = f; }
    // the mixed-precision filter (single-precision tcgen05 kernel while the residuals are above 1e-3) is built for
    // the single-GPU backend; the distributed HEMM needs the M-major operand orientation the kind::tf32 kernel lacks
    void use_mixed_precision(bool) {}
    std::size_t sp_filter_cols() const { return 0; }
    std::size_t heev_sweeps() const { return heev_sweeps_; }
    std::size_t hemm_cols() const { return hemm_cols_; } // columns actually multiplied by the filter
    std::size_t gather_passes() const { return gathers_; }
    std::size_t collectives() const { return comm_.collectives(); }
    std::size_t local_rows() const { return m_loc_; }
    std::size_t local_cols() const { return n_loc_; }
    const b200::Grid2D& grid() const { return grid_; }
    cudaStream_t stream() const { return stream_; }
    T* device_H() { return dH_; }
    std::size_t device_lda() const { return lda_; }
    void mark_matrix_on_device()
    {
        matrix_on_device_ = true;
        wide_valid_ = false;
    }

private:
    enum class NextOp
    {
        bAc, // next HEMM maps V1 (column layout) -> W1 (row layout) with A^H
        cAb  // next HEMM maps W1 -> V1 with A
    };
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
    int64_t* upload_map(const std::vector<int64_t>& h)
    {
        int64_t* d = alloc<int64_t>(h.size());
        if (!h.empty())
            CB2_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        return d;
    }
    static std::vector<int64_t> expand(const std::vector<b200::RowCopy>& list, std::size_t dst_rows)
    {
        std::vector<int64_t> m(dst_rows, -1);
        for (const auto& c : list)
            for (int64_t t = 0; t < c.len; ++t)
                m[(std::size_t)(c.dst0 + t)] = c.src0 + t;
        return m;
    }
    void build_maps()
    {
        const b200::Dist1D full((int64_t)N_, 1, 0);
        // my row-layout piece out of the column-layout pieces gathered over the grid column
        map_v2w_ = upload_map(expand(b200::redistribution_list(Dr_, (int64_t)ldv_, Dc_, grid_.j), n_loc_));
        // full-length vectors out of gathered column-layout / row-layout pieces
        map_v2full_ = upload_map(expand(b200::redistribution_list(Dr_, (int64_t)ldv_, full, 0), N_));
        map_w2full_ = upload_map(expand(b200::redistribution_list(Dc_, (int64_t)ldw_, full, 0), N_));
        // my column-layout rows out of a full-length vector (= their global indices); same for the row layout
        map_full2v_ = upload_map(Dr_.global_indices(grid_.i));
        map_full2w_ = upload_map(Dc_.global_indices(grid_.j));
        // local copies of global diagonal entries
        std::vector<int64_t> lin;
        const auto gcols = Dc_.global_indices(grid_.j);
        for (std::size_t lc = 0; lc < gcols.size(); ++lc)
        {
            const int64_t g = gcols[lc];
            if (Dr_.owner(g) == grid_.i)
                lin.push_back(Dr_.local_index(g) + (int64_t)lc * (int64_t)lda_);
        }
        ndiag_ = lin.size();
        diag_lin_ = upload_map(lin);
    }

    // W(n_loc x k, row layout) <- alpha A_loc^H V(m_loc x k) + beta W, summed inside the grid column
    void hemm_v2w(T alpha, const T* V, T beta, T* W, std::size_t k)
    {
        const T b = (grid_.i == 0) ? beta : T(0);
        CB2_KCHECK(KK::hemm_rect(1, (int64_t)n_loc_, (int64_t)m_loc_, (int64_t)k, b200::re_of(alpha), b200::im_of(alpha),
                                 dH_, (int64_t)lda_, V, (int64_t)ldv_, b200::re_of(b), b200::im_of(b), W, (int64_t)ldw_,
                                 stream_));
        if (grid_.r > 1)
            comm_.allreduce_sum(W, ldw_ * k, comm_.col(), stream_);
    }
    // V(m_loc x k, column layout) <- alpha A_loc W(n_loc x k) + beta V, summed inside the grid row
    void hemm_w2v(T alpha, const T* W, T beta, T* V, std::size_t k)
    {
        const T b = (grid_.j == 0) ? beta : T(0);
        CB2_KCHECK(KK::hemm_rect(0, (int64_t)m_loc_, (int64_t)n_loc_, (int64_t)k, b200::re_of(alpha), b200::im_of(alpha),
                                 dH_, (int64_t)lda_, W, (int64_t)ldw_, b200::re_of(b), b200::im_of(b), V, (int64_t)ldv_,
                                 stream_));
        if (grid_.c > 1)
            comm_.allreduce_sum(V, ldv_ * k, comm_.row(), stream_);
    }
    // column layout -> row layout: all-gather inside the grid column + one gather kernel
    void redistribute_v2w(const T* V, T* W, std::size_t k)
    {
        if (grid_.r == 1)
        {
            CB2_KCHECK(KK::gather_rows((int64_t)n_loc_, (int64_t)k, map_v2w_, V, (int64_t)ldv_, 0, W, (int64_t)ldw_,
                                       stream_));
            return;
        }
        comm_.allgather(V, dGath_, ldv_ * k, comm_.col(), stream_);
        CB2_KCHECK(KK::gather_rows((int64_t)n_loc_, (int64_t)k, map_v2w_, dGath_, (int64_t)ldv_, (int64_t)(ldv_ * k), W,
                                   (int64_t)ldw_, stream_));
    }
    // full-length replicated copies (Lanczos)
    void v_to_full(const T* V, T* F, std::size_t k)
    {
        if (grid_.r == 1)
        {
            CB2_KCHECK(KK::gather_rows((int64_t)N_, (int64_t)k, map_v2full_, V, (int64_t)ldv_, 0, F, (int64_t)ldn_,
                                       stream_));
            return;
        }
        comm_.allgather(V, dGath_, ldv_ * k, comm_.col(), stream_);
        CB2_KCHECK(KK::gather_rows((int64_t)N_, (int64_t)k, map_v2full_, dGath_, (int64_t)ldv_, (int64_t)(ldv_ * k), F,
                                   (int64_t)ldn_, stream_));
    }
    void w_to_full(const T* W, T* F, std::size_t k)
    {
        if (grid_.c == 1)
        {
            CB2_KCHECK(KK::gather_rows((int64_t)N_, (int64_t)k, map_w2full_, W, (int64_t)ldw_, 0, F, (int64_t)ldn_,
                                       stream_));
            return;
        }
        comm_.allgather(W, dGath_, ldw_ * k, comm_.row(), stream_);
        CB2_KCHECK(KK::gather_rows((int64_t)N_, (int64_t)k, map_w2full_, dGath_, (int64_t)ldw_, (int64_t)(ldw_ * k), F,
                                   (int64_t)ldn_, stream_));
    }
    void full_to_v(const T* F, T* V, std::size_t k)
    {
        CB2_KCHECK(KK::gather_rows((int64_t)m_loc_, (int64_t)k, map_full2v_, F, (int64_t)ldn_, 0, V, (int64_t)ldv_,
                                   stream_));
    }

    void reset_perm()
    {
        for (std::size_t i = 0; i < nc_; ++i)
            perm_[i] = (int)i;
        perm_dirty_ = false;
    }
    void flush_perm()
    {
        if (!perm_dirty_)
            return;
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
            std::vector<int> idx = src;
            idx.insert(idx.end(), dst.begin(), dst.end());
            CB2_CHECK(cudaMemcpyAsync(dIdx_, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
            CB2_KCHECK(KK::gather_cols((int64_t)m_loc_, cnt, dIdx_, dIdx_ + cnt, dV1_, (int64_t)ldv_, dVs_,
                                       (int64_t)ldv_, stream_));
            CB2_KCHECK(KK::gather_cols((int64_t)m_loc_, cnt, dIdx_ + cnt, dIdx_ + cnt, dVs_, (int64_t)ldv_, dV1_,
                                       (int64_t)ldv_, stream_));
            CB2_CHECK(cudaStreamSynchronize(stream_));
            gathers_++;
        }
        reset_perm();
    }

    // one CholQR round on all nev+nex columns (nccl/cholqr.hpp:59-256): local Gram, allreduce inside the grid
    // column, replicated Cholesky, local TRSM
    // CHASE_B200_QR_TIMING=1: wall-clock split of the CholQR rounds (synchronises after every step; diagnostics only)
    struct QrTimer
    {
        bool on = std::getenv("CHASE_B200_QR_TIMING") != nullptr;
        double t[5] = {0, 0, 0, 0, 0}, tmax[5] = {0, 0, 0, 0, 0};
        int rounds = 0;
        std::chrono::high_resolution_clock::time_point t0;
        void start(cudaStream_t s)
        {
            if (!on)
                return;
            cudaStreamSynchronize(s);
            t0 = std::chrono::high_resolution_clock::now();
        }
        void lap(cudaStream_t s, int slot)
        {
            if (!on)
                return;
            cudaStreamSynchronize(s);
            const auto t1 = std::chrono::high_resolution_clock::now();
            const double dt = std::chrono::duration<double>(t1 - t0).count();
            t[slot] += dt;
            tmax[slot] = std::max(tmax[slot], dt);
            t0 = t1;
        }
    } qrt_;

    // Sum of the per-rank contributions to a Hermitian n x n matrix of which only one triangle is needed afterwards
    // (upper: potrf; lower: heev), made bit-identical on every rank: the packed triangle (n (n+1) / 2 elements, half
    // the padded square) is all-reduced inside one communicator and broadcast inside the other, like the reference's
    // extractUpperTriangular + allreduce + unpack (nccl/cholqr.hpp:152-157, nccl/rayleighRitz.hpp:124-132).
    void sum_triangle(T* G, int64_t n, bool lower, bool reduce_in_col)
    {
        const bool red = reduce_in_col ? grid_.r > 1 : grid_.c > 1;
        const bool bc = reduce_in_col ? grid_.c > 1 : grid_.r > 1;
        if (!red && !bc)
            return;
        const std::size_t cnt = (std::size_t)n * (std::size_t)(n + 1) / 2;
        if (dPack_ == nullptr)
            dPack_ = alloc<T>(nc_ * (nc_ + 1) / 2);
        CB2_KCHECK(KK::tri_pack(n, G, (int64_t)ldg_, dPack_, lower ? 1 : 0, stream_));
        if (red)
            comm_.allreduce_sum(dPack_, cnt, reduce_in_col ? comm_.col() : comm_.row(), stream_);
        if (bc)
            comm_.broadcast(dPack_, cnt, 0, reduce_in_col ? comm_.row() : comm_.col(), stream_);
        CB2_KCHECK(KK::tri_unpack(n, dPack_, G, (int64_t)ldg_, lower ? 1 : 0, stream_));
    }

    int chol_round(bool shifted, double shift_boost)
    {
        const int64_t n = (int64_t)nc_;
        qrt_.start(stream_);
        CB2_KCHECK(KK::gemm(1, 0, n, n, (int64_t)m_loc_, 1.0, 0.0, dV1_, (int64_t)ldv_, dV1_, (int64_t)ldv_, 0.0, 0.0,
                            dG_, (int64_t)ldg_, 1, splitk_ws_, splitk_ws_bytes_, stream_));
        qrt_.lap(stream_, 0);
        sum_triangle(dG_, n, false, true); // potrf reads the upper triangle
        qrt_.lap(stream_, 1);
        if (shifted)
        {
            const double scale = (sizeof(R) == 8) ? std::sqrt((double)N_) * 2.220446049250313e-16
                                                  : 10.0 * 1.1920928955078125e-07;
            CB2_KCHECK(KK::shift_abstrace(n, dG_, (int64_t)ldg_, scale * shift_boost, nullptr, stream_));
        }
        CB2_CHECK(cudaMemsetAsync(dInfo_, 0, sizeof(int), stream_));
        CB2_KCHECK(KK::potrf(n, dG_, (int64_t)ldg_, dInfo_, stream_));
        int info = 0;
        CB2_CHECK(cudaMemcpyAsync(&info, dInfo_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        qrt_.lap(stream_, 2);
        if (info != 0)
            return info;
        CB2_KCHECK(KK::trsm((int64_t)m_loc_, n, dG_, (int64_t)ldg_, dV1_, (int64_t)ldv_, dVs_, (int64_t)ldv_, trsm_ws_,
                            trsm_ws_bytes_, stream_));
        std::swap(dV1_, dVs_);
        qrt_.lap(stream_, 3);
        qrt_.rounds++;
        return 0;
    }
    // Householder fallback.  The reference runs a distributed panel factorisation (nccl/householder_qr.hpp, 3000
    // lines of scalar allreduces per column); the fallback is rare and the whole N x nc panel fits on one GPU
    // (C4: 2.7 GB), so here the column-layout pieces are all-gathered inside the grid column, every GPU factorises the
    // full-length copy with the single-GPU kernel (identical inputs -> identical results) and keeps its own rows.
    void householder()
    {
        if (hh_a_ == nullptr)
        {
            hh_a_ = alloc<T>(ldn_ * nc_);
            hh_q_ = alloc<T>(ldn_ * nc_);
            hh_ws_bytes_ = chase_b200_hhqr_ws_bytes((int64_t)N_, (int64_t)nc_, (int)sizeof(T));
            hh_ws_ = alloc<unsigned char>(hh_ws_bytes_);
        }
        v_to_full(dV1_, hh_a_, nc_);
        CB2_KCHECK(KK::hhqr((int64_t)N_, (int64_t)nc_, hh_a_, (int64_t)ldn_, hh_q_, (int64_t)ldn_, hh_ws_, hh_ws_bytes_,
                            stream_));
        full_to_v(hh_q_, dV1_, nc_);
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

    // S X on a distributed panel: negate the local rows whose global index lies in the lower half
    void flip_v(T* X, std::size_t k)
    {
        CB2_KCHECK(KK::scale_rows_map((int64_t)m_loc_, (int64_t)k, map_full2v_, (int64_t)(N_ / 2), X, (int64_t)ldv_, -1.0,
                                      stream_));
    }
    void flip_w(T* X, std::size_t k)
    {
        CB2_KCHECK(KK::scale_rows_map((int64_t)n_loc_, (int64_t)k, map_full2w_, (int64_t)(N_ / 2), X, (int64_t)ldw_, -1.0,
                                      stream_));
    }
    // W (row layout) <- H V for the pseudo-Hermitian H: H = S H^H S, so the A^H product of the Hermitian path serves
    void times_H_v2w(T* V, T* W, std::size_t k)
    {
        flip_v(V, k);
        hemm_v2w(T(1), V, T(0), W, k);
        flip_v(V, k);
        flip_w(W, k);
    }

    // Distributed rayleighRitz_v2 (reference linalg/internal/nccl/pseudo_hermitian_rayleighRitz.hpp; statement:
    // cpu/rayleighRitz.hpp:284-392) on Q = V1[:, locked_ ... locked_ + 2 block): the two N-long contractions are local
    // GEMMs + one allreduce each, the n x n part is replicated (and broadcast from one replica: identical decisions).
    void rr_pseudo(R* ritzv, std::size_t block)
    {
        const std::size_t nn = 2 * block;
        const int64_t n = (int64_t)nn, ldg = (int64_t)ldg_;
        T* Q = dV1_ + locked_ * ldv_;
        T* W1 = dW_[0] + locked_ * ldw_; // S H Q = A^H (S Q)   (row layout)
        T* W2 = dW_[1] + locked_ * ldw_; // Q                   (row layout)
        T* SQ = dVs_ + locked_ * ldv_;   // S Q                 (column layout)
        if (grid_.c > 1)
            comm_.broadcast(Q, ldv_ * nn, 0, comm_.row(), stream_);
        CB2_KCHECK(KK::lacpy((int64_t)m_loc_, n, Q, (int64_t)ldv_, SQ, (int64_t)ldv_, stream_));
        flip_v(SQ, nn);
        hemm_v2w(T(1), SQ, T(0), W1, nn);
        redistribute_v2w(Q, W2, nn);
        // G = Q^H S H Q
        CB2_KCHECK(KK::gemm(1, 0, n, n, (int64_t)n_loc_, 1.0, 0.0, W2, (int64_t)ldw_, W1, (int64_t)ldw_, 0.0, 0.0, dG_, ldg,
                            0, splitk_ws_, splitk_ws_bytes_, stream_));
        sum_triangle(dG_, n, false, false); // factorised by potrf: upper triangle
        // M0 = Q^H S Q (= I - 2 Q2^H Q2 for orthonormal Q)
        CB2_KCHECK(KK::gemm(1, 0, n, n, (int64_t)m_loc_, 1.0, 0.0, Q, (int64_t)ldv_, SQ, (int64_t)ldv_, 0.0, 0.0, dM_, ldg,
                            0, splitk_ws_, splitk_ws_bytes_, stream_));
        if (grid_.r > 1)
            comm_.allreduce_sum(dM_, ldg_ * nn, comm_.col(), stream_);
        if (grid_.c > 1)
            comm_.broadcast(dM_, ldg_ * nn, 0, comm_.row(), stream_);
        CB2_CHECK(cudaMemsetAsync(dInfo_, 0, sizeof(int), stream_));
        CB2_KCHECK(KK::potrf(n, dG_, ldg, dInfo_, stream_));
        int info = 0;
        CB2_CHECK(cudaMemcpyAsync(&info, dInfo_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
        CB2_CHECK(cudaStreamSynchronize(stream_));
        if (info != 0)
            throw std::runtime_error("chase_b200: Q^H S H Q is not positive definite in the pseudo-Hermitian RR "
                                     "(potrf info=" + std::to_string(info) + "): S H must be positive definite");
        CB2_CHECK(cudaMemsetAsync(dT_, 0, ldg_ * nn * sizeof(T), stream_));
        CB2_KCHECK(KK::shift_diag(n, dT_, ldg, 1.0, stream_));
        CB2_KCHECK(KK::trsm(n, n, dG_, ldg, dT_, ldg, dRinv_, ldg, trsm_ws_, trsm_ws_bytes_, stream_));
        // M = -R^-H M0 R^-1
        CB2_KCHECK(KK::gemm(0, 0, n, n, n, 1.0, 0.0, dM_, ldg, dRinv_, ldg, 0.0, 0.0, dT_, ldg, 0, nullptr, 0, stream_));
        CB2_KCHECK(KK::gemm(1, 0, n, n, n, -1.0, 0.0, dRinv_, ldg, dT_, ldg, 0.0, 0.0, dM_, ldg, 0, nullptr, 0,
                            stream_));
        std::vector<double> w(nn);
        int sweeps = 0;
        const int rc = KK::heev(n, dM_, ldg, dZ_, ldg, w.data(), heev_ws_, heev_ws_bytes_, &sweeps, stream_);
        if (rc < 0)
            throw std::runtime_error("chase_b200: Hermitian eigensolver failed in the pseudo-Hermitian RR (rc=" +
                                     std::to_string(rc) + ")");
        heev_sweeps_ += sweeps;
        if (rc > 0) // sweep limit reached: the decomposition is still usable, the residual check guards accuracy
            std::fprintf(stderr, "chase_b200: warning: Jacobi eigensolver stopped at its sweep limit (%d sweeps)\n", sweeps);
        for (std::size_t i = 0; i < nn; ++i)
            ritzv[i] = R(1.0) / (R)(-w[i]);
        CB2_KCHECK(KK::gemm(0, 0, n, (int64_t)block, n, 1.0, 0.0, dRinv_, ldg, dZ_, ldg, 0.0, 0.0, dT_, ldg, 0, nullptr,
                            0, stream_));
        CB2_KCHECK(KK::normalize_cols(n, (int64_t)block, dT_, ldg, stream_));
        CB2_KCHECK(KK::gemm(0, 0, (int64_t)m_loc_, (int64_t)block, n, 1.0, 0.0, Q, (int64_t)ldv_, dT_, ldg, 0.0, 0.0,
                            dV2_ + locked_ * ldv_, (int64_t)ldv_, 0, nullptr, 0, stream_));
        std::swap(dV1_, dV2_);
    }

    // Lanczos in the S H inner product on full-length replicated vectors (cpu/lanczos.hpp:332-516); H x = S A^H S x
    // is the only distributed step.
    void run_lanczos_pseudo(std::size_t M, std::size_t numvec, R* upperb, R* Theta, R* Tau, R* ritzV, bool multi)
    {
        flush_perm();
        resid_ready_ = false;
        const int nv = (int)numvec, m = (int)M;
        ensure_lanczos_buffers(M, numvec);
        T* v0 = lan_v_;
        T* v1 = lan_v_ + ldn_ * numvec;
        T* v2 = lan_v_ + 2 * ldn_ * numvec;
        const int64_t N = (int64_t)N_, half = (int64_t)(N_ / 2), ldn = (int64_t)ldn_;
        T* xloc = dVs_;
        T* ypart = dW_[2];
        auto matvec = [&](T* x, T* y)
        {
            CB2_KCHECK(KK::scale_rows(N - half, nv, x + half, ldn, -1.0, stream_));
            full_to_v(x, xloc, numvec);
            CB2_KCHECK(KK::scale_rows(N - half, nv, x + half, ldn, -1.0, stream_));
            CB2_KCHECK(KK::gemv_conjt((int64_t)m_loc_, (int64_t)n_loc_, dH_, (int64_t)lda_, xloc, (int64_t)ldv_, nv,
                                      ypart, (int64_t)ldw_, stream_));
            if (grid_.r > 1)
                comm_.allreduce_sum(ypart, ldw_ * numvec, comm_.col(), stream_);
            w_to_full(ypart, y, numvec);
            CB2_KCHECK(KK::scale_rows(N - half, nv, y + half, ldn, -1.0, stream_));
        };
        CB2_CHECK(cudaMemsetAsync(lan_d_, 0, M * numvec * sizeof(double), stream_));
        CB2_CHECK(cudaMemsetAsync(lan_e_, 0, M * numvec * sizeof(double), stream_));
        v_to_full(dV1_, v1, numvec);
        matvec(v1, v2);
        CB2_KCHECK(KK::lanczos_pseudo_norm(N, nv, -1, m, v1, v2, ldn, lan_e_, lan_rb_, stream_));
        for (int k = 0; k < m; ++k)
        {
            if (multi)
                full_to_v(v1 + (std::size_t)(nv - 1) * ldn_, dV1_ + (std::size_t)k * ldv_, 1);
            CB2_KCHECK(KK::lanczos_pseudo_step(N, nv, k, m, v0, v1, v2, ldn, lan_d_, lan_rb_, stream_));
            if (k == m - 1)
                break;
            T* t = v0;
            v0 = v1;
            v1 = v2;
            v2 = t;
            matvec(v1, v2);
            CB2_KCHECK(KK::lanczos_pseudo_norm(N, nv, k, m, v1, v2, ldn, lan_e_, lan_rb_, stream_));
        }
        if (multi)
            full_to_v(v1, dV1_, numvec);
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
            lan_v_ = alloc<T>(3 * ldn_ * numvec);
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

    // Lanczos with the numvec vectors replicated in full length on every rank: A v is the only distributed step
    // (local A_loc^H x, allreduce inside the grid column, all-gather inside the grid row); the fused vector update
    // of the single-GPU backend then runs identically everywhere.  Formulas: cpu/lanczos.hpp:45-209.
    void run_lanczos(std::size_t M, std::size_t numvec, R* upperb, R* Theta, R* Tau, R* ritzV, bool multi)
    {
        flush_perm();
        resid_ready_ = false;
        const int nv = (int)numvec, m = (int)M;
        ensure_lanczos_buffers(M, numvec);
        T* v0 = lan_v_;
        T* v1 = lan_v_ + ldn_ * numvec;
        T* v2 = lan_v_ + 2 * ldn_ * numvec;
        CB2_CHECK(cudaMemsetAsync(lan_d_, 0, M * numvec * sizeof(double), stream_));
        CB2_CHECK(cudaMemsetAsync(lan_e_, 0, M * numvec * sizeof(double), stream_));
        v_to_full(dV1_, v1, numvec);
        CB2_KCHECK(KK::normalize_cols((int64_t)N_, nv, v1, (int64_t)ldn_, stream_));
        T* xloc = dVs_;      // my rows of the current vectors (column layout)
        T* ypart = dW_[2];   // A_loc^H x  (row layout)
        for (int k = 0; k < m; ++k)
        {
            full_to_v(v1, xloc, numvec);
            if (multi) // V1[:, k] <- current vector of the LAST run (cpu/lanczos.hpp:85-88)
                CB2_KCHECK(KK::lacpy((int64_t)m_loc_, 1, xloc + (std::size_t)(nv - 1) * ldv_, (int64_t)ldv_,
                                     dV1_ + (std::size_t)k * ldv_, (int64_t)ldv_, stream_));
            CB2_KCHECK(KK::gemv_conjt((int64_t)m_loc_, (int64_t)n_loc_, dH_, (int64_t)lda_, xloc, (int64_t)ldv_, nv,
                                      ypart, (int64_t)ldw_, stream_));
            if (grid_.r > 1)
                comm_.allreduce_sum(ypart, ldw_ * numvec, comm_.col(), stream_);
            w_to_full(ypart, v2, numvec);
            CB2_KCHECK(KK::lanczos_step((int64_t)N_, nv, k, m, v0, v1, v2, (int64_t)ldn_, lan_d_, lan_e_, lan_rb_,
                                        stream_));
            if (k == m - 1)
                break;
            T* t = v0;
            v0 = v1;
            v1 = v2;
            v2 = t;
        }
        if (multi)
            full_to_v(v1, dV1_, numvec);
        std::vector<double> w, Z, rb(numvec);
        CB2_CHECK(cudaMemcpyAsync(rb.data(), lan_rb_, rb.size() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
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
    std::size_t ldvh_;
    R* ritzv_;
    b200::Grid2D grid_;
    b200::GridComm comm_;
    b200::Dist1D Dr_, Dc_;
    ChaseConfig<T> config_;
    cudaStream_t stream_ = nullptr;
    std::size_t m_loc_ = 0, n_loc_ = 0, lda_ = 0, ldv_ = 0, ldw_ = 0, ldn_ = 0, ldg_ = 0, gath_elems_ = 0;
    T *dH_ = nullptr, *dV1_ = nullptr, *dV2_ = nullptr, *dVs_ = nullptr, *dGath_ = nullptr, *dG_ = nullptr,
      *dZ_ = nullptr;
    T* dW_[4] = {nullptr, nullptr, nullptr, nullptr};
    T *dM_ = nullptr, *dRinv_ = nullptr, *dT_ = nullptr, *dFull_ = nullptr; // pseudo-Hermitian only
    T* dPack_ = nullptr; // packed triangle of a Gram / projected matrix (allreduce payload)
    T* dHl_ = nullptr;   // lo part of the TF32 split of an FP32 local block (A_loc^H V on the tcgen05 kernel)
    unsigned char* tf32_scratch_ = nullptr;
    std::size_t tf32_scratch_bytes_ = 0;
    unsigned char *dHw_ = nullptr, *wide_scratch_ = nullptr; // FP64 copy of an FP32 local block + panel scratch
    std::size_t wide_scratch_bytes_ = 0;
    bool wide_valid_ = false;
    double* ones_ = nullptr;
    T* dV0_ = nullptr; // device copy of this rank's rows of the reference start block (parity mode)
    unsigned char *heev_ws_ = nullptr, *trsm_ws_ = nullptr, *splitk_ws_ = nullptr, *hh_ws_ = nullptr;
    std::size_t heev_ws_bytes_ = 0, trsm_ws_bytes_ = 0, splitk_ws_bytes_ = 0, hh_ws_bytes_ = 0;
    T *hh_a_ = nullptr, *hh_q_ = nullptr; // full-length copies for the Householder fallback (allocated on first use)
    double *dTheta_ = nullptr, *dNorms_ = nullptr;
    int *dInfo_ = nullptr, *dIdx_ = nullptr;
    int64_t *map_v2w_ = nullptr, *map_v2full_ = nullptr, *map_w2full_ = nullptr, *map_full2v_ = nullptr,
            *map_full2w_ = nullptr, *diag_lin_ = nullptr;
    std::size_t ndiag_ = 0;
    T* lan_v_ = nullptr;
    double *lan_d_ = nullptr, *lan_e_ = nullptr, *lan_w_ = nullptr, *lan_Z_ = nullptr, *lan_rb_ = nullptr;
    std::size_t lan_nv_ = 0, lan_m_ = 0, lan_z_ = 0;
    std::vector<void*> allocs_;
    std::vector<R> resid_;
    std::vector<int> perm_;
    bool perm_dirty_ = false;
    std::size_t locked_ = 0;
    NextOp next_ = NextOp::bAc;
    bool resid_ready_ = false;
    std::size_t resid_block_ = 0;
    std::size_t lanczosIter_ = 0, numLanczos_ = 0;
    bool device_rng_ = false, is_sym_ = true, keep_device_matrix_ = false, matrix_on_device_ = false;
    std::string last_qr_;
    std::vector<std::string> qr_log_;
    std::size_t heev_sweeps_ = 0, hemm_cols_ = 0, swaps_ = 0, gathers_ = 0;
};

} // namespace Impl
} // namespace chase
