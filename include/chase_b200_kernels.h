/* chase_b200 — kernel-level C ABI (device pointers in, status out).
 *
 * This is the thin layer the C++ host orchestration (chase_b200/host) calls;
 * every entry point is a hand-written sm_100a CUDA kernel launcher and stands
 * in for one vendor-library call (or helper kernel) of the reference's GPU
 * backends.  No cuBLAS / cuSOLVER / cuRAND is linked.
 *
 * Conventions
 *   - one function per ChASE value type, suffix _s/_d/_c/_z
 *     (float, double, complex<float>, complex<double>; complex = interleaved
 *     re,im exactly like std::complex / the reference);
 *   - all matrices column-major with an explicit leading dimension (elements);
 *   - complex scalars cross the ABI as (re, im) doubles;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - no allocation inside: workspaces are passed in;
 *   - return 0 = ok, <0 = CUDA / argument error (message on stderr),
 *     >0 = numerical info where stated (LAPACK convention).
 */
#ifndef CHASE_B200_KERNELS_H
#define CHASE_B200_KERNELS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define CHASE_B200_KERNEL_API(X)                                                                                       \
    /* C <- alpha op(A) op(B) + beta C; ta/tb: 0 = N, 1 = conjugate transpose; uplo: 0 all, 1 upper tiles only,       \
       2 lower tiles only. Replaces cublasTgemm / cublasTsyherk                                                        \
       (reference linalg/internal/cuda/rayleighRitz.hpp:125-214, cuda/cholqr.hpp:110-117). */                         \
    int chase_b200_gemm_##X(int ta, int tb, int64_t M, int64_t N, int64_t K, double alpha_re, double alpha_im,        \
                            const void* A, int64_t lda, const void* B, int64_t ldb, double beta_re, double beta_im,   \
                            void* C, int64_t ldc, int uplo, void* ws, size_t ws_bytes, void* stream);                 \
    /* Filter step: C <- alpha (A - shift I) B + beta C on an n x k panel (A is n x n Hermitian and is NOT            \
       modified).  If theta != NULL (device, k doubles) the shift is per column: C <- alpha (A B - B diag(theta))     \
       + beta C, which with alpha=1, beta=0 is the residual block A V - V Theta.                                      \
       Replaces Shift + cublasTgemm in ChASEGPU::HEMM (reference Impl/chase_gpu/chase_gpu.hpp:599-603, 656-678)       \
       and the GEMM of cuda::residuals (linalg/internal/cuda/residuals.hpp:92-110). */                                \
    int chase_b200_hemm_##X(int64_t n, int64_t k, double alpha_re, double alpha_im, const void* A, int64_t lda,       \
                            const void* B, int64_t ldb, double beta_re, double beta_im, void* C, int64_t ldc,         \
                            double shift, const double* theta, void* stream);                                         \
    /* Distributed filter step on a local block: C(M x k) <- alpha op(A) B(K x k) + beta C, op(A) = A (stored M x K)  \
       for ta == 0 or A^H (stored K x M) for ta == 1.  Same TMA + DMMA kernel as chase_b200_hemm.  Replaces the two   \
       cublasTgemm of cuda_nccl::MatrixMultiplyMultiVectors (reference linalg/internal/nccl/hemm.hpp:325-332,         \
       382-389). */                                                                                                    \
    int chase_b200_hemm_rect_##X(int ta, int64_t M, int64_t K, int64_t k, double alpha_re, double alpha_im,           \
                                 const void* A, int64_t lda, const void* B, int64_t ldb, double beta_re,              \
                                 double beta_im, void* C, int64_t ldc, void* stream);                                 \
    /* Upper Cholesky G = R^H R in place (strict lower part untouched). *info_dev (device int, must be zeroed by      \
       the caller) receives the 1-based index of the first non-positive pivot, LAPACK ?potrf convention.              \
       Replaces cusolverDnTpotrf (reference cuda/cholqr.hpp:119-125). */                                              \
    int chase_b200_potrf_##X(int64_t n, void* G, int64_t ldg, int* info_dev, void* stream);                           \
    /* X <- V R^-1 with R upper triangular n x n (V is overwritten with intermediates).  ws: at least                 \
       chase_b200_trsm_ws_bytes(n) bytes.  Replaces cublasTtrsm(RIGHT, UPPER, N) (reference cuda/cholqr.hpp:127-132) \
     */                                                                                                                \
    int chase_b200_trsm_##X(int64_t rows, int64_t n, const void* R, int64_t ldr, void* V, int64_t ldv, void* Xout,    \
                            int64_t ldx, void* ws, size_t ws_bytes, void* stream);                                    \
    /* G_ii += scale * sum_i |G_ii| ; the shift is also stored to *shift_out_dev (device double, may be NULL).        \
       Replaces absTrace + shiftDiagonalFromDeviceShift (reference cuda/cholqr.hpp:391-420). */                       \
    int chase_b200_shift_abstrace_##X(int64_t n, void* G, int64_t ldg, double scale, double* shift_out_dev,           \
                                      void* stream);                                                                   \
    /* Hermitian eigen-decomposition of the n x n matrix G (LOWER triangle referenced): eigenvalues ascending to      \
       w_host (host, n doubles), eigenvectors to Z (device).  Synchronises the stream.  ws: at least                   \
       chase_b200_heev_ws_bytes(n, is_complex).  *sweeps (host, may be NULL) gets the Jacobi sweep count.             \
       Returns >0 if not converged.  Replaces cusolverDnTheevd (reference cuda/rayleighRitz.hpp:169-209). */          \
    int chase_b200_heev_##X(int64_t n, const void* G, int64_t ldg, void* Z, int64_t ldz, double* w_host, void* ws,    \
                            size_t ws_bytes, int* sweeps, void* stream);                                              \
    /* out[j] = ||X[:, j]||_2 (take_sqrt != 0) or its square.  Replaces residual_gpu's reduction                      \
       (reference linalg/internal/cuda/residuals.cu:113-296). */                                                      \
    int chase_b200_colnorms_##X(int64_t rows, int64_t cols, const void* Xm, int64_t ldx, double* out_dev,             \
                                int take_sqrt, void* stream);                                                         \
    /* dst <- src (rows x cols).  Replaces cuda::t_lacpy('A') (reference cuda/lacpy.cu:62-496). */                    \
    int chase_b200_lacpy_##X(int64_t rows, int64_t cols, const void* src, int64_t lds, void* dst, int64_t ldd,        \
                             void* stream);                                                                            \
    /* Packed triangle of the n x n matrix G (LAPACK 'U' / 'L' column-major packing, n (n+1) / 2 elements) and back:  \
       the payload of the Gram / projected-matrix allreduces.  Replaces cuda::extractUpperTriangular /                \
       unpackUpperTriangular (reference cuda/lacpy.cu:837-, 956-; nccl/cholqr.hpp:152-157). */                        \
    int chase_b200_tri_pack_##X(int64_t n, const void* G, int64_t ldg, void* P, int lower, void* stream);             \
    int chase_b200_tri_unpack_##X(int64_t n, const void* P, void* G, int64_t ldg, int lower, void* stream);           \
    /* dst[:, dcols[t]] <- src[:, scols[t]], t < cnt (index arrays on the device; src != dst).  Batched form of       \
       the reference's per-call cublasTswap (Impl/chase_gpu/chase_gpu.hpp:1000-1006). */                              \
    int chase_b200_gather_cols_##X(int64_t rows, int cnt, const int* scols_dev, const int* dcols_dev,                 \
                                   const void* src, int64_t lds, void* dst, int64_t ldd, void* stream);               \
    /* dst[i, :] <- src[src_row[i], :] for every i < rows with src_row[i] >= 0 (device int64 map).  piece_stride > 0: \
       src is an all-gathered stack of (lds x cols) panels piece_stride elements apart and src_row = piece * lds +     \
       row.  The one layout primitive of the distributed backend; replaces the broadcast loops of redistributeImpl    \
       (reference linalg/distMatrix/distMultiVector.hpp:2817-2909). */                                                \
    int chase_b200_gather_rows_##X(int64_t rows, int64_t cols, const int64_t* src_row_dev, const void* src,           \
                                   int64_t lds, int64_t piece_stride, void* dst, int64_t ldd, void* stream);          \
    /* C[:, j] += g * gvec[j] * E[:, j] (gvec: device, cols doubles).  With g = -1, gvec = theta it turns A V into    \
       the residual block A V - V Theta (reference cuda/residuals.cu:113-251 does the subtraction inside its norm     \
       kernel). */                                                                                                     \
    int chase_b200_axpy_cols_##X(int64_t rows, int64_t cols, const double* gvec_dev, double g_re, double g_im,        \
                                 const void* E, int64_t lde, void* Cm, int64_t ldc, void* stream);                    \
    /* A[lin[t]] += c, t < cnt: local copies of global diagonal entries (reference chase_shift_mgpu_matrix,           \
       cuda/shiftDiagonal.cu:100-150). */                                                                              \
    int chase_b200_shift_diag_list_##X(int64_t cnt, const int64_t* lin_dev, void* A, double c, void* stream);         \
    /* Philox N(0,1) fill addressed by global element index grow[i] + j * nglobal (layout-independent start block;   \
       the reference seeds cuRAND per grid row instead, Impl/pchase_gpu/pchase_gpu.hpp:652-688). */                   \
    int chase_b200_rng_normal_rows_##X(int64_t rows, int64_t cols, const int64_t* grow_dev, int64_t nglobal,          \
                                       void* Xm, int64_t ldx, uint64_t seed, void* stream);                           \
    /* Y[:, v] <- A^H X[:, v], v < nv (A is rows x cols).  HBM-bound Lanczos product; replaces the 4-column           \
       cublasTgemm(OP_C) of cuda::lanczos (reference linalg/internal/cuda/lanczos.hpp:178-186). */                    \
    int chase_b200_gemv_conjt_##X(int64_t rows, int64_t cols, const void* A, int64_t lda, const void* Xm,             \
                                  int64_t ldx, int nv, void* Y, int64_t ldy, void* stream);                           \
    /* One fused Lanczos step for nv vectors (dot, two axpys, norm, scale; scalars stay on the device).               \
       Replaces the batched dot/axpy/norm kernels of reference cuda/lanczos_kernels.cu. */                             \
    int chase_b200_lanczos_step_##X(int64_t rows, int nv, int k, int M, const void* v0, const void* v1, void* v2,     \
                                    int64_t ld, double* d_dev, double* e_dev, double* rbeta_dev, void* stream);       \
    /* X[:, j] /= ||X[:, j]||  (reference cuda/lanczos.hpp:112-131) */                                                \
    int chase_b200_normalize_cols_##X(int64_t rows, int64_t cols, void* Xm, int64_t ldx, void* stream);               \
    /* Philox4x32-10 + Box-Muller N(0,1) fill (production-mode start vectors; reference                               \
       cuda/random_normal_distribution.cu:21-93 uses cuRAND Philox). */                                               \
    int chase_b200_rng_normal_##X(int64_t rows, int64_t cols, void* Xm, int64_t ldx, uint64_t seed, void* stream);    \
    /* *bad_dev (device, zeroed by caller) += #entries of the lower triangle with |A_ij - conj(A_ji)| >               \
       tol (|A_ij| + |A_ji|).  Replaces checkSymmetryEasy (reference chase_gpu.hpp:463-470). */                       \
    int chase_b200_herm_check_##X(int64_t n, const void* A, int64_t lda, double tol, unsigned long long* bad_dev,     \
                                  void* stream);                                                                       \
    /* A_ii += c (real part).  Replaces chase_shift_matrix (reference cuda/shiftDiagonal.cu:23-50). */                \
    int chase_b200_shift_diag_##X(int64_t n, void* A, int64_t lda, double c, void* stream);                           \
    /* Mirror one triangle onto the other (symOrHermMatrix, reference chase_gpu.hpp:472-505). */                      \
    int chase_b200_herm_mirror_##X(int64_t n, void* A, int64_t lda, int from_upper, void* stream);                    \
    /* Householder QR of the rows x n matrix A (rows >= n): Q <- orthonormal factor (LAPACK ?geqrf + ?orgqr/?ungqr       \
       conventions: R_jj real, sign -sign(Re x_j)), A <- R (upper triangle) and the reflectors.  ws: at least            \
       chase_b200_hhqr_ws_bytes(rows, n, sizeof(element)).  Replaces cusolverDnTgeqrf + cusolverDnTorgqr/ungqr of        \
       cuda::houseHoulderQR (reference linalg/internal/cuda/cholqr.hpp:524-556), the fallback of ChASEGPU::QR when       \
       CholQR breaks down or qr == 'H' (Impl/chase_gpu/chase_gpu.hpp:822-919). */                                        \
    int chase_b200_hhqr_##X(int64_t rows, int64_t n, void* A, int64_t lda, void* Q, int64_t ldq, void* ws,            \
                            size_t ws_bytes, void* stream);                                                           \
    /* X[0:nrows, 0:cols] *= a (X points at the first row to scale).  a = -1 on rows [N/2, N) is S X of the           \
       pseudo-Hermitian path (reference cuda/flipSign.cu, chase_gpu.hpp:752-782); a = 1e-3 is the start-vector        \
       damping of the lower block (chase_gpu.hpp:518-529). */                                                         \
    int chase_b200_scale_rows_##X(int64_t nrows, int64_t cols, void* Xm, int64_t ldx, double a, void* stream);        \
    /* Same on a row-split (distributed) panel: X[i, :cols] *= a for the local rows whose global index               \
       grow_dev[i] >= g0 (device int64 map, the rows' global indices). */                                             \
    int chase_b200_scale_rows_map_##X(int64_t rows, int64_t cols, const int64_t* grow_dev, int64_t g0, void* Xm,      \
                                      int64_t ldx, double a, void* stream);                                           \
    /* K-conjugate partner vectors dst[:, j] = conj([src[rows/2:, j]; src[:rows/2, j]]), rows even, src/dst           \
       disjoint.  Replaces the two lacpy + conjugate kernel of ChASEGPU::ApplyKconjugate (chase_gpu.hpp:718-742). */  \
    int chase_b200_kconj_##X(int64_t rows, int64_t cols, const void* src, int64_t lds, void* dst, int64_t ldd,        \
                             void* stream);                                                                            \
    /* S-inner-product Lanczos (pseudo-Hermitian H, v2 = H v1 on entry) for nv vectors, scalars on the device:        \
       norm: beta = sqrt(Re <v1, S v2>), v1 /= beta, v2 /= beta, e[ke] = beta if ke >= 0, bnorm[v] = beta;            \
       step: alpha = <v2, S v2>, v2 -= alpha v1, d[k] = alpha, and v2 -= beta v0 for 0 < k < M-1.                     \
       Replace the dot/scal/axpy kernels of reference cuda/lanczos.hpp:547-785 (CPU: cpu/lanczos.hpp:332-516). */     \
    int chase_b200_lanczos_pseudo_norm_##X(int64_t rows, int nv, int ke, int M, void* v1, void* v2, int64_t ld,       \
                                           double* e_dev, double* bnorm_dev, void* stream);                           \
    int chase_b200_lanczos_pseudo_step_##X(int64_t rows, int nv, int k, int M, const void* v0, const void* v1,        \
                                           void* v2, int64_t ld, double* d_dev, const double* bnorm_dev,              \
                                           void* stream);

    CHASE_B200_KERNEL_API(s)
    CHASE_B200_KERNEL_API(d)
    CHASE_B200_KERNEL_API(c)
    CHASE_B200_KERNEL_API(z)

    /* Batched symmetric tridiagonal eigen-solve (n <= 48) fully on the device: matrix b has diagonal
       d[b*ldde + 0..n-1] and off-diagonal e[b*ldde + 0..n-2]; w[b*n + i] ascending, Z[b*n*n + i + j*n].
       Replaces the host LAPACK ?stemr of the reference (linalg/internal/cuda/lanczos.hpp:270-299). */
    int chase_b200_tridiag_eig(int n, int batch, const double* d_dev, const double* e_dev, int ldde, double* w_dev,
                               double* Z_dev, void* stream);

    /* FP32 storage on the FP64 TMA pipeline: register an FP64 copy A_wide (same ld, caller-owned, device) of the
       float / complex<float> matrix A plus a scratch area; chase_b200_hemm[_rect]_{s,c} calls whose A is this pointer
       then widen their panels into the scratch, run the TMA + DMMA kernel of the wide type on A_wide and narrow the
       result back (needs scratch_bytes >= (roundup16(K) + roundup16(M)) * k * sizeof(wide element)).
       chase_b200_widen_sync(type 's' | 'c') re-converts A -> A_wide after A changed.  Host-side registry, one entry
       per matrix; unregister before freeing. */
    int chase_b200_widen_register(const void* A, void* A_wide, int64_t ld, int64_t rows, int64_t cols, void* scratch,
                                  size_t scratch_bytes);
    int chase_b200_widen_unregister(const void* A);
    int chase_b200_widen_sync(char type, const void* A, void* stream);
    /* Single-precision value types on the 5th-generation tensor cores (tcgen05.mma kind::tf32, TMEM accumulators,
       TMA-fed; csrc/hemm_tf32.cuh):  C(M x k) <- alpha S (A^s)^H S B + beta C - alpha shift_j B  with the stored matrix
       A^s K x M column-major, 3 (or 4) TF32 partial products per operand pair.  Alo = the lo part of the TF32 split of
       A^s (same shape and ld; chase_b200_tf32_sync writes it), sflip > 0 negates rows >= sflip of the panel and of the
       product (pseudo-Hermitian H = S H^H S), scratch >= chase_b200_hemm_tf32_scratch_bytes(K, k, sizeof element).
       Replaces cublasSgemm / cublasCgemm (reference external/cublaspp/cublaspp.hpp:563, 623).  Returns -5 when the
       shape / alignment is not supported (callers fall back to the generic kernel). */
    int chase_b200_hemm_tf32_s(int64_t M, int64_t K, int64_t k, double are, double aim, const void* A, const void* Alo,
                               int64_t lda, const void* B, int64_t ldb, double bre, double bim, void* C, int64_t ldc,
                               double shift, const double* theta, int64_t sflip, int terms, void* scratch,
                               size_t scratch_bytes, void* stream);
    int chase_b200_hemm_tf32_c(int64_t M, int64_t K, int64_t k, double are, double aim, const void* A, const void* Alo,
                               int64_t lda, const void* B, int64_t ldb, double bre, double bim, void* C, int64_t ldc,
                               double shift, const double* theta, int64_t sflip, int terms, void* scratch,
                               size_t scratch_bytes, void* stream);
    size_t chase_b200_hemm_tf32_scratch_bytes(int64_t K, int64_t k, int elem_bytes);
    /* Registry: chase_b200_hemm[_rect]_{s,c} calls whose A is a registered pointer run on the kernel above.
       kind 0: A Hermitian (A B = A^H B), 1: pseudo-Hermitian (S A Hermitian), 2: general block (only op(A) = A^H).
       chase_b200_tf32_sync(type 's' | 'c') recomputes Alo after A changed; chase_b200_tf32_set_terms(3 | 4) selects
       the number of partial products of the following calls (4: A_lo B_lo included, for the RR / residual products). */
    int chase_b200_tf32_register(const void* A, void* Alo, int64_t ld, int64_t rows, int64_t cols, int kind,
                                 void* scratch, size_t scratch_bytes);
    int chase_b200_tf32_unregister(const void* A);
    int chase_b200_tf32_sync(char type, const void* A, void* stream);
    /* same for `cnt` elements given by linear element indices (device array): after an in-place diagonal shift */
    int chase_b200_tf32_sync_list(char type, const void* A, int64_t cnt, const int64_t* lin_dev, void* stream);
    void chase_b200_tf32_set_terms(int terms);
    /* Precision change of a rows x cols column-major array, (from, to) in {'d'->'s', 's'->'d', 'z'->'c', 'c'->'z'}:
       the copies behind the mixed-precision filter (reference linalg/internal/cuda/precision_conversion.cu:20-55). */
    int chase_b200_convert(char from, char to, int64_t rows, int64_t cols, const void* src, int64_t lds, void* dst,
                           int64_t ldd, void* stream);
    /* The filter HEMM keeps stream-K hand-over slots per (device, stream); the owner of a stream releases them before
       cudaStreamDestroy (a later stream with the same handle value must not inherit them). */
    int chase_b200_stream_release(void* stream);
    size_t chase_b200_trsm_ws_bytes(int64_t n, int elem_bytes);
    size_t chase_b200_hhqr_ws_bytes(int64_t rows, int64_t n, int elem_bytes);
    size_t chase_b200_heev_ws_bytes(int64_t n, int is_complex);
    /* Which HEMM implementation handles this shape: 0 = generic DMMA tiles, 1 = TMA-fed DMMA pipeline. */
    int chase_b200_hemm_path(int dtype_code, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc);
    /* FP64 tensor-pipe microbenchmark: runs `iters` register-resident DMMA.8x8x4 per warp on every SM and
       returns the achieved FLOP/s (the roofline denominator for the filter; MEASURED_PEAKS.json has no FP64). */
    double chase_b200_dmma_peak(int iters, void* stream);
    /* Virtual -> raster tile index of the stream-K filter HEMM (host copy of the device function, for tests):
       a bijection on [0, ntiles) that places the s-th tile of every CTA next to each other (L2 sharing of A). */
    long long chase_b200_hemm_tile_remap(long long v, long long ntiles, long long nctas);
    /* Host replay of the work list of one CTA of the filter HEMM (stream-K / hybrid schedule; same code as the
       kernel): out[3 i ..] = raster tile, first k-block, one-past-last k-block of part i.  Returns the number of
       parts, -1 if cta is outside the grid.  For tests of coverage and of the head/tail hand-over order. */
    long long chase_b200_hemm_walk(long long ntiles, long long nkt, int sms, int cta, long long* out, long long cap);
    /* Number of product kernels launched by this library so far (process-wide; bench.py's gpu_launches). */
    unsigned long long chase_b200_launch_count(void);
    /* Per-launch CUDA-event timing of the filter HEMM kernel on its launching stream.  enable(1) resets and
       starts recording, read() synchronises the device and returns {launches, total ms, total algorithmic
       flop (2 f n^2 k per launch), longest single launch ms}. */
    int chase_b200_hemm_profile_enable(int on);
    int chase_b200_hemm_profile_read(double* out4);
    const char* chase_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
