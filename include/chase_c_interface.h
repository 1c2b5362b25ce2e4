/* chase_b200 — the reference-compatible C / Fortran-callable interface of the
 * sequential (single-GPU) solver.  Names, argument order, by-pointer Fortran
 * convention and trailing underscore are those of the reference's
 * interface/chase_c_interface.h:17-41,177-180,217-239, so a C or Fortran
 * application linked against the reference's libchase_c (GPU build) can link
 * against libchase_b200.so instead.
 *
 *   ?chase_init_(N, nev, nex, H, ldh, V, ritzv, init)   reference :17-24
 *       constructs the process-global solver for that scalar type; the caller
 *       owns H (column-major, ldh >= N; read at every solve), V (N x (nev+nex),
 *       ld = N) and ritzv (nev+nex); V / ritzv may be NULL (then allocated
 *       internally, fetch results with ?chase_get_eigenpairs_).  *init = 1.
 *   ?chase_init_internal_(N, nev, nex, H, ldh, init)     reference :25-33
 *   ?chase_(deg, tol, mode, opt, qr)                     reference :38-41
 *       mode 'R' random / 'A' approximate start vectors; opt 'S' optimise
 *       degrees / 'N'; qr 'C' CholeskyQR / 'H' Householder.
 *   ?chase_finalize_(flag)                               reference :34-37
 *   ?chase_get_eigenpairs_(V, ld, ritzv)                 reference :177-180
 *   chase_set_*_ / chase_has_*_ / chase_get_version_     reference :217-239
 *
 * Not re-entrant and not thread-safe (one process-global solver per type), as
 * in the reference (interface/chase_c_interface.cpp:71-97).
 */
#ifndef CHASE_B200_C_INTERFACE_H
#define CHASE_B200_C_INTERFACE_H

#include <stddef.h>
#include <stdint.h>

#include "mpi_shim.h" /* MPI_Comm for MPI-less builds; include <mpi.h> first to use the real one */

#ifdef __cplusplus
extern "C"
{
#define CHASE_B200_CF float
#define CHASE_B200_CD double
#else
#define CHASE_B200_CF float _Complex
#define CHASE_B200_CD double _Complex
#endif

    void dchase_init_(int* N, int* nev, int* nex, double* H, int* ldh, double* V, double* ritzv, int* init);
    void schase_init_(int* N, int* nev, int* nex, float* H, int* ldh, float* V, float* ritzv, int* init);
    void cchase_init_(int* N, int* nev, int* nex, CHASE_B200_CF* H, int* ldh, CHASE_B200_CF* V, float* ritzv,
                      int* init);
    void zchase_init_(int* N, int* nev, int* nex, CHASE_B200_CD* H, int* ldh, CHASE_B200_CD* V, double* ritzv,
                      int* init);
    void dchase_init_internal_(int* N, int* nev, int* nex, double* H, int* ldh, int* init);
    void schase_init_internal_(int* N, int* nev, int* nex, float* H, int* ldh, int* init);
    void cchase_init_internal_(int* N, int* nev, int* nex, CHASE_B200_CF* H, int* ldh, int* init);
    void zchase_init_internal_(int* N, int* nev, int* nex, CHASE_B200_CD* H, int* ldh, int* init);

    /* Sequential pseudo-Hermitian (BSE) problems, H = [[A, B], [-conj(B), -conj(A)]] with S H positive definite
       (reference interface/chase_c_interface.h:42-58).  The caller allocates V with 2 (nev+nex) columns and ritzv
       with 2 (nev+nex) entries; on return the first nev columns / entries hold the smallest positive eigenpairs.
       Once a pseudo solver is initialised, ?chase_ / ?chase_get_eigenpairs_ / ?chase_finalize_ act on it, exactly as
       in the reference (interface/chase_c_interface.cpp:2204-2231). */
    void cchase_init_pseudo_(int* N, int* nev, int* nex, CHASE_B200_CF* H, int* ldh, CHASE_B200_CF* V, float* ritzv,
                             int* init);
    void cchase_init_pseudo_internal_(int* N, int* nev, int* nex, CHASE_B200_CF* H, int* ldh, int* init);
    void zchase_init_pseudo_(int* N, int* nev, int* nex, CHASE_B200_CD* H, int* ldh, CHASE_B200_CD* V, double* ritzv,
                             int* init);
    void zchase_init_pseudo_internal_(int* N, int* nev, int* nex, CHASE_B200_CD* H, int* ldh, int* init);
    void cchase_pseudo_(int* deg, float* tol, char* mode, char* opt, char* qr);
    void zchase_pseudo_(int* deg, double* tol, char* mode, char* opt, char* qr);

    void dchase_finalize_(int* flag);
    void schase_finalize_(int* flag);
    void cchase_finalize_(int* flag);
    void zchase_finalize_(int* flag);

    void dchase_(int* deg, double* tol, char* mode, char* opt, char* qr);
    void schase_(int* deg, float* tol, char* mode, char* opt, char* qr);
    void zchase_(int* deg, double* tol, char* mode, char* opt, char* qr);
    void cchase_(int* deg, float* tol, char* mode, char* opt, char* qr);

    void dchase_get_eigenpairs_(double* LEigsV, int* ld, double* ritzv);
    void schase_get_eigenpairs_(float* LEigsV, int* ld, float* ritzv);
    void cchase_get_eigenpairs_(CHASE_B200_CF* LEigsV, int* ld, float* ritzv);
    void zchase_get_eigenpairs_(CHASE_B200_CD* LEigsV, int* ld, double* ritzv);

    /* unified configuration setters: act on the first initialised solver in the order d, s, z, c */
    void chase_set_tol_(double* tol);
    void chase_set_deg_(int* deg);
    void chase_set_max_deg_(int* max_deg);
    void chase_set_deg_extra_(int* deg_extra);
    void chase_set_max_iter_(int* max_iter);
    void chase_set_lanczos_iter_(int* lanczos_iter);
    void chase_set_num_lanczos_(int* num_lanczos);
    void chase_set_approx_(int* flag);
    void chase_set_opt_(int* flag);
    void chase_set_cholqr_(int* flag);
    void chase_enable_sym_check_(int* flag);
    void chase_set_decaying_rate_(float* decaying_rate);
    void chase_set_cluster_aware_degrees_(int* flag);
    void chase_set_upperb_scale_rate_(float* upperb_scale_rate);

    void chase_get_version_(char* version, int* len);
    void chase_has_cuda_(int* flag);
    void chase_has_nccl_(int* flag);
    void chase_has_scalapack_(int* flag);
    void chase_has_mpi_(int* flag);
    void chase_print_config_(void);

    /* ---- distributed (multi-GPU) entry points: reference interface/chase_c_interface.h:61-195 -------------------
       One process per GPU on a dim0 x dim1 grid (dim0 >= dim1; grid_major 'R' | 'C' gives the rank order of
       MPI_Cart_create, grid/mpiGrid2D.hpp:402-430).  H / V are the LOCAL pieces: H is m x n with
       m = local rows, n = local columns of the block layout (distMatrix.hpp:1992-2039) or of the block-cyclic layout
       (numroc, distMatrix.hpp:44-67; irsrc = icsrc = 0); V holds the m local rows of the nev+nex vectors (ld = m).
       *comm is the handle made by chase_b200_comm_init (include/chase_b200_comm.h). */
#define CHASE_B200_DIST_API(X, CT, RT)                                                                                 \
    void p##X##chase_init_(int* N, int* nev, int* nex, int* m, int* n, CT* H, int* ldh, CT* V, RT* ritzv, int* dim0,  \
                           int* dim1, char* grid_major, MPI_Comm* comm, int* init);                                    \
    void p##X##chase_init_internal_(int* N, int* nev, int* nex, int* m, int* n, CT* H, int* ldh, int* dim0, int* dim1,\
                                    char* grid_major, MPI_Comm* comm, int* init);                                      \
    void p##X##chase_init_blockcyclic_(int* N, int* nev, int* nex, int* mbsize, int* nbsize, CT* H, int* ldh, CT* V,  \
                                       RT* ritzv, int* dim0, int* dim1, char* grid_major, int* irsrc, int* icsrc,      \
                                       MPI_Comm* comm, int* init);                                                     \
    void p##X##chase_init_blockcyclic_internal_(int* N, int* nev, int* nex, int* mbsize, int* nbsize, CT* H, int* ldh, \
                                                int* dim0, int* dim1, char* grid_major, int* irsrc, int* icsrc,        \
                                                MPI_Comm* comm, int* init);                                            \
    void p##X##chase_(int* deg, RT* tol, char* mode, char* opt, char* qr);                                             \
    void p##X##chase_finalize_(int* flag);                                                                             \
    void p##X##chase_get_eigenpairs_(CT* V, int* ld, RT* ritzv);                                                       \
    void p##X##chase_get_resid_(RT* resid);
    CHASE_B200_DIST_API(d, double, double)
    CHASE_B200_DIST_API(s, float, float)
    CHASE_B200_DIST_API(z, CHASE_B200_CD, double)
    CHASE_B200_DIST_API(c, CHASE_B200_CF, float)

    /* Distributed pseudo-Hermitian (BSE) problems (reference interface/chase_c_interface.h:105-123, 158-175): same
       arguments as the Hermitian initialisers; V holds the m local rows of 2 (nev+nex) vectors and ritzv 2 (nev+nex)
       entries.  While a pseudo solver exists, p?chase_ / p?chase_get_eigenpairs_ / p?chase_finalize_ act on it. */
#define CHASE_B200_DIST_PSEUDO_API(X, CT, RT)                                                                          \
    void p##X##chase_init_pseudo_(int* N, int* nev, int* nex, int* m, int* n, CT* H, int* ldh, CT* V, RT* ritzv,      \
                                  int* dim0, int* dim1, char* grid_major, MPI_Comm* comm, int* init);                  \
    void p##X##chase_init_pseudo_internal_(int* N, int* nev, int* nex, int* m, int* n, CT* H, int* ldh, int* dim0,    \
                                           int* dim1, char* grid_major, MPI_Comm* comm, int* init);                    \
    void p##X##chase_init_pseudo_blockcyclic_(int* N, int* nev, int* nex, int* mbsize, int* nbsize, CT* H, int* ldh,  \
                                              CT* V, RT* ritzv, int* dim0, int* dim1, char* grid_major, int* irsrc,    \
                                              int* icsrc, MPI_Comm* comm, int* init);                                  \
    void p##X##chase_init_pseudo_blockcyclic_internal_(int* N, int* nev, int* nex, int* mbsize, int* nbsize, CT* H,   \
                                                       int* ldh, int* dim0, int* dim1, char* grid_major, int* irsrc,   \
                                                       int* icsrc, MPI_Comm* comm, int* init);
    CHASE_B200_DIST_PSEUDO_API(z, CHASE_B200_CD, double)
    CHASE_B200_DIST_PSEUDO_API(c, CHASE_B200_CF, float)

    /* Matrix file I/O (reference interface/chase_c_interface.h:196-214): raw column-major N x N binary files, no
       header (Matrix::readFromBinaryFile / saveToBinaryFile, linalg/matrix/matrix.hpp:276-352).  They act on the host
       matrix buffer given at init (every rank reads / writes the pieces of its own local block); the next solve uploads
       it.  The un-prefixed readHam names are aliases, as in the reference. */
    void pschase_wrtHam_(const char* filename);
    void pdchase_wrtHam_(const char* filename);
    void pcchase_wrtHam_(const char* filename);
    void pzchase_wrtHam_(const char* filename);
    void pschase_readHam_(const char* filename);
    void pdchase_readHam_(const char* filename);
    void pcchase_readHam_(const char* filename);
    void pzchase_readHam_(const char* filename);
    void schase_readHam_(const char* filename);
    void dchase_readHam_(const char* filename);
    void cchase_readHam_(const char* filename);
    void zchase_readHam_(const char* filename);

    /* ---- chase_b200 additions (introspection for parity tests and benchmarks) ---- */
    /* residuals of the last solve (nev+nex values, ordered like ritzv) */
    void dchase_get_resid_(double* resid);
    void schase_get_resid_(float* resid);
    void cchase_get_resid_(float* resid);
    void zchase_get_resid_(double* resid);
    /* out[0..15]: iterations, filtered_vecs, hemm_calls, swaps, t_all, t_initvecs, t_lanczos, t_filter, t_qr, t_rr,
       t_resid, gflop_filter, gflop_total, heev_sweeps, gather_passes, error_flag  (seconds / GFLOP, last solve) */
    void chase_b200_get_stats_(double* out, int* n);
    /* record the ChaseBase call trace of subsequent solves (format of oracle/ref_driver.cpp) */
    void chase_b200_trace_enable_(int* flag);
    /* copies the '\n'-joined trace of the last solve; returns the byte count needed (without NUL) */
    size_t chase_b200_trace_copy_(char* buf, size_t cap);
    /* '\n'-joined CholQR variants chosen during the last solve */
    size_t chase_b200_qr_log_copy_(char* buf, size_t cap);
    /* host-side start block used by initVecs in parity mode: the reference CPU backend's mt19937(1337) +
       normal_distribution stream (reference Impl/chase_cpu/chase_cpu.hpp:296-309), N x m column-major */
    void chase_b200_start_vectors_d(int64_t N, int64_t m, double* V, int64_t ldv);
    void chase_b200_start_vectors_s(int64_t N, int64_t m, float* V, int64_t ldv);
    void chase_b200_start_vectors_z(int64_t N, int64_t m, CHASE_B200_CD* V, int64_t ldv);
    void chase_b200_start_vectors_c(int64_t N, int64_t m, CHASE_B200_CF* V, int64_t ldv);
    /* message of the last failed init/solve ("" if none); stats[15] is 1 after a failed solve */
    size_t chase_b200_last_error_copy_(char* buf, size_t cap);
    /* Symmetric tridiagonal eigen-decomposition on the host (implicit QL; w ascending, Z n x n column-major): what the
       Lanczos bound estimation uses beyond the on-device solver's 48 steps (reference: host LAPACK ?stemr,
       linalg/internal/cuda/lanczos.hpp:270-299).  Returns 0 or the 1-based index of a non-converged eigenvalue. */
    int chase_b200_tridiag_eig_host(int n, const double* d, const double* e, double* w, double* Z);
    int chase_b200_device_sync(void);
    /* flag != 0: subsequent solves do NOT re-upload the host matrix H when a copy from an earlier solve is already
       on the device (default 0 = the reference's behaviour: H is re-read at every solve, chase_gpu.hpp:536) */
    void chase_b200_set_matrix_resident_(int* flag);
    /* flag != 0: random start vectors come from the on-device Philox generator (what the reference GPU backend
       does with cuRAND, chase_gpu.hpp:509-533); 0 (default): the reference CPU backend's mt19937 stream, which is
       what makes iteration counts identical to the reference CPU solver */
    void chase_b200_set_device_rng_(int* flag);
    /* flag != 0: double-precision problems (d, z; sequential solver) filter in single precision -- on the tcgen05
       kind::tf32 kernel, FP32-accurate -- while the smallest residual of the wanted, unlocked pairs is above 1e-3, and in
       double precision afterwards: the reference's compile-time option ENABLE_MIXED_PRECISION
       (Impl/pchase_gpu/pchase_gpu.hpp:785-881), here a run-time switch (also CHASE_B200_MIXED_PRECISION=1).  Off by
       default, like the reference: the degree schedule of a mixed-precision run differs from the double-precision one.
       chase_b200_last_sp_filter_cols_: matrix-vector products the last solve did in single precision. */
    void chase_b200_set_mixed_precision_(int* flag);
    double chase_b200_last_sp_filter_cols_(void);

#ifdef __cplusplus
}
#endif
#endif
