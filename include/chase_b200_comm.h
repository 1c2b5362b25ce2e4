/* chase_b200 — communicator bootstrap for the distributed entry points (p?chase_init_ ...).
 *
 * The reference builds its NCCL communicators inside chase::grid::MpiGrid2D from an MPI communicator: rank 0 calls
 * ncclGetUniqueId, MPI_Bcast ships the 128-byte id, every rank calls ncclCommInitRank
 * (reference grid/mpiGrid2D.hpp:449-485).  This library is MPI-free (one process per GPU, NCCL over NVLink for every
 * collective on the path), so the id exchange is the launcher's job and the "MPI_Comm*" argument of the reference's
 * distributed C interface carries the handle created here:
 *
 *     unsigned char id[CHASE_B200_COMM_ID_BYTES];
 *     if (rank == 0) chase_b200_comm_unique_id(id);
 *     MPI_Bcast(id, sizeof id, MPI_BYTE, 0, MPI_COMM_WORLD);      // or torch.distributed / a file / a socket
 *     MPI_Comm comm;                                              // typedef void* in include/mpi_shim.h
 *     chase_b200_comm_init(rank, nranks, id, local_device, &comm);
 *     pdchase_init_(&N, &nev, &nex, &m, &n, H, &ldh, V, ritzv, &dim0, &dim1, &major, &comm, &init);
 *
 * All functions return 0 on success, -1 on failure (message on stderr).
 */
#ifndef CHASE_B200_COMM_H
#define CHASE_B200_COMM_H

#ifdef __cplusplus
extern "C"
{
#endif

#define CHASE_B200_COMM_ID_BYTES 128

    int chase_b200_comm_unique_id(void* id_out);
    /* binds the calling process to CUDA device `device` (the reference uses the node-local rank,
       grid/mpiGrid2D.hpp:225-233) and joins the world communicator */
    int chase_b200_comm_init(int rank, int nranks, const void* id, int device, void** comm_out);
    int chase_b200_comm_free(void* comm);
    int chase_b200_comm_rank(void* comm);
    int chase_b200_comm_size(void* comm);
    /* index maps of the reference's layouts (distMatrix.hpp:44-67 numroc, :1992-2039 block layout):
       local extent of process p for a dimension of N split over nprocs; nb = 0 block layout, nb > 0 block-cyclic */
    long long chase_b200_local_size(long long N, int nprocs, long long nb, int p);
    /* global index of each local row/column of process p (out: local_size entries) */
    int chase_b200_global_indices(long long N, int nprocs, long long nb, int p, long long* out);
    /* Layout change as a gather map (what the distributed backend uploads once per direction): the pieces of a
       distribution (N over src_nprocs, src_nb) are stacked src_stride rows apart (an all-gather); out[t] is the stacked
       row that holds local row t of process pd of the destination distribution (dst_nprocs, dst_nb).  out has
       chase_b200_local_size(N, dst_nprocs, dst_nb, pd) entries.  Replaces the index arithmetic of the reference's
       redistributeImpl (linalg/distMatrix/distMultiVector.hpp:2817-2909). */
    int chase_b200_redistribution_map(long long N, int src_nprocs, long long src_nb, long long src_stride,
                                      int dst_nprocs, long long dst_nb, int pd, long long* out);
    /* Device-resident input: copies this rank's column-major device block (leading dimension *ld_src) into the active
       distributed solver of scalar type *type ('s','d','c','z') and marks it resident; subsequent p?chase_ calls do not
       read the host matrix pointer (which may then be NULL at init).  For matrices that must never exist on the host
       (BASELINE config C4: 28.8 GB per GPU). */
    int chase_b200_dist_load_device_matrix_(char* type, const void* src_dev, long long* ld_src);
    /* In-place variant for blocks that fit only once in HBM: the solver's own device buffer of the local block
       (column-major, leading dimension *ld_out elements) to be filled by the caller, then declared valid. */
    int chase_b200_dist_device_matrix_(char* type, void** ptr_out, long long* ld_out);
    int chase_b200_dist_mark_device_matrix_(char* type);
    /* Fortran callers hold the communicator as an INTEGER (the reference's MPI_Fint arguments of the `_f_` entry points,
       chase_c_interface.cpp:2425-2900): chase_b200_comm_c2f registers the handle and returns its index (>= 1),
       chase_b200_comm_f2c maps it back (NULL if unknown).  Counterparts of MPI_Comm_c2f / MPI_Comm_f2c. */
    int chase_b200_comm_c2f(void* comm);
    void* chase_b200_comm_f2c(int fcomm);
    /* grid coordinates of `rank` in a dim0 x dim1 grid with 'R'ow- or 'C'olumn-major rank order */
    int chase_b200_grid_coords(int dim0, int dim1, char grid_major, int rank, int* row_out, int* col_out);

#ifdef __cplusplus
}
#endif
#endif
