/* Minimal stand-in for <mpi.h> so that the reference's distributed prototypes (interface/chase_c_interface.h:61-195,
 * which take an MPI_Comm*) compile unchanged in MPI-less builds.  The handle is created by chase_b200_comm_init
 * (include/chase_b200_comm.h).  If a real <mpi.h> was included first, nothing is defined here. */
#ifndef CHASE_B200_MPI_SHIM_H
#define CHASE_B200_MPI_SHIM_H
#ifndef MPI_VERSION
typedef void* MPI_Comm;
#endif
#endif
