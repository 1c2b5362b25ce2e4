// chase_b200 host layer — index maps of the reference's distributed data structures, as plain arithmetic.
//
//   Grid2D  <- chase::grid::MpiGrid2D<RowMajor|ColMajor>      (grid/mpiGrid2D.hpp:200-256, 402-447)
//   Dist1D  <- BlockBlockMatrix (ceil(N/p)-sized blocks, last rank takes the remainder,
//              linalg/distMatrix/distMatrix.hpp:1992-2039) and BlockCyclicMatrix
//              (ScaLAPACK numroc with source process 0, distMatrix.hpp:44-67, 2899-2912)
//
// A is distributed by Dist1D "rows" over the grid rows and Dist1D "cols" over the grid columns; a column-layout
// multivector (V) is split like A's rows and replicated over grid columns, a row-layout multivector (W) is split like
// A's columns and replicated over grid rows (linalg/distMatrix/distMultiVector.hpp:1108-1120).
//
// Pure host code (no CUDA, no NCCL): unit-tested on CPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace chase
{
namespace b200
{

// ScaLAPACK NUMROC with isrcproc = 0 (reference distMatrix.hpp:44-67); returns the local extent only.
inline int64_t numroc(int64_t n, int64_t nb, int iproc, int nprocs)
{
    const int64_t nblocks = n / nb;
    int64_t loc = (nblocks / nprocs) * nb;
    const int64_t extra = nblocks % nprocs;
    if (iproc < extra)
        loc += nb;
    else if (iproc == extra)
        loc += n % nb;
    return loc;
}

struct Segment
{
    int64_t g0;  // first global index
    int64_t l0;  // first local index on the owner
    int64_t len;
};

struct Dist1D
{
    int64_t N = 0;
    int nprocs = 1;
    int64_t nb = 0; // 0: block layout; > 0: block-cyclic with this block size

    Dist1D() = default;
    Dist1D(int64_t N_, int nprocs_, int64_t nb_) : N(N_), nprocs(nprocs_), nb(nb_)
    {
        if (N < 0 || nprocs < 1 || nb < 0)
            throw std::invalid_argument("Dist1D: bad arguments");
    }
    int64_t block_len() const // block layout: distMatrix.hpp:2000-2008
    {
        return (N % nprocs == 0) ? N / nprocs : std::min<int64_t>(N, N / nprocs + 1);
    }
    int64_t local_size(int p) const
    {
        if (nb > 0)
            return numroc(N, nb, p, nprocs);
        const int64_t len = block_len();
        if (p < nprocs - 1)
            return std::max<int64_t>(0, std::min(len, N - (int64_t)p * len));
        return std::max<int64_t>(0, N - (int64_t)(nprocs - 1) * len);
    }
    int64_t max_local_size() const
    {
        int64_t m = 0;
        for (int p = 0; p < nprocs; ++p)
            m = std::max(m, local_size(p));
        return m;
    }
    int owner(int64_t g) const
    {
        if (nb > 0)
            return (int)((g / nb) % nprocs);
        return (int)std::min<int64_t>(g / block_len(), nprocs - 1);
    }
    int64_t local_index(int64_t g) const
    {
        if (nb > 0)
            return (g / (nb * nprocs)) * nb + g % nb;
        return g - (int64_t)owner(g) * block_len();
    }
    // contiguous runs owned by process p, ascending
    std::vector<Segment> segments(int p) const
    {
        std::vector<Segment> out;
        if (nb == 0)
        {
            const int64_t len = local_size(p);
            if (len > 0)
                out.push_back({(int64_t)p * block_len(), 0, len});
            return out;
        }
        int64_t l0 = 0;
        for (int64_t g0 = (int64_t)p * nb; g0 < N; g0 += nb * nprocs)
        {
            const int64_t len = std::min(nb, N - g0);
            out.push_back({g0, l0, len});
            l0 += len;
        }
        return out;
    }
    // global index of every local row of process p
    std::vector<int64_t> global_indices(int p) const
    {
        std::vector<int64_t> g;
        for (const auto& s : segments(p))
            for (int64_t t = 0; t < s.len; ++t)
                g.push_back(s.g0 + t);
        return g;
    }
};

// Row copy list: dst rows [dst0, dst0+len) <- src rows [src0, src0+len)
struct RowCopy
{
    int64_t src0, dst0, len;
};

// Copy list that assembles the piece `dst` (process pd of distribution D_dst) from the all-gathered pieces of
// distribution D_src, where the gathered buffer stacks the pieces of processes 0..nprocs-1 with stride `src_stride`
// rows.  Covers every redistribution on the path:
//   column layout -> row layout  (D_src = rows, D_dst = cols)     distMultiVector.hpp:2817-2909 redistributeImpl
//   a piece      -> global order (D_dst = trivial 1-process layout)
inline std::vector<RowCopy> redistribution_list(const Dist1D& src, int64_t src_stride, const Dist1D& dst, int pd)
{
    std::vector<RowCopy> out;
    for (const auto& sd : dst.segments(pd))
    {
        int64_t g = sd.g0;
        const int64_t gend = sd.g0 + sd.len;
        while (g < gend)
        {
            const int ps = src.owner(g);
            const int64_t ls = src.local_index(g);
            // run length inside the source segment
            int64_t run;
            if (src.nb > 0)
                run = std::min<int64_t>(src.nb - g % src.nb, gend - g);
            else
            {
                const int64_t seg_end = (ps == src.nprocs - 1) ? src.N : (int64_t)(ps + 1) * src.block_len();
                run = std::min<int64_t>(seg_end - g, gend - g);
            }
            const int64_t s0 = (int64_t)ps * src_stride + ls, d0 = sd.l0 + (g - sd.g0);
            if (!out.empty() && out.back().src0 + out.back().len == s0 && out.back().dst0 + out.back().len == d0)
                out.back().len += run;
            else
                out.push_back({s0, d0, run});
            g += run;
        }
    }
    return out;
}

// Process grid (reference MpiGrid2D: dims[0] = rows >= dims[1] = cols; MPI_Cart_create row-major ordering of the
// (possibly swapped) dimensions, mpiGrid2D.hpp:402-430).
struct Grid2D
{
    int r = 1, c = 1; // grid rows / columns
    int i = 0, j = 0; // my coordinates
    int rank = 0, nranks = 1;
    char major = 'R';

    static Grid2D make(int dim0, int dim1, char grid_major, int rank, int nranks)
    {
        if (dim0 <= 0 || dim1 <= 0)
            throw std::invalid_argument("Row and column dimensions of 2D grid must be greater than 0");
        if (dim0 < dim1)
            throw std::invalid_argument("Row dimension of 2D grid must be greater than or equal to column dimension");
        if (grid_major != 'R' && grid_major != 'C')
            throw std::runtime_error("Invalid grid major type, expected 'C' or 'R'.");
        if (dim0 * dim1 != nranks)
            throw std::invalid_argument("grid dimensions do not match the communicator size");
        Grid2D g;
        g.r = dim0;
        g.c = dim1;
        g.rank = rank;
        g.nranks = nranks;
        g.major = grid_major;
        if (grid_major == 'R')
        {
            g.i = rank / dim1;
            g.j = rank % dim1;
        }
        else
        {
            g.j = rank / dim0;
            g.i = rank % dim0;
        }
        return g;
    }
    int rank_of(int ii, int jj) const { return major == 'R' ? ii * c + jj : jj * r + ii; }
};

} // namespace b200
} // namespace chase
