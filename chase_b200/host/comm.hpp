// chase_b200 host layer — the communication component of the distributed backend.
//
// Replaces chase::grid::MpiGrid2D's three NCCL communicators (world, row, column; grid/mpiGrid2D.hpp:449-485) and
// the wrappers of grid/nccl_utils.hpp:121-204 (allreduce / broadcast with complex counted as 2 x real).  One process
// per GPU; every collective on the path is NCCL over NVLink on the backend's stream.  The only job MPI had on the
// reference's GPU path — shipping the ncclUniqueId — is left to the launcher (include/chase_b200_comm.h).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the single-GPU library has no NCCL dependency, and inside a
// PyTorch process the already-loaded NCCL is shared instead of loading a second copy.
#pragma once
#include "dist_layout.hpp"

#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <nccl.h>

#include <complex>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace chase
{
namespace b200
{

struct NcclApi
{
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;

    static NcclApi& get()
    {
        static NcclApi api = load();
        return api;
    }
    static NcclApi load()
    {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h)
            h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h)
            throw std::runtime_error(std::string("chase_b200: cannot load NCCL (libnccl.so.2): ") + dlerror());
        auto sym = [&](const char* n)
        {
            void* p = dlsym(h, n);
            if (!p)
                throw std::runtime_error(std::string("chase_b200: NCCL symbol missing: ") + n);
            return p;
        };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommSplit = reinterpret_cast<decltype(a.CommSplit)>(sym("ncclCommSplit"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
        a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
        return a;
    }
};

// reference error convention: print and exit (grid/nccl_utils.hpp:13-27)
#define CB2_NCCL(call)                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
        ncclResult_t r__ = (call);                                                                                     \
        if (r__ != ncclSuccess)                                                                                        \
        {                                                                                                              \
            std::fprintf(stderr, "chase_b200: NCCL failure '%s' at %s:%d\n",                                          \
                         ::chase::b200::NcclApi::get().GetErrorString(r__), __FILE__, __LINE__);                       \
            std::exit(EXIT_FAILURE);                                                                                   \
        }                                                                                                              \
    } while (0)

// World communicator handle handed out through the C ABI (what the p?chase_init_ calls receive as "MPI_Comm*").
struct WorldComm
{
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1, device = 0;
};

template <class T>
struct NcclType;
template <>
struct NcclType<float>
{
    static constexpr ncclDataType_t dt = ncclFloat32;
    static constexpr int mul = 1;
};
template <>
struct NcclType<double>
{
    static constexpr ncclDataType_t dt = ncclFloat64;
    static constexpr int mul = 1;
};
template <>
struct NcclType<std::complex<float>>
{
    static constexpr ncclDataType_t dt = ncclFloat32;
    static constexpr int mul = 2;
};
template <>
struct NcclType<std::complex<double>>
{
    static constexpr ncclDataType_t dt = ncclFloat64;
    static constexpr int mul = 2;
};
template <>
struct NcclType<int>
{
    static constexpr ncclDataType_t dt = ncclInt32;
    static constexpr int mul = 1;
};

// The grid's communicators: world + "row" (same grid row i, size c) + "column" (same grid column j, size r).
class GridComm
{
public:
    GridComm(const WorldComm& w, const Grid2D& g) : grid_(g), world_(w.comm)
    {
        auto& n = NcclApi::get();
        // row communicator: ranks with the same i, ordered by j; column communicator: same j, ordered by i
        CB2_NCCL(n.CommSplit(world_, g.i, g.j, &row_, nullptr));
        CB2_NCCL(n.CommSplit(world_, g.j, g.i, &col_, nullptr));
    }
    GridComm(const GridComm&) = delete;
    ~GridComm()
    {
        auto& n = NcclApi::get();
        if (row_)
            n.CommDestroy(row_);
        if (col_)
            n.CommDestroy(col_);
    }
    const Grid2D& grid() const { return grid_; }
    ncclComm_t world() const { return world_; }
    ncclComm_t row() const { return row_; }
    ncclComm_t col() const { return col_; }

    template <class T>
    void allreduce_sum(T* buf, size_t count, ncclComm_t c, cudaStream_t st) const
    {
        CB2_NCCL(NcclApi::get().AllReduce(buf, buf, count * NcclType<T>::mul, NcclType<T>::dt, ncclSum, c, st));
        ++collectives_;
    }
    template <class T>
    void broadcast(T* buf, size_t count, int root, ncclComm_t c, cudaStream_t st) const
    {
        CB2_NCCL(NcclApi::get().Broadcast(buf, buf, count * NcclType<T>::mul, NcclType<T>::dt, root, c, st));
        ++collectives_;
    }
    template <class T>
    void allgather(const T* send, T* recv, size_t sendcount, ncclComm_t c, cudaStream_t st) const
    {
        CB2_NCCL(NcclApi::get().AllGather(send, recv, sendcount * NcclType<T>::mul, NcclType<T>::dt, c, st));
        ++collectives_;
    }
    std::size_t collectives() const { return collectives_; }

private:
    Grid2D grid_;
    ncclComm_t world_ = nullptr, row_ = nullptr, col_ = nullptr;
    mutable std::size_t collectives_ = 0;
};

} // namespace b200
} // namespace chase
