// chase_b200 — implementation of include/chase_c_interface.h.
// Mirrors the reference's sequential C interface
// (interface/chase_c_interface.cpp:105-150 Initialize, :443-466 solve wrapper,
// :493-513 copy_first_nev_results, :3804-... unified setters): one
// process-global solver per scalar type, wrapped in the performance decorator
// for every solve.
#include "../../include/chase_c_interface.h"

#include "../../include/chase_b200_comm.h"

#include "algorithm.hpp"
#include "chase_gpu.hpp"
#include "pchase_gpu.hpp"
#include "performance.hpp"
#include "tridiag_host.hpp"

#include <complex>
#include <cstring>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

extern "C" int chase_b200_device_sync(void) { return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1; }

namespace
{

struct LastRun
{
    double stats[16] = {0};
    std::string trace;
    std::string qr_log;
    std::string error;
};
LastRun g_last;
bool g_trace = false;
int g_matrix_resident = 0; // chase_b200_set_matrix_resident_
int g_device_rng = -1;     // chase_b200_set_device_rng_ (-1: leave the backend's default / env)
int g_mixed = -1;          // chase_b200_set_mixed_precision_ (-1: leave the backend's default / env)
double g_last_sp_cols = 0; // columns the last solve filtered in single precision

// shared by the sequential and the distributed entry points (reference chase_c_interface.cpp:443-491)
template <class T, class Backend>
void run_solve(Backend* solver, std::size_t N, int deg, chase::Base<T> tol, char mode, char opt, char qr,
               double /*unused*/)
{
    auto& config = solver->GetConfig();
    config.SetTol(tol);
    config.SetDeg((std::size_t)deg);
    config.SetOpt(opt == 'S');
    config.SetApprox(mode == 'A');
    config.SetCholQR(qr == 'C');
    solver->clear_logs();
    solver->keep_device_matrix(g_matrix_resident != 0);
    if (g_device_rng >= 0)
        solver->use_device_rng(g_device_rng != 0);
    if (g_mixed >= 0)
        solver->use_mixed_precision(g_mixed != 0);
    chase::PerformanceDecoratorChase<T> perf(solver);
    perf.EnableTrace(g_trace);
    g_last.error.clear();
    try
    {
        if (solver->isPseudoHerm())
            chase::Solve_pseudo(&perf); // reference ChASE_SEQ_Solve, chase_c_interface.cpp:461-466
        else
            chase::Solve(&perf);
    }
    catch (const std::exception& e)
    {
        // C callers cannot catch: report, flag (stats[15] = 1) and return
        std::fprintf(stderr, "chase_b200: solve failed: %s\n", e.what());
        g_last.error = e.what();
    }

    auto& pd = perf.GetPerfData();
    const int factor = chase::is_complex_t<T>::value ? 4 : 1;
    double* s = g_last.stats;
    s[0] = (double)pd.get_iter_count();
    s[1] = (double)pd.get_filtered_vecs();
    s[2] = (double)perf.HemmCalls();
    s[3] = (double)perf.Swaps();
    s[4] = pd.seconds(chase::ChasePerfData::All);
    s[5] = pd.seconds(chase::ChasePerfData::InitVecs);
    s[6] = pd.seconds(chase::ChasePerfData::Lanczos);
    s[7] = pd.seconds(chase::ChasePerfData::Filter);
    s[8] = pd.seconds(chase::ChasePerfData::Qr);
    s[9] = pd.seconds(chase::ChasePerfData::Rr);
    s[10] = pd.seconds(chase::ChasePerfData::Resid);
    // filter work actually done: 2 f N^2 per multiplied column.  Equal to the reference's bookkeeping
    // (performance.hpp:559-570, filtered_vecs) for Hermitian problems; for pseudo-Hermitian ones the reference books
    // 2 * block per HEMM_H2 call even after columns have retired, which would overstate the TFLOP/s.
    s[11] = solver->isPseudoHerm() ? 2.0 * factor * (double)N * (double)N * (double)solver->hemm_cols() / 1e9
                                   : pd.get_filter_flops(N, factor);
    s[12] = pd.get_flops(N, config.GetLanczosIter(), config.GetNumLanczos(), factor);
    s[13] = (double)solver->heev_sweeps();
    s[14] = (double)solver->gather_passes();
    s[15] = g_last.error.empty() ? 0.0 : 1.0;
    g_last_sp_cols = (double)solver->sp_filter_cols();
    g_last.trace.clear();
    for (auto& l : perf.Trace())
    {
        g_last.trace += l;
        g_last.trace += '\n';
    }
    g_last.qr_log.clear();
    for (auto& l : solver->qr_log())
    {
        g_last.qr_log += l;
        g_last.qr_log += '\n';
    }
    if (std::getenv("CHASE_B200_VERBOSE") && solver->get_rank() == 0)
        pd.print(N, factor);
}

// MT = chase::matrix::PseudoHermitianMatrix<T>: the ?chase_init_pseudo_ singletons (reference
// chase_c_interface.cpp:249-320); V then has 2 (nev+nex) columns and ritzv 2 (nev+nex) entries.
template <class T, class MT = chase::matrix::Matrix<T, chase::platform::GPU>>
struct Seq
{
    using R = chase::Base<T>;
    static constexpr std::size_t kWidth =
        std::is_same<MT, chase::matrix::PseudoHermitianMatrix<T, chase::platform::GPU>>::value ? 2 : 1;
    std::unique_ptr<chase::Impl::ChASEGPU<T, MT>> solver;
    std::vector<T> vec;   // internal V when the caller passed NULL
    std::vector<R> ritz;  // internal ritzv when the caller passed NULL
    T* V = nullptr;
    R* ritzv = nullptr;
    std::size_t N = 0;

    static Seq& get()
    {
        static Seq s;
        return s;
    }
    int init(int N_, int nev, int nex, T* H, int ldh, T* V_, R* ritzv_)
    {
        solver.reset();
        N = (std::size_t)N_;
        V = V_;
        if (V == nullptr)
        {
            vec.assign((std::size_t)N_ * kWidth * (std::size_t)(nev + nex), T(0));
            V = vec.data();
        }
        ritzv = ritzv_;
        if (ritzv == nullptr)
        {
            ritz.assign(kWidth * (std::size_t)(nev + nex), R(0));
            ritzv = ritz.data();
        }
        try
        {
            solver.reset(new chase::Impl::ChASEGPU<T, MT>((std::size_t)N_, (std::size_t)nev, (std::size_t)nex, H,
                                                          (std::size_t)ldh, V, (std::size_t)N_, ritzv));
        }
        catch (const std::exception& e)
        {
            std::fprintf(stderr, "chase_b200: init failed: %s\n", e.what());
            g_last.error = e.what();
            return 0;
        }
        return 1;
    }
    void finalize()
    {
        solver.reset();
        std::vector<T>().swap(vec);
        std::vector<R>().swap(ritz);
    }
    void solve(int deg, R tol, char mode, char opt, char qr)
    {
        if (solver)
            run_solve<T>(solver.get(), N, deg, tol, mode, opt, qr, 1.0);
    }
    void get_eigenpairs(T* out, int ld, R* ritz_out)
    {
        if (!solver || out == nullptr || ritz_out == nullptr || ld <= 0)
            return;
        const std::size_t nev = solver->GetNev();
        for (std::size_t j = 0; j < nev; ++j)
            std::memcpy(out + j * (std::size_t)ld, V + j * N, N * sizeof(T));
        std::memcpy(ritz_out, ritzv, nev * sizeof(R));
    }
    void get_resid(R* out)
    {
        if (!solver || !out)
            return;
        const std::size_t nevex = solver->GetNev() + solver->GetNex();
        std::memcpy(out, solver->GetResid(), nevex * sizeof(R));
    }
};

// Distributed solver singleton per scalar type (reference ChASE_DIST, chase_c_interface.cpp:535-640).
template <class T, class MT = chase::matrix::Matrix<T, chase::platform::GPU>>
struct Dist
{
    using R = chase::Base<T>;
    static constexpr std::size_t kWidth =
        std::is_same<MT, chase::matrix::PseudoHermitianMatrix<T, chase::platform::GPU>>::value ? 2 : 1;
    std::unique_ptr<chase::Impl::pChASEGPU<T, MT>> solver;
    std::vector<T> vec;
    std::vector<R> ritz;
    T* V = nullptr;
    R* ritzv = nullptr;
    std::size_t N = 0, m_loc = 0;

    static Dist& get()
    {
        static Dist s;
        return s;
    }
    // mb/nb = 0: block layout (p?chase_init_), else block-cyclic (p?chase_init_blockcyclic_)
    int init(int N_, int nev, int nex, long long mb, long long nb, T* H, int ldh, T* V_, R* ritzv_, int dim0, int dim1,
             char grid_major, void* comm)
    {
        solver.reset();
        try
        {
            if (comm == nullptr)
                throw std::invalid_argument("communicator handle is NULL (create it with chase_b200_comm_init)");
            auto* w = static_cast<chase::b200::WorldComm*>(comm);
            const auto g = chase::b200::Grid2D::make(dim0, dim1, grid_major, w->rank, w->size);
            const chase::b200::Dist1D dr((int64_t)N_, dim0, mb);
            N = (std::size_t)N_;
            m_loc = (std::size_t)dr.local_size(g.i);
            V = V_;
            if (V == nullptr)
            {
                vec.assign(std::max<std::size_t>(m_loc, 1) * kWidth * (std::size_t)(nev + nex), T(0));
                V = vec.data();
            }
            ritzv = ritzv_;
            if (ritzv == nullptr)
            {
                ritz.assign(kWidth * (std::size_t)(nev + nex), R(0));
                ritzv = ritz.data();
            }
            solver.reset(new chase::Impl::pChASEGPU<T, MT>((std::size_t)N_, (std::size_t)nev, (std::size_t)nex, *w, dim0,
                                                       dim1, grid_major, (std::size_t)mb, (std::size_t)nb, H,
                                                       (std::size_t)ldh, V, std::max<std::size_t>(m_loc, 1), ritzv));
        }
        catch (const std::exception& e)
        {
            std::fprintf(stderr, "chase_b200: distributed init failed: %s\n", e.what());
            g_last.error = e.what();
            return 0;
        }
        return 1;
    }
    void finalize()
    {
        solver.reset();
        std::vector<T>().swap(vec);
        std::vector<R>().swap(ritz);
    }
    void solve(int deg, R tol, char mode, char opt, char qr)
    {
        if (solver)
            run_solve<T>(solver.get(), N, deg, tol, mode, opt, qr, 1.0);
    }
    // local rows of the first nev eigenvectors (reference copy_first_nev_results, chase_c_interface.cpp:493-513)
    void get_eigenpairs(T* out, int ld, R* ritz_out)
    {
        if (!solver || out == nullptr || ritz_out == nullptr || ld <= 0)
            return;
        const std::size_t nev = solver->GetNev();
        const std::size_t ldv = std::max<std::size_t>(m_loc, 1);
        for (std::size_t j = 0; j < nev; ++j)
            std::memcpy(out + j * (std::size_t)ld, V + j * ldv, m_loc * sizeof(T));
        std::memcpy(ritz_out, ritzv, nev * sizeof(R));
    }
    void get_resid(R* out)
    {
        if (!solver || !out)
            return;
        std::memcpy(out, solver->GetResid(), (solver->GetNev() + solver->GetNex()) * sizeof(R));
    }
};

using SD = Seq<double>;
using SS = Seq<float>;
using SZ = Seq<std::complex<double>>;
using SC = Seq<std::complex<float>>;
using SZP = Seq<std::complex<double>, chase::matrix::PseudoHermitianMatrix<std::complex<double>, chase::platform::GPU>>;
using SCP = Seq<std::complex<float>, chase::matrix::PseudoHermitianMatrix<std::complex<float>, chase::platform::GPU>>;
using PD = Dist<double>;
using PS = Dist<float>;
using PZ = Dist<std::complex<double>>;
using PC = Dist<std::complex<float>>;
using PZP = Dist<std::complex<double>, chase::matrix::PseudoHermitianMatrix<std::complex<double>, chase::platform::GPU>>;
using PCP = Dist<std::complex<float>, chase::matrix::PseudoHermitianMatrix<std::complex<float>, chase::platform::GPU>>;

// Fortran callers hold their communicator as an INTEGER: index + 1 into this table (see chase_b200_comm_c2f)
inline std::vector<void*>& fortran_comm_table()
{
    static std::vector<void*> t;
    return t;
}

template <class F>
void with_active_config(F&& f)
{
    if (SD::get().solver)
        f(SD::get().solver->GetConfig());
    else if (SS::get().solver)
        f(SS::get().solver->GetConfig());
    else if (SZ::get().solver)
        f(SZ::get().solver->GetConfig());
    else if (SC::get().solver)
        f(SC::get().solver->GetConfig());
    else if (SZP::get().solver)
        f(SZP::get().solver->GetConfig());
    else if (SCP::get().solver)
        f(SCP::get().solver->GetConfig());
    else if (PD::get().solver)
        f(PD::get().solver->GetConfig());
    else if (PS::get().solver)
        f(PS::get().solver->GetConfig());
    else if (PZ::get().solver)
        f(PZ::get().solver->GetConfig());
    else if (PC::get().solver)
        f(PC::get().solver->GetConfig());
    else if (PZP::get().solver)
        f(PZP::get().solver->GetConfig());
    else if (PCP::get().solver)
        f(PCP::get().solver->GetConfig());
}

// keeps the matrix-object constructors of the backends (reference chase_gpu.hpp:195-267) compiled
[[maybe_unused]] void* instantiate_matrix_constructors(chase::matrix::Matrix<double>* Hd,
                                                       chase::matrix::PseudoHermitianMatrix<std::complex<double>>* Hz,
                                                       double* Vd, std::complex<double>* Vz, double* r)
{
    using PH = chase::matrix::PseudoHermitianMatrix<std::complex<double>>;
    if (Hd)
        return new chase::Impl::ChASEGPU<double>(Hd->rows(), 1, 0, Hd, Vd, Hd->rows(), r);
    return new chase::Impl::ChASEGPU<std::complex<double>, PH>(Hz->rows(), 1, 0, Hz, Vz, Hz->rows(), r);
}

size_t copy_out(const std::string& s, char* buf, size_t cap)
{
    if (buf && cap > 0)
    {
        const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size();
}

} // namespace

extern "C"
{
    using cf = std::complex<float>;
    using cd = std::complex<double>;

    void dchase_init_(int* N, int* nev, int* nex, double* H, int* ldh, double* V, double* ritzv, int* init)
    {
        *init = SD::get().init(*N, *nev, *nex, H, *ldh, V, ritzv);
    }
    void schase_init_(int* N, int* nev, int* nex, float* H, int* ldh, float* V, float* ritzv, int* init)
    {
        *init = SS::get().init(*N, *nev, *nex, H, *ldh, V, ritzv);
    }
    void cchase_init_(int* N, int* nev, int* nex, CHASE_B200_CF* H, int* ldh, CHASE_B200_CF* V, float* ritzv,
                      int* init)
    {
        *init = SC::get().init(*N, *nev, *nex, reinterpret_cast<cf*>(H), *ldh, reinterpret_cast<cf*>(V), ritzv);
    }
    void zchase_init_(int* N, int* nev, int* nex, CHASE_B200_CD* H, int* ldh, CHASE_B200_CD* V, double* ritzv,
                      int* init)
    {
        *init = SZ::get().init(*N, *nev, *nex, reinterpret_cast<cd*>(H), *ldh, reinterpret_cast<cd*>(V), ritzv);
    }
    void dchase_init_internal_(int* N, int* nev, int* nex, double* H, int* ldh, int* init)
    {
        *init = SD::get().init(*N, *nev, *nex, H, *ldh, nullptr, nullptr);
    }
    void schase_init_internal_(int* N, int* nev, int* nex, float* H, int* ldh, int* init)
    {
        *init = SS::get().init(*N, *nev, *nex, H, *ldh, nullptr, nullptr);
    }
    void cchase_init_internal_(int* N, int* nev, int* nex, CHASE_B200_CF* H, int* ldh, int* init)
    {
        *init = SC::get().init(*N, *nev, *nex, reinterpret_cast<cf*>(H), *ldh, nullptr, nullptr);
    }
    void zchase_init_internal_(int* N, int* nev, int* nex, CHASE_B200_CD* H, int* ldh, int* init)
    {
        *init = SZ::get().init(*N, *nev, *nex, reinterpret_cast<cd*>(H), *ldh, nullptr, nullptr);
    }

    // pseudo-Hermitian (BSE) singletons, reference chase_c_interface.h:42-58
    void cchase_init_pseudo_(int* N, int* nev, int* nex, CHASE_B200_CF* H, int* ldh, CHASE_B200_CF* V, float* ritzv,
                             int* init)
    {
        *init = SCP::get().init(*N, *nev, *nex, reinterpret_cast<cf*>(H), *ldh, reinterpret_cast<cf*>(V), ritzv);
    }
    void zchase_init_pseudo_(int* N, int* nev, int* nex, CHASE_B200_CD* H, int* ldh, CHASE_B200_CD* V, double* ritzv,
                             int* init)
    {
        *init = SZP::get().init(*N, *nev, *nex, reinterpret_cast<cd*>(H), *ldh, reinterpret_cast<cd*>(V), ritzv);
    }
    void cchase_init_pseudo_internal_(int* N, int* nev, int* nex, CHASE_B200_CF* H, int* ldh, int* init)
    {
        *init = SCP::get().init(*N, *nev, *nex, reinterpret_cast<cf*>(H), *ldh, nullptr, nullptr);
    }
    void zchase_init_pseudo_internal_(int* N, int* nev, int* nex, CHASE_B200_CD* H, int* ldh, int* init)
    {
        *init = SZP::get().init(*N, *nev, *nex, reinterpret_cast<cd*>(H), *ldh, nullptr, nullptr);
    }

    void dchase_finalize_(int* flag)
    {
        SD::get().finalize();
        *flag = 0;
    }
    void schase_finalize_(int* flag)
    {
        SS::get().finalize();
        *flag = 0;
    }
    void cchase_finalize_(int* flag)
    {
        SC::get().finalize();
        SCP::get().finalize();
        *flag = 0;
    }
    void zchase_finalize_(int* flag)
    {
        SZ::get().finalize();
        SZP::get().finalize();
        *flag = 0;
    }

    void dchase_(int* deg, double* tol, char* mode, char* opt, char* qr)
    {
        SD::get().solve(*deg, *tol, *mode, *opt, *qr);
    }
    void schase_(int* deg, float* tol, char* mode, char* opt, char* qr)
    {
        SS::get().solve(*deg, *tol, *mode, *opt, *qr);
    }
    // the pseudo-Hermitian singleton takes the call when it exists (reference chase_c_interface.cpp:2204-2231)
    void zchase_(int* deg, double* tol, char* mode, char* opt, char* qr)
    {
        if (SZP::get().solver)
            SZP::get().solve(*deg, *tol, *mode, *opt, *qr);
        else
            SZ::get().solve(*deg, *tol, *mode, *opt, *qr);
    }
    void cchase_(int* deg, float* tol, char* mode, char* opt, char* qr)
    {
        if (SCP::get().solver)
            SCP::get().solve(*deg, *tol, *mode, *opt, *qr);
        else
            SC::get().solve(*deg, *tol, *mode, *opt, *qr);
    }
    void zchase_pseudo_(int* deg, double* tol, char* mode, char* opt, char* qr)
    {
        SZP::get().solve(*deg, *tol, *mode, *opt, *qr);
    }
    void cchase_pseudo_(int* deg, float* tol, char* mode, char* opt, char* qr)
    {
        SCP::get().solve(*deg, *tol, *mode, *opt, *qr);
    }

    void dchase_get_eigenpairs_(double* V, int* ld, double* ritzv)
    {
        if (ld)
            SD::get().get_eigenpairs(V, *ld, ritzv);
    }
    void schase_get_eigenpairs_(float* V, int* ld, float* ritzv)
    {
        if (ld)
            SS::get().get_eigenpairs(V, *ld, ritzv);
    }
    void cchase_get_eigenpairs_(CHASE_B200_CF* V, int* ld, float* ritzv)
    {
        if (ld && SCP::get().solver)
            SCP::get().get_eigenpairs(reinterpret_cast<cf*>(V), *ld, ritzv);
        else if (ld)
            SC::get().get_eigenpairs(reinterpret_cast<cf*>(V), *ld, ritzv);
    }
    void zchase_get_eigenpairs_(CHASE_B200_CD* V, int* ld, double* ritzv)
    {
        if (ld && SZP::get().solver)
            SZP::get().get_eigenpairs(reinterpret_cast<cd*>(V), *ld, ritzv);
        else if (ld)
            SZ::get().get_eigenpairs(reinterpret_cast<cd*>(V), *ld, ritzv);
    }

    void dchase_get_resid_(double* r) { SD::get().get_resid(r); }
    void schase_get_resid_(float* r) { SS::get().get_resid(r); }
    void cchase_get_resid_(float* r)
    {
        if (SCP::get().solver)
            SCP::get().get_resid(r);
        else
            SC::get().get_resid(r);
    }
    void zchase_get_resid_(double* r)
    {
        if (SZP::get().solver)
            SZP::get().get_resid(r);
        else
            SZ::get().get_resid(r);
    }

    void chase_set_tol_(double* tol)
    {
        with_active_config([&](auto& c) { c.SetTol(*tol); });
    }
    void chase_set_deg_(int* deg)
    {
        with_active_config([&](auto& c) { c.SetDeg((std::size_t)*deg); });
    }
    void chase_set_max_deg_(int* v)
    {
        with_active_config([&](auto& c) { c.SetMaxDeg((std::size_t)*v); });
    }
    void chase_set_deg_extra_(int* v)
    {
        with_active_config([&](auto& c) { c.SetDegExtra((std::size_t)*v); });
    }
    void chase_set_max_iter_(int* v)
    {
        with_active_config([&](auto& c) { c.SetMaxIter((std::size_t)*v); });
    }
    void chase_set_lanczos_iter_(int* v)
    {
        with_active_config([&](auto& c) { c.SetLanczosIter((std::size_t)*v); });
    }
    void chase_set_num_lanczos_(int* v)
    {
        with_active_config([&](auto& c) { c.SetNumLanczos((std::size_t)*v); });
    }
    void chase_set_approx_(int* flag)
    {
        with_active_config([&](auto& c) { c.SetApprox(*flag != 0); });
    }
    void chase_set_opt_(int* flag)
    {
        with_active_config([&](auto& c) { c.SetOpt(*flag != 0); });
    }
    void chase_set_cholqr_(int* flag)
    {
        with_active_config([&](auto& c) { c.SetCholQR(*flag != 0); });
    }
    void chase_enable_sym_check_(int* flag)
    {
        with_active_config([&](auto& c) { c.EnableSymCheck(*flag != 0); });
    }
    void chase_set_decaying_rate_(float* v)
    {
        with_active_config([&](auto& c) { c.SetDecayingRate(*v); });
    }
    void chase_set_cluster_aware_degrees_(int* flag)
    {
        with_active_config([&](auto& c) { c.SetClusterAwareDegrees(*flag != 0); });
    }
    void chase_set_upperb_scale_rate_(float* v)
    {
        with_active_config([&](auto& c) { c.SetUpperbScaleRate(*v); });
    }

    void chase_get_version_(char* version, int* len)
    {
        const char* v = chase_b200_version();
        if (version && len && *len > 0)
        {
            std::strncpy(version, v, (size_t)*len - 1);
            version[*len - 1] = 0;
        }
    }
    void chase_has_cuda_(int* flag) { *flag = 1; }
    void chase_has_nccl_(int* flag) { *flag = 1; }
    void chase_has_scalapack_(int* flag) { *flag = 0; }
    void chase_has_mpi_(int* flag) { *flag = 0; }
    void chase_print_config_(void)
    {
        std::printf("%s: CUDA yes (hand-written sm_100a kernels, no cuBLAS/cuSOLVER), NCCL yes (bound at run time), ScaLAPACK no, MPI no\n",
                    chase_b200_version());
    }

    // ---- Fortran twins (reference chase_c_interface.cpp:2296-2327, 2425-2900: the `_f_` entry points the Fortran
    // module binds to).  A Fortran caller holds its communicator as an INTEGER (MPI_Fint); here that integer is an
    // index into a table of the handles made by chase_b200_comm_init (chase_b200_comm_c2f / _f2c play the role of
    // MPI_Comm_c2f / MPI_Comm_f2c).
    int chase_b200_comm_c2f(void* comm)
    {
        auto& tab = fortran_comm_table();
        for (std::size_t i = 0; i < tab.size(); ++i)
            if (tab[i] == comm)
                return (int)i + 1;
        tab.push_back(comm);
        return (int)tab.size();
    }
    void* chase_b200_comm_f2c(int fcomm)
    {
        auto& tab = fortran_comm_table();
        return (fcomm >= 1 && (std::size_t)fcomm <= tab.size()) ? tab[(std::size_t)fcomm - 1] : nullptr;
    }
    void cchase_init_pseudo_f_(int* N, int* nev, int* nex, CHASE_B200_CF* H, int* ldh, CHASE_B200_CF* V, float* ritzv,
                               int* init)
    {
        cchase_init_pseudo_(N, nev, nex, H, ldh, V, ritzv, init);
    }
    void zchase_init_pseudo_f_(int* N, int* nev, int* nex, CHASE_B200_CD* H, int* ldh, CHASE_B200_CD* V, double* ritzv,
                               int* init)
    {
        zchase_init_pseudo_(N, nev, nex, H, ldh, V, ritzv, init);
    }
    void cchase_pseudo_f_(int* deg, float* tol, char* mode, char* opt, char* qr) { cchase_pseudo_(deg, tol, mode, opt, qr); }
    void zchase_pseudo_f_(int* deg, double* tol, char* mode, char* opt, char* qr) { zchase_pseudo_(deg, tol, mode, opt, qr); }

    // ---- distributed entry points (reference chase_c_interface.h:61-195) ---------------------------------
#define CB2_DIST_INIT_API(X, PSEUDO, TT, CT, RT, SINGLETON)                                                            \
    void p##X##chase_init_##PSEUDO(int* N, int* nev, int* nex, int* m, int* n, CT* H, int* ldh, CT* V, RT* ritzv, int* dim0,  \
                           int* dim1, char* grid_major, MPI_Comm* comm, int* init)                                     \
    {                                                                                                                  \
        (void)m;                                                                                                       \
        (void)n;                                                                                                       \
        *init = SINGLETON::get().init(*N, *nev, *nex, 0, 0, reinterpret_cast<TT*>(H), *ldh, reinterpret_cast<TT*>(V), \
                                      ritzv, *dim0, *dim1, *grid_major, comm ? *comm : nullptr);                       \
    }                                                                                                                  \
    void p##X##chase_init_##PSEUDO##internal_(int* N, int* nev, int* nex, int* m, int* n, CT* H, int* ldh, int* dim0, int* dim1,\
                                    char* grid_major, MPI_Comm* comm, int* init)                                       \
    {                                                                                                                  \
        (void)m;                                                                                                       \
        (void)n;                                                                                                       \
        *init = SINGLETON::get().init(*N, *nev, *nex, 0, 0, reinterpret_cast<TT*>(H), *ldh, nullptr, nullptr, *dim0,  \
                                      *dim1, *grid_major, comm ? *comm : nullptr);                                     \
    }                                                                                                                  \
    void p##X##chase_init_##PSEUDO##blockcyclic_(int* N, int* nev, int* nex, int* mbsize, int* nbsize, CT* H, int* ldh, CT* V,  \
                                       RT* ritzv, int* dim0, int* dim1, char* grid_major, int* irsrc, int* icsrc,      \
                                       MPI_Comm* comm, int* init)                                                      \
    {                                                                                                                  \
        if ((irsrc && *irsrc != 0) || (icsrc && *icsrc != 0))                                                          \
        {                                                                                                              \
            std::fprintf(stderr, "chase_b200: block-cyclic source process must be 0 (as in the reference's numroc)\n"); \
            *init = 0;                                                                                                 \
            return;                                                                                                    \
        }                                                                                                              \
        *init = SINGLETON::get().init(*N, *nev, *nex, *mbsize, *nbsize, reinterpret_cast<TT*>(H), *ldh,               \
                                      reinterpret_cast<TT*>(V), ritzv, *dim0, *dim1, *grid_major,                      \
                                      comm ? *comm : nullptr);                                                         \
    }                                                                                                                  \
    void p##X##chase_init_##PSEUDO##blockcyclic_internal_(int* N, int* nev, int* nex, int* mbsize, int* nbsize, CT* H, int* ldh, \
                                                int* dim0, int* dim1, char* grid_major, int* irsrc, int* icsrc,        \
                                                MPI_Comm* comm, int* init)                                             \
    {                                                                                                                  \
        (void)irsrc;                                                                                                   \
        (void)icsrc;                                                                                                   \
        *init = SINGLETON::get().init(*N, *nev, *nex, *mbsize, *nbsize, reinterpret_cast<TT*>(H), *ldh, nullptr,      \
                                      nullptr, *dim0, *dim1, *grid_major, comm ? *comm : nullptr);                     \
    }                                                                                                                  \
    void p##X##chase_init_##PSEUDO##f_(int* N, int* nev, int* nex, int* m, int* n, CT* H, int* ldh, CT* V, RT* ritzv,  \
                                       int* dim0, int* dim1, char* grid_major, int* fcomm, int* init)                  \
    {                                                                                                                  \
        MPI_Comm c = chase_b200_comm_f2c(*fcomm);                                                                      \
        p##X##chase_init_##PSEUDO(N, nev, nex, m, n, H, ldh, V, ritzv, dim0, dim1, grid_major, &c, init);              \
    }                                                                                                                  \
    void p##X##chase_init_##PSEUDO##internal_f_(int* N, int* nev, int* nex, int* m, int* n, CT* H, int* ldh, int* dim0,\
                                                int* dim1, char* grid_major, int* fcomm, int* init)                    \
    {                                                                                                                  \
        MPI_Comm c = chase_b200_comm_f2c(*fcomm);                                                                      \
        p##X##chase_init_##PSEUDO##internal_(N, nev, nex, m, n, H, ldh, dim0, dim1, grid_major, &c, init);             \
    }                                                                                                                  \
    void p##X##chase_init_##PSEUDO##blockcyclic_f_(int* N, int* nev, int* nex, int* mbsize, int* nbsize, CT* H,        \
                                                   int* ldh, CT* V, RT* ritzv, int* dim0, int* dim1, char* grid_major, \
                                                   int* irsrc, int* icsrc, int* fcomm, int* init)                      \
    {                                                                                                                  \
        MPI_Comm c = chase_b200_comm_f2c(*fcomm);                                                                      \
        p##X##chase_init_##PSEUDO##blockcyclic_(N, nev, nex, mbsize, nbsize, H, ldh, V, ritzv, dim0, dim1, grid_major, \
                                                irsrc, icsrc, &c, init);                                               \
    }                                                                                                                  \
    void p##X##chase_init_##PSEUDO##blockcyclic_internal_f_(int* N, int* nev, int* nex, int* mbsize, int* nbsize,      \
                                                            CT* H, int* ldh, int* dim0, int* dim1, char* grid_major,   \
                                                            int* irsrc, int* icsrc, int* fcomm, int* init)             \
    {                                                                                                                  \
        MPI_Comm c = chase_b200_comm_f2c(*fcomm);                                                                      \
        p##X##chase_init_##PSEUDO##blockcyclic_internal_(N, nev, nex, mbsize, nbsize, H, ldh, dim0, dim1, grid_major,  \
                                                         irsrc, icsrc, &c, init);                                      \
    }
#define CB2_DIST_RUN_API(X, TT, CT, RT, SINGLETON)                                                                     \
    void p##X##chase_(int* deg, RT* tol, char* mode, char* opt, char* qr)                                              \
    {                                                                                                                  \
        SINGLETON::get().solve(*deg, *tol, *mode, *opt, *qr);                                                          \
    }                                                                                                                  \
    void p##X##chase_finalize_(int* flag)                                                                              \
    {                                                                                                                  \
        SINGLETON::get().finalize();                                                                                   \
        *flag = 0;                                                                                                     \
    }                                                                                                                  \
    void p##X##chase_get_eigenpairs_(CT* V, int* ld, RT* ritzv)                                                        \
    {                                                                                                                  \
        if (ld)                                                                                                        \
            SINGLETON::get().get_eigenpairs(reinterpret_cast<TT*>(V), *ld, ritzv);                                     \
    }                                                                                                                  \
    void p##X##chase_get_resid_(RT* r) { SINGLETON::get().get_resid(r); }
    // complex types: the pseudo-Hermitian singleton takes the call while it exists (reference
    // chase_c_interface.cpp:1976-1996, 2017-2037)
#define CB2_DIST_RUN_API_CPLX(X, TT, CT, RT, SINGLETON, PSINGLETON)                                                    \
    void p##X##chase_(int* deg, RT* tol, char* mode, char* opt, char* qr)                                              \
    {                                                                                                                  \
        if (PSINGLETON::get().solver)                                                                                  \
            PSINGLETON::get().solve(*deg, *tol, *mode, *opt, *qr);                                                     \
        else                                                                                                           \
            SINGLETON::get().solve(*deg, *tol, *mode, *opt, *qr);                                                      \
    }                                                                                                                  \
    void p##X##chase_finalize_(int* flag)                                                                              \
    {                                                                                                                  \
        SINGLETON::get().finalize();                                                                                   \
        PSINGLETON::get().finalize();                                                                                  \
        *flag = 0;                                                                                                     \
    }                                                                                                                  \
    void p##X##chase_get_eigenpairs_(CT* V, int* ld, RT* ritzv)                                                        \
    {                                                                                                                  \
        if (ld && PSINGLETON::get().solver)                                                                            \
            PSINGLETON::get().get_eigenpairs(reinterpret_cast<TT*>(V), *ld, ritzv);                                    \
        else if (ld)                                                                                                   \
            SINGLETON::get().get_eigenpairs(reinterpret_cast<TT*>(V), *ld, ritzv);                                     \
    }                                                                                                                  \
    void p##X##chase_get_resid_(RT* r)                                                                                 \
    {                                                                                                                  \
        if (PSINGLETON::get().solver)                                                                                  \
            PSINGLETON::get().get_resid(r);                                                                            \
        else                                                                                                           \
            SINGLETON::get().get_resid(r);                                                                             \
    }

    CB2_DIST_INIT_API(d, , double, double, double, PD)
    CB2_DIST_INIT_API(s, , float, float, float, PS)
    CB2_DIST_INIT_API(z, , cd, CHASE_B200_CD, double, PZ)
    CB2_DIST_INIT_API(c, , cf, CHASE_B200_CF, float, PC)
    CB2_DIST_INIT_API(z, pseudo_, cd, CHASE_B200_CD, double, PZP)
    CB2_DIST_INIT_API(c, pseudo_, cf, CHASE_B200_CF, float, PCP)
    CB2_DIST_RUN_API(d, double, double, double, PD)
    CB2_DIST_RUN_API(s, float, float, float, PS)
    CB2_DIST_RUN_API_CPLX(z, cd, CHASE_B200_CD, double, PZ, PZP)
    CB2_DIST_RUN_API_CPLX(c, cf, CHASE_B200_CF, float, PC, PCP)
#undef CB2_DIST_INIT_API
#undef CB2_DIST_RUN_API
#undef CB2_DIST_RUN_API_CPLX

    // Device-resident input for matrices that should never exist on the host (C4: 28.8 GB per GPU): copies a
    // column-major device block (m_loc x n_loc, leading dimension ld_src) into the active distributed solver and marks
    // it resident, so that p?chase_ does not read the host pointer given at init.
    // In-place variant of the hand-over below for blocks that only fit once (BASELINE config C4 on 2 GPUs: 115 GB per
    // GPU): returns the solver's own device buffer (column-major local block, leading dimension *ld_out) so that the
    // caller generates the block directly into it, then chase_b200_dist_mark_device_matrix_ declares it valid.
    int chase_b200_dist_device_matrix_(char* type, void** ptr_out, long long* ld_out)
    {
        auto get = [&](auto& inst) -> int
        {
            if (!inst.solver)
                return -1;
            *ptr_out = (void*)inst.solver->device_H();
            *ld_out = (long long)inst.solver->device_lda();
            return 0;
        };
        switch (*type)
        {
            case 'd': return get(PD::get());
            case 's': return get(PS::get());
            case 'z': return PZP::get().solver ? get(PZP::get()) : get(PZ::get());
            case 'c': return PCP::get().solver ? get(PCP::get()) : get(PC::get());
        }
        return -1;
    }
    int chase_b200_dist_mark_device_matrix_(char* type)
    {
        auto mark = [&](auto& inst) -> int
        {
            if (!inst.solver)
                return -1;
            inst.solver->mark_matrix_on_device();
            g_matrix_resident = 1;
            return 0;
        };
        switch (*type)
        {
            case 'd': return mark(PD::get());
            case 's': return mark(PS::get());
            case 'z': return PZP::get().solver ? mark(PZP::get()) : mark(PZ::get());
            case 'c': return PCP::get().solver ? mark(PCP::get()) : mark(PC::get());
        }
        return -1;
    }
    int chase_b200_dist_load_device_matrix_(char* type, const void* src_dev, long long* ld_src)
    {
        auto load = [&](auto& inst) -> int
        {
            using S = std::remove_reference_t<decltype(*inst.solver)>;
            if (!inst.solver)
                return -1;
            auto* sv = inst.solver.get();
            using TT = std::remove_pointer_t<decltype(sv->device_H())>;
            if (sv->local_rows() > 0 && sv->local_cols() > 0)
            {
                if (cudaMemcpy2DAsync(sv->device_H(), sv->device_lda() * sizeof(TT), src_dev,
                                      (size_t)*ld_src * sizeof(TT), sv->local_rows() * sizeof(TT), sv->local_cols(),
                                      cudaMemcpyDeviceToDevice, sv->stream()) != cudaSuccess)
                    return -1;
                if (cudaStreamSynchronize(sv->stream()) != cudaSuccess)
                    return -1;
            }
            sv->mark_matrix_on_device();
            g_matrix_resident = 1;
            (void)sizeof(S);
            return 0;
        };
        switch (*type)
        {
            case 'd': return load(PD::get());
            case 's': return load(PS::get());
            case 'z': return PZP::get().solver ? load(PZP::get()) : load(PZ::get());
            case 'c': return PCP::get().solver ? load(PCP::get()) : load(PC::get());
        }
        return -1;
    }

    // ---- matrix file I/O (reference chase_c_interface.h:196-214; the un-prefixed names are aliases) -----------------
#define CB2_IO_API(X, SEQ0, SEQ1, DIST0, DIST1)                                                                        \
    void p##X##chase_readHam_(const char* filename)                                                                    \
    {                                                                                                                  \
        try                                                                                                            \
        {                                                                                                              \
            if (DIST1::get().solver)                                                                                   \
                DIST1::get().solver->loadProblemFromFile(filename);                                                    \
            else if (DIST0::get().solver)                                                                              \
                DIST0::get().solver->loadProblemFromFile(filename);                                                    \
            else if (SEQ1::get().solver)                                                                               \
                SEQ1::get().solver->loadProblemFromFile(filename);                                                     \
            else if (SEQ0::get().solver)                                                                               \
                SEQ0::get().solver->loadProblemFromFile(filename);                                                     \
        }                                                                                                              \
        catch (const std::exception& e)                                                                                \
        {                                                                                                              \
            std::fprintf(stderr, "%s\n", e.what());                                                                    \
            g_last.error = e.what();                                                                                   \
        }                                                                                                              \
    }                                                                                                                  \
    void X##chase_readHam_(const char* filename) { p##X##chase_readHam_(filename); }                                   \
    void p##X##chase_wrtHam_(const char* filename)                                                                     \
    {                                                                                                                  \
        try                                                                                                            \
        {                                                                                                              \
            if (DIST1::get().solver)                                                                                   \
                DIST1::get().solver->saveProblemToFile(filename);                                                      \
            else if (DIST0::get().solver)                                                                              \
                DIST0::get().solver->saveProblemToFile(filename);                                                      \
            else if (SEQ1::get().solver)                                                                               \
                SEQ1::get().solver->saveProblemToFile(filename);                                                       \
            else if (SEQ0::get().solver)                                                                               \
                SEQ0::get().solver->saveProblemToFile(filename);                                                       \
        }                                                                                                              \
        catch (const std::exception& e)                                                                                \
        {                                                                                                              \
            std::fprintf(stderr, "%s\n", e.what());                                                                    \
            g_last.error = e.what();                                                                                   \
        }                                                                                                              \
    }
    CB2_IO_API(d, SD, SD, PD, PD)
    CB2_IO_API(s, SS, SS, PS, PS)
    CB2_IO_API(z, SZ, SZP, PZ, PZP)
    CB2_IO_API(c, SC, SCP, PC, PCP)
#undef CB2_IO_API

    // host tridiagonal eigensolver of the Lanczos step (exported for the CPU tests)
    int chase_b200_tridiag_eig_host(int n, const double* d, const double* e, double* w, double* Z)
    {
        return chase::b200::tridiag_eig_host(n, d, e, w, Z);
    }

    // ---- communicator bootstrap (include/chase_b200_comm.h) ------------------------------------------------
    int chase_b200_comm_unique_id(void* id_out)
    {
        try
        {
            static_assert(sizeof(ncclUniqueId) == CHASE_B200_COMM_ID_BYTES, "ncclUniqueId size");
            ncclUniqueId id;
            if (chase::b200::NcclApi::get().GetUniqueId(&id) != ncclSuccess)
                return -1;
            std::memcpy(id_out, &id, sizeof id);
            return 0;
        }
        catch (const std::exception& e)
        {
            std::fprintf(stderr, "%s\n", e.what());
            return -1;
        }
    }
    int chase_b200_comm_init(int rank, int nranks, const void* id_in, int device, void** comm_out)
    {
        try
        {
            if (cudaSetDevice(device) != cudaSuccess)
            {
                std::fprintf(stderr, "chase_b200: cudaSetDevice(%d) failed\n", device);
                return -1;
            }
            ncclUniqueId id;
            std::memcpy(&id, id_in, sizeof id);
            auto* w = new chase::b200::WorldComm;
            w->rank = rank;
            w->size = nranks;
            w->device = device;
            ncclResult_t r = chase::b200::NcclApi::get().CommInitRank(&w->comm, nranks, id, rank);
            if (r != ncclSuccess)
            {
                std::fprintf(stderr, "chase_b200: ncclCommInitRank failed: %s\n",
                             chase::b200::NcclApi::get().GetErrorString(r));
                delete w;
                return -1;
            }
            *comm_out = w;
            return 0;
        }
        catch (const std::exception& e)
        {
            std::fprintf(stderr, "%s\n", e.what());
            return -1;
        }
    }
    int chase_b200_comm_free(void* comm)
    {
        auto* w = static_cast<chase::b200::WorldComm*>(comm);
        if (!w)
            return 0;
        if (w->comm)
            chase::b200::NcclApi::get().CommDestroy(w->comm);
        delete w;
        return 0;
    }
    int chase_b200_comm_rank(void* comm) { return comm ? static_cast<chase::b200::WorldComm*>(comm)->rank : -1; }
    int chase_b200_comm_size(void* comm) { return comm ? static_cast<chase::b200::WorldComm*>(comm)->size : -1; }
    long long chase_b200_local_size(long long N, int nprocs, long long nb, int p)
    {
        return chase::b200::Dist1D(N, nprocs, nb).local_size(p);
    }
    int chase_b200_global_indices(long long N, int nprocs, long long nb, int p, long long* out)
    {
        const auto g = chase::b200::Dist1D(N, nprocs, nb).global_indices(p);
        for (std::size_t i = 0; i < g.size(); ++i)
            out[i] = g[i];
        return 0;
    }
    int chase_b200_redistribution_map(long long N, int src_nprocs, long long src_nb, long long src_stride,
                                      int dst_nprocs, long long dst_nb, int pd, long long* out)
    {
        const chase::b200::Dist1D src(N, src_nprocs, src_nb), dst(N, dst_nprocs, dst_nb);
        for (const auto& c : chase::b200::redistribution_list(src, src_stride, dst, pd))
            for (long long t = 0; t < c.len; ++t)
                out[c.dst0 + t] = c.src0 + t;
        return 0;
    }
    int chase_b200_grid_coords(int dim0, int dim1, char grid_major, int rank, int* row_out, int* col_out)
    {
        try
        {
            const auto g = chase::b200::Grid2D::make(dim0, dim1, grid_major, rank, dim0 * dim1);
            *row_out = g.i;
            *col_out = g.j;
            return 0;
        }
        catch (const std::exception& e)
        {
            std::fprintf(stderr, "chase_b200: %s\n", e.what());
            return -1;
        }
    }

    void chase_b200_start_vectors_d(int64_t N, int64_t m, double* V, int64_t ldv)
    {
        chase::Impl::fill_start_vectors<double>((size_t)N, (size_t)m, V, (size_t)ldv);
    }
    void chase_b200_start_vectors_s(int64_t N, int64_t m, float* V, int64_t ldv)
    {
        chase::Impl::fill_start_vectors<float>((size_t)N, (size_t)m, V, (size_t)ldv);
    }
    void chase_b200_start_vectors_z(int64_t N, int64_t m, CHASE_B200_CD* V, int64_t ldv)
    {
        chase::Impl::fill_start_vectors<cd>((size_t)N, (size_t)m, reinterpret_cast<cd*>(V), (size_t)ldv);
    }
    void chase_b200_start_vectors_c(int64_t N, int64_t m, CHASE_B200_CF* V, int64_t ldv)
    {
        chase::Impl::fill_start_vectors<cf>((size_t)N, (size_t)m, reinterpret_cast<cf*>(V), (size_t)ldv);
    }

    void chase_b200_get_stats_(double* out, int* n)
    {
        const int cnt = (*n < 16) ? *n : 16;
        for (int i = 0; i < cnt; ++i)
            out[i] = g_last.stats[i];
    }
    void chase_b200_trace_enable_(int* flag) { g_trace = (*flag != 0); }
    void chase_b200_set_matrix_resident_(int* flag) { g_matrix_resident = *flag; }
    void chase_b200_set_device_rng_(int* flag) { g_device_rng = *flag; }
    void chase_b200_set_mixed_precision_(int* flag) { g_mixed = *flag; }
    double chase_b200_last_sp_filter_cols_(void) { return g_last_sp_cols; }
    size_t chase_b200_trace_copy_(char* buf, size_t cap) { return copy_out(g_last.trace, buf, cap); }
    size_t chase_b200_qr_log_copy_(char* buf, size_t cap) { return copy_out(g_last.qr_log, buf, cap); }
    size_t chase_b200_last_error_copy_(char* buf, size_t cap) { return copy_out(g_last.error, buf, cap); }
}
