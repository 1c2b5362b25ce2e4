// TEST INFRASTRUCTURE — not product code.
//
// Cross-check of the NEW host driver (chase_b200/host/algorithm.hpp: Algorithm<T>::solve / solve_pseudo) against the
// reference's own driver (/root/reference/algorithm/algorithm.inc:1376-1788, 1834-2220).  Both drivers run on the
// UNMODIFIED reference CPU backend (chase::Impl::ChASECPU, Impl/chase_cpu/chase_cpu.hpp:67) with identical inputs;
// because the backend arithmetic is then bit-identical, every ChaseBase call (arguments included), every Ritz value
// and residual and the returned eigenvectors must be EXACTLY equal.  This pins the decision logic of the new driver
// on the CPU, independently of the CUDA kernels.
//
// The new host headers live in namespace chase as well (drop-in), so they are included under a renamed namespace.
//
// Usage: xcheck_<d|z|pz|pc> --N n --nev k --nex x --matrix clement|file:<path> [--tol t] [--deg d] [--opt 0|1]
//                           [--maxiter i] [--numlanczos n] [--lanczositer m]
// Exit code 0 = identical, 1 = mismatch (first difference on stderr).
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "algorithm/performance.hpp"
#include "Impl/chase_cpu/chase_cpu.hpp"

#include "trace_backend.hpp"

#define chase chase_b2
#include "../chase_b200/host/algorithm.hpp"
#undef chase

#ifndef REF_T
#define REF_T double
#endif
using T = REF_T;
using R = chase::Base<T>;

// new-driver view of a reference backend
template <class S>
class Adapter : public chase_b2::ChaseBase<S>
{
    using B = chase_b2::Base<S>;

public:
    Adapter(chase::ChaseBase<S>* ref) : ref_(ref), cfg_(ref->GetN(), ref->GetNev(), ref->GetNex())
    {
        auto& c = ref->GetConfig();
        cfg_.SetTol(c.GetTol());
        cfg_.SetDeg(c.GetDeg());
        cfg_.SetMaxDeg(c.GetMaxDeg());
        cfg_.SetDegExtra(c.GetDegExtra());
        cfg_.SetMaxIter(c.GetMaxIter());
        cfg_.SetOpt(c.DoOptimization());
        cfg_.SetApprox(c.UseApprox());
        cfg_.SetLanczosIter(c.GetLanczosIter());
        cfg_.SetNumLanczos(c.GetNumLanczos());
        cfg_.SetCholQR(c.DoCholQR());
        cfg_.SetDecayingRate(c.GetDecayingRate());
        cfg_.SetUpperbScaleRate(c.GetUpperbScaleRate());
        cfg_.SetClusterAwareDegrees(c.UseClusterAwareDegrees());
    }
    void Shift(S c, bool un = false) override { ref_->Shift(c, un); }
    void HEMM(std::size_t n, S a, S b, std::size_t ol, std::size_t orr = 0) override { ref_->HEMM(n, a, b, ol, orr); }
    void HEMM_H2(std::size_t n, S a, S b, S g, std::size_t ol, std::size_t orr = 0) override
    {
        ref_->HEMM_H2(n, a, b, g, ol, orr);
    }
    void ApplyKconjugate(std::size_t b) override { ref_->ApplyKconjugate(b); }
    void FilterPhaseStart() override { ref_->FilterPhaseStart(); }
    void FilterPhaseEnd() override { ref_->FilterPhaseEnd(); }
    void QR(std::size_t f, B cond) override { ref_->QR(f, cond); }
    void RR(B* r, std::size_t b) override { ref_->RR(r, b); }
    void Sort(B* a, B* b, B* c) override { ref_->Sort(a, b, c); }
    void Resd(B* r, B* d, std::size_t f) override { ref_->Resd(r, d, f); }
    void Lanczos(std::size_t m, B* ub) override { ref_->Lanczos(m, ub); }
    void Lanczos(std::size_t M, std::size_t nv, B* ub, B* rv, B* tau, B* rV) override
    {
        ref_->Lanczos(M, nv, ub, rv, tau, rV);
    }
    void LanczosDos(std::size_t idx, std::size_t m, S* rvc) override { ref_->LanczosDos(idx, m, rvc); }
    void Swap(std::size_t i, std::size_t j) override { ref_->Swap(i, j); }
    void Lock(std::size_t n) override { ref_->Lock(n); }
    bool checkSymmetryEasy() override { return ref_->checkSymmetryEasy(); }
    bool isSym() override { return ref_->isSym(); }
    bool checkPseudoHermicityEasy() override { return ref_->checkPseudoHermicityEasy(); }
    bool isPseudoHerm() override { return ref_->isPseudoHerm(); }
    void symOrHermMatrix(char u) override { ref_->symOrHermMatrix(u); }
    void Start() override { ref_->Start(); }
    void End() override { ref_->End(); }
    void initVecs(bool random) override { ref_->initVecs(random); }
    std::size_t GetN() const override { return ref_->GetN(); }
    std::size_t GetNev() override { return ref_->GetNev(); }
    std::size_t GetNex() override { return ref_->GetNex(); }
    std::size_t GetLanczosIter() override { return ref_->GetLanczosIter(); }
    std::size_t GetNumLanczos() override { return ref_->GetNumLanczos(); }
    std::size_t GetRitzvBlockSize() const override { return ref_->GetRitzvBlockSize(); }
    B* GetRitzv() override { return ref_->GetRitzv(); }
    B* GetResid() override { return ref_->GetResid(); }
    chase_b2::ChaseConfig<S>& GetConfig() override { return cfg_; }
    int get_nprocs() override { return ref_->get_nprocs(); }
    int get_rank() override { return ref_->get_rank(); }
    void set_early_locked_residuals(std::vector<B> v) override { ref_->set_early_locked_residuals(v); }

private:
    chase::ChaseBase<S>* ref_;
    chase_b2::ChaseConfig<S> cfg_;
};

template <typename U>
struct mk
{
    static U f(double re) { return U(re); }
};

int main(int argc, char** argv)
{
    std::size_t N = 256, nev = 24, nex = 16, maxiter = 25;
    std::string matrix = "clement";
    double tol = -1;
    long deg = -1, numlanczos = -1, lanczositer = -1;
    int opt = 1;
    for (int i = 1; i + 1 < argc; i += 2)
    {
        std::string a = argv[i], v = argv[i + 1];
        if (a == "--N") N = std::stoul(v);
        else if (a == "--nev") nev = std::stoul(v);
        else if (a == "--nex") nex = std::stoul(v);
        else if (a == "--matrix") matrix = v;
        else if (a == "--tol") tol = std::stod(v);
        else if (a == "--deg") deg = std::stol(v);
        else if (a == "--opt") opt = std::stoi(v);
        else if (a == "--maxiter") maxiter = std::stoul(v);
        else if (a == "--numlanczos") numlanczos = std::stol(v);
        else if (a == "--lanczositer") lanczositer = std::stol(v);
        else
        {
            std::cerr << "unknown arg " << a << "\n";
            return 2;
        }
    }
    const std::size_t nevex = nev + nex;
#ifdef REF_PSEUDO
    const std::size_t ncols = 2 * nevex;
    using Backend = chase::Impl::ChASECPU<T, chase::matrix::PseudoHermitianMatrix<T>>;
#else
    const std::size_t ncols = nevex;
    using Backend = chase::Impl::ChASECPU<T>;
#endif
    std::vector<T> H0(N * N, T(0));
    if (matrix == "clement")
    {
        for (std::size_t i = 0; i + 1 < N; ++i)
            H0[i + 1 + N * i] = H0[i + N * (i + 1)] = T(R(std::sqrt(double(i * (N + 1 - i)))));
    }
    else if (matrix.rfind("file:", 0) == 0)
    {
        std::ifstream f(matrix.substr(5), std::ios::binary);
        if (!f)
        {
            std::cerr << "cannot open " << matrix << "\n";
            return 2;
        }
        f.read(reinterpret_cast<char*>(H0.data()), sizeof(T) * N * N);
    }
    else
    {
        std::cerr << "unknown matrix " << matrix << "\n";
        return 2;
    }

    struct Run
    {
        std::vector<T> H, V;
        std::vector<R> Lambda, resid;
        std::vector<std::string> calls;
        std::size_t swaps = 0;
    } run[2];

    for (int which = 0; which < 2; ++which)
    {
        Run& r = run[which];
        r.H = H0;
        r.V.assign(N * ncols, T(0));
        r.Lambda.assign(ncols, R(0));
        Backend single(N, nev, nex, r.H.data(), N, r.V.data(), N, r.Lambda.data());
        auto& config = single.GetConfig();
        if (tol > 0) config.SetTol(tol);
        if (deg > 0) config.SetDeg(deg);
        config.SetOpt(opt != 0);
        config.SetMaxIter(maxiter);
        config.SetApprox(false);
        if (numlanczos > 0) config.SetNumLanczos(numlanczos);
        if (lanczositer > 0) config.SetLanczosIter(lanczositer);
        TraceBackend<T> trace(&single);
        if (which == 0)
        {
#ifdef REF_PSEUDO
            chase::Solve_pseudo(&trace);
#else
            chase::Solve(&trace);
#endif
        }
        else
        {
            Adapter<T> adapter(&trace);
#ifdef REF_PSEUDO
            chase_b2::Solve_pseudo(&adapter);
#else
            chase_b2::Solve(&adapter);
#endif
        }
        r.calls = trace.calls;
        r.swaps = trace.swaps;
        r.resid.assign(single.GetResid(), single.GetResid() + nevex);
    }

    int bad = 0;
    const std::size_t nc = std::min(run[0].calls.size(), run[1].calls.size());
    for (std::size_t i = 0; i < nc && !bad; ++i)
        if (run[0].calls[i] != run[1].calls[i])
        {
            std::cerr << "call " << i << " differs:\n  reference driver: " << run[0].calls[i].substr(0, 300)
                      << "\n  new driver      : " << run[1].calls[i].substr(0, 300) << "\n";
            bad = 1;
        }
    if (!bad && run[0].calls.size() != run[1].calls.size())
    {
        std::cerr << "call counts differ: " << run[0].calls.size() << " vs " << run[1].calls.size() << "\n";
        bad = 1;
    }
    if (run[0].swaps != run[1].swaps)
    {
        std::cerr << "swap counts differ: " << run[0].swaps << " vs " << run[1].swaps << "\n";
        bad = 1;
    }
    if (std::memcmp(run[0].Lambda.data(), run[1].Lambda.data(), sizeof(R) * nevex) != 0)
    {
        std::cerr << "eigenvalues differ\n";
        bad = 1;
    }
    if (std::memcmp(run[0].resid.data(), run[1].resid.data(), sizeof(R) * nevex) != 0)
    {
        std::cerr << "residuals differ\n";
        bad = 1;
    }
    if (std::memcmp(run[0].V.data(), run[1].V.data(), sizeof(T) * N * nevex) != 0)
    {
        std::cerr << "eigenvectors differ\n";
        bad = 1;
    }
    std::size_t locks = 0;
    for (auto& c : run[0].calls)
        if (c.rfind("Lock ", 0) == 0)
            locks++;
    std::cout << "{\"identical\": " << (bad ? "false" : "true") << ", \"calls\": " << run[0].calls.size()
              << ", \"iterations\": " << locks << ", \"swaps\": " << run[0].swaps << "}\n";
    return bad;
}
