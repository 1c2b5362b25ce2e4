// TEST INFRASTRUCTURE — not product code.
//
// Generic command-line driver around the UNMODIFIED reference CPU backend
// (chase::Impl::ChASECPU, /root/reference/Impl/chase_cpu/chase_cpu.hpp:67) and
// the reference's own algorithm driver (chase::Solve,
// /root/reference/algorithm/algorithm.hpp:345).  It is compiled from the
// reference sources where they lie (see oracle/Makefile) into
// oracle/_ref/chase_ref_cpu_<type>; nothing from /root/reference is copied
// into this repository.
//
// It is used (a) to generate the golden fixtures under tests/golden/
// (tests/golden/make_golden.py), (b) to pin oracle/chase_oracle.py, and
// (c) as the `cpu_baseline` / `--impl reference` arm of bench.py.
//
// Every ChaseBase<T> call the reference algorithm makes is recorded through
// TraceBackend so that the degree schedule / locking decisions can be compared
// call by call with the B200 backend.
//
// Usage:
//   chase_ref_cpu_<d|z|s|c> --N n --nev k --nex x --matrix clement|uniform|uniform_dense|file:<path>
//        [--tol t] [--deg d] [--opt 0|1] [--maxiter i] [--seq k] [--perturb p]
//        [--vecs file] [--out result.json] [--dump-eigvecs file] [--initvecs-only file]
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "algorithm/performance.hpp"
#include "Impl/chase_cpu/chase_cpu.hpp"

#ifndef REF_T
#define REF_T double
#endif
using T = REF_T;
using R = chase::Base<T>;

template <typename U>
struct is_cplx : std::false_type
{
};
template <typename U>
struct is_cplx<std::complex<U>> : std::true_type
{
};

static std::string fmt(double v)
{
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.17g", v);
    return buf;
}

// Forwards every virtual to the wrapped reference backend and logs the call.
template <class S>
class TraceBackend : public chase::ChaseBase<S>
{
    using B = chase::Base<S>;

public:
    explicit TraceBackend(chase::ChaseBase<S>* inner) : in_(inner) {}
    std::vector<std::string> calls;
    std::size_t swaps = 0, hemm_cols = 0, hemm_calls = 0;

    void Shift(S c, bool un = false) override
    {
        calls.push_back("Shift " + fmt(std::real(c)) + (un ? " 1" : " 0"));
        in_->Shift(c, un);
    }
    void HEMM(std::size_t nev, S a, S b, std::size_t ol,
              std::size_t orr = 0) override
    {
        calls.push_back("HEMM " + std::to_string(nev) + " " +
                        fmt(std::real(a)) + " " + fmt(std::real(b)) + " " +
                        std::to_string(ol) + " " + std::to_string(orr));
        hemm_calls++;
        hemm_cols += nev - orr;
        in_->HEMM(nev, a, b, ol, orr);
    }
    void HEMM_H2(std::size_t nev, S a, S b, S g, std::size_t ol,
                 std::size_t orr = 0) override
    {
        in_->HEMM_H2(nev, a, b, g, ol, orr);
    }
    void ApplyKconjugate(std::size_t b) override { in_->ApplyKconjugate(b); }
    void FilterPhaseStart() override { in_->FilterPhaseStart(); }
    void FilterPhaseEnd() override { in_->FilterPhaseEnd(); }
    void QR(std::size_t f, B cond) override
    {
        calls.push_back("QR " + std::to_string(f) + " " + fmt(cond));
        in_->QR(f, cond);
    }
    void RR(B* ritzv, std::size_t block) override
    {
        in_->RR(ritzv, block);
        std::string s = "RR " + std::to_string(block);
        calls.push_back(s);
        std::string v = "RITZV";
        for (std::size_t i = 0; i < block; ++i)
            v += " " + fmt(ritzv[i]);
        calls.push_back(v);
    }
    void Sort(B* a, B* b, B* c) override { in_->Sort(a, b, c); }
    void Resd(B* ritzv, B* resd, std::size_t f) override
    {
        in_->Resd(ritzv, resd, f);
        std::size_t nevex = in_->GetNev() + in_->GetNex();
        std::string v = "RESID " + std::to_string(f);
        for (std::size_t i = 0; i + f < nevex; ++i)
            v += " " + fmt(resd[i]);
        calls.push_back(v);
    }
    void Lanczos(std::size_t m, B* ub) override
    {
        in_->Lanczos(m, ub);
        calls.push_back("Lanczos1 " + std::to_string(m) + " " + fmt(*ub));
    }
    void Lanczos(std::size_t M, std::size_t nv, B* ub, B* rv, B* tau,
                 B* rV) override
    {
        in_->Lanczos(M, nv, ub, rv, tau, rV);
        std::string s = "Lanczos " + std::to_string(M) + " " +
                        std::to_string(nv) + " " + fmt(*ub);
        calls.push_back(s);
        std::string t = "THETA";
        for (std::size_t i = 0; i < M * nv; ++i)
            t += " " + fmt(rv[i]);
        calls.push_back(t);
        t = "TAU";
        for (std::size_t i = 0; i < M * nv; ++i)
            t += " " + fmt(tau[i]);
        calls.push_back(t);
    }
    void LanczosDos(std::size_t idx, std::size_t m, S* rvc) override
    {
        calls.push_back("LanczosDos " + std::to_string(idx) + " " +
                        std::to_string(m));
        in_->LanczosDos(idx, m, rvc);
    }
    void Swap(std::size_t i, std::size_t j) override
    {
        swaps++;
        in_->Swap(i, j);
    }
    void Lock(std::size_t n) override
    {
        calls.push_back("Lock " + std::to_string(n) + " swaps " +
                        std::to_string(swaps));
        in_->Lock(n);
    }
    bool checkSymmetryEasy() override { return in_->checkSymmetryEasy(); }
    bool isSym() override { return in_->isSym(); }
    bool checkPseudoHermicityEasy() override
    {
        return in_->checkPseudoHermicityEasy();
    }
    bool isPseudoHerm() override { return in_->isPseudoHerm(); }
    void symOrHermMatrix(char u) override { in_->symOrHermMatrix(u); }
    void Start() override
    {
        calls.push_back("Start");
        in_->Start();
    }
    void End() override
    {
        calls.push_back("End swaps " + std::to_string(swaps));
        in_->End();
    }
    void initVecs(bool random) override
    {
        calls.push_back(std::string("initVecs ") + (random ? "1" : "0"));
        in_->initVecs(random);
    }
    std::size_t GetN() const override { return in_->GetN(); }
    std::size_t GetNev() override { return in_->GetNev(); }
    std::size_t GetNex() override { return in_->GetNex(); }
    std::size_t GetLanczosIter() override { return in_->GetLanczosIter(); }
    std::size_t GetNumLanczos() override { return in_->GetNumLanczos(); }
    std::size_t GetRitzvBlockSize() const override
    {
        return in_->GetRitzvBlockSize();
    }
    B* GetRitzv() override { return in_->GetRitzv(); }
    B* GetResid() override { return in_->GetResid(); }
    chase::ChaseConfig<S>& GetConfig() override { return in_->GetConfig(); }
    int get_nprocs() override { return in_->get_nprocs(); }
    int get_rank() override { return in_->get_rank(); }
    void set_early_locked_residuals(std::vector<B> v) override
    {
        calls.push_back("early_locked " + std::to_string(v.size()));
        in_->set_early_locked_residuals(v);
    }
#ifdef CHASE_OUTPUT
    void Output(chase::LogLevel l, std::string s,
                const char* c = "algorithm") override
    {
        in_->Output(l, s, c);
    }
#endif
private:
    chase::ChaseBase<S>* in_;
};

template <typename U>
static U cj_(const U& x) { return x; }
template <typename U>
static std::complex<U> cj_(const std::complex<U>& x) { return std::conj(x); }
static T cj(const T& x) { return cj_(x); }

template <typename U>
struct mk { static U f(double re, double) { return U(re); } };
template <typename U>
struct mk<std::complex<U>> { static std::complex<U> f(double re, double im) { return std::complex<U>(U(re), U(im)); } };
static T make_T(double re, double im) { return mk<T>::f(re, im); }

// Dense uniform-spectrum generator used by the CPU baseline timing:
// A = Q diag(lambda) Q^H, Q = product of 3 Householder reflectors with
// deterministic unit vectors; lambda_k as in the reference's --isMatGen
// generator (examples/2_input_output/2_input_output.cpp:250-262).
static void gen_uniform(std::vector<T>& H, std::size_t N, bool dense)
{
    const double dmax = 100, eps = 1e-4;
    std::fill(H.begin(), H.end(), T(0));
    for (std::size_t k = 0; k < N; ++k)
        H[k + N * k] = make_T(dmax * (eps + double(k) * (1.0 - eps) / double(N)), 0);
    if (!dense)
        return;
    std::mt19937_64 g(1337);
    std::normal_distribution<double> nd;
    std::vector<T> v(N), w(N);
    for (int p = 0; p < 3; ++p)
    {
        double nrm = 0;
        for (std::size_t i = 0; i < N; ++i)
        {
            double re = nd(g), im = is_cplx<T>::value ? nd(g) : 0.0;
            v[i] = make_T(re, im);
            nrm += re * re + im * im;
        }
        nrm = std::sqrt(nrm);
        for (auto& x : v)
            x /= (R)nrm;
        // A <- (I-2vv^H) A (I-2vv^H) = A - 2 v (A v)^H - 2 (A v) v^H + 4 (v^H A v) v v^H
#pragma omp parallel for
        for (std::size_t i = 0; i < N; ++i)
        {
            T s = T(0);
            for (std::size_t j = 0; j < N; ++j)
                s += H[i + N * j] * v[j];
            w[i] = s;
        }
        T vAv = T(0);
        for (std::size_t i = 0; i < N; ++i)
            vAv += cj(v[i]) * w[i];
#pragma omp parallel for
        for (std::size_t j = 0; j < N; ++j)
            for (std::size_t i = 0; i < N; ++i)
                H[i + N * j] += -(R)2 * v[i] * cj(w[j]) -
                                (R)2 * w[i] * cj(v[j]) +
                                (R)4 * vAv * v[i] * cj(v[j]);
    }
}

int main(int argc, char** argv)
{
    std::size_t N = 1001, nev = 100, nex = 40, maxiter = 25, seq = 1;
    std::string matrix = "clement", out, dump_vecs, vecs_in, initvecs_only;
    double tol = -1, perturb = 1e-4;
    long deg = -1;
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
        else if (a == "--seq") seq = std::stoul(v);
        else if (a == "--perturb") perturb = std::stod(v);
        else if (a == "--out") out = v;
        else if (a == "--dump-eigvecs") dump_vecs = v;
        else if (a == "--vecs") vecs_in = v;
        else if (a == "--initvecs-only") initvecs_only = v;
        else
        {
            std::cerr << "unknown arg " << a << "\n";
            return 2;
        }
    }
    const std::size_t nevex = nev + nex;
    std::vector<T> V(N * nevex), H(N * N, T(0));
    std::vector<R> Lambda(nevex);

    if (matrix == "clement")
    {
        // tests/noinput.cpp:67-74 of the reference
        for (std::size_t i = 0; i < N; ++i)
        {
            H[i + N * i] = 0;
            if (i != N - 1)
            {
                H[i + 1 + N * i] = make_T(std::sqrt(double(i * (N + 1 - i))), 0);
                H[i + N * (i + 1)] = make_T(std::sqrt(double(i * (N + 1 - i))), 0);
            }
        }
    }
    else if (matrix == "uniform")
        gen_uniform(H, N, false);
    else if (matrix == "uniform_dense")
        gen_uniform(H, N, true);
    else if (matrix.rfind("file:", 0) == 0)
    {
        std::ifstream f(matrix.substr(5), std::ios::binary);
        if (!f)
        {
            std::cerr << "cannot open " << matrix << "\n";
            return 2;
        }
        f.read(reinterpret_cast<char*>(H.data()), sizeof(T) * N * N);
    }
    else
    {
        std::cerr << "unknown matrix " << matrix << "\n";
        return 2;
    }

    chase::Impl::ChASECPU<T> single(N, nev, nex, H.data(), N, V.data(), N,
                                    Lambda.data());
    auto& config = single.GetConfig();
    if (tol > 0) config.SetTol(tol);
    if (deg > 0) config.SetDeg(deg);
    config.SetOpt(opt != 0);
    config.SetMaxIter(maxiter);
    config.SetApprox(false);

    if (!initvecs_only.empty())
    {
        single.initVecs(true);
        std::ofstream f(initvecs_only, std::ios::binary);
        f.write(reinterpret_cast<char*>(V.data()), sizeof(T) * N * nevex);
        return 0;
    }
    if (!vecs_in.empty())
    {
        std::ifstream f(vecs_in, std::ios::binary);
        f.read(reinterpret_cast<char*>(V.data()), sizeof(T) * N * nevex);
        config.SetApprox(true);
    }

    std::mt19937 gen(1337.0);
    std::normal_distribution<> d;

    std::ostringstream js;
    js << "{\"type\": \"" << (is_cplx<T>::value ? (sizeof(R) == 8 ? "z" : "c") : (sizeof(R) == 8 ? "d" : "s"))
       << "\", \"N\": " << N << ", \"nev\": " << nev << ", \"nex\": " << nex
       << ", \"matrix\": \"" << matrix << "\", \"tol\": " << fmt(config.GetTol())
       << ", \"deg\": " << config.GetDeg() << ", \"opt\": " << opt
       << ", \"problems\": [";

    for (std::size_t idx = 0; idx < seq; ++idx)
    {
        chase::PerformanceDecoratorChase<T> perf(&single);
        TraceBackend<T> trace(&perf);
        auto t0 = std::chrono::high_resolution_clock::now();
        chase::Solve(&trace);
        auto t1 = std::chrono::high_resolution_clock::now();
        auto& pd = perf.GetPerfData();
        // print() is the only public path that folds the time points into
        // the timings vector (performance.hpp:352-356)
        pd.print(N, config.GetLanczosIter(), config.GetNumLanczos());
        auto tm = pd.get_timings();
        R* resid = single.GetResid();
        if (idx) js << ", ";
        js << "{\"iterations\": " << pd.get_iter_count()
           << ", \"filtered_vecs\": " << pd.get_filtered_vecs()
           << ", \"hemm_calls\": " << trace.hemm_calls
           << ", \"swaps\": " << trace.swaps << ", \"wall_s\": "
           << fmt(std::chrono::duration<double>(t1 - t0).count())
           << ", \"timings\": {";
        const char* names[8] = {"All", "InitVecs", "Lanczos", "Filter",
                                "ApplyKconjugate", "QR", "RR", "Resid"};
        for (int k = 0; k < 8; ++k)
            js << (k ? ", " : "") << "\"" << names[k] << "\": "
               << fmt(tm[k].count());
        js << "}, \"gflop_total\": "
           << pd.get_flops(N, config.GetLanczosIter(), config.GetNumLanczos())
           << ", \"gflop_filter\": " << pd.get_filter_flops(N)
           << ", \"ritzv\": [";
        for (std::size_t i = 0; i < nevex; ++i)
            js << (i ? ", " : "") << fmt(Lambda[i]);
        js << "], \"resid\": [";
        for (std::size_t i = 0; i < nevex; ++i)
            js << (i ? ", " : "") << fmt(resid[i]);
        js << "], \"trace\": [";
        for (std::size_t i = 0; i < trace.calls.size(); ++i)
            js << (i ? ", " : "") << "\"" << trace.calls[i] << "\"";
        js << "]}";

        std::cerr << "problem " << idx << ": iterations " << pd.get_iter_count()
                  << " filtered_vecs " << pd.get_filtered_vecs() << " All "
                  << tm[0].count() << " s Filter " << tm[3].count() << " s\n";

        config.SetApprox(true);
        if (idx + 1 < seq)
        {
            // element-wise Hermitian perturbation, tests/noinput.cpp:120-134
            for (std::size_t i = 1; i < N; ++i)
                for (std::size_t j = 1; j < i; ++j)
                {
                    double re = d(gen);
                    double im = is_cplx<T>::value ? d(gen) : 0.0;
                    T e = make_T(re * perturb, im * perturb);
                    H[j + N * i] += e;
                    H[i + N * j] += cj(e);
                }
        }
    }
    js << "]}\n";
    if (!out.empty())
    {
        std::ofstream f(out);
        f << js.str();
    }
    else
        std::cout << js.str();
    if (!dump_vecs.empty())
    {
        std::ofstream f(dump_vecs, std::ios::binary);
        f.write(reinterpret_cast<char*>(V.data()), sizeof(T) * N * nevex);
    }
    return 0;
}
