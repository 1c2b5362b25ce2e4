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
//        [--tol t] [--deg d] [--opt 0|1] [--maxiter i] [--seq k] [--perturb p] [--repeat k]
//        [--vecs file] [--out result.json] [--dump-eigvecs file] [--initvecs-only file]
//        [--numlanczos n] [--lanczositer m]
//
// Built with -DREF_PSEUDO (chase_ref_cpu_p<z|c>) the backend is
// ChASECPU<T, PseudoHermitianMatrix<T>> driven by chase::Solve_pseudo
// (algorithm/algorithm.hpp:359): V holds 2 (nev+nex) columns, --matrix must be
// file:<path> (e.g. the reference's tests/linalg/internal/BSE_matrices/*.bin
// or a matrix written by oracle/chase_oracle.py: bse_matrix).
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
#ifdef REF_GPU
// chase_ref_gpu_<d|z>: the reference's OWN single-GPU backend (cuBLAS / cuSOLVER / cuRAND), built by `make refgpu`
// for the same-box GPU baseline of bench.py.  Its start vectors come from cuRAND, so iteration counts
// differ slightly from the CPU reference / from chase_b200's parity mode.
#include "Impl/chase_gpu/chase_gpu.hpp"
#endif

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

#include "trace_backend.hpp"

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
    long deg = -1, numlanczos = -1, lanczositer = -1;
    int opt = 1;
    bool fresh = false; // --repeat k: the SAME problem k times from fresh random vectors (warm-up + timed repeats)
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
        else if (a == "--repeat") { seq = std::stoul(v); fresh = true; }
        else if (a == "--perturb") perturb = std::stod(v);
        else if (a == "--out") out = v;
        else if (a == "--dump-eigvecs") dump_vecs = v;
        else if (a == "--vecs") vecs_in = v;
        else if (a == "--initvecs-only") initvecs_only = v;
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
    const std::size_t ncols = 2 * nevex; // [positive | K-conjugate] halves
    using Backend = chase::Impl::ChASECPU<T, chase::matrix::PseudoHermitianMatrix<T>>;
#elif defined(REF_GPU)
    const std::size_t ncols = nevex;
    using Backend = chase::Impl::ChASEGPU<T>;
#else
    const std::size_t ncols = nevex;
    using Backend = chase::Impl::ChASECPU<T>;
#endif
    std::vector<T> V(N * ncols), H(N * N, T(0));
    std::vector<R> Lambda(ncols);

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

    Backend single(N, nev, nex, H.data(), N, V.data(), N, Lambda.data());
    auto& config = single.GetConfig();
    if (tol > 0) config.SetTol(tol);
    if (deg > 0) config.SetDeg(deg);
    config.SetOpt(opt != 0);
    config.SetMaxIter(maxiter);
    config.SetApprox(false);
    if (numlanczos > 0) config.SetNumLanczos(numlanczos);
    if (lanczositer > 0) config.SetLanczosIter(lanczositer);

    if (!initvecs_only.empty())
    {
        single.initVecs(true);
        std::ofstream f(initvecs_only, std::ios::binary);
        f.write(reinterpret_cast<char*>(V.data()), sizeof(T) * N * ncols);
        return 0;
    }
    if (!vecs_in.empty())
    {
        std::ifstream f(vecs_in, std::ios::binary);
        f.read(reinterpret_cast<char*>(V.data()), sizeof(T) * N * ncols);
        config.SetApprox(true);
    }

    std::mt19937 gen(1337.0);
    std::normal_distribution<> d;

    std::ostringstream js;
    js << "{\"type\": \"" << (is_cplx<T>::value ? (sizeof(R) == 8 ? "z" : "c") : (sizeof(R) == 8 ? "d" : "s"))
       << "\", \"N\": " << N << ", \"nev\": " << nev << ", \"nex\": " << nex
       << ", \"matrix\": \"" << matrix << "\", \"tol\": " << fmt(config.GetTol())
       << ", \"deg\": " << config.GetDeg() << ", \"opt\": " << opt
#ifdef REF_PSEUDO
       << ", \"pseudo\": 1, \"numlanczos\": " << config.GetNumLanczos()
       << ", \"lanczositer\": " << config.GetLanczosIter()
#endif
       << ", \"problems\": [";

    for (std::size_t idx = 0; idx < seq; ++idx)
    {
        chase::PerformanceDecoratorChase<T> perf(&single);
        TraceBackend<T> trace(&perf);
        auto t0 = std::chrono::high_resolution_clock::now();
#ifdef REF_PSEUDO
        chase::Solve_pseudo(&trace);
#else
        chase::Solve(&trace);
#endif
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

        config.SetApprox(!fresh);
        if (idx + 1 < seq && !fresh)
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
        f.write(reinterpret_cast<char*>(V.data()), sizeof(T) * N * ncols);
    }
    return 0;
}
