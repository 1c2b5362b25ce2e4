// chase_b200 host layer — per-phase timers, FLOP model and call trace.
//
// API mirror of the reference's ChasePerfData / PerformanceDecoratorChase<T>
// (algorithm/performance.hpp:43-516, 537-700): the decorator wraps any
// ChaseBase<T>, times InitVecs / Lanczos / Filter / QR / RR / Resid+Locking and
// counts iterations and filtered vectors with the reference's definitions
// (filtered vectors += nev - offset_right per HEMM call, performance.hpp:559-564;
// filter flops = 2 f N^2 * filtered_vecs, :248-260).
// Device work is asynchronous, so every time point first drains the device.
#pragma once
#include "interface.hpp"

#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

extern "C" int chase_b200_device_sync(void);

namespace chase
{

class ChasePerfData
{
public:
    enum TimePtrs
    {
        All = 0,
        InitVecs,
        Lanczos,
        Filter,
        ApplyKconjugate,
        Qr,
        Rr,
        Resid,
        End
    };

    ChasePerfData() { Reset(); }
    void Reset()
    {
        iter_count_ = 0;
        filtered_vecs_ = 0;
        blocksizes_.clear();
        early_locked_.clear();
        timings_.fill(0.0);
    }
    std::size_t get_iter_count() const { return iter_count_; }
    std::size_t get_filtered_vecs() const { return filtered_vecs_; }
    std::vector<std::chrono::duration<double>> get_timings() const
    {
        std::vector<std::chrono::duration<double>> t;
        for (double s : timings_)
            t.emplace_back(s);
        return t;
    }
    double seconds(TimePtrs p) const { return timings_[p]; }
    void add_iter_count(std::size_t a = 1) { iter_count_ += a; }
    void add_iter_blocksize(std::size_t b) { blocksizes_.push_back(b); }
    void add_filtered_vecs(std::size_t v) { filtered_vecs_ += v; }
    void set_early_locked_residuals(std::vector<double> v) { early_locked_ = std::move(v); }
    const std::vector<double>& early_locked() const { return early_locked_; }

    void start(TimePtrs p)
    {
        chase_b200_device_sync();
        t0_[p] = std::chrono::high_resolution_clock::now();
    }
    void stop(TimePtrs p)
    {
        chase_b200_device_sync();
        timings_[p] += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0_[p]).count();
    }

    // GFLOP by the reference's model; factor = 1 real, 4 complex
    double get_flops(std::size_t N, std::size_t lanczosIter, std::size_t numLanczos, int factor) const
    {
        double f = (double)lanczosIter * 2 * N * numLanczos * N;
        f += (double)lanczosIter * lanczosIter * numLanczos * numLanczos;
        for (auto block : blocksizes_)
        {
            const double b = (double)block, n = (double)N;
            f += 2. * n * b * b + 2. * b * b * b + 2. * n * b * b; // QR (CholQR-2 model)
            f += 2 * n * b * n + 2 * b * b * n + 4 * b * b * b + 2 * n * b * b; // RR
            f += 2 * n * b * n + 3 * b * n + n * b;                             // residuals
        }
        f += 2.0 * N * (double)filtered_vecs_ * N;
        return f * factor / 1e9;
    }
    double get_filter_flops(std::size_t N, int factor) const
    {
        return 2.0 * factor * (double)N * (double)filtered_vecs_ * (double)N / 1e9;
    }

    void print(std::size_t N, int factor, std::ostream& os = std::cout) const
    {
        os << " | Iterations | Vecs | All | InitVecs | Lanczos | Filter | QR | RR | Resid | Filter GFLOP | Filter "
              "TFLOP/s |\n | "
           << iter_count_ << " | " << filtered_vecs_;
        for (int p : {All, InitVecs, Lanczos, Filter, Qr, Rr, Resid})
            os << " | " << std::scientific << std::setprecision(3) << timings_[p];
        const double gf = get_filter_flops(N, factor);
        os << " | " << gf << " | " << (timings_[Filter] > 0 ? gf / timings_[Filter] / 1e3 : 0.0) << " |\n";
    }

private:
    std::size_t iter_count_, filtered_vecs_;
    std::vector<std::size_t> blocksizes_;
    std::vector<double> early_locked_;
    std::array<double, End> timings_;
    std::array<std::chrono::high_resolution_clock::time_point, End> t0_;
};

template <class T>
class PerformanceDecoratorChase : public ChaseBase<T>
{
    using R = Base<T>;

public:
    explicit PerformanceDecoratorChase(ChaseBase<T>* chase) : chase_(chase) {}
    ChasePerfData& GetPerfData() { return perf_; }
    // optional call trace in the format of oracle/ref_driver.cpp's TraceBackend
    void EnableTrace(bool on) { trace_on_ = on; }
    const std::vector<std::string>& Trace() const { return trace_; }
    std::size_t Swaps() const { return swaps_; }
    std::size_t HemmCalls() const { return hemm_calls_; }

    void Shift(T c, bool isunshift = false) override
    {
        if (trace_on_)
            trace_.push_back("Shift " + fmt(std::real(c)) + (isunshift ? " 1" : " 0"));
        chase_->Shift(c, isunshift);
    }
    void HEMM(std::size_t nev, T alpha, T beta, std::size_t offset_left, std::size_t offset_right = 0) override
    {
        if (trace_on_)
            trace_.push_back("HEMM " + std::to_string(nev) + " " + fmt(std::real(alpha)) + " " + fmt(std::real(beta)) +
                             " " + std::to_string(offset_left) + " " + std::to_string(offset_right));
        hemm_calls_++;
        chase_->HEMM(nev, alpha, beta, offset_left, offset_right);
        perf_.add_filtered_vecs(nev - offset_right);
    }
    void HEMM_H2(std::size_t nev, T alpha, T beta, T gamma, std::size_t ol, std::size_t orr = 0) override
    {
        if (trace_on_)
            trace_.push_back("HEMM_H2 " + std::to_string(nev) + " " + fmt(std::real(alpha)) + " " + fmt(std::real(beta)) +
                             " " + fmt(std::real(gamma)) + " " + std::to_string(ol) + " " + std::to_string(orr));
        hemm_calls_++;
        chase_->HEMM_H2(nev, alpha, beta, gamma, ol, orr);
        perf_.add_filtered_vecs(2 * (nev - orr)); // reference performance.hpp:565-570
    }
    void ApplyKconjugate(std::size_t block) override
    {
        if (trace_on_)
            trace_.push_back("ApplyK " + std::to_string(block));
        chase_->ApplyKconjugate(block);
    }
    void FilterPhaseStart() override
    {
        perf_.start(ChasePerfData::Filter);
        chase_->FilterPhaseStart();
    }
    void FilterPhaseEnd() override
    {
        chase_->FilterPhaseEnd();
        perf_.stop(ChasePerfData::Filter);
    }
    void QR(std::size_t fixednev, R cond) override
    {
        if (trace_on_)
            trace_.push_back("QR " + std::to_string(fixednev) + " " + fmt(cond));
        perf_.start(ChasePerfData::Qr);
        chase_->QR(fixednev, cond);
        perf_.stop(ChasePerfData::Qr);
    }
    void RR(R* ritzv, std::size_t block) override
    {
        perf_.start(ChasePerfData::Rr);
        chase_->RR(ritzv, block);
        perf_.stop(ChasePerfData::Rr);
        perf_.add_iter_count();
        perf_.add_iter_blocksize(block);
        if (trace_on_)
        {
            trace_.push_back("RR " + std::to_string(block));
            std::string v = "RITZV";
            for (std::size_t i = 0; i < block; ++i)
                v += " " + fmt(ritzv[i]);
            trace_.push_back(v);
        }
    }
    void Sort(R* a, R* b, R* c) override { chase_->Sort(a, b, c); }
    void Resd(R* ritzv, R* resd, std::size_t fixednev) override
    {
        perf_.start(ChasePerfData::Resid);
        chase_->Resd(ritzv, resd, fixednev);
        perf_.stop(ChasePerfData::Resid);
        if (trace_on_)
        {
            const std::size_t nevex = chase_->GetNev() + chase_->GetNex();
            std::string v = "RESID " + std::to_string(fixednev);
            for (std::size_t i = 0; i + fixednev < nevex; ++i)
                v += " " + fmt(resd[i]);
            trace_.push_back(v);
        }
    }
    void Lanczos(std::size_t m, R* upperb) override
    {
        perf_.start(ChasePerfData::Lanczos);
        chase_->Lanczos(m, upperb);
        perf_.stop(ChasePerfData::Lanczos);
        if (trace_on_)
            trace_.push_back("Lanczos1 " + std::to_string(m) + " " + fmt(*upperb));
    }
    void Lanczos(std::size_t M, std::size_t numvec, R* upperb, R* ritzv, R* Tau, R* ritzV) override
    {
        perf_.start(ChasePerfData::Lanczos);
        chase_->Lanczos(M, numvec, upperb, ritzv, Tau, ritzV);
        perf_.stop(ChasePerfData::Lanczos);
        if (trace_on_)
        {
            trace_.push_back("Lanczos " + std::to_string(M) + " " + std::to_string(numvec) + " " + fmt(*upperb));
            std::string t = "THETA";
            for (std::size_t i = 0; i < M * numvec; ++i)
                t += " " + fmt(ritzv[i]);
            trace_.push_back(t);
            t = "TAU";
            for (std::size_t i = 0; i < M * numvec; ++i)
                t += " " + fmt(Tau[i]);
            trace_.push_back(t);
        }
    }
    void LanczosDos(std::size_t idx, std::size_t m, T* ritzVc) override
    {
        if (trace_on_)
            trace_.push_back("LanczosDos " + std::to_string(idx) + " " + std::to_string(m));
        perf_.start(ChasePerfData::Lanczos);
        chase_->LanczosDos(idx, m, ritzVc);
        perf_.stop(ChasePerfData::Lanczos);
    }
    void Swap(std::size_t i, std::size_t j) override
    {
        swaps_++;
        chase_->Swap(i, j);
    }
    void Lock(std::size_t new_converged) override
    {
        if (trace_on_)
            trace_.push_back("Lock " + std::to_string(new_converged) + " swaps " + std::to_string(swaps_));
        chase_->Lock(new_converged);
    }
    bool checkSymmetryEasy() override { return chase_->checkSymmetryEasy(); }
    bool isSym() override { return chase_->isSym(); }
    bool checkPseudoHermicityEasy() override { return chase_->checkPseudoHermicityEasy(); }
    bool isPseudoHerm() override { return chase_->isPseudoHerm(); }
    void symOrHermMatrix(char uplo) override { chase_->symOrHermMatrix(uplo); }
    void Start() override
    {
        perf_.Reset();
        trace_.clear();
        swaps_ = hemm_calls_ = 0;
        if (trace_on_)
            trace_.push_back("Start");
        perf_.start(ChasePerfData::All);
        chase_->Start();
    }
    void End() override
    {
        if (trace_on_)
            trace_.push_back("End swaps " + std::to_string(swaps_));
        chase_->End();
        perf_.stop(ChasePerfData::All);
    }
    void initVecs(bool random) override
    {
        if (trace_on_)
            trace_.push_back(std::string("initVecs ") + (random ? "1" : "0"));
        perf_.start(ChasePerfData::InitVecs);
        chase_->initVecs(random);
        perf_.stop(ChasePerfData::InitVecs);
    }
    std::size_t GetN() const override { return chase_->GetN(); }
    std::size_t GetNev() override { return chase_->GetNev(); }
    std::size_t GetNex() override { return chase_->GetNex(); }
    std::size_t GetLanczosIter() override { return chase_->GetLanczosIter(); }
    std::size_t GetNumLanczos() override { return chase_->GetNumLanczos(); }
    std::size_t GetRitzvBlockSize() const override { return chase_->GetRitzvBlockSize(); }
    R* GetRitzv() override { return chase_->GetRitzv(); }
    R* GetResid() override { return chase_->GetResid(); }
    ChaseConfig<T>& GetConfig() override { return chase_->GetConfig(); }
    int get_nprocs() override { return chase_->get_nprocs(); }
    int get_rank() override { return chase_->get_rank(); }
    void set_early_locked_residuals(std::vector<R> v) override
    {
        if (trace_on_)
            trace_.push_back("early_locked " + std::to_string(v.size()));
        perf_.set_early_locked_residuals(std::vector<double>(v.begin(), v.end()));
        chase_->set_early_locked_residuals(v);
    }
    void Output(LogLevel l, std::string s, const char* c = "algorithm") override { chase_->Output(l, s, c); }

private:
    static std::string fmt(double v)
    {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%.17g", v);
        return buf;
    }
    ChaseBase<T>* chase_;
    ChasePerfData perf_;
    bool trace_on_ = false;
    std::vector<std::string> trace_;
    std::size_t swaps_ = 0, hemm_calls_ = 0;
};

} // namespace chase
