// TEST INFRASTRUCTURE — not product code.
//
// TraceBackend: forwards every chase::ChaseBase<T> virtual to a wrapped REFERENCE backend and records the call,
// so that degree schedules / locking decisions can be compared call by call (oracle/ref_driver.cpp,
// oracle/xcheck_driver.cpp).  Include after the reference's algorithm/interface.hpp.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

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
        calls.push_back("HEMM_H2 " + std::to_string(nev) + " " + fmt(std::real(a)) + " " + fmt(std::real(b)) + " " +
                        fmt(std::real(g)) + " " + std::to_string(ol) + " " + std::to_string(orr));
        hemm_calls++;
        hemm_cols += 2 * (nev - ol - orr);
        in_->HEMM_H2(nev, a, b, g, ol, orr);
    }
    void ApplyKconjugate(std::size_t b) override
    {
        calls.push_back("ApplyK " + std::to_string(b));
        in_->ApplyKconjugate(b);
    }
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

