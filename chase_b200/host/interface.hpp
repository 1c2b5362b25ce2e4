// chase_b200 host layer — the backend contract.  Same virtuals, names, argument
// meaning and defaults as the reference's chase::ChaseBase<T>
// (algorithm/interface.hpp:46-434) so that a backend written for one driver
// works with the other.
#pragma once
#include "configuration.hpp"
#include "types.hpp"

#include <cstddef>
#include <string>
#include <vector>

namespace chase
{

template <class T>
class ChaseBase
{
public:
    virtual ~ChaseBase() = default;

    virtual void Shift(T c, bool isunshift = false) = 0;                                   // interface.hpp:60
    virtual void HEMM(std::size_t nev, T alpha, T beta, std::size_t offset_left,
                      std::size_t offset_right = 0) = 0;                                   // :78
    virtual void HEMM_H2(std::size_t nev, T alpha, T beta, T gamma, std::size_t offset_left,
                         std::size_t offset_right = 0) = 0;                                // :86
    virtual void ApplyKconjugate(std::size_t block) = 0;                                   // :100
    virtual void FilterPhaseStart() {}                                                     // :106
    virtual void FilterPhaseEnd() {}                                                       // :112
    virtual void QR(std::size_t fixednev, Base<T> cond) = 0;                               // :124
    virtual void RR(Base<T>* ritzv, std::size_t block) = 0;                                // :135
    virtual void Sort(Base<T>* ritzv, Base<T>* residLast, Base<T>* resid) = 0;             // :147
    virtual void Resd(Base<T>* ritzv, Base<T>* resd, std::size_t fixednev) = 0;            // :159
    virtual void Lanczos(std::size_t m, Base<T>* upperb) = 0;                              // :170
    virtual void Lanczos(std::size_t M, std::size_t numvec, Base<T>* upperb, Base<T>* ritzv, Base<T>* Tau,
                         Base<T>* ritzV) = 0;                                              // :185
    virtual void LanczosDos(std::size_t idx, std::size_t m, T* ritzVc) = 0;                // :198
    virtual void Swap(std::size_t i, std::size_t j) = 0;                                   // :208
    virtual void Lock(std::size_t new_converged) = 0;                                      // :217
    virtual bool checkSymmetryEasy() = 0;                                                  // :225
    virtual bool isSym() = 0;                                                              // :233
    virtual bool checkPseudoHermicityEasy() = 0;                                           // :241
    virtual bool isPseudoHerm() = 0;                                                       // :249
    virtual void symOrHermMatrix(char uplo) = 0;                                           // :259
    virtual void Start() = 0;                                                              // :265
    virtual void End() = 0;                                                                // :271
    virtual void initVecs(bool random) = 0;                                                // :280
    virtual void ReinitColumns(std::size_t, std::size_t const*, std::size_t) {}            // :292

    virtual std::size_t GetN() const = 0;
    virtual std::size_t GetNev() = 0;
    virtual std::size_t GetNex() = 0;
    virtual std::size_t GetLanczosIter() = 0;
    virtual std::size_t GetNumLanczos() = 0;
    virtual std::size_t GetRitzvBlockSize() const = 0;
    virtual Base<T>* GetRitzv() = 0;
    virtual Base<T>* GetResid() = 0;
    virtual ChaseConfig<T>& GetConfig() = 0;
    virtual int get_nprocs() = 0;
    virtual int get_rank() = 0;
    virtual void set_early_locked_residuals(std::vector<Base<T>>) {}                       // :416
    virtual void Output(LogLevel, std::string, const char* = "algorithm") {}               // :431
};

} // namespace chase
