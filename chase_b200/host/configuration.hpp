// chase_b200 host layer — solver configuration (API mirror of the reference's
// chase::ChaseConfig<T>, algorithm/configuration.hpp:155-669; defaults from
// configuration.hpp:34-129: tol 1e-10 / 1e-5, deg 20 / 10, max_deg 36 / 18,
// lanczos_iter 25 / 12, num_lanczos 4, deg_extra 2, max_iter 25).
#pragma once
#include "types.hpp"

#include <cstddef>
#include <ostream>
#include <type_traits>

namespace chase
{

enum class LogLevel
{
    Error = 0,
    Warn = 1,
    Info = 2,
    Debug = 3,
    Trace = 4
};

template <class T>
class ChaseConfig
{
    static constexpr bool kDouble = std::is_same<Base<T>, double>::value;

public:
    ChaseConfig(std::size_t N, std::size_t nev, std::size_t nex)
        : N_(N), nev_(nev), nex_(nex), optimization_(true), approx_(false), deg_extra_(2), max_iter_(25),
          num_lanczos_(4), cholqr_(true), sym_check_(true), decaying_rate_(1.0f), cluster_aware_degrees_(true),
          upperb_scale_rate_(1.0f), log_level_(LogLevel::Info), log_rank_(0)
    {
        max_deg_ = kDouble ? 36 : 18;
        deg_ = kDouble ? 20 : 10;
        lanczos_iter_ = kDouble ? 25 : 12;
        tol_ = kDouble ? 1e-10 : 1e-5;
    }

    bool UseApprox() const { return approx_; }
    bool DoOptimization() const { return optimization_; }
    void SetApprox(bool flag) { approx_ = flag; }
    void SetOpt(bool flag) { optimization_ = flag; }

    std::size_t GetMaxDeg() const { return max_deg_; }
    void SetMaxDeg(std::size_t d)
    {
        max_deg_ = d;
        max_deg_ += max_deg_ % 2;
    }
    std::size_t GetDegExtra() const { return deg_extra_; }
    void SetDegExtra(std::size_t d) { deg_extra_ = d; }
    std::size_t GetMaxIter() const { return max_iter_; }
    void SetMaxIter(std::size_t m) { max_iter_ = m; }
    std::size_t GetDeg() const { return deg_; }
    void SetDeg(std::size_t d)
    {
        deg_ = d;
        deg_ += deg_ % 2;
    }
    double GetTol() const { return tol_; }
    void SetTol(double t) { tol_ = t; }
    std::size_t GetLanczosIter() const { return lanczos_iter_; }
    void SetLanczosIter(std::size_t m) { lanczos_iter_ = m; }
    std::size_t GetNumLanczos() const { return num_lanczos_; }
    void SetNumLanczos(std::size_t m) { num_lanczos_ = m; }

    std::size_t GetN() const { return N_; }
    std::size_t GetNev() const { return nev_; }
    std::size_t GetNex() const { return nex_; }

    void SetCholQR(bool flag) { cholqr_ = flag; }
    bool DoCholQR() { return cholqr_; }
    void EnableSymCheck(bool flag) { sym_check_ = flag; }
    bool DoSymCheck() { return sym_check_; }
    float GetDecayingRate() const { return decaying_rate_; }
    void SetDecayingRate(float r) { decaying_rate_ = r; }
    bool UseClusterAwareDegrees() const { return cluster_aware_degrees_; }
    void SetClusterAwareDegrees(bool flag) { cluster_aware_degrees_ = flag; }
    float GetUpperbScaleRate() const { return upperb_scale_rate_; }
    void SetUpperbScaleRate(float r) { upperb_scale_rate_ = r; }
    void SetVerbosity(LogLevel l) { log_level_ = l; }
    LogLevel GetLogLevel() const { return log_level_; }
    void SetLogRank(int r) { log_rank_ = r; }
    int GetLogRank() const { return log_rank_; }

    friend std::ostream& operator<<(std::ostream& os, const ChaseConfig<T>& c)
    {
        os << "ChASE config: N=" << c.N_ << " nev=" << c.nev_ << " nex=" << c.nex_ << " opt=" << c.optimization_
           << " approx=" << c.approx_ << " deg=" << c.deg_ << " max_deg=" << c.max_deg_ << " deg_extra=" << c.deg_extra_
           << " tol=" << c.tol_ << " max_iter=" << c.max_iter_ << " lanczos_iter=" << c.lanczos_iter_
           << " num_lanczos=" << c.num_lanczos_ << " cholqr=" << c.cholqr_ << "\n";
        return os;
    }

private:
    std::size_t N_, nev_, nex_;
    bool optimization_, approx_;
    std::size_t deg_, max_deg_, deg_extra_, max_iter_, lanczos_iter_, num_lanczos_;
    double tol_;
    bool cholqr_, sym_check_;
    float decaying_rate_;
    bool cluster_aware_degrees_;
    float upperb_scale_rate_;
    LogLevel log_level_;
    int log_rank_;
};

} // namespace chase
