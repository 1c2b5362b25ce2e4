// chase_b200 host layer — Chebyshev-filtered subspace iteration driver.
//
// Written from scratch; decision logic (bounds, degree schedule, locking,
// condition estimate, final ordering) follows the reference's
// chase::Algorithm<T> step by step so that iteration counts and HEMM schedules
// are identical on identical inputs:
//   solve         /root/reference/algorithm/algorithm.inc:1376-1788
//   filter        :942-1009          calc_degrees :136-193
//   locking       :519-578           lanczos/DoS  :1067-1214
// Pseudo-Hermitian (BSE) problems — filter on H^2, K-conjugate pairs, symmetric
// locking — are driven by solve_pseudo:
//   solve_pseudo  :1834-2220         filter_H2    :1012-1064
//   calc_degrees_pseudo_H2 :196-317  detect_eigenvalue_clusters :19-135
//   locking_pseudo_v3 :730-816       lanczos_for_H2 :1217-1373
#pragma once
#include "interface.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <iostream>
#include <limits>
#include <numeric>
#include <sstream>
#include <vector>

namespace chase
{

template <class T>
class Algorithm
{
    using R = Base<T>;

public:
    // Chebyshev ratio: max |t -+ sqrt(t^2 - 1)|
    static R cheb_rho(R t, bool abs_arg)
    {
        const R q = abs_arg ? std::sqrt(std::abs(t * t - 1)) : std::sqrt(t * t - 1);
        return std::max(std::abs(t - q), std::abs(t + q));
    }

    static std::size_t calc_degrees(ChaseBase<T>* single, std::size_t unconverged, std::size_t nex, R upperb, R lowerb,
                                    R tol, R* ritzv, R* resid, std::size_t* degrees, std::size_t locked)
    {
        ChaseConfig<T>& conf = single->GetConfig();
        const R c = (upperb + lowerb) / 2;
        const R e = (upperb - lowerb) / 2;
        for (std::size_t i = 0; i < unconverged - nex; ++i)
        {
            const R t = (ritzv[i] - c) / e;
            const R rho = cheb_rho(t, true);
            degrees[i] = (std::size_t)std::ceil(std::abs(std::log(resid[i] / tol) / std::log(rho)));
            if (std::is_same<R, float>::value)
                degrees[i] = std::max(degrees[i], std::size_t(8));
            degrees[i] = std::min(degrees[i] + conf.GetDegExtra(), conf.GetMaxDeg());
        }
        for (std::size_t i = unconverged - nex; i < unconverged; ++i)
            degrees[i] = degrees[unconverged - 1 - nex];
        for (std::size_t i = 0; i < unconverged; ++i)
            degrees[i] += degrees[i] % 2;
        // ascending by degree with the reference's exchange sort (the order of
        // equal-degree columns is part of the parity contract)
        for (std::size_t j = 0; j + 1 < unconverged; ++j)
            for (std::size_t k = j; k < unconverged; ++k)
                if (degrees[k] < degrees[j])
                {
                    std::swap(degrees[k], degrees[j]);
                    std::swap(ritzv[k], ritzv[j]);
                    std::swap(resid[k], resid[j]);
                    single->Swap(k + locked, j + locked);
                }
        return degrees[unconverged - 1];
    }

    static std::size_t filter(ChaseBase<T>* single, std::size_t unprocessed, std::size_t deg, std::size_t* degrees,
                              R lambda_1, R lower, R upper)
    {
        const R c = (upper + lower) / 2;
        const R e = (upper - lower) / 2;
        const R sigma_1 = e / (lambda_1 - c);
        R sigma = sigma_1, sigma_new;
        std::size_t offset = 0, num_mult = 0, Av = 0;

        single->FilterPhaseStart();
        single->Shift(T(-c));
        T alpha = T(sigma_1 / e);
        T beta = T(0.0);
        single->HEMM(unprocessed, alpha, beta, offset);
        Av += unprocessed;
        num_mult++;
        while (*degrees <= num_mult) // cannot trigger: degrees are >= 2
        {
            degrees++;
            unprocessed--;
            offset++;
        }
        for (std::size_t i = 2; i <= deg; ++i)
        {
            sigma_new = R(1.0 / (2.0 / sigma_1 - sigma));
            alpha = T(R(2.0 * sigma_new / e));
            beta = T(-sigma * sigma_new);
            single->HEMM(unprocessed, alpha, beta, offset);
            sigma = sigma_new;
            Av += unprocessed;
            num_mult++;
            while (unprocessed != 0 && *degrees <= num_mult)
            {
                degrees++;
                unprocessed--;
                offset++;
            }
        }
        single->Shift(T(c), true);
        single->FilterPhaseEnd();
        return Av;
    }

    static std::size_t locking(ChaseBase<T>* single, std::size_t unconverged, R tol, R* Lritzv, R* resid, R* residLast,
                               std::vector<R>* early, std::size_t locked)
    {
        std::vector<int> index(unconverged);
        std::iota(index.begin(), index.end(), 0);
        std::sort(index.begin(), index.end(), [&](const int& a, const int& b) { return Lritzv[a] < Lritzv[b]; });
        std::size_t converged = 0;
        for (std::size_t k = 0; k < unconverged; ++k)
        {
            const std::size_t j = (std::size_t)index[k];
            const bool early_lock = single->isSym() && resid[j] >= residLast[j] && resid[j] < 100.0 * tol;
            if (resid[j] <= tol || early_lock)
            {
                if (resid[j] > tol && early_lock)
                    early->push_back(resid[j]);
                if (j != converged)
                {
                    std::swap(resid[j], resid[converged]);
                    std::swap(residLast[j], residLast[converged]);
                    std::swap(Lritzv[j], Lritzv[converged]);
                    single->Swap(j + locked, converged + locked);
                }
                converged++;
            }
        }
        return converged;
    }

    // Spectral-bound estimation; returns the number of DoS vectors extracted.
    static std::size_t lanczos(ChaseBase<T>* single, int N, int numvec, int m, int nevex, R* upperb, bool mode,
                               R* ritzv_)
    {
        assert(m >= 1);
        if (!mode)
        {
            single->Lanczos(m, upperb);
            return 0;
        }
        std::vector<R> Theta((std::size_t)numvec * m, R(0)), Tau((std::size_t)numvec * m, R(0));
        std::vector<R> ritzV((std::size_t)m * m, R(0));
        R lowerb = R(0), lambda;
        single->Lanczos(m, numvec, upperb, Theta.data(), Tau.data(), ritzV.data());

        std::vector<double> ThetaSorted(Theta.begin(), Theta.end());
        std::sort(ThetaSorted.begin(), ThetaSorted.end());
        lambda = R(ThetaSorted[0]);

        double curr, prev = 0;
        const double sigma = 0.25;
        const double threshold = 2 * sigma * sigma / 10;
        const double search = static_cast<double>(nevex) / static_cast<double>(N);
        const auto G = [&](double x) -> double { return 0.5 * (1 + std::erf(x / std::sqrt(2 * sigma * sigma))); };
        const int bound = m; // halved only for pseudo-Hermitian problems
        for (int i = 0; i < numvec * bound - 1; ++i)
        {
            curr = 0;
            for (int j = 0; j < numvec * bound; ++j)
            {
                if (ThetaSorted[i] < (Theta[j] - threshold))
                    curr += 0;
                else if (ThetaSorted[i] > (Theta[j] + threshold))
                    curr += Tau[j] * 1;
                else
                    curr += Tau[j] * G(ThetaSorted[i] - Theta[j]);
            }
            curr = curr / numvec;
            if (curr > search)
            {
                if (std::abs(curr - search) < std::abs(prev - search))
                    lowerb = R((i + 1 < numvec * bound) ? ThetaSorted[i + 1] : ThetaSorted[i]);
                else
                    lowerb = R(ThetaSorted[i]);
                break;
            }
            prev = curr;
        }

        int idx = 0;
        for (int i = 0; i < m; ++i)
            if (Theta[(std::size_t)(numvec - 1) * m + i] > lowerb)
            {
                idx = i - 1;
                break;
            }
        if (idx > 0)
        {
            std::vector<T> ritzVc((std::size_t)m * m);
            for (std::size_t i = 0; i < (std::size_t)m * m; ++i)
                ritzVc[i] = T(ritzV[i]);
            single->LanczosDos(idx, m, ritzVc.data());
        }
        for (int i = 0; i < idx; ++i)
            ritzv_[i] = Theta[(std::size_t)(numvec - 1) * m + i];
        for (int i = std::max(idx, 0); i < nevex - 1; ++i)
            ritzv_[i] = lambda;
        ritzv_[nevex - 1] = lowerb;
        for (int i = 1; i < idx; ++i)
        {
            const int j = i * (nevex / idx);
            single->Swap(i, j);
            std::swap(ritzv_[i], ritzv_[j]);
        }
        return (std::size_t)std::max(idx, 0);
    }

    static void solve(ChaseBase<T>* single)
    {
        ChaseConfig<T>& config = single->GetConfig();
        single->Start();

        const std::size_t N = config.GetN();
        const std::size_t nev = config.GetNev();
        const std::size_t nex = config.GetNex();
        const std::size_t num_lanczos = config.GetNumLanczos();
        R* resid_ = single->GetResid();
        R* ritzv_ = single->GetRitzv();
        const double tol = config.GetTol();
        const std::size_t nevex = nev + nex;
        std::size_t unconverged = nevex;
        R lowerb, upperb, lambda, cond;

        std::vector<std::size_t> degrees_(nevex);
        std::vector<R> residLast_(nevex);
        std::vector<R> early_locked;
        for (std::size_t i = 0; i < nevex; ++i)
        {
            residLast_[i] = std::numeric_limits<R>::max();
            resid_[i] = std::numeric_limits<R>::max();
        }
        std::size_t deg = config.GetDeg();
        deg += deg % 2;
        std::size_t* degrees = degrees_.data();
        R* ritzv = ritzv_;
        R* resid = resid_;
        R* residLast = residLast_.data();
        deg = std::min(deg, config.GetMaxDeg());
        for (std::size_t i = 0; i < nevex; ++i)
            degrees[i] = deg;

        const bool random = !config.UseApprox();
        single->initVecs(random);
        if (random)
            single->QR(0, R(1.0));

        std::size_t lanczos_iter = std::min(nevex, std::min(N / 2, config.GetLanczosIter()));
        if (2.0 * (lanczos_iter / 2) < lanczos_iter)
        {
            config.SetLanczosIter(lanczos_iter - 1);
            lanczos_iter = config.GetLanczosIter();
        }
        lanczos(single, (int)N, (int)num_lanczos, (int)lanczos_iter, (int)nevex, &upperb, random,
                random ? ritzv : nullptr);

        std::size_t locked = 0, iteration = 0;
        lowerb = *std::max_element(ritzv, ritzv + unconverged);
        lambda = *std::min_element(ritzv_, ritzv_ + nevex);
        lowerb = lowerb * config.GetDecayingRate();
        std::size_t new_converged = 0;

        while (unconverged > nex && iteration < config.GetMaxIter())
        {
            std::size_t cnt = 0;
            for (; cnt < unconverged; ++cnt)
                if (resid[cnt] > R(5e-1))
                    break;
            if (single->isSym() && cnt == unconverged)
                lowerb = ritzv[unconverged - 1];

            if (config.GetLogLevel() >= LogLevel::Debug)
            {
                std::ostringstream oss;
                oss << std::scientific << "iteration: " << iteration << "\t" << lambda << "\t" << lowerb << "\t"
                    << upperb << "\t" << unconverged << "\n";
                single->Output(LogLevel::Debug, oss.str(), "algorithm");
            }
            if (lowerb > upperb)
            {
                std::cout << "ASSERTION FAILURE lowerb > upperb\n";
                lowerb = upperb;
            }
            if (single->isSym())
                for (std::size_t i = 0; i < unconverged; ++i)
                    residLast[i] = std::min(residLast[i], resid[i]);

            if (config.DoOptimization() && iteration != 0)
                deg = calc_degrees(single, unconverged, nex, upperb, lowerb, R(tol), ritzv, resid, degrees, locked);

            filter(single, unconverged, deg, degrees, lambda, lowerb, upperb);

            // condition estimate of the filtered block -> CholQR variant
            const R cc = (upperb + lowerb) / 2;
            const R ee = (upperb - lowerb) / 2;
            const R t_1 = (single->GetRitzv()[0] - cc) / ee;
            const R t_k = (ritzv[0] - cc) / ee;
            const R rho_1 = cheb_rho(t_1, false);
            const R rho_k = cheb_rho(t_k, false);
            cond = R(std::pow(rho_k, degrees[0]) *
                     std::pow(rho_1, (*std::max_element(degrees, degrees + nevex - locked) - degrees[0])));

            single->QR(locked, cond);
            single->RR(ritzv, unconverged);
            single->Resd(ritzv, resid, locked);

            new_converged = locking(single, unconverged - nex, R(tol), ritzv, resid, residLast, &early_locked, locked);
            single->Lock(new_converged);

            locked += new_converged;
            unconverged -= new_converged;
            resid += new_converged;
            residLast += new_converged;
            ritzv += new_converged;
            degrees += new_converged;
            iteration++;
        }

        // final ordering of the first nev pairs by eigenvalue, realised as
        // swaps along the permutation cycles
        std::vector<std::size_t> perm(nev);
        std::iota(perm.begin(), perm.end(), 0);
        std::sort(perm.begin(), perm.end(), [&](std::size_t i, std::size_t j) { return ritzv_[i] < ritzv_[j]; });
        std::vector<bool> visited(nev, false);
        for (std::size_t i = 0; i < nev; ++i)
        {
            if (visited[i] || perm[i] == i)
                continue;
            std::size_t current = i;
            const R temp_ritz = ritzv_[i];
            const R temp_resid = resid_[i];
            std::vector<std::size_t> cyc;
            while (!visited[current])
            {
                visited[current] = true;
                cyc.push_back(current);
                current = perm[current];
            }
            for (std::size_t k = 0; k + 1 < cyc.size(); ++k)
            {
                ritzv_[cyc[k]] = ritzv_[cyc[k + 1]];
                resid_[cyc[k]] = resid_[cyc[k + 1]];
            }
            ritzv_[cyc.back()] = temp_ritz;
            resid_[cyc.back()] = temp_resid;
            for (std::size_t k = 0; k + 1 < cyc.size(); ++k)
                single->Swap(cyc[k], cyc[k + 1]);
        }
        single->set_early_locked_residuals(early_locked);
        single->End();
    }
    // ---------------------------------------------------------------------
    // Pseudo-Hermitian (BSE) driver.  The subspace holds 2 (nev+nex) columns:
    // [locked+ | active | K-conjugates of active | K-conjugates of locked].
    // Expressions keep the reference's operand types (double literals mixed with
    // Base<T>) so that FP32 problems take the same decisions too.
    // ---------------------------------------------------------------------

    // Chebyshev ratio with a complex square root (valid inside the damped interval too)
    static R cheb_rho_c(R t)
    {
        const std::complex<R> z(t * t - 1, 0);
        const std::complex<R> q = std::sqrt(z);
        return std::max(std::abs(std::complex<R>(t, 0) - q), std::abs(std::complex<R>(t, 0) + q));
    }

    // per-vector degree multipliers from local spectral density and relative residual size
    static void detect_eigenvalue_clusters(const R* ritzv, const R* resid, R tol, std::size_t unconverged,
                                           std::size_t nex, R upperb, R lowerb, std::vector<R>& factors)
    {
        const std::size_t na = unconverged - nex;
        factors.assign(na, 1.0);
        const R near = std::abs(upperb - lowerb) * 1e-6;
        const R f_lo = 0.5, f_hi = 3.0;

        std::vector<R> weight(na);
        R mean_res = 0.0;
        for (std::size_t i = 0; i < na; ++i)
            mean_res += resid[i];
        mean_res /= na;
        for (std::size_t i = 0; i < na; ++i)
        {
            const R rel = resid[i] / (mean_res + 1e-14);
            weight[i] = 1.0 + std::log(1.0 + rel);
            weight[i] = std::min(weight[i], R(2.5));
        }
        for (std::size_t i = 0; i < na; ++i)
        {
            R spatial = 1.0;
            R density = 0.0;
            std::size_t neighbours = 0;
            for (std::size_t j = 0; j < na; ++j)
            {
                if (i == j)
                    continue;
                const R dist = std::abs(ritzv[i] - ritzv[j]);
                if (dist < near)
                {
                    density += weight[j] / (dist + 1e-14);
                    neighbours++;
                }
            }
            if (neighbours > 0)
                spatial = 1.0 + std::log(1.0 + density * 0.1);
            R f = spatial * weight[i];
            if (neighbours > 2 && resid[i] > 2.0 * mean_res)
                f *= 1.2;
            if (resid[i] > 10.0 * tol)
                f *= 1.15;
            factors[i] = std::min(f_hi, std::max(f_lo, f));
        }
        const std::vector<R> raw = factors;
        for (std::size_t i = 1; i + 1 < na; ++i)
            factors[i] = 0.25 * raw[i - 1] + 0.5 * raw[i] + 0.25 * raw[i + 1];
        for (std::size_t i = 0; i < na; ++i)
            factors[i] = std::min(f_hi, std::max(f_lo, factors[i]));
    }

    static std::size_t calc_degrees_pseudo_H2(ChaseBase<T>* single, std::size_t unconverged, std::size_t nex, R upperb,
                                              R lowerb, R tol, R* ritzv, R* resid, R* residLast, std::size_t* degrees,
                                              std::size_t locked)
    {
        ChaseConfig<T> conf = single->GetConfig();
        const std::size_t deg_extra = conf.GetDegExtra();
        const std::size_t deg_cap = conf.GetMaxDeg();
        const bool aware = conf.UseClusterAwareDegrees();

        std::vector<R> factors;
        if (aware)
            detect_eigenvalue_clusters(ritzv, resid, tol, unconverged, nex, upperb, lowerb, factors);

        const R c = (upperb + lowerb) / 2;
        const R e = (upperb - lowerb) / 2;
        if (e <= R(0))
        {
            for (std::size_t i = 0; i < unconverged; ++i)
                degrees[i] = deg_cap + (deg_cap % 2);
            return deg_cap + (deg_cap % 2);
        }
        for (std::size_t i = 0; i < unconverged; ++i)
        {
            const R mu = ritzv[i] * ritzv[i]; // eigenvalue of H^2
            const R rho = cheb_rho_c((mu - c) / e);
            std::size_t deg;
            if (!std::isfinite(rho) || rho <= R(1))
                deg = deg_cap;
            else
            {
                const R steps = std::log(resid[i] / tol) / std::log(rho);
                if (!std::isfinite(steps))
                    deg = deg_cap;
                else
                {
                    deg = static_cast<std::size_t>(std::ceil(std::abs(static_cast<double>(steps))));
                    if (aware)
                    {
                        if (i < factors.size())
                            deg = static_cast<std::size_t>(deg * factors[i]);
                        else
                            deg = static_cast<std::size_t>(deg * 1.0);
                        // residual stuck within 10 tol: push harder
                        const R near_tol = tol * 10.0;
                        if (resid[i] <= near_tol)
                        {
                            const R change = std::abs(resid[i] - residLast[i]);
                            const R rel_change = change / (resid[i] + 1e-14);
                            const R stuck = 0.1;
                            if (rel_change < stuck)
                                deg += 6;
                        }
                        // small |lambda| relative to the damped interval
                        if (std::abs(ritzv[i]) < std::abs(upperb - lowerb) * 0.1)
                            deg += 2;
                    }
                    deg = std::min(deg + deg_extra, deg_cap);
                }
            }
            if (std::is_same<R, float>::value)
                deg = std::max(deg, std::size_t(8));
            degrees[i] = deg + (deg % 2);
        }
        for (std::size_t j = 0; j + 1 < unconverged; ++j)
            for (std::size_t k = j; k < unconverged; ++k)
                if (degrees[k] < degrees[j])
                {
                    std::swap(degrees[k], degrees[j]);
                    std::swap(ritzv[k], ritzv[j]);
                    std::swap(resid[k], resid[j]);
                    single->Swap(k + locked, j + locked);
                }
        return *std::max_element(degrees, degrees + unconverged);
    }

    // Chebyshev filter in H^2 on the first `unconverged` active columns; columns retire from the left.
    static std::size_t filter_H2(ChaseBase<T>* single, std::size_t unconverged, const std::size_t* degrees, R lambda_1,
                                 R lower, R upper)
    {
        if (lower >= upper)
            std::swap(lower, upper);
        const R c = (upper + lower) / 2;
        const R e = (upper - lower) / 2;
        const R sigma_1 = e / (lambda_1 - c);
        R sigma = sigma_1;
        std::size_t deg_max = 0;
        for (std::size_t i = 0; i < unconverged; ++i)
            deg_max = std::max(deg_max, degrees[i]);

        std::size_t Av = 0;
        const T a1 = T(sigma_1 / e);
        single->HEMM_H2(unconverged, a1, T(0), T(-a1 * c), 0, 0);
        Av += 2 * unconverged;
        std::size_t s = 0;
        for (std::size_t t = 2; t <= deg_max; ++t)
        {
            if (s >= unconverged)
                break;
            const R tau = 1.0 / (2.0 / sigma_1 - sigma);
            const T alpha = T(2.0 * tau / e);
            const R beta = sigma * tau;
            single->HEMM_H2(unconverged, alpha, T(-beta), T(-alpha * c), s, 0);
            Av += 2 * (unconverged - s);
            sigma = tau;
            while (s < unconverged && degrees[s] <= t)
                ++s;
        }
        return Av;
    }

    // Locks converged positive pairs (symmetric partner handled by ApplyKconjugate afterwards).
    static std::size_t locking_pseudo(ChaseBase<T>* single, std::size_t unconverged, std::size_t nex, R tol,
                                      const std::size_t* index, R* Lritzv, R* resid, R* residLast,
                                      std::vector<R>* early, std::size_t locked, std::size_t iteration)
    {
        std::vector<std::size_t> open_idx;
        const std::vector<R> resid_in(resid, resid + 2 * unconverged);
        std::size_t converged = 0;
        for (std::size_t k = 0; k < unconverged - nex; ++k)
        {
            const std::size_t j = index[k];
            const bool early_lock =
                resid[j] > tol && resid[j] >= residLast[k] && resid[j] <= 1000.0 * tol && iteration >= 4;
            if (resid[j] <= tol || early_lock)
            {
                if (early_lock)
                    early->push_back(resid[j]);
                if (j != converged)
                {
                    std::swap(resid[j], resid[converged]);
                    std::swap(Lritzv[j], Lritzv[converged]);
                    single->Swap(j + locked, converged + locked);
                }
                converged++;
            }
            else
                open_idx.push_back(j);
        }
        for (std::size_t k = unconverged - nex; k < unconverged; ++k)
            open_idx.push_back(index[k]);
        for (std::size_t i = converged; i < unconverged; ++i)
            residLast[i] = resid_in[open_idx[i - converged]];
        return converged;
    }

    // Bounds of H^2 from Lanczos in the S-inner product + DoS start vectors.
    static std::size_t lanczos_for_H2(ChaseBase<T>* single, int N, int numvec, int m, int nevex, R* upperb, bool mode,
                                      R* ritzv_)
    {
        assert(m >= 1);
        if (!mode)
        {
            single->Lanczos(m, upperb);
            if (upperb)
                *upperb = (*upperb) * (*upperb);
            return 0;
        }
        const std::size_t nt = (std::size_t)numvec * m;
        std::vector<R> Theta(nt, R(0)), Tau(nt, R(0)), ritzV((std::size_t)m * m, R(0));
        single->Lanczos(m, numvec, upperb, Theta.data(), Tau.data(), ritzV.data());

        std::vector<double> ThetaSorted(Theta.begin(), Theta.end());
        std::sort(ThetaSorted.begin(), ThetaSorted.end(), std::less<double>());
        const double sigma = 0.25;
        const double thresh = 2 * sigma * sigma / 10;
        const auto G = [&](double x) -> double { return 0.5 * (1 + std::erf(x / std::sqrt(2 * sigma * sigma))); };

        R max_abs = 0;
        R min_abs = std::abs(Theta[0]);
        std::size_t i_min = 0;
        for (std::size_t i = 0; i < nt; ++i)
        {
            const R a = std::abs(Theta[i]);
            if (a > max_abs)
                max_abs = a;
            if (a < min_abs)
            {
                min_abs = a;
                i_min = i;
            }
        }
        const R mu_1 = Theta[i_min] * Theta[i_min];
        if (upperb)
            *upperb = max_abs * max_abs;

        ChaseConfig<T>& config = single->GetConfig();
        double search = (static_cast<double>(N) / 2 - static_cast<double>(config.GetNev()) -
                         static_cast<double>(config.GetNex()) - 1) /
                        static_cast<double>(N);
        search = std::min(1.0, std::max(0.0, search));

        R lambda_q = static_cast<R>(ThetaSorted[nt - 1]);
        double prev = 0;
        for (std::size_t i = 0; i < nt; ++i)
        {
            double curr = 0;
            for (std::size_t j = 0; j < nt; ++j)
            {
                if (ThetaSorted[i] < (Theta[j] - thresh))
                    curr += 0;
                else if (ThetaSorted[i] > (Theta[j] + thresh))
                    curr += Tau[j] * 1;
                else
                    curr += Tau[j] * G(ThetaSorted[i] - Theta[j]);
            }
            curr /= numvec;
            if (curr > search)
            {
                if (std::abs(curr - search) < std::abs(prev - search))
                    lambda_q = static_cast<R>(ThetaSorted[i]);
                else
                    lambda_q = static_cast<R>(i > 0 ? ThetaSorted[i - 1] : ThetaSorted[i]);
                break;
            }
            prev = curr;
            lambda_q = static_cast<R>(ThetaSorted[i]);
        }
        const R mu_q = lambda_q * lambda_q;

        const R* last = Theta.data() + (std::size_t)(numvec - 1) * m;
        int idx = 0;
        for (int i = 0; i < m; ++i)
        {
            if (last[i] > lambda_q)
            {
                idx = i - 1;
                break;
            }
            idx = i + 1;
        }
        idx = std::max(idx, 0);
        if (idx > 0)
        {
            std::vector<T> ritzVc((std::size_t)m * m);
            for (std::size_t i = 0; i < (std::size_t)m * m; ++i)
                ritzVc[i] = T(ritzV[i]);
            single->LanczosDos(idx, m, ritzVc.data());
        }
        if (ritzv_)
        {
            for (int i = 0; i < idx; ++i)
                ritzv_[i] = last[i] * last[i];
            for (int i = idx; i < nevex - 1; ++i)
                ritzv_[i] = mu_1;
            ritzv_[nevex - 1] = mu_q;
        }
        for (int i = 1; i < idx; ++i)
        {
            const int j = i * (nevex / idx);
            single->Swap(i, j);
            if (ritzv_)
                std::swap(ritzv_[i], ritzv_[j]);
        }
        return static_cast<std::size_t>(idx);
    }

    static void solve_pseudo(ChaseBase<T>* single)
    {
        ChaseConfig<T>& config = single->GetConfig();
        single->Start();

        const std::size_t N = config.GetN();
        const std::size_t nev = config.GetNev();
        const std::size_t nex = config.GetNex();
        const std::size_t nevex = nev + nex;
        std::size_t unconverged = nevex;

        R* ritzv_ = single->GetRitzv();
        R* ritzv = ritzv_;
        std::size_t deg = config.GetDeg();
        deg += deg % 2;
        deg = std::min(deg, config.GetMaxDeg());
        std::vector<std::size_t> degrees_(2 * nevex);
        for (std::size_t i = 0; i < unconverged; ++i)
            degrees_[i] = deg;
        std::size_t* degrees = degrees_.data();

        const double tol = config.GetTol();
        R* resid_ = single->GetResid();
        std::vector<R> residLast_(2 * nevex);
        for (std::size_t i = 0; i < 2 * unconverged; ++i)
        {
            resid_[i] = std::numeric_limits<R>::max();
            residLast_[i] = std::numeric_limits<R>::max();
        }
        R* resid = resid_;
        R* residLast = residLast_.data();

        const bool random = !config.UseApprox();
        single->initVecs(random);
        if (random)
            single->QR(0, 1.0);

        R upperb = R(0);
        std::size_t lanczos_iter = std::min(nevex, std::min(N / 2, config.GetLanczosIter()));
        if (2u * (lanczos_iter / 2u) < lanczos_iter)
        {
            config.SetLanczosIter(lanczos_iter - 1);
            lanczos_iter = config.GetLanczosIter();
        }
        lanczos_for_H2(single, static_cast<int>(N), static_cast<int>(config.GetNumLanczos()),
                       static_cast<int>(lanczos_iter), static_cast<int>(nevex), &upperb, true, ritzv_);

        // interval of H^2: [lambda_1 | lower ... b_sup]; only `lower` moves afterwards
        const R lambda_1 = *std::min_element(ritzv_, ritzv_ + nevex - 1);
        R lower = ritzv_[nevex - 1];
        if (upperb > 0)
            upperb = upperb * config.GetUpperbScaleRate();
        else
            upperb = upperb / config.GetUpperbScaleRate();
        const R b_sup = upperb;
        R next_lower = lower;
        lower = lower * config.GetDecayingRate();

        std::vector<std::size_t> index(2 * unconverged);
        std::vector<std::size_t> order(unconverged);
        std::vector<R> early_locked;
        R cond;
        std::size_t locked = 0, new_converged = 0, iteration = 0;

        while (locked < nev && unconverged > 0 && iteration < config.GetMaxIter())
        {
            if (iteration > 0)
            {
                next_lower = next_lower * next_lower;
                if (next_lower < lower && next_lower > lambda_1)
                    lower = next_lower;
            }
            if (config.GetLogLevel() >= LogLevel::Debug)
            {
                std::ostringstream oss;
                oss << std::scientific << "iteration: " << iteration << "\t" << lambda_1 << "\t" << lower << "\t"
                    << b_sup << "\t" << unconverged << "\n";
                single->Output(LogLevel::Debug, oss.str(), "algorithm");
            }
            if (config.DoOptimization() && iteration != 0)
                deg = calc_degrees_pseudo_H2(single, unconverged, nex, b_sup, lower, tol, ritzv, resid, residLast,
                                             degrees, locked);

            single->FilterPhaseStart();
            filter_H2(single, unconverged, degrees, lambda_1, lower, b_sup);
            single->FilterPhaseEnd();
            single->ApplyKconjugate(unconverged);

            const R cc = (b_sup + lower) / 2;
            R ee = (b_sup - lower) / 2;
            if (ee <= R(0))
                ee = std::abs(lower - b_sup) / 2;
            const R t_1 = (lambda_1 - cc) / ee;
            const R t_k = (iteration > 0) ? (ritzv[0] * ritzv[0] - cc) / ee : t_1;
            const R rho_1 = cheb_rho_c(t_1);
            const R rho_k = cheb_rho_c(t_k);
            const std::size_t deg_max = *std::max_element(degrees, degrees + unconverged);
            cond = std::pow(rho_k, static_cast<R>(degrees[0])) * std::pow(rho_1, static_cast<R>(deg_max - degrees[0]));

            single->QR(locked, cond);
            single->RR(ritzv, unconverged);
            single->Resd(ritzv, resid, locked);

            std::iota(index.begin(), index.begin() + 2 * unconverged, 0);
            std::iota(order.begin(), order.begin() + unconverged, 0);
            std::sort(order.begin(), order.begin() + std::size_t(1.0 * unconverged),
                      [&](int a, int b) { return (ritzv[a] < ritzv[b]); });
            // reference: order[size_t(unconverged * 0.95) - 1] (algorithm.inc:2130) reads order[SIZE_MAX] when a single
            // pair is left; clamped here (identical for unconverged >= 2)
            next_lower = ritzv[order[std::max<std::size_t>(std::size_t(unconverged * 0.95), 1) - 1]] *
                         config.GetDecayingRate();

            new_converged = locking_pseudo(single, unconverged, nex, tol, index.data(), ritzv, resid, residLast,
                                           &early_locked, locked, iteration);
            if (new_converged > 0)
                single->ApplyKconjugate(new_converged);
            single->Lock(new_converged);

            locked += new_converged;
            unconverged -= new_converged;
            resid += new_converged;
            residLast += new_converged;
            ritzv += new_converged;
            degrees += new_converged;
            ++iteration;
        }

        // positive Ritz values first (ascending), then the rest (ascending)
        std::size_t n_reorder = locked + unconverged;
        if (n_reorder == 0)
            n_reorder = 1;
        std::vector<std::size_t> perm(n_reorder);
        std::iota(perm.begin(), perm.end(), 0);
        std::sort(perm.begin(), perm.end(), [&](std::size_t i, std::size_t j) {
            const bool ip = ritzv_[i] > R(0), jp = ritzv_[j] > R(0);
            if (ip != jp)
                return ip;
            return ritzv_[i] < ritzv_[j];
        });
        std::vector<bool> visited(n_reorder, false);
        for (std::size_t i = 0; i < n_reorder; ++i)
        {
            if (visited[i] || perm[i] == i)
                continue;
            std::vector<std::size_t> cyc;
            for (std::size_t cur = i; !visited[cur]; cur = perm[cur])
            {
                visited[cur] = true;
                cyc.push_back(cur);
            }
            const R ritz0 = ritzv_[i], resid0 = resid_[i];
            for (std::size_t k = 0; k + 1 < cyc.size(); ++k)
            {
                ritzv_[cyc[k]] = ritzv_[cyc[k + 1]];
                resid_[cyc[k]] = resid_[cyc[k + 1]];
                single->Swap(cyc[k], cyc[k + 1]);
            }
            ritzv_[cyc.back()] = ritz0;
            resid_[cyc.back()] = resid0;
        }
        single->set_early_locked_residuals(early_locked);
        single->End();
    }
};

// Entry point, same name and meaning as the reference's chase::Solve
// (algorithm/algorithm.hpp:345-349).
template <class T>
void Solve(ChaseBase<T>* single)
{
    Algorithm<T>::solve(single);
}

// Pseudo-Hermitian (BSE) entry point (algorithm/algorithm.hpp:359-363): the backend must hold 2 (nev+nex)
// columns and GetRitzvBlockSize() == 2 (nev+nex).
template <class T>
void Solve_pseudo(ChaseBase<T>* single)
{
    Algorithm<T>::solve_pseudo(single);
}

} // namespace chase
