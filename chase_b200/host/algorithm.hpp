// chase_b200 host layer — Chebyshev-filtered subspace iteration driver.
//
// Written from scratch; decision logic (bounds, degree schedule, locking,
// condition estimate, final ordering) follows the reference's
// chase::Algorithm<T> step by step so that iteration counts and HEMM schedules
// are identical on identical inputs:
//   solve         /root/reference/algorithm/algorithm.inc:1376-1788
//   filter        :942-1009          calc_degrees :136-193
//   locking       :519-578           lanczos/DoS  :1067-1214
// Only the Hermitian path is implemented here (pseudo-Hermitian is a later row
// of the scope table).
#pragma once
#include "interface.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
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
};

// Entry point, same name and meaning as the reference's chase::Solve
// (algorithm/algorithm.hpp:345-349).
template <class T>
void Solve(ChaseBase<T>* single)
{
    Algorithm<T>::solve(single);
}

} // namespace chase
