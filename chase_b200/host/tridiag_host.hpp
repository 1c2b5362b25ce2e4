// chase_b200 host layer — symmetric tridiagonal eigen-decomposition on the host (implicit-shift QL with
// eigenvectors), used by the Lanczos bound estimation when the number of steps exceeds the on-device solver's
// 48 x 48 shared-memory tile.  The reference always solves these on the host (LAPACK ?stemr,
// linalg/internal/cuda/lanczos.hpp:270-299); the matrices are M x M with M <= nev+nex Lanczos steps, O(M^3) flop.
#pragma once
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace chase
{
namespace b200
{

// d[0..n-1] diagonal, e[0..n-2] off-diagonal.  w: eigenvalues ascending; Z: n x n column-major, column k = unit
// eigenvector of w[k].  Returns 0, or the index (1-based) of an eigenvalue that did not converge in 60 sweeps.
inline int tridiag_eig_host(int n, const double* d_in, const double* e_in, double* w, double* Z)
{
    if (n <= 0)
        return 0;
    std::vector<double> d(d_in, d_in + n), e((std::size_t)n, 0.0), z((std::size_t)n * n, 0.0);
    for (int i = 0; i + 1 < n; ++i)
        e[i] = e_in[i];
    for (int i = 0; i < n; ++i)
        z[(std::size_t)i + (std::size_t)i * n] = 1.0;
    for (int l = 0; l < n; ++l)
    {
        int iter = 0, m;
        do
        {
            for (m = l; m + 1 < n; ++m)
            {
                const double dd = std::abs(d[m]) + std::abs(d[m + 1]);
                if (std::abs(e[m]) <= 2.220446049250313e-16 * dd)
                    break;
            }
            if (m != l)
            {
                if (iter++ == 60)
                    return l + 1;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::abs(r) : -std::abs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i)
                {
                    double f = s * e[i];
                    const double b = c * e[i];
                    r = std::hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0)
                    {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    for (int k = 0; k < n; ++k)
                    {
                        double* zk = &z[(std::size_t)k];
                        f = zk[(std::size_t)(i + 1) * n];
                        zk[(std::size_t)(i + 1) * n] = s * zk[(std::size_t)i * n] + c * f;
                        zk[(std::size_t)i * n] = c * zk[(std::size_t)i * n] - s * f;
                    }
                }
                if (r == 0.0 && i >= l)
                    continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    std::vector<int> order((std::size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return d[a] < d[b]; });
    for (int k = 0; k < n; ++k)
    {
        w[k] = d[order[k]];
        for (int i = 0; i < n; ++i)
            Z[(std::size_t)i + (std::size_t)k * n] = z[(std::size_t)i + (std::size_t)order[k] * n];
    }
    return 0;
}

} // namespace b200
} // namespace chase
