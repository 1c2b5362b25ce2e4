// chase_b200 — dense Hermitian eigensolver for the projected Rayleigh-Ritz
// matrix (replaces cusolverDn?heevd / ?syevd called by the reference at
// /root/reference/linalg/internal/cuda/rayleighRitz.hpp:169-214) and for the
// Lanczos tridiagonals (replaces host LAPACK ?stemr,
// /root/reference/linalg/internal/cuda/lanczos.hpp:270-299).
//
// Method: two-sided cyclic Jacobi with the round-robin (tournament) ordering.
// One round holds n/2 disjoint index pairs (p_a, q_a).  Because the pairs are
// disjoint, G <- J^H G J decomposes into (n/2)^2 independent 2x2 block updates
//        G[{p_a,q_a},{p_b,q_b}] <- J_a^H G[{p_a,q_a},{p_b,q_b}] J_b
// so a round is ONE fully parallel in-place kernel for G (plus Z <- Z J).
// Hermitian 2x2 [[a,b],[conj b,d]], b=|b|u:  J = [[c, s],[-s conj(u), c conj(u)]]
// with t = sgn(th)/(|th|+sqrt(th^2+1)), th=(d-a)/(2|b|), c=1/sqrt(1+t^2), s=t c.
#pragma once
#include "common.cuh"

namespace cb2
{

struct JRot
{
    double c, s;
    double ure, uim; // unit phase u
    double tapq;     // t * |a_pq|: the diagonal moves by -/+ this amount
    int active;
    int pad;
};

__host__ __device__ inline void rr_pair(int r, int a, int np, int& p, int& q)
{
    // np even; round r in [0, np-1); pair a in [0, np/2)
    const int m = np - 1;
    if (a == 0)
    {
        p = m;
        q = r;
    }
    else
    {
        p = (r + a) % m;
        q = (r - a + m) % m;
    }
}

__device__ __forceinline__ cxd phase_of(cxd b, double ab) { return cxd{b.re / ab, b.im / ab}; }
__device__ __forceinline__ double phase_of(double b, double) { return b >= 0 ? 1.0 : -1.0; }
__device__ __forceinline__ void store_u(JRot& R, cxd u)
{
    R.ure = u.re;
    R.uim = u.im;
}
__device__ __forceinline__ void store_u(JRot& R, double u)
{
    R.ure = u;
    R.uim = 0.0;
}
template <class C>
__device__ __forceinline__ C load_u(const JRot& R);
template <>
__device__ __forceinline__ double load_u<double>(const JRot& R)
{
    return R.ure;
}
template <>
__device__ __forceinline__ cxd load_u<cxd>(const JRot& R)
{
    return cxd{R.ure, R.uim};
}

__device__ __forceinline__ double jabs(double a) { return fabs(a); }
__device__ __forceinline__ double jabs(cxd a) { return sqrt(a.re * a.re + a.im * a.im); }

template <class C>
__device__ __forceinline__ JRot make_rot(double app, double aqq, C apq, double thresh)
{
    JRot R;
    R.c = 1.0;
    R.s = 0.0;
    R.ure = 1.0;
    R.uim = 0.0;
    R.tapq = 0.0;
    R.active = 0;
    R.pad = 0;
    const double ab = jabs(apq);
    if (!(ab > thresh))
        return R;
    // t = sign(zeta) 2|a_pq| / (|zeta| + sqrt(zeta^2 + 4|a_pq|^2)), zeta = a_qq - a_pp: the smaller root of
    // t^2 + 2 theta t - 1 = 0 with theta = zeta / (2|a_pq|), written with one square root, one division and one
    // reciprocal square root (the rotation sits on the critical path of every Jacobi step: 8 threads compute while
    // the CTA waits)
    const double zeta = aqq - app;
    const double t = (zeta >= 0 ? 2.0 : -2.0) * ab / (fabs(zeta) + sqrt(zeta * zeta + 4.0 * ab * ab));
    R.c = rsqrt(t * t + 1.0);
    R.s = t * R.c;
    R.tapq = t * ab;
    store_u(R, phase_of(apq, ab));
    R.active = 1;
    return R;
}

// B <- Ja^H B Jb on a 2x2 block b = [[b00,b01],[b10,b11]] (row index = pair a)
template <class C>
__device__ __forceinline__ void rot_block(C& b00, C& b01, C& b10, C& b11, const JRot& Ra, const JRot& Rb)
{
    // columns: B J_b
    {
        const C ub = cconj(load_u<C>(Rb));
        const C t00 = csub(cmul(Rb.c, b00), cmul(Rb.s, cmul(ub, b01)));
        const C t01 = cadd(cmul(Rb.s, b00), cmul(Rb.c, cmul(ub, b01)));
        const C t10 = csub(cmul(Rb.c, b10), cmul(Rb.s, cmul(ub, b11)));
        const C t11 = cadd(cmul(Rb.s, b10), cmul(Rb.c, cmul(ub, b11)));
        b00 = t00;
        b01 = t01;
        b10 = t10;
        b11 = t11;
    }
    // rows: J_a^H B ;  J^H = [[c, -s u],[s, c u]]
    {
        const C ua = load_u<C>(Ra);
        const C t00 = csub(cmul(Ra.c, b00), cmul(Ra.s, cmul(ua, b10)));
        const C t10 = cadd(cmul(Ra.s, b00), cmul(Ra.c, cmul(ua, b10)));
        const C t01 = csub(cmul(Ra.c, b01), cmul(Ra.s, cmul(ua, b11)));
        const C t11 = cadd(cmul(Ra.s, b01), cmul(Ra.c, cmul(ua, b11)));
        b00 = t00;
        b10 = t10;
        b01 = t01;
        b11 = t11;
    }
}

__device__ __forceinline__ double real_only(double a) { return a; }
__device__ __forceinline__ cxd real_only(cxd a) { return cxd{a.re, 0.0}; }

// ---------------------------------------------------------------------------
// large-n path: kernels launched once per round
// ---------------------------------------------------------------------------
template <class T>
__global__ void jacobi_init_kernel(int n, const T* G, long long ldg, typename Traits<T>::comp* Gw,
                                   typename Traits<T>::comp* Zw, double* fro2)
{
    using C = typename Traits<T>::comp;
    __shared__ double sh[32];
    double acc = 0.0;
    const long long total = (long long)n * n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
    {
        const int i = (int)(idx % n), j = (int)(idx / n);
        C v;
        if (i > j)
            v = widen(G[i + j * ldg]); // lower triangle is the reference ('L' in heevd)
        else if (i < j)
            v = cconj(widen(G[j + i * ldg]));
        else
            v = real_only(widen(G[i + j * ldg]));
        Gw[idx] = v;
        Zw[idx] = from_real<C>(i == j ? 1.0 : 0.0);
        acc += cabs2(v);
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0)
        atomicAdd(fro2, acc);
}

template <class C>
__global__ void jacobi_rot_kernel(int n, int np, int r, const C* Gw, JRot* rots, const double* fro2, int* nrot)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= np / 2)
        return;
    int p, q;
    rr_pair(r, a, np, p, q);
    JRot R;
    R.c = 1.0;
    R.s = 0.0;
    R.ure = 1.0;
    R.uim = 0.0;
    R.tapq = 0.0;
    R.active = 0;
    R.pad = 0;
    if (p < n && q < n)
    {
        const double thresh = 8.0 * 2.220446049250313e-16 * sqrt(*fro2 / (double)n);
        R = make_rot<C>(creal(Gw[p + (long long)p * n]), creal(Gw[q + (long long)q * n]), Gw[p + (long long)q * n],
                        thresh);
        if (R.active)
            atomicAdd(nrot, 1);
    }
    rots[a] = R;
}

// grid: (ceil(max(n, np/2)/128), np/2, 2); z=0 updates G blocks, z=1 updates Z
template <class C>
__global__ void jacobi_apply_kernel(int n, int np, int r, C* Gw, C* Zw, const JRot* rots)
{
    const int b = blockIdx.y;
    const JRot Rb = rots[b];
    int pb, qb;
    rr_pair(r, b, np, pb, qb);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.z == 0)
    {
        if (t >= np / 2)
            return;
        const JRot Ra = rots[t];
        if (!Ra.active && !Rb.active)
            return;
        int pa, qa;
        rr_pair(r, t, np, pa, qa);
        const bool va_p = pa < n, va_q = qa < n, vb_p = pb < n, vb_q = qb < n;
        C b00 = czero<C>(), b01 = czero<C>(), b10 = czero<C>(), b11 = czero<C>();
        if (va_p && vb_p)
            b00 = Gw[pa + (long long)pb * n];
        if (va_p && vb_q)
            b01 = Gw[pa + (long long)qb * n];
        if (va_q && vb_p)
            b10 = Gw[qa + (long long)pb * n];
        if (va_q && vb_q)
            b11 = Gw[qa + (long long)qb * n];
        const double app = creal(b00), aqq = creal(b11);
        rot_block<C>(b00, b01, b10, b11, Ra, Rb);
        if (t == b)
        {
            // pivot block: annihilated element set exactly; the diagonal is updated with the
            // classical small-correction formulas (a_pp - t|a_pq|, a_qq + t|a_pq|) so that its
            // rounding error scales with the correction, not with |a_pp|
            b01 = czero<C>();
            b10 = czero<C>();
            b00 = from_real<C>(app - Ra.tapq);
            b11 = from_real<C>(aqq + Ra.tapq);
        }
        if (va_p && vb_p)
            Gw[pa + (long long)pb * n] = b00;
        if (va_p && vb_q)
            Gw[pa + (long long)qb * n] = b01;
        if (va_q && vb_p)
            Gw[qa + (long long)pb * n] = b10;
        if (va_q && vb_q)
            Gw[qa + (long long)qb * n] = b11;
    }
    else
    {
        if (t >= n || !Rb.active)
            return;
        // pb, qb < n whenever the rotation is active
        const C ub = cconj(load_u<C>(Rb));
        const C zp = Zw[t + (long long)pb * n], zq = Zw[t + (long long)qb * n];
        Zw[t + (long long)pb * n] = csub(cmul(Rb.c, zp), cmul(Rb.s, cmul(ub, zq)));
        Zw[t + (long long)qb * n] = cadd(cmul(Rb.s, zp), cmul(Rb.c, cmul(ub, zq)));
    }
}

template <class C>
__global__ void jacobi_diag_kernel(int n, const C* Gw, double* w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        w[i] = creal(Gw[i + (long long)i * n]);
}

// Z[:, j] <- Zw[:, perm[j]]  (narrowed to the storage type)
template <class T>
__global__ void jacobi_gather_kernel(int n, const typename Traits<T>::comp* Zw, long long ldzw, const int* perm, T* Z,
                                     long long ldz)
{
    const int j = blockIdx.y;
    const int src = perm[j];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        Z[i + j * ldz] = narrow<T>(Zw[i + (long long)src * ldzw]);
}

// ---------------------------------------------------------------------------
// small-n path (n <= 48): one CTA per matrix, everything in shared memory.
// Used for the Lanczos tridiagonals: batch of `gridDim.x` real symmetric
// tridiagonal matrices given by (d, e); outputs ascending eigenvalues and the
// full eigenvector matrices (column-major, n x n).
// ---------------------------------------------------------------------------
constexpr int JSMALL_MAX = 48;

__global__ void __launch_bounds__(256) jacobi_small_tridiag_kernel(int n, const double* d, const double* e, int ldde,
                                                                   double* w, double* Z, int* sweeps_out)
{
    __shared__ double G[JSMALL_MAX * (JSMALL_MAX + 1)];
    __shared__ double V[JSMALL_MAX * (JSMALL_MAX + 1)];
    __shared__ JRot rots[JSMALL_MAX / 2];
    __shared__ int nrot;
    __shared__ double sh[32];
    const int LD = JSMALL_MAX + 1;
    const int tid = threadIdx.x;
    const double* dd = d + (long long)blockIdx.x * ldde;
    const double* ee = e + (long long)blockIdx.x * ldde;
    double acc = 0.0;
    for (int idx = tid; idx < n * n; idx += blockDim.x)
    {
        const int i = idx % n, j = idx / n;
        double v = 0.0;
        if (i == j)
            v = dd[i];
        else if (i == j + 1)
            v = ee[j];
        else if (j == i + 1)
            v = ee[i];
        G[i + j * LD] = v;
        V[i + j * LD] = (i == j) ? 1.0 : 0.0;
        acc += v * v;
    }
    const double fro2 = block_sum(acc, sh);
    const double thresh = 8.0 * 2.220446049250313e-16 * sqrt(fro2 / (double)n);
    const int np = (n + 1) & ~1, h = np / 2;
    int sweep = 0;
    for (; sweep < 60; ++sweep)
    {
        if (tid == 0)
            nrot = 0;
        __syncthreads();
        for (int r = 0; r < np - 1; ++r)
        {
            if (tid < h)
            {
                int p, q;
                rr_pair(r, tid, np, p, q);
                JRot R;
                R.c = 1.0;
                R.s = 0.0;
                R.ure = 1.0;
                R.uim = 0.0;
                R.tapq = 0.0;
                R.active = 0;
                R.pad = 0;
                if (p < n && q < n)
                {
                    R = make_rot<double>(G[p + p * LD], G[q + q * LD], G[p + q * LD], thresh);
                    if (R.active)
                        atomicAdd(&nrot, 1);
                }
                rots[tid] = R;
            }
            __syncthreads();
            for (int idx = tid; idx < h * h; idx += blockDim.x)
            {
                const int a = idx % h, b = idx / h;
                const JRot Ra = rots[a], Rb = rots[b];
                if (!Ra.active && !Rb.active)
                    continue;
                int pa, qa, pb, qb;
                rr_pair(r, a, np, pa, qa);
                rr_pair(r, b, np, pb, qb);
                const bool vap = pa < n, vaq = qa < n, vbp = pb < n, vbq = qb < n;
                double b00 = (vap && vbp) ? G[pa + pb * LD] : 0.0;
                double b01 = (vap && vbq) ? G[pa + qb * LD] : 0.0;
                double b10 = (vaq && vbp) ? G[qa + pb * LD] : 0.0;
                double b11 = (vaq && vbq) ? G[qa + qb * LD] : 0.0;
                const double app = b00, aqq = b11;
                rot_block<double>(b00, b01, b10, b11, Ra, Rb);
                if (a == b)
                {
                    b01 = b10 = 0.0;
                    b00 = app - Ra.tapq;
                    b11 = aqq + Ra.tapq;
                }
                if (vap && vbp)
                    G[pa + pb * LD] = b00;
                if (vap && vbq)
                    G[pa + qb * LD] = b01;
                if (vaq && vbp)
                    G[qa + pb * LD] = b10;
                if (vaq && vbq)
                    G[qa + qb * LD] = b11;
            }
            for (int idx = tid; idx < n * h; idx += blockDim.x)
            {
                const int i = idx % n, b = idx / n;
                const JRot Rb = rots[b];
                if (!Rb.active)
                    continue;
                int pb, qb;
                rr_pair(r, b, np, pb, qb);
                const double zp = V[i + pb * LD], zq = V[i + qb * LD];
                V[i + pb * LD] = Rb.c * zp - Rb.s * Rb.ure * zq;
                V[i + qb * LD] = Rb.s * zp + Rb.c * Rb.ure * zq;
            }
            __syncthreads();
        }
        const int done = (nrot == 0);
        __syncthreads();
        if (done)
            break;
    }
    // ascending order by rank counting (n <= 48; ties broken by index)
    for (int j = tid; j < n; j += blockDim.x)
    {
        const double v = G[j + j * LD];
        int rank = 0;
        for (int l = 0; l < n; ++l)
        {
            const double u = G[l + l * LD];
            rank += (u < v) || (u == v && l < j);
        }
        w[(long long)blockIdx.x * n + rank] = v;
        for (int i = 0; i < n; ++i)
            Z[(long long)blockIdx.x * n * n + i + (long long)rank * n] = V[i + j * LD];
    }
    if (tid == 0 && sweeps_out)
        sweeps_out[blockIdx.x] = sweep;
}

} // namespace cb2
