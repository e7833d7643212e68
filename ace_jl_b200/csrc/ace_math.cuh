// ace_math.cuh -- per-neighbour device math: distance transform, enveloped orthogonal-polynomial
// recursion, and column-wise associated-Legendre / spherical-harmonic recurrences.
//
// What is computed follows the reference (file:line cited at each function); how it is computed is
// organised for one-thread-per-neighbour register residency: no arrays of harmonics are ever formed,
// the (l, m) values are produced one m-column at a time and consumed immediately.
#pragma once

#include "ace_platform.cuh"

namespace aceb200 {

constexpr int kMaxRad = 32;   // n_rad limit (kernel parameter space)
constexpr int kMaxL = 12;     // highest l limit
constexpr int kMaxP = (kMaxL + 1) * (kMaxL + 2) / 2;
constexpr int kMaxOrdDev = 6;   // highest correlation order the adjoint trees handle

// Radial basis parameters, passed by value in the kernel parameter (constant) bank.
struct RadialParams {
    int N, pl, pr, tkind;
    double tl, tr;
    double tpar[4];
    double A[kMaxRad], B[kMaxRad], C[kMaxRad];
};

// ALP recursion coefficients (sphericalharmonics.jl:146-159), index_p(l, m) = m + l(l+1)/2.
struct AlpParams {
    int L;                  // highest l referenced by the model's one-particle basis
    double A[kMaxP], B[kMaxP];
    double diagc[kMaxL + 2];   // diagc[m] = sqrt(1 + 1/(2m)) (m >= 2), sqrt(1.5) (m = 1): P_m^m from P_{m-1}^{m-1}
};

ACE_HD inline int index_p(int l, int m) { return m + (l * (l + 1)) / 2; }

ACE_HD inline double ipow(double x, int n)
{
    double y = 1.0;
    for (int i = 0; i < n; ++i) y *= x;
    return y;
}

// x^p for the exponents a transform carries: small integers by repeated multiplication (what Julia's
// `^` does for an Int literal), anything else through pow().
ACE_HD inline double powp(double x, double p)
{
    if (p == floor(p) && fabs(p) <= 16.0) {
        double y = ipow(x, (int)fabs(p));
        return p < 0 ? 1.0 / y : y;
    }
    return pow(x, p);
}

// t(r), t'(r): src/transforms/distancetransforms.jl:16-25 (derivative: closed form of what
// ForwardDiff returns, src/transforms/lambdas.jl:28-34).
ACE_HD inline void transform_ed(const RadialParams& rp, double r, double& t, double& dt)
{
    switch (rp.tkind) {
    case 0: t = r; dt = 1.0; break;
    case 1: {  // ((1+r0)/(1+r))^p
        double p = rp.tpar[0], ir = 1.0 / (1.0 + r), s = (1.0 + rp.tpar[1]) * ir;
        if (p == 2.0) { t = s * s; dt = -2.0 * t * ir; }
        else { t = powp(s, p); dt = -p * t * ir; }
        break;
    }
    case 2: {  // exp(-lambda (r/r0 - 1))
        double lam = rp.tpar[0], r0 = rp.tpar[1];
        t = exp(-lam * (r / r0 - 1.0));
        dt = -(lam / r0) * t;
        break;
    }
    default: {  // 1/(1 + a (r/r0)^p)
        double r0 = rp.tpar[0], p = rp.tpar[1], a = rp.tpar[2], x = r / r0;
        double xp1 = powp(x, p - 1.0), d = 1.0 + a * xp1 * x;
        t = 1.0 / d;
        dt = -(a * p * xp1 / r0) * t * t;
        break;
    }
    }
}

// Envelope (t-tl)^pl (t-tr)^pr and its t-derivative, zero outside the cut side(s)
// (src/polynomials/orthpolys.jl:32-53).
ACE_HD inline void envelope_ed(const RadialParams& rp, double t, double& f, double& df)
{
    if ((rp.pl > 0 && t < rp.tl) || (rp.pr > 0 && t > rp.tr)) { f = 0.0; df = 0.0; return; }
    double a = t - rp.tl, b = t - rp.tr;
    double al = ipow(a, rp.pl > 0 ? rp.pl - 1 : 0), br = ipow(b, rp.pr > 0 ? rp.pr - 1 : 0);
    double fa = rp.pl > 0 ? al * a : 1.0, fb = rp.pr > 0 ? br * b : 1.0;
    f = fa * fb;
    df = (rp.pl > 0 ? rp.pl * al * fb : 0.0) + (rp.pr > 0 ? rp.pr * fa * br : 0.0);
}

// R_n(r) = P_n(t(r)), n < N (src/polynomials/orthpolys.jl:253-263).  NMAX is a compile-time bound so
// that R stays in registers.
template <int NMAX>
ACE_HD inline void radial_e(const RadialParams& rp, double r, double (&R)[NMAX])
{
    double t, dt, f, df;
    transform_ed(rp, r, t, dt);
    envelope_ed(rp, t, f, df);
    R[0] = rp.A[0] * f;
    if (NMAX > 1) R[1] = (rp.A[1] * t + rp.B[1]) * R[0];
#pragma unroll
    for (int n = 2; n < NMAX; ++n)
        R[n] = (n < rp.N) ? (rp.A[n] * t + rp.B[n]) * R[n - 1] + rp.C[n] * R[n - 2] : 0.0;
}

// R_n and dR_n/dr = P_n'(t) t'(r) (src/polynomials/orthpolys.jl:287-306, chain rule :242-247).
template <int NMAX>
ACE_HD inline void radial_ed(const RadialParams& rp, double r, double (&R)[NMAX], double (&dR)[NMAX])
{
    double t, dt, f, df;
    transform_ed(rp, r, t, dt);
    envelope_ed(rp, t, f, df);
    R[0] = rp.A[0] * f;
    dR[0] = rp.A[0] * df;
    if (NMAX > 1) {
        double al = rp.A[1] * t + rp.B[1];
        R[1] = al * R[0];
        dR[1] = al * dR[0] + rp.A[1] * R[0];
    }
#pragma unroll
    for (int n = 2; n < NMAX; ++n) {
        if (n < rp.N) {
            double al = rp.A[n] * t + rp.B[n];
            R[n] = al * R[n - 1] + rp.C[n] * R[n - 2];
            dR[n] = al * dR[n - 1] + rp.C[n] * dR[n - 2] + rp.A[n] * R[n - 1];
        } else { R[n] = 0.0; dR[n] = 0.0; }
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n) dR[n] *= dt;
}

// radial_ed with dR_n/dr written to dRs[n * stride] as soon as it is formed (two rolling registers instead of
// an NMAX-register array): k_forces parks the derivatives in shared memory.
template <int NMAX>
ACE_HD inline void radial_ed_park(const RadialParams& rp, double r, double (&R)[NMAX], double* dRs, int stride)
{
    double t, dt, f, df;
    transform_ed(rp, r, t, dt);
    envelope_ed(rp, t, f, df);
    R[0] = rp.A[0] * f;
    double d2 = rp.A[0] * df, d1 = 0.0;      // dP_{n-2}/dt, dP_{n-1}/dt
    dRs[0] = d2 * dt;
    if (NMAX > 1) {
        const double al = rp.A[1] * t + rp.B[1];
        R[1] = al * R[0];
        d1 = al * d2 + rp.A[1] * R[0];
        dRs[stride] = d1 * dt;
    }
#pragma unroll
    for (int n = 2; n < NMAX; ++n) {
        if (n < rp.N) {
            const double al = rp.A[n] * t + rp.B[n];
            R[n] = al * R[n - 1] + rp.C[n] * R[n - 2];
            const double d = al * d1 + rp.C[n] * d2 + rp.A[n] * R[n - 1];
            d2 = d1; d1 = d;
            dRs[n * stride] = d * dt;
        } else { R[n] = 0.0; dRs[n * stride] = 0.0; }
    }
}

// Spherical coordinates of a neighbour (src/polynomials/sphericalharmonics.jl:43-51).  The reference
// goes through atan + sincos; x/rho, y/rho are the same numbers to round-off, with the rho = 0
// convention atan(0, 0) = 0 kept (cos = 1, sin = 0).
struct Spher {
    double r, rinv, cphi, sphi, cth, sth;
};

ACE_HD inline Spher cart2spher(double x, double y, double z)
{
    // two reciprocal square roots instead of two square roots and two divisions (each a ~20-instruction
    // FP64 sequence); r = r2 * rsqrt(r2) is within 2 ulp of sqrt(r2)
    Spher S;
    const double rho2 = x * x + y * y;
    const double r2 = rho2 + z * z;
#ifdef __CUDA_ARCH__
    S.rinv = rsqrt(r2);
#else
    S.rinv = 1.0 / sqrt(r2);
#endif
    S.r = r2 * S.rinv;
    if (rho2 > 0.0) {
#ifdef __CUDA_ARCH__
        const double ir = rsqrt(rho2);
#else
        const double ir = 1.0 / sqrt(rho2);
#endif
        S.cphi = x * ir; S.sphi = y * ir;
        S.sth = (rho2 * ir) * S.rinv;
    } else { S.cphi = 1.0; S.sphi = 0.0; S.sth = 0.0; }
    S.cth = z * S.rinv;
    return S;
}

// Walks the harmonics one m-column at a time, calling f(l, m, Pv, epr, epi) with
//   Pv = P_l^m(cos th)  (src/polynomials/sphericalharmonics.jl:175-194, reorganised by column)
//   ep = exp(i m phi) / sqrt(2)  (:385-393),  so that  Y_l^m = ep * Pv  and  Y_l^{-m} = (-1)^m conj(Y_l^m).
// The coefficient tables hold A_{m+1}^m = sqrt(2m+3), B_{m+1}^m = 0 (:179, :191) so that one recurrence
// serves every l > m and the callback is instantiated once.
// Up to kStaticL the walk is fully unrolled behind warp-uniform `l <= L` exits: (l, m) are then compile-time
// constants in the callback, and the recursion coefficients become constant-bank operands of the FP64
// instructions instead of indexed loads.  Models referencing a higher l take the rolled loop.
constexpr int kStaticL = 6;

// WALK selects the code that is generated: kWalkAny (both, chosen at run time), kWalkStatic (the model is known
// to have L <= kStaticL), kWalkRolled.  The hot kernels are instantiated per walk so that each carries one copy.
constexpr int kWalkAny = 0, kWalkStatic = 1, kWalkRolled = 2;

template <int WALK = kWalkAny, class F>
ACE_HD inline void for_each_lm(const AlpParams& ap, const Spher& S, F&& f)
{
    const int L = ap.L;
    const double P00 = 0.39894228040143268;   // sqrt(0.5/pi) (:175)
    const double is2 = 0.70710678118654752;   // 1/sqrt(2)
    double dg = P00;                          // P_m^m
    double epr = is2, epi = 0.0;
    if (WALK == kWalkStatic || (WALK == kWalkAny && L <= kStaticL)) {
#pragma unroll
        for (int m = 0; m <= kStaticL; ++m) {
            if (m > L) break;
            if (m > 0) {
                dg = -ap.diagc[m] * S.sth * dg;
                const double nr = epr * S.cphi - epi * S.sphi;
                epi = epr * S.sphi + epi * S.cphi;
                epr = nr;
            }
            double p2 = 0.0, p1 = dg;
#pragma unroll
            for (int l = m; l <= kStaticL; ++l) {
                if (l > L) break;
                if (l > m) {
                    const double p = (l == m + 1) ? ap.A[index_p(l, m)] * (S.cth * p1)
                                                  : ap.A[index_p(l, m)] * (S.cth * p1 + ap.B[index_p(l, m)] * p2);
                    p2 = p1; p1 = p;
                }
                f(l, m, p1, epr, epi);
            }
        }
        return;
    }
    if (WALK == kWalkStatic) return;
    for (int m = 0; m <= L; ++m) {
        if (m > 0) {
            dg = -ap.diagc[m] * S.sth * dg;   // :180, :192
            const double nr = epr * S.cphi - epi * S.sphi;
            epi = epr * S.sphi + epi * S.cphi;
            epr = nr;
        }
        double p2 = 0.0, p1 = dg;
        for (int l = m; l <= L; ++l) {
            if (l > m) {
                const int ip = index_p(l, m);
                const double p = ap.A[ip] * (S.cth * p1 + ap.B[ip] * p2);   // :188-189
                p2 = p1; p1 = p;
            }
            f(l, m, p1, epr, epi);
        }
    }
}

// Gradient variant: f(l, m, Pt, dP, epr, epi) with
//   Pt = P_l^m / sin th for m >= 1, P_l^0 for m = 0   (the pole-stable storage of :208-267)
//   dP = d P_l^m / d theta
// from which  Y_l^m = ep Pt sin th (m >= 1),  dY/dphi / sin th = i m ep Pt,  dY/dtheta = ep dP  (:410-443).
template <int WALK = kWalkAny, class F>
ACE_HD inline void for_each_lm_ed(const AlpParams& ap, const Spher& S, F&& f)
{
    const int L = ap.L;
    const double P00 = 0.39894228040143268;
    const double is2 = 0.70710678118654752;
    double dgt = P00, dgd = 0.0;              // diagonal Pt_m^m, dP_m^m (temp1, temp_d of :229-263)
    double epr = is2, epi = 0.0;
    const double s2 = S.sth * S.sth;
    if (WALK == kWalkStatic || (WALK == kWalkAny && L <= kStaticL)) {
#pragma unroll
        for (int m = 0; m <= kStaticL; ++m) {
            if (m > L) break;
            const double sfac = (m == 0) ? S.sth : s2;
            if (m == 1) {
                dgd = -ap.diagc[1] * (S.cth * dgt + S.sth * dgd);
                dgt = -ap.diagc[1] * dgt;
            } else if (m > 1) {
                const double nd = -ap.diagc[m] * (S.cth * dgt * S.sth + S.sth * dgd);
                dgt = -ap.diagc[m] * S.sth * dgt;
                dgd = nd;
            }
            if (m > 0) {
                const double nr = epr * S.cphi - epi * S.sphi;
                epi = epr * S.sphi + epi * S.cphi;
                epr = nr;
            }
            double p2 = 0.0, d2 = 0.0, p1 = dgt, d1 = dgd;
#pragma unroll
            for (int l = m; l <= kStaticL; ++l) {
                if (l > L) break;
                if (l > m) {
                    const int ip = index_p(l, m);
                    double p, d;
                    if (l == m + 1) {     // B_{m+1}^m = 0
                        p = ap.A[ip] * (S.cth * p1);
                        d = ap.A[ip] * (-sfac * p1 + S.cth * d1);
                    } else {
                        p = ap.A[ip] * (S.cth * p1 + ap.B[ip] * p2);
                        d = ap.A[ip] * (-sfac * p1 + S.cth * d1 + ap.B[ip] * d2);
                    }
                    p2 = p1; d2 = d1; p1 = p; d1 = d;
                }
                f(l, m, p1, d1, epr, epi);
            }
        }
        return;
    }
    if (WALK == kWalkStatic) return;
    for (int m = 0; m <= L; ++m) {
        const double sfac = (m == 0) ? S.sth : s2;   // -sin th P (m = 0, :241) vs -sin^2 th Pt (:251)
        if (m == 1) {
            dgd = -ap.diagc[1] * (S.cth * dgt + S.sth * dgd);   // :229-230 (uses P_0^0, undivided)
            dgt = -ap.diagc[1] * dgt;
        } else if (m > 1) {
            const double nd = -ap.diagc[m] * (S.cth * dgt * S.sth + S.sth * dgd);   // :259-261
            dgt = -ap.diagc[m] * S.sth * dgt;
            dgd = nd;
        }
        if (m > 0) {
            const double nr = epr * S.cphi - epi * S.sphi;
            epi = epr * S.sphi + epi * S.cphi;
            epr = nr;
        }
        double p2 = 0.0, d2 = 0.0, p1 = dgt, d1 = dgd;
        for (int l = m; l <= L; ++l) {
            if (l > m) {
                const int ip = index_p(l, m);
                const double p = ap.A[ip] * (S.cth * p1 + ap.B[ip] * p2);                    // :236-238, :246-248, :255
                const double d = ap.A[ip] * (-sfac * p1 + S.cth * d1 + ap.B[ip] * d2);       // :239-243, :249-253, :256
                p2 = p1; d2 = d1; p1 = p; d1 = d;
            }
            f(l, m, p1, d1, epr, epi);
        }
    }
}

}  // namespace aceb200
