/*
 * ace_oracle.c -- CPU restatement of ACE.jl's basis/model evaluation.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file's library.  The product (ace_jl_b200/) never does; it fails loudly without CUDA.
 *
 * PARITY STATUS: the reference (ACEsuit/ACE.jl v0.12.46, 100 % Julia) cannot be run in the build
 * container (no Julia, ACEbase 0.2.4 not vendored) and ships no golden vectors for Rn, A, AA, B,
 * energies or forces.  This oracle is therefore pinned only by the reference's
 * own known-answer tests, re-expressed in tests/test_oracle_*.py:
 *   - closed-form Y_l^m for l <= 3 (test/polynomials/test_ylm.jl:14-33), also near the pole,
 *   - closed-form distance transforms (test/transforms/test_transforms.jl:21,34,49),
 *   - the one-hot categorical basis (test/test_discrete.jl:38-45),
 *   - and the reference's randomised identities (finite differences, A(cfg)=sum_j A(X_j),
 *     naive == product evaluator, rotation / permutation invariance),
 * and, independently of this repository's reading of the reference, by third-party arithmetic
 * (tests/test_oracle_independent.py): Y_l^m for all l <= 8 against scipy.special.sph_harm_y and mpmath at 50 digits,
 * grad Y_l^m against mpmath derivatives, and one whole model (A, AA, B, E, forces) re-evaluated with mpmath at 50 digits
 * from the same tables.  tests/golden/export_golden.jl dumps the same quantities from a real ACE.jl installation;
 * tests/test_julia_golden.py checks this oracle (and the CUDA path) against such dumps when they are present.
 * Until one is: "parity unpinned by the reference" for Rn, A, AA, B, energies and forces.
 *
 * Every function follows the reference line by line, INCLUDING its inefficiencies (the full
 * (maxL+1)^2 harmonics, the materialised dA matrix, atan/sincos) -- this is what the reference's
 * CPU path costs, and it is what bench.py times as the CPU baseline.  Environments are distributed
 * over OpenMP threads, which is the reference's user-level `Threads.@threads` over configurations
 * (src/utils/pools.jl:44-75).
 */
#include <complex.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/aceb200.h"

typedef double complex cplx;

/* ------------------------------------------------------------------------------------------
 * distance transforms: src/transforms/distancetransforms.jl:16-25; derivative = what ForwardDiff
 * returns for these expressions (src/transforms/lambdas.jl:28-34)
 * ---------------------------------------------------------------------------------------- */
static double ipow_or_pow(double x, double p)
{
    if (p == floor(p) && fabs(p) <= 64) {          /* Julia: x^Int is repeated multiplication */
        int n = (int)fabs(p);
        double y = 1.0;
        for (int i = 0; i < n; i++) y *= x;
        return p < 0 ? 1.0 / y : y;
    }
    return pow(x, p);
}

double oracle_transform(int kind, const double *q, double r)
{
    switch (kind) {
    case ACEB200_TRANS_ID:     return r;
    case ACEB200_TRANS_POLY:   return ipow_or_pow((1.0 + q[1]) / (1.0 + r), q[0]);
    case ACEB200_TRANS_MORSE:  return exp(-q[0] * (r / q[1] - 1.0));
    case ACEB200_TRANS_AGNESI: return 1.0 / (1.0 + q[2] * ipow_or_pow(r / q[0], q[1]));
    }
    return NAN;
}

double oracle_transform_d(int kind, const double *q, double r)
{
    switch (kind) {
    case ACEB200_TRANS_ID:   return 1.0;
    case ACEB200_TRANS_POLY: {
        double p = q[0], x = (1.0 + q[1]) / (1.0 + r);
        return p * ipow_or_pow(x, p - 1.0) * (-(1.0 + q[1]) / ((1.0 + r) * (1.0 + r)));
    }
    case ACEB200_TRANS_MORSE:
        return exp(-q[0] * (r / q[1] - 1.0)) * (-q[0] / q[1]);
    case ACEB200_TRANS_AGNESI: {
        double r0 = q[0], p = q[1], a = q[2], x = r / r0;
        double d = 1.0 + a * ipow_or_pow(x, p);
        return -(a * p * ipow_or_pow(x, p - 1.0) / r0) / (d * d);
    }
    }
    return NAN;
}

/* ------------------------------------------------------------------------------------------
 * radial basis: src/polynomials/orthpolys.jl:32-53 (envelope), :250-306 (recursions)
 * ---------------------------------------------------------------------------------------- */
static double fcut(int pl, double tl, int pr, double tr, double t)
{
    if ((pl > 0 && t < tl) || (pr > 0 && t > tr)) return 0.0;
    return ipow_or_pow(t - tl, pl) * ipow_or_pow(t - tr, pr);
}

static double fcut_d(int pl, double tl, int pr, double tr, double t)
{
    if ((pl > 0 && t < tl) || (pr > 0 && t > tr)) return 0.0;
    double a = 0.0, b = 0.0;
    if (pl > 0) a = pl * ipow_or_pow(t - tl, pl - 1) * ipow_or_pow(t - tr, pr);
    if (pr > 0) b = pr * ipow_or_pow(t - tl, pl) * ipow_or_pow(t - tr, pr - 1);
    return a + b;
}

/* evaluate!(P, J, t), orthpolys.jl:253-263 */
static void orthpoly_eval(const aceb200_desc *d, double t, double *P)
{
    int N = d->n_rad;
    P[0] = d->rad_A[0] * fcut(d->pl, d->tl, d->pr, d->tr, t);
    if (N == 1) return;
    P[1] = (d->rad_A[1] * t + d->rad_B[1]) * P[0];
    for (int n = 2; n < N; n++)
        P[n] = (d->rad_A[n] * t + d->rad_B[n]) * P[n - 1] + d->rad_C[n] * P[n - 2];
}

/* evaluate_ed!(P, dP, J, t), orthpolys.jl:287-306 */
static void orthpoly_eval_ed(const aceb200_desc *d, double t, double *P, double *dP)
{
    int N = d->n_rad;
    P[0] = d->rad_A[0] * fcut(d->pl, d->tl, d->pr, d->tr, t);
    dP[0] = d->rad_A[0] * fcut_d(d->pl, d->tl, d->pr, d->tr, t);
    if (N == 1) return;
    double al = d->rad_A[1] * t + d->rad_B[1];
    P[1] = al * P[0];
    dP[1] = al * dP[0] + d->rad_A[1] * P[0];
    for (int n = 2; n < N; n++) {
        al = d->rad_A[n] * t + d->rad_B[n];
        P[n] = al * P[n - 1] + d->rad_C[n] * P[n - 2];
        dP[n] = al * dP[n - 1] + d->rad_C[n] * dP[n - 2] + d->rad_A[n] * P[n - 1];
    }
}

/* Rn = chain(norm, trans, OrthPoly) (src/b1pcomponents/Rn.jl:21, src/chain.jl:22-52) */
void oracle_rn(const aceb200_desc *d, const double *rr, double *P)
{
    double r = sqrt(rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2]);
    orthpoly_eval(d, oracle_transform(d->trans_kind, d->trans_par, r), P);
}

/* value and Cartesian gradient dP[n][3] = dP_n/dt * t'(r) * rr/r (orthpolys.jl:242-247) */
void oracle_rn_ed(const aceb200_desc *d, const double *rr, double *P, double *dP3)
{
    int N = d->n_rad;
    double r = sqrt(rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2]);
    double t = oracle_transform(d->trans_kind, d->trans_par, r);
    double dt = oracle_transform_d(d->trans_kind, d->trans_par, r);
    double *dP = (double *)malloc(sizeof(double) * N);
    orthpoly_eval_ed(d, t, P, dP);
    double g[3] = { dt * (rr[0] / r), dt * (rr[1] / r), dt * (rr[2] / r) };
    for (int n = 0; n < N; n++)
        for (int k = 0; k < 3; k++) dP3[3 * n + k] = dP[n] * g[k];
    free(dP);
}

/* ------------------------------------------------------------------------------------------
 * spherical harmonics: src/polynomials/sphericalharmonics.jl
 * ---------------------------------------------------------------------------------------- */
typedef struct { double r, cosphi, sinphi, costh, sinth; } sphc;

static int index_p(int l, int m) { return m + (l * (l + 1)) / 2; }      /* :93, 0-based here  */
static int index_y(int l, int m) { return m + l + l * l; }              /* :102, 0-based here */

static sphc cart2spher(const double *R)                                 /* :43-51 */
{
    sphc S;
    S.r = sqrt(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
    double phi = atan2(R[1], R[0]);
    S.sinphi = sin(phi);
    S.cosphi = cos(phi);
    S.costh = R[2] / S.r;
    S.sinth = sqrt(R[0] * R[0] + R[1] * R[1]) / S.r;
    return S;
}

static void alp_coeffs(int L, double *A, double *B)                     /* :146-159 */
{
    int n = (L + 1) * (L + 2) / 2;
    for (int i = 0; i < n; i++) A[i] = B[i] = 0.0;
    for (int l = 2; l <= L; l++) {
        double ls = (double)l * l, lm1s = (double)(l - 1) * (l - 1);
        for (int m = 0; m <= l - 2; m++) {
            double ms = (double)m * m;
            A[index_p(l, m)] = sqrt((4 * ls - 1.0) / (ls - ms));
            B[index_p(l, m)] = -sqrt((lm1s - ms) / (4 * lm1s - 1.0));
        }
    }
}

static void alp_eval(int L, const double *A, const double *B, sphc S, double *P)   /* :167-197 */
{
    double temp = sqrt(0.5 / M_PI);
    P[index_p(0, 0)] = temp;
    if (L == 0) return;
    P[index_p(1, 0)] = S.costh * sqrt(3.0) * temp;
    temp = -sqrt(1.5) * S.sinth * temp;
    P[index_p(1, 1)] = temp;
    for (int l = 2; l <= L; l++) {
        int il = (l * (l + 1)) / 2, ilm1 = il - l, ilm2 = ilm1 - l + 1;
        for (int m = 0; m <= l - 2; m++)
            P[il + m] = A[il + m] * (S.costh * P[ilm1 + m] + B[il + m] * P[ilm2 + m]);
        P[il + l - 1] = S.costh * sqrt(2.0 * (l - 1) + 3.0) * temp;
        temp = -sqrt(1.0 + 0.5 / l) * S.sinth * temp;
        P[il + l] = temp;
    }
}

/* P holds P (m = 0) or P / sin(theta) (m > 0); dP holds dP/dtheta (:212-267) */
static void alp_eval_ed(int L, const double *A, const double *B, sphc S, double *P, double *dP)
{
    double temp = sqrt(0.5 / M_PI), temp_d = 0.0, temp1;
    P[index_p(0, 0)] = temp;
    dP[index_p(0, 0)] = temp_d;
    if (L == 0) return;
    P[index_p(1, 0)] = S.costh * sqrt(3.0) * temp;
    dP[index_p(1, 0)] = -S.sinth * sqrt(3.0) * temp + S.costh * sqrt(3.0) * temp_d;
    temp1 = -sqrt(1.5) * temp;
    temp_d = -sqrt(1.5) * (S.costh * temp + S.sinth * temp_d);
    P[index_p(1, 1)] = temp1;
    dP[index_p(1, 1)] = temp_d;
    for (int l = 2; l <= L; l++) {
        int m = 0;
        P[index_p(l, m)] = A[index_p(l, m)] * (S.costh * P[index_p(l - 1, m)] + B[index_p(l, m)] * P[index_p(l - 2, m)]);
        dP[index_p(l, m)] = A[index_p(l, m)] * (-S.sinth * P[index_p(l - 1, m)] + S.costh * dP[index_p(l - 1, m)]
                                                 + B[index_p(l, m)] * dP[index_p(l - 2, m)]);
        for (m = 1; m <= l - 2; m++) {
            P[index_p(l, m)] = A[index_p(l, m)] * (S.costh * P[index_p(l - 1, m)] + B[index_p(l, m)] * P[index_p(l - 2, m)]);
            dP[index_p(l, m)] = A[index_p(l, m)] * (-(S.sinth * S.sinth) * P[index_p(l - 1, m)] + S.costh * dP[index_p(l - 1, m)]
                                                     + B[index_p(l, m)] * dP[index_p(l - 2, m)]);
        }
        P[index_p(l, l - 1)] = sqrt(2.0 * (l - 1) + 3.0) * S.costh * temp1;
        dP[index_p(l, l - 1)] = sqrt(2.0 * (l - 1) + 3.0) * (-(S.sinth * S.sinth) * temp1 + S.costh * temp_d);
        double c = -sqrt(1.0 + 0.5 / l);
        double n1 = c * S.sinth * temp1;
        double nd = c * (S.costh * temp1 * S.sinth + S.sinth * temp_d);
        temp1 = n1;
        temp_d = nd;
        P[index_p(l, l)] = temp1;
        dP[index_p(l, l)] = temp_d;
    }
}

/* cYlm!, :379-403 */
static void cylm(int L, sphc S, const double *P, cplx *Y)
{
    cplx ep = 1.0 / sqrt(2.0);
    for (int l = 0; l <= L; l++) Y[index_y(l, 0)] = P[index_p(l, 0)] * ep;
    double sig = 1.0;
    cplx ep_fact = S.cosphi + I * S.sinphi;
    for (int m = 1; m <= L; m++) {
        sig *= -1.0;
        ep *= ep_fact;
        cplx em = sig * conj(ep);
        for (int l = m; l <= L; l++) {
            double p = P[index_p(l, m)];
            Y[index_y(l, -m)] = em * p;
            Y[index_y(l, m)] = ep * p;
        }
    }
}

/* dspher_to_dcart, :60-65 */
static void dspher_to_dcart(sphc S, cplx f_phi_div_sinth, cplx f_th, cplx *out)
{
    double r = S.r + DBL_EPSILON;
    out[0] = (-(S.sinphi * f_phi_div_sinth) + (S.cosphi * S.costh * f_th)) / r;
    out[1] = ((S.cosphi * f_phi_div_sinth) + (S.sinphi * S.costh * f_th)) / r;
    out[2] = (-(S.sinth * f_th)) / r;
}

/* cYlm_ed!, :410-443 */
static void cylm_ed(int L, sphc S, const double *P, const double *dP, cplx *Y, cplx *dY)
{
    cplx ep = 1.0 / sqrt(2.0);
    for (int l = 0; l <= L; l++) {
        Y[index_y(l, 0)] = P[index_p(l, 0)] * ep;
        dspher_to_dcart(S, 0.0, dP[index_p(l, 0)] * ep, dY + 3 * index_y(l, 0));
    }
    double sig = 1.0;
    cplx ep_fact = S.cosphi + I * S.sinphi;
    for (int m = 1; m <= L; m++) {
        sig *= -1.0;
        ep *= ep_fact;
        cplx em = sig * conj(ep);
        cplx dep_dphi = I * (double)m * ep;
        cplx dem_dphi = I * (double)(-m) * em;
        for (int l = m; l <= L; l++) {
            double p_div_sinth = P[index_p(l, m)];
            Y[index_y(l, -m)] = em * p_div_sinth * S.sinth;
            Y[index_y(l, m)] = ep * p_div_sinth * S.sinth;
            double dp_dth = dP[index_p(l, m)];
            dspher_to_dcart(S, dem_dphi * p_div_sinth, em * dp_dth, dY + 3 * index_y(l, -m));
            dspher_to_dcart(S, dep_dphi * p_div_sinth, ep * dp_dth, dY + 3 * index_y(l, m));
        }
    }
}

/* evaluate!(Y, SH, R), :340-348.  Y: (L+1)^2 complex */
void oracle_ylm(int L, const double *R, double *Yout)
{
    int nP = (L + 1) * (L + 2) / 2;
    double *A = (double *)malloc(sizeof(double) * nP * 3), *B = A + nP, *P = B + nP;
    alp_coeffs(L, A, B);
    sphc S = cart2spher(R);
    alp_eval(L, A, B, S, P);
    cylm(L, S, P, (cplx *)Yout);
    free(A);
}

/* evaluate_ed!(Y, dY, SH, R), :364-373.  dY: (L+1)^2 x 3 complex */
void oracle_ylm_ed(int L, const double *R, double *Yout, double *dYout)
{
    int nP = (L + 1) * (L + 2) / 2;
    double *A = (double *)malloc(sizeof(double) * nP * 4), *B = A + nP, *P = B + nP, *dP = P + nP;
    alp_coeffs(L, A, B);
    sphc S = cart2spher(R);
    alp_eval_ed(L, A, B, S, P, dP);
    cylm_ed(L, S, P, dP, (cplx *)Yout, (cplx *)dYout);
    free(A);
}

/* ------------------------------------------------------------------------------------------
 * per-thread scratch
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int N, nY, nP;
    double *alpA, *alpB, *P, *dP;   /* ALP */
    double *Rn, *dRn3;              /* N, N x 3 */
    cplx *Y, *dY;                   /* nY, nY x 3 */
    double *onehot;                 /* n_cat */
    cplx *A, *dA;                   /* nA, nA x J x 3 (column-major over [iA, j]) */
    cplx *dAAdA;
    cplx *dAco;
    int Jcap;
} scratch;

static void scratch_init(scratch *s, const aceb200_desc *d)
{
    memset(s, 0, sizeof(*s));
    s->N = d->n_rad;
    s->nY = (d->maxL + 1) * (d->maxL + 1);
    s->nP = (d->maxL + 1) * (d->maxL + 2) / 2;
    s->alpA = (double *)malloc(sizeof(double) * s->nP * 4);
    s->alpB = s->alpA + s->nP; s->P = s->alpB + s->nP; s->dP = s->P + s->nP;
    alp_coeffs(d->maxL, s->alpA, s->alpB);
    s->Rn = (double *)malloc(sizeof(double) * s->N * 4);
    s->dRn3 = s->Rn + s->N;
    s->Y = (cplx *)malloc(sizeof(cplx) * s->nY * 4);
    s->dY = s->Y + s->nY;
    s->onehot = (double *)calloc(d->n_cat > 0 ? d->n_cat : 1, sizeof(double));
    s->A = (cplx *)malloc(sizeof(cplx) * d->nA);
    s->dAAdA = (cplx *)malloc(sizeof(cplx) * (d->maxord > 0 ? d->maxord : 1));
    s->dAco = (cplx *)malloc(sizeof(cplx) * (size_t)d->nA * d->nprop * d->ncomp);
    s->dA = NULL; s->Jcap = 0;
}

static void scratch_need_dA(scratch *s, const aceb200_desc *d, int J)
{
    if (J > s->Jcap) {
        free(s->dA);
        s->dA = (cplx *)malloc(sizeof(cplx) * 3 * (size_t)d->nA * J);
        s->Jcap = J;
    }
}

static void scratch_free(scratch *s)
{
    free(s->alpA); free(s->Rn); free(s->Y); free(s->onehot); free(s->A); free(s->dAAdA); free(s->dAco); free(s->dA);
}

/* ------------------------------------------------------------------------------------------
 * one-particle basis: src/product_1pbasis.jl
 * ---------------------------------------------------------------------------------------- */

/* evaluate the components for one state X = (rr, species); with_d: also gradients */
static int eval_components(const aceb200_desc *d, scratch *s, const double *rr, int species, int with_d)
{
    for (int ib = 0; ib < d->n_comp; ib++) {
        switch (d->comp_kind[ib]) {
        case ACEB200_COMP_RN:
            if (with_d) oracle_rn_ed(d, rr, s->Rn, s->dRn3); else oracle_rn(d, rr, s->Rn);
            break;
        case ACEB200_COMP_YLM: {
            sphc S = cart2spher(rr);
            if (with_d) { alp_eval_ed(d->maxL, s->alpA, s->alpB, S, s->P, s->dP); cylm_ed(d->maxL, S, s->P, s->dP, s->Y, s->dY); }
            else { alp_eval(d->maxL, s->alpA, s->alpB, S, s->P); cylm(d->maxL, S, s->P, s->Y); }
            break;
        }
        case ACEB200_COMP_CAT:            /* discrete1pbasis.jl:108-112 */
            if (species < 1 || species > d->n_cat) return ACEB200_ECATEGORY;
            for (int q = 0; q < d->n_cat; q++) s->onehot[q] = 0.0;
            s->onehot[species - 1] = 1.0;
            break;
        default: return ACEB200_EUNSUPPORTED;
        }
    }
    return 0;
}

static cplx comp_val(const aceb200_desc *d, const scratch *s, int ib, int idx1)
{
    switch (d->comp_kind[ib]) {
    case ACEB200_COMP_RN:  return s->Rn[idx1 - 1];
    case ACEB200_COMP_YLM: return s->Y[idx1 - 1];
    default:               return s->onehot[idx1 - 1];
    }
}

/* add_into_A!, product_1pbasis.jl:99-118 */
static int add_into_A(const aceb200_desc *d, scratch *s, const double *rr, int species, cplx *A)
{
    int rc = eval_components(d, s, rr, species, 0);
    if (rc) return rc;
    for (int iA = 0; iA < d->nA; iA++) {
        const int32_t *phi = d->indices + (size_t)iA * d->n_comp;
        cplx v = comp_val(d, s, 0, phi[0]);
        for (int ib = 1; ib < d->n_comp; ib++) v *= comp_val(d, s, ib, phi[ib]);
        A[iA] += v;
    }
    return 0;
}

/* _add_into_A_dA!, product_1pbasis.jl:169-221; dAcol is nA x 3 */
static int add_into_A_dA(const aceb200_desc *d, scratch *s, const double *rr, int species, cplx *A, cplx *dAcol)
{
    int rc = eval_components(d, s, rr, species, 1);
    if (rc) return rc;
    for (int iA = 0; iA < d->nA; iA++) {
        const int32_t *phi = d->indices + (size_t)iA * d->n_comp;
        cplx v = comp_val(d, s, 0, phi[0]);
        for (int ib = 1; ib < d->n_comp; ib++) v *= comp_val(d, s, ib, phi[ib]);
        A[iA] += v;
        cplx g[3] = { 0, 0, 0 };
        for (int a = 0; a < d->n_comp; a++) {
            if (d->comp_kind[a] == ACEB200_COMP_CAT) continue;
            cplx dt[3];
            for (int k = 0; k < 3; k++)
                dt[k] = (d->comp_kind[a] == ACEB200_COMP_RN) ? (cplx)s->dRn3[3 * (phi[a] - 1) + k]
                                                               : s->dY[3 * (phi[a] - 1) + k];
            for (int b = 0; b < d->n_comp; b++) {
                if (b == a) continue;
                cplx Bb = comp_val(d, s, b, phi[b]);
                for (int k = 0; k < 3; k++) dt[k] *= Bb;
            }
            for (int k = 0; k < 3; k++) g[k] += dt[k];
        }
        for (int k = 0; k < 3; k++) dAcol[3 * iA + k] = g[k];
    }
    return 0;
}

/* evaluate(basis1p, cfg), product_1pbasis.jl:123-134 */
static int eval_A_env(const aceb200_desc *d, scratch *s, const aceb200_batch *b, int64_t e, cplx *A)
{
    int64_t j0 = b->offsets[e], j1 = b->offsets[e + 1];
    if (j1 <= j0) return ACEB200_EEMPTY;
    for (int i = 0; i < d->nA; i++) A[i] = 0;
    for (int64_t j = j0; j < j1; j++) {
        int rc = add_into_A(d, s, b->R + 3 * j, b->species ? b->species[j] : 0, A);
        if (rc) return rc;
    }
    return 0;
}

/* evaluate_ed(basis1p, cfg), product_1pbasis.jl:234-244; dA [j][iA][3] */
static int eval_A_dA_env(const aceb200_desc *d, scratch *s, const aceb200_batch *b, int64_t e, cplx *A, cplx *dA)
{
    int64_t j0 = b->offsets[e], j1 = b->offsets[e + 1];
    if (j1 <= j0) return ACEB200_EEMPTY;
    for (int i = 0; i < d->nA; i++) A[i] = 0;
    for (int64_t j = j0; j < j1; j++) {
        int rc = add_into_A_dA(d, s, b->R + 3 * j, b->species ? b->species[j] : 0, A, dA + 3 * (size_t)d->nA * (j - j0));
        if (rc) return rc;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * product basis: src/pibasis.jl
 * ---------------------------------------------------------------------------------------- */
#define IAA(d, i, t) ((d)->iAA2iA[(size_t)(t) * (d)->nAA + (i)])   /* column-major, value is 1-based */

/* evaluate!(AA, basis, A), pibasis.jl:265-275; out real (pireal) or complex */
static void eval_AA_from_A(const aceb200_desc *d, const cplx *A, double *AAout)
{
    for (int i = 0; i < d->nAA; i++) {
        cplx aa = 1.0;
        for (int t = 0; t < d->orders[i]; t++) aa *= A[IAA(d, i, t) - 1];
        if (d->pireal) AAout[i] = creal(aa);
        else { AAout[2 * i] = creal(aa); AAout[2 * i + 1] = cimag(aa); }
    }
}

/* _AA_local_adjoints!, pibasis.jl:338-390; returns the (unprojected) product */
static cplx AA_local_adjoints(const aceb200_desc *d, const cplx *A, int i, int ord, cplx *dAAdA)
{
    if (ord == 1) { dAAdA[0] = 1.0; return A[IAA(d, i, 0) - 1]; }
    if (ord == 2) {
        cplx A1 = A[IAA(d, i, 0) - 1], A2 = A[IAA(d, i, 1) - 1];
        dAAdA[0] = A2; dAAdA[1] = A1;
        return A1 * A2;
    }
    cplx A1 = A[IAA(d, i, 0) - 1], A2 = A[IAA(d, i, 1) - 1];
    dAAdA[0] = 1.0; dAAdA[1] = A1;
    cplx fwd = A1 * A2;
    for (int a = 2; a < ord - 1; a++) { dAAdA[a] = fwd; fwd *= A[IAA(d, i, a) - 1]; }
    dAAdA[ord - 1] = fwd;
    cplx Aend = A[IAA(d, i, ord - 1) - 1];
    cplx aa = fwd * Aend;
    cplx bwd = Aend;
    for (int a = ord - 2; a >= 2; a--) { dAAdA[a] *= bwd; bwd *= A[IAA(d, i, a) - 1]; }
    dAAdA[1] *= bwd;
    bwd *= A2;
    dAAdA[0] *= bwd;
    return aa;
}

/* _evaluate_ed!, pibasis.jl:402-432.  dAAout: [j][iAA][3] real or complex */
static void eval_AA_dAA(const aceb200_desc *d, scratch *s, const cplx *A, const cplx *dA, int J, double *AAout, double *dAAout)
{
    int cs = d->pireal ? 1 : 2;
    int i0 = 0;
    if (d->nAA > 0 && d->orders[0] == 0) {
        i0 = 1;
        if (AAout) { AAout[0] = 1.0; if (cs == 2) AAout[1] = 0.0; }
        for (int j = 0; j < J; j++)
            for (int k = 0; k < 3 * cs; k++) dAAout[((size_t)j * d->nAA + 0) * 3 * cs + k] = 0.0;
    }
    for (int i = i0; i < d->nAA; i++) {
        int ord = d->orders[i];
        cplx aa = AA_local_adjoints(d, A, i, ord, s->dAAdA);
        if (AAout) { if (cs == 1) AAout[i] = creal(aa); else { AAout[2 * i] = creal(aa); AAout[2 * i + 1] = cimag(aa); } }
        for (int j = 0; j < J; j++) {
            cplx g[3] = { 0, 0, 0 };
            for (int a = 0; a < ord; a++) {
                const cplx *da = dA + 3 * ((size_t)d->nA * j + (IAA(d, i, a) - 1));
                for (int k = 0; k < 3; k++) g[k] += s->dAAdA[a] * da[k];
            }
            double *o = dAAout + ((size_t)j * d->nAA + i) * 3 * cs;
            for (int k = 0; k < 3; k++) {
                if (cs == 1) o[k] = creal(g[k]);
                else { o[2 * k] = creal(g[k]); o[2 * k + 1] = cimag(g[k]); }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * symmetric basis: genmul!, src/symmbasis.jl:248-264, with the mulops of :312-316 and :330-334
 * ---------------------------------------------------------------------------------------- */
static cplx get_AA(const aceb200_desc *d, const double *AA, size_t i)
{
    return d->pireal ? (cplx)AA[i] : AA[2 * i] + I * AA[2 * i + 1];
}

static void eval_B_from_AA(const aceb200_desc *d, const double *AA, double *Bout)
{
    int cs = d->symreal ? 1 : 2;
    size_t n = (size_t)d->nB * d->ncomp * cs;
    for (size_t k = 0; k < n; k++) Bout[k] = 0.0;
    const cplx *nz = (const cplx *)d->nzval;
    for (int col = 0; col < d->nAA; col++) {
        cplx x = get_AA(d, AA, col);
        for (int k = d->colptr[col] - 1; k < d->colptr[col + 1] - 1; k++) {
            int row = d->rowval[k] - 1;
            for (int c = 0; c < d->ncomp; c++) {
                cplx v = nz[(size_t)k * d->ncomp + c] * x;
                double *o = Bout + ((size_t)row * d->ncomp + c) * cs;
                o[0] += creal(v);
                if (cs == 2) o[1] += cimag(v);
            }
        }
    }
}

/* dB[j][iB][xyz][comp] += real?(nz[comp] * dAA[j][col][xyz]) */
static void eval_dB_from_dAA(const aceb200_desc *d, const double *dAA, int J, double *dBout)
{
    int cs = d->symreal ? 1 : 2, ca = d->pireal ? 1 : 2;
    size_t per_j = (size_t)d->nB * 3 * d->ncomp * cs;
    for (size_t k = 0; k < per_j * J; k++) dBout[k] = 0.0;
    const cplx *nz = (const cplx *)d->nzval;
    for (int j = 0; j < J; j++) {
        for (int col = 0; col < d->nAA; col++) {
            const double *x = dAA + ((size_t)j * d->nAA + col) * 3 * ca;
            for (int k = d->colptr[col] - 1; k < d->colptr[col + 1] - 1; k++) {
                int row = d->rowval[k] - 1;
                for (int xyz = 0; xyz < 3; xyz++) {
                    cplx xv = (ca == 1) ? (cplx)x[xyz] : x[2 * xyz] + I * x[2 * xyz + 1];
                    for (int c = 0; c < d->ncomp; c++) {
                        cplx v = nz[(size_t)k * d->ncomp + c] * xv;
                        double *o = dBout + per_j * j + (((size_t)row * 3 + xyz) * d->ncomp + c) * cs;
                        o[0] += creal(v);
                        if (cs == 2) o[1] += cimag(v);
                    }
                }
            }
        }
    }
}

/* c~ = transpose(A2Bmap) * c, src/evaluator.jl:59-60 with src/symmbasis.jl:267-285.
 * ctilde: [nAA][nprop][ncomp] complex */
int oracle_eff_coeffs(const aceb200_desc *d, const double *c, double *ctilde_out)
{
    cplx *ct = (cplx *)ctilde_out;
    const cplx *nz = (const cplx *)d->nzval;
    for (int col = 0; col < d->nAA; col++) {
        for (int p = 0; p < d->nprop; p++)
            for (int cc = 0; cc < d->ncomp; cc++) {
                cplx tmp = 0;
                for (int k = d->colptr[col] - 1; k < d->colptr[col + 1] - 1; k++)
                    tmp += nz[(size_t)k * d->ncomp + cc] * (c ? c[(size_t)(d->rowval[k] - 1) * d->nprop + p] : 0.0);
                ct[((size_t)col * d->nprop + p) * d->ncomp + cc] = tmp;
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * batch drivers
 * ---------------------------------------------------------------------------------------- */
#define OMP_ENV_LOOP_BEGIN                                                   \
    int rc_all = 0;                                                          \
    _Pragma("omp parallel")                                                  \
    {                                                                        \
        scratch s; scratch_init(&s, d);                                      \
        _Pragma("omp for schedule(dynamic, 8)")                              \
        for (int64_t e = 0; e < b->nenv; e++) {                              \
            int rc = 0;

#define OMP_ENV_LOOP_END                                                     \
            if (rc) { _Pragma("omp critical") { if (!rc_all) rc_all = rc; } } \
        }                                                                    \
        scratch_free(&s);                                                    \
    }                                                                        \
    return rc_all;

int oracle_eval_A(const aceb200_desc *d, const aceb200_batch *b, double *A)
{
    OMP_ENV_LOOP_BEGIN
        rc = eval_A_env(d, &s, b, e, (cplx *)A + (size_t)e * d->nA);
    OMP_ENV_LOOP_END
}

int oracle_eval_AA(const aceb200_desc *d, const aceb200_batch *b, double *AA)
{
    int cs = d->pireal ? 1 : 2;
    OMP_ENV_LOOP_BEGIN
        rc = eval_A_env(d, &s, b, e, s.A);
        if (!rc) eval_AA_from_A(d, s.A, AA + (size_t)e * d->nAA * cs);
    OMP_ENV_LOOP_END
}

int oracle_eval_B(const aceb200_desc *d, const aceb200_batch *b, double *B)
{
    int cs = d->symreal ? 1 : 2, ca = d->pireal ? 1 : 2;
    OMP_ENV_LOOP_BEGIN
        double *AA = (double *)malloc(sizeof(double) * d->nAA * ca);
        rc = eval_A_env(d, &s, b, e, s.A);
        if (!rc) {
            eval_AA_from_A(d, s.A, AA);
            eval_B_from_AA(d, AA, B + (size_t)e * d->nB * d->ncomp * cs);
        }
        free(AA);
    OMP_ENV_LOOP_END
}

int oracle_eval_dA(const aceb200_desc *d, const aceb200_batch *b, double *A, double *dA)
{
    OMP_ENV_LOOP_BEGIN
        rc = eval_A_dA_env(d, &s, b, e, s.A, (cplx *)dA + 3 * (size_t)d->nA * b->offsets[e]);
        if (!rc && A) memcpy((cplx *)A + (size_t)e * d->nA, s.A, sizeof(cplx) * d->nA);
    OMP_ENV_LOOP_END
}

int oracle_eval_dAA(const aceb200_desc *d, const aceb200_batch *b, double *AA, double *dAA)
{
    int ca = d->pireal ? 1 : 2;
    OMP_ENV_LOOP_BEGIN
        int J = (int)(b->offsets[e + 1] - b->offsets[e]);
        scratch_need_dA(&s, d, J);
        rc = eval_A_dA_env(d, &s, b, e, s.A, s.dA);
        if (!rc) eval_AA_dAA(d, &s, s.A, s.dA, J, AA ? AA + (size_t)e * d->nAA * ca : NULL,
                             dAA + (size_t)b->offsets[e] * d->nAA * 3 * ca);
    OMP_ENV_LOOP_END
}

int oracle_eval_dB(const aceb200_desc *d, const aceb200_batch *b, double *B, double *dB)
{
    int cs = d->symreal ? 1 : 2, ca = d->pireal ? 1 : 2;
    OMP_ENV_LOOP_BEGIN
        int J = (int)(b->offsets[e + 1] - b->offsets[e]);
        scratch_need_dA(&s, d, J);
        rc = eval_A_dA_env(d, &s, b, e, s.A, s.dA);
        if (!rc) {
            double *AA = (double *)malloc(sizeof(double) * d->nAA * ca);
            double *dAA = (double *)malloc(sizeof(double) * (size_t)d->nAA * J * 3 * ca);
            eval_AA_dAA(d, &s, s.A, s.dA, J, AA, dAA);
            if (B) eval_B_from_AA(d, AA, B + (size_t)e * d->nB * d->ncomp * cs);
            eval_dB_from_dAA(d, dAA, J, dB + (size_t)b->offsets[e] * d->nB * 3 * d->ncomp * cs);
            free(AA); free(dAA);
        }
    OMP_ENV_LOOP_END
}

/* evaluate(V::ProductEvaluator, cfg), src/evaluator.jl:121-147.  ct: [nAA][nprop][ncomp] complex */
static void energy_from_A(const aceb200_desc *d, const cplx *ct, const cplx *A, double *Eout)
{
    int cs = d->symreal ? 1 : 2, P = d->nprop * d->ncomp;
    cplx *val = (cplx *)calloc(P, sizeof(cplx));
    int i0 = 0;
    if (d->nAA > 0 && d->orders[0] == 0) { for (int q = 0; q < P; q++) val[q] += ct[q]; i0 = 1; }
    for (int i = i0; i < d->nAA; i++) {
        cplx aa = A[IAA(d, i, 0) - 1];
        for (int t = 1; t < d->orders[i]; t++) aa *= A[IAA(d, i, t) - 1];
        if (d->pireal) aa = creal(aa);
        for (int q = 0; q < P; q++) {
            cplx v = aa * ct[(size_t)i * P + q];
            val[q] += d->symreal ? (cplx)creal(v) : v;
        }
    }
    for (int q = 0; q < P; q++) { Eout[q * cs] = creal(val[q]); if (cs == 2) Eout[q * cs + 1] = cimag(val[q]); }
    free(val);
}

int oracle_energy(const aceb200_desc *d, const aceb200_batch *b, const double *ctilde, double *E)
{
    int cs = d->symreal ? 1 : 2, P = d->nprop * d->ncomp;
    OMP_ENV_LOOP_BEGIN
        rc = eval_A_env(d, &s, b, e, s.A);
        if (!rc) energy_from_A(d, (const cplx *)ctilde, s.A, E + (size_t)e * P * cs);
    OMP_ENV_LOOP_END
}

/* _rrule_evaluate(_One(), m, V, cfg) = grad_config, src/evaluator.jl:161-200.
 * G: [j][prop][xyz][comp], real if symreal else complex. */
int oracle_energy_forces(const aceb200_desc *d, const aceb200_batch *b, const double *ctilde, double *E, double *G)
{
    int cs = d->symreal ? 1 : 2, P = d->nprop * d->ncomp;
    const cplx *ct = (const cplx *)ctilde;
    OMP_ENV_LOOP_BEGIN
        int J = (int)(b->offsets[e + 1] - b->offsets[e]);
        scratch_need_dA(&s, d, J);
        /* stage 1 */
        rc = eval_A_dA_env(d, &s, b, e, s.A, s.dA);
        if (!rc) {
            if (E) energy_from_A(d, ct, s.A, E + (size_t)e * P * cs);
            /* stage 2 */
            for (size_t k = 0; k < (size_t)d->nA * P; k++) s.dAco[k] = 0;
            int i0 = (d->nAA > 0 && d->orders[0] == 0) ? 1 : 0;
            for (int i = i0; i < d->nAA; i++) {
                int ord = d->orders[i];
                AA_local_adjoints(d, s.A, i, ord, s.dAAdA);
                for (int t = 0; t < ord; t++) {
                    cplx *dst = s.dAco + (size_t)(IAA(d, i, t) - 1) * P;
                    for (int q = 0; q < P; q++) dst[q] += s.dAAdA[t] * ct[(size_t)i * P + q];
                }
            }
            /* stage 3 */
            double *g = G + (size_t)b->offsets[e] * d->nprop * 3 * d->ncomp * cs;
            for (size_t k = 0; k < (size_t)J * d->nprop * 3 * d->ncomp * cs; k++) g[k] = 0.0;
            for (int j = 0; j < J; j++)
                for (int iA = 0; iA < d->nA; iA++) {
                    const cplx *da = s.dA + 3 * ((size_t)d->nA * j + iA);
                    for (int p = 0; p < d->nprop; p++)
                        for (int xyz = 0; xyz < 3; xyz++)
                            for (int c = 0; c < d->ncomp; c++) {
                                cplx v = s.dAco[(size_t)iA * P + p * d->ncomp + c] * da[xyz];
                                double *o = g + ((((size_t)j * d->nprop + p) * 3 + xyz) * d->ncomp + c) * cs;
                                o[0] += creal(v);
                                if (cs == 2) o[1] += cimag(v);
                            }
                }
        }
    OMP_ENV_LOOP_END
}

/* grad_config(m, NaiveEvaluator, cfg) and evaluate(m, NaiveEvaluator, cfg), src/linearmodel.jl:141-158:
 * E = sum_i c_i B_i,  g_j = sum_i c_i dB[i, j].  Used by tests as the reference's own cross-check
 * (test/test_linearmodel.jl:47-78). */
int oracle_naive_energy_forces(const aceb200_desc *d, const aceb200_batch *b, const double *c, double *E, double *G)
{
    int cs = d->symreal ? 1 : 2, ca = d->pireal ? 1 : 2;
    OMP_ENV_LOOP_BEGIN
        int J = (int)(b->offsets[e + 1] - b->offsets[e]);
        scratch_need_dA(&s, d, J);
        rc = eval_A_dA_env(d, &s, b, e, s.A, s.dA);
        if (!rc) {
            double *AA = (double *)malloc(sizeof(double) * d->nAA * ca);
            double *dAA = (double *)malloc(sizeof(double) * (size_t)d->nAA * J * 3 * ca);
            double *B = (double *)malloc(sizeof(double) * (size_t)d->nB * d->ncomp * cs);
            double *dB = (double *)malloc(sizeof(double) * (size_t)d->nB * J * 3 * d->ncomp * cs);
            eval_AA_dAA(d, &s, s.A, s.dA, J, AA, dAA);
            eval_B_from_AA(d, AA, B);
            eval_dB_from_dAA(d, dAA, J, dB);
            for (int p = 0; p < d->nprop; p++)
                for (int cc = 0; cc < d->ncomp * cs; cc++) {
                    double acc = 0.0;
                    for (int i = 0; i < d->nB; i++) acc += c[(size_t)i * d->nprop + p] * B[(size_t)i * d->ncomp * cs + cc];
                    if (E) E[((size_t)e * d->nprop + p) * d->ncomp * cs + cc] = acc;
                }
            size_t per = (size_t)3 * d->ncomp * cs;
            double *g = G + (size_t)b->offsets[e] * d->nprop * per;
            for (int j = 0; j < J; j++)
                for (int p = 0; p < d->nprop; p++)
                    for (size_t k = 0; k < per; k++) {
                        double acc = 0.0;
                        for (int i = 0; i < d->nB; i++)
                            acc += c[(size_t)i * d->nprop + p] * dB[((size_t)j * d->nB + i) * per + k];
                        g[((size_t)j * d->nprop + p) * per + k] = acc;
                    }
            free(AA); free(dAA); free(B); free(dB);
        }
    OMP_ENV_LOOP_END
}

/* adjoint_EVAL_D(m, V::ProductEvaluator, cfg, w), src/evaluator.jl:204-244.
 * w: [neighbour][3] real; out: [nenv][nB][ncomp] complex */
int oracle_adjoint_eval_d(const aceb200_desc *d, const aceb200_batch *b, const double *w, double *out)
{
    OMP_ENV_LOOP_BEGIN
        int J = (int)(b->offsets[e + 1] - b->offsets[e]);
        scratch_need_dA(&s, d, J);
        rc = eval_A_dA_env(d, &s, b, e, s.A, s.dA);           /* [1] */
        if (!rc) {
            cplx *dAw = (cplx *)calloc(d->nA, sizeof(cplx));
            cplx *dAAw = (cplx *)calloc(d->nAA, sizeof(cplx));
            const double *we = w + 3 * b->offsets[e];
            for (int k = 0; k < d->nA; k++)
                for (int j = 0; j < J; j++)
                    for (int x = 0; x < 3; x++) dAw[k] += we[3 * j + x] * s.dA[3 * ((size_t)d->nA * j + k) + x];   /* contract: no conjugation */
            int i0 = (d->nAA > 0 && d->orders[0] == 0) ? 1 : 0;    /* [2] */
            for (int i = i0; i < d->nAA; i++) {
                int ord = d->orders[i];
                AA_local_adjoints(d, s.A, i, ord, s.dAAdA);
                for (int t = 0; t < ord; t++) {
                    cplx v = dAw[IAA(d, i, t) - 1] * s.dAAdA[t];
                    dAAw[i] += d->symreal ? (cplx)creal(v) : v;
                }
            }
            const cplx *nz = (const cplx *)d->nzval;              /* [3] dB = A2Bmap * dAAw */
            cplx *o = (cplx *)out + (size_t)e * d->nB * d->ncomp;
            for (size_t k = 0; k < (size_t)d->nB * d->ncomp; k++) o[k] = 0;
            for (int col = 0; col < d->nAA; col++)
                for (int k = d->colptr[col] - 1; k < d->colptr[col + 1] - 1; k++)
                    for (int c = 0; c < d->ncomp; c++)
                        o[(size_t)(d->rowval[k] - 1) * d->ncomp + c] += nz[(size_t)k * d->ncomp + c] * dAAw[col];
            free(dAw); free(dAAw);
        }
    OMP_ENV_LOOP_END
}

void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
