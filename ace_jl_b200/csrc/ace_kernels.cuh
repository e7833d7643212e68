// ace_kernels.cuh -- the sm_100a kernels of the ACE evaluation path.
//
// Data flow for energies + forces:
//
//   k_pool     one thread per neighbour evaluates R_n and Y_l^m (m >= 0) into shared-memory staging;
//              the same CTA then pools  A_{q n l m} = sum_j R_n(r_j) Y_l^m(r_j)  with one thread per
//              (environment, (q,l,m)-column) holding the radial prefix in registers.  No atomics: the
//              reduction over neighbours is a serial loop over the staged tile.
//              [replaces evaluate(basis1p, cfg), src/product_1pbasis.jl:123-134]
//   k_adjoint  one LANE per environment, 32 environments per warp.  A lives in shared memory as
//              [slot][lane] (conflict-free), the adjoint trees are walked with warp-uniform control
//              flow, every lane accumulates dE/dA_a for its own environment in registers.
//              [replaces the AA loop of evaluate (src/evaluator.jl:137-143) and stage 2 of
//               _rrule_evaluate (src/evaluator.jl:180-185)]
//   k_forces   one thread per neighbour recomputes R_n, dR_n and walks Y_l^m / grad Y_l^m column by column,
//              contracting with the environment's dE/dA on the fly; the |A| x J x 3 complex matrix dA
//              that the reference materialises (src/product_1pbasis.jl:234-244) never exists.
//              [replaces stages 1 and 3 of _rrule_evaluate, src/evaluator.jl:169, 190-192]
//
// The intermediates A and dE/dA cross between kernels through a workspace laid out [slot][environment]
// so that k_adjoint's per-lane accesses are coalesced 512-byte rows.
#pragma once

#include "ace_math.cuh"

namespace aceb200 {

struct c2 { double x, y; };   // complex value, 16 bytes

ACE_HD inline c2 cmul(c2 a, c2 b) { return c2{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

// flip sign bits with integer ops (keeps the FP64 pipe free): used for conj / (-1)^m of canonical slots
ACE_HD inline double flip_sign(double v, unsigned long long mask)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double(__double_as_longlong(v) ^ (long long)mask);
#else
    union { double d; unsigned long long u; } w; w.d = v; w.u ^= mask; return w.d;
#endif
}

// A_a from its canonical slot value: code = slot*4 + neg + 2*odd  (ace_tables.h: make_code)
//   m >= 0 : A ;   m < 0, m even : conj(A) ;   m < 0, m odd : -conj(A)
ACE_HD inline c2 decode_A(c2 v, int code)
{
    const unsigned long long SB = 0x8000000000000000ull;
    unsigned long long mx = (code & 2) ? SB : 0ull;                 // real part flips iff odd negative m
    unsigned long long my = ((code & 1) && !(code & 2)) ? SB : 0ull; // imag part flips iff even negative m
    return c2{flip_sign(v.x, mx), flip_sign(v.y, my)};
}

// ------------------------------------------------------------------------------------------------
// device views of the tables
// ------------------------------------------------------------------------------------------------
struct ColumnsDev {
    int ncols, nS, nPused, nQ;
    const int* q; const int* l; const int* m; const int* cnt; const int* base; const int* ip;
    const int* colmap;      // [nQ][nPused]
};

// Adjoint list of one correlation order: records of `stride` bytes, sorted by target.
//   record = 4 x uint16 A-codes of the other factors (unused ones 0), then the weights:
//            Ppad doubles (real weights) or Ppad (re, im) pairs (complex weights)
struct ListDev {
    const int* ptr;               // [nA+1]
    const unsigned char* rec;
    int stride;
};

struct BatchDev {
    long long nenv;          // environments in this chunk
    const long long* off;    // [nenv+1] absolute neighbour offsets (off[0] = first neighbour of the chunk)
    const double* R;         // neighbour positions, indexed by ABSOLUTE neighbour index minus jbase
    const int* species;      // same indexing, or null
    long long jbase;         // absolute index of R[0]
};

// ------------------------------------------------------------------------------------------------
// k_pool: A_{slot}[env] for a chunk of environments
// ------------------------------------------------------------------------------------------------
struct PoolParams {
    RadialParams rp;
    AlpParams ap;
    ColumnsDev C;
    BatchDev B;
    c2* Ac;                 // [nS][ldA]
    long long ldA;
    int* errflag;           // set to ACEB200_EEMPTY / ACEB200_ECATEGORY on bad input
    int TE;                 // environments per CTA
    int SK;                 // staging row stride in doubles (odd)
    int colpass_size;       // columns handled per blockIdx.y
};

constexpr int kPoolThreads = 128;

template <int NMAX>
__global__ void __launch_bounds__(kPoolThreads) k_pool(const PoolParams p)
{
    ACE_DYN_SMEM(double, smem);
    double* S = smem;                                      // [kPoolThreads][SK]
    int* sq = reinterpret_cast<int*>(S + (size_t)kPoolThreads * p.SK);   // [kPoolThreads] species of staged neighbour
    const int tid = threadIdx.x;
    const int N = p.rp.N;
    const long long e0 = (long long)blockIdx.x * p.TE;
    if (e0 >= p.B.nenv) return;
    const int ne = (int)((p.B.nenv - e0) < p.TE ? (p.B.nenv - e0) : p.TE);
    const int col0 = blockIdx.y * p.colpass_size;
    const int ncol_here = (p.C.ncols - col0) < p.colpass_size ? (p.C.ncols - col0) : p.colpass_size;

    // work item of this thread in the pooling phase: (local environment, column)
    const int el = tid / ncol_here;
    const bool has_item = el < ne;
    int cq = 0, ccnt = 0, cbase = 0, cip = 0;
    long long jlo = 0, jhi = 0;
    if (has_item) {
        int col = col0 + tid % ncol_here;
        cq = p.C.q[col]; ccnt = p.C.cnt[col]; cbase = p.C.base[col]; cip = p.C.ip[col];
        jlo = p.B.off[e0 + el]; jhi = p.B.off[e0 + el + 1];
        if (jhi <= jlo && blockIdx.y == 0 && tid % ncol_here == 0) atomicMax(p.errflag, 5);   // EEMPTY
    }
    c2 acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = c2{0.0, 0.0};

    const long long jbeg = p.B.off[e0], jend = p.B.off[e0 + ne];
    for (long long c0 = jbeg; c0 < jend; c0 += kPoolThreads) {
        // ---- phase a: one thread per neighbour of this tile
        const long long j = c0 + tid;
        if (j < jend) {
            const double* r = p.B.R + 3 * (j - p.B.jbase);
            const double x = r[0], y = r[1], z = r[2];
            int q = 0;
            if (p.B.species) {
                q = p.B.species[j - p.B.jbase] - 1;
                if (q < 0 || q >= p.C.nQ) { atomicMax(p.errflag, 6); q = 0; }   // ECATEGORY
            }
            sq[tid] = q;
            const Spher sp = cart2spher(x, y, z);
            double Rn[NMAX];
            radial_e<NMAX>(p.rp, sp.r, Rn);
            double* row = S + (size_t)tid * p.SK;
#pragma unroll
            for (int n = 0; n < NMAX; ++n) if (n < N) row[n] = Rn[n];
            for_each_lm(p.ap, sp, [&](int l, int m, double Pv, double epr, double epi) {
                const int ip = index_p(l, m);
                row[N + 2 * ip] = epr * Pv;
                row[N + 2 * ip + 1] = epi * Pv;
            });
        }
        __syncthreads();
        // ---- phase b: pool this tile's neighbours into the thread's column
        if (has_item) {
            long long a = jlo > c0 ? jlo : c0;
            long long b = jhi < c0 + kPoolThreads ? jhi : c0 + kPoolThreads;
            for (long long jj = a; jj < b; ++jj) {
                const int t = (int)(jj - c0);
                if (sq[t] != cq) continue;
                const double* row = S + (size_t)t * p.SK;
                const double yr = row[N + 2 * cip], yi = row[N + 2 * cip + 1];
#pragma unroll
                for (int n = 0; n < NMAX; ++n)
                    if (n < ccnt) { const double rn = row[n]; acc[n].x += rn * yr; acc[n].y += rn * yi; }
            }
        }
        __syncthreads();
    }
    if (has_item) {
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
            if (n < ccnt) p.Ac[(size_t)(cbase + n) * p.ldA + (e0 + el)] = acc[n];
    }
}

// ------------------------------------------------------------------------------------------------
// k_adjoint: per environment  E = sum_AA c~ prod A   and   D~_slot = dE/dA folded onto m >= 0
// ------------------------------------------------------------------------------------------------
struct AdjointParams {
    int nS, nA, maxord, P, Ppad, has_const, want_D;
    const int* slot_pos; const int* slot_neg;    // [nS] target index or -1
    const int* code;                             // [nA] A-code of each target
    const double* w1;                            // [nA][Ppad](x2): order-1 weights (0 if no such AA)
    const double* w0;                            // [Ppad](x2): the constant
    ListDev list[kMaxOrdDev + 1];                // index nu = 2..maxord
    const c2* Ac; long long ldA;                 // [nS][ldA]
    c2* Dt;                                      // [nS][P][ldA]
    double* E;                                   // [nenv][P] (real part)
    long long nenv;
};

__device__ __forceinline__ c2 fetch_A(const c2* As, int lane, unsigned code)
{
    return decode_A(As[(code >> 2) * 32 + lane], (int)code);
}

// out[p] += sum over the leaves of target a:  w[leaf][p] * prod_{k < NU-1} A[code_k]
// The loop body has no loop-carried dependence except the accumulators, the record addresses do not
// depend on data, and control flow is warp-uniform: unrolling lets the table loads and the shared-memory
// gathers of several leaves be in flight together.
template <int NU, int PB, bool CW>
__device__ __forceinline__ void list_eval(const ListDev& T, int a, const c2* As, int lane, int pb, c2 (&out)[PB])
{
    const int i0 = __ldg(T.ptr + a), i1 = __ldg(T.ptr + a + 1);
    const unsigned char* r = T.rec + (size_t)i0 * T.stride;
    const int stride = T.stride;
#pragma unroll 4
    for (int i = i0; i < i1; ++i, r += stride) {
        c2 prod;
        if (PB == 1 && !CW) {
            // 16-byte record: one 128-bit uniform load brings the codes and the weight
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(r));
            prod = fetch_A(As, lane, q.x & 0xffffu);
            if (NU >= 3) prod = cmul(prod, fetch_A(As, lane, q.x >> 16));
            if (NU >= 4) prod = cmul(prod, fetch_A(As, lane, q.y & 0xffffu));
            if (NU >= 5) prod = cmul(prod, fetch_A(As, lane, q.y >> 16));
            const double w = __hiloint2double((int)q.w, (int)q.z);
            out[0].x += w * prod.x;
            out[0].y += w * prod.y;
        } else {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(r));
            prod = fetch_A(As, lane, q.x & 0xffffu);
            if (NU >= 3) prod = cmul(prod, fetch_A(As, lane, q.x >> 16));
            if (NU >= 4) prod = cmul(prod, fetch_A(As, lane, q.y & 0xffffu));
            if (NU >= 5) prod = cmul(prod, fetch_A(As, lane, q.y >> 16));
            const double* w = reinterpret_cast<const double*>(r + 8) + (size_t)pb * (CW ? 2 : 1);
#pragma unroll
            for (int p = 0; p < PB; ++p) {
                if (CW) {
                    const double wr = __ldg(w + 2 * p), wi = __ldg(w + 2 * p + 1);
                    out[p].x += wr * prod.x - wi * prod.y;
                    out[p].y += wr * prod.y + wi * prod.x;
                } else {
                    const double wr = __ldg(w + p);
                    out[p].x += wr * prod.x;
                    out[p].y += wr * prod.y;
                }
            }
        }
    }
}

// sum over orders of dE/dA_a; also accumulates the energy  Re(A_a S_nu) / nu  (Euler: sum_a A_a dF_nu/dA_a = nu F_nu)
template <int PB, bool CW>
__device__ __forceinline__ void target_eval(const AdjointParams& p, int a, const c2* As, int lane, int pb, c2 (&S)[PB], double (&E)[PB])
{
    const c2 Aa = fetch_A(As, lane, (unsigned)__ldg(p.code + a));
    // order 1: dE/dA_a = c~ ; energy Re(A_a c~)
    {
        const double* w = p.w1 + ((size_t)a * p.Ppad + pb) * (CW ? 2 : 1);
#pragma unroll
        for (int q = 0; q < PB; ++q) {
            const double wr = CW ? __ldg(w + 2 * q) : __ldg(w + q);
            const double wi = CW ? __ldg(w + 2 * q + 1) : 0.0;
            S[q] = c2{wr, wi};
            E[q] += Aa.x * wr - Aa.y * wi;
        }
    }
#define ACE_ORDER(NU)                                                                              \
    if (p.maxord >= NU) {                                                                          \
        c2 s[PB];                                                                                  \
        _Pragma("unroll") for (int q = 0; q < PB; ++q) s[q] = c2{0.0, 0.0};                        \
        list_eval<NU, PB, CW>(p.list[NU], a, As, lane, pb, s);                                     \
        _Pragma("unroll") for (int q = 0; q < PB; ++q) {                                           \
            S[q].x += s[q].x; S[q].y += s[q].y;                                                    \
            E[q] += (Aa.x * s[q].x - Aa.y * s[q].y) * (1.0 / NU);                                  \
        }                                                                                          \
    }
    ACE_ORDER(2) ACE_ORDER(3) ACE_ORDER(4) ACE_ORDER(5)
#undef ACE_ORDER
}

template <int PB, bool CW>
__global__ void __launch_bounds__(32) k_adjoint(const AdjointParams p)
{
    ACE_DYN_SMEM(c2, As);   // [nS][32]
    const int lane = threadIdx.x;
    const long long ntiles = (p.nenv + 31) / 32;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long e = tile * 32 + lane;      // the workspace is padded to a multiple of 32 columns
        for (int s = 0; s < p.nS; ++s) As[s * 32 + lane] = p.Ac[(size_t)s * p.ldA + e];
        __syncwarp();
        for (int pb = 0; pb < p.P; pb += PB) {
            double E[PB];
#pragma unroll
            for (int q = 0; q < PB; ++q) E[q] = p.has_const ? __ldg(p.w0 + (size_t)(pb + q) * (CW ? 2 : 1)) : 0.0;
            for (int s = 0; s < p.nS; ++s) {
                c2 D[PB];
#pragma unroll
                for (int q = 0; q < PB; ++q) D[q] = c2{0.0, 0.0};
                const int ap = __ldg(p.slot_pos + s), an = __ldg(p.slot_neg + s);
                if (ap >= 0) {
                    c2 S[PB];
                    target_eval<PB, CW>(p, ap, As, lane, pb, S, E);
#pragma unroll
                    for (int q = 0; q < PB; ++q) { D[q].x += S[q].x; D[q].y += S[q].y; }
                }
                if (an >= 0) {
                    // Re(D- grad(phi_-m)) = Re((-1)^m conj(D-) grad(phi_m)): fold onto the m > 0 slot
                    c2 S[PB];
                    target_eval<PB, CW>(p, an, As, lane, pb, S, E);
                    const double sg = (__ldg(p.code + an) & 2) ? -1.0 : 1.0;
#pragma unroll
                    for (int q = 0; q < PB; ++q) { D[q].x += sg * S[q].x; D[q].y -= sg * S[q].y; }
                }
                if (p.want_D && e < p.nenv) {
#pragma unroll
                    for (int q = 0; q < PB; ++q)
                        if (pb + q < p.P) p.Dt[((size_t)s * p.P + pb + q) * p.ldA + e] = D[q];
                }
            }
            if (e < p.nenv) {
#pragma unroll
                for (int q = 0; q < PB; ++q)
                    if (pb + q < p.P) p.E[(size_t)e * p.P + pb + q] = E[q];
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// k_adjoint_stream: the single-channel, real-weight fast path (energies + forces of an invariant model)
// ------------------------------------------------------------------------------------------------
// The adjoint lists of all targets and orders are flattened by the host into ONE stream of 16-byte
// records, in exactly the order the kernel consumes them, grouped in blocks of 4 leaves of the same
// (target, order):
//     record = { u16 c0, u16 c1, u16 c2, u16 ctl, f64 w }        leaf value = w * A[c0] * A[c1] (* A[c2])
// unused factors point at an extra slot that holds 1.  The ctl fields of a block's 4 records carry its
// header: ctl0 = flags | order, ctl1 = A-code of the target, ctl2 = target index.
// Because the stream is read strictly sequentially and identically by every lane, the warp fetches it
// cooperatively -- 32 records (512 contiguous bytes) per coalesced load, two chunks ahead of use -- into a
// small shared-memory ring, and reads records back as broadcasts.  Table latency is thereby hidden and
// the leaf loop is branch-free; control (end of order segment / target / slot) is a warp-uniform branch
// taken once per few dozen leaves.
constexpr unsigned kSegEnd = 8, kTgtEnd = 16, kTgtNeg = 32, kTgtOdd = 64, kSlotEnd = 128;

struct StreamParams {
    int nS, has_const, want_D, nchunks;      // nchunks: stream length in chunks of 8 blocks = 32 records
    const uint4* stream;
    const double* w1;                        // [nA] order-1 weights
    double w0;
    const c2* Ac; long long ldA;
    c2* Dt;                                  // [nS][ldA]
    double* E;                               // [nenv]
    long long nenv;
};

template <int NF>
__global__ void __launch_bounds__(32) k_adjoint_stream(const StreamParams p)
{
    ACE_DYN_SMEM(c2, As);                                       // [nS + 1][32]; slot nS holds 1
    uint4* ring = reinterpret_cast<uint4*>(As + (size_t)(p.nS + 1) * 32);   // [3][32]
    const int lane = threadIdx.x;
    const long long ntiles = (p.nenv + 31) / 32;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long e = tile * 32 + lane;
        for (int s = 0; s < p.nS; ++s) As[s * 32 + lane] = p.Ac[(size_t)s * p.ldA + e];
        As[p.nS * 32 + lane] = c2{1.0, 0.0};
        ring[lane] = __ldg(p.stream + lane);
        if (p.nchunks > 1) ring[32 + lane] = __ldg(p.stream + 32 + lane);
        __syncwarp();
        double E = p.has_const ? p.w0 : 0.0;
        c2 D = c2{0.0, 0.0}, S = c2{0.0, 0.0}, acc0 = c2{0.0, 0.0}, acc1 = c2{0.0, 0.0};
        int slot = 0;
        for (int ch = 0; ch < p.nchunks; ++ch) {
            const bool havepre = ch + 2 < p.nchunks;
            uint4 pre = uint4{0u, 0u, 0u, 0u};
            if (havepre) pre = __ldg(p.stream + (size_t)(ch + 2) * 32 + lane);
            const uint4* rb = ring + (ch % 3) * 32;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                uint4 q[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) q[k] = rb[4 * b + k];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    c2 prod = cmul(fetch_A(As, lane, q[k].x & 0xffffu), fetch_A(As, lane, q[k].x >> 16));
                    if (NF >= 3) prod = cmul(prod, fetch_A(As, lane, q[k].y & 0xffffu));
                    const double w = __hiloint2double((int)q[k].w, (int)q[k].z);
                    if (k & 1) { acc1.x += w * prod.x; acc1.y += w * prod.y; }
                    else { acc0.x += w * prod.x; acc0.y += w * prod.y; }
                }
                const unsigned flags = q[0].y >> 16;
                if (flags & 0xf8u) {
                    const c2 Aa = fetch_A(As, lane, q[1].y >> 16);
                    if (flags & kSegEnd) {
                        const c2 sg = c2{acc0.x + acc1.x, acc0.y + acc1.y};
                        acc0 = c2{0.0, 0.0}; acc1 = c2{0.0, 0.0};
                        const unsigned nu = flags & 7u;
                        const double inv = nu == 2 ? 0.5 : (nu == 3 ? (1.0 / 3.0) : 0.25);
                        S.x += sg.x; S.y += sg.y;
                        E += (Aa.x * sg.x - Aa.y * sg.y) * inv;    // Euler: sum_a A_a dF_nu/dA_a = nu F_nu
                    }
                    if (flags & kTgtEnd) {
                        const double w = __ldg(p.w1 + (q[2].y >> 16));
                        S.x += w;
                        E += Aa.x * w;
                        if (flags & kTgtNeg) {
                            // Re(D- grad(phi_-m)) = Re((-1)^m conj(D-) grad(phi_m)): fold onto the m > 0 slot
                            const double sgn = (flags & kTgtOdd) ? -1.0 : 1.0;
                            D.x += sgn * S.x; D.y -= sgn * S.y;
                        } else { D.x += S.x; D.y += S.y; }
                        S = c2{0.0, 0.0};
                    }
                    if (flags & kSlotEnd) {
                        if (p.want_D && e < p.nenv) p.Dt[(size_t)slot * p.ldA + e] = D;
                        D = c2{0.0, 0.0};
                        ++slot;
                    }
                }
            }
            if (havepre) ring[((ch + 2) % 3) * 32 + lane] = pre;
            __syncwarp();
        }
        if (e < p.nenv) p.E[e] = E;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// k_forces: one thread per neighbour
// ------------------------------------------------------------------------------------------------
struct ForceParams {
    RadialParams rp;
    AlpParams ap;
    ColumnsDev C;
    BatchDev B;
    const c2* Dt; long long ldA;
    int P, nprop, ncomp;
    int TE;                  // environments per CTA
    double* G;               // [neighbour][nprop][3][ncomp], chunk-relative
};

constexpr int kForceThreads = 128;

// A CTA owns TE consecutive environments: it stages their folded adjoints D~ in shared memory
// ([slot][channel][local env]), then runs one thread per neighbour of those environments.
template <int NMAX, int PB>
__global__ void __launch_bounds__(kForceThreads) k_forces(const ForceParams p)
{
    ACE_DYN_SMEM(c2, Ds);   // [nS][PB][TE]
    const int tid = threadIdx.x;
    const int TE = p.TE;
    const long long e0 = (long long)blockIdx.x * TE;
    if (e0 >= p.B.nenv) return;
    const int ne = (int)((p.B.nenv - e0) < TE ? (p.B.nenv - e0) : TE);
    const long long jbeg = p.B.off[e0], jend = p.B.off[e0 + ne];
    const int nS = p.C.nS;

    for (int pb = 0; pb < p.P; pb += PB) {
        __syncthreads();
        for (int idx = tid; idx < nS * PB * ne; idx += kForceThreads) {
            const int el = idx % ne, sc = idx / ne, c = sc % PB, s = sc / PB;
            Ds[(size_t)sc * TE + el] = (pb + c < p.P) ? p.Dt[((size_t)s * p.P + pb + c) * p.ldA + e0 + el] : c2{0.0, 0.0};
        }
        __syncthreads();
        for (long long jabs = jbeg + tid; jabs < jend; jabs += kForceThreads) {
            int el = 0;
            while (el + 1 < ne && p.B.off[e0 + el + 1] <= jabs) ++el;
            const double* r = p.B.R + 3 * (jabs - p.B.jbase);
            const double x = r[0], y = r[1], z = r[2];
            int q = 0;
            if (p.B.species) { q = p.B.species[jabs - p.B.jbase] - 1; if (q < 0 || q >= p.C.nQ) q = 0; }
            const Spher sp = cart2spher(x, y, z);
            double Rn[NMAX], dRn[NMAX];
            radial_ed<NMAX>(p.rp, sp.r, Rn, dRn);
            const int* cmap = p.C.colmap + (size_t)q * p.C.nPused;
            double S0[PB], S1[PB], S2[PB];
#pragma unroll
            for (int c = 0; c < PB; ++c) { S0[c] = 0.0; S1[c] = 0.0; S2[c] = 0.0; }
            for_each_lm_ed(p.ap, sp, [&](int l, int m, double Pt, double dP, double epr, double epi) {
                const int col = __ldg(cmap + index_p(l, m));
                if (col < 0) return;
                const int cnt = __ldg(p.C.cnt + col), base = __ldg(p.C.base + col);
                const double f0 = (m == 0) ? Pt : Pt * sp.sth;    // |Y| factor:  Y = ep * f0
                const double f1 = (double)m * Pt;
#pragma unroll
                for (int c = 0; c < PB; ++c) {
                    double ur = 0.0, ui = 0.0, vr = 0.0, vi = 0.0;
                    const c2* D = Ds + ((size_t)base * PB + c) * TE + el;
#pragma unroll
                    for (int n = 0; n < NMAX; ++n) {
                        if (n < cnt) {
                            const c2 d = D[(size_t)n * PB * TE];
                            ur += d.x * Rn[n]; ui += d.y * Rn[n];
                            vr += d.x * dRn[n]; vi += d.y * dRn[n];
                        }
                    }
                    // z = u * ep ;  Re(v * ep)
                    const double zr = ur * epr - ui * epi, zi = ur * epi + ui * epr;
                    const double ve = vr * epr - vi * epi;
                    S0[c] += f0 * ve;          // radial:   rhat * Re(v Y)
                    S1[c] += f1 * zi;          // azimuth:  m Pt Im(u ep)
                    S2[c] += dP * zr;          // polar:    dP Re(u ep)
                }
            });
            // g = rhat S0 + (1/r) [ sphi S1 + cphi cth S2,  -cphi S1 + sphi cth S2,  -sth S2 ]
            const double rx = x * sp.rinv, ry = y * sp.rinv, rz = z * sp.rinv;
            const long long jl = jabs - p.B.off[0];
#pragma unroll
            for (int c = 0; c < PB; ++c) {
                if (pb + c >= p.P) break;
                const double gx = rx * S0[c] + sp.rinv * (sp.sphi * S1[c] + sp.cphi * sp.cth * S2[c]);
                const double gy = ry * S0[c] + sp.rinv * (-sp.cphi * S1[c] + sp.sphi * sp.cth * S2[c]);
                const double gz = rz * S0[c] - sp.rinv * sp.sth * S2[c];
                const int ch = pb + c, prop = ch / p.ncomp, comp = ch % p.ncomp;
                double* g = p.G + (((size_t)jl * p.nprop + prop) * 3) * p.ncomp + comp;
                g[0] = gx; g[p.ncomp] = gy; g[2 * p.ncomp] = gz;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// basis-value kernels (evaluate on Product1pBasis / PIBasis / SymmetricBasis)
// ------------------------------------------------------------------------------------------------

// canonical slots -> the reference's A vector: A[e][iA] (src/product_1pbasis.jl:123-134)
__global__ void k_expand_A(long long nenv, int nA, const int* code, const c2* Ac, long long ldA, c2* A)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nenv * nA) return;
    const long long e = t / nA;
    const int a = (int)(t % nA);
    const int cd = __ldg(code + a);
    A[t] = decode_A(Ac[(size_t)(cd >> 2) * ldA + e], cd);
}

// AA[e][i] = real?(prod_t A[e][spec[i][t]])  (src/pibasis.jl:265-275)
__global__ void k_AA(long long nenv, int nA, int nAA, int maxord, const int* orders, const int* spec,
                     const c2* A, int pireal, double* AA)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nenv * nAA) return;
    const long long e = t / nAA;
    const int i = (int)(t % nAA);
    const c2* Ae = A + (size_t)e * nA;
    c2 aa = c2{1.0, 0.0};
    const int o = __ldg(orders + i);
    for (int k = 0; k < o; ++k) aa = cmul(aa, Ae[__ldg(spec + (size_t)i * maxord + k)]);
    if (pireal) AA[t] = aa.x;
    else { AA[2 * t] = aa.x; AA[2 * t + 1] = aa.y; }
}

// B[e][row][c] = real?(sum_k A2B[row,k][c] * AA[e][col_k])  (src/symmbasis.jl:248-264, 312-316), row-parallel CSR
__global__ void k_B(long long nenv, int nB, int nAA, int ncomp, const int* ptr, const int* col, const c2* val,
                    const double* AA, int pireal, int symreal, double* B)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nenv * nB * ncomp) return;
    const int c = (int)(t % ncomp);
    const int row = (int)((t / ncomp) % nB);
    const long long e = t / ((long long)ncomp * nB);
    double br = 0.0, bi = 0.0;
    for (int k = __ldg(ptr + row); k < __ldg(ptr + row + 1); ++k) {
        const c2 v = val[(size_t)k * ncomp + c];
        const size_t ia = (size_t)e * nAA + __ldg(col + k);
        const double ar = pireal ? AA[ia] : AA[2 * ia], ai = pireal ? 0.0 : AA[2 * ia + 1];
        br += v.x * ar - v.y * ai;
        bi += v.x * ai + v.y * ar;
    }
    if (symreal) B[t] = br;
    else { B[2 * t] = br; B[2 * t + 1] = bi; }
}

// ------------------------------------------------------------------------------------------------
// Jacobian kernels (evaluate_d / evaluate_ed)
// ------------------------------------------------------------------------------------------------
struct dAParams {
    RadialParams rp;
    AlpParams ap;
    ColumnsDev C;
    BatchDev B;
    const int* slot_pos; const int* slot_neg;
    int nA;
    c2* dA;                  // [neighbour][nA][3], chunk-relative
    long long nJ;
};

// dA[j][iA][:] = grad phi_iA(r_j) (src/product_1pbasis.jl:169-221), one thread per neighbour
template <int NMAX>
__global__ void __launch_bounds__(128) k_dA(const dAParams p)
{
    const long long jl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (jl >= p.nJ) return;
    const long long jabs = p.B.off[0] + jl;
    const double* r = p.B.R + 3 * (jabs - p.B.jbase);
    const double x = r[0], y = r[1], z = r[2];
    int q = 0;
    if (p.B.species) { q = p.B.species[jabs - p.B.jbase] - 1; if (q < 0 || q >= p.C.nQ) q = 0; }
    const Spher sp = cart2spher(x, y, z);
    double Rn[NMAX], dRn[NMAX];
    radial_ed<NMAX>(p.rp, sp.r, Rn, dRn);
    c2* out = p.dA + (size_t)jl * p.nA * 3;
    // functions of other species are identically zero for this neighbour
    for (int a = 0; a < p.nA * 3; ++a) out[a] = c2{0.0, 0.0};
    const int* cmap = p.C.colmap + (size_t)q * p.C.nPused;
    const double rx = x * sp.rinv, ry = y * sp.rinv, rz = z * sp.rinv;
    for_each_lm_ed(p.ap, sp, [&](int l, int m, double Pt, double dP, double epr, double epi) {
        const int col = __ldg(cmap + index_p(l, m));
        if (col < 0) return;
        const int cnt = __ldg(p.C.cnt + col), base = __ldg(p.C.base + col);
        const double f0 = (m == 0) ? Pt : Pt * sp.sth;
        const c2 Y = c2{epr * f0, epi * f0};
        // grad Y = dspher_to_dcart(S, i m ep Pt, ep dP)  (sphericalharmonics.jl:60-65, 429-438)
        const c2 F1 = c2{-(double)m * epi * Pt, (double)m * epr * Pt};
        const c2 F2 = c2{epr * dP, epi * dP};
        c2 gY[3];
        gY[0] = c2{(-sp.sphi * F1.x + sp.cphi * sp.cth * F2.x) * sp.rinv, (-sp.sphi * F1.y + sp.cphi * sp.cth * F2.y) * sp.rinv};
        gY[1] = c2{(sp.cphi * F1.x + sp.sphi * sp.cth * F2.x) * sp.rinv, (sp.cphi * F1.y + sp.sphi * sp.cth * F2.y) * sp.rinv};
        gY[2] = c2{(-sp.sth * F2.x) * sp.rinv, (-sp.sth * F2.y) * sp.rinv};
        const double rh[3] = {rx, ry, rz};
#pragma unroll
        for (int n = 0; n < NMAX; ++n) {
            if (n < cnt) {
                const int apos = __ldg(p.slot_pos + base + n), aneg = __ldg(p.slot_neg + base + n);
                c2 g[3];
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    g[k] = c2{dRn[n] * rh[k] * Y.x + Rn[n] * gY[k].x, dRn[n] * rh[k] * Y.y + Rn[n] * gY[k].y};
                if (apos >= 0) for (int k = 0; k < 3; ++k) out[(size_t)apos * 3 + k] = g[k];
                if (aneg >= 0) {
                    const double sg = (m & 1) ? -1.0 : 1.0;
                    for (int k = 0; k < 3; ++k) out[(size_t)aneg * 3 + k] = c2{sg * g[k].x, -sg * g[k].y};
                }
            }
        }
    });
}

// dAA[j][i][:] = real?(sum_t (prod_{s != t} A_{v_s}) dA[j][v_t][:])  (src/pibasis.jl:402-432)
constexpr int kMaxOrdDevK = 8;
__global__ void k_dAA(long long nenv, const long long* off, int nA, int nAA, int maxord, const int* orders, const int* spec,
                      const c2* A, const c2* dA, int pireal, double* dAA)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nenv * nAA) return;
    const long long e = t / nAA;
    const int i = (int)(t % nAA);
    const c2* Ae = A + (size_t)e * nA;
    const int o = __ldg(orders + i);
    int v[kMaxOrdDevK];
    c2 adj[kMaxOrdDevK];
    for (int k = 0; k < o; ++k) v[k] = __ldg(spec + (size_t)i * maxord + k);
    // adj[k] = prod_{s != k} A[v_s]: forward prefix products, then a backward sweep (src/pibasis.jl:362-390)
    c2 run = c2{1.0, 0.0};
    for (int k = 0; k < o; ++k) { adj[k] = run; run = cmul(run, Ae[v[k]]); }
    run = c2{1.0, 0.0};
    for (int k = o - 1; k >= 0; --k) { adj[k] = cmul(adj[k], run); run = cmul(run, Ae[v[k]]); }
    const int cs = pireal ? 1 : 2;
    const long long j0 = off[e] - off[0], j1 = off[e + 1] - off[0];
    for (long long j = j0; j < j1; ++j) {
        c2 g[3] = {c2{0, 0}, c2{0, 0}, c2{0, 0}};
        for (int k = 0; k < o; ++k) {
            const c2* da = dA + ((size_t)j * nA + v[k]) * 3;
            for (int d = 0; d < 3; ++d) { const c2 m = cmul(adj[k], da[d]); g[d].x += m.x; g[d].y += m.y; }
        }
        double* out = dAA + ((size_t)j * nAA + i) * 3 * cs;
        for (int d = 0; d < 3; ++d) {
            if (pireal) out[d] = g[d].x;
            else { out[2 * d] = g[d].x; out[2 * d + 1] = g[d].y; }
        }
    }
}

// dB[j][row][xyz][c] = real?(sum_k A2B[row,k][c] * dAA[j][col_k][xyz])  (src/symmbasis.jl:330-334, src/properties.jl:53-59)
__global__ void k_dB(long long nJ, int nB, int nAA, int ncomp, const int* ptr, const int* col, const c2* val,
                     const double* dAA, int pireal, int symreal, double* dB)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nJ * nB * 3) return;
    const int d = (int)(t % 3);
    const int row = (int)((t / 3) % nB);
    const long long j = t / ((long long)3 * nB);
    const int ca = pireal ? 1 : 2, cs = symreal ? 1 : 2;
    for (int c = 0; c < ncomp; ++c) {
        double br = 0.0, bi = 0.0;
        for (int k = __ldg(ptr + row); k < __ldg(ptr + row + 1); ++k) {
            const c2 v = val[(size_t)k * ncomp + c];
            const double* xa = dAA + (((size_t)j * nAA + __ldg(col + k)) * 3 + d) * ca;
            const double ar = xa[0], ai = pireal ? 0.0 : xa[1];
            br += v.x * ar - v.y * ai;
            bi += v.x * ai + v.y * ar;
        }
        double* o = dB + ((((size_t)j * nB + row) * 3 + d) * ncomp + c) * cs;
        o[0] = br;
        if (!symreal) o[1] = bi;
    }
}

}  // namespace aceb200
