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

#ifdef ACEB200_EMU
struct alignas(16) c2 { double x, y; };   // complex value, 16 bytes
#else
struct __align__(16) c2 { double x, y; };   // complex value: 16-byte aligned so that it moves as one 128-bit access
#endif

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
    const int* slot_n; const int* slot_ip; const int* slot_q;   // [nS] decode of a canonical slot
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
    const int* gate;         // the call's error flag: 1 = the offsets of a DEVICE batch failed k_check_offsets, so no
                             // kernel may index with them (every offset-consuming kernel returns at once)
};
#define ACE_GATE(B) do { if ((B).gate && *(B).gate == 1) return; } while (0)

// ------------------------------------------------------------------------------------------------
// k_pool: A_{slot}[env] for a chunk of environments
// ------------------------------------------------------------------------------------------------
// A CTA owns TE consecutive environments and works through them in sub-tiles of at most 128 neighbours:
// as many WHOLE environments as fit (or a 128-neighbour piece of one that does not).
//   phase a  one thread per neighbour: R_n -> SR[n][row], Y_l^m (m >= 0) -> SY[ip][row]   (registers -> smem)
//   phase b  one thread per (environment, slot) item, flat over the sub-tile: the reduction over the
//            environment's neighbours is a serial loop over its staged rows -- no atomics, deterministic.
// Staging is [function][row] with a row pitch of 129: phase a writes and phase b reads are both
// bank-conflict free, and consecutive neighbours are adjacent so the unrolled reduction uses immediate offsets.
struct PoolParams {
    RadialParams rp;
    AlpParams ap;
    ColumnsDev C;
    BatchDev B;
    c2* Ac;                 // [nS][ldA]
    long long ldA;
    int* errflag;           // set to ACEB200_EEMPTY / ACEB200_ECATEGORY on bad input
    int TE;                 // environments per CTA (<= kPoolTEmax)
    int nP;                 // harmonics staged per neighbour: sizeP(L)
    const int4* blk;        // [nblk] 2 x 2 register blocks of slots (two radial indices x two columns of one species):
    int nblk, nbp;          //   x = n0 | n1 << 8 | q << 16,  y = ip0 | ip1 << 16,
                            //   z = slot(n0, col0) | slot(n1, col0) << 16,  w = slot(n0, col1) | slot(n1, col1) << 16  (0xffff: none)
                            // nbp >= nblk: lanes reserved per environment in phase b (a power of two up to 32, or a multiple
                            // of 32), so that the lanes that share a shared-memory wavefront belong to one environment
};

constexpr int kPoolThreads = 128;
constexpr int kPoolBItems = 2;      // (environment, slot block) items a thread can own per sub-tile of k_pool
constexpr int kPoolItems = 4;       // (environment, slot) items a thread can own per sub-tile
constexpr int kPoolPitch = 129;     // staging row pitch (elements)
constexpr int kPoolTEmax = 16;

template <int NMAX, bool SPECIES, int WALK>
__global__ void __launch_bounds__(kPoolThreads) k_pool(const PoolParams p)
{
    ACE_DYN_SMEM(c2, smem);
    c2* SY = smem;                                                            // [nP][129]
    double* SR = reinterpret_cast<double*>(SY + (size_t)p.nP * kPoolPitch);    // [N][129]
    int* sq = reinterpret_cast<int*>(SR + (size_t)p.rp.N * kPoolPitch);        // [128] species of the staged neighbour
    int* joff = sq + kPoolThreads;                                             // [TE + 1] neighbour offsets relative to the CTA's first
    const int tid = threadIdx.x;
    const int N = p.rp.N, nblk = p.nblk, nbp = p.nbp;
    const long long e0 = (long long)blockIdx.x * p.TE;
    if (e0 >= p.B.nenv) return;
    ACE_GATE(p.B);
    const int ne = (int)((p.B.nenv - e0) < p.TE ? (p.B.nenv - e0) : p.TE);
    const long long jbeg = p.B.off[e0];
    if (tid <= ne) joff[tid] = (int)(p.B.off[e0 + tid] - jbeg);
    const double* Rb = p.B.R + 3 * (jbeg - p.B.jbase);
    const int* spb = SPECIES ? p.B.species + (jbeg - p.B.jbase) : nullptr;
    __syncthreads();

    c2 acc[kPoolBItems][4];
#pragma unroll
    for (int it = 0; it < kPoolBItems; ++it)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[it][k] = c2{0.0, 0.0};

    int e = 0, j0 = 0;
    while (e < ne) {
        // ---- choose the sub-tile [j0, j1): whole environments e..e2-1, or a piece of environment e
        const int jend_e = joff[e + 1];
        int e2 = e + 1, j1;
        bool done_e;                      // the environments of this sub-tile are complete after it
        if (j0 == joff[e] && jend_e - j0 <= kPoolThreads) {
            while (e2 < ne && joff[e2 + 1] - j0 <= kPoolThreads && (e2 + 1 - e) * nbp <= kPoolThreads * kPoolBItems) ++e2;
            j1 = joff[e2];
            done_e = true;
        } else {
            j1 = (j0 + kPoolThreads < jend_e) ? j0 + kPoolThreads : jend_e;
            done_e = (j1 == jend_e);
        }
        const int nrows = j1 - j0;

        // ---- phase a
        if (tid < nrows) {
            const int j = j0 + tid;
            const double x = Rb[3 * j], y = Rb[3 * j + 1], z = Rb[3 * j + 2];
            if (SPECIES) {
                int q = spb[j] - 1;
                if (q < 0 || q >= p.C.nQ) { atomicMax(p.errflag, 6); q = 0; }   // ECATEGORY (src/discrete1pbasis.jl:39)
                sq[tid] = q;
            }
            const Spher sp = cart2spher(x, y, z);
            double Rn[NMAX];
            radial_e<NMAX>(p.rp, sp.r, Rn);
#pragma unroll
            for (int n = 0; n < NMAX; ++n) if (n < N) SR[n * kPoolPitch + tid] = Rn[n];
            for_each_lm<WALK>(p.ap, sp, [&](int l, int m, double Pv, double epr, double epi) {
                SY[index_p(l, m) * kPoolPitch + tid] = c2{epr * Pv, epi * Pv};
            });
        }
        __syncthreads();

        // ---- phase b: one (environment, 2 x 2 slot block) item accumulates r_{n0,n1}[j] * y_{col0,col1}[j] over
        // the environment's staged rows: four shared-memory loads feed eight FP64 FMAs
        const int nitems = (e2 - e) * nbp;
#pragma unroll
        for (int it = 0; it < kPoolBItems; ++it) {
            const int idx = tid + it * kPoolThreads;
            const int el = idx / nbp, b = idx - el * nbp;
            if (idx < nitems && b < nblk) {
                const int4 d = __ldg(p.blk + b);
                int ra = joff[e + el] - j0, rb = joff[e + el + 1] - j0;
                if (ra == rb && b == 0) atomicMax(p.errflag, 5);          // EEMPTY (src/product_1pbasis.jl:124)
                if (ra < 0) ra = 0;
                if (rb > nrows) rb = nrows;
                const double* pr0 = SR + (d.x & 0xff) * kPoolPitch + ra;
                const double* pr1 = SR + ((d.x >> 8) & 0xff) * kPoolPitch + ra;
                const c2* py0 = SY + (d.y & 0xffff) * kPoolPitch + ra;
                const c2* py1 = SY + ((d.y >> 16) & 0xffff) * kPoolPitch + ra;
                c2 a00 = acc[it][0], a10 = acc[it][1], a01 = acc[it][2], a11 = acc[it][3];
                const int nr = rb - ra;
                if (SPECIES) {
                    const int q = (d.x >> 16) & 0xffff;
                    const int* ps = sq + ra;
                    for (int r = 0; r < nr; ++r) {
                        const bool on = ps[r] == q;
                        const double r0 = on ? pr0[r] : 0.0, r1 = on ? pr1[r] : 0.0;
                        const c2 y0 = py0[r], y1 = py1[r];
                        a00.x += r0 * y0.x; a00.y += r0 * y0.y; a10.x += r1 * y0.x; a10.y += r1 * y0.y;
                        a01.x += r0 * y1.x; a01.y += r0 * y1.y; a11.x += r1 * y1.x; a11.y += r1 * y1.y;
                    }
                } else {
                    int r = 0;
                    for (; r + 3 < nr; r += 4) {            // four rows per trip: half the pointer bumps and branches (-2.6 %)
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const double r0 = pr0[u], r1 = pr1[u];
                            const c2 y0 = py0[u], y1 = py1[u];
                            a00.x += r0 * y0.x; a00.y += r0 * y0.y; a10.x += r1 * y0.x; a10.y += r1 * y0.y;
                            a01.x += r0 * y1.x; a01.y += r0 * y1.y; a11.x += r1 * y1.x; a11.y += r1 * y1.y;
                        }
                        pr0 += 4; pr1 += 4; py0 += 4; py1 += 4;
                    }
                    for (; r + 1 < nr; r += 2) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const double r0 = pr0[u], r1 = pr1[u];
                            const c2 y0 = py0[u], y1 = py1[u];
                            a00.x += r0 * y0.x; a00.y += r0 * y0.y; a10.x += r1 * y0.x; a10.y += r1 * y0.y;
                            a01.x += r0 * y1.x; a01.y += r0 * y1.y; a11.x += r1 * y1.x; a11.y += r1 * y1.y;
                        }
                        pr0 += 2; pr1 += 2; py0 += 2; py1 += 2;
                    }
                    if (r < nr) {
                        const double r0 = pr0[0], r1 = pr1[0];
                        const c2 y0 = py0[0], y1 = py1[0];
                        a00.x += r0 * y0.x; a00.y += r0 * y0.y; a10.x += r1 * y0.x; a10.y += r1 * y0.y;
                        a01.x += r0 * y1.x; a01.y += r0 * y1.y; a11.x += r1 * y1.x; a11.y += r1 * y1.y;
                    }
                }
                if (done_e) {
                    c2* out = p.Ac + (e0 + e + el);
                    const int s00 = d.z & 0xffff, s10 = (d.z >> 16) & 0xffff, s01 = d.w & 0xffff, s11 = (d.w >> 16) & 0xffff;
                    if (s00 != 0xffff) out[(size_t)s00 * p.ldA] = a00;
                    if (s10 != 0xffff) out[(size_t)s10 * p.ldA] = a10;
                    if (s01 != 0xffff) out[(size_t)s01 * p.ldA] = a01;
                    if (s11 != 0xffff) out[(size_t)s11 * p.ldA] = a11;
                    a00 = a10 = a01 = a11 = c2{0.0, 0.0};
                }
                acc[it][0] = a00; acc[it][1] = a10; acc[it][2] = a01; acc[it][3] = a11;
            }
        }
        __syncthreads();
        j0 = j1;
        if (done_e) e = e2;
    }
}

// ------------------------------------------------------------------------------------------------
// FP64 tensor-core building block: D(8x8) += A(8x4, row-major) * B(4x8, column-major), mma.sync.m8n8k4.f64 (SASS DMMA).
// Fragments: lane i holds A[i / 4][i % 4], B[i % 4][i / 4] and C[i / 4][2 (i % 4) + {0, 1}].
// One operand pair (two 8-byte shared-memory loads per lane) feeds 256 FMAs of the warp, against 8 FMAs for the four
// loads of the register-blocked FMA formulation: the pooling / force contractions stop being shared-memory bound.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
#ifdef ACEB200_EMU
    emu::mma_m8n8k4(c0, c1, a, b);
#else
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
#endif
}

// ------------------------------------------------------------------------------------------------
// k_pool_mma: the pooling  A_{n, lm} = sum_j R_n(r_j) Y_lm(r_j)  of one environment as a small dense product
//     [n] x [j] . [j] x [(lm, re / im)]
// on the FP64 tensor cores.  Phase a is k_pool's (one thread per neighbour: R_n and Y_l^m, m >= 0, into shared-memory
// planes); phase b hands (environment, column tile) units to the four warps.  A column tile is four (l, m) columns
// = eight real columns of the B operand; the radial index runs over up to NT row tiles of eight, of which only those
// below the tile's longest column are computed (columns are sorted by length, so the staircase (n, l) pattern of a
// level-truncated one-particle basis wastes as little as an 8 x 8 tiling allows).  Each lane ends up with the complex
// A value of (n = 8 nt + lane / 4, column = lane % 4) -- exactly one canonical slot -- and stores it.
// The plane pitch is 4 mod 16 doubles: the A and B fragment loads (lane -> row lane / 4 | lane % 4 apart by one
// pitch) then touch every bank exactly twice, the minimum for 32 eight-byte loads.
// ------------------------------------------------------------------------------------------------
constexpr int kPoolMmaThreads = 256;     // k_pool_mma: 128 staged neighbours per sub-tile, two threads each in phase a, eight warps in phase b
constexpr int kMmaPitch = 132;           // doubles per staged plane: >= 128 + 3 (tail reads), = 4 mod 16
struct PoolTile { int ip[4], base[4], cnt[4], nnt, pad[3]; };   // four columns: harmonic index, first slot, length; row tiles

struct PoolMmaParams {
    RadialParams rp;
    AlpParams ap;
    BatchDev B;
    c2* Ac; long long ldA;
    int* errflag;
    int TE, nP;
    const PoolTile* tiles; int ntiles;
};

// acc[nt] += R-tile(nt)[8 x nr] . Y-tile[nr x 8] over the nr staged rows of one environment, NNT row tiles
template <int NNT, int NT>
__device__ __forceinline__ void pool_unit(const double* pa, const double* pb, const int (&rowA)[NT], int nr, int k4, double (&acc)[NT][2])
{
    const double* qa[NNT];
#pragma unroll
    for (int nt = 0; nt < NNT; ++nt) qa[nt] = pa + rowA[nt];
    // One accumulator chain per row tile, k-steps strictly in order: appending neighbours that contribute exact zeros
    // (beyond the cutoff, src/polynomials/orthpolys.jl:41-46) then leaves every bit of A unchanged, as in the reference's
    // sequential sum.  (The 26-cycle DMMA latency is hidden by the ten resident warps per SMSP, not by a second chain.)
    int k = 0;
#pragma unroll 1
    for (; k + 8 <= nr; k += 8) {
        const double b0 = pb[0], b1 = pb[4];
        double a0[NNT], a1[NNT];
#pragma unroll
        for (int nt = 0; nt < NNT; ++nt) { a0[nt] = qa[nt][0]; a1[nt] = qa[nt][4]; }
#pragma unroll
        for (int nt = 0; nt < NNT; ++nt) dmma(acc[nt][0], acc[nt][1], a0[nt], b0);
#pragma unroll
        for (int nt = 0; nt < NNT; ++nt) dmma(acc[nt][0], acc[nt][1], a1[nt], b1);
        pb += 8;
#pragma unroll
        for (int nt = 0; nt < NNT; ++nt) qa[nt] += 8;
    }
    if (k + 4 <= nr) {
        const double b0 = pb[0];
#pragma unroll
        for (int nt = 0; nt < NNT; ++nt) dmma(acc[nt][0], acc[nt][1], qa[nt][0], b0);
        pb += 4; k += 4;
#pragma unroll
        for (int nt = 0; nt < NNT; ++nt) qa[nt] += 4;
    }
    if (k < nr) {                                       // last, partial k-step: rows beyond the environment count as zero
        const bool ok = k + k4 < nr;
        const double b0 = ok ? pb[0] : 0.0;
#pragma unroll
        for (int nt = 0; nt < NNT; ++nt) dmma(acc[nt][0], acc[nt][1], ok ? qa[nt][0] : 0.0, b0);
    }
}

template <int NMAX, int WALK>
__global__ void __launch_bounds__(kPoolMmaThreads) k_pool_mma(const PoolMmaParams p)
{
    constexpr int NT = (NMAX + 7) / 8;
    constexpr int P = kMmaPitch;
    ACE_DYN_SMEM(double, smem);
    const int N = p.rp.N, nP = p.nP, ntiles = p.ntiles;
    double* SYr = smem;                                        // [nP][P]
    double* SYi = SYr + (size_t)nP * P;                        // [nP][P]
    double* SR = SYi + (size_t)nP * P;                         // [N][P]
    PoolTile* tl = reinterpret_cast<PoolTile*>(SR + (size_t)N * P);   // [ntiles]
    int* joff = reinterpret_cast<int*>(tl + ntiles);           // [TE + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long e0 = (long long)blockIdx.x * p.TE;
    if (e0 >= p.B.nenv) return;
    ACE_GATE(p.B);
    const int ne = (int)((p.B.nenv - e0) < p.TE ? (p.B.nenv - e0) : p.TE);
    const long long jbeg = p.B.off[e0];
    if (tid <= ne) joff[tid] = (int)(p.B.off[e0 + tid] - jbeg);
    for (int i = tid; i < ntiles * (int)(sizeof(PoolTile) / sizeof(int)); i += kPoolMmaThreads)
        reinterpret_cast<int*>(tl)[i] = __ldg(reinterpret_cast<const int*>(p.tiles) + i);
    const double* Rb = p.B.R + 3 * (jbeg - p.B.jbase);
    const int k4 = lane & 3, g = lane >> 2;
    int rowA[NT];                                         // plane offset of this lane's A-fragment row in each row tile
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { const int n = nt * 8 + g; rowA[nt] = (n < N ? n : N - 1) * P; }   // rows >= N: finite, never stored
    __syncthreads();

    int e = 0, j0 = 0;
    while (e < ne) {
        // ---- the sub-tile [j0, j1): whole environments e..e2-1, or a 128-neighbour piece of environment e
        const int jend_e = joff[e + 1];
        int e2 = e + 1, j1;
        bool done_e, first_piece = true;
        if (j0 == joff[e] && jend_e - j0 <= kPoolThreads) {
            while (e2 < ne && joff[e2 + 1] - j0 <= kPoolThreads) ++e2;
            j1 = joff[e2];
            done_e = true;
        } else {
            first_piece = (j0 == joff[e]);
            j1 = (j0 + kPoolThreads < jend_e) ? j0 + kPoolThreads : jend_e;
            done_e = (j1 == jend_e);
        }
        const int nrows = j1 - j0;

        // ---- phase a: two threads per neighbour (threads 0..127: R_n, threads 128..255: Y_l^m of neighbour tid % 128)
        {
            const int rowa = tid & (kPoolThreads - 1);
            if (rowa < nrows) {
                const int j = j0 + rowa;
                const double x = Rb[3 * j], y = Rb[3 * j + 1], z = Rb[3 * j + 2];
                if (tid < kPoolThreads) {
                    const double r2 = x * x + y * y + z * z;
#ifdef __CUDA_ARCH__
                    const double r = r2 * rsqrt(r2);
#else
                    const double r = r2 * (1.0 / sqrt(r2));
#endif
                    double Rn[NMAX];
                    radial_e<NMAX>(p.rp, r, Rn);
#pragma unroll
                    for (int n = 0; n < NMAX; ++n) if (n < N) SR[n * P + rowa] = Rn[n];
                } else {
                    const Spher sp = cart2spher(x, y, z);
                    for_each_lm<WALK>(p.ap, sp, [&](int l, int m, double Pv, double epr, double epi) {
                        SYr[index_p(l, m) * P + rowa] = epr * Pv;
                        SYi[index_p(l, m) * P + rowa] = epi * Pv;
                    });
                }
            }
        }
        __syncthreads();

        // ---- phase b: (environment, column tile) units, heaviest tiles first, dealt round-robin to the warps.
        // The k-loop is the hot spot and is issue-bound: operands come through per-operand running pointers with
        // immediate offsets (two k-steps per trip) and the number of row tiles is a compile-time constant per branch,
        // so that a trip is 2 (1 + NNT) LDS.64 + 2 NNT DMMA + a handful of integer instructions.
        const int nes = e2 - e;
        {
            int t = 0, el = warp;                          // unit u = t * nes + el, u = warp, warp + 4, ...
            while (el >= nes) { el -= nes; ++t; }
            while (t < ntiles) {
                const PoolTile& T = tl[t];
                int ra = joff[e + el] - j0, rb = joff[e + el + 1] - j0;
                if (ra == rb && t == 0 && lane == 0) atomicMax(p.errflag, 5);      // EEMPTY (src/product_1pbasis.jl:124)
                if (ra < 0) ra = 0;
                if (rb > nrows) rb = nrows;
                const int nr = rb - ra, nnt = T.nnt;
                const double* pb = ((g & 1) ? SYi : SYr) + T.ip[g >> 1] * P + ra + k4;
                const double* pa = SR + ra + k4;
                double acc[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = 0.0; acc[nt][1] = 0.0; }
                switch (nnt) {
                case 1: pool_unit<1, NT>(pa, pb, rowA, nr, k4, acc); break;
                case 2: if (NT >= 2) pool_unit<(NT >= 2 ? 2 : 1), NT>(pa, pb, rowA, nr, k4, acc); break;
                case 3: if (NT >= 3) pool_unit<(NT >= 3 ? 3 : 1), NT>(pa, pb, rowA, nr, k4, acc); break;
                case 4: if (NT >= 4) pool_unit<(NT >= 4 ? 4 : 1), NT>(pa, pb, rowA, nr, k4, acc); break;
                default: break;
                }
                const int cnt = T.cnt[k4];
                c2* out = p.Ac + (size_t)(T.base[k4] + g) * p.ldA + (e0 + e + el);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    if (nt < nnt && nt * 8 + g < cnt) {
                        c2* o = out + (size_t)(nt * 8) * p.ldA;
                        c2 v = c2{acc[nt][0], acc[nt][1]};
                        if (!first_piece) { const c2 w = *o; v.x += w.x; v.y += w.y; }   // a later piece of a > 128-neighbour environment
                        *o = v;
                    }
                }
                el += kPoolMmaThreads / 32;
                while (el >= nes) { el -= nes; ++t; }
            }
        }
        __syncthreads();
        j0 = j1;
        if (done_e) e = e2;
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
// k_adjoint_stream: energies and folded adjoints for PB output channels per pass
// ------------------------------------------------------------------------------------------------
// The adjoint lists of all targets and orders are flattened by the host into streams in exactly the order
// the kernel consumes them.  The kernel is bound by shared-memory -> register bandwidth (every byte a lane
// receives costs the same, broadcast or not), so the per-leaf stream is as small as it can be:
//   leaf block (4 leaves of one (target, order) segment) = { u32 code[4] (x2 for NF > 2), w[4][PB] }
//     code  = slot1 | flipIm<<15 | slot2<<16 | conj2<<31        code' = slot3 | conj3<<15 | slot4<<16 | conj4<<31
//     X     = A~[slot1] * cj(A~[slot2]) (* cj(A~[slot3]) * cj(A~[slot4])),  Im X negated if flipIm
//     value = w[p] * X   for each channel p (w real, or complex when CW)
//     A~ are the canonical (m >= 0) slots; unused factors point at an extra slot that holds 1; the conjugation
//     of the first operand and all (-1)^m signs are folded into flipIm / w by the host:
//       prod_k (s_k cj^{k_k} A~_k) = (prod s_k) cj^{k_1}( A~_1 prod_{k>1} cj^{k_k xor k_1} A~_k ).
//   Grouped form (one real channel, PB == 1 && !CW): within a segment the host greedily groups the leaves that
//   share a factor, makes that factor the first one, and marks each group's last leaf (kGroupEnd).  The kernel
//   accumulates  T += w * cj^{k_2}(A~[slot2]) (* ...)  per leaf (two FMAs instead of a complex product plus two
//   FMAs) and multiplies by the shared factor once per group:  S += cj^{k_1}(A~[slot1]) * T.  Bit 15 is then the
//   conjugation of the first factor alone, the other conj bits apply to their own factor, and w carries the
//   (-1)^m signs.  Order-2 leaves (one other factor) form a single group whose first factor is the slot of ones.
//   ctl[block]  (u32, global)  0 for most blocks; else flags | order | slot<<8: end of segment / target / slot
//   tinfo[k]    (global) consumed in order, one per non-zero ctl: target slot + masks, 1/order, order-1 weights
// The leaf blocks are read strictly sequentially and identically by every lane, so the warp fetches them
// cooperatively (coalesced 128-bit loads, two chunks ahead of use) into a 2-slot shared-memory ring and reads
// them back as broadcasts.  Table latency is hidden, the leaves of a block are independent, and control is a
// warp-uniform branch on ctl.  The energy falls out of the same pass by Euler's identity
// sum_a A_a dF_nu/dA_a = nu F_nu.
constexpr unsigned kSegEnd = 8, kTgtEnd = 16, kTgtNeg = 32, kTgtOdd = 64, kSlotEnd = 128;
constexpr unsigned kGroupEnd = 0x4000u;     // leaf code bit 14 (grouped form)
constexpr unsigned kSubEnd = 0x4000u;       // second code word bit 14 (two-level grouped form, NF == 3)
constexpr int kBlkLeaves = 4;               // leaves per block
constexpr int kStreamWarps = 12;            // most warps a CTA of k_adjoint_stream can have (StreamGeom::NW); each walks its own
                                            // sub-stream over the same A tile

template <int NF, int PB, bool CW>
struct StreamGeom {
    static constexpr int CS = CW ? 2 : 1;
    static constexpr int CWORDS = (NF == 2) ? 1 : 2;                     // code words per leaf
    static constexpr int QB = CWORDS + 2 * PB * CS;                      // uint4 per block (4 leaves)
    // warps per CTA: the single-channel real stream has small rings, so 12 warps share one A tile and two such CTAs still
    // fit on an SM (24 resident warps instead of 16); the multi-channel streams keep 8 (their rings are 2.5 KB per slot)
    static constexpr int NW = (PB == 1 && !CW) ? 12 : 8;
    static constexpr int KB = (QB <= 4) ? 16 : 8;                        // blocks per ring chunk
    static constexpr int QBP = ((QB + 3) / 4) * 4;                       // padded so that a chunk is a multiple of 32 uint4
    static constexpr int CH = KB * QBP;                                  // uint4 per chunk
    static constexpr int LPC = CH / 32;                                  // uint4 each lane loads per chunk
    static constexpr int TIQ = 1 + (1 + PB * CS + 1) / 2;                // uint4 per tinfo record
};

struct StreamParams {
    int nS, has_const, want_D, nchunks, ntinfo;
    int P, pb0;                              // channels [pb0, pb0 + PB) of P are computed by this launch
    int nblk[kStreamWarps];                  // [NW] blocks of each sub-stream before its inert padding
    const uint4* stream;                     // [kStreamWarps][nchunks][CH]
    const unsigned* ctl;                     // [kStreamWarps][nchunks * KB]
    const uint4* tinfo;                      // [kStreamWarps][ntinfo][TIQ]
    const double* w0;                        // [PB * CS] the constant term of each channel
    const c2* Ac; long long ldA;
    c2* Dt;                                  // [nS][P][ldA]
    double* E;                               // [nenv][P]
    long long nenv;
};

__device__ __forceinline__ c2 lds_c2(const unsigned char* base, unsigned off)
{
    return *reinterpret_cast<const c2*>(base + off);
}

// a = *(c2*)(base + off) only if pred: a predicated LDS.128 costs no shared-memory bandwidth when it is off.
// (Written as PTX because the compiler otherwise turns the uniform branch into load + select.)
__device__ __forceinline__ void lds_c2_if(c2& a, const unsigned char* base, unsigned off, bool pred)
{
#ifdef __CUDA_ARCH__
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(base) + off;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p ld.shared.v2.f64 {%0, %1}, [%2];\n\t}"
                 : "+d"(a.x), "+d"(a.y) : "r"(saddr), "r"((int)pred));
#else
    if (pred) a = *reinterpret_cast<const c2*>(base + off);
#endif
}

__device__ __forceinline__ double xor_hi(double v, unsigned mask)
{
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
#else
    union { double d; unsigned long long u; } w; w.d = v; w.u ^= ((unsigned long long)mask << 32); return w.d;
#endif
}


// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on mbarriers: how the A tile and the stream
// chunks reach shared memory without passing through registers.  The emulated test build uses plain loads.
#if !defined(ACEB200_EMU) && !defined(ACEB200_NO_TMA)
#define ACEB200_TMA 1
#else
#define ACEB200_TMA 0
#endif
#if ACEB200_TMA
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity)
{
    unsigned ok = 0;
    int spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (!ok && ++spins > (1 << 22)) __trap();      // a lost transaction must fail loudly, never hang the GPU
    } while (!ok);
}
#endif

// One CTA = kStreamWarps warps sharing one shared-memory tile of A (32 * EPL environments, EPL per lane).  The
// host splits the targets into kStreamWarps balanced sub-streams; warp w walks sub-stream w.  More warps
// per byte of shared memory is what hides the FP64 and shared-memory latencies of the leaf products; EPL = 2
// amortises the (warp-uniform) record decode over two environments and doubles the independent work per lane.
template <int NF, int PB, bool CW, int EPL>
__global__ void __launch_bounds__(32 * StreamGeom<NF, PB, CW>::NW, (PB == 1 && !CW) ? 2 : 1) k_adjoint_stream(const StreamParams p)
{
    typedef StreamGeom<NF, PB, CW> G;
    constexpr int CS = G::CS, CH = G::CH, QBP = G::QBP, KB = G::KB, LPC = G::LPC, TIQ = G::TIQ, NW = G::NW;
    constexpr int TW = 32 * EPL;                                // environments per tile
    constexpr int RSH = (EPL == 1) ? 9 : 10;                    // log2 of the tile row pitch in bytes
    ACE_DYN_SMEM(c2, As);                                       // [nS + 1][TW]; slot nS holds 1
    uint4* rings = reinterpret_cast<uint4*>(As + (size_t)(p.nS + 1) * TW);   // [NW][2][CH]
    double* Epart = reinterpret_cast<double*>(rings + NW * 2 * CH); // [NW][PB][EPL][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* ring = rings + warp * 2 * CH;
#if ACEB200_TMA
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(Epart + NW * PB * EPL * 32);   // [1 + 2 NW]
    unsigned long long* barA = bars;
    unsigned long long* barR = bars + 1 + 2 * warp;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 1 + 2 * NW; ++i) mbar_init(bars + i, 1);
        fence_barrier_init();
    }
    unsigned phA = 0, phR0 = 0, phR1 = 0;
#endif
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) As[p.nS * TW + lane + 32 * j] = c2{1.0, 0.0};
    }
    __syncthreads();
    constexpr bool GR = (PB == 1 && !CW);                       // grouped leaves (see the stream layout above)
    const int nblkw = p.nblk[warp];                             // this warp's sub-stream stops at its last real block
    const int nchw = (nblkw + KB - 1) / KB;
    const uint4* stream = p.stream + (size_t)warp * p.nchunks * CH;
    const unsigned* ctl = p.ctl + (size_t)warp * p.nchunks * KB;
    const uint4* tinfo = p.tinfo + (size_t)warp * p.ntinfo * TIQ;
    const unsigned char* Ab = reinterpret_cast<const unsigned char*>(As) + lane * 16;   // this lane's first column
    const long long ntiles = (p.nenv + TW - 1) / TW;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long e = tile * TW + lane;                   // environments e, e + 32, ...
#if ACEB200_TMA
        // one thread arms the tile barrier and issues one bulk copy per slot row (TW * 16 contiguous bytes each);
        // lane 0 of every warp primes its two ring slots
        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(barA, (unsigned)(p.nS * TW * sizeof(c2)));
            for (int s = 0; s < p.nS; ++s) bulk_g2s(As + s * TW, p.Ac + (size_t)s * p.ldA + tile * TW, TW * sizeof(c2), barA);
        }
        if (lane == 0 && nchw > 0) {
            fence_proxy_async();
            mbar_expect_tx(barR, CH * sizeof(uint4));
            bulk_g2s(ring, stream, CH * sizeof(uint4), barR);
            if (nchw > 1) {
                mbar_expect_tx(barR + 1, CH * sizeof(uint4));
                bulk_g2s(ring + CH, stream + CH, CH * sizeof(uint4), barR + 1);
            }
        }
        mbar_wait(barA, phA);
        phA ^= 1u;
#else
        for (int s = warp; s < p.nS; s += NW) {
#pragma unroll
            for (int j = 0; j < EPL; ++j) As[s * TW + lane + 32 * j] = p.Ac[(size_t)s * p.ldA + e + 32 * j];
        }
        if (nchw > 0) {
#pragma unroll
            for (int k = 0; k < LPC; ++k) ring[k * 32 + lane] = __ldg(stream + k * 32 + lane);
        }
        if (nchw > 1) {
#pragma unroll
            for (int k = 0; k < LPC; ++k) ring[CH + k * 32 + lane] = __ldg(stream + CH + k * 32 + lane);
        }
        __syncthreads();
#endif
        double E[PB][EPL];
        c2 D[PB][EPL], S[PB][EPL];
        c2 Tg[EPL];                      // GR: sum of w * (other factors) over the current group
        c2 Ug[EPL];                      // GR, order 4: the same over the current sub-group
#pragma unroll
        for (int j = 0; j < EPL; ++j) { Tg[j] = c2{0.0, 0.0}; Ug[j] = c2{0.0, 0.0}; }
#pragma unroll
        for (int q = 0; q < PB; ++q)
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                E[q][j] = (p.has_const && warp == 0) ? __ldg(p.w0 + q * CS) : 0.0;
                D[q][j] = c2{0.0, 0.0}; S[q][j] = c2{0.0, 0.0};
            }
        int ti = 0;
        for (int ch = 0; ch < nchw; ++ch) {
            const bool havepre = ch + 2 < nchw;
#if ACEB200_TMA
            if (ch & 1) { mbar_wait(barR + 1, phR1); phR1 ^= 1u; }
            else { mbar_wait(barR, phR0); phR0 ^= 1u; }
#else
            uint4 pre[LPC];
#pragma unroll
            for (int k = 0; k < LPC; ++k) pre[k] = uint4{0u, 0u, 0u, 0u};
            if (havepre) {
#pragma unroll
                for (int k = 0; k < LPC; ++k) pre[k] = __ldg(stream + (size_t)(ch + 2) * CH + k * 32 + lane);
            }
#endif
            const uint4* rb = ring + (ch & 1) * CH;
            const unsigned* cb = ctl + (size_t)ch * KB;
            const int nb = (nblkw - ch * KB < KB) ? nblkw - ch * KB : KB;
#pragma unroll 4
            for (int b = 0; b < nb; ++b) {
                const uint4* blk = rb + b * QBP;
                const unsigned flags = __ldg(cb + b);
                const uint4 cw = blk[0];
                const unsigned code[4] = {cw.x, cw.y, cw.z, cw.w};
                unsigned code2[4] = {0u, 0u, 0u, 0u};
                if (NF > 2) { const uint4 c3 = blk[1]; code2[0] = c3.x; code2[1] = c3.y; code2[2] = c3.z; code2[3] = c3.w; }
                const double* wb = reinterpret_cast<const double*>(blk + G::CWORDS);     // [4][PB][CS]
                c2 acc1[EPL];                    // second accumulator (PB == 1 only): halves the dependent chain
                c2 a1[EPL];
#pragma unroll
                for (int j = 0; j < EPL; ++j) { acc1[j] = c2{0.0, 0.0}; a1[j] = c2{0.0, 0.0}; }
                unsigned s1prev = 0xffffffffu;
#pragma unroll
                for (int k = 0; k < kBlkLeaves; ++k) {
                    const unsigned c = code[k];
                    if (GR && NF == 3) {
                        // Two-level (Horner) form for correlation order 4: leaves are grouped by a shared factor a1 and,
                        // inside a group, sub-grouped by a second shared factor a2:
                        //     U += w * a3^{c3}                     per leaf        (one load, two FMAs)
                        //     T += a2^{c2} * U,  U = 0             per sub-group   (kSubEnd, bit 14 of the second code word)
                        //     S += a1^{c1} * T,  T = 0             per group       (kGroupEnd)
                        // instead of one three-factor product per leaf.  Shared sub-products are formed once -- what the
                        // reference's (disabled) recursive graph evaluator does (src/grapheval.jl; src/linearmodel.jl:50-51),
                        // here in the order the stream is walked, with no intermediate storage.  Leaves of order 2 and 3 use the
                        // same form with the slot of ones as the missing factors.
                        const unsigned o3 = (code2[k] & 0x3fffu) << RSH, m3 = (code2[k] & 0x8000u) << 16;
                        const double w = wb[k];
#pragma unroll
                        for (int j = 0; j < EPL; ++j) {
                            c2 a3 = lds_c2(Ab + 512 * j, o3);
                            a3.y = xor_hi(a3.y, m3);
                            Ug[j].x += w * a3.x; Ug[j].y += w * a3.y;
                        }
                        if (code2[k] & kSubEnd) {
                            const unsigned o2 = ((c >> 16) & 0x3fffu) << RSH, m2 = c & 0x80000000u;
#pragma unroll
                            for (int j = 0; j < EPL; ++j) {
                                c2 a = lds_c2(Ab + 512 * j, o2);
                                a.y = xor_hi(a.y, m2);
                                Tg[j].x += a.x * Ug[j].x - a.y * Ug[j].y;
                                Tg[j].y += a.x * Ug[j].y + a.y * Ug[j].x;
                                Ug[j] = c2{0.0, 0.0};
                            }
                        }
                        if (c & kGroupEnd) {
                            const unsigned o1 = (c & 0x3fffu) << RSH, m1 = (c & 0x8000u) << 16;
#pragma unroll
                            for (int j = 0; j < EPL; ++j) {
                                c2 a = lds_c2(Ab + 512 * j, o1);
                                a.y = xor_hi(a.y, m1);
                                S[0][j].x += a.x * Tg[j].x - a.y * Tg[j].y;
                                S[0][j].y += a.x * Tg[j].y + a.y * Tg[j].x;
                                Tg[j] = c2{0.0, 0.0};
                            }
                        }
                        continue;
                    }
                    if (GR) {
                        // T += w * a2^{c2} [* a3^{c3} * a4^{c4}];  at the group's last leaf  S += a1^{c1} * T
                        const unsigned o2 = ((c >> 16) & 0x3fffu) << RSH, m2 = c & 0x80000000u;
                        c2 X[EPL];
#pragma unroll
                        for (int j = 0; j < EPL; ++j) { X[j] = lds_c2(Ab + 512 * j, o2); X[j].y = xor_hi(X[j].y, m2); }
                        if (NF > 2) {
                            const unsigned o3 = (code2[k] & 0x3fffu) << RSH, m3 = (code2[k] & 0x8000u) << 16;
#pragma unroll
                            for (int j = 0; j < EPL; ++j) { c2 a3 = lds_c2(Ab + 512 * j, o3); a3.y = xor_hi(a3.y, m3); X[j] = cmul(X[j], a3); }
                        }
                        if (NF > 3) {
                            const unsigned o4 = ((code2[k] >> 16) & 0x3fffu) << RSH, m4 = code2[k] & 0x80000000u;
#pragma unroll
                            for (int j = 0; j < EPL; ++j) { c2 a4 = lds_c2(Ab + 512 * j, o4); a4.y = xor_hi(a4.y, m4); X[j] = cmul(X[j], a4); }
                        }
                        const double w = wb[k];
#pragma unroll
                        for (int j = 0; j < EPL; ++j) { Tg[j].x += w * X[j].x; Tg[j].y += w * X[j].y; }
                        if (c & kGroupEnd) {
                            const unsigned o1 = (c & 0x3fffu) << RSH, m1 = (c & 0x8000u) << 16;
#pragma unroll
                            for (int j = 0; j < EPL; ++j) {
                                c2 a = lds_c2(Ab + 512 * j, o1);
                                a.y = xor_hi(a.y, m1);
                                S[0][j].x += a.x * Tg[j].x - a.y * Tg[j].y;
                                S[0][j].y += a.x * Tg[j].y + a.y * Tg[j].x;
                                Tg[j] = c2{0.0, 0.0};
                            }
                        }
                        continue;
                    }
                    const unsigned s1 = c & 0x3fffu;
                    const unsigned o1 = s1 << RSH, o2 = ((c >> 16) & 0x3fffu) << RSH;
                    const unsigned m2 = c & 0x80000000u, mf = (c & 0x8000u) << 16;
                    c2 X[EPL];
#pragma unroll
                    for (int j = 0; j < EPL; ++j) {
                        // leaves are sorted by their first factor: re-fetch it only when it changes (warp-uniform)
                        lds_c2_if(a1[j], Ab + 512 * j, o1, s1 != s1prev);
                        c2 a2 = lds_c2(Ab + 512 * j, o2);
                        a2.y = xor_hi(a2.y, m2);
                        X[j] = cmul(a1[j], a2);
                    }
                    s1prev = s1;
                    if (NF > 2) {
                        const unsigned o3 = (code2[k] & 0x3fffu) << RSH, m3 = (code2[k] & 0x8000u) << 16;
#pragma unroll
                        for (int j = 0; j < EPL; ++j) { c2 a3 = lds_c2(Ab + 512 * j, o3); a3.y = xor_hi(a3.y, m3); X[j] = cmul(X[j], a3); }
                    }
                    if (NF > 3) {
                        const unsigned o4 = ((code2[k] >> 16) & 0x3fffu) << RSH, m4 = code2[k] & 0x80000000u;
#pragma unroll
                        for (int j = 0; j < EPL; ++j) { c2 a4 = lds_c2(Ab + 512 * j, o4); a4.y = xor_hi(a4.y, m4); X[j] = cmul(X[j], a4); }
                    }
#pragma unroll
                    for (int j = 0; j < EPL; ++j) X[j].y = xor_hi(X[j].y, mf);
#pragma unroll
                    for (int q = 0; q < PB; ++q) {
                        if (CW) {
                            const double wr = wb[(k * PB + q) * 2], wi = wb[(k * PB + q) * 2 + 1];
#pragma unroll
                            for (int j = 0; j < EPL; ++j) {
                                S[q][j].x += wr * X[j].x - wi * X[j].y;
                                S[q][j].y += wr * X[j].y + wi * X[j].x;
                            }
                        } else {
                            const double w = wb[k * PB + q];
#pragma unroll
                            for (int j = 0; j < EPL; ++j) {
                                if (PB == 1 && (k & 1)) { acc1[j].x += w * X[j].x; acc1[j].y += w * X[j].y; }
                                else { S[q][j].x += w * X[j].x; S[q][j].y += w * X[j].y; }
                            }
                        }
                    }
                }
                if (PB == 1 && !CW && !GR) {
#pragma unroll
                    for (int j = 0; j < EPL; ++j) { S[0][j].x += acc1[j].x; S[0][j].y += acc1[j].y; }
                }
                if (flags) {
                    const uint4 t0 = __ldg(tinfo + (size_t)ti * TIQ);
                    const double* td = reinterpret_cast<const double*>(tinfo + (size_t)ti * TIQ + 1);   // scale, w1[PB][CS]
                    ++ti;
                    const bool neg = (flags & kTgtNeg) != 0u, odd = (flags & kTgtOdd) != 0u;
                    // fold onto the m >= 0 slot: Re(D- grad(phi_-m)) = Re((-1)^m conj(D-) grad(phi_m))
                    const double fx = (neg && odd) ? -1.0 : 1.0, fy = neg ? (odd ? 1.0 : -1.0) : 1.0;
                    const double scale = __ldg(td);
#pragma unroll
                    for (int j = 0; j < EPL; ++j) {
                        c2 Aa = lds_c2(Ab + 512 * j, EPL == 1 ? t0.x : t0.x * 2u);
                        Aa.x = xor_hi(Aa.x, t0.y);
                        Aa.y = xor_hi(Aa.y, t0.z);
                        if (flags & kSegEnd) {
#pragma unroll
                            for (int q = 0; q < PB; ++q) {
                                E[q][j] += (Aa.x * S[q][j].x - Aa.y * S[q][j].y) * scale;
                                D[q][j].x += fx * S[q][j].x;
                                D[q][j].y += fy * S[q][j].y;
                                S[q][j] = c2{0.0, 0.0};
                            }
                        }
                        if (flags & kTgtEnd) {
                            // order-1 term of this target: dE/dA_a += c~, E += Re(A_a c~)
#pragma unroll
                            for (int q = 0; q < PB; ++q) {
                                const double wr = __ldg(td + 1 + q * CS), wi = CW ? __ldg(td + 2 + q * CS) : 0.0;
                                E[q][j] += Aa.x * wr - Aa.y * wi;
                                D[q][j].x += fx * wr;
                                D[q][j].y += fy * wi;
                            }
                        }
                        if (flags & kSlotEnd) {
#pragma unroll
                            for (int q = 0; q < PB; ++q) {
                                if (p.want_D && e + 32 * j < p.nenv && p.pb0 + q < p.P)
                                    p.Dt[((size_t)(flags >> 8) * p.P + p.pb0 + q) * p.ldA + e + 32 * j] = D[q][j];
                                D[q][j] = c2{0.0, 0.0};
                            }
                        }
                    }
                }
            }
            __syncwarp();            // every lane is done reading this ring slot
#if ACEB200_TMA
            if (havepre && lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(barR + (ch & 1), CH * sizeof(uint4));
                bulk_g2s(ring + (ch & 1) * CH, stream + (size_t)(ch + 2) * CH, CH * sizeof(uint4), barR + (ch & 1));
            }
#else
            if (havepre) {
#pragma unroll
                for (int k = 0; k < LPC; ++k) ring[(ch & 1) * CH + k * 32 + lane] = pre[k];
            }
            __syncwarp();
#endif
        }
#pragma unroll
        for (int q = 0; q < PB; ++q)
#pragma unroll
            for (int j = 0; j < EPL; ++j) Epart[((warp * PB + q) * EPL + j) * 32 + lane] = E[q][j];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < PB; ++q)
#pragma unroll
                for (int j = 0; j < EPL; ++j) {
                    if (p.pb0 + q < p.P && e + 32 * j < p.nenv) {
                        double Et = 0.0;
#pragma unroll
                        for (int w = 0; w < NW; ++w) Et += Epart[((w * PB + q) * EPL + j) * 32 + lane];
                        p.E[(size_t)(e + 32 * j) * p.P + p.pb0 + q] = Et;
                    }
                }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// k_basis_stream: evaluate(basis::SymmetricBasis, cfg) fused -- B = real?(A2Bmap . AA), AA = prod A, without AA in HBM
// ------------------------------------------------------------------------------------------------
// [replaces evaluate!(AA, pibasis, A) (src/pibasis.jl:265-275) + genmul! (src/symmbasis.jl:248-264, 312-316)]
// Same mapping as k_adjoint_stream: the pooled A of 32 environments sits in shared memory as [slot][lane] (brought in by
// TMA bulk copies), one lane = one environment, and every warp walks its own pre-flattened stream of leaves with
// warp-uniform control flow.  A leaf is one non-zero of A2Bmap: up to NFAC slot codes (the AA function's factors) and, per
// real output channel, a weight pair (p, q):  out += p Re(X) - q Im(X),  X = prod of the factors.  The host folds into
// (p, q): the A2Bmap value (complex in general), all (-1)^m signs and the conjugation of the first factor, the real /
// imaginary part selection of a complex B, and -- for a real B -- the mirror partner of the AA function (the function
// with every m negated equals +-conj(AA), so one product serves both non-zeros).  Rows of B are dealt to the warps in
// contiguous ranges balanced by leaf count; a leaf carries a row-end bit.  Finished rows go to a per-warp shared-memory
// staging area [environment][W] and leave the SM as contiguous >= 128-byte segments per environment: B is
// environment-major ([env][row][component]) while the lanes are environments, so direct stores would be 8-byte pieces
// 2-55 KB apart.  Only B crosses HBM (config 4b: 55 KB per environment out, 0.7 KB in).
constexpr unsigned kRowEnd = 0x4000u;      // leaf code bit 14
constexpr int kBasisMaxWarps = 16;

template <int NFAC, int NCH, bool CW>
struct BasisGeom {
    static constexpr int CS = CW ? 2 : 1;
    static constexpr bool MASKED = (NFAC == 3 && NCH >= 3 && NCH <= 9);  // bits 16.. of the second code word: channel mask of the leaf
    // a leaf is one contiguous record: header (the two code words; padded to 16 bytes when the weights are read with LDS.128),
    // then NCH x CS weights.  The walk is a running pointer with compile-time offsets (the earlier block-of-four layout cost
    // ~8 index instructions per leaf).
    static constexpr int NWD = NCH * CS;                                 // weights (doubles) per leaf
    static constexpr int HDR = (NWD == 1) ? 8 : 16;
    static constexpr int LB = (HDR + 8 * NWD + 15) / 16 * 16;            // bytes per leaf
    static constexpr int LPC = (LB <= 32) ? 512 / LB : ((1024 / LB > 0) ? 1024 / LB : 1);   // leaves per ring chunk (~1 KB; 2 KB chunks measured the
                                                                         // same at 4b; 512 B for the 16-byte leaves keeps the rings small)
    static constexpr int CH = LPC * LB / 16;                             // uint4 per chunk
    static constexpr int NSLOT = 4;                                      // ring slots: three chunks in flight ahead of the one being read (a
                                                                         // two-slot ring left the L2 latency of the ~1 KB chunks exposed: the
                                                                         // mbarrier wait was the top stall of the 9- and 16-channel streams)
    static constexpr int ROWS = (NCH >= 9) ? 1 : (NCH <= 2) ? 4 / NCH : (16 + NCH - 1) / NCH;   // rows staged per flush: >= 72 contiguous bytes per
                                                                         // environment; one 32-byte sector for 1 / 2 channels, where a small
                                                                         // staging area lets two CTAs share an SM (the tile loads then overlap)
    static constexpr int W = ROWS * NCH;                                 // doubles staged per environment
    static constexpr int WP = W | 1;                                     // odd pitch: conflict-free staging
};

struct BasisParams {
    int nS, nw, nchunks;
    int nleaf[kBasisMaxWarps];               // leaves of each warp's sub-stream
    int row0[kBasisMaxWarps];                // first B row of each warp
    const uint4* stream;                     // [nw][nchunks][CH]
    const c2* Ac; long long ldA;
    double* out;                             // [nenv][rowlen]
    long long rowlen;                        // nB * NCH
    long long nenv;
};

// EPL environments per lane (1 or 2): the (warp-uniform) leaf decode and the broadcast weight loads are shared by the
// lane's environments, which also doubles the independent work per lane.
template <int NFAC, int NCH, bool CW, int EPL>
__global__ void __launch_bounds__((EPL == 2 && NCH >= 9) ? 384 : 32 * kBasisMaxWarps, 1) k_basis_stream(const BasisParams p)
{
    typedef BasisGeom<NFAC, NCH, CW> G;
    constexpr int CS = G::CS, CH = G::CH, LPC = G::LPC, W = G::W, WP = G::WP, NSLOT = G::NSLOT;
    constexpr int TW = 32 * EPL;                                        // environments per tile
    constexpr int RSH = (EPL == 1) ? 9 : 10;                            // log2 of the tile row pitch in bytes
    const int NW = p.nw;
    ACE_DYN_SMEM(c2, As);                                               // [nS + 1][TW]; slot nS holds 1
    uint4* rings = reinterpret_cast<uint4*>(As + (size_t)(p.nS + 1) * TW);   // [NW][NSLOT][CH]
    double* stg_all = reinterpret_cast<double*>(rings + (size_t)NW * NSLOT * CH);  // [NW][TW][WP]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* ring = rings + (size_t)warp * NSLOT * CH;
    double* stg = stg_all + (size_t)warp * TW * WP;
#if ACEB200_TMA
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(stg_all + (size_t)NW * TW * WP);   // [1 + NSLOT NW]
    unsigned long long* barA = bars;
    unsigned long long* barR = bars + 1 + NSLOT * warp;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 1 + NSLOT * NW; ++i) mbar_init(bars + i, 1);
        fence_barrier_init();
    }
    unsigned phA = 0, phR = 0;                  // bit s of phR: parity the next wait on ring slot s expects
#endif
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) As[p.nS * TW + lane + 32 * j] = c2{1.0, 0.0};
    }
    __syncthreads();
    const int nlw = p.nleaf[warp];
    const int nchw = (nlw + LPC - 1) / LPC;
    const uint4* stream = p.stream + (size_t)warp * p.nchunks * CH;
    const unsigned char* Ab = reinterpret_cast<const unsigned char*>(As) + lane * 16;
    const long long ntiles = (p.nenv + TW - 1) / TW;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#if ACEB200_TMA
        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(barA, (unsigned)(p.nS * TW * sizeof(c2)));
            for (int s = 0; s < p.nS; ++s) bulk_g2s(As + s * TW, p.Ac + (size_t)s * p.ldA + tile * TW, TW * sizeof(c2), barA);
        }
        if (lane == 0) {
            fence_proxy_async();
            for (int c = 0; c < NSLOT - 1 && c < nchw; ++c) {
                mbar_expect_tx(barR + c, CH * sizeof(uint4));
                bulk_g2s(ring + c * CH, stream + (size_t)c * CH, CH * sizeof(uint4), barR + c);
            }
        }
        mbar_wait(barA, phA);
        phA ^= 1u;
#else
        for (int s = warp; s < p.nS; s += NW) {
#pragma unroll
            for (int j = 0; j < EPL; ++j) As[s * TW + lane + 32 * j] = p.Ac[(size_t)s * p.ldA + tile * TW + lane + 32 * j];
        }
        for (int c = 0; c < NSLOT - 1 && c < nchw; ++c)
            for (int k = lane; k < CH; k += 32) ring[c * CH + k] = __ldg(stream + (size_t)c * CH + k);
        __syncthreads();
#endif
        double S[NCH][EPL];
#pragma unroll
        for (int q = 0; q < NCH; ++q)
#pragma unroll
            for (int j = 0; j < EPL; ++j) S[q][j] = 0.0;
        int pos = 0;                                   // doubles staged per environment
        long long o0 = (long long)p.row0[warp] * NCH;  // output offset (within an environment's row) of stg[.][0]
        // write the staged [environment][pos] block: a store instruction covers the contiguous runs of 32 / n whole
        // environments (lane -> (environment el, position kl) once per flush; then running pointers, no index arithmetic
        // in the loop: the general (env, k) walk cost 17 instructions per trip, 20 per leaf at config 4b)
        auto flush = [&]() {
            __syncwarp();
            const int n = pos;                                           // == W except for the last flush of a warp
            int epi, el;
            if (n == W) { epi = 32 / W; el = lane / W; } else { epi = 32 / n; el = lane / n; }
            const int kl = lane - el * n;
            const int envmax = (int)((p.nenv - tile * TW) < TW ? (p.nenv - tile * TW) : TW);
            if (el < epi) {
                double* op = p.out + ((size_t)(tile * TW) + el) * p.rowlen + o0 + kl;
                const double* sp = stg + el * WP + kl;
                const size_t ostep = (size_t)epi * p.rowlen;
                const int sstep = epi * WP;
#pragma unroll 4
                for (int env = el; env < envmax; env += epi, op += ostep, sp += sstep) *op = *sp;
            }
            __syncwarp();
            o0 += n;
            pos = 0;
        };
        for (int ch = 0; ch < nchw; ++ch) {
            const int slot = ch % NSLOT;
            // the slot chunk ch - 1 occupied is free (every lane passed the __syncwarp that ended it): refill it with
            // chunk ch + NSLOT - 1, three chunks ahead of the one about to be read
            const int pre = ch + NSLOT - 1, pslot = pre % NSLOT;
#if ACEB200_TMA
            if (pre < nchw && lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(barR + pslot, CH * sizeof(uint4));
                bulk_g2s(ring + pslot * CH, stream + (size_t)pre * CH, CH * sizeof(uint4), barR + pslot);
            }
            mbar_wait(barR + slot, (phR >> slot) & 1u);
            phR ^= 1u << slot;
#else
            if (pre < nchw) for (int k = lane; k < CH; k += 32) ring[pslot * CH + k] = __ldg(stream + (size_t)pre * CH + k);
            __syncwarp();
#endif
            const unsigned char* lp = reinterpret_cast<const unsigned char*>(ring + slot * CH);
            const int nl = (nlw - ch * LPC < LPC) ? nlw - ch * LPC : LPC;
            // One leaf per trip and NOT unrolled: the body (product + NCH channel updates for EPL environments + the
            // row-end path with its flush) is ~100-250 instructions; unrolled four times it overflowed the instruction
            // cache with a dozen warps at different places in it (ncu: "no instruction" became the top stall).
#pragma unroll 1
            for (int lf = 0; lf < nl; ++lf, lp += G::LB) {
                {
                    const uint2 cw2 = *reinterpret_cast<const uint2*>(lp);
                    const unsigned c = cw2.x, cc2 = cw2.y;
                    const double* wl = reinterpret_cast<const double*>(lp + G::HDR);                            // [NCH][CS]
                    c2 X[EPL];
#pragma unroll
                    for (int j = 0; j < EPL; ++j) X[j] = lds_c2(Ab + 512 * j, (c & 0x3fffu) << RSH);
                    if (NFAC >= 2) {
                        const unsigned o2 = ((c >> 16) & 0x3fffu) << RSH, m2 = c & 0x80000000u;
#pragma unroll
                        for (int j = 0; j < EPL; ++j) { c2 a2 = lds_c2(Ab + 512 * j, o2); a2.y = xor_hi(a2.y, m2); X[j] = cmul(X[j], a2); }
                    }
                    if (NFAC >= 3) {
                        const unsigned o3 = (cc2 & 0x3fffu) << RSH, m3 = (cc2 & 0x8000u) << 16;
#pragma unroll
                        for (int j = 0; j < EPL; ++j) { c2 a3 = lds_c2(Ab + 512 * j, o3); a3.y = xor_hi(a3.y, m3); X[j] = cmul(X[j], a3); }
                    }
                    if (NFAC >= 4) {
                        const unsigned o4 = ((cc2 >> 16) & 0x3fffu) << RSH, m4 = cc2 & 0x80000000u;
#pragma unroll
                        for (int j = 0; j < EPL; ++j) { c2 a4 = lds_c2(Ab + 512 * j, o4); a4.y = xor_hi(a4.y, m4); X[j] = cmul(X[j], a4); }
                    }
#pragma unroll
                    for (int q = 0; q < NCH; ++q) {
                        // channels whose weights are exactly zero for this leaf (warp-uniform mask in the unused fourth-factor
                        // field): vector / matrix-valued couplings touch 1.6 of 3 / 3.7 of 9 components per non-zero on average
                        if (G::MASKED && !((cc2 >> (16 + q)) & 1u)) continue;
                        if (CW) {        // two FMAs (the host stores -q): written out so that no separate multiply / add is formed
                            const double w0 = wl[q * 2], w1 = wl[q * 2 + 1];
#pragma unroll
                            for (int j = 0; j < EPL; ++j) { S[q][j] = fma(w0, X[j].x, S[q][j]); S[q][j] = fma(w1, X[j].y, S[q][j]); }
                        } else {
                            const double w0 = wl[q];
#pragma unroll
                            for (int j = 0; j < EPL; ++j) S[q][j] = fma(w0, X[j].x, S[q][j]);
                        }
                    }
                    if (c & kRowEnd) {                 // warp-uniform: every lane walks the same stream
#pragma unroll
                        for (int j = 0; j < EPL; ++j)
#pragma unroll
                            for (int q = 0; q < NCH; ++q) { stg[(lane + 32 * j) * WP + pos + q] = S[q][j]; S[q][j] = 0.0; }
                        pos += NCH;
                        if (pos + NCH > W) flush();
                    }
                }
            }
            __syncwarp();            // every lane is done reading this ring slot
        }
        if (pos > 0) flush();
        __syncthreads();             // the A tile is free for the next tile's copies
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
    int TE, dpitch;          // environments per CTA; c2 elements per staged environment (odd: conflict-free broadcasts)
    double* G;               // [neighbour][nprop][3][ncomp], chunk-relative
};

#ifndef ACE_FORCE_MINB
#define ACE_FORCE_MINB 5      // resident CTAs per SM the register allocation of k_forces is held to
#endif
constexpr int kForceThreads = 128;
constexpr int kForceTEmax = 32;      // environments per CTA (ForceParams::TE), chosen by the host so that the CTA's
                                     // neighbours fill whole passes of kForceThreads

// u += D[n] R[n], v += D[n] dR[n] for n < cnt, with n a compile-time register index: a fall-through
// switch on the (warp-uniform) column length replaces a per-n predicate.  R_n lives in registers; dR_n/dr is
// parked in shared memory ([n][thread]) so that the register budget allows ACE_FORCE_MINB resident CTAs.
// dR_n/dr is re-read from shared memory at every use: through a plain pointer ptxas hoists all NMAX loads above the column
// loop and keeps them in 2 NMAX registers for the whole harmonics walk, which is what pushed k_forces into spilling.
#ifdef ACEB200_EMU
#define ACE_DR_LOAD(p) (*(p))
#else
#define ACE_DR_LOAD(p) (*reinterpret_cast<const volatile double*>(p))
#endif
// REALONLY (the m = 0 columns: Y_l^0 and e^{i 0 phi} are real, and the azimuthal term carries the factor m = 0): only
// Re u and Re v are ever used, so only Re D~ is loaded (8 bytes instead of 16) and two of the four FMAs are issued --
// 34 of the 74 slots of config 2 (measured: 3.45 -> 3.41 ms; the kernel is issue-bound, not FP64-bound).
template <int NMAX, int STRIDE, bool REALONLY>
__device__ __forceinline__ void column_dot(const c2* D, int cnt, const double (&Rn)[NMAX], const double* dRs,
                                           double& ur, double& ui, double& vr, double& vi)
{
#define ACE_TERM(n)                                                                      \
    case (n) + 1:                                                                        \
        if ((n) < NMAX) {                                                                \
            const double dr = ACE_DR_LOAD(dRs + (n) * kForceThreads);                    \
            if (REALONLY) {                                                              \
                const double dx = D[(size_t)(n) * STRIDE].x;                             \
                ur += dx * Rn[(n) < NMAX ? (n) : 0]; vr += dx * dr;                      \
            } else {                                                                     \
                const c2 d = D[(size_t)(n) * STRIDE];                                    \
                ur += d.x * Rn[(n) < NMAX ? (n) : 0]; ui += d.y * Rn[(n) < NMAX ? (n) : 0];  \
                vr += d.x * dr; vi += d.y * dr;                                          \
            }                                                                            \
        }
    switch (cnt < NMAX ? cnt : NMAX) {      // the clamp tells ptxas that the cases above NMAX are dead (-1.4 % on k_forces)
        ACE_TERM(31) ACE_TERM(30) ACE_TERM(29) ACE_TERM(28) ACE_TERM(27) ACE_TERM(26) ACE_TERM(25) ACE_TERM(24)
        ACE_TERM(23) ACE_TERM(22) ACE_TERM(21) ACE_TERM(20) ACE_TERM(19) ACE_TERM(18) ACE_TERM(17) ACE_TERM(16)
        ACE_TERM(15) ACE_TERM(14) ACE_TERM(13) ACE_TERM(12) ACE_TERM(11) ACE_TERM(10) ACE_TERM(9) ACE_TERM(8)
        ACE_TERM(7) ACE_TERM(6) ACE_TERM(5) ACE_TERM(4) ACE_TERM(3) ACE_TERM(2) ACE_TERM(1) ACE_TERM(0)
    default: break;
    }
#undef ACE_TERM
}

// A CTA owns TE consecutive environments: it stages their folded adjoints D~ in shared memory
// ([local env][slot][channel]: the slots of a column are a compile-time stride apart) together with the
// (species, l, m) -> column table and the environments' neighbour offsets, then runs one thread per neighbour
// of those environments.
template <int NMAX, int PB, bool SPECIES, int WALK>
__global__ void __launch_bounds__(kForceThreads, ACE_FORCE_MINB) k_forces(const ForceParams p)
{
    ACE_DYN_SMEM(c2, Ds);   // [TE][dpitch >= nS * PB], double dR[NMAX][kForceThreads], int colinfo[nQ * nPused], int joff[TE + 1]
    const int TE = p.TE, dpitch = p.dpitch;
    const int tid = threadIdx.x;
    const long long e0 = (long long)blockIdx.x * TE;
    if (e0 >= p.B.nenv) return;
    ACE_GATE(p.B);
    const int ne = (int)((p.B.nenv - e0) < TE ? (p.B.nenv - e0) : TE);
    const long long jbeg = p.B.off[e0];
    const int nS = p.C.nS, ncol = p.C.nQ * p.C.nPused;
    double* dRs = reinterpret_cast<double*>(Ds + (size_t)TE * dpitch) + tid;   // [NMAX][kForceThreads]
    int* colinfo = reinterpret_cast<int*>(dRs - tid + NMAX * kForceThreads);
    int* joff = colinfo + ncol;
    for (int i = tid; i < ncol; i += kForceThreads) {
        const int col = __ldg(p.C.colmap + i);
        colinfo[i] = col < 0 ? -1 : (__ldg(p.C.base + col) | (__ldg(p.C.cnt + col) << 16));
    }
    if (tid <= ne) joff[tid] = (int)(p.B.off[e0 + tid] - jbeg);
    const double* Rb = p.B.R + 3 * (jbeg - p.B.jbase);
    const int* spb = SPECIES ? p.B.species + (jbeg - p.B.jbase) : nullptr;
    double* Gb = p.G + (size_t)(jbeg - p.B.off[0]) * p.nprop * 3 * p.ncomp;

    for (int pb = 0; pb < p.P; pb += PB) {
        __syncthreads();
        for (int idx = tid; idx < nS * PB * TE; idx += kForceThreads) {
            const int el = idx % TE, sc = idx / TE, c = sc % PB, s = sc / PB;
            Ds[el * dpitch + sc] = (pb + c < p.P && el < ne) ? p.Dt[((size_t)s * p.P + pb + c) * p.ldA + e0 + el] : c2{0.0, 0.0};
        }
        __syncthreads();
        const int nj = joff[ne];
        int el = 0;                          // j only grows from pass to pass: the environment search resumes where it stopped
        for (int j = tid; j < nj; j += kForceThreads) {
            while (el + 1 < ne && joff[el + 1] <= j) ++el;
            const double x = Rb[3 * j], y = Rb[3 * j + 1], z = Rb[3 * j + 2];
            int q = 0;
            if (SPECIES) { q = spb[j] - 1; if (q < 0 || q >= p.C.nQ) q = 0; }
            const Spher sp = cart2spher(x, y, z);
            double Rn[NMAX];
            radial_ed_park<NMAX>(p.rp, sp.r, Rn, dRs, kForceThreads);
            const int* cinfo = colinfo + (SPECIES ? q * p.C.nPused : 0);
            const c2* Dj = Ds + el * dpitch;
            double S0[PB], S1[PB], S2[PB];
#pragma unroll
            for (int c = 0; c < PB; ++c) { S0[c] = 0.0; S1[c] = 0.0; S2[c] = 0.0; }
            for_each_lm_ed<WALK>(p.ap, sp, [&](int l, int m, double Pt, double dP, double epr, double epi) {
                const int ci = cinfo[index_p(l, m)];
                if (ci < 0) return;
                const int cnt = ci >> 16, base = ci & 0xffff;
                const double f0 = (m == 0) ? Pt : Pt * sp.sth;    // |Y| factor:  Y = ep * f0
                const double f1 = (double)m * Pt;
#pragma unroll
                for (int c = 0; c < PB; ++c) {
                    double ur = 0.0, ui = 0.0, vr = 0.0, vi = 0.0;
                    if (m == 0) {              // (compile-time in the statically unrolled walk)
                        column_dot<NMAX, PB, true>(Dj + base * PB + c, cnt, Rn, dRs, ur, ui, vr, vi);
                        S0[c] += f0 * (vr * epr);  // radial:   rhat * Re(v Y),  ep real
                        S2[c] += dP * (ur * epr);  // polar:    dP Re(u ep);  the azimuthal term has the factor m = 0
                    } else {
                        column_dot<NMAX, PB, false>(Dj + base * PB + c, cnt, Rn, dRs, ur, ui, vr, vi);
                        // z = u * ep ;  Re(v * ep)
                        const double zr = ur * epr - ui * epi;
                        const double zi = ur * epi + ui * epr;
                        const double ve = vr * epr - vi * epi;
                        S0[c] += f0 * ve;          // radial:   rhat * Re(v Y)
                        S1[c] += f1 * zi;          // azimuth:  m Pt Im(u ep)
                        S2[c] += dP * zr;          // polar:    dP Re(u ep)
                    }
                }
            });
            // g = rhat S0 + (1/r) [ sphi S1 + cphi cth S2,  -cphi S1 + sphi cth S2,  -sth S2 ]
            const double rx = sp.sth * sp.cphi, ry = sp.sth * sp.sphi, rz = sp.cth;   // rhat (x, y, z need not stay live)
#pragma unroll
            for (int c = 0; c < PB; ++c) {
                if (pb + c >= p.P) break;
                const double gx = rx * S0[c] + sp.rinv * (sp.sphi * S1[c] + sp.cphi * sp.cth * S2[c]);
                const double gy = ry * S0[c] + sp.rinv * (-sp.cphi * S1[c] + sp.sphi * sp.cth * S2[c]);
                const double gz = rz * S0[c] - sp.rinv * sp.sth * S2[c];
                const int ch = pb + c, prop = ch / p.ncomp, comp = ch % p.ncomp;
                double* g = Gb + (((size_t)j * p.nprop + prop) * 3) * p.ncomp + comp;
                g[0] = gx; g[p.ncomp] = gy; g[2 * p.ncomp] = gz;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_forces_mma: the force contraction on the FP64 tensor cores (single output channel, single species)
// ------------------------------------------------------------------------------------------------
//   u_{lm}(j) = sum_n D~_{n,lm} R_n(r_j),   v_{lm}(j) = sum_n D~_{n,lm} R_n'(r_j)
// is, per environment, the product  [j] x [n] . [n] x [(lm, re / im)]  (twice: R and R').  With eight neighbours as the
// rows of an m8n8k4 tile, four radial indices per k-step and four (l, m) columns as its eight real columns, every lane
// ends up holding the complex u and v of ONE (neighbour, column) pair.  The harmonics that u and v are contracted with
//     S0 += f0 Re(v ep),  S1 += m Pt Im(u ep),  S2 += dP Re(u ep)          (f0 = |Y| factor, ep = e^{i m phi} / sqrt 2)
// are staged by phase a in shared-memory planes [function][row] and read back in the fragment layout (lane -> row
// lane / 4, column lane % 4: conflict-free at a pitch of 4 mod 16); the three scalars are then summed over the four
// lanes of a quad and over the column tiles, and lanes 0..2 of each quad write one Cartesian component
//     g = rhat S0 + (1 / r) [ sphi S1 + cphi cth S2,  -cphi S1 + sphi cth S2,  -sth S2 ].
// Phase a is one thread per neighbour as in k_forces (R_n, R_n' by the recurrence, the harmonics walk with the
// pole-stable derivative), but it only produces operands -- no dot products, no column switch.
constexpr int kFmmaEnvs = 8;             // environments staged per sub-tile (their D~)
constexpr int kFmmaRows = 128;           // neighbours per sub-tile
constexpr int kFmmaThreads = 256;        // phase a: threads 0..127 radial part, 128..255 angular part of neighbour tid % 128;
                                         // phase b: eight warps
struct ForceTile { int offF[4], offE[4], base[4], cnt[4], ks, pad[3]; };   // four columns: double offsets of the column's harmonic
                                                                           // planes (F: [ip][3][P], E: [m][2][P]), first slot, length;
                                                                           // ks = k-steps (of four radial indices)

struct ForceMmaParams {
    RadialParams rp;
    AlpParams ap;
    BatchDev B;
    const c2* Dt; long long ldA;
    int nS, dpitch, TE, nP;
    const ForceTile* tiles; int ntiles;
    double* G;               // [neighbour][3], chunk-relative
};

// R_n and dR_n/dr straight into shared-memory planes (two rolling registers each)
template <int NMAX>
__device__ __forceinline__ void radial_ed_planes(const RadialParams& rp, double r, double* sR, double* sD, int stride)
{
    double t, dt, f, df;
    transform_ed(rp, r, t, dt);
    envelope_ed(rp, t, f, df);
    double r2 = rp.A[0] * f, d2 = rp.A[0] * df, r1 = 0.0, d1 = 0.0;
    sR[0] = r2; sD[0] = d2 * dt;
    if (NMAX > 1 && rp.N > 1) {
        const double al = rp.A[1] * t + rp.B[1];
        r1 = al * r2;
        d1 = al * d2 + rp.A[1] * r2;
        sR[stride] = r1; sD[stride] = d1 * dt;
    }
#pragma unroll
    for (int n = 2; n < NMAX; ++n) {
        if (n < rp.N) {
            const double al = rp.A[n] * t + rp.B[n];
            const double rn = al * r1 + rp.C[n] * r2;
            const double dn = al * d1 + rp.C[n] * d2 + rp.A[n] * r1;
            r2 = r1; r1 = rn; d2 = d1; d1 = dn;
            sR[n * stride] = rn; sD[n * stride] = dn * dt;
        }
    }
}

// u, v of one column tile for JT row tiles: KS k-steps of four radial indices (compile-time, so the chain is straight-line)
template <int KS, int JT, int NT4>
__device__ __forceinline__ void force_tile_mma(const double* dB, int cntB, int k4, const double (&aR)[JT][NT4], const double (&aD)[JT][NT4],
                                               double (&ur)[JT], double (&ui)[JT], double (&vr)[JT], double (&vi)[JT])
{
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
        const int n = kk * 4 + k4;
        const double b = n < cntB ? dB[2 * n] : 0.0;
#pragma unroll
        for (int s = 0; s < JT; ++s) {
            dmma(ur[s], ui[s], aR[s][kk], b);
            dmma(vr[s], vi[s], aD[s][kk], b);
        }
    }
}

// One unit: JT row tiles (eight neighbours each) of one environment against every column tile.
template <int JT, int NT4>
__device__ __forceinline__ void force_unit(const ForceTile* tl, int ntiles, const c2* De, const double* SR, const double* SD, const double* SF,
                                           const double* SE, const double* SGm, int N, int r0, int rb, int nrows, int k4, int g, double* Gout)
{
    constexpr int P = kMmaPitch;
    int row[JT];
#pragma unroll
    for (int s = 0; s < JT; ++s) { const int r = r0 + s * 8 + g; row[s] = r < nrows ? r : nrows - 1; }
    // A fragments (rows = neighbours, k = radial index): R and R'
    double aR[JT][NT4], aD[JT][NT4];
#pragma unroll
    for (int kk = 0; kk < NT4; ++kk) {
        const int n = kk * 4 + k4, nc = n < N ? n : N - 1;     // n >= N meets a zero of D~: any finite value will do
#pragma unroll
        for (int s = 0; s < JT; ++s) { aR[s][kk] = SR[nc * P + row[s]]; aD[s][kk] = SD[nc * P + row[s]]; }
    }
    double S0[JT], S1[JT], S2[JT];
#pragma unroll
    for (int s = 0; s < JT; ++s) { S0[s] = 0.0; S1[s] = 0.0; S2[s] = 0.0; }
    const int colB = g >> 1, part = g & 1;
#pragma unroll 1
    for (int t = 0; t < ntiles; ++t) {
        const ForceTile& T = tl[t];
        const int cntB = T.cnt[colB];
        const double* dB = reinterpret_cast<const double*>(De + T.base[colB]) + part;
        double ur[JT], ui[JT], vr[JT], vi[JT];
#pragma unroll
        for (int s = 0; s < JT; ++s) { ur[s] = 0.0; ui[s] = 0.0; vr[s] = 0.0; vi[s] = 0.0; }
        switch (T.ks) {
        case 1: force_tile_mma<1, JT, NT4>(dB, cntB, k4, aR, aD, ur, ui, vr, vi); break;
        case 2: force_tile_mma<(NT4 >= 2 ? 2 : 1), JT, NT4>(dB, cntB, k4, aR, aD, ur, ui, vr, vi); break;
        case 3: force_tile_mma<(NT4 >= 3 ? 3 : 1), JT, NT4>(dB, cntB, k4, aR, aD, ur, ui, vr, vi); break;
        case 4: force_tile_mma<(NT4 >= 4 ? 4 : 1), JT, NT4>(dB, cntB, k4, aR, aD, ur, ui, vr, vi); break;
        case 5: force_tile_mma<(NT4 >= 5 ? 5 : 1), JT, NT4>(dB, cntB, k4, aR, aD, ur, ui, vr, vi); break;
        case 6: force_tile_mma<(NT4 >= 6 ? 6 : 1), JT, NT4>(dB, cntB, k4, aR, aD, ur, ui, vr, vi); break;
        case 7: force_tile_mma<(NT4 >= 7 ? 7 : 1), JT, NT4>(dB, cntB, k4, aR, aD, ur, ui, vr, vi); break;
        default: force_tile_mma<NT4, JT, NT4>(dB, cntB, k4, aR, aD, ur, ui, vr, vi); break;
        }
        // this lane now holds u, v of (neighbour row[s], column k4 of the tile)
        const double* pF = SF + T.offF[k4];
        const double* pE = SE + T.offE[k4];
#pragma unroll
        for (int s = 0; s < JT; ++s) {
            const double er = pE[row[s]], ei = pE[P + row[s]];
            const double f0 = pF[row[s]], f1 = pF[P + row[s]], f2 = pF[2 * P + row[s]];
            const double ve = vr[s] * er - vi[s] * ei;
            const double zr = ur[s] * er - ui[s] * ei;
            const double zi = ur[s] * ei + ui[s] * er;
            S0[s] += f0 * ve; S1[s] += f1 * zi; S2[s] += f2 * zr;
        }
    }
    // sum over the four columns of the quad, then lanes 0..2 of the quad write x, y, z
#pragma unroll
    for (int s = 0; s < JT; ++s) {
#ifdef ACEB200_EMU
        S0[s] += emu::shfl_xor(S0[s], 1); S1[s] += emu::shfl_xor(S1[s], 1); S2[s] += emu::shfl_xor(S2[s], 1);
        S0[s] += emu::shfl_xor(S0[s], 2); S1[s] += emu::shfl_xor(S1[s], 2); S2[s] += emu::shfl_xor(S2[s], 2);
#else
        S0[s] += __shfl_xor_sync(0xffffffffu, S0[s], 1); S1[s] += __shfl_xor_sync(0xffffffffu, S1[s], 1); S2[s] += __shfl_xor_sync(0xffffffffu, S2[s], 1);
        S0[s] += __shfl_xor_sync(0xffffffffu, S0[s], 2); S1[s] += __shfl_xor_sync(0xffffffffu, S1[s], 2); S2[s] += __shfl_xor_sync(0xffffffffu, S2[s], 2);
#endif
        const int r = r0 + s * 8 + g;
        if (k4 < 3 && r < rb) {
            const double* gm = SGm + (3 * k4) * P + r;
            Gout[(size_t)r * 3 + k4] = gm[0] * S0[s] + gm[P] * S1[s] + gm[2 * P] * S2[s];
        }
    }
}

template <int NMAX, int WALK>
__global__ void __launch_bounds__(kFmmaThreads) k_forces_mma(const ForceMmaParams p)
{
    constexpr int P = kMmaPitch;
    constexpr int NT4 = (NMAX + 3) / 4;             // k-steps covering every radial index
    constexpr int NW = kFmmaThreads / 32;
    ACE_DYN_SMEM(double, smem);
    const int N = p.rp.N, nP = p.nP, ntiles = p.ntiles, L = p.ap.L;
    double* SR = smem;                                  // [N][P]        R_n
    double* SD = SR + (size_t)N * P;                    // [N][P]        dR_n/dr
    double* SF = SD + (size_t)N * P;                    // [nP][3][P]    Pt sin(th) (m > 0) | P_l^0;  m Pt;  dP/dtheta
    double* SE = SF + (size_t)3 * nP * P;               // [L + 1][2][P] Re, Im of e^{i m phi} / sqrt 2
    double* SGm = SE + (size_t)2 * (L + 1) * P;         // [9][P]        Cartesian assembly coefficients
    c2* Ds = reinterpret_cast<c2*>(SGm + 9 * P);        // [kFmmaEnvs][dpitch]
    ForceTile* tl = reinterpret_cast<ForceTile*>(Ds + (size_t)kFmmaEnvs * p.dpitch);   // [ntiles]
    int* joff = reinterpret_cast<int*>(tl + ntiles);    // [TE + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long e0 = (long long)blockIdx.x * p.TE;
    if (e0 >= p.B.nenv) return;
    ACE_GATE(p.B);
    const int ne = (int)((p.B.nenv - e0) < p.TE ? (p.B.nenv - e0) : p.TE);
    const long long jbeg = p.B.off[e0];
    if (tid <= ne) joff[tid] = (int)(p.B.off[e0 + tid] - jbeg);
    for (int i = tid; i < ntiles * (int)(sizeof(ForceTile) / sizeof(int)); i += kFmmaThreads)
        reinterpret_cast<int*>(tl)[i] = __ldg(reinterpret_cast<const int*>(p.tiles) + i);
    const double* Rb = p.B.R + 3 * (jbeg - p.B.jbase);
    double* Gb = p.G + (size_t)(jbeg - p.B.off[0]) * 3;
    const int k4 = lane & 3, g = lane >> 2;
    __syncthreads();

    int e = 0, j0 = 0;
    while (e < ne) {
        const int jend_e = joff[e + 1];
        int e2 = e + 1, j1;
        bool done_e;
        if (j0 == joff[e] && jend_e - j0 <= kFmmaRows) {
            while (e2 < ne && e2 - e < kFmmaEnvs && joff[e2 + 1] - j0 <= kFmmaRows) ++e2;
            j1 = joff[e2];
            done_e = true;
        } else {
            j1 = (j0 + kFmmaRows < jend_e) ? j0 + kFmmaRows : jend_e;
            done_e = (j1 == jend_e);
        }
        const int nrows = j1 - j0, nes = e2 - e;

        // ---- stage D~ of the sub-tile's environments: Ds[el][slot]
        for (int idx = tid; idx < p.nS * nes; idx += kFmmaThreads) {
            const int el = idx % nes, s = idx / nes;
            Ds[el * p.dpitch + s] = p.Dt[(size_t)s * p.ldA + e0 + e + el];
        }
        // ---- phase a: two threads per neighbour -> operand planes
        {
            const int rowa = tid & (kFmmaRows - 1);
            if (rowa < nrows) {
                const int j = j0 + rowa;
                const double x = Rb[3 * j], y = Rb[3 * j + 1], z = Rb[3 * j + 2];
                if (tid < kFmmaRows) {            // radial part
                    const double r2 = x * x + y * y + z * z;
#ifdef __CUDA_ARCH__
                    const double r = r2 * rsqrt(r2);
#else
                    const double r = r2 * (1.0 / sqrt(r2));
#endif
                    radial_ed_planes<NMAX>(p.rp, r, SR + rowa, SD + rowa, P);
                } else {                          // angular part + Cartesian assembly coefficients
                    const Spher sp = cart2spher(x, y, z);
                    double* gm = SGm + rowa;
                    gm[0 * P] = sp.sth * sp.cphi; gm[1 * P] = sp.rinv * sp.sphi;  gm[2 * P] = sp.rinv * sp.cphi * sp.cth;
                    gm[3 * P] = sp.sth * sp.sphi; gm[4 * P] = -sp.rinv * sp.cphi; gm[5 * P] = sp.rinv * sp.sphi * sp.cth;
                    gm[6 * P] = sp.cth;           gm[7 * P] = 0.0;                gm[8 * P] = -sp.rinv * sp.sth;
                    double* sf = SF + rowa;
                    double* se = SE + rowa;
                    for_each_lm_ed<WALK>(p.ap, sp, [&](int l, int m, double Pt, double dP, double epr, double epi) {
                        const int ip = index_p(l, m);
                        sf[(3 * ip) * P] = (m == 0) ? Pt : Pt * sp.sth;
                        sf[(3 * ip + 1) * P] = (double)m * Pt;
                        sf[(3 * ip + 2) * P] = dP;
                        if (l == m) { se[(2 * m) * P] = epr; se[(2 * m + 1) * P] = epi; }
                    });
                }
            }
        }
        __syncthreads();

        // ---- phase b: units of up to two row tiles (8 neighbours each) of one environment, dealt in contiguous runs to the warps
        int nunits = 0;
        for (int el = 0; el < nes; ++el) {
            int ra = joff[e + el] - j0, rb = joff[e + el + 1] - j0;
            if (ra < 0) ra = 0;
            if (rb > nrows) rb = nrows;
            nunits += (rb - ra + 15) >> 4;
        }
        const int per = (nunits + NW - 1) / NW;
        const int u0 = warp * per, u1 = (u0 + per < nunits) ? u0 + per : nunits;
        int el = 0, ubase = 0;                 // environment of the current unit and the index of its first unit
        for (int u = u0; u < u1; ++u) {
            int ra, rb;
            for (;;) {
                ra = joff[e + el] - j0; rb = joff[e + el + 1] - j0;
                if (ra < 0) ra = 0;
                if (rb > nrows) rb = nrows;
                const int nu_e = (rb - ra + 15) >> 4;
                if (u < ubase + nu_e) break;
                ubase += nu_e; ++el;
            }
            const int r0 = ra + (u - ubase) * 16;          // first staged row of this unit
            const c2* De = Ds + el * p.dpitch;
            if (r0 + 8 < rb) force_unit<2, NT4>(tl, ntiles, De, SR, SD, SF, SE, SGm, N, r0, rb, nrows, k4, g, Gb + (size_t)j0 * 3);
            else force_unit<1, NT4>(tl, ntiles, De, SR, SD, SF, SE, SGm, N, r0, rb, nrows, k4, g, Gb + (size_t)j0 * 3);
        }
        __syncthreads();
        j0 = j1;
        if (done_e) e = e2;
    }
}

// ------------------------------------------------------------------------------------------------
// basis-value kernels (evaluate on Product1pBasis / PIBasis / SymmetricBasis)
// ------------------------------------------------------------------------------------------------

// canonical slots -> the reference's A vector: A[e][iA] (src/product_1pbasis.jl:123-134)
static __global__ void k_expand_A(long long nenv, int nA, const int* code, const c2* Ac, long long ldA, c2* A)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nenv * nA) return;
    const long long e = t / nA;
    const int a = (int)(t % nA);
    const int cd = __ldg(code + a);
    A[t] = decode_A(Ac[(size_t)(cd >> 2) * ldA + e], cd);
}

// AA[e][i] = real?(prod_t A[e][spec[i][t]])  (src/pibasis.jl:265-275)
static __global__ void k_AA(long long nenv, int nA, int nAA, int maxord, const int* orders, const int* spec,
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
static __global__ void k_B(long long nenv, int nB, int nAA, int ncomp, const int* ptr, const int* col, const c2* val,
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
// adjoint_EVAL_D (src/evaluator.jl:204-244): dB_k = sum_j w_j . dB_k/dr_j without forming any Jacobian
// ------------------------------------------------------------------------------------------------
// Step 1 pools  dAw_s = sum_j w_j . grad phi_s(r_j)  exactly like k_pool pools A: with
//   w . grad(R_n Y_lm) = (R_n' w.rhat) Y_lm + R_n (w . grad Y_lm),   w . grad Y = ep (dP w_theta + i m Pt w_phi)
// two radial and two angular rows are staged per neighbour.  w is real, so dAw has the same m -> -m symmetry as A
// and lives in the same canonical slots.
struct PoolWParams {
    RadialParams rp;
    AlpParams ap;
    ColumnsDev C;
    BatchDev B;
    const double* W;        // [neighbour][3], indexed like R
    c2* Aw;                 // [nS][ldA]
    long long ldA;
    int TE, nP;
    int rows;               // neighbours staged per sub-tile: 128 unless the staging would not fit in shared memory
};

template <int NMAX, bool SPECIES>
__global__ void __launch_bounds__(kPoolThreads) k_pool_w(const PoolWParams p)
{
    ACE_DYN_SMEM(c2, smem);
    const int rows = p.rows, pitch = p.rows + 1;   // staged rows per sub-tile (<= kPoolThreads) and the row pitch
    c2* SY = smem;                                                             // [2 nP][129]: Y, w.grad Y
    double* SR = reinterpret_cast<double*>(SY + (size_t)2 * p.nP * pitch); // [2 N][129]: R' w.rhat, R
    int* sq = reinterpret_cast<int*>(SR + (size_t)2 * p.rp.N * pitch);
    int* joff = sq + kPoolThreads;
    const int tid = threadIdx.x;
    const int N = p.rp.N, nS = p.C.nS, nP = p.nP;
    const long long e0 = (long long)blockIdx.x * p.TE;
    if (e0 >= p.B.nenv) return;
    ACE_GATE(p.B);
    const int ne = (int)((p.B.nenv - e0) < p.TE ? (p.B.nenv - e0) : p.TE);
    const long long jbeg = p.B.off[e0];
    if (tid <= ne) joff[tid] = (int)(p.B.off[e0 + tid] - jbeg);
    const double* Rb = p.B.R + 3 * (jbeg - p.B.jbase);
    const double* Wb = p.W + 3 * (jbeg - p.B.jbase);
    const int* spb = SPECIES ? p.B.species + (jbeg - p.B.jbase) : nullptr;
    __syncthreads();
    c2 acc[kPoolItems];
#pragma unroll
    for (int it = 0; it < kPoolItems; ++it) acc[it] = c2{0.0, 0.0};
    int e = 0, j0 = 0;
    while (e < ne) {
        const int jend_e = joff[e + 1];
        int e2 = e + 1, j1;
        bool done_e;
        if (j0 == joff[e] && jend_e - j0 <= rows) {
            while (e2 < ne && joff[e2 + 1] - j0 <= rows && (e2 + 1 - e) * nS <= kPoolThreads * kPoolItems) ++e2;
            j1 = joff[e2];
            done_e = true;
        } else {
            j1 = (j0 + rows < jend_e) ? j0 + rows : jend_e;
            done_e = (j1 == jend_e);
        }
        const int nrows = j1 - j0;
        if (tid < nrows) {
            const int j = j0 + tid;
            const double x = Rb[3 * j], y = Rb[3 * j + 1], z = Rb[3 * j + 2];
            const double wx = Wb[3 * j], wy = Wb[3 * j + 1], wz = Wb[3 * j + 2];
            if (SPECIES) { int q = spb[j] - 1; if (q < 0 || q >= p.C.nQ) q = 0; sq[tid] = q; }
            const Spher sp = cart2spher(x, y, z);
            double Rn[NMAX], dRn[NMAX];
            radial_ed<NMAX>(p.rp, sp.r, Rn, dRn);
            const double wr = (wx * x + wy * y + wz * z) * sp.rinv;                                     // w . rhat
            const double wphi = (-sp.sphi * wx + sp.cphi * wy) * sp.rinv;
            const double wth = (sp.cphi * sp.cth * wx + sp.sphi * sp.cth * wy - sp.sth * wz) * sp.rinv;
#pragma unroll
            for (int n = 0; n < NMAX; ++n)
                if (n < N) { SR[n * pitch + tid] = dRn[n] * wr; SR[(N + n) * pitch + tid] = Rn[n]; }
            for_each_lm_ed(p.ap, sp, [&](int l, int m, double Pt, double dP, double epr, double epi) {
                const int ip = index_p(l, m);
                const double f0 = (m == 0) ? Pt : Pt * sp.sth;
                SY[ip * pitch + tid] = c2{epr * f0, epi * f0};
                const double gr = dP * wth, gi = (double)m * Pt * wphi;          // (gr + i gi) * ep
                SY[(nP + ip) * pitch + tid] = c2{epr * gr - epi * gi, epr * gi + epi * gr};
            });
        }
        __syncthreads();
        const int nitems = (e2 - e) * nS;
#pragma unroll
        for (int it = 0; it < kPoolItems; ++it) {
            const int idx = tid + it * kPoolThreads;
            if (idx < nitems) {
                const int el = idx / nS, s = idx - el * nS;
                const int n = __ldg(p.C.slot_n + s), ip = __ldg(p.C.slot_ip + s), q = SPECIES ? __ldg(p.C.slot_q + s) : 0;
                int ra = joff[e + el] - j0, rb = joff[e + el + 1] - j0;
                if (ra < 0) ra = 0;
                if (rb > nrows) rb = nrows;
                const double* pr1 = SR + n * pitch; const double* pr2 = SR + (N + n) * pitch;
                const c2* py1 = SY + ip * pitch; const c2* py2 = SY + (nP + ip) * pitch;
                c2 a = acc[it];
                for (int r = ra; r < rb; ++r) {
                    if (SPECIES && sq[r] != q) continue;
                    const double r1 = pr1[r], r2 = pr2[r];
                    const c2 y1 = py1[r], y2 = py2[r];
                    a.x += r1 * y1.x + r2 * y2.x;
                    a.y += r1 * y1.y + r2 * y2.y;
                }
                if (done_e) { p.Aw[(size_t)s * p.ldA + (e0 + e + el)] = a; a = c2{0.0, 0.0}; }
                acc[it] = a;
            }
        }
        __syncthreads();
        j0 = j1;
        if (done_e) e = e2;
    }
}

// dAAw[e][i] = real?( sum_t dAw[v_t] prod_{s != t} A[v_s] )   (src/evaluator.jl:228-235)
static __global__ void k_AAw(long long nenv, int nA, int nAA, int maxord, const int* orders, const int* spec,
                      const c2* A, const c2* Aw, int symreal, double* out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nenv * nAA) return;
    const long long e = t / nAA;
    const int i = (int)(t % nAA);
    const c2* Ae = A + (size_t)e * nA;
    const c2* We = Aw + (size_t)e * nA;
    const int o = __ldg(orders + i);
    int v[8];
    c2 adj[8];
    for (int k = 0; k < o; ++k) v[k] = __ldg(spec + (size_t)i * maxord + k);
    c2 run = c2{1.0, 0.0};
    for (int k = 0; k < o; ++k) { adj[k] = run; run = cmul(run, Ae[v[k]]); }
    run = c2{1.0, 0.0};
    for (int k = o - 1; k >= 0; --k) { adj[k] = cmul(adj[k], run); run = cmul(run, Ae[v[k]]); }
    c2 acc = c2{0.0, 0.0};
    for (int k = 0; k < o; ++k) { const c2 m = cmul(adj[k], We[v[k]]); acc.x += m.x; acc.y += m.y; }
    if (symreal) acc.y = 0.0;
    out[2 * t] = acc.x; out[2 * t + 1] = acc.y;
}

// out[e][row][c] = sum_k A2B[row,k][c] * dAAw[e][col_k]   (dB = A2Bmap * dAAw, src/evaluator.jl:237); complex
static __global__ void k_Bw(long long nenv, int nB, int nAA, int ncomp, const int* ptr, const int* col, const c2* val,
                     const double* AAw, double* out)
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
        const double ar = AAw[2 * ia], ai = AAw[2 * ia + 1];
        br += v.x * ar - v.y * ai;
        bi += v.x * ai + v.y * ar;
    }
    out[2 * t] = br; out[2 * t + 1] = bi;
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
    c2* dA;                  // [neighbour][nA][3], chunk-relative; canon: [neighbour / 32][nS][3][neighbour % 32] over the canonical
                             // slots (m >= 0) -- the plane layout of k_dB_env, written and read with full 512 B transactions
    long long nJ;
    int canon;
};

// dA[j][iA][:] = grad phi_iA(r_j) (src/product_1pbasis.jl:169-221), one thread per neighbour
template <int NMAX>
__global__ void __launch_bounds__(128, 4) k_dA(const dAParams p)   // 4 CTAs per SM (128 registers, 168 B of spills at NMAX = 12): 0.38 -> 0.26 ms per 6 x 10^5 neighbours; 6 CTAs (80 registers) 0.50 ms
{
    const long long jl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (jl >= p.nJ) return;
    ACE_GATE(p.B);
    const long long jabs = p.B.off[0] + jl;
    const double* r = p.B.R + 3 * (jabs - p.B.jbase);
    const double x = r[0], y = r[1], z = r[2];
    int q = 0;
    if (p.B.species) { q = p.B.species[jabs - p.B.jbase] - 1; if (q < 0 || q >= p.C.nQ) q = 0; }
    const Spher sp = cart2spher(x, y, z);
    double Rn[NMAX], dRn[NMAX];
    radial_ed<NMAX>(p.rp, sp.r, Rn, dRn);
    // canon: element (slot, xyz) of neighbour jl is out[(slot * 3 + xyz) * 32]
    c2* out = p.canon ? p.dA + (size_t)(jl >> 5) * p.C.nS * 3 * 32 + (jl & 31) : p.dA + (size_t)jl * p.nA * 3;
    // functions of other species are identically zero for this neighbour (one species: every canonical slot is written below)
    if (p.C.nQ <= 1) {}
    else if (!p.canon) for (int a = 0; a < p.nA * 3; ++a) out[a] = c2{0.0, 0.0};
    else for (int a = 0; a < p.C.nS * 3; ++a) out[(size_t)a * 32] = c2{0.0, 0.0};
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
                if (p.canon) { for (int k = 0; k < 3; ++k) out[((size_t)(base + n) * 3 + k) * 32] = g[k]; continue; }
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
static __global__ void k_dAA(long long nenv, const long long* off, int nA, int nAA, int maxord, const int* orders, const int* spec,
                      const c2* A, const c2* dA, int pireal, double* dAA, const int* gate)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nenv * nAA) return;
    if (gate && *gate == 1) return;
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
static __global__ void k_dB(long long nJ, int nB, int nAA, int ncomp, const int* ptr, const int* col, const c2* val,
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

// dB straight from A and dA -- dAA (|AA| x J x 3, 549 KB per environment for config 1) is never materialised
// [src/pibasis.jl:402-432 + src/symmbasis.jl:330-334 in one pass].  Written as a sparse matrix applied to the one-particle
// Jacobian:
//     dB[j][row][xyz][c] = Re( sum_a W_e[row][a][c] * dA[j][a][xyz] ),
//     W_e[row][a][c] = sum_{k in row, t : v_t(k) = a}  A2B[row,k][c] * prod_{s != t} A_e[v_s(k)]       (Re(A2B) if the PI basis is real)
// The sparsity pattern of W (entries = distinct (row, a), contributions = (k, t)) is static and packed on the host (DbPack);
// its values depend on the environment only, not on the neighbour, so they are computed once per environment and row tile
// (phase W, one thread per entry) and then applied to all neighbours (phase D): one CTA per environment, lane = neighbour,
// warp = group of RW consecutive rows.  dA of the <= 32 neighbours of a pass lives in shared memory as planes [slot][xyz][lane]
// over the canonical slots (k_dA writes that layout, in 32-neighbour tiles of the chunk); the weight of an entry is one broadcast LDS.128.  A warp stages the results of its group in a private
// strip of shared memory and writes them out itself, several neighbours' contiguous (RW x 3 x NC) runs per instruction, so
// phase D needs no CTA barrier: the warps of a tile run decoupled, groups dealt longest-first.
// Bound: shared-memory bandwidth -- 48 B of dA per lane and entry for 6 NC DFMAs.
struct DbEnvParams {
    long long nenv; const long long* off; const int* gate;
    int nA, nS, nB, nT, pireal, ET;
    const int* tile_grp;      // [nT + 1] position of each row tile's first group in grp_list
    const int* tile_ent;      // [nT + 1] first entry of each row tile (<= ET entries per tile)
    const int4* grp_list;     // (first row, first entry, end of row 0, end of row 1) of each group, entries relative to the tile;
                              // dealt longest-first within a tile (warp w takes positions w, w + nwarps, ...)
    const int4* ent_rec;      // [nE] (plane offset of the entry's canonical slot, first contribution, end of contributions, 0)
    const int4* con_rec;      // [nC] (4 * (non-zero of A2Bmap) + the neg / odd bits of the A-code of the differentiated factor,
                              //       the up to three other factors of the product (index into A), -1 = none)
    const c2* val; const c2* A; const c2* dA; double* dB;
};
constexpr int kDbPitch = 32;
ACE_HD constexpr int db_threads(int NC) { return NC == 1 ? 1024 : NC == 3 ? 512 : 256; }   // one CTA per SM (the dA planes fill shared memory): as many warps as the registers allow
ACE_HD constexpr int db_rows(int NC) { return NC == 1 ? 2 : 1; }           // rows per group
ACE_HD constexpr int db_strip(int NC) { return (db_rows(NC) * 3 * NC) | 1; }   // doubles per lane in a warp's staging strip
inline size_t db_env_fixed_smem(int nA, int nS, int NC, int threads)
{
    return ((size_t)nS * 3 * kDbPitch + nA) * sizeof(c2) + (size_t)(threads / 32) * 32 * db_strip(NC) * sizeof(double);
}
inline size_t db_env_smem(int nA, int nS, int ET, int NC, int threads) { return db_env_fixed_smem(nA, nS, NC, threads) + (size_t)ET * (NC * sizeof(c2) + sizeof(int)); }

template <int NC>
__global__ void __launch_bounds__(db_threads(NC), 1) k_dB_env(const DbEnvParams p)
{
    constexpr int RW = db_rows(NC), L = RW * 3 * NC, SP = db_strip(NC), JPER = 32 / L;
    ACE_DYN_SMEM(c2, planes);                                   // [nS * 3][kDbPitch]
    const long long e = blockIdx.x;
    if (e >= p.nenv) return;
    if (p.gate && *p.gate == 1) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    c2* As = planes + (size_t)p.nS * 3 * kDbPitch;              // [nA]
    c2* Ws = As + p.nA;                                         // [ET][NC]  (m0, m1): Re(w dA) = m0 Re dA_slot + m1 Im dA_slot
    double* strip = reinterpret_cast<double*>(Ws + (size_t)p.ET * NC) + (size_t)warp * 32 * SP;    // [32][SP], private to the warp
    int* Ea = reinterpret_cast<int*>(reinterpret_cast<double*>(Ws + (size_t)p.ET * NC) + (size_t)nwarps * 32 * SP);   // [ET] plane offset of the entry
    const int nS3 = p.nS * 3;
    const long long j0 = p.off[e] - p.off[0];
    const int J = (int)(p.off[e + 1] - p.off[e]);
    for (int a = tid; a < p.nA; a += blockDim.x) As[a] = p.A[(size_t)e * p.nA + a];
    for (int jt = 0; jt < J; jt += 32) {
        const int nj = (J - jt < 32) ? (J - jt) : 32;
        __syncthreads();                                         // the planes of the previous pass are no longer read
        {   // k_dA wrote dA in 32-neighbour tiles of the chunk: a plane row is one or two contiguous runs
            const long long gj = j0 + jt + lane;
            const c2* src = p.dA + (size_t)(gj >> 5) * nS3 * 32 + (gj & 31);
#pragma unroll 4
            for (int x = warp; x < nS3; x += nwarps) planes[x * kDbPitch + lane] = lane < nj ? src[(size_t)x * 32] : c2{0.0, 0.0};
        }
        for (int t = 0; t < p.nT; ++t) {
            const int g0 = __ldg(p.tile_grp + t), g1 = __ldg(p.tile_grp + t + 1);
            const int e0 = __ldg(p.tile_ent + t), e1 = __ldg(p.tile_ent + t + 1);
            __syncthreads();                                     // planes / A visible; the weights of the previous tile consumed
            // phase W: one thread per entry
            for (int i = e0 + tid; i < e1; i += blockDim.x) {
                double m0[NC], m1[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) { m0[c] = 0.0; m1[c] = 0.0; }
                const int4 er = __ldg(p.ent_rec + i);
                for (int q = er.y; q < er.z; ++q) {
                    const int4 cr = __ldg(p.con_rec + q);
                    const c2* v = p.val + (size_t)(cr.x >> 2) * NC;
                    c2 g = c2{1.0, 0.0};
                    if (cr.y >= 0) g = As[cr.y];
                    if (cr.z >= 0) g = cmul(g, As[cr.z]);
                    if (cr.w >= 0) g = cmul(g, As[cr.w]);
                    // the differentiated factor is decode_A(slot value): Re flips for odd negative m, Im for even negative m
                    const double sx = (cr.x & 2) ? -1.0 : 1.0, sy = ((cr.x & 1) && !(cr.x & 2)) ? -1.0 : 1.0;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const c2 vc = v[c];
                        double wr, wi;
                        if (p.pireal) { wr = vc.x * g.x; wi = vc.x * g.y; }
                        else { wr = vc.x * g.x - vc.y * g.y; wi = vc.x * g.y + vc.y * g.x; }
                        m0[c] += sx * wr; m1[c] -= sy * wi;
                    }
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) Ws[(size_t)(i - e0) * NC + c] = c2{m0[c], m1[c]};
                Ea[i - e0] = er.x;
            }
            __syncthreads();
            // phase D: one warp per group of RW rows, one lane per neighbour; no CTA barrier inside
            int4 gd_next = g0 + warp < g1 ? __ldg(p.grp_list + g0 + warp) : int4{0, 0, 0, 0};
            for (int pos = g0 + warp; pos < g1; pos += nwarps) {
                const int4 gd = gd_next;
                if (pos + nwarps < g1) gd_next = __ldg(p.grp_list + pos + nwarps);     // the next descriptor arrives during this group
                const int r0 = gd.x;
                const int nr = (p.nB - r0 < RW) ? (p.nB - r0) : RW;
#pragma unroll
                for (int rr = 0; rr < RW; ++rr) {
                    if (rr < nr) {
                        double acc[3][NC];
#pragma unroll
                        for (int d = 0; d < 3; ++d)
#pragma unroll
                            for (int c = 0; c < NC; ++c) acc[d][c] = 0.0;
                        const int i1 = rr == 0 ? gd.z : gd.w;
#pragma unroll 2
                        for (int i = rr == 0 ? gd.y : gd.z; i < i1; ++i) {
                            const c2* pl = planes + Ea[i] + lane;
                            const c2 x0 = pl[0], x1 = pl[kDbPitch], x2 = pl[2 * kDbPitch];
                            const c2* wr = Ws + (size_t)i * NC;
#pragma unroll
                            for (int c = 0; c < NC; ++c) {
                                const c2 w = wr[c];
                                acc[0][c] = fma(w.x, x0.x, fma(w.y, x0.y, acc[0][c]));
                                acc[1][c] = fma(w.x, x1.x, fma(w.y, x1.y, acc[1][c]));
                                acc[2][c] = fma(w.x, x2.x, fma(w.y, x2.y, acc[2][c]));
                            }
                        }
                        double* o = strip + lane * SP + rr * 3 * NC;
#pragma unroll
                        for (int d = 0; d < 3; ++d)
#pragma unroll
                            for (int c = 0; c < NC; ++c) o[d * NC + c] = acc[d][c];
                    }
                }
                __syncwarp();
                // JPER neighbours' contiguous runs of (nr x 3 x NC) doubles per store instruction
                const int len = nr * 3 * NC, jj = lane / L, x = lane - jj * L;
                if (jj < JPER && x < len) {
                    const size_t dstep = (size_t)JPER * p.nB * 3 * NC;
                    double* dst = p.dB + ((size_t)(j0 + jt + jj) * p.nB + r0) * 3 * NC + x;
                    const double* src = strip + jj * SP + x;
#pragma unroll 4
                    for (int jb = jj; jb < nj; jb += JPER, dst += dstep, src += JPER * SP) *dst = *src;
                }
                __syncwarp();
            }
        }
    }
}




// ------------------------------------------------------------------------------------------------
// caller side of the path: an atomic structure + neighbour list in, site energies and
// atomic forces out.  In the reference ecosystem this loop lives in JuLIP / ACEatoms.jl
// (forces(V, at): for each centre i, dV = evaluate_d(V, Rs); frc[j] -= dV_j; frc[i] += dV_j), around the
// per-environment calls of ACE.jl.  Doing it on the device removes the 24 B/pair each way that the
// per-environment interface has to move over PCIe.
// ------------------------------------------------------------------------------------------------
constexpr int kPairAtoms = 32;      // centres per CTA of k_build_pairs

struct CellDev { double c[9]; };

// R_p = X[nbr_p] + S_p . cell - X[i(p)],  species_p = species[nbr_p]   for the pairs of centres [a0, a0 + na)
// Packed neighbour words (aceb200_structure.flags & ACEB200_NBR_PACKED): bits 0..25 the neighbour index, bits 26..31 the
// image shift, two bits per component holding S + 1 (S in {-1, 0, 1}): 4 bytes per pair over PCIe instead of 7.
constexpr unsigned kNbrMask = 0x03ffffffu;
__device__ __forceinline__ long long nbr_index(int w, int packed) { return packed ? (long long)((unsigned)w & kNbrMask) : (long long)w; }
__device__ __forceinline__ int nbr_shift(int w, int k) { return (int)(((unsigned)w >> (26 + 2 * k)) & 3u) - 1; }

static __global__ void k_build_pairs(long long a0, long long na, long long natoms, const long long* first, const int* nbr,
                                     const signed char* image, int packed, const CellDev cell, const double* X, const int* spc, double* R, int* sp,
                                     int* errflag)
{
    ACE_DYN_SMEM(long long, f);                 // [kPairAtoms + 1]
    const long long c0 = a0 + (long long)blockIdx.x * kPairAtoms;
    if (c0 >= a0 + na) return;
    const int nc = (int)((a0 + na - c0) < kPairAtoms ? (a0 + na - c0) : kPairAtoms);
    if ((int)threadIdx.x <= nc) f[threadIdx.x] = first[c0 + threadIdx.x];
    __syncthreads();
    for (long long p = f[0] + threadIdx.x; p < f[nc]; p += blockDim.x) {
        int lo = 0, hi = nc;                       // centre of pair p: the last c with f[c] <= p
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (f[mid] <= p) lo = mid; else hi = mid; }
        const long long i = c0 + lo;
        const int wj = nbr[p];
        long long j = nbr_index(wj, packed);
        if (j < 0 || j >= natoms) { atomicMax(errflag, 1); j = i; }
        double x = X[3 * j] - X[3 * i], y = X[3 * j + 1] - X[3 * i + 1], z = X[3 * j + 2] - X[3 * i + 2];
        if (image || packed) {
            const double s0 = packed ? (double)nbr_shift(wj, 0) : (double)image[3 * p];
            const double s1 = packed ? (double)nbr_shift(wj, 1) : (double)image[3 * p + 1];
            const double s2 = packed ? (double)nbr_shift(wj, 2) : (double)image[3 * p + 2];
            x += s0 * cell.c[0] + s1 * cell.c[3] + s2 * cell.c[6];
            y += s0 * cell.c[1] + s1 * cell.c[4] + s2 * cell.c[7];
            z += s0 * cell.c[2] + s1 * cell.c[5] + s2 * cell.c[8];
        }
        R[3 * p] = x; R[3 * p + 1] = y; R[3 * p + 2] = z;
        if (sp) sp[p] = spc[j];
    }
}

// rev[p] for the pairs of 32 centres per CTA: scan the pair list of j = nbr[p] for (neighbour i, image -S).  The table is
// symmetric, so only one pair of each (p, rev p) couple searches -- the one whose (centre, image) is smaller than its
// reverse's -- and writes both entries; rev is pre-set to -1 (no reverse pair) by the caller.
static __global__ void k_find_rev(long long natoms, const long long* first, const int* nbr, const signed char* image, int packed, int* rev)
{
    ACE_DYN_SMEM(long long, f);                 // [kPairAtoms + 1]
    const long long c0 = (long long)blockIdx.x * kPairAtoms;
    if (c0 >= natoms) return;
    const int nc = (int)((natoms - c0) < kPairAtoms ? (natoms - c0) : kPairAtoms);
    if ((int)threadIdx.x <= nc) f[threadIdx.x] = first[c0 + threadIdx.x];
    __syncthreads();
    for (long long p = f[0] + threadIdx.x; p < f[nc]; p += blockDim.x) {
        int lo = 0, hi = nc;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (f[mid] <= p) lo = mid; else hi = mid; }
        const int i = (int)(c0 + lo);
        const int wj = __ldg(nbr + p);
        const long long j = nbr_index(wj, packed);
        if (j < 0 || j >= natoms) continue;
        int s0 = 0, s1 = 0, s2 = 0;             // the reverse pair's image
        if (image) { s0 = -__ldg(image + 3 * p); s1 = -__ldg(image + 3 * p + 1); s2 = -__ldg(image + 3 * p + 2); }
        if (packed) { s0 = -nbr_shift(wj, 0); s1 = -nbr_shift(wj, 1); s2 = -nbr_shift(wj, 2); }
        // packed: the reverse pair is one exact word (neighbour i, shift -S)
        const int wrev = (int)((unsigned)i | ((unsigned)(s0 + 1) << 26) | ((unsigned)(s1 + 1) << 28) | ((unsigned)(s2 + 1) << 30));
        // who searches: the pair with the smaller centre; for a self-image pair (i == j) the one whose image is
        // lexicographically smaller than its reverse's (a pair that is its own reverse cannot occur: S = 0 means r = 0)
        if (j < i) continue;
        if (j == i) {
            const int t0 = -s0, t1 = -s1, t2 = -s2;     // this pair's own image
            const bool smaller = t0 != s0 ? t0 < s0 : (t1 != s1 ? t1 < s1 : t2 < s2);
            if (!smaller) continue;
        }
        int found = -1;
        // neighbour lists are normally sorted by neighbour index within a centre: binary search for the first
        // entry >= i and walk the (few) images of i; an unsorted list falls back to the linear scan below
        long long lo2 = __ldg(first + j), hi2 = __ldg(first + j + 1);
        const long long qb = lo2, qe = hi2;
        while (lo2 < hi2) { const long long mid = (lo2 + hi2) >> 1; if (nbr_index(__ldg(nbr + mid), packed) < i) lo2 = mid + 1; else hi2 = mid; }
        for (long long q = lo2; q < qe && nbr_index(__ldg(nbr + q), packed) == i; ++q) {
            if (image && (__ldg(image + 3 * q) != s0 || __ldg(image + 3 * q + 1) != s1 || __ldg(image + 3 * q + 2) != s2)) continue;
            if (packed && __ldg(nbr + q) != wrev) continue;
            found = (int)q;
            break;
        }
        if (found < 0) {
            for (long long q = qb; q < qe; ++q) {
                if (nbr_index(__ldg(nbr + q), packed) != i) continue;
                if (image && (__ldg(image + 3 * q) != s0 || __ldg(image + 3 * q + 1) != s1 || __ldg(image + 3 * q + 2) != s2)) continue;
                if (packed && __ldg(nbr + q) != wrev) continue;
                found = (int)q;
                break;
            }
        }
        if (found >= 0) { rev[p] = found; rev[found] = (int)p; }
    }
}

// F_i[k] = sum_{p in env(i)} ( g_p[k] - g_{rev(p)}[k] ),  k over the K = nprop * 3 * ncomp components of a pair
// gradient: a gather over the reverse-pair table, no atomics, fixed summation order.
static __global__ void k_assemble_rev(long long natoms, int K, const long long* first, const int* rev, const double* G,
                                      double* F, int* errflag, long long npairs)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= natoms * K) return;
    const long long i = t / K;
    const int k = (int)(t - i * K);
    double acc = 0.0;
    for (long long p = first[i]; p < first[i + 1]; ++p) {
        const long long q = rev[p];
        double v = G[(size_t)p * K + k];
        if (q >= 0 && q < npairs) v -= G[(size_t)q * K + k];
        else if (q >= npairs) atomicMax(errflag, 1);
        acc += v;                                   // q < 0: the neighbour is not a centre (no reverse pair)
    }
    F[t] = acc;
}

// virial  W[prop][a][b] = - sum_p g_p[prop][a] R_p[b]  (JuLIP: site_virial = -sum dV_j (x) R_j), two deterministic stages
constexpr int kVirThreads = 256;
static __global__ void k_virial_partial(long long npairs, int nprop, const double* G, const double* R, double* part)
{
    ACE_DYN_SMEM(double, red);                  // [kVirThreads]
    const int prop = blockIdx.y;
    double w[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
        const double* g = G + ((size_t)p * nprop + prop) * 3;
        const double* r = R + (size_t)p * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) w[a * 3 + b] -= g[a] * r[b];
    }
    for (int c = 0; c < 9; ++c) {
        red[threadIdx.x] = w[c];
        __syncthreads();
        for (int s = kVirThreads / 2; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) part[((size_t)prop * gridDim.x + blockIdx.x) * 9 + c] = red[0];
        __syncthreads();
    }
}

static __global__ void k_virial_final(int nblocks, const double* part, double* W)
{
    const int prop = blockIdx.x, c = threadIdx.x;
    if (c >= 9) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += part[((size_t)prop * nblocks + b) * 9 + c];
    W[prop * 9 + c] = acc;
}

}  // namespace aceb200
