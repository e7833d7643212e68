// aceb200.cu -- the C ABI of include/aceb200.h on top of the kernels in ace_kernels.cuh.
//
// Host-side responsibilities: validate and flatten the descriptor (ace_tables.h), upload tables,
// keep c~ and the tree weights in sync with set_params, split a batch into chunks that fit the
// workspace, move host batches to the device and results back, launch.  There is no CPU evaluation
// path in this file: every entry point needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <map>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <thread>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include "ace_platform.cuh"
#include "ace_kernels.cuh"
#include "ace_tables.h"
#include "ace_launch.h"

using namespace aceb200;

// ----------------------------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local int g_device = 0;

static int fail(int code, const std::string& msg) { g_err = msg; return code; }

// CU(call): throw ModelError on a CUDA error (ace_launch.h)

// ----------------------------------------------------------------------------------------------
// device buffers
// ----------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    void reserve(size_t n)
    {
        if (n <= bytes) return;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        CU(cudaMalloc(&p, n));
        bytes = n;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

template <class T>
static T* upload(std::vector<DevBuf>& pool, const std::vector<T>& v)
{
    pool.emplace_back();
    DevBuf& b = pool.back();
    b.reserve(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (!v.empty()) CU(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return b.as<T>();
}

// ----------------------------------------------------------------------------------------------
// the model handle
// ----------------------------------------------------------------------------------------------
constexpr int kLanes = 4;

struct Lane {
    DevBuf ws_Ac, ws_Dt, ws_E, ws_G, ws_A, ws_AA, ws_dA, ws_dAA, ws_out;
    DevBuf in_off, in_R, in_sp, in_w, ws_Aw, ws_A2;
    cudaStream_t stream = nullptr;       // the stream this lane currently launches on
    cudaStream_t own_stream = nullptr;   // private non-blocking stream (host-batch pipeline)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr;
    bool busy = false;                   // has un-synchronised work with pending timing events
    bool timed_ef = false;
    void release()
    {
        DevBuf* bufs[] = {&ws_Ac, &ws_Dt, &ws_E, &ws_G, &ws_A, &ws_AA, &ws_dA, &ws_dAA, &ws_out, &in_off, &in_R, &in_sp, &in_w, &ws_Aw, &ws_A2};
        for (DevBuf* b : bufs) b->release();
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (evA) cudaEventDestroy(evA);
        if (evB) cudaEventDestroy(evB);
        if (own_stream) cudaStreamDestroy(own_stream);
    }
};

// Per-call state.  Evaluation calls on one handle may run concurrently from several host threads (the reference's
// per-thread pools, src/utils/pools.jl:44-75): each call leases a context -- workspaces, pipeline lanes with their
// streams and events, an error flag -- from the handle's pool (created on demand) and returns it when it is done.
// The device tables are shared and read-only; set_params takes the handle exclusively.
struct Ctx {
    Lane lanes[kLanes];
    DevBuf ws_err;             // device int: EEMPTY / ECATEGORY / EDESC raised by kernels; zero between calls
    int* h_err = nullptr;      // pinned host copy
    bool in_use = false;
    void create()
    {
        for (Lane& L : lanes) {
            CU(cudaEventCreate(&L.ev0)); CU(cudaEventCreate(&L.ev1)); CU(cudaEventCreate(&L.evA)); CU(cudaEventCreate(&L.evB));
            CU(cudaStreamCreateWithFlags(&L.own_stream, cudaStreamNonBlocking));
        }
        ws_err.reserve(sizeof(int));
        CU(cudaMemset(ws_err.p, 0, sizeof(int)));
        CU(cudaMallocHost((void**)&h_err, sizeof(int)));
        *h_err = 0;
    }
    void release()
    {
        for (Lane& L : lanes) L.release();
        ws_err.release();
        if (h_err) cudaFreeHost(h_err);
        h_err = nullptr;
    }
};
static thread_local Ctx* t_ctx = nullptr;     // the context the calling thread holds
static thread_local Lane* t_cur = nullptr;    // the lane the calling thread is launching on

struct StreamPass { DevBuf blocks, tinfo, w0; int pb0 = 0; };

// A packed leaf stream for k_basis_stream and its geometry
struct BStream {
    DevBuf buf;
    int nw = 0, nchunks = 0, nfac = 0, nch = 0, LB = 0, LPC = 0, W = 0, epl = 1, nrows = 0;
    bool cw = false;
    int nleaf[kBasisMaxWarps] = {0}, row0[kBasisMaxWarps] = {0};
    size_t smem = 0;
};
struct BLeaf { unsigned c0, c1; std::vector<double> w; };     // slot codes; w: [nch][cs] = (p, -q) pairs

struct aceb200_model {
    int device = 0;
    HostTables T;
    RadialParams rp;
    AlpParams ap;
    std::vector<cplx> ctilde;       // [nAA][P]
    bool cw = false;                // complex tree weights
    int PB = 1, Ppad = 1, NMAX = 8;
    // device tables
    std::vector<DevBuf> pool;
    ColumnsDev C;
    const int *d_slot_pos = nullptr, *d_slot_neg = nullptr, *d_code = nullptr;
    const int4* d_pool_blk = nullptr;
    int n_pool_blk = 0;
    const PoolTile* d_pool_tiles = nullptr;    // k_pool_mma column tiles (single-species models)
    int n_pool_tiles = 0;
    const ForceTile* d_force_tiles = nullptr;  // k_forces_mma column tiles (same column order)
    // k_basis_stream (fused B = A2Bmap . prod A): leaf stream per warp; depends on the tables only, not on c
    BStream bs;                                // B = A2Bmap . AA
    // k_dB_env (dB = W_e . dA): static sparsity pattern of W; depends on the tables only
    struct DbPack { bool ok = false; int nT = 0, ET = 0, threads = 0; size_t smem = 0;
                    const int *tile_grp = nullptr, *tile_ent = nullptr; const int4 *grp_list = nullptr, *ent_rec = nullptr, *con_rec = nullptr; } db;
    const int *d_orders = nullptr, *d_spec = nullptr;
    const int *d_csr_ptr = nullptr, *d_csr_col = nullptr;
    const c2* d_csr_val = nullptr;
    ListDev list[kMaxOrdDev + 1];
    DevBuf d_w0, d_w1;
    DevBuf d_lw[kMaxOrdDev + 1];
    DevBuf d_ctl;                      // k_adjoint_stream control words (shared by all passes)
    std::vector<StreamPass> passes;    // per-pass leaf blocks and target records (PB channels each)
    int stream_nblk[kStreamWarps] = {0};
    // energy-only stream (aceb200_energy): every AA function once, under its first factor (see upload_stream)
    DevBuf e_ctl;
    std::vector<StreamPass> e_passes;
    int e_stream_nblk[kStreamWarps] = {0};
    int e_stream_chunks = 0, e_stream_ntinfo = 0;
    int stream_chunks = 0, stream_nf = 0, stream_ntinfo = 0, stream_pb = 1, stream_epl = 1;   // 0 chunks: use the generic k_adjoint
    // per-call workspace (guarded by mu).  Three lanes: device-resident batches use lane 0 on the caller's
    // stream; host-resident batches are pipelined chunk by chunk over all lanes (H2D copy, kernels and D2H
    // copy of consecutive chunks overlap on the lanes' private streams).
    std::shared_mutex params_mu;           // evaluation: shared; set_params: exclusive
    std::mutex ctx_mu;
    std::condition_variable ctx_cv;
    std::vector<std::unique_ptr<Ctx>> ctxs;
    std::mutex stat_mu;                    // last_ms / stage_ms
    std::vector<double> c_host;            // the coefficients (for replicas on other devices)
    std::vector<aceb200_model*> replicas;  // aceb200_set_devices: the same model on further GPUs
    std::vector<int> devices;              // device of this handle followed by the replicas' 
    // structure path (aceb200_structure_energy_forces): inputs are copied on copy_stream, chunk by chunk, while
    // the evaluation of earlier chunks runs on the caller's stream
    std::mutex mu_s;
    DevBuf s_X, s_first, s_nbr, s_img, s_spc, s_rev, s_R, s_sp, s_G, s_E, s_F, s_W, s_part, s_err;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> s_ev;
    cudaStream_t user_stream = nullptr;
    double last_ms = 0.0;
    double stage_ms[3] = {0.0, 0.0, 0.0};   // pool, adjoint, forces of the last energy(_forces) call
    std::atomic<long long> launches{0};
    int sm_count = 148;
    int smem_optin = 227 * 1024;
};

static int pick_nmax(int n)
{
    const int opts[] = {4, 8, 12, 16, 20, 24, 32};
    for (int o : opts) if (n <= o) return o;
    return 32;
}

static int pick_pb(int P)
{
    if (P == 1) return 1;
    if (P <= 2) return 2;
    if (P == 3 || P == 9) return 3;
    return 4;
}

static void fill_params(aceb200_model* m, const aceb200_desc& d)
{
    RadialParams& rp = m->rp;
    memset(&rp, 0, sizeof(rp));
    rp.N = d.n_rad; rp.pl = d.pl; rp.pr = d.pr; rp.tkind = d.trans_kind; rp.tl = d.tl; rp.tr = d.tr;
    for (int i = 0; i < 4; ++i) rp.tpar[i] = d.trans_par[i];
    for (int i = 0; i < d.n_rad; ++i) { rp.A[i] = d.rad_A[i]; rp.B[i] = d.rad_B[i]; rp.C[i] = d.rad_C[i]; }
    AlpParams& ap = m->ap;
    memset(&ap, 0, sizeof(ap));
    ap.L = m->T.Lused;
    // src/polynomials/sphericalharmonics.jl:146-159
    for (int l = 2; l <= ap.L; ++l)
        for (int mm = 0; mm <= l - 2; ++mm) {
            double ls = (double)l * l, lm1s = (double)(l - 1) * (l - 1), ms = (double)mm * mm;
            ap.A[index_p(l, mm)] = sqrt((4 * ls - 1.0) / (ls - ms));
            ap.B[index_p(l, mm)] = -sqrt((lm1s - ms) / (4 * lm1s - 1.0));
        }
    ap.diagc[0] = 0.0;
    ap.diagc[1] = sqrt(1.5);                                   // :180
    for (int mm = 2; mm <= kMaxL + 1; ++mm) ap.diagc[mm] = sqrt(1.0 + 0.5 / mm);   // :192
    for (int mm = 0; mm + 1 <= ap.L; ++mm) { ap.A[index_p(mm + 1, mm)] = sqrt(2.0 * mm + 3.0); ap.B[index_p(mm + 1, mm)] = 0.0; }   // :179, :191
}

static void upload_stream(aceb200_model* m, bool energy_only = false);

// c~ and the weights that depend on it: order-0/1 weights and the leaf weights of every tree
static void upload_weights(aceb200_model* m, const double* c)
{
    HostTables& T = m->T;
    eff_coeffs(T, c, m->ctilde);
    bool cw = false;
    for (const cplx& z : m->ctilde) if (z.imag() != 0.0) { cw = true; break; }
    if (cw && T.pireal)
        throw ModelError(ACEB200_EUNSUPPORTED, "complex effective coefficients with a real AA basis (pireal) are not supported");
    m->cw = cw;
    const int P = T.P, Ppad = m->Ppad, cs = cw ? 2 : 1;
    auto put = [&](std::vector<double>& w, size_t row, int aa, double mult) {
        for (int p = 0; p < P; ++p) {
            cplx z = m->ctilde[(size_t)aa * P + p] * mult;
            w[(row * Ppad + p) * cs] = z.real();
            if (cw) w[(row * Ppad + p) * cs + 1] = z.imag();
        }
    };
    std::vector<double> w0((size_t)Ppad * cs, 0.0), w1((size_t)T.nA * Ppad * cs, 0.0);
    if (T.has_const) put(w0, 0, 0, 1.0);
    for (int a = 0; a < T.nA; ++a) if (T.aa1_of_target[a] >= 0) put(w1, a, T.aa1_of_target[a], 1.0);
    m->d_w0.reserve(w0.size() * sizeof(double));
    m->d_w1.reserve(w1.size() * sizeof(double));
    CU(cudaMemcpy(m->d_w0.p, w0.data(), w0.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m->d_w1.p, w1.data(), w1.size() * sizeof(double), cudaMemcpyHostToDevice));
    for (int nu = 2; nu <= T.maxord; ++nu) {
        const Tree& tr = T.trees[nu];
        const size_t nleaf = tr.laa.size();
        const int stride = 8 + 8 * Ppad * cs;
        std::vector<unsigned char> rec(std::max<size_t>(nleaf, 1) * stride, 0);
        std::vector<double> w((size_t)Ppad * cs);
        for (size_t i = 0; i < nleaf; ++i) {
            std::fill(w.begin(), w.end(), 0.0);
            put(w, 0, tr.laa[i], (double)tr.lmult[i]);
            memcpy(&rec[i * stride], &tr.codes[4 * i], 8);
            memcpy(&rec[i * stride + 8], w.data(), w.size() * sizeof(double));
        }
        m->d_lw[nu].reserve(rec.size());
        CU(cudaMemcpy(m->d_lw[nu].p, rec.data(), rec.size(), cudaMemcpyHostToDevice));
        m->list[nu].rec = m->d_lw[nu].as<unsigned char>();
        m->list[nu].stride = stride;
    }
    upload_stream(m);
    upload_stream(m, true);
}

// Host mirror of StreamGeom (ace_kernels.cuh)
struct HostGeom { int CS, CWORDS, QB, KB, QBP, CH, TIQ, NW; };
static HostGeom stream_geom(int NF, int PB, bool CW)
{
    HostGeom g;
    g.CS = CW ? 2 : 1;
    g.CWORDS = (NF == 2) ? 1 : 2;
    g.QB = g.CWORDS + 2 * PB * g.CS;
    g.NW = (PB == 1 && !CW) ? 12 : 8;
    g.KB = (g.QB <= 4) ? 16 : 8;
    g.QBP = ((g.QB + 3) / 4) * 4;
    g.CH = g.KB * g.QBP;
    g.TIQ = 1 + (1 + PB * g.CS + 1) / 2;
    return g;
}

static size_t stream_smem(int nS, const HostGeom& g, int pb, int epl = 1)
{
    return (size_t)(nS + 1) * 32 * epl * sizeof(c2)                // A tile (+ the row of ones)
         + (size_t)g.NW * 2 * g.CH * sizeof(uint4)                  // per-warp stream rings
         + (size_t)g.NW * pb * epl * 32 * sizeof(double)            // per-warp energy partials
         + (size_t)(1 + 2 * g.NW) * 8;                              // mbarriers
}

// Flatten the adjoint lists into the streams k_adjoint_stream consumes (layout: ace_kernels.cuh).
//
// Mirror folding.  With A_{n l -m} = (-1)^m conj(A_{n l m}) the AA function whose factors all have their m
// negated equals s conj(AA), s = (-1)^{sum m}.  Under the real part that every output of this path takes,
//     Re(c~_i AA_i) + Re(c~_i' AA_i') = Re((c~_i + s conj(c~_i')) AA_i),
// so only one function of each mirror pair is kept, with the folded weight.  This halves the stream; energy and
// gradient are unchanged as functions of the positions (only the order of floating-point additions differs).
//
// Energy-only stream (energy_only = true).  evaluate(model, cfg) needs no adjoints, so the Euler pass over every target
// (each AA function visited once per distinct factor) is nu times more work than necessary.  The same kernel walks a
// second stream in which an AA function appears exactly once -- under its first factor, with multiplicity 1 and
// segment scale 1 -- so that  E += Re(A_a * w * prod(other factors)) = w Re(AA)  [src/evaluator.jl:137-143].
static void upload_stream(aceb200_model* m, bool energy_only)
{
    HostTables& T = m->T;
    std::vector<StreamPass>& passes = energy_only ? m->e_passes : m->passes;
    DevBuf& d_ctl = energy_only ? m->e_ctl : m->d_ctl;
    int* stream_nblk = energy_only ? m->e_stream_nblk : m->stream_nblk;
    int& stream_chunks = energy_only ? m->e_stream_chunks : m->stream_chunks;
    int& stream_ntinfo = energy_only ? m->e_stream_ntinfo : m->stream_ntinfo;
    stream_chunks = 0;
    for (StreamPass& sp : passes) { sp.blocks.release(); sp.tinfo.release(); sp.w0.release(); }
    passes.clear();
    if (energy_only && (m->stream_chunks == 0 || getenv("ACEB200_NO_ENERGY_STREAM"))) return;
    if (T.maxord < 2 || T.maxord > 5 || !T.symreal || getenv("ACEB200_NO_STREAM")) return;
    if (T.nS + 1 >= (1 << 14)) return;                   // 14-bit slot fields
    const int NF = T.maxord <= 3 ? 2 : (T.maxord == 4 ? 3 : 4);
    const bool CW = m->cw;
    const int P = T.P;
    // channels per pass: the largest power of two <= 8 whose ring still fits next to the A tile
    int PB = 1;
    while (PB < P && PB < 8) PB *= 2;
    while (PB > 1 && stream_smem(T.nS, stream_geom(NF, PB, CW), PB) > (size_t)m->smem_optin) PB /= 2;
    const HostGeom g = stream_geom(NF, PB, CW);
    if (stream_smem(T.nS, g, PB) > (size_t)m->smem_optin) return;    // falls back to the list kernel
    const bool fold = !getenv("ACEB200_NO_MIRROR_FOLD");

    // ---- effective (mirror-folded) coefficients; non-canonical partners are dropped
    std::vector<cplx> ceff(m->ctilde);
    std::vector<char> keep(T.nAA, 1);
    if (fold) {
        std::map<std::tuple<int, int, int, int>, int> inv1p;
        for (int a = 0; a < T.nA; ++a) inv1p[std::make_tuple(T.iA_q[a], T.iA_n[a], T.iA_l[a], T.iA_m[a])] = a;
        std::vector<int> mirA(T.nA, -1);
        for (int a = 0; a < T.nA; ++a) {
            auto it = inv1p.find(std::make_tuple(T.iA_q[a], T.iA_n[a], T.iA_l[a], -T.iA_m[a]));
            if (it != inv1p.end()) mirA[a] = it->second;
        }
        std::map<std::vector<int>, int> invAA;
        auto keyof = [&](int i, bool mirror, bool& ok) {
            std::vector<int> k;
            ok = true;
            for (int t = 0; t < T.orders[i]; ++t) {
                int a = T.spec[(size_t)i * T.maxord + t];
                if (mirror) { a = mirA[a]; if (a < 0) { ok = false; a = 0; } }
                k.push_back(a);
            }
            std::sort(k.begin(), k.end(), std::greater<int>());
            return k;
        };
        bool ok;
        for (int i = 0; i < T.nAA; ++i) invAA[keyof(i, false, ok)] = i;
        for (int i = 0; i < T.nAA; ++i) {
            if (!keep[i]) continue;
            std::vector<int> k = keyof(i, true, ok);
            if (!ok) continue;
            auto it = invAA.find(k);
            if (it == invAA.end() || it->second <= i) continue;     // no partner, self-mirror, or already handled
            const int ip = it->second;
            int summ = 0;
            for (int t = 0; t < T.orders[i]; ++t) summ += T.iA_m[T.spec[(size_t)i * T.maxord + t]];
            const double sg = (summ & 1) ? -1.0 : 1.0;
            for (int pch = 0; pch < P; ++pch) {
                ceff[(size_t)i * P + pch] += sg * std::conj(m->ctilde[(size_t)ip * P + pch]);
                ceff[(size_t)ip * P + pch] = cplx(0, 0);
            }
            keep[ip] = 0;
        }
    }

    const unsigned ONE = (unsigned)T.nS;                 // slot index of the constant 1
    struct Leaf { unsigned code, code2; double sg; int aa, mult; };
    auto make_leaf = [&](const uint16_t* codes, int nf, int aa, int mult) {
        unsigned slot[4] = {ONE, ONE, ONE, ONE}, cj[4] = {0u, 0u, 0u, 0u};
        double sg = 1.0;
        unsigned k1 = 0;
        for (int f = 0; f < nf; ++f) {
            const unsigned c = codes[f];
            const unsigned neg = c & 1u, odd = (c >> 1) & 1u;
            if (neg && odd) sg = -sg;                     // (-1)^m
            if (f == 0) k1 = neg;
            slot[f] = c >> 2;
            cj[f] = (f > 0 && (neg ^ k1)) ? 1u : 0u;
        }
        Leaf L;
        L.code = slot[0] | (k1 << 15) | (slot[1] << 16) | (cj[1] << 31);   // flipIm = conj of the whole product
        L.code2 = slot[2] | (cj[2] << 15) | (slot[3] << 16) | (cj[3] << 31);
        L.sg = sg; L.aa = aa; L.mult = mult;
        return L;
    };
    // grouped form: `first` is the A-code of the group's shared factor (-1: the slot of ones), `others` the rest
    const bool GR = (PB == 1 && !CW);
    auto make_leaf_gr = [&](int first, const uint16_t* others, int no, int aa, int mult, bool gend) {
        unsigned slot[4] = {ONE, ONE, ONE, ONE}, cj[4] = {0u, 0u, 0u, 0u};
        double sg = 1.0;
        auto put = [&](int f, unsigned c) {
            const unsigned neg = c & 1u, odd = (c >> 1) & 1u;
            if (neg && odd) sg = -sg;
            slot[f] = c >> 2; cj[f] = neg;
        };
        if (first >= 0) put(0, (unsigned)first);
        for (int f = 0; f < no; ++f) put(1 + f, others[f]);
        Leaf L;
        L.code = slot[0] | (gend ? kGroupEnd : 0u) | (cj[0] << 15) | (slot[1] << 16) | (cj[1] << 31);
        L.code2 = slot[2] | (cj[2] << 15) | (slot[3] << 16) | (cj[3] << 31);
        L.sg = sg; L.aa = aa; L.mult = mult;
        return L;
    };
    // two-level grouped form (NF == 3): group factor, sub-group factor (either may be absent: the slot of ones), leaf factor
    auto make_leaf_nested = [&](int gkey, int skey, unsigned leafc, int aa, int mult, bool gend, bool send) {
        unsigned slot[3] = {ONE, ONE, ONE}, cj[3] = {0u, 0u, 0u};
        double sg = 1.0;
        auto put = [&](int f, unsigned c) {
            const unsigned neg = c & 1u, odd = (c >> 1) & 1u;
            if (neg && odd) sg = -sg;
            slot[f] = c >> 2; cj[f] = neg;
        };
        if (gkey >= 0) put(0, (unsigned)gkey);
        if (skey >= 0) put(1, (unsigned)skey);
        put(2, leafc);
        Leaf L;
        L.code = slot[0] | (gend ? kGroupEnd : 0u) | (cj[0] << 15) | (slot[1] << 16) | (cj[1] << 31);
        L.code2 = slot[2] | (send ? kSubEnd : 0u) | (cj[2] << 15) | (ONE << 16);
        L.sg = sg; L.aa = aa; L.mult = mult;
        return L;
    };
    // The block structure (codes, ctl, which target a tinfo record belongs to) is the same for every pass;
    // only the weights differ.  Build the structure once per sub-stream, then emit the passes.
    // energy-only: an AA function belongs to the target that is its first factor
    auto mine = [&](int a, int aa) { return !energy_only || T.spec[(size_t)aa * T.maxord] == a; };
    struct BlockRef { Leaf L[kBlkLeaves]; int nleaf; unsigned flags; int target; double invnu; };
    std::vector<std::vector<BlockRef>> perslot(T.nS);
    auto add_block = [&](std::vector<BlockRef>& v, const std::vector<Leaf>* leaves, size_t i0, unsigned flags, int target, double invnu) {
        BlockRef b;
        b.nleaf = 0; b.flags = flags; b.target = target; b.invnu = invnu;
        for (int k = 0; k < kBlkLeaves; ++k) {
            if (leaves && i0 + k < leaves->size()) { b.L[k] = (*leaves)[i0 + k]; b.nleaf = k + 1; }
            else { b.L[k].code = ONE | (ONE << 16); b.L[k].code2 = ONE | (ONE << 16); b.L[k].sg = 0.0; b.L[k].aa = -1; b.L[k].mult = 0; }
        }
        v.push_back(b);
    };
    for (int s = 0; s < T.nS; ++s) {
        std::vector<BlockRef>& S = perslot[s];
        const unsigned sbits = (unsigned)s << 8;
        int tg[2] = {T.slot_pos[s], T.slot_neg[s]};
        int last = -1;
        for (int k = 0; k < 2; ++k) if (tg[k] >= 0) last = k;
        if (last < 0) { add_block(S, nullptr, 0, kSlotEnd | sbits, -1, 0.0); continue; }
        for (int k = 0; k < 2; ++k) {
            const int a = tg[k];
            if (a < 0) continue;
            unsigned tflags = kTgtEnd | (k == last ? (kSlotEnd | sbits) : 0u);
            if (T.iA_code[a] & 1) tflags |= kTgtNeg;
            if (T.iA_code[a] & 2) tflags |= kTgtOdd;
            std::vector<std::vector<Leaf>> per(T.maxord + 1);
            int lastnu = 0;
            for (int nu = 2; nu <= T.maxord; ++nu) {
                const Tree& tr = T.trees[nu];
                if (!GR) {
                    for (int i = tr.ptr[a]; i < tr.ptr[a + 1]; ++i) {
                        if (!keep[tr.laa[i]] || !mine(a, tr.laa[i])) continue;
                        per[nu].push_back(make_leaf(&tr.codes[4 * (size_t)i], nu - 1, tr.laa[i], energy_only ? 1 : tr.lmult[i]));
                    }
                } else {
                    // greedy grouping: repeatedly take the factor shared by the most remaining leaves
                    std::vector<int> rest;
                    for (int i = tr.ptr[a]; i < tr.ptr[a + 1]; ++i) if (keep[tr.laa[i]] && mine(a, tr.laa[i])) rest.push_back(i);
                    const int nf = nu - 1;
                    while (!rest.empty()) {
                        int key = -1;
                        std::vector<int> grp;
                        if (nf == 1) grp.swap(rest);          // order 2: one group behind the slot of ones
                        else {
                            std::map<int, int> cnt;
                            for (int i : rest) {
                                const uint16_t* cd = &tr.codes[4 * (size_t)i];
                                for (int f = 0; f < nf; ++f) {
                                    bool seen = false;
                                    for (int f2 = 0; f2 < f; ++f2) seen |= cd[f2] == cd[f];
                                    if (!seen) cnt[cd[f]]++;
                                }
                            }
                            int bestc = 0;
                            for (auto& kv : cnt) if (kv.second > bestc) { bestc = kv.second; key = kv.first; }
                            std::vector<int> keepv;
                            for (int i : rest) {
                                const uint16_t* cd = &tr.codes[4 * (size_t)i];
                                bool has = false;
                                for (int f = 0; f < nf; ++f) has |= cd[f] == key;
                                (has ? grp : keepv).push_back(i);
                            }
                            rest.swap(keepv);
                        }
                        // the factors of each leaf of the group other than (one occurrence of) the group key
                        std::vector<std::vector<uint16_t>> oth(grp.size());
                        for (size_t gi = 0; gi < grp.size(); ++gi) {
                            const uint16_t* cd = &tr.codes[4 * (size_t)grp[gi]];
                            bool taken = false;
                            for (int f = 0; f < nf; ++f) {
                                if (key >= 0 && !taken && cd[f] == key) { taken = true; continue; }
                                oth[gi].push_back(cd[f]);
                            }
                        }
                        if (NF == 3) {
                            // two-level form: sub-group the group's leaves by a second shared factor (greedy, as above)
                            std::vector<size_t> rest2(grp.size());
                            for (size_t gi = 0; gi < grp.size(); ++gi) rest2[gi] = gi;
                            while (!rest2.empty()) {
                                int skey = -1;
                                std::vector<size_t> sub, keep2;
                                if (oth[rest2[0]].size() < 2) sub.swap(rest2);          // order 2 / 3: a single factor left, one sub-group
                                else {
                                    std::map<int, int> cnt;
                                    for (size_t gi : rest2) { cnt[oth[gi][0]]++; if (oth[gi][1] != oth[gi][0]) cnt[oth[gi][1]]++; }
                                    int bestc = 0;
                                    for (auto& kv : cnt) if (kv.second > bestc) { bestc = kv.second; skey = kv.first; }
                                    for (size_t gi : rest2) ((oth[gi][0] == skey || oth[gi][1] == skey) ? sub : keep2).push_back(gi);
                                    rest2.swap(keep2);
                                }
                                for (size_t si = 0; si < sub.size(); ++si) {
                                    const size_t gi = sub[si];
                                    unsigned leafc = oth[gi][0];
                                    if (skey >= 0) leafc = (oth[gi][0] == skey) ? oth[gi][1] : oth[gi][0];
                                    const bool send = si + 1 == sub.size(), gend = send && rest2.empty();
                                    per[nu].push_back(make_leaf_nested(key, skey, leafc, tr.laa[grp[gi]], energy_only ? 1 : tr.lmult[grp[gi]], gend, send));
                                }
                            }
                        } else {
                            for (size_t gi = 0; gi < grp.size(); ++gi) {
                                uint16_t ord[4] = {0, 0, 0, 0};
                                for (size_t f = 0; f < oth[gi].size(); ++f) ord[f] = oth[gi][f];
                                per[nu].push_back(make_leaf_gr(key, ord, (int)oth[gi].size(), tr.laa[grp[gi]], energy_only ? 1 : tr.lmult[grp[gi]], gi + 1 == grp.size()));
                            }
                        }
                    }
                }
                if (!per[nu].empty()) lastnu = nu;
            }
            if (lastnu == 0) { add_block(S, nullptr, 0, tflags, a, 0.0); continue; }
            for (int nu = 2; nu <= T.maxord; ++nu) {
                const std::vector<Leaf>& lv = per[nu];
                for (size_t i = 0; i < lv.size(); i += kBlkLeaves) {
                    unsigned flags = 0u;
                    if (i + kBlkLeaves >= lv.size()) {
                        flags = (unsigned)nu | kSegEnd | (tflags & (kTgtNeg | kTgtOdd));
                        if (nu == lastnu) flags |= tflags;
                    }
                    add_block(S, &lv, i, flags, a, energy_only ? 1.0 : 1.0 / nu);
                }
            }
        }
    }
    // LPT split of the slots over the g.NW sub-streams
    std::vector<int> order(T.nS);
    for (int s = 0; s < T.nS; ++s) order[s] = s;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return perslot[x].size() > perslot[y].size(); });
    std::vector<std::vector<BlockRef>> sub(g.NW);
    for (int s : order) {
        int best = 0;
        for (int w = 1; w < g.NW; ++w) if (sub[w].size() < sub[best].size()) best = w;
        sub[best].insert(sub[best].end(), perslot[s].begin(), perslot[s].end());
    }
    size_t longest = 1, ntinfo = 1;
    for (auto& v : sub) {
        longest = std::max(longest, v.size());
        size_t nt = 0;
        for (auto& b : v) nt += b.flags != 0;
        ntinfo = std::max(ntinfo, nt);
    }
    const size_t nchunks = (longest + g.KB - 1) / g.KB;
    for (int w = 0; w < kStreamWarps; ++w) stream_nblk[w] = w < g.NW ? (int)sub[w].size() : 0;
    for (auto& v : sub) while (v.size() < nchunks * g.KB) add_block(v, nullptr, 0, 0u, -1, 0.0);   // inert padding blocks

    // ctl (shared by all passes)
    std::vector<uint32_t> ctl;
    for (auto& v : sub) for (auto& b : v) ctl.push_back(b.flags);
    d_ctl.reserve(ctl.size() * 4 + 256);
    CU(cudaMemcpy(d_ctl.p, ctl.data(), ctl.size() * 4, cudaMemcpyHostToDevice));

    // passes
    const int npass = (P + PB - 1) / PB;
    passes.resize(npass);
    for (int ps = 0; ps < npass; ++ps) {
        const int pb0 = ps * PB;
        std::vector<uint32_t> blocks((size_t)g.NW * nchunks * g.CH * 4, 0u);
        std::vector<uint32_t> tinfo((size_t)g.NW * ntinfo * g.TIQ * 4, 0u);
        auto chan = [&](int aa, int q) { return (aa >= 0 && pb0 + q < P) ? ceff[(size_t)aa * P + pb0 + q] : cplx(0, 0); };
        for (int w = 0; w < g.NW; ++w) {
            size_t ti = 0;
            for (size_t ib = 0; ib < sub[w].size(); ++ib) {
                const BlockRef& b = sub[w][ib];
                uint32_t* dst = &blocks[(((size_t)w * nchunks * g.KB) + ib) * g.QBP * 4];
                for (int k = 0; k < kBlkLeaves; ++k) {
                    dst[k] = b.L[k].code;
                    if (g.CWORDS == 2) dst[4 + k] = b.L[k].code2;
                }
                double* wd = reinterpret_cast<double*>(dst + 4 * g.CWORDS);
                for (int k = 0; k < kBlkLeaves; ++k)
                    for (int q = 0; q < PB; ++q) {
                        const cplx z = chan(b.L[k].aa, q) * (b.L[k].sg * (double)b.L[k].mult);
                        if (CW) { wd[(k * PB + q) * 2] = z.real(); wd[(k * PB + q) * 2 + 1] = z.imag(); }
                        else wd[k * PB + q] = z.real();
                    }
                if (b.flags) {
                    uint32_t* t = &tinfo[(((size_t)w * ntinfo) + ti) * g.TIQ * 4];
                    ++ti;
                    unsigned toff = ONE * 512u, mx = 0u, my = 0u;
                    int aa1 = -1;
                    if (b.target >= 0) {
                        const unsigned c = (unsigned)T.iA_code[b.target];
                        toff = (c >> 2) * 512u;
                        mx = (c & 2u) ? 0x80000000u : 0u;
                        my = ((c & 1u) && !(c & 2u)) ? 0x80000000u : 0u;
                        aa1 = T.aa1_of_target[b.target];
                    }
                    t[0] = toff; t[1] = mx; t[2] = my; t[3] = 0u;
                    double* td = reinterpret_cast<double*>(t + 4);
                    td[0] = b.invnu;
                    for (int q = 0; q < PB; ++q) {
                        const cplx z = chan(aa1, q);
                        td[1 + q * g.CS] = z.real();
                        if (CW) td[2 + q * g.CS] = z.imag();
                    }
                }
            }
        }
        std::vector<double> w0((size_t)PB * g.CS, 0.0);
        if (T.has_const) for (int q = 0; q < PB; ++q) { const cplx z = chan(0, q); w0[q * g.CS] = z.real(); if (CW) w0[q * g.CS + 1] = z.imag(); }
        StreamPass& sp = passes[ps];
        sp.blocks.reserve(blocks.size() * 4 + 4096);
        sp.tinfo.reserve(tinfo.size() * 4 + 256);
        sp.w0.reserve(w0.size() * 8 + 64);
        CU(cudaMemcpy(sp.blocks.p, blocks.data(), blocks.size() * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(sp.tinfo.p, tinfo.data(), tinfo.size() * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(sp.w0.p, w0.data(), w0.size() * 8, cudaMemcpyHostToDevice));
        sp.pb0 = pb0;
    }
    stream_chunks = (int)nchunks;
    stream_ntinfo = (int)ntinfo;
    if (energy_only) return;                     // geometry (NF, PB, EPL) is the full stream's
    m->stream_nf = NF;
    m->stream_pb = PB;
    // two environments per lane for the single-channel real path when two such CTAs still fit on an SM
    m->stream_epl = (PB == 1 && !CW && 2 * (stream_smem(T.nS, g, PB, 2) + 1024) <= (size_t)m->smem_optin) ? 2 : 1;
    if (const char* ov = getenv("ACEB200_EPL")) m->stream_epl = (atoi(ov) == 2 && PB == 1 && !CW && stream_smem(T.nS, g, PB, 2) <= (size_t)m->smem_optin) ? 2 : 1;
    if (getenv("ACEB200_VERBOSE")) {
        size_t kept = 0, total = 0;
        for (int i = 0; i < T.nAA; ++i) { total += T.orders[i] >= 2; kept += (T.orders[i] >= 2 && keep[i]); }
        fprintf(stderr, "[aceb200] stream: NF=%d PB=%d CW=%d passes=%d, %d sub-streams x %zu chunks of %d blocks, AA functions of order >= 2 kept %zu of %zu, smem %zu B\n",
                NF, PB, (int)CW, npass, g.NW, nchunks, g.KB, kept, total, stream_smem(T.nS, g, PB, m->stream_epl));
    }
}

// Mirror partner of every AA function (all m negated), or -1: AA' = s conj(AA), s = (-1)^{sum m}
static void mirror_partners(const HostTables& T, std::vector<int>& partner, std::vector<int>& sign)
{
    partner.assign(T.nAA, -1); sign.assign(T.nAA, 1);
    std::map<std::tuple<int, int, int, int>, int> inv1p;
    for (int a = 0; a < T.nA; ++a) inv1p[std::make_tuple(T.iA_q[a], T.iA_n[a], T.iA_l[a], T.iA_m[a])] = a;
    std::vector<int> mirA(T.nA, -1);
    for (int a = 0; a < T.nA; ++a) {
        auto it = inv1p.find(std::make_tuple(T.iA_q[a], T.iA_n[a], T.iA_l[a], -T.iA_m[a]));
        if (it != inv1p.end()) mirA[a] = it->second;
    }
    std::map<std::vector<int>, int> invAA;
    auto keyof = [&](int i, bool mirror, bool& ok) {
        std::vector<int> k;
        ok = true;
        for (int t = 0; t < T.orders[i]; ++t) {
            int a = T.spec[(size_t)i * T.maxord + t];
            if (mirror) { a = mirA[a]; if (a < 0) { ok = false; a = 0; } }
            k.push_back(a);
        }
        std::sort(k.begin(), k.end(), std::greater<int>());
        return k;
    };
    bool ok;
    for (int i = 0; i < T.nAA; ++i) invAA[keyof(i, false, ok)] = i;
    for (int i = 0; i < T.nAA; ++i) {
        std::vector<int> k = keyof(i, true, ok);
        if (!ok) continue;
        auto it = invAA.find(k);
        if (it == invAA.end()) continue;
        partner[i] = it->second;
        int summ = 0;
        for (int t = 0; t < T.orders[i]; ++t) summ += T.iA_m[T.spec[(size_t)i * T.maxord + t]];
        sign[i] = (summ & 1) ? -1 : 1;
    }
}

// slot codes and sign bookkeeping of one AA function:  AA = sg * flip^{k1}(Y),  Y = A~_1 * prod_{f > 1} cj^{k_f xor k_1}(A~_f),
// so that  p Re(AA) - q Im(AA) = (sg p) Re(Y) - (sg fy q) Im(Y)
static void aa_codes(const HostTables& T, int aa, unsigned& c0, unsigned& c1, double& sg, double& fy)
{
    const unsigned ONE = (unsigned)T.nS;
    unsigned slot[4] = {ONE, ONE, ONE, ONE}, cj[4] = {0u, 0u, 0u, 0u};
    sg = 1.0;
    unsigned k1 = 0;
    for (int f = 0; f < T.orders[aa]; ++f) {
        const unsigned c = (unsigned)T.iA_code[T.spec[(size_t)aa * T.maxord + f]];
        const unsigned neg = c & 1u, odd = (c >> 1) & 1u;
        if (neg && odd) sg = -sg;
        if (f == 0) k1 = neg;
        slot[f] = c >> 2;
        cj[f] = (f > 0 && (neg ^ k1)) ? 1u : 0u;
    }
    fy = k1 ? -1.0 : 1.0;
    c0 = slot[0] | (slot[1] << 16) | (cj[1] << 31);
    c1 = slot[2] | (cj[2] << 15) | (slot[3] << 16) | (cj[3] << 31);
}

// Pack rows of leaves into the per-warp streams of k_basis_stream and choose the launch geometry.
static bool pack_bstream(aceb200_model* m, std::vector<std::vector<BLeaf>>& rows, int nfac, int nch, bool cw, BStream& out, const char* what)
{
    HostTables& T = m->T;
    out.nw = 0;
    int LB, LPC, HDR, W;
    if (!basis_geom(nfac, nch, cw, LB, LPC, HDR, W)) return false;
    const int cs = cw ? 2 : 1, nrows = (int)rows.size();
    const unsigned ONE = (unsigned)T.nS;
    const bool masked = (nfac == 3 && nch >= 3 && nch <= 9);     // BasisGeom::MASKED: channel mask in bits 16.. of the second code word
    size_t nleaves = 0;
    for (auto& r : rows) {
        if (r.empty()) {      // a structurally empty row still produces its zeros
            BLeaf L; L.c0 = ONE | (ONE << 16); L.c1 = masked ? ONE : (ONE | (ONE << 16)); L.w.assign((size_t)nch * cs, 0.0);
            r.push_back(L);
        }
        r.back().c0 |= kRowEnd;
        nleaves += r.size();
    }
    auto smem_of = [&](int nw, int epl) {
        return (size_t)(T.nS + 1) * 32 * epl * sizeof(c2) + (size_t)nw * 4 * LPC * LB + (size_t)nw * 32 * epl * (W | 1) * sizeof(double)
             + (size_t)(1 + 4 * nw) * 8;
    };
    // The kernel is latency-bound (dependent FP64 chains): what counts is warps per SM, so take as many as fit in shared
    // memory (and in the register file: 12 for the 9-channel, two-environment variant, which needs 168 registers).
    const bool epl2ok = basis_epl2(nch, cw);
    int epl = epl2ok ? 2 : 1;
    if (const char* ov = getenv("ACEB200_BASIS_EPL")) epl = (atoi(ov) == 2 && epl2ok) ? 2 : 1;
    if (epl == 2 && smem_of(8, 2) > (size_t)m->smem_optin) epl = 1;
    int nw = (epl == 2 && nch >= 9) ? 12 : kBasisMaxWarps;
    while (nw > 1 && smem_of(nw, epl) > (size_t)m->smem_optin) --nw;
    // two CTAs of 12 (10, 8) warps per SM when they fit: the TMA load of one CTA's next A tile then overlaps the other's walk
    // (config 1: 39 leaves per warp and tile -- with one CTA per SM the tile load was exposed, issue slots 32 % busy)
    if (!getenv("ACEB200_BASIS_ONE_CTA"))
        for (int w2 : {12, 10, 8})
            if (2 * (smem_of(w2, epl) + 1024) <= (size_t)m->smem_optin + 1024
                && basis_blocks_per_sm(nfac, nch, cw, epl, 32 * w2, smem_of(w2, epl)) >= 2) { nw = w2; break; }   // (the register file must hold both)
    if (const char* ov = getenv("ACEB200_BASIS_WARPS")) nw = std::max(1, std::min(nw, atoi(ov)));
    if (smem_of(nw, epl) > (size_t)m->smem_optin) return false;
    nw = std::min(nw, std::max(1, nrows));
    // contiguous row ranges balanced by cost: a leaf costs its product + nch channel updates, a row its share of a flush
    std::vector<int> cut(nw + 1, nrows);
    cut[0] = 0;
    { const double cfix = 14.0 + 6.0 * (nfac - 1) + (masked ? 2.0 * nch : 0.0), cchn = cw ? 3.0 : 2.0, crow = 8.0 + 14.0 * nch;
      auto rcost = [&](int r) {
          double c = crow;
          for (const BLeaf& L : rows[r]) c += cfix + cchn * (masked ? __builtin_popcount(L.c1 >> 16) : nch);
          return c;
      };
      double tot = 0.0, acc = 0.0;
      for (int r = 0; r < nrows; ++r) tot += rcost(r);
      int w = 1;
      for (int r = 0; r < nrows && w < nw; ++r) {
          acc += rcost(r);
          while (w < nw && acc * nw >= tot * w) cut[w++] = r + 1;
      } }
    size_t longest = 1;
    for (int w = 0; w < nw; ++w) { size_t nl = 0; for (int r = cut[w]; r < cut[w + 1]; ++r) nl += rows[r].size(); longest = std::max(longest, nl); }
    const size_t nchunks = (longest + LPC - 1) / LPC;
    const size_t chunk_bytes = (size_t)LPC * LB;
    std::vector<uint32_t> blocks((size_t)nw * nchunks * chunk_bytes / 4, 0u);
    for (int w = 0; w < nw; ++w) {
        size_t il = 0;
        for (int r = cut[w]; r < cut[w + 1]; ++r)
            for (const BLeaf& L : rows[r]) {
                unsigned char* rec = reinterpret_cast<unsigned char*>(blocks.data()) + ((size_t)w * nchunks + il / LPC) * chunk_bytes + (il % LPC) * LB;
                uint32_t hdr[2] = {L.c0, L.c1};
                memcpy(rec, hdr, 8);
                memcpy(rec + HDR, L.w.data(), (size_t)nch * cs * sizeof(double));
                ++il;
            }
        out.nleaf[w] = (int)il;
        out.row0[w] = cut[w];
    }
    out.buf.reserve(blocks.size() * 4 + 4096);
    CU(cudaMemcpy(out.buf.p, blocks.data(), blocks.size() * 4, cudaMemcpyHostToDevice));
    out.nw = nw; out.nchunks = (int)nchunks; out.nfac = nfac; out.nch = nch; out.cw = cw; out.LB = LB; out.LPC = LPC; out.W = W; out.epl = epl;
    out.nrows = nrows; out.smem = smem_of(nw, epl);
    if (getenv("ACEB200_VERBOSE")) {
        size_t active = 0;
        for (auto& r : rows) for (const BLeaf& L : r) active += masked ? __builtin_popcount(L.c1 >> 16) : nch;
        fprintf(stderr, "[aceb200] %s stream: NFAC=%d NCH=%d CW=%d EPL=%d, %d warps x %zu chunks of %d leaves, %zu leaves in %d rows (%.2f active channels per leaf), smem %zu B\n",
                what, nfac, nch, (int)cw, epl, nw, nchunks, LPC, nleaves, nrows, (double)active / std::max<size_t>(nleaves, 1), out.smem);
    }
    return true;
}

// Flatten A2Bmap (CSR) into the leaf streams k_basis_stream walks (layout: ace_kernels.cuh).
static void upload_basis_stream(aceb200_model* m)
{
    HostTables& T = m->T;
    m->bs.nw = 0;
    if (getenv("ACEB200_NO_BASIS_STREAM")) return;
    if (T.nB == 0 || T.maxord > 4 || T.nS + 1 >= (1 << 14)) return;
    const int nfac = std::max(1, T.maxord);
    const bool cw = !T.pireal;
    const int nch = T.ncomp * (T.symreal ? 1 : 2);
    const int cs = cw ? 2 : 1;
    // ---- leaves per row, mirror partners folded for a real B
    std::vector<int> partner, psign;
    const bool fold = T.symreal && !getenv("ACEB200_NO_MIRROR_FOLD");
    if (fold) mirror_partners(T, partner, psign);
    std::vector<std::vector<BLeaf>> rows(T.nB);
    for (int r = 0; r < T.nB; ++r) {
        // (AA column -> value) of this row
        std::map<int, const cplx*> ent;
        for (int k = T.csr_ptr[r]; k < T.csr_ptr[r + 1]; ++k) ent[T.csr_col[k]] = &T.csr_val[(size_t)k * T.ncomp];
        std::map<int, bool> done;
        for (auto& kv : ent) {
            const int aa = kv.first;
            if (done[aa]) continue;
            done[aa] = true;
            BLeaf L;
            double sg, fy;
            aa_codes(T, aa, L.c0, L.c1, sg, fy);
            const cplx* vp = nullptr;
            double ps = 0.0;
            if (fold && partner[aa] >= 0 && partner[aa] != aa) {
                auto it = ent.find(partner[aa]);
                if (it != ent.end()) { vp = it->second; ps = (double)psign[aa]; done[partner[aa]] = true; }
            }
            L.w.assign((size_t)nch * cs, 0.0);
            for (int c = 0; c < T.ncomp; ++c) {
                // value of this non-zero, with the mirror partner folded in:  v AA + v' s conj(AA) -> (p, q) on (Re AA, Im AA)
                const cplx v = kv.second[c];
                double pr = v.real(), qr = v.imag();            // Re(v X) = pr X.x - qr X.y
                double pi = v.imag(), qi = -v.real();           // Im(v X) = pi X.x - qi X.y
                if (vp) { const cplx v2 = vp[c]; pr += ps * v2.real(); qr -= ps * v2.imag(); }   // Re(v2 s conj X) = s (v2r X.x + v2i X.y)
                if (T.pireal) { qr = 0.0; qi = 0.0; }           // AA = Re(prod A): the imaginary part of the product is dropped
                auto put = [&](int chn, double pp, double qq) {
                    L.w[(size_t)chn * cs] = sg * pp;
                    if (cw) L.w[(size_t)chn * cs + 1] = -sg * fy * qq;       // the kernel adds w1 * Im: store -q
                };
                if (T.symreal) put(c, pr, qr);
                else { put(2 * c, pr, qr); put(2 * c + 1, pi, qi); }
            }
            unsigned mask = 0u;
            for (int q = 0; q < nch; ++q) for (int i = 0; i < cs; ++i) if (L.w[(size_t)q * cs + i] != 0.0) mask |= 1u << q;
            if (mask == 0u) continue;                            // a non-zero of A2Bmap whose value is exactly zero in every component
            if (nfac == 3 && nch >= 3 && nch <= 9) L.c1 = (L.c1 & 0xffffu) | (mask << 16);   // BasisGeom::MASKED
            rows[r].push_back(L);
        }
    }
    pack_bstream(m, rows, nfac, nch, cw, m->bs, "basis");
}

// one k_basis_stream launch: pooled A (ws_Ac) -> out [ne][S.nrows][S.nch]
static bool launch_bstream(aceb200_model* m, const BStream& S, long long ne, long long ldA, double* out)
{
    if (S.nw == 0) return false;
    HostTables& T = m->T;
    BasisParams p;
    memset(&p, 0, sizeof(p));
    p.nS = T.nS; p.nw = S.nw; p.nchunks = S.nchunks;
    for (int w = 0; w < S.nw; ++w) { p.nleaf[w] = S.nleaf[w]; p.row0[w] = S.row0[w]; }
    p.stream = S.buf.as<uint4>();
    p.Ac = t_cur->ws_Ac.as<c2>(); p.ldA = ldA; p.out = out; p.rowlen = (long long)S.nrows * S.nch; p.nenv = ne;
    const long long ntiles = (ne + 32 * S.epl - 1) / (32 * S.epl);
    const int per_sm = std::max<int>(1, std::min<int>(2048 / (32 * p.nw), (int)((size_t)m->smem_optin / (S.smem + 1024))));
    const int grid = (int)std::min<long long>(ntiles, (long long)m->sm_count * per_sm);
    if (!launch_basis_inst(S.nfac, S.nch, S.cw, S.epl, p, grid, S.smem, t_cur->stream)) return false;
    CU(cudaGetLastError());
    m->launches++;
    return true;
}

// B for a chunk: out [ne][nB][ncomp] (real or complex)
static bool launch_basis(aceb200_model* m, long long ne, long long ldA, double* out) { return launch_bstream(m, m->bs, ne, ldA, out); }

// k_dB_env's view of the tables: for every row of A2Bmap the distinct one-particle indices a that any of its products
// contains (an "entry"), and for every entry the (non-zero k, position t) pairs that contribute
// A2B[row,k] * prod_{s != t} A[v_s(k)] to its weight.  Row tiles are cut so that the weights and the staged results of
// one tile fit beside the dA planes in shared memory.
static void upload_db_pack(aceb200_model* m)
{
    const auto& T = m->T;
    m->db = aceb200_model::DbPack();
    // (a contribution record holds up to three other factors: correlation order <= 4)
    if (!T.symreal || T.nB == 0 || T.maxord < 1 || T.maxord > 4 || !(T.ncomp == 1 || T.ncomp == 3 || T.ncomp == 9)) return;
    const int NC = T.ncomp;
    std::vector<int> row_ent(1, 0), ent_a;
    std::vector<int4> ent_rec, con_rec;
    int emax = 0;
    for (int r = 0; r < T.nB; ++r) {
        std::map<int, std::vector<std::pair<int, int>>> by_a;            // canonical slot -> (k, t), ascending, contributions in (k, t) order
        for (int k = T.csr_ptr[r]; k < T.csr_ptr[r + 1]; ++k) {
            const int i = T.csr_col[k];
            for (int t = 0; t < T.orders[i]; ++t) by_a[T.iA_code[T.spec[(size_t)i * T.maxord + t]] >> 2].push_back({k, t});
        }
        for (auto& kv : by_a) {
            ent_a.push_back(kv.first);
            const int q0 = (int)con_rec.size();
            for (auto& kt : kv.second) {
                const int i = T.csr_col[kt.first];
                int f[3] = {-1, -1, -1}, nf = 0;
                for (int s2 = 0; s2 < T.orders[i]; ++s2) if (s2 != kt.second) f[nf++] = T.spec[(size_t)i * T.maxord + s2];
                con_rec.push_back(int4{kt.first * 4 + (T.iA_code[T.spec[(size_t)i * T.maxord + kt.second]] & 3), f[0], f[1], f[2]});
            }
            ent_rec.push_back(int4{kv.first * 3 * kDbPitch, q0, (int)con_rec.size(), 0});
        }
        row_ent.push_back((int)ent_a.size());
        emax = std::max(emax, row_ent[r + 1] - row_ent[r]);
    }
    const int RW = db_rows(NC), ngrp = (T.nB + RW - 1) / RW;
    auto gent = [&](int g) { return row_ent[std::min(T.nB, (g + 1) * RW)] - row_ent[g * RW]; };
    std::vector<int> tile_grp, tile_ent;
    std::vector<int4> grp_list;
    int threads = 0, ET = 0;
    // cut the rows into tiles whose weights fit beside the planes and the warps' strips; false: a group does not fit
    auto cut = [&](int thr, size_t limit) {
        const size_t base = db_env_fixed_smem(T.nA, T.nS, NC, thr);
        if (base + 4096 > limit) return false;
        const int nw = thr / 32;
        ET = (int)std::min<size_t>((limit - base) / (NC * sizeof(c2) + sizeof(int)), std::max<size_t>(ent_a.size(), 1));
        tile_grp.assign(1, 0); tile_ent.assign(1, 0); grp_list.clear();
        for (int g = 0; g < ngrp;) {
            int g1 = g;
            while (g1 < ngrp && row_ent[std::min(T.nB, (g1 + 1) * RW)] - row_ent[g * RW] <= ET) ++g1;
            if (g1 == g) return false;
            // deal the groups of the tile to the warps longest-first, in snake order: position w + k nw belongs to warp w
            std::vector<int> order(g1 - g);
            for (int i = 0; i < g1 - g; ++i) order[i] = g + i;
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return gent(x) > gent(y); });
            for (size_t i = 0; i < order.size(); ++i) {
                const size_t round = i / nw, k = i % nw;
                const size_t src = round * nw + ((round & 1) ? std::min<size_t>(order.size() - round * nw, nw) - 1 - k : k);
                const int r0 = order[src] * RW, eb = row_ent[g * RW];
                grp_list.push_back(int4{r0, row_ent[r0] - eb, row_ent[std::min(T.nB, r0 + 1)] - eb, row_ent[std::min(T.nB, r0 + 2)] - eb});
            }
            tile_grp.push_back(g1);
            tile_ent.push_back(row_ent[std::min(T.nB, g1 * RW)]);
            g = g1;
        }
        threads = thr;
        return true;
    };
    // two CTAs of half the threads per SM when at most three tiles result (the phases of one environment then overlap with
    // the other's), else one CTA with all the warps the registers allow; neither: the generic k_dAA + k_dB path runs
    const bool two = !getenv("ACEB200_DB_ONE_CTA") && cut(db_threads(NC) / 2, ((size_t)m->smem_optin + 1024) / 2 - 1024 - 256) && tile_grp.size() <= 4;
    if (!two && !cut(db_threads(NC), (size_t)m->smem_optin - 1024)) return;
    auto& D = m->db;
    D.nT = (int)tile_grp.size() - 1; D.ET = std::max(ET, 1);
    D.threads = threads;
    D.smem = db_env_smem(T.nA, T.nS, D.ET, NC, threads);
    D.tile_grp = upload(m->pool, tile_grp); D.tile_ent = upload(m->pool, tile_ent); D.grp_list = upload(m->pool, grp_list);
    if (ent_rec.empty()) ent_rec.push_back(int4{0, 0, 0, 0});
    if (con_rec.empty()) con_rec.push_back(int4{0, -1, -1, -1});
    D.ent_rec = upload(m->pool, ent_rec); D.con_rec = upload(m->pool, con_rec);
    D.ok = true;
}

static void upload_tables(aceb200_model* m)
{
    HostTables& T = m->T;
    std::vector<int32_t> cq, cl, cm, cc, cb, ci;
    for (const Column& c : T.cols) { cq.push_back(c.q); cl.push_back(c.l); cm.push_back(c.m); cc.push_back(c.cnt); cb.push_back(c.base); ci.push_back(c.ip); }
    ColumnsDev& C = m->C;
    C.ncols = T.ncols; C.nS = T.nS; C.nPused = (T.Lused + 1) * (T.Lused + 2) / 2; C.nQ = T.nQ;
    C.q = upload(m->pool, cq); C.l = upload(m->pool, cl); C.m = upload(m->pool, cm);
    C.cnt = upload(m->pool, cc); C.base = upload(m->pool, cb); C.ip = upload(m->pool, ci);
    C.colmap = upload(m->pool, T.colmap);
    {
        std::vector<int32_t> sn(T.nS), sip(T.nS), sqv(T.nS);
        for (const Column& c : T.cols)
            for (int n = 0; n < c.cnt; ++n) { sn[c.base + n] = n; sip[c.base + n] = c.ip; sqv[c.base + n] = c.q; }
        C.slot_n = upload(m->pool, sn); C.slot_ip = upload(m->pool, sip); C.slot_q = upload(m->pool, sqv);
    }
    {
        // 2 x 2 slot blocks for k_pool: per species, columns sorted by length and paired; a block is two
        // consecutive radial indices of such a pair (missing corners carry slot 0xffff and are never written)
        std::vector<int4> blk;
        for (int q = 0; q < T.nQ; ++q) {
            std::vector<const Column*> cs;
            for (const Column& c : T.cols) if (c.q == q) cs.push_back(&c);
            std::stable_sort(cs.begin(), cs.end(), [](const Column* a, const Column* b) { return a->cnt > b->cnt; });
            for (size_t i = 0; i < cs.size(); i += 2) {
                const Column* a = cs[i];
                const Column* b = i + 1 < cs.size() ? cs[i + 1] : nullptr;
                for (int n0 = 0; n0 < a->cnt; n0 += 2) {
                    const int n1 = n0 + 1 < a->cnt ? n0 + 1 : n0;
                    auto slot = [&](const Column* c, int n, bool ok) { return (c && ok && n < c->cnt) ? c->base + n : 0xffff; };
                    int4 d;
                    d.x = n0 | (n1 << 8) | (q << 16);
                    d.y = a->ip | ((b ? b->ip : a->ip) << 16);
                    d.z = slot(a, n0, true) | (slot(a, n1, n1 != n0) << 16);
                    d.w = slot(b, n0, true) | (slot(b, n1, n1 != n0) << 16);
                    blk.push_back(d);
                }
            }
        }
        m->n_pool_blk = (int)blk.size();
        m->d_pool_blk = upload(m->pool, blk);
    }
    if (T.nQ == 1) {
        // column tiles for k_pool_mma: columns sorted by length (longest first), four to a tile
        std::vector<const Column*> cs;
        for (const Column& c : T.cols) cs.push_back(&c);
        std::stable_sort(cs.begin(), cs.end(), [](const Column* a, const Column* b) { return a->cnt > b->cnt; });
        std::vector<PoolTile> tiles;
        for (size_t i = 0; i < cs.size(); i += 4) {
            PoolTile t;
            memset(&t, 0, sizeof(t));
            int longest = 0;
            for (int k = 0; k < 4; ++k) {
                const Column* c = i + k < cs.size() ? cs[i + k] : nullptr;
                t.ip[k] = c ? c->ip : 0; t.base[k] = c ? c->base : 0; t.cnt[k] = c ? c->cnt : 0;   // a missing column stores nothing
                longest = std::max(longest, t.cnt[k]);
            }
            t.nnt = (longest + 7) / 8;
            tiles.push_back(t);
        }
        m->n_pool_tiles = (int)tiles.size();
        m->d_pool_tiles = upload(m->pool, tiles);
        std::vector<ForceTile> ftiles;
        for (size_t i = 0; i < cs.size(); i += 4) {
            ForceTile t;
            memset(&t, 0, sizeof(t));
            int longest = 0;
            for (int k = 0; k < 4; ++k) {
                const Column* c = i + k < cs.size() ? cs[i + k] : nullptr;
                t.offF[k] = (c ? c->ip : 0) * 3 * kMmaPitch; t.offE[k] = (c ? c->m : 0) * 2 * kMmaPitch;
                t.base[k] = c ? c->base : 0; t.cnt[k] = c ? c->cnt : 0;
                longest = std::max(longest, t.cnt[k]);
            }
            t.ks = (longest + 3) / 4;
            ftiles.push_back(t);
        }
        m->d_force_tiles = upload(m->pool, ftiles);
    }
    m->d_slot_pos = upload(m->pool, T.slot_pos);
    m->d_slot_neg = upload(m->pool, T.slot_neg);
    m->d_code = upload(m->pool, T.iA_code);
    m->d_orders = upload(m->pool, T.orders);
    m->d_spec = upload(m->pool, T.spec);
    m->d_csr_ptr = upload(m->pool, T.csr_ptr);
    m->d_csr_col = upload(m->pool, T.csr_col);
    std::vector<c2> val(T.csr_val.size());
    for (size_t k = 0; k < val.size(); ++k) val[k] = c2{T.csr_val[k].real(), T.csr_val[k].imag()};
    m->d_csr_val = upload(m->pool, val);
    for (int nu = 0; nu <= kMaxOrdDev; ++nu) memset(&m->list[nu], 0, sizeof(ListDev));
    for (int nu = 2; nu <= T.maxord; ++nu) m->list[nu].ptr = upload(m->pool, T.trees[nu].ptr);
    upload_basis_stream(m);
    upload_db_pack(m);
}

// ----------------------------------------------------------------------------------------------
// batch handling
// ----------------------------------------------------------------------------------------------
struct Chunk { long long e0, e1, j0, j1; };

struct BatchView {
    const aceb200_batch* b;
    std::vector<long long> host_off;    // chunk boundary offsets when the batch is on the device; all offsets when on the host
    long long nenv, nJ;
};

// environments per chunk: bounded by a workspace budget per environment
static long long chunk_envs(long long nenv, size_t bytes_per_env, size_t budget)
{
    long long n = (long long)(budget / std::max<size_t>(bytes_per_env, 1));
    n = std::max<long long>(32, (n / 32) * 32);
    return std::min<long long>(n, ((nenv + 31) / 32) * 32);
}

__global__ void k_gather_offsets(const long long* off, long long nenv, long long step, long long nb, long long* out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nb) return;
    long long e = i * step;
    if (e > nenv) e = nenv;
    out[i] = off[e];
}

// Per-call state: where the inputs of a chunk live on the device.
struct Staged {
    const long long* off;   // device, points at the chunk's first offset
    const double* R;        // device
    const int* species;     // device or null
    long long jbase;        // absolute neighbour index of R[0]
};

static void validate_batch(const aceb200_batch* b)
{
    if (!b) throw ModelError(ACEB200_EDESC, "null batch");
    if (b->nenv < 0) throw ModelError(ACEB200_EDESC, "negative nenv");
    if (b->nenv > 0 && (!b->offsets || !b->R)) throw ModelError(ACEB200_EDESC, "null offsets / R");
    if (b->space != ACEB200_HOST && b->space != ACEB200_DEVICE) throw ModelError(ACEB200_EDESC, "batch.space must be HOST or DEVICE");
    if (b->space == ACEB200_HOST && b->nenv > 0) {
        // every offset, not only the chunk boundaries: a malformed interior offset would make the kernels read out of bounds
        const int64_t* off = b->offsets;
        bool ok = off[0] >= 0;
        for (int64_t e = 0; e < b->nenv; ++e) ok &= off[e + 1] >= off[e];
        if (!ok) throw ModelError(ACEB200_EDESC, "offsets must be non-decreasing");
        if (b->nJ > 0 && off[b->nenv] - off[0] != b->nJ) throw ModelError(ACEB200_EDESC, "batch.nJ does not match offsets");
    }
}

// boundary offsets for chunks of `step` environments
static std::vector<long long> boundary_offsets(aceb200_model* m, const aceb200_batch* b, long long step)
{
    long long nb = (b->nenv + step - 1) / step;
    std::vector<long long> out(nb + 1);
    if (b->space == ACEB200_HOST) {
        for (long long i = 0; i <= nb; ++i) out[i] = b->offsets[std::min(i * step, (long long)b->nenv)];
    } else {
        t_cur->ws_out.reserve((nb + 1) * sizeof(long long));
        auto kfn = k_gather_offsets;
        long long* dst = t_cur->ws_out.as<long long>();
        ACE_LAUNCH(kfn, dim3((unsigned)((nb + 1 + 127) / 128)), dim3(128), 0, t_cur->stream, reinterpret_cast<const long long*>(b->offsets), (long long)b->nenv, step, nb, dst);
        CU(cudaGetLastError());
        m->launches++;
        CU(cudaMemcpyAsync(out.data(), t_cur->ws_out.p, (nb + 1) * sizeof(long long), cudaMemcpyDeviceToHost, t_cur->stream));
        CU(cudaStreamSynchronize(t_cur->stream));
    }
    for (long long i = 0; i < nb; ++i)
        if (out[i + 1] < out[i]) throw ModelError(ACEB200_EDESC, "offsets must be non-decreasing");
    return out;
}

static Staged stage_chunk(aceb200_model* m, const aceb200_batch* b, const Chunk& c)
{
    Staged s;
    long long ne = c.e1 - c.e0, nj = c.j1 - c.j0;
    if (b->space == ACEB200_DEVICE) {
        s.off = reinterpret_cast<const long long*>(b->offsets) + c.e0;
        s.R = b->R + 3 * c.j0;
        s.species = b->species ? b->species + c.j0 : nullptr;
        s.jbase = c.j0;
        return s;
    }
    t_cur->in_off.reserve((ne + 1) * sizeof(long long));
    t_cur->in_R.reserve(std::max<long long>(nj, 1) * 3 * sizeof(double));
    CU(cudaMemcpyAsync(t_cur->in_off.p, b->offsets + c.e0, (ne + 1) * sizeof(long long), cudaMemcpyHostToDevice, t_cur->stream));
    if (nj > 0) CU(cudaMemcpyAsync(t_cur->in_R.p, b->R + 3 * c.j0, nj * 3 * sizeof(double), cudaMemcpyHostToDevice, t_cur->stream));
    s.off = t_cur->in_off.as<long long>();
    s.R = t_cur->in_R.as<double>();
    s.species = nullptr;
    if (b->species) {
        t_cur->in_sp.reserve(std::max<long long>(nj, 1) * sizeof(int));
        if (nj > 0) CU(cudaMemcpyAsync(t_cur->in_sp.p, b->species + c.j0, nj * sizeof(int), cudaMemcpyHostToDevice, t_cur->stream));
        s.species = t_cur->in_sp.as<int>();
    }
    s.jbase = c.j0;
    return s;
}

static BatchDev batch_dev(const Staged& s, const Chunk& c)
{
    BatchDev B;
    B.nenv = c.e1 - c.e0; B.off = s.off; B.R = s.R; B.species = s.species; B.jbase = s.jbase;
    B.gate = t_ctx ? t_ctx->ws_err.as<int>() : nullptr;
    return B;
}

// ----------------------------------------------------------------------------------------------
// kernel launch helpers
// ----------------------------------------------------------------------------------------------
// pooling on the FP64 tensor cores (single-species models; ACEB200_POOL_MMA=0 selects the FMA kernel for comparisons)
static bool launch_pool_mma(aceb200_model* m, const BatchDev& B, long long nJ, long long ldA)
{
    HostTables& T = m->T;
    static const bool off = getenv("ACEB200_POOL_MMA") && atoi(getenv("ACEB200_POOL_MMA")) == 0;
    if (off || T.nQ != 1 || B.species || m->n_pool_tiles == 0) return false;
    PoolMmaParams p;
    p.rp = m->rp; p.ap = m->ap; p.B = B;
    p.Ac = t_cur->ws_Ac.as<c2>(); p.ldA = ldA; p.errflag = t_ctx->ws_err.as<int>();
    p.tiles = m->d_pool_tiles; p.ntiles = m->n_pool_tiles;
    p.nP = (T.Lused + 1) * (T.Lused + 2) / 2;
    // environments per CTA: a multiple of the whole environments one 128-row sub-tile holds, so that the last
    // sub-tile of a CTA is as full as the others
    const double Jbar = std::max(1.0, (double)nJ / (double)std::max<long long>(1, B.nenv));
    int k = std::max(1, (int)(kPoolThreads / Jbar));
    p.TE = k >= 8 ? std::min(k, kPoolTEmax) : k * ((8 + k - 1) / k);
    if (p.TE > kPoolTEmax) p.TE = (kPoolTEmax / k) * k;
    const size_t smem = (size_t)kMmaPitch * (2 * p.nP + m->rp.N) * sizeof(double) + (size_t)p.ntiles * sizeof(PoolTile)
                      + (kPoolTEmax + 1) * sizeof(int);
    if (smem > (size_t)m->smem_optin) return false;
    launch_pool_mma_inst(m->NMAX, p.ap.L <= kStaticL, p, (unsigned)((B.nenv + p.TE - 1) / p.TE), smem, t_cur->stream);
    CU(cudaGetLastError());
    m->launches++;
    return true;
}

static void launch_pool(aceb200_model* m, const BatchDev& B, long long nJ, long long ldA)
{
    HostTables& T = m->T;
    if (launch_pool_mma(m, B, nJ, ldA)) return;
    PoolParams p;
    p.rp = m->rp; p.ap = m->ap; p.C = m->C; p.B = B;
    p.Ac = t_cur->ws_Ac.as<c2>(); p.ldA = ldA; p.errflag = t_ctx->ws_err.as<int>();
    if (m->n_pool_blk > kPoolThreads * kPoolBItems || T.nS >= 0xffff || T.nQ > 0xffff)
        throw ModelError(ACEB200_EUNSUPPORTED, "one-particle basis too large for k_pool (more than 256 slot blocks)");
    p.blk = m->d_pool_blk; p.nblk = m->n_pool_blk;
    p.nbp = 1;
    while (p.nbp < p.nblk && p.nbp < 32) p.nbp *= 2;
    if (p.nblk > 32) p.nbp = ((p.nblk + 31) / 32) * 32;
    // environments per CTA: a multiple of the whole environments one 128-row sub-tile holds, so that the last
    // sub-tile of a CTA is as full as the others
    {
        const double Jbar = std::max(1.0, (double)nJ / (double)std::max<long long>(1, B.nenv));
        int k = (int)(kPoolThreads / Jbar);
        k = std::max(1, std::min(k, kPoolThreads * kPoolBItems / std::max(1, p.nbp)));
        p.TE = k >= 8 ? std::min(k, kPoolTEmax) : k * ((8 + k - 1) / k);
        if (p.TE > kPoolTEmax) p.TE = (kPoolTEmax / k) * k;
    }
    p.nP = (T.Lused + 1) * (T.Lused + 2) / 2;
    const size_t smem = (size_t)kPoolPitch * (p.nP * sizeof(c2) + m->rp.N * sizeof(double)) + (kPoolThreads + kPoolTEmax + 1) * sizeof(int);
    dim3 grid((unsigned)((B.nenv + p.TE - 1) / p.TE));
    launch_pool_inst(m->NMAX, p.B.species != nullptr, p.ap.L <= kStaticL, p, grid.x, smem, t_cur->stream);
    CU(cudaGetLastError());
    m->launches++;
}

template <int NMAX>
static void launch_pool_w_t(aceb200_model* m, const PoolWParams& p, dim3 grid, size_t smem)
{
    if (p.B.species) {
        auto kfn = k_pool_w<NMAX, true>;
        CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ACE_LAUNCH(kfn, grid, dim3(kPoolThreads), smem, t_cur->stream, p);
        return;
    }
    auto kfn = k_pool_w<NMAX, false>;
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACE_LAUNCH(kfn, grid, dim3(kPoolThreads), smem, t_cur->stream, p);
}

static void launch_pool_w(aceb200_model* m, const BatchDev& B, const double* W, long long ldA)
{
    HostTables& T = m->T;
    PoolWParams p;
    p.rp = m->rp; p.ap = m->ap; p.C = m->C; p.B = B; p.W = W;
    p.Aw = t_cur->ws_Aw.as<c2>(); p.ldA = ldA; p.TE = 8;
    p.nP = (T.Lused + 1) * (T.Lused + 2) / 2;
    const size_t rowbytes = (size_t)2 * (p.nP * sizeof(c2) + m->rp.N * sizeof(double)), misc = (kPoolThreads + kPoolTEmax + 1) * sizeof(int);
    p.rows = kPoolThreads;
    if ((p.rows + 1) * rowbytes + misc > (size_t)m->smem_optin) p.rows = (int)(((size_t)m->smem_optin - misc) / rowbytes) - 1;
    if (p.rows < 8) throw ModelError(ACEB200_EUNSUPPORTED, "one-particle basis too large for the shared-memory staging of k_pool_w");
    if (p.rows & 1) p.rows -= 1;             // an odd pitch (rows + 1) keeps the staging rows conflict-free
    const size_t smem = (size_t)(p.rows + 1) * rowbytes + misc;
    dim3 grid((unsigned)((B.nenv + p.TE - 1) / p.TE));
    switch (m->NMAX) {
    case 4: launch_pool_w_t<4>(m, p, grid, smem); break;
    case 8: launch_pool_w_t<8>(m, p, grid, smem); break;
    case 12: launch_pool_w_t<12>(m, p, grid, smem); break;
    case 16: launch_pool_w_t<16>(m, p, grid, smem); break;
    case 20: launch_pool_w_t<20>(m, p, grid, smem); break;
    case 24: launch_pool_w_t<24>(m, p, grid, smem); break;
    default: launch_pool_w_t<32>(m, p, grid, smem); break;
    }
    CU(cudaGetLastError());
    m->launches++;
}

template <int PB, bool CW>
static void launch_adjoint_t(aceb200_model* m, const AdjointParams& p, int grid, size_t smem)
{
    auto kfn = k_adjoint<PB, CW>;
    CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACE_LAUNCH(kfn, dim3(grid), dim3(32), smem, t_cur->stream, p);
}

static void launch_adjoint(aceb200_model* m, long long nenv, long long ldA, bool want_D)
{
    HostTables& T = m->T;
    if (m->stream_chunks > 0) {
        const HostGeom g = stream_geom(m->stream_nf, m->stream_pb, m->cw);
        const size_t smem = stream_smem(T.nS, g, m->stream_pb, m->stream_epl);
        const long long ntiles = (nenv + 32 * m->stream_epl - 1) / (32 * m->stream_epl);
        const int per_sm = std::max<int>(1, std::min<int>(8, (int)((size_t)m->smem_optin / (smem + 1024))));
        const int grid = (int)std::min<long long>(ntiles, (long long)m->sm_count * per_sm);
        const bool eo = !want_D && m->e_stream_chunks > 0;       // energy only: the short stream
        for (const StreamPass& sp : (eo ? m->e_passes : m->passes)) {
            StreamParams p;
            p.nS = T.nS; p.has_const = T.has_const; p.want_D = want_D ? 1 : 0;
            p.nchunks = eo ? m->e_stream_chunks : m->stream_chunks; p.ntinfo = eo ? m->e_stream_ntinfo : m->stream_ntinfo;
            for (int w = 0; w < kStreamWarps; ++w) p.nblk[w] = eo ? m->e_stream_nblk[w] : m->stream_nblk[w];
            p.P = T.P; p.pb0 = sp.pb0;
            p.stream = sp.blocks.as<uint4>(); p.ctl = (eo ? m->e_ctl : m->d_ctl).as<unsigned>(); p.tinfo = sp.tinfo.as<uint4>(); p.w0 = sp.w0.as<double>();
            p.Ac = t_cur->ws_Ac.as<c2>(); p.ldA = ldA; p.Dt = t_cur->ws_Dt.as<c2>(); p.E = t_cur->ws_E.as<double>(); p.nenv = nenv;
            launch_stream_inst(m->stream_nf, m->stream_pb, m->cw, m->stream_epl, p, grid, smem, t_cur->stream);
            CU(cudaGetLastError());
            m->launches++;
        }
        return;
    }
    AdjointParams p;
    memset(&p, 0, sizeof(p));
    p.nS = T.nS; p.nA = T.nA; p.maxord = T.maxord; p.P = T.P; p.Ppad = m->Ppad; p.has_const = T.has_const; p.want_D = want_D ? 1 : 0;
    p.slot_pos = m->d_slot_pos; p.slot_neg = m->d_slot_neg; p.code = m->d_code;
    p.w1 = m->d_w1.as<double>(); p.w0 = m->d_w0.as<double>();
    for (int nu = 2; nu <= T.maxord; ++nu) p.list[nu] = m->list[nu];
    p.Ac = t_cur->ws_Ac.as<c2>(); p.ldA = ldA; p.Dt = t_cur->ws_Dt.as<c2>(); p.E = t_cur->ws_E.as<double>(); p.nenv = nenv;
    size_t smem = (size_t)T.nS * 32 * sizeof(c2);
    if (smem > (size_t)m->smem_optin)
        throw ModelError(ACEB200_EUNSUPPORTED, "one-particle basis too large for the shared-memory tile of k_adjoint");
    long long ntiles = (nenv + 31) / 32;
    int per_sm = std::max<int>(1, std::min<int>(32, (int)((size_t)m->smem_optin / (smem + 1024))));
    int grid = (int)std::min<long long>(ntiles, (long long)m->sm_count * per_sm);
    const bool cw = m->cw;
#define ACE_ADJ(PBV) { if (cw) launch_adjoint_t<PBV, true>(m, p, grid, smem); else launch_adjoint_t<PBV, false>(m, p, grid, smem); }
    switch (m->PB) {
    case 1: ACE_ADJ(1) break;
    case 2: ACE_ADJ(2) break;
    case 3: ACE_ADJ(3) break;
    default: ACE_ADJ(4) break;
    }
#undef ACE_ADJ
    CU(cudaGetLastError());
    m->launches++;
}

// The force contraction on the FP64 tensor cores (one channel, one species).  OPT-IN (ACEB200_FORCES_MMA=1): measured on
// B200 at config 2 it loses to k_forces, 7.5 ms against 3.45 ms per 10^6 environments (profiles/r2_dmma_study.md): staging
// 87 operand planes per neighbour and the per-(neighbour, column) epilogue cost more instructions than the 296 FMAs per
// neighbour they replace, and 103 KB of planes per CTA leaves 2 CTAs per SM.  Kept for the record and for wider bases.
static bool launch_forces_mma(aceb200_model* m, const BatchDev& B, long long nJ, long long ldA, double* G)
{
    HostTables& T = m->T;
    static const bool on = getenv("ACEB200_FORCES_MMA") && atoi(getenv("ACEB200_FORCES_MMA")) == 1;
    if (!on || T.P != 1 || T.nQ != 1 || B.species || m->n_pool_tiles == 0) return false;
    ForceMmaParams p;
    p.rp = m->rp; p.ap = m->ap; p.B = B;
    p.Dt = t_cur->ws_Dt.as<c2>(); p.ldA = ldA; p.nS = T.nS; p.dpitch = T.nS | 1; p.G = G;
    p.tiles = m->d_force_tiles; p.ntiles = m->n_pool_tiles;
    p.nP = (T.Lused + 1) * (T.Lused + 2) / 2;
    const double Jbar = std::max(1.0, (double)nJ / (double)std::max<long long>(1, B.nenv));
    int k = std::max(1, std::min(kFmmaEnvs, (int)(kFmmaRows / Jbar)));       // whole environments per 128-row sub-tile
    p.TE = k >= 8 ? k : k * ((8 + k - 1) / k);
    if (p.TE > kForceTEmax) p.TE = (kForceTEmax / k) * k;
    const size_t smem = (size_t)kMmaPitch * (2 * m->rp.N + 3 * p.nP + 2 * (T.Lused + 1) + 9) * sizeof(double)
                      + (size_t)kFmmaEnvs * p.dpitch * sizeof(c2) + (size_t)p.ntiles * sizeof(ForceTile) + (kForceTEmax + 1) * sizeof(int);
    if (smem > (size_t)m->smem_optin) return false;
    launch_forces_mma_inst(m->NMAX, p.ap.L <= kStaticL, p, (unsigned)((B.nenv + p.TE - 1) / p.TE), smem, t_cur->stream);
    CU(cudaGetLastError());
    m->launches++;
    return true;
}

// D2[s][comp][e] = sum_prop dp[prop] * D[s][prop * ncomp + comp][e]: the pullback of _rrule_evaluate(dp::SVector, ...)
// (src/evaluator.jl:183, contract(dp, c~)) applied after the adjoint pass, so that ONE force field is assembled
struct DpVec { double v[32]; };
__global__ void k_contract_D(int nS, int nprop, int ncomp, long long ldA, long long nenv, DpVec dp, const c2* D, c2* D2)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int sc = blockIdx.y, s = sc / ncomp, comp = sc - s * ncomp;
    if (e >= nenv) return;
    c2 acc = c2{0.0, 0.0};
    for (int pr = 0; pr < nprop; ++pr) {
        const c2 d = D[((size_t)s * nprop * ncomp + pr * ncomp + comp) * ldA + e];
        acc.x += dp.v[pr] * d.x; acc.y += dp.v[pr] * d.y;
    }
    D2[((size_t)s * ncomp + comp) * ldA + e] = acc;
}

// contracted == true: D~ has been contracted over the properties (k_contract_D): ncomp channels, one "property"
static void launch_forces(aceb200_model* m, const BatchDev& B, long long nJ, long long ldA, double* G, bool contracted = false)
{
    if (nJ == 0) return;
    if (!contracted && launch_forces_mma(m, B, nJ, ldA, G)) return;
    ForceParams p;
    p.rp = m->rp; p.ap = m->ap; p.C = m->C; p.B = B;
    p.Dt = contracted ? t_cur->ws_A2.as<c2>() : t_cur->ws_Dt.as<c2>(); p.ldA = ldA;
    p.P = contracted ? m->T.ncomp : m->T.P; p.nprop = contracted ? 1 : m->T.nprop; p.ncomp = m->T.ncomp; p.G = G;
    const int pbm = contracted ? pick_pb(m->T.ncomp) : m->PB;
    const int pb = pbm == 1 ? 1 : (pbm == 3 ? 3 : 2);
    p.dpitch = (m->T.nS * pb) | 1;
    // environments per CTA: the largest TE (within shared memory for 5 CTAs per SM, and leaving two waves of
    // CTAs) whose neighbours waste the fewest lanes of the last pass over kForceThreads
    const size_t misc = ((size_t)m->C.nQ * m->C.nPused + kForceTEmax + 1) * sizeof(int) + (size_t)m->NMAX * kForceThreads * sizeof(double);
    const double Jbar = (double)nJ / (double)B.nenv;
    int best = 1;
    double besteff = 0.0;
    int minb = ACE_FORCE_MINB;                       // resident CTAs per SM the shared-memory budget is sized for
    if (const char* ov = getenv("ACEB200_FORCE_MINB")) minb = std::max(1, atoi(ov));
    for (int te = 1; te <= kForceTEmax; ++te) {
        const size_t sm = (size_t)te * p.dpitch * sizeof(c2) + misc;
        if (te > 1 && (sm > (size_t)m->smem_optin / minb - 1024 || (B.nenv + te - 1) / te < 2LL * 5 * m->sm_count)) break;
        const double rows = te * Jbar, eff = rows / (kForceThreads * ceil(rows / kForceThreads));
        if (eff >= besteff - 1e-9) { besteff = eff; best = te; }
    }
    p.TE = best;
    const size_t smem = (size_t)p.TE * p.dpitch * sizeof(c2) + misc;
    const bool sp = B.species != nullptr;
    if (smem > (size_t)m->smem_optin)
        throw ModelError(ACEB200_EUNSUPPORTED, "one-particle basis too large for the shared-memory staging of k_forces");
    const unsigned grid = (unsigned)((B.nenv + p.TE - 1) / p.TE);
    launch_forces_inst(m->NMAX, pb, sp, p.ap.L <= kStaticL, p, grid, smem, t_cur->stream);
    CU(cudaGetLastError());
    m->launches++;
}

template <int NMAX>
static void launch_dA_t(aceb200_model* m, const dAParams& p)
{
    auto kfn = k_dA<NMAX>;
    ACE_LAUNCH(kfn, dim3((unsigned)((p.nJ + 127) / 128)), dim3(128), 0, t_cur->stream, p);
}

static void launch_dA(aceb200_model* m, const BatchDev& B, long long nJ, c2* dA, bool canon = false)
{
    if (nJ == 0) return;
    dAParams p;
    p.rp = m->rp; p.ap = m->ap; p.C = m->C; p.B = B; p.slot_pos = m->d_slot_pos; p.slot_neg = m->d_slot_neg;
    p.nA = m->T.nA; p.dA = dA; p.nJ = nJ; p.canon = canon ? 1 : 0;
    switch (m->NMAX) {
    case 4: launch_dA_t<4>(m, p); break;
    case 8: launch_dA_t<8>(m, p); break;
    case 12: launch_dA_t<12>(m, p); break;
    case 16: launch_dA_t<16>(m, p); break;
    case 20: launch_dA_t<20>(m, p); break;
    case 24: launch_dA_t<24>(m, p); break;
    default: launch_dA_t<32>(m, p); break;
    }
    CU(cudaGetLastError());
    m->launches++;
}

static unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

constexpr int kMaxCtx = 16;

// RAII lease of a context; nested use by the same thread (run_structure -> run) shares the outer lease
struct CtxLease {
    aceb200_model* m;
    bool owner = false;
    explicit CtxLease(aceb200_model* mm) : m(mm)
    {
        if (t_ctx) return;
        std::unique_lock<std::mutex> lk(m->ctx_mu);
        for (;;) {
            for (auto& c : m->ctxs) if (!c->in_use) { c->in_use = true; t_ctx = c.get(); break; }
            if (t_ctx) break;
            if ((int)m->ctxs.size() < kMaxCtx) {
                std::unique_ptr<Ctx> c(new Ctx());
                c->create();
                c->in_use = true;
                t_ctx = c.get();
                m->ctxs.push_back(std::move(c));
                break;
            }
            m->ctx_cv.wait(lk);
        }
        owner = true;
        t_cur = &t_ctx->lanes[0];
    }
    ~CtxLease()
    {
        if (!owner) return;
        if (std::uncaught_exceptions() > 0) {
            // the call is failing: nothing of it may still be in flight or flagged when the context is leased again
            cudaDeviceSynchronize();
            cudaMemset(t_ctx->ws_err.p, 0, sizeof(int));
            *t_ctx->h_err = 0;
            for (Lane& L : t_ctx->lanes) L.busy = false;
        }
        { std::lock_guard<std::mutex> lk(m->ctx_mu); t_ctx->in_use = false; }
        t_ctx = nullptr; t_cur = nullptr;
        m->ctx_cv.notify_one();
    }
};

static void throw_errflag(int flag)
{
    if (flag == 1) throw ModelError(ACEB200_EDESC, "offsets must be non-decreasing and within the neighbour arrays");
    if (flag == 5) throw ModelError(ACEB200_EEMPTY, "Product1pBasis can only be evaluated with non-empty configurations");
    if (flag == 6) throw ModelError(ACEB200_ECATEGORY, "species code not found in the category list");
}

// every environment of a DEVICE batch: off[e] <= off[e+1], all within [off[0], off[0] + nJ] when nJ is known
__global__ void k_check_offsets(const long long* off, long long nenv, long long nJ, int* errflag)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nenv) return;
    const long long a = off[e], b = off[e + 1];
    if (b < a || a < off[0] || (nJ > 0 && b - off[0] > nJ)) atomicMax(errflag, 1);
}

// Copy `bytes` of a result to the user's buffer (host or device space).
static void deliver(aceb200_model* m, const aceb200_batch* b, void* user, const void* dev, size_t bytes)
{
    if (!user || bytes == 0 || user == dev) return;
    CU(cudaMemcpyAsync(user, dev, bytes, b->space == ACEB200_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, t_cur->stream));
}

// ----------------------------------------------------------------------------------------------
// the evaluation driver
// ----------------------------------------------------------------------------------------------
enum Want { W_A = 1, W_AA = 2, W_B = 4, W_dA = 8, W_dAA = 16, W_dB = 32, W_E = 64, W_G = 128, W_ADJ = 256, W_DP = 512 };

struct Outputs {
    const double* w = nullptr;      // adjoint_EVAL_D: [sum J][3], same space as the batch
    const double* dp = nullptr;     // W_DP: [nprop] host, the property contraction of the pullback
    double* adj = nullptr;          // adjoint_EVAL_D: [nenv][nB][ncomp] complex
    double *A = nullptr, *AA = nullptr, *B = nullptr, *dA = nullptr, *dAA = nullptr, *dB = nullptr, *E = nullptr, *G = nullptr;
};

// Collect the timing events of a lane whose work has been synchronised.
static void harvest(aceb200_model* m, Lane& L, double& kernel_ms, double (&stage_ms)[3])
{
    if (!L.busy) return;
    CU(cudaStreamSynchronize(L.stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, L.ev0, L.ev1));
    kernel_ms += ms;
    if (L.timed_ef) {
        float a = 0.f, bb = 0.f, cc = 0.f;
        CU(cudaEventElapsedTime(&a, L.ev0, L.evA));
        CU(cudaEventElapsedTime(&bb, L.evA, L.evB));
        CU(cudaEventElapsedTime(&cc, L.evB, L.ev1));
        stage_ms[0] += a; stage_ms[1] += bb; stage_ms[2] += cc;
    }
    L.busy = false;
}

// What an internal caller (the structure driver) already knows about a DEVICE batch, so that run() need not
// synchronise with the GPU to learn it: the host copy of the offsets, and that the error flag is checked later.
struct RunHints {
    const long long* host_offsets = nullptr;   // [nenv + 1], the same numbers as the device-resident batch.offsets
    bool defer_errflag = false;                // do not reset / read back ws_err in this call, and do not wait for the
                                               // kernels either (no per-call timing): the caller synchronises once
};

static void run(aceb200_model* m, const aceb200_batch* b, int want, const Outputs& o, const RunHints* hints = nullptr)
{
    validate_batch(b);
    HostTables& T = m->T;
    if ((want & (W_E | W_G)) && !T.symreal)
        throw ModelError(ACEB200_EUNSUPPORTED, "energy / forces need a real symmetric basis (SymmetricBasis.real === real)");
    if (b->nenv == 0) return;
    CU(cudaSetDevice(m->device));
    std::shared_lock<std::shared_mutex> plock(m->params_mu, std::defer_lock);
    if (!t_ctx) plock.lock();                      // (a nested call runs under its caller's lock and lease)
    CtxLease lease(m);
    const int P = T.P, nA = T.nA, nAA = T.nAA, nB = T.nB, ncomp = T.ncomp;
    const int ca = T.pireal ? 1 : 2, cs = T.symreal ? 1 : 2;
    const bool fusedB = (want & W_B) && m->bs.nw > 0;       // B straight from the pooled A (k_basis_stream): AA never reaches HBM
    const bool need_full_A = (want & (W_A | W_AA | W_dA | W_dAA | W_dB | W_ADJ)) || ((want & W_B) && !fusedB);
    const bool dB_fusable = (want & W_dB) && !(want & (W_dAA | W_dA)) && m->db.ok && !getenv("ACEB200_NO_FUSED_DB");   // k_dB_env: dB from A and dA, no AA / dAA
    const bool need_AA = (want & (W_AA | W_dAA)) || ((want & W_dB) && !dB_fusable) || ((want & W_B) && !fusedB);
    const bool need_dA = want & (W_dA | W_dAA | W_dB);
    const bool need_dAA = want & (W_dAA | W_dB);
    const bool host = b->space == ACEB200_HOST;

    // lanes: a device-resident batch runs on the caller's stream; a host-resident batch is pipelined
    int nlanes = host ? kLanes : 1;
    if (host) if (const char* ov = getenv("ACEB200_LANES")) nlanes = std::max(1, std::min(kLanes, atoi(ov)));
    for (int l = 0; l < kLanes; ++l) { t_ctx->lanes[l].stream = host ? t_ctx->lanes[l].own_stream : m->user_stream; t_ctx->lanes[l].busy = false; }
    t_cur = &t_ctx->lanes[0];

    // workspace bytes per environment (J-dependent parts use the batch average, bounded below)
    auto bounds = [&](long long stp) {
        if (hints && hints->host_offsets) {
            const long long nb = (b->nenv + stp - 1) / stp;
            std::vector<long long> out(nb + 1);
            for (long long i = 0; i <= nb; ++i) out[i] = hints->host_offsets[std::min(i * stp, (long long)b->nenv)];
            return out;
        }
        return boundary_offsets(m, b, stp);
    };
    // total neighbour count: from the host offsets, from the caller (batch.nJ, offsets[0] = 0) or -- the only case that
    // costs a device -> host round trip before any kernel is launched -- read back from a DEVICE batch
    long long nJ_tot, j_first = 0;
    bool offsets_known = true;
    if (host) { j_first = b->offsets[0]; nJ_tot = b->offsets[b->nenv] - j_first; }
    else if (hints && hints->host_offsets) { j_first = hints->host_offsets[0]; nJ_tot = hints->host_offsets[b->nenv] - j_first; }
    else if (b->nJ > 0) nJ_tot = b->nJ;
    else { std::vector<long long> ends = bounds(b->nenv); j_first = ends[0]; nJ_tot = ends[1] - ends[0]; offsets_known = false; }
    (void)offsets_known;
    const double Jav = std::max(1.0, (double)nJ_tot / (double)b->nenv);
    size_t per_env = (size_t)T.nS * 16 + 64;
    if (want & W_G) per_env += (size_t)T.nS * P * 16 + (size_t)(Jav * (4 + 24.0 * P));
    if (want & (W_E | W_G)) per_env += (size_t)P * 8;
    if (need_full_A) per_env += (size_t)nA * 16;
    if (need_AA) per_env += (size_t)nAA * 8 * ca;
    if (want & W_B) per_env += (size_t)nB * ncomp * 8 * cs;
    if (need_dA) per_env += (size_t)(Jav * nA * 48.0);
    if (need_dAA && !dB_fusable) per_env += (size_t)(Jav * nAA * 24.0 * ca);
    if ((want & W_dB) && (host || !dB_fusable)) per_env += (size_t)(Jav * nB * 24.0 * ncomp * cs);
    if (want & W_ADJ) per_env += (size_t)T.nS * 16 + (size_t)nA * 16 + (size_t)nAA * 16 + (size_t)nB * ncomp * 16 + (size_t)(Jav * 24.0);
    if (host) per_env += (size_t)(Jav * 28.0) + 8;
    long long step = chunk_envs(b->nenv, per_env, (size_t)8 << 30);
    if (host) {
        // pipeline granularity: a few MiB of positions per chunk, at least ~6 chunks when the batch is large
        double pipe_mb = 64.0;
        if (const char* ov = getenv("ACEB200_PIPE_MB")) pipe_mb = std::max(1.0, atof(ov));
        long long pipe = std::max<long long>(4096, (long long)((pipe_mb * 1048576.0) / (24.0 * Jav)));
        pipe = std::min<long long>(pipe, std::max<long long>(4096, (b->nenv + 5) / 6));
        step = std::min<long long>(step, ((pipe + 31) / 32) * 32);
    }
    if (const char* ov = getenv("ACEB200_CHUNK_ENVS")) {   // test hook: force the multi-chunk path on small batches
        long long v = atoll(ov);
        if (v >= 32) step = std::min<long long>(step, (v / 32) * 32);
    }
    // chunk boundaries in environments (uniform; ramping the first and last chunks of a host batch down to step/8 was
    // measured and changes nothing: 23.6 ms vs 23.5 ms per 10^6 environments, the link itself is the bound)
    std::vector<long long> bo;
    if (step >= b->nenv) bo = {j_first, j_first + nJ_tot};          // one chunk: nothing to look up
    else bo = bounds(step);
    std::vector<long long> eb;
    for (long long e = 0; e < b->nenv; e += step) eb.push_back(e);
    eb.push_back(b->nenv);
    if (!host && !(hints && hints->host_offsets)) {
        // a DEVICE batch: its offsets are checked where they live (no synchronisation); the flag is read with the results
        auto kfn = k_check_offsets;
        int* flagp = t_ctx->ws_err.as<int>();
        ACE_LAUNCH(kfn, dim3((unsigned)((b->nenv + 255) / 256)), dim3(256), 0, t_ctx->lanes[0].stream,
                   reinterpret_cast<const long long*>(b->offsets), (long long)b->nenv, nJ_tot, flagp);
        CU(cudaGetLastError());
        m->launches++;
    }
    double kernel_ms = 0.0;
    double stage_ms[3] = {0.0, 0.0, 0.0};
    const long long nchunks = (long long)eb.size() - 1;
    for (long long ic = 0; ic < nchunks; ++ic) {
        Lane& L = t_ctx->lanes[ic % nlanes];
        t_cur = &L;
        if (!(hints && hints->defer_errflag && !host)) harvest(m, L, kernel_ms, stage_ms);      // the lane's previous chunk must be done before its buffers are
                                                      // reused (a deferred device batch runs on one stream: stream order suffices)
        Chunk c;
        c.e0 = eb[ic]; c.e1 = eb[ic + 1]; c.j0 = bo[ic]; c.j1 = bo[ic + 1];
        const long long ne = c.e1 - c.e0, nj = c.j1 - c.j0;
        const long long ldA = ((ne + 63) / 64) * 64;
        Staged st = stage_chunk(m, b, c);
        BatchDev B = batch_dev(st, c);
        L.ws_Ac.reserve((size_t)T.nS * ldA * sizeof(c2));
        if (want & W_G) L.ws_Dt.reserve((size_t)T.nS * P * ldA * sizeof(c2));
        CU(cudaEventRecord(L.ev0, L.stream));
        launch_pool(m, B, nj, ldA);
        L.timed_ef = (want & (W_E | W_G)) != 0;

        if (want & (W_E | W_G)) {
            L.ws_E.reserve((size_t)ne * P * sizeof(double));
            double* Gdev = nullptr;
            CU(cudaEventRecord(L.evA, L.stream));
            launch_adjoint(m, ne, ldA, (want & W_G) != 0);
            CU(cudaEventRecord(L.evB, L.stream));
            const size_t gper = (size_t)((want & W_DP) ? ncomp : P) * 3;      // doubles of gradient per neighbour
            if (want & W_G) {
                if (!host) Gdev = o.G + (size_t)c.j0 * gper;
                else { L.ws_G.reserve(std::max<long long>(nj, 1) * gper * sizeof(double)); Gdev = L.ws_G.as<double>(); }
                if (want & W_DP) {
                    DpVec dv;
                    for (int i = 0; i < 32; ++i) dv.v[i] = i < T.nprop ? o.dp[i] : 0.0;
                    L.ws_A2.reserve((size_t)T.nS * ncomp * ldA * sizeof(c2));
                    auto kfn = k_contract_D;
                    ACE_LAUNCH(kfn, dim3(blocks_for(ne, 128), (unsigned)(T.nS * ncomp)), dim3(128), 0, L.stream, T.nS, T.nprop, ncomp, ldA, ne, dv,
                               (const c2*)L.ws_Dt.as<c2>(), L.ws_A2.as<c2>());
                    CU(cudaGetLastError()); m->launches++;
                }
                launch_forces(m, B, nj, ldA, Gdev, (want & W_DP) != 0);
            }
            CU(cudaEventRecord(L.ev1, L.stream));
            if (o.E) deliver(m, b, o.E + (size_t)c.e0 * P, L.ws_E.p, (size_t)ne * P * sizeof(double));
            if ((want & W_G) && host)
                deliver(m, b, o.G + (size_t)c.j0 * gper, Gdev, (size_t)nj * gper * sizeof(double));
        } else {
            // basis values / Jacobians
            c2* dA_dev = nullptr; double* AA_dev = nullptr; double* dAA_dev = nullptr;
            double* B_dev = nullptr;
            if (fusedB) {
                CU(cudaEventRecord(L.evA, L.stream));
                if (!host) B_dev = o.B + (size_t)c.e0 * nB * ncomp * cs;
                else { L.ws_out.reserve((size_t)ne * nB * ncomp * 8 * cs); B_dev = L.ws_out.as<double>(); }
                if (!launch_basis(m, ne, ldA, B_dev)) throw ModelError(ACEB200_ECUDA, "k_basis_stream: no instantiation for this basis");
                CU(cudaEventRecord(L.evB, L.stream));
                L.timed_ef = true;
            }
            if (need_full_A) {
                L.ws_A.reserve((size_t)ne * nA * sizeof(c2));
                auto kfn = k_expand_A;
                ACE_LAUNCH(kfn, dim3(blocks_for(ne * nA, 256)), dim3(256), 0, L.stream, ne, nA, m->d_code, L.ws_Ac.as<c2>(), ldA, L.ws_A.as<c2>());
                CU(cudaGetLastError()); m->launches++;
            }
            if (want & W_ADJ) {
                // adjoint_EVAL_D: pool dAw like A, contract through the product basis, apply A2Bmap
                const double* Wdev = nullptr;
                if (!host) Wdev = o.w + 3 * c.j0;
                else {
                    L.in_w.reserve(std::max<long long>(nj, 1) * 3 * sizeof(double));
                    if (nj > 0) CU(cudaMemcpyAsync(L.in_w.p, o.w + 3 * c.j0, nj * 3 * sizeof(double), cudaMemcpyHostToDevice, L.stream));
                    Wdev = L.in_w.as<double>();
                }
                L.ws_Aw.reserve((size_t)T.nS * ldA * sizeof(c2));
                L.ws_A2.reserve((size_t)ne * nA * sizeof(c2));
                L.ws_AA.reserve((size_t)ne * nAA * 16);
                L.ws_out.reserve((size_t)ne * nB * ncomp * 16);
                launch_pool_w(m, B, Wdev, ldA);
                { auto kfn = k_expand_A;
                  ACE_LAUNCH(kfn, dim3(blocks_for(ne * nA, 256)), dim3(256), 0, L.stream, ne, nA, m->d_code, (const c2*)L.ws_Aw.as<c2>(), ldA, L.ws_A2.as<c2>());
                  CU(cudaGetLastError()); m->launches++; }
                { auto kfn = k_AAw;
                  ACE_LAUNCH(kfn, dim3(blocks_for(ne * nAA, 128)), dim3(128), 0, L.stream, ne, nA, nAA, std::max(1, T.maxord), m->d_orders, m->d_spec,
                             (const c2*)L.ws_A.as<c2>(), (const c2*)L.ws_A2.as<c2>(), T.symreal, L.ws_AA.as<double>());
                  CU(cudaGetLastError()); m->launches++; }
                if (nB > 0) { auto kfn = k_Bw;
                  ACE_LAUNCH(kfn, dim3(blocks_for(ne * nB * ncomp, 256)), dim3(256), 0, L.stream, ne, nB, nAA, ncomp, m->d_csr_ptr, m->d_csr_col, m->d_csr_val,
                             (const double*)L.ws_AA.as<double>(), L.ws_out.as<double>());
                  CU(cudaGetLastError()); m->launches++; }
                CU(cudaEventRecord(L.ev1, L.stream));
                deliver(m, b, o.adj + (size_t)c.e0 * nB * ncomp * 2, L.ws_out.p, (size_t)ne * nB * ncomp * 16);
                L.busy = true;
                continue;
            }
            if (need_AA) {
                L.ws_AA.reserve((size_t)ne * nAA * 8 * ca);
                AA_dev = L.ws_AA.as<double>();
                auto kfn = k_AA;
                ACE_LAUNCH(kfn, dim3(blocks_for(ne * nAA, 256)), dim3(256), 0, L.stream, ne, nA, nAA, std::max(1, T.maxord), m->d_orders, m->d_spec,
                           (const c2*)L.ws_A.as<c2>(), T.pireal, AA_dev);
                CU(cudaGetLastError()); m->launches++;
            }
            if ((want & W_B) && !fusedB) {
                L.ws_out.reserve((size_t)ne * nB * ncomp * 8 * cs);
                B_dev = L.ws_out.as<double>();
                auto kfn = k_B;
                ACE_LAUNCH(kfn, dim3(blocks_for(ne * nB * ncomp, 256)), dim3(256), 0, L.stream, ne, nB, nAA, ncomp, m->d_csr_ptr, m->d_csr_col, m->d_csr_val,
                           (const double*)AA_dev, T.pireal, T.symreal, B_dev);
                CU(cudaGetLastError()); m->launches++;
            }
            if (need_dA) {
                L.ws_dA.reserve(std::max<long long>((nj + 31) / 32 * 32, 1) * nA * 3 * sizeof(c2));
                dA_dev = L.ws_dA.as<c2>();
                launch_dA(m, B, nj, dA_dev, dB_fusable);          // k_dB_env reads the canonical slots only
            }
            double* dB_dev = nullptr;
            const bool fused_dB = dB_fusable && nj > 0;
            if (fused_dB) {
                // dB without dAA in HBM (k_dB_env): one CTA per environment; a DEVICE batch is written in place
                if (!host) dB_dev = o.dB + (size_t)c.j0 * nB * 3 * ncomp * cs;
                else { L.ws_G.reserve((size_t)nj * nB * 24 * ncomp * cs); dB_dev = L.ws_G.as<double>(); }
                DbEnvParams q;
                q.nenv = ne; q.off = st.off; q.gate = t_ctx->ws_err.as<int>();
                q.nA = nA; q.nS = T.nS; q.nB = nB; q.nT = m->db.nT; q.pireal = T.pireal; q.ET = m->db.ET;
                q.tile_grp = m->db.tile_grp; q.tile_ent = m->db.tile_ent; q.grp_list = m->db.grp_list; q.ent_rec = m->db.ent_rec; q.con_rec = m->db.con_rec;
                q.val = m->d_csr_val; q.A = L.ws_A.as<c2>(); q.dA = dA_dev; q.dB = dB_dev;
                const size_t smem = m->db.smem;
#define ACE_DBF(NCV) { auto kfn = k_dB_env<NCV>; CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                ACE_LAUNCH(kfn, dim3((unsigned)ne), dim3(m->db.threads), smem, L.stream, q); }
                if (ncomp == 1) ACE_DBF(1) else if (ncomp == 3) ACE_DBF(3) else ACE_DBF(9)
#undef ACE_DBF
                CU(cudaGetLastError()); m->launches++;
            }
            if (need_dAA && !dB_fusable) {
                L.ws_dAA.reserve(std::max<long long>(nj, 1) * nAA * 24 * ca);
                dAA_dev = L.ws_dAA.as<double>();
                auto kfn = k_dAA;
                const int* gate = t_ctx->ws_err.as<int>();      // (a local: launch arguments are evaluated by the launching thread)
                ACE_LAUNCH(kfn, dim3(blocks_for(ne * nAA, 128)), dim3(128), 0, L.stream, ne, st.off, nA, nAA, std::max(1, T.maxord), m->d_orders, m->d_spec,
                           (const c2*)L.ws_A.as<c2>(), (const c2*)dA_dev, T.pireal, dAA_dev, gate);
                CU(cudaGetLastError()); m->launches++;
            }
            if ((want & W_dB) && nj > 0 && !dB_fusable) {
                L.ws_G.reserve((size_t)nj * nB * 24 * ncomp * cs);
                dB_dev = L.ws_G.as<double>();
                auto kfn = k_dB;
                ACE_LAUNCH(kfn, dim3(blocks_for(nj * nB * 3, 128)), dim3(128), 0, L.stream, nj, nB, nAA, ncomp, m->d_csr_ptr, m->d_csr_col, m->d_csr_val,
                           (const double*)dAA_dev, T.pireal, T.symreal, dB_dev);
                CU(cudaGetLastError()); m->launches++;
            }
            CU(cudaEventRecord(L.ev1, L.stream));
            if (o.A) deliver(m, b, o.A + (size_t)c.e0 * nA * 2, L.ws_A.p, (size_t)ne * nA * sizeof(c2));
            if (o.AA) deliver(m, b, o.AA + (size_t)c.e0 * nAA * ca, AA_dev, (size_t)ne * nAA * 8 * ca);
            if (o.B && (host || !fusedB)) deliver(m, b, o.B + (size_t)c.e0 * nB * ncomp * cs, B_dev, (size_t)ne * nB * ncomp * 8 * cs);
            if (o.dA) deliver(m, b, o.dA + (size_t)c.j0 * nA * 6, dA_dev, (size_t)nj * nA * 3 * sizeof(c2));
            if (o.dAA) deliver(m, b, o.dAA + (size_t)c.j0 * nAA * 3 * ca, dAA_dev, (size_t)nj * nAA * 24 * ca);
            if (o.dB) deliver(m, b, o.dB + (size_t)c.j0 * nB * 3 * ncomp * cs, dB_dev, (size_t)nj * nB * 24 * ncomp * cs);
        }
        L.busy = true;
    }
    t_cur = &t_ctx->lanes[0];
    if (hints && hints->defer_errflag) { for (int l = 0; l < nlanes; ++l) t_ctx->lanes[l].busy = false; return; }
    // One synchronisation per lane ends the call.  The error flag travels with it: for a single-stream (DEVICE) call
    // its copy and reset are enqueued behind the kernels before that synchronisation; a pipelined HOST call reads it
    // once every lane has finished.
    if (!host) {
        CU(cudaMemcpyAsync(t_ctx->h_err, t_ctx->ws_err.p, sizeof(int), cudaMemcpyDeviceToHost, t_cur->stream));
        CU(cudaMemsetAsync(t_ctx->ws_err.p, 0, sizeof(int), t_cur->stream));
    }
    for (int l = 0; l < nlanes; ++l) harvest(m, t_ctx->lanes[l], kernel_ms, stage_ms);
    if (host) {
        CU(cudaMemcpyAsync(t_ctx->h_err, t_ctx->ws_err.p, sizeof(int), cudaMemcpyDeviceToHost, t_cur->stream));
        CU(cudaMemsetAsync(t_ctx->ws_err.p, 0, sizeof(int), t_cur->stream));
    }
    CU(cudaStreamSynchronize(t_cur->stream));
    const int flag = *t_ctx->h_err;
    *t_ctx->h_err = 0;
    { std::lock_guard<std::mutex> lk(m->stat_mu);
      m->last_ms = kernel_ms;
      for (int i = 0; i < 3; ++i) m->stage_ms[i] = stage_ms[i]; }
    throw_errflag(flag);
}

// ----------------------------------------------------------------------------------------------
// the structure driver: neighbour list in, site energies / atomic forces / virial out
// ----------------------------------------------------------------------------------------------
static void run_structure(aceb200_model* m, const aceb200_structure* s, double* Esite, double* F, double* W)
{
    if (!s) throw ModelError(ACEB200_EDESC, "null structure");
    if (s->natoms < 0 || s->npairs < 0) throw ModelError(ACEB200_EDESC, "negative natoms / npairs");
    if (s->space != ACEB200_HOST && s->space != ACEB200_DEVICE) throw ModelError(ACEB200_EDESC, "structure.space must be HOST or DEVICE");
    HostTables& T = m->T;
    if (!T.symreal) throw ModelError(ACEB200_EUNSUPPORTED, "structure forces need a real symmetric basis");
    if (W && T.ncomp != 1) throw ModelError(ACEB200_EUNSUPPORTED, "the virial is defined for scalar (ncomp = 1) properties");
    if (T.has_cat && !s->species) throw ModelError(ACEB200_EDESC, "the model has a categorical basis: structure.species is required");
    if (s->npairs >= (1ll << 31)) throw ModelError(ACEB200_EUNSUPPORTED, "structure: npairs must be below 2^31");
    const int packed = (s->flags & ACEB200_NBR_PACKED) ? 1 : 0;
    if (packed && s->image) throw ModelError(ACEB200_EDESC, "structure: packed neighbour words carry the image shift; image must be NULL");
    if (packed && s->natoms > (1ll << 26)) throw ModelError(ACEB200_EUNSUPPORTED, "structure: packed neighbour words address at most 2^26 atoms");
    if (s->natoms == 0) return;
    if (!s->X || !s->first || (s->npairs > 0 && !s->nbr) || !F) throw ModelError(ACEB200_EDESC, "null X / first / nbr / F");
    CU(cudaSetDevice(m->device));
    std::lock_guard<std::mutex> lock(m->mu_s);                          // the structure buffers (s_*) are per handle
    std::shared_lock<std::shared_mutex> plock(m->params_mu);
    CtxLease lease(m);                                                   // held across the deferred per-chunk evaluations
    const bool host = s->space == ACEB200_HOST;
    const long long na = s->natoms, np = s->npairs;
    const int P = T.P, K = P * 3;
    const bool species = T.has_cat;
    cudaStream_t st = m->user_stream, cs = m->copy_stream;

    // chunks of centres: ~32 MiB of pair tables each when the structure comes from the host (so that the copy
    // of chunk k+1 overlaps the evaluation of chunk k); one chunk when it is already on the device
    std::vector<long long> cut{0};
    if (host) {
        if (s->first[0] != 0 || s->first[na] != np) throw ModelError(ACEB200_EDESC, "first[0] must be 0 and first[natoms] = npairs");
        for (long long i = 0; i < na; ++i)
            if (s->first[i + 1] < s->first[i]) throw ModelError(ACEB200_EDESC, "structure.first must be non-decreasing");
        const double pair_bytes = 4.0 + (s->image ? 3.0 : 0.0);
        double chunk_mb = 64.0;
        if (const char* ov = getenv("ACEB200_STRUCT_MB")) chunk_mb = std::max(0.001, atof(ov));
        const long long ppc = std::max<long long>(1024, (long long)(chunk_mb * 1048576.0 / pair_bytes));
        long long a = 0;
        int grow = getenv("ACEB200_NO_RAMP") ? 1 : 8;      // the first chunks are ppc/8, ppc/4, ppc/2: evaluation starts early
        while (a < na) {
            const long long target = s->first[a] + std::max<long long>(1024, ppc / grow);
            if (grow > 1) grow /= 2;
            long long b = std::upper_bound(s->first + a + 1, s->first + na + 1, target) - s->first - 1;   // last b with first[b] <= target
            b = std::max(a + 1, std::min(b, na));
            cut.push_back(b);
            a = b;
        }
    } else cut.push_back(na);
    const size_t nch = cut.size() - 1;
    while (m->s_ev.size() < nch + 3) { cudaEvent_t e; CU(cudaEventCreate(&e)); m->s_ev.push_back(e); }

    m->s_err.reserve(sizeof(int));
    CU(cudaMemsetAsync(m->s_err.p, 0, sizeof(int), st));
    m->s_R.reserve(std::max<long long>(np, 1) * 3 * sizeof(double));
    m->s_G.reserve(std::max<long long>(np, 1) * (size_t)K * sizeof(double));
    if (species) m->s_sp.reserve(std::max<long long>(np, 1) * sizeof(int));
    if (host || !s->rev) m->s_rev.reserve(std::max<long long>(np, 1) * sizeof(int));
    const double* dX; const long long* dfirst; const int* dnbr; const signed char* dimg; const int* dspc; const int* drev;
    CellDev cell;
    for (int i = 0; i < 9; ++i) cell.c[i] = (s->image || packed) ? s->cell[i] : 0.0;
    double *dE, *dF, *dW;
    if (host) {
        m->s_X.reserve(na * 3 * sizeof(double));
        m->s_first.reserve((na + 1) * sizeof(long long));
        m->s_nbr.reserve(std::max<long long>(np, 1) * sizeof(int));
        if (s->image) m->s_img.reserve(std::max<long long>(np, 1) * 3);
        if (species) m->s_spc.reserve(na * sizeof(int));

        m->s_E.reserve((size_t)na * P * sizeof(double));
        m->s_F.reserve((size_t)na * K * sizeof(double));
        m->s_W.reserve((size_t)T.nprop * 9 * sizeof(double));
        CU(cudaMemcpyAsync(m->s_X.p, s->X, na * 3 * sizeof(double), cudaMemcpyHostToDevice, cs));
        CU(cudaMemcpyAsync(m->s_first.p, s->first, (na + 1) * sizeof(long long), cudaMemcpyHostToDevice, cs));
        if (species) CU(cudaMemcpyAsync(m->s_spc.p, s->species, na * sizeof(int), cudaMemcpyHostToDevice, cs));
        for (size_t k = 0; k < nch; ++k) {
            const long long j0 = s->first[cut[k]], j1 = s->first[cut[k + 1]];
            if (j1 > j0) {
                CU(cudaMemcpyAsync(m->s_nbr.as<int>() + j0, s->nbr + j0, (j1 - j0) * sizeof(int), cudaMemcpyHostToDevice, cs));
                if (s->image) CU(cudaMemcpyAsync(m->s_img.as<signed char>() + 3 * j0, s->image + 3 * j0, (j1 - j0) * 3, cudaMemcpyHostToDevice, cs));
            }
            CU(cudaEventRecord(m->s_ev[k], cs));
        }
        if (s->rev && np > 0) CU(cudaMemcpyAsync(m->s_rev.p, s->rev, np * sizeof(int), cudaMemcpyHostToDevice, cs));
        CU(cudaEventRecord(m->s_ev[nch], cs));
        dX = m->s_X.as<double>(); dfirst = m->s_first.as<long long>(); dnbr = m->s_nbr.as<int>();
        dimg = s->image ? m->s_img.as<signed char>() : nullptr; dspc = species ? m->s_spc.as<int>() : nullptr;
        drev = s->rev ? m->s_rev.as<int>() : nullptr;
        dE = m->s_E.as<double>(); dF = m->s_F.as<double>(); dW = m->s_W.as<double>();
    } else {
        dX = s->X; dfirst = reinterpret_cast<const long long*>(s->first); dnbr = s->nbr; dimg = reinterpret_cast<const signed char*>(s->image);
        dspc = species ? s->species : nullptr; drev = s->rev;
        if (!Esite) m->s_E.reserve((size_t)na * P * sizeof(double));
        dE = Esite ? Esite : m->s_E.as<double>(); dF = F; dW = W;
    }

    CU(cudaMemsetAsync(t_ctx->ws_err.p, 0, sizeof(int), st));
    CU(cudaEventRecord(m->s_ev[nch + 1], st));         // the whole device-side sequence is timed as one interval
    double kernel_ms = 0.0, stage_ms[3] = {0.0, 0.0, 0.0};
    for (size_t k = 0; k < nch; ++k) {
        const long long a0 = cut[k], a1 = cut[k + 1];
        if (host) CU(cudaStreamWaitEvent(st, m->s_ev[k], 0));
        { auto kfn = k_build_pairs;
          ACE_LAUNCH(kfn, dim3(blocks_for(a1 - a0, kPairAtoms)), dim3(128), (kPairAtoms + 1) * sizeof(long long), st, a0, a1 - a0, na, dfirst, dnbr, dimg, packed, cell, dX, dspc,
                     m->s_R.as<double>(), species ? m->s_sp.as<int>() : (int*)nullptr, m->s_err.as<int>());
          CU(cudaGetLastError()); m->launches++; }
        aceb200_batch sub;
        sub.nenv = a1 - a0; sub.offsets = reinterpret_cast<const int64_t*>(dfirst + a0); sub.R = m->s_R.as<double>();
        sub.species = species ? m->s_sp.as<int>() : nullptr; sub.space = ACEB200_DEVICE; sub._pad = 0; sub.nJ = 0;
        Outputs o; o.E = dE + (size_t)a0 * P; o.G = m->s_G.as<double>();
        RunHints hints;
        hints.host_offsets = host ? reinterpret_cast<const long long*>(s->first) + a0 : nullptr;
        hints.defer_errflag = true;              // one read-back for the whole structure, below
        run(m, &sub, W_E | W_G, o, &hints);
    }
    if (host) CU(cudaStreamWaitEvent(st, m->s_ev[nch], 0));
    if (!drev) {
        // no reverse table from the caller: find each pair's reverse on the device (all pair tables are resident now).
        // (Running this search on a second, high-priority stream underneath the evaluation kernels was measured and
        // is slower: the kernels contend for the same SMs, 21 ms vs 17 ms end to end on the benchmark structure.)
        CU(cudaMemsetAsync(m->s_rev.p, 0xff, std::max<long long>(np, 1) * sizeof(int), st));      // -1: no reverse pair
        auto kfn = k_find_rev;
        ACE_LAUNCH(kfn, dim3(blocks_for(na, kPairAtoms)), dim3(128), (kPairAtoms + 1) * sizeof(long long), st, na, dfirst, dnbr, dimg, packed, m->s_rev.as<int>());
        CU(cudaGetLastError()); m->launches++;
        drev = m->s_rev.as<int>();
    }
    { auto kfn = k_assemble_rev;
      ACE_LAUNCH(kfn, dim3(blocks_for(na * K, 256)), dim3(256), 0, st, na, K, dfirst, drev, (const double*)m->s_G.as<double>(), dF, m->s_err.as<int>(), np);
      CU(cudaGetLastError()); m->launches++; }
    if (W) {
        const int nblk = (int)std::max<long long>(1, std::min<long long>((long long)m->sm_count * 4, (np + kVirThreads - 1) / kVirThreads));
        m->s_part.reserve((size_t)T.nprop * nblk * 9 * sizeof(double));
        auto kfn = k_virial_partial;
        ACE_LAUNCH(kfn, dim3(nblk, T.nprop), dim3(kVirThreads), kVirThreads * sizeof(double), st, np, T.nprop, (const double*)m->s_G.as<double>(),
                   (const double*)m->s_R.as<double>(), m->s_part.as<double>());
        CU(cudaGetLastError()); m->launches++;
        auto kf2 = k_virial_final;
        ACE_LAUNCH(kf2, dim3(T.nprop), dim3(32), 0, st, nblk, (const double*)m->s_part.as<double>(), dW);
        CU(cudaGetLastError()); m->launches++;
    }
    CU(cudaEventRecord(m->s_ev[nch + 2], st));
    int flag = 0;
    if (host) {
        if (Esite) CU(cudaMemcpyAsync(Esite, dE, (size_t)na * P * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(F, dF, (size_t)na * K * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (W) CU(cudaMemcpyAsync(W, dW, (size_t)T.nprop * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    int eflag = 0;
    CU(cudaMemcpyAsync(&flag, m->s_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&eflag, t_ctx->ws_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemsetAsync(t_ctx->ws_err.p, 0, sizeof(int), st));
    CU(cudaStreamSynchronize(st));
    { float ms = 0.f; CU(cudaEventElapsedTime(&ms, m->s_ev[nch + 1], m->s_ev[nch + 2])); kernel_ms = ms; }
    { std::lock_guard<std::mutex> lk(m->stat_mu);
      m->last_ms = kernel_ms;
      for (int i = 0; i < 3; ++i) m->stage_ms[i] = stage_ms[i]; }
    if (flag) throw ModelError(ACEB200_EDESC, "structure: neighbour or reverse-pair index out of range");
    throw_errflag(eflag);
}

// ----------------------------------------------------------------------------------------------
// several GPUs behind one handle (aceb200_set_devices): environments are independent, so a HOST batch is cut into
// contiguous shards balanced by neighbour count, one per device, each evaluated by a replica of the model on its own
// host thread, streams and pinned pipeline; every shard writes its slice of the caller's buffers.  No collective is
// involved: per-environment outputs need none, and a total energy is a host-side sum of what was just copied back.
// ----------------------------------------------------------------------------------------------
static aceb200_model* clone_on_device(const aceb200_model* m, int dev)
{
    aceb200_model* r = new aceb200_model();
    try {
        r->device = dev;
        CU(cudaSetDevice(dev));
        r->T = m->T; r->rp = m->rp; r->ap = m->ap;
        r->NMAX = m->NMAX; r->PB = m->PB; r->Ppad = m->Ppad;
        CU(cudaDeviceGetAttribute(&r->sm_count, cudaDevAttrMultiProcessorCount, dev));
        CU(cudaDeviceGetAttribute(&r->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        CU(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
        r->devices.assign(1, dev);
        upload_tables(r);
        upload_weights(r, m->c_host.empty() ? nullptr : m->c_host.data());
        r->c_host = m->c_host;
    } catch (...) { aceb200_model_destroy(r); throw; }
    return r;
}

static void run_any(aceb200_model* m, const aceb200_batch* b, int want, const Outputs& o)
{
    const size_t ndev = 1 + m->replicas.size();
    if (ndev == 1 || !b || b->space != ACEB200_HOST || b->nenv < (int64_t)(2 * ndev)) { run(m, b, want, o); return; }
    validate_batch(b);
    HostTables& T = m->T;
    const int64_t* off = b->offsets;
    const int64_t j0 = off[0], nJ = off[b->nenv] - j0;
    std::vector<int64_t> cut(ndev + 1, b->nenv);
    cut[0] = 0;
    for (size_t k = 1; k < ndev; ++k) {
        const int64_t target = j0 + (int64_t)((double)nJ * (double)k / (double)ndev);
        int64_t e = std::lower_bound(off, off + b->nenv + 1, target) - off;
        cut[k] = std::max(cut[k - 1], std::min<int64_t>(e, b->nenv));
    }
    const int ca = T.pireal ? 1 : 2, cs = T.symreal ? 1 : 2;
    std::vector<int> codes(ndev, ACEB200_OK);
    std::vector<std::string> msgs(ndev);
    std::vector<std::thread> th;
    for (size_t k = 0; k < ndev; ++k) {
        const int64_t e0 = cut[k], e1 = cut[k + 1];
        if (e1 <= e0) continue;
        aceb200_model* mk = k == 0 ? m : m->replicas[k - 1];
        th.emplace_back([=, &codes, &msgs]() {
            try {
                aceb200_batch sub = *b;
                sub.nenv = e1 - e0; sub.offsets = off + e0; sub.nJ = 0;
                Outputs oo = o;                       // neighbour-indexed outputs (G, dA, dAA, dB, w) use absolute offsets
                if (o.E) oo.E = o.E + (size_t)e0 * T.P;
                if (o.A) oo.A = o.A + (size_t)e0 * T.nA * 2;
                if (o.AA) oo.AA = o.AA + (size_t)e0 * T.nAA * ca;
                if (o.B) oo.B = o.B + (size_t)e0 * T.nB * T.ncomp * cs;
                if (o.adj) oo.adj = o.adj + (size_t)e0 * T.nB * T.ncomp * 2;
                run(mk, &sub, want, oo);
            } catch (const ModelError& e) { codes[k] = e.code; msgs[k] = e.what(); }
            catch (const std::exception& e) { codes[k] = ACEB200_EDESC; msgs[k] = e.what(); }
        });
    }
    for (std::thread& t : th) t.join();
    double ms = 0.0, st[3] = {0.0, 0.0, 0.0};
    for (size_t k = 0; k < ndev; ++k) {
        aceb200_model* mk = k == 0 ? m : m->replicas[k - 1];
        std::lock_guard<std::mutex> lk(mk->stat_mu);
        ms = std::max(ms, mk->last_ms);
        for (int i = 0; i < 3; ++i) st[i] = std::max(st[i], mk->stage_ms[i]);
    }
    { std::lock_guard<std::mutex> lk(m->stat_mu); m->last_ms = ms; for (int i = 0; i < 3; ++i) m->stage_ms[i] = st[i]; }
    for (size_t k = 0; k < ndev; ++k) if (codes[k] != ACEB200_OK) throw ModelError(codes[k], msgs[k]);
}

// FP64 FMA throughput probe: 8 independent dependent-FMA chains per thread.  The roofline
// denominator of this path is the FP64 pipe, which MEASURED_PEAKS.json does not hold.
__global__ void k_fp64_peak(int iters, double* sink)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double x = 1.0000001, y = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * x + y; a1 = a1 * x + y; a2 = a2 * x + y; a3 = a3 * x + y;
        a4 = a4 * x + y; a5 = a5 * x + y; a6 = a6 * x + y; a7 = a7 * x + y;
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[threadIdx.x] = s;
}


// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) throughput probe: 8 independent accumulator pairs per thread.
// One warp-level m8n8k4 is 8 * 8 * 4 FMAs = 512 flops.  north_star asks for the tensor-core roofline of the
// multi-property readout to be reported against a measured figure.
__global__ void k_dmma_peak(int iters, double* sink)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k][0] = k * 1e-3; c[k][1] = -k * 1e-3; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) dmma(c[k][0], c[k][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    if (s == 123.456) sink[threadIdx.x] = s;
}

// ----------------------------------------------------------------------------------------------
// extern "C"
// ----------------------------------------------------------------------------------------------
#define API_BEGIN try {
#define API_END                                                                  \
    } catch (const ModelError& e) { return fail(e.code, e.what()); }             \
    catch (const std::bad_alloc&) { return fail(ACEB200_ENOMEM, "host out of memory"); } \
    catch (const std::exception& e) { return fail(ACEB200_EDESC, e.what()); }    \
    return ACEB200_OK;

extern "C" {

int aceb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int aceb200_set_device(int device)
{
    API_BEGIN
    int n = aceb200_device_count();
    if (device < 0 || device >= n) throw ModelError(ACEB200_ECUDA, "no such CUDA device");
    g_device = device;
    API_END
}

int aceb200_set_devices(aceb200_model* m, int n, const int* devices)
{
    API_BEGIN
    if (!m || n < 1 || !devices) throw ModelError(ACEB200_EDESC, "set_devices: null argument or empty device list");
    const int ndev = aceb200_device_count();
    bool has_own = false;
    for (int i = 0; i < n; ++i) {
        if (devices[i] < 0 || devices[i] >= ndev) throw ModelError(ACEB200_ECUDA, "set_devices: no such CUDA device");
        for (int k = 0; k < i; ++k) if (devices[k] == devices[i]) throw ModelError(ACEB200_EDESC, "set_devices: duplicate device");
        has_own |= devices[i] == m->device;
    }
    if (!has_own) throw ModelError(ACEB200_EDESC, "set_devices: the list must contain the device the model was created on");
    std::unique_lock<std::shared_mutex> lock(m->params_mu);
    for (aceb200_model* r : m->replicas) aceb200_model_destroy(r);
    m->replicas.clear();
    m->devices.assign(1, m->device);
    for (int i = 0; i < n; ++i) {
        if (devices[i] == m->device) continue;
        m->replicas.push_back(clone_on_device(m, devices[i]));
        m->devices.push_back(devices[i]);
    }
    CU(cudaSetDevice(m->device));
    API_END
}

int aceb200_last_error(char* buf, int n)
{
    if (!buf || n <= 0) return ACEB200_EDESC;
    snprintf(buf, (size_t)n, "%s", g_err.c_str());
    return ACEB200_OK;
}

int aceb200_model_create(const aceb200_desc* desc, aceb200_model** out)
{
    aceb200_model* m = nullptr;
    try {
        if (!desc || !out) throw ModelError(ACEB200_EDESC, "null argument");
        *out = nullptr;
        if (aceb200_device_count() <= 0) throw ModelError(ACEB200_ECUDA, "no CUDA device available: ace_b200 has no CPU path");
        m = new aceb200_model();
        m->device = g_device;
        CU(cudaSetDevice(m->device));
        build_tables(*desc, m->T);
        fill_params(m, *desc);
        m->NMAX = pick_nmax(desc->n_rad);
        m->PB = pick_pb(m->T.P);
        m->Ppad = ((m->T.P + m->PB - 1) / m->PB) * m->PB;
        CU(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device));
        CU(cudaDeviceGetAttribute(&m->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
        m->devices.assign(1, m->device);
        CU(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        upload_tables(m);
        upload_weights(m, desc->c);
        if (desc->c) m->c_host.assign(desc->c, desc->c + (size_t)m->T.nB * m->T.nprop);
        *out = m;
    } catch (const ModelError& e) { delete m; return fail(e.code, e.what()); }
    catch (const std::bad_alloc&) { delete m; return fail(ACEB200_ENOMEM, "host out of memory"); }
    catch (const std::exception& e) { delete m; return fail(ACEB200_EDESC, e.what()); }
    return ACEB200_OK;
}

int aceb200_model_destroy(aceb200_model* m)
{
    if (!m) return ACEB200_OK;
    for (aceb200_model* r : m->replicas) aceb200_model_destroy(r);
    m->replicas.clear();
    cudaSetDevice(m->device);
    for (DevBuf& b : m->pool) b.release();
    m->d_w0.release(); m->d_w1.release(); m->d_ctl.release(); m->bs.buf.release();
    for (StreamPass& sp : m->passes) { sp.blocks.release(); sp.tinfo.release(); sp.w0.release(); }
    for (StreamPass& sp : m->e_passes) { sp.blocks.release(); sp.tinfo.release(); sp.w0.release(); }
    m->e_ctl.release();
    for (int nu = 0; nu <= kMaxOrdDev; ++nu) m->d_lw[nu].release();
    for (auto& c : m->ctxs) c->release();
    m->ctxs.clear();
    { DevBuf* bufs[] = {&m->s_X, &m->s_first, &m->s_nbr, &m->s_img, &m->s_spc, &m->s_rev, &m->s_R, &m->s_sp, &m->s_G, &m->s_E,
                        &m->s_F, &m->s_W, &m->s_part, &m->s_err};
      for (DevBuf* b : bufs) b->release(); }
    for (cudaEvent_t e : m->s_ev) cudaEventDestroy(e);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    delete m;
    return ACEB200_OK;
}

int aceb200_set_params(aceb200_model* m, const double* c, int64_t n)
{
    API_BEGIN
    if (!m || !c) throw ModelError(ACEB200_EDESC, "null argument");
    if (n != (int64_t)m->T.nB * m->T.nprop) throw ModelError(ACEB200_EDESC, "set_params: expected nB*nprop coefficients");
    CU(cudaSetDevice(m->device));
    { std::unique_lock<std::shared_mutex> lock(m->params_mu);       // waits for running evaluations (src/evaluator.jl:35-38 mutates in place)
      upload_weights(m, c);
      m->c_host.assign(c, c + n); }
    for (aceb200_model* r : m->replicas) { const int rc = aceb200_set_params(r, c, n); if (rc != ACEB200_OK) return rc; }
    API_END
}

int aceb200_get_eff_coeffs(aceb200_model* m, double* ctilde)
{
    API_BEGIN
    if (!m || !ctilde) throw ModelError(ACEB200_EDESC, "null argument");
    memcpy(ctilde, m->ctilde.data(), m->ctilde.size() * sizeof(cplx));
    API_END
}

int aceb200_set_stream(aceb200_model* m, void* cuda_stream)
{
    if (!m) return fail(ACEB200_EDESC, "null model");
    m->user_stream = (cudaStream_t)cuda_stream;
    return ACEB200_OK;
}

int64_t aceb200_launch_count(const aceb200_model* m) { return m ? (int64_t)m->launches.load() : 0; }

int aceb200_model_sizes(const aceb200_model* m, aceb200_sizes* out)
{
    if (!m || !out) return fail(ACEB200_EDESC, "null argument");
    out->nA = m->T.nA; out->nAA = m->T.nAA; out->nB = m->T.nB; out->ncomp = m->T.ncomp; out->nprop = m->T.nprop;
    out->maxord = m->T.maxord; out->pireal = m->T.pireal; out->symreal = m->T.symreal;
    return ACEB200_OK;
}

int aceb200_last_kernel_ms(const aceb200_model* m, double* ms)
{
    if (!m || !ms) return fail(ACEB200_EDESC, "null argument");
    *ms = m->last_ms;
    return ACEB200_OK;
}

int aceb200_last_stage_ms(const aceb200_model* m, double* ms3)
{
    if (!m || !ms3) return fail(ACEB200_EDESC, "null argument");
    for (int i = 0; i < 3; ++i) ms3[i] = m->stage_ms[i];
    return ACEB200_OK;
}

int aceb200_measure_fp64(double* tflops)
{
    API_BEGIN
    if (!tflops) throw ModelError(ACEB200_EDESC, "null argument");
    if (aceb200_device_count() <= 0) throw ModelError(ACEB200_ECUDA, "no CUDA device");
    CU(cudaSetDevice(g_device));
    int sms = 148;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g_device));
    double* sink = nullptr;
    CU(cudaMalloc((void**)&sink, 1024 * sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int iters = 1 << 15, threads = 256, blocks = sms * 8;
    auto kfn = k_fp64_peak;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0, nullptr));
        ACE_LAUNCH(kfn, dim3(blocks), dim3(threads), 0, (cudaStream_t) nullptr, iters, sink);
        CU(cudaEventRecord(e1, nullptr));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        double fl = (double)blocks * threads * (double)iters * 8.0 * 2.0;
        if (ms > 0.f) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *tflops = best;
    API_END
}


int aceb200_measure_dmma(double* tflops)
{
    API_BEGIN
    if (!tflops) throw ModelError(ACEB200_EDESC, "null argument");
    if (aceb200_device_count() <= 0) throw ModelError(ACEB200_ECUDA, "no CUDA device");
    CU(cudaSetDevice(g_device));
    int sms = 148;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g_device));
    double* sink = nullptr;
    CU(cudaMalloc((void**)&sink, 1024 * sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int iters = 1 << 13, threads = 256, blocks = sms * 8;
    auto kfn = k_dmma_peak;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0, nullptr));
        ACE_LAUNCH(kfn, dim3(blocks), dim3(threads), 0, (cudaStream_t) nullptr, iters, sink);
        CU(cudaEventRecord(e1, nullptr));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        double fl = (double)blocks * (threads / 32) * (double)iters * 8.0 * 512.0;
        if (ms > 0.f) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *tflops = best;
    API_END
}

#define NEED(m, b) if (!(m) || !(b)) throw ModelError(ACEB200_EDESC, "null argument")

int aceb200_eval_A(aceb200_model* m, const aceb200_batch* b, double* A)
{ API_BEGIN NEED(m, b); Outputs o; o.A = A; run_any(m, b, W_A, o); API_END }

int aceb200_eval_AA(aceb200_model* m, const aceb200_batch* b, double* AA)
{ API_BEGIN NEED(m, b); Outputs o; o.AA = AA; run_any(m, b, W_AA, o); API_END }

int aceb200_eval_B(aceb200_model* m, const aceb200_batch* b, double* B)
{ API_BEGIN NEED(m, b); Outputs o; o.B = B; run_any(m, b, W_B, o); API_END }

int aceb200_eval_dA(aceb200_model* m, const aceb200_batch* b, double* A, double* dA)
{ API_BEGIN NEED(m, b); Outputs o; o.A = A; o.dA = dA; run_any(m, b, W_dA | (A ? W_A : 0), o); API_END }

int aceb200_eval_dAA(aceb200_model* m, const aceb200_batch* b, double* AA, double* dAA)
{ API_BEGIN NEED(m, b); Outputs o; o.AA = AA; o.dAA = dAA; run_any(m, b, W_dAA | (AA ? W_AA : 0), o); API_END }

int aceb200_eval_dB(aceb200_model* m, const aceb200_batch* b, double* B, double* dB)
{ API_BEGIN NEED(m, b); Outputs o; o.B = B; o.dB = dB; run_any(m, b, W_dB | (B ? W_B : 0), o); API_END }

int aceb200_adjoint_eval_d(aceb200_model* m, const aceb200_batch* b, const double* w, double* out)
{ API_BEGIN NEED(m, b); if (!w || !out) throw ModelError(ACEB200_EDESC, "null argument"); Outputs o; o.w = w; o.adj = out; run_any(m, b, W_ADJ, o); API_END }

int aceb200_energy(aceb200_model* m, const aceb200_batch* b, double* E)
{ API_BEGIN NEED(m, b); Outputs o; o.E = E; run_any(m, b, W_E, o); API_END }

int aceb200_energy_forces(aceb200_model* m, const aceb200_batch* b, double* E, double* G)
{ API_BEGIN NEED(m, b); if (!G) throw ModelError(ACEB200_EDESC, "null G"); Outputs o; o.E = E; o.G = G; run_any(m, b, W_E | W_G, o); API_END }

int aceb200_energy_forces_dp(aceb200_model* m, const aceb200_batch* b, const double* dp, double* E, double* G)
{ API_BEGIN NEED(m, b); if (!G || !dp) throw ModelError(ACEB200_EDESC, "null G / dp");
  if (m->T.nprop > 32) throw ModelError(ACEB200_EUNSUPPORTED, "energy_forces_dp: more than 32 properties");
  Outputs o; o.E = E; o.G = G; o.dp = dp; run_any(m, b, W_E | W_G | W_DP, o); API_END }

int aceb200_structure_energy_forces(aceb200_model* m, const aceb200_structure* s, double* Esite, double* F, double* W)
{ API_BEGIN if (!m) throw ModelError(ACEB200_EDESC, "null argument"); run_structure(m, s, Esite, F, W); API_END }

}  // extern "C"
