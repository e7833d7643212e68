// ace_tables.h -- host-side derivation of the device tables from an aceb200_desc.
//
// Everything here is integer bookkeeping done once per model (or once per set_params for the
// weights).  It turns the reference's data structures
//     Product1pBasis.indices  (src/product_1pbasis.jl:5-8)
//     PIBasisSpec.orders / iAA2iA  (src/pibasis.jl:10-13)
//     A2Bmap (CSC)  (src/symmbasis.jl:33-38)  and  c  (src/linearmodel.jl:36-40)
// into the three layouts the kernels want:
//   * canonical columns: the one-particle functions grouped by (species q, l, m >= 0); within a column the
//     radial index n is a dense prefix 0..cnt-1.  Functions with m < 0 are never stored: the pooled
//     A_{n l -m} equals (-1)^m conj(A_{n l m}) because Y_l^{-m} = (-1)^m conj(Y_l^m)
//     (src/polynomials/sphericalharmonics.jl:394-398), so one complex "slot" serves both.
//   * adjoint lists: for every one-particle function a ("target") and every correlation order nu, the flat
//     list of the AA functions that contain a:  dE/dA_a |_nu = sum_leaves w * prod_{others} A_other,
//     with weights w = multiplicity * c~ (the stage-2 loop of src/evaluator.jl:180-185 regrouped by
//     target, which turns its scatter-add into a gather with no atomics).
//   * CSR copy of A2Bmap for the row-parallel B = A2B * AA product (src/symmbasis.jl:248-264).
#pragma once

#include <algorithm>
#include <cmath>
#include <tuple>
#include <complex>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/aceb200.h"

namespace aceb200 {

typedef std::complex<double> cplx;

struct ModelError : std::runtime_error {
    int code;
    ModelError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

constexpr int kMaxOrd = 5;   // correlation orders supported by the adjoint lists (4 other factors per leaf)

struct Column { int q, l, m, cnt, base, ip; };

// One adjoint list per correlation order nu >= 2, sorted by target:
//   ptr[a]..ptr[a+1]   leaves of target a
//   codes[4*i + k]     A-code of the k-th other factor of leaf i (k < nu-1)
//   laa[i], lmult[i]   AA index (0-based) and multiplicity: weight = lmult * c~[laa]
struct Tree {
    int nu = 0;
    std::vector<int32_t> ptr;
    std::vector<uint16_t> codes;
    std::vector<int32_t> laa, lmult;
};

struct HostTables {
    // sizes and flags
    int nA = 0, nAA = 0, nB = 0, ncomp = 1, nprop = 1, P = 1, maxord = 0, nS = 0, ncols = 0, Lused = 0, nQ = 1;
    bool has_cat = false;               // the one-particle basis has a Categorical1pBasis component
    int pireal = 0, symreal = 0, has_const = 0;
    bool cw = false;   // complex weights
    // one-particle decode
    std::vector<int32_t> iA_q, iA_n, iA_l, iA_m, iA_slot, iA_code;
    std::vector<Column> cols;
    std::vector<int32_t> colmap;        // [nQ][nPused] -> column index or -1
    std::vector<int32_t> slot_pos, slot_neg;  // iA with m >= 0 / m < 0 stored in the slot, or -1
    // product basis (0-based copy, row-major [nAA][maxord])
    std::vector<int32_t> orders, spec;
    // A2B as CSR
    std::vector<int32_t> csr_ptr, csr_col;
    std::vector<cplx> csr_val;          // [nnz][ncomp]
    // CSC copy (for c~)
    std::vector<int32_t> colptr, rowval;
    std::vector<cplx> nzval;
    // trees for nu = 2..maxord (index nu)
    std::vector<Tree> trees;
    std::vector<int32_t> aa1_of_target;  // order-1 AA index of target a or -1
};

// A-code: how a kernel fetches A_a from the canonical slots:  slot*4 + (m<0 ? 1 : 0) + (m<0 && m odd ? 2 : 0)
inline int32_t make_code(int slot, int m) { return slot * 4 + (m < 0 ? 1 : 0) + ((m < 0 && (m & 1)) ? 2 : 0); }

inline void idx2lm(int i1, int& l, int& m)
{
    l = (int)std::floor(std::sqrt((double)(i1 - 1)) + 1e-10);   // sphericalharmonics.jl:104-108
    m = i1 - (l + l * l + 1);
}

inline void build_tables(const aceb200_desc& d, HostTables& T)
{
    if (d.abi_version != ACEB200_ABI_VERSION || d.struct_bytes != (int)sizeof(aceb200_desc))
        throw ModelError(ACEB200_EDESC, "descriptor ABI version / size mismatch");
    if (d.n_rad < 1 || d.n_rad > 32) throw ModelError(ACEB200_EUNSUPPORTED, "n_rad must be in 1..32");
    if (d.trans_kind < 0 || d.trans_kind > 3) throw ModelError(ACEB200_EUNSUPPORTED, "unknown distance transform kind (only id/poly/morse/agnesi cross the ABI)");
    if (d.n_comp < 2 || d.n_comp > ACEB200_MAX_COMP) throw ModelError(ACEB200_EUNSUPPORTED, "Product1pBasis needs 2..4 components");
    if (d.nA < 1 || d.nAA < 1 || d.nB < 0 || d.nnz < 0) throw ModelError(ACEB200_EDESC, "empty basis");
    if (!(d.ncomp == 1 || d.ncomp == 3 || d.ncomp == 9)) throw ModelError(ACEB200_EUNSUPPORTED, "ncomp must be 1, 3 or 9");
    if (d.nprop < 1) throw ModelError(ACEB200_EDESC, "nprop < 1");
    if (d.maxord > kMaxOrd) throw ModelError(ACEB200_EUNSUPPORTED, "correlation order > 5");
    if (!d.rad_A || !d.rad_B || !d.rad_C || !d.indices || !d.orders || (d.maxord > 0 && !d.iAA2iA) || !d.colptr || (d.nnz > 0 && (!d.rowval || !d.nzval)))
        throw ModelError(ACEB200_EDESC, "null table pointer");
    int ib_rn = -1, ib_ylm = -1, ib_cat = -1;
    for (int i = 0; i < d.n_comp; ++i) {
        int k = d.comp_kind[i];
        if (k == ACEB200_COMP_RN && ib_rn < 0) ib_rn = i;
        else if (k == ACEB200_COMP_YLM && ib_ylm < 0) ib_ylm = i;
        else if (k == ACEB200_COMP_CAT && ib_cat < 0) ib_cat = i;
        else throw ModelError(ACEB200_EUNSUPPORTED, "unsupported one-particle component (supported: one Rn, one Ylm, at most one Categorical)");
    }
    if (ib_rn < 0 || ib_ylm < 0) throw ModelError(ACEB200_EUNSUPPORTED, "the one-particle basis needs an Rn and a Ylm component");
    if (ib_cat >= 0 && d.n_cat < 1) throw ModelError(ACEB200_EDESC, "categorical component with n_cat < 1");

    T.nA = d.nA; T.nAA = d.nAA; T.nB = d.nB; T.ncomp = d.ncomp; T.nprop = d.nprop; T.P = d.nprop * d.ncomp;
    T.maxord = d.maxord; T.pireal = d.pireal; T.symreal = d.symreal;
    T.nQ = ib_cat >= 0 ? d.n_cat : 1;
    T.has_cat = ib_cat >= 0;

    // ---- decode the one-particle functions
    T.iA_q.resize(T.nA); T.iA_n.resize(T.nA); T.iA_l.resize(T.nA); T.iA_m.resize(T.nA);
    T.Lused = 0;
    for (int a = 0; a < T.nA; ++a) {
        const int32_t* phi = d.indices + (size_t)a * d.n_comp;
        int n1 = phi[ib_rn], y1 = phi[ib_ylm], q1 = ib_cat >= 0 ? phi[ib_cat] : 1;
        if (n1 < 1 || n1 > d.n_rad) throw ModelError(ACEB200_EDESC, "radial index out of range in indices");
        if (y1 < 1 || y1 > (d.maxL + 1) * (d.maxL + 1)) throw ModelError(ACEB200_EDESC, "Ylm index out of range in indices");
        if (q1 < 1 || q1 > T.nQ) throw ModelError(ACEB200_EDESC, "category index out of range in indices");
        int l, m; idx2lm(y1, l, m);
        T.iA_q[a] = q1 - 1; T.iA_n[a] = n1 - 1; T.iA_l[a] = l; T.iA_m[a] = m;
        T.Lused = std::max(T.Lused, l);
    }
    if (T.Lused > 12) throw ModelError(ACEB200_EUNSUPPORTED, "l > 12 referenced by the one-particle basis");

    // ---- canonical columns sorted by (q, m, l)
    std::map<std::tuple<int, int, int>, int> colcnt;
    for (int a = 0; a < T.nA; ++a) {
        auto key = std::make_tuple(T.iA_q[a], std::abs(T.iA_m[a]), T.iA_l[a]);
        int& c = colcnt[key];
        c = std::max(c, T.iA_n[a] + 1);
    }
    int nPused = (T.Lused + 1) * (T.Lused + 2) / 2;
    T.colmap.assign((size_t)T.nQ * nPused, -1);
    T.cols.clear();
    int base = 0;
    for (auto& kv : colcnt) {
        Column c;
        c.q = std::get<0>(kv.first); c.m = std::get<1>(kv.first); c.l = std::get<2>(kv.first);
        c.cnt = kv.second; c.base = base; c.ip = c.m + c.l * (c.l + 1) / 2;
        base += c.cnt;
        T.colmap[(size_t)c.q * nPused + c.ip] = (int)T.cols.size();
        T.cols.push_back(c);
    }
    T.ncols = (int)T.cols.size();
    T.nS = base;
    T.iA_slot.resize(T.nA); T.iA_code.resize(T.nA);
    T.slot_pos.assign(T.nS, -1); T.slot_neg.assign(T.nS, -1);
    for (int a = 0; a < T.nA; ++a) {
        int l = T.iA_l[a], m = T.iA_m[a], am = std::abs(m);
        int ci = T.colmap[(size_t)T.iA_q[a] * nPused + am + l * (l + 1) / 2];
        int s = T.cols[ci].base + T.iA_n[a];
        T.iA_slot[a] = s;
        T.iA_code[a] = make_code(s, m);
        int32_t& dst = (m < 0) ? T.slot_neg[s] : T.slot_pos[s];
        if (dst >= 0) throw ModelError(ACEB200_EDESC, "duplicate one-particle basis function in indices");
        dst = a;
    }

    // ---- product basis, 0-based row-major copy
    T.orders.assign(d.orders, d.orders + T.nAA);
    T.spec.assign((size_t)T.nAA * std::max(1, T.maxord), -1);
    T.has_const = 0;
    for (int i = 0; i < T.nAA; ++i) {
        int o = T.orders[i];
        if (o < 0 || o > T.maxord) throw ModelError(ACEB200_EDESC, "orders[i] outside 0..maxord");
        if (o == 0) {
            if (i != 0) throw ModelError(ACEB200_EDESC, "the order-0 function must be the first AA function (src/pibasis.jl:409)");
            T.has_const = 1;
        }
        for (int t = 0; t < o; ++t) {
            int v = d.iAA2iA[(size_t)t * T.nAA + i];
            if (v < 1 || v > T.nA) throw ModelError(ACEB200_EDESC, "iAA2iA entry out of range");
            T.spec[(size_t)i * T.maxord + t] = v - 1;
        }
    }

    // ---- A2B: CSC copy and CSR transpose
    T.colptr.assign(d.colptr, d.colptr + T.nAA + 1);
    if (T.colptr[0] != 1 || T.colptr[T.nAA] != d.nnz + 1) throw ModelError(ACEB200_EDESC, "colptr is not a 1-based CSC pointer array");
    T.rowval.assign(d.rowval, d.rowval + d.nnz);
    T.nzval.resize((size_t)d.nnz * T.ncomp);
    for (size_t k = 0; k < T.nzval.size(); ++k) T.nzval[k] = cplx(d.nzval[2 * k], d.nzval[2 * k + 1]);
    std::vector<int32_t> cnt(T.nB + 1, 0);
    for (int64_t k = 0; k < d.nnz; ++k) {
        if (T.rowval[k] < 1 || T.rowval[k] > T.nB) throw ModelError(ACEB200_EDESC, "rowval out of range");
        cnt[T.rowval[k]]++;
    }
    T.csr_ptr.assign(T.nB + 1, 0);
    for (int r = 0; r < T.nB; ++r) T.csr_ptr[r + 1] = T.csr_ptr[r] + cnt[r + 1];
    T.csr_col.resize(d.nnz); T.csr_val.resize((size_t)d.nnz * T.ncomp);
    std::vector<int32_t> fill(T.csr_ptr.begin(), T.csr_ptr.end() - 1);
    for (int col = 0; col < T.nAA; ++col)
        for (int k = T.colptr[col] - 1; k < T.colptr[col + 1] - 1; ++k) {
            int r = T.rowval[k] - 1, dst = fill[r]++;
            T.csr_col[dst] = col;
            for (int c = 0; c < T.ncomp; ++c) T.csr_val[(size_t)dst * T.ncomp + c] = T.nzval[(size_t)k * T.ncomp + c];
        }

    // ---- adjoint trees
    T.trees.assign(T.maxord + 1, Tree());
    T.aa1_of_target.assign(T.nA, -1);
    // gather (target, others...) -> (AA index, multiplicity)
    std::vector<std::map<std::vector<int32_t>, std::pair<int32_t, int32_t>>> entries(T.maxord + 1);
    for (int i = 0; i < T.nAA; ++i) {
        int o = T.orders[i];
        const int32_t* v = &T.spec[(size_t)i * T.maxord];
        if (o == 1) {
            if (T.aa1_of_target[v[0]] >= 0) throw ModelError(ACEB200_EDESC, "duplicate order-1 AA function");
            T.aa1_of_target[v[0]] = i;
        }
        if (o < 2) continue;
        for (int t = 0; t < o; ++t) {
            std::vector<int32_t> key;
            key.push_back(v[t]);
            for (int s = 0; s < o; ++s) if (s != t) key.push_back(v[s]);
            auto it = entries[o].find(key);
            if (it == entries[o].end()) entries[o][key] = std::make_pair(i, 1);
            else {
                if (it->second.first != i) throw ModelError(ACEB200_EDESC, "duplicate AA function in the product basis spec");
                it->second.second++;
            }
        }
    }
    if (T.nS * 4 + 3 > 65535) throw ModelError(ACEB200_EUNSUPPORTED, "too many one-particle slots for 16-bit A-codes");
    for (int nu = 2; nu <= T.maxord; ++nu) {
        Tree& tr = T.trees[nu];
        tr.nu = nu;
        tr.ptr.assign(T.nA + 1, 0);
        // std::map iterates the keys (target, others...) in lexicographic order: leaves come out sorted by target
        for (auto& kv : entries[nu]) {
            const std::vector<int32_t>& key = kv.first;
            for (int k = 0; k < 4; ++k) tr.codes.push_back(k < nu - 1 ? (uint16_t)T.iA_code[key[k + 1]] : (uint16_t)0);
            tr.laa.push_back(kv.second.first);
            tr.lmult.push_back(kv.second.second);
            tr.ptr[key[0] + 1]++;
        }
        for (int a = 0; a < T.nA; ++a) tr.ptr[a + 1] += tr.ptr[a];
    }
}

// c~ = transpose(A2Bmap) * c  (src/evaluator.jl:59-60; src/symmbasis.jl:267-285): [nAA][nprop][ncomp]
inline void eff_coeffs(const HostTables& T, const double* c, std::vector<cplx>& ct)
{
    ct.assign((size_t)T.nAA * T.P, cplx(0, 0));
    if (!c) return;
    for (int col = 0; col < T.nAA; ++col)
        for (int p = 0; p < T.nprop; ++p)
            for (int cc = 0; cc < T.ncomp; ++cc) {
                cplx tmp(0, 0);
                for (int k = T.colptr[col] - 1; k < T.colptr[col + 1] - 1; ++k)
                    tmp += T.nzval[(size_t)k * T.ncomp + cc] * c[(size_t)(T.rowval[k] - 1) * T.nprop + p];
                ct[((size_t)col * T.nprop + p) * T.ncomp + cc] = tmp;
            }
}

}  // namespace aceb200
