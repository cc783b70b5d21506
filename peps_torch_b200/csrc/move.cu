// Move driver + C ABI of libctmb: sequences the batched kernels of one CTM move on one stream.
// Reference call stack being replaced: SURVEY.md section 3.1 / 3.2
//   ctm_MOVE (ctm/generic/ctmrg.py:179-319) -> ctm_get_projectors_4x4 (ctm_projectors.py:14-64)
//   -> halves_of_4x4_CTM_MOVE_* -> c2x2_*_sl_c -> ctm_get_projectors_from_matrices
//   -> absorb_truncate_CTM_MOVE_* -> move_normalize_c ;  ctm_MOVE_sl (ctmrg_c4v.py:325-463).
#include "contract.h"
#include "../../include/ctmb.h"
#include <cmath>
#include <cstring>
#include <cctype>

namespace ctmb {
const std::string& get_error();
void jacobi_stats(unsigned long long out[2]);

struct Handle {
    Engine eng;
    explicit Handle(int dev) : eng(dev) {}
};

// ------------------------------------------------------------------------------------------
// tables (every string was checked against the reference's *_c functions
// through oracle/ctm_oracle.py, which uses the same tables)
// ------------------------------------------------------------------------------------------
struct CornerSpec { int c, t1, t2; const char* lc; const char* l1; const char* l2; const char* la; const char* out; };
// indices into ctmb_site::C / ::T
static const CornerSpec CORNERS[4] = {
    /* LU */ {0, 0, 1, "ab", "buc", "ael", "ulfg", "efcg"},
    /* RU */ {1, 3, 0, "ab", "brc", "eua", "ulfr", "elcf"},
    /* RD */ {2, 2, 3, "ab", "feb", "cra", "ulfr", "cuel"},
    /* LD */ {3, 1, 2, "ab", "cal", "fbe", "ulfr", "cuer"},
};
struct HalfSpec { int kind[4]; bool tr[4]; };
static const HalfSpec HALVES[4] = {
    /* UP    */ {{CTMB_RU, CTMB_RD, CTMB_LU, CTMB_LD}, {false, false, true, false}},
    /* LEFT  */ {{CTMB_LU, CTMB_RU, CTMB_LD, CTMB_RD}, {false, false, false, true}},
    /* DOWN  */ {{CTMB_LD, CTMB_LU, CTMB_RD, CTMB_RU}, {true, false, true, true}},
    /* RIGHT */ {{CTMB_RD, CTMB_LD, CTMB_RU, CTMB_LU}, {false, true, true, true}},
};
struct AbsorbSpec {
    int C1, T1, T, T2, C2;
    const char *lC1, *lT1, *lPt1, *oC1;     // nC1 = C1 . T1 . Pt1
    const char *lC2, *lT2, *lP2, *oC2;      // nC2 = C2 . T2 . P2
    const char *lT, *lPa, *lA, *lPb, *oT;   // nT = T . Pa . a . conj(a) . Pb
    bool pa_is_P1;                          // Pa = P1 (and Pb = Pt2) else Pa = Pt2 (and Pb = P1)
};
static const AbsorbSpec ABSORB[4] = {
    /* UP    */ {1, 3, 0, 1, 0, "ab", "brc", "arx", "xc", "ab", "ael", "blx", "ex", "auc", "alx", "uldr", "cry", "xdy", false},
    /* LEFT  */ {0, 0, 1, 2, 3, "ab", "buc", "aux", "xc", "ab", "fbe", "afx", "xe", "acl", "aux", "uldr", "cdy", "xyr", true},
    /* DOWN  */ {3, 1, 2, 3, 2, "ab", "cal", "blx", "cx", "ab", "cra", "brx", "cx", "fbe", "blx", "ulfr", "ery", "uxy", true},
    /* RIGHT */ {2, 2, 3, 0, 1, "ab", "feb", "afx", "xe", "ab", "eua", "bux", "ex", "arc", "aux", "uldr", "cdy", "xly", false},
};

// dims[0] == 0 marks a DOUBLE-LAYER on-site tensor A[u,l,d,r] whose legs are already the fused (ket,bra) pairs
// (what ctmrg.run builds under ctm_force_dl, ctmrg.py:51-61, run_overlap :137-147 and ctm_MOVE_dl ctmrg_c4v.py:229-233)
static inline bool is_dl(const ctmb_site& s) { return s.dims[0] == 0; }
// extent of the environment leg facing auxiliary leg `leg` (0..3 = u,l,d,r)
static inline int64_t aux2(const ctmb_site& s, int leg) {
    const int64_t D = s.dims[1 + leg];
    return is_dl(s) ? D : D * D;
}
static void t_dims(const ctmb_site& s, int which, int chi, int64_t d[3]) {
    switch (which) {
        case 0: d[0] = chi; d[1] = aux2(s, 0); d[2] = chi; break;
        case 1: d[0] = chi; d[1] = chi; d[2] = aux2(s, 1); break;
        case 2: d[0] = aux2(s, 2); d[1] = chi; d[2] = chi; break;
        default: d[0] = chi; d[1] = aux2(s, 3); d[2] = chi; break;
    }
}
static Tn env_T(const ctmb_site& s, int which, int chi, const char* lab) {
    int64_t d[3]; t_dims(s, which, chi, d);
    return make_tn(const_cast<void*>(s.T[which]), lab, {d[0], d[1], d[2]});
}
static Tn env_C(const ctmb_site& s, int which, int chi, const char* lab) {
    return make_tn(const_cast<void*>(s.C[which]), lab, {chi, chi});
}

// Build the single-layer chain for a double-layer spec: every label of `la` (the aux legs
// u,l,d,r of the on-site tensor, in that order) found on an operand is split into (ket,bra) =
// (lower, upper case); a and conj(a) are inserted at position `apos` of the operand list.
// open_phys: the physical legs stay open (ket 's', bra 'S') and are appended to the output, as the
// enlarged corners of the reduced density matrices need them (ctm/generic/rdm.py:1362-1592).
static Engine::ChainJob sl_job(std::vector<Tn> ops, size_t apos, const char* la, const ctmb_site& s,
                               const char* out_lab, void* out_ptr, unsigned long long* amax, bool open_phys = false) {
    Engine::ChainJob job;
    if (is_dl(s)) {
        CTMB_CHECK(!open_phys, "a double-layer on-site tensor has no open physical legs");
        // double-layer site: A enters as ONE operand with the fused labels, nothing is split
        Tn A = make_tn(const_cast<void*>(s.a), la, {s.dims[1], s.dims[2], s.dims[3], s.dims[4]});
        std::map<char, int64_t> ext;
        for (int d = 0; d < A.nd; ++d) ext[A.idx[d]] = A.dim[d];
        for (size_t i = 0; i < ops.size(); ++i) {
            if (i == apos) { job.ops.push_back(A); job.conj.push_back(false); }
            for (int d = 0; d < ops[i].nd; ++d) ext[ops[i].idx[d]] = ops[i].dim[d];
            job.ops.push_back(ops[i]); job.conj.push_back(false);
        }
        if (apos >= ops.size()) { job.ops.push_back(A); job.conj.push_back(false); }
        std::vector<int64_t> od;
        for (const char* p = out_lab; *p; ++p) { CTMB_CHECK(ext.count(*p), "output label not found"); od.push_back(ext[*p]); }
        job.out = make_tn(out_ptr, std::string(out_lab), od);
        job.amax = amax;
        return job;
    }
    auto Dof = [&](char c) -> int64_t { const char* p = strchr(la, c); return p ? s.dims[1 + (p - la)] : 0; };
    std::string lab_a = std::string("s") + la, lab_ac = lab_a;
    for (size_t i = open_phys ? 0 : 1; i < lab_ac.size(); ++i) lab_ac[i] = (char)toupper(lab_ac[i]);
    Tn a = make_tn(const_cast<void*>(s.a), lab_a, {s.dims[0], s.dims[1], s.dims[2], s.dims[3], s.dims[4]});
    Tn ac = relabel(a, lab_ac.c_str());
    std::map<char, int64_t> ext;
    for (size_t i = 0; i < ops.size(); ++i) {
        if (i == apos) { job.ops.push_back(a); job.conj.push_back(false); job.ops.push_back(ac); job.conj.push_back(true); }
        Tn t = ops[i];
        for (const char* p = la; *p; ++p)
            if (t.find(*p) >= 0) t = split_mode(t, *p, *p, (char)toupper(*p), Dof(*p), Dof(*p));
        for (int d = 0; d < t.nd; ++d) ext[t.idx[d]] = t.dim[d];
        job.ops.push_back(t); job.conj.push_back(false);
    }
    if (apos >= ops.size()) { job.ops.push_back(a); job.conj.push_back(false); job.ops.push_back(ac); job.conj.push_back(true); }
    std::string ol; std::vector<int64_t> od;
    for (const char* p = out_lab; *p; ++p) {
        if (strchr(la, *p)) { ol.push_back(*p); od.push_back(Dof(*p)); ol.push_back((char)toupper(*p)); od.push_back(Dof(*p)); }
        else { ol.push_back(*p); CTMB_CHECK(ext.count(*p), "output label not found"); od.push_back(ext[*p]); }
    }
    if (open_phys) { ol += "sS"; od.push_back(s.dims[0]); od.push_back(s.dims[0]); }
    job.out = make_tn(out_ptr, ol, od);
    job.amax = amax;
    return job;
}

static void corner_shape(int kind, const ctmb_site& s, int chi, int64_t& rows, int64_t& cols) {
    const CornerSpec& cs = CORNERS[kind];
    auto D2 = [&](char c) -> int64_t { const char* p = strchr(cs.la, c); return aux2(s, (int)(p - cs.la)); };
    rows = chi * D2(cs.out[1]);
    cols = chi * D2(cs.out[3]);
}

static Engine::ChainJob corner_job(int kind, const ctmb_site& s, int chi, void* out) {
    const CornerSpec& cs = CORNERS[kind];
    std::vector<Tn> ops = {env_C(s, cs.c, chi, cs.lc), env_T(s, cs.t1, chi, cs.l1), env_T(s, cs.t2, chi, cs.l2)};
    return sl_job(ops, 3, cs.la, s, cs.out, out, nullptr);
}

struct CornerReq { int kind; const ctmb_site* site; void* out; };

// label roles of a corner: k1/k2 = auxiliary legs of a contracted with T1/T2, x1/x2 = the free
// environment legs of T1/T2, o1/o2 = the open auxiliary legs in output order
struct CornerRoles { char k1, k2, x1, x2, o1, o2; };
static CornerRoles corner_roles(const CornerSpec& cs) {
    CornerRoles r{};
    for (const char* p = cs.l1; *p; ++p) { if (strchr(cs.la, *p)) r.k1 = *p; else if (!strchr(cs.lc, *p)) r.x1 = *p; }
    for (const char* p = cs.l2; *p; ++p) { if (strchr(cs.la, *p)) r.k2 = *p; else if (!strchr(cs.lc, *p)) r.x2 = *p; }
    for (const char* p = cs.out; *p; ++p) if (strchr(cs.la, *p)) { if (!r.o1) r.o1 = *p; else r.o2 = *p; }
    return r;
}

static bool corner_fusable(const Engine& e, int kind, const ctmb_site& s) {
    if (is_dl(s)) return false;
    const CornerSpec& cs = CORNERS[kind];
    const CornerRoles r = corner_roles(cs);
    auto D = [&](char c) { return s.dims[1 + (strchr(cs.la, c) - cs.la)]; };
    return dl_corner_supported(D(r.k1), D(r.k2), D(r.o1), D(r.o2), s.dims[0], e.cplx);
}

// Enlarged corners of a list of requests.  Large real corners (D = 8) go through the fused
// double-layer kernel: chain C.T1.T2 -> X (laid out per environment-index pair), then dl_corner_kernel;
// everything else is the four-step contraction chain.
static void corners_run(Engine& e, int chi, const std::vector<CornerReq>& reqs) {
    std::vector<Engine::ChainJob> plain;
    std::vector<size_t> fused;
    for (size_t i = 0; i < reqs.size(); ++i) {
        if (corner_fusable(e, reqs[i].kind, *reqs[i].site)) fused.push_back(i);
        else plain.push_back(corner_job(reqs[i].kind, *reqs[i].site, chi, reqs[i].out));
    }
    if (!plain.empty()) e.chain_multi(plain);
    const size_t group = 4;                              // X buffers alive at a time (2.1 GB each at D=8, chi=256)
    for (size_t g0 = 0; g0 < fused.size(); g0 += group) {
        const size_t g1 = std::min(fused.size(), g0 + group);
        const size_t mark = e.ws.mark();
        std::vector<Engine::ChainJob> jobs;
        std::vector<DlParams> dl;
        for (size_t g = g0; g < g1; ++g) {
            const CornerReq& rq = reqs[fused[g]];
            const ctmb_site& s = *rq.site;
            const CornerSpec& cs = CORNERS[rq.kind];
            const CornerRoles r = corner_roles(cs);
            auto pos = [&](char c) { return (int)(strchr(cs.la, c) - cs.la); };
            auto D = [&](char c) -> int64_t { return s.dims[1 + pos(c)]; };
            const int64_t dk1 = D(r.k1), dk2 = D(r.k2);
            const char K1 = (char)toupper(r.k1), K2 = (char)toupper(r.k2);
            Engine::ChainJob job;
            job.ops = {env_C(s, cs.c, chi, cs.lc),
                       split_mode(env_T(s, cs.t1, chi, cs.l1), r.k1, r.k1, K1, dk1, dk1),
                       split_mode(env_T(s, cs.t2, chi, cs.l2), r.k2, r.k2, K2, dk2, dk2)};
            job.conj = {false, false, false};
            void* X = e.ws.alloc((size_t)chi * chi * dk1 * dk1 * dk2 * dk2 * e.esize());
            const char xl[7] = {r.x1, r.x2, K1, K2, r.k1, r.k2, 0};
            job.out = make_tn(X, std::string(xl), {chi, chi, dk1, dk2, dk1, dk2});
            jobs.push_back(job);
            // the output view and the strides of a, from the same label machinery the chain uses
            const Tn out = corner_job(rq.kind, s, chi, rq.out).out;
            DlParams p{};
            p.X = (const double*)X; p.a = (const double*)s.a; p.out = (double*)rq.out;
            p.npairs = chi * chi; p.n2 = chi;
            p.st1 = out.str[out.find(r.x1)]; p.st2 = out.str[out.find(r.x2)];
            int64_t astr[5]; astr[4] = 1;
            for (int q = 3; q >= 0; --q) astr[q] = astr[q + 1] * s.dims[q + 1];
            p.as_s = (int)astr[0];
            p.as_k1 = (int)astr[1 + pos(r.k1)]; p.as_k2 = (int)astr[1 + pos(r.k2)];
            p.as_o1 = (int)astr[1 + pos(r.o1)]; p.as_o2 = (int)astr[1 + pos(r.o2)];
            const int do1 = (int)D(r.o1), do2 = (int)D(r.o2);
            const char O1 = (char)toupper(r.o1), O2 = (char)toupper(r.o2);
            for (int i1 = 0; i1 < do1; ++i1)
                for (int i2 = 0; i2 < do2; ++i2) {
                    p.ro[i1 * do2 + i2] = (int)(i1 * out.str[out.find(r.o1)] + i2 * out.str[out.find(r.o2)]);
                    p.co[i1 * do2 + i2] = (int)(i1 * out.str[out.find(O1)] + i2 * out.str[out.find(O2)]);
                }
            dl.push_back(p);
        }
        e.chain_multi(jobs);
        if (!e.ws.dry())
            for (size_t g = 0; g < dl.size(); ++g) {
                const ctmb_site& s = *reqs[fused[g0 + g]].site;
                const double d2 = 64.0, n3 = 128.0;
                ProfScope ps(e, Engine::CAT_GEMM, 2.0 * dl[g].npairs * (d2 * d2 * n3 + d2 * d2 * n3),
                             8.0 * dl[g].npairs * 2.0 * d2 * d2);
                e.flops += 2.0 * dl[g].npairs * 2.0 * d2 * d2 * n3;
                dl_corner_launch(dl[g], 8, 8, s.dims[0], e.stream);
            }
        e.ws.release(mark);
    }
}

// ------------------------------------------------------------------------------------------
// randomised truncated SVD / Hermitian EVD, batched over nb equally-shaped problems
// ------------------------------------------------------------------------------------------
struct Rsvd {
    int nb = 0, m = 0, n = 0, chi = 0, k = 0;
    std::vector<void*> U, V, S;       // U: m x chi col-major, V: n x chi col-major, S: k doubles (sorted)
};

// Sketch width.  Small problems (min(m,n) <= 128) are decomposed EXACTLY (k = min(m,n): the
// "sketch" spans the whole space, no power iteration is needed); this also covers the
// SU(2)-symmetric fixtures of the reference's tests, whose flat, exactly degenerate spectra a
// subspace iteration resolves only slowly.
static int sketch_width(int m, int n, int chi, const ctmb_options& o) {
    int mn = std::min(m, n);
    if (mn <= 160) return mn;
    // rank factor 0 = automatic.  Measured on B200 (profiles/r1_sketch_width.md): where the move is bound by the serial
    // column steps of the QR and the Jacobi rounds (n <= 1024: config c2, 432 x 432) a narrower sketch with one more
    // power iteration is faster (k = 1.75 chi, q = 3: 333 moves/s; 2 chi, q = 2: 296; 1.5 chi, q = 5: 314); where it is
    // bound by the n x n x k products (config c5) the iteration count explodes below 2 chi (1.5 chi: q = 11, 10 % slower).
    // Accuracy does not depend on the choice: the residual test decides the number of iterations.
    double f = o.rsvd_rank_factor;
    if (!(f > 0.0)) f = std::max(m, n) <= 1024 ? 1.75 : 2.0;
    int k = (int)std::ceil(f * chi);
    k = std::max(k, chi + 1);
    return std::min(k, mn);
}

// Blocked Householder QR (explicit thin Q) for sketches that fit neither the registers nor the shared
// memory of one cluster.  Two levels: LEAF panels of width w.b (what one cluster can hold: 8 columns at
// 16384 rows) are factored in WY form by qr_panel_launch and applied only inside their SUPER panel of
// width w.B <= 128; the reflectors of a super panel are then aggregated (T from the Gram matrix of its
// V, wy_tsolve) and applied to the rest of the matrix with three GEMMs  A2 -= V (T^H (V^H A2)).
// Q = H_1 .. H_P E is accumulated backwards per super panel the same way.  (With leaf panels only, the
// 16384 x 512 sketch of config c5 paid 64 full-width trailing updates with K = 8 per QR: 1.4 s per move.)
// V is kept in its own buffer (zeros above each leaf's diagonal), the factor R stays in the upper part of A.
struct QrBlockedWs { PtrBatch Q{}, V{}, Tau{}, Rpp{}, G{}, Tl{}, T{}, W{}, W2{}; int b = 0, B = 0; void* tall = nullptr; };

static void qr_blocked_alloc(Engine& e, QrBlockedWs& w, int bidx, int rows_max, int k, int b) {
    const size_t es = e.esize();
    const int B = std::max(b, std::min(e.cplx ? 64 : 128, k));      // wy_tsolve keeps a B x B matrix in shared memory
    w.b = b; w.B = B;
    const int nl = (k + b - 1) / b, ns = (k + B - 1) / B;
    w.Q.p[bidx] = e.ws.alloc((size_t)rows_max * k * es);
    w.V.p[bidx] = e.ws.alloc((size_t)rows_max * k * es);
    w.Tau.p[bidx] = e.ws.alloc((size_t)k * es);
    w.Rpp.p[bidx] = e.ws.alloc((size_t)nl * b * b * es);
    w.G.p[bidx] = e.ws.alloc((size_t)B * B * es);
    w.Tl.p[bidx] = e.ws.alloc((size_t)b * b * es);
    w.T.p[bidx] = e.ws.alloc((size_t)ns * B * B * es);
    w.W.p[bidx] = e.ws.alloc((size_t)B * k * es);
    w.W2.p[bidx] = e.ws.alloc((size_t)B * k * es);
}

static void qr_blocked(Engine& e, PtrBatch& cur, const PtrBatch& Rout, int nb, int rows, int k, QrBlockedWs& w) {
    const size_t es = e.esize();
    const int b = w.b, B = w.B;
    auto off = [&](const PtrBatch& P, int i, size_t elems) { return (void*)((char*)P.p[i] + elems * es); };
    auto sub = [&](const PtrBatch& P, size_t elems) { PtrBatch r{}; for (int i = 0; i < nb; ++i) r.p[i] = off(P, i, elems); return r; };
    // A2[:, c0:c0+nc] (rows r0..) -= V (T^H (V^H A2)) resp. (T (V^H A2)) with V = Vbuf[r0:, v0:v0+vw], Tm stored [c][s] = T[s,c]
    auto apply = [&](const PtrBatch& Tgt, int r0, int v0, int vw, const PtrBatch& Tm, int c0, int nc, bool adjoint) {
        if (nc <= 0) return;
        const int prow = rows - r0;
        for (int i = 0; i < nb; ++i) {
            Tn V = make_strided(off(w.V, i, (size_t)v0 * rows + r0), "si", {vw, prow}, {rows, 1});
            Tn A2 = make_strided(off(Tgt, i, (size_t)c0 * rows + r0), "ci", {nc, prow}, {rows, 1});
            e.contract(V, true, A2, false, make_tn(w.W.p[i], "cs", {nc, vw}));
        }
        e.flush();
        for (int i = 0; i < nb; ++i) {
            if (adjoint)    // W2[c][t] = sum_s conj(T[s,t]) W[c][s]
                e.contract(make_tn(Tm.p[i], "ts", {vw, vw}), true, make_tn(w.W.p[i], "cs", {nc, vw}), false, make_tn(w.W2.p[i], "ct", {nc, vw}));
            else            // W2[c][t] = sum_s T[t,s] W[c][s]
                e.contract(make_tn(Tm.p[i], "st", {vw, vw}), false, make_tn(w.W.p[i], "cs", {nc, vw}), false, make_tn(w.W2.p[i], "ct", {nc, vw}));
        }
        e.flush();
        for (int i = 0; i < nb; ++i) {
            Tn V = make_strided(off(w.V, i, (size_t)v0 * rows + r0), "ti", {vw, prow}, {rows, 1});
            Tn A2 = make_strided(off(Tgt, i, (size_t)c0 * rows + r0), "ci", {nc, prow}, {rows, 1});
            e.contract(V, false, make_tn(w.W2.p[i], "ct", {nc, vw}), false, A2, nullptr, -1.0, true);
        }
        e.flush();
    };
    // T (stored [c][s]) of the reflectors V[r0:, v0:v0+vw] with the scalars Tau[v0..]
    auto make_T = [&](int r0, int v0, int vw, const PtrBatch& Tm) {
        const int prow = rows - r0;
        for (int i = 0; i < nb; ++i) {                  // G[t][s] = v_s^H v_t
            Tn V = make_strided(off(w.V, i, (size_t)v0 * rows + r0), "si", {vw, prow}, {rows, 1});
            e.contract(V, true, relabel(V, "ti"), false, make_tn(w.G.p[i], "ts", {vw, vw}));
        }
        e.flush();
        ProfScope ps(e, Engine::CAT_QR, 0, 3.0 * es * nb * (double)vw * vw);
        wy_tsolve_launch(w.G, sub(w.Tau, v0), w.G, Tm, nb, vw, 0, e.cplx, e.stream);
    };
    e.flush();
    for (int J0 = 0; J0 < k; J0 += B) {
        const int BW = std::min(B, k - J0), J1 = J0 + BW;
        for (int j0 = J0; j0 < J1; j0 += b) {
            const int bw = std::min(b, J1 - j0), prow = rows - j0;
            PtrBatch Ap = sub(cur, (size_t)j0 * rows + j0), Rp = sub(w.Rpp, (size_t)(j0 / b) * b * b);
            { ProfScope ps(e, Engine::CAT_QR, (e.cplx ? 4.0 : 1.0) * nb * 2.0 * prow * bw * bw, 2.0 * es * nb * (double)prow * bw);
              // tall panels: rows split over the SMs (qr_tall.cu); the cluster kernel takes over where that does not apply
              if (!(w.tall && qr_tall_panel_launch(Ap, Rp, sub(w.Tau, j0), nb, prow, bw, rows, e.cplx, w.tall, e.stream)))
                  qr_panel_launch(Ap, Rp, sub(w.Tau, j0), nb, prow, bw, rows, e.cplx, e.stream); }
            { ProfScope ps(e, Engine::CAT_MISC);          // V buffer <- the leaf's reflectors, zeros above its diagonal (rows J0..)
              qr_copy_v_launch(cur, w.V, nb, rows, J0, j0, bw, e.cplx, e.stream); }
            if (b < B || J1 < k) {
                if (b == B) make_T(j0, j0, bw, sub(w.T, (size_t)(J0 / B) * B * B));
                else { make_T(j0, j0, bw, w.Tl); apply(cur, j0, j0, bw, w.Tl, j0 + bw, J1 - j0 - bw, true); }
            }
        }
        if (b < B) make_T(J0, J0, BW, sub(w.T, (size_t)(J0 / B) * B * B));
        else if (J1 >= k) make_T(J0, J0, BW, sub(w.T, (size_t)(J0 / B) * B * B));
        apply(cur, J0, J0, BW, sub(w.T, (size_t)(J0 / B) * B * B), J1, k - J1, true);
        if (Rout.p[0] != nullptr) {
            ProfScope ps(e, Engine::CAT_MISC);
            for (int j0 = J0; j0 < J1; j0 += b)
                qr_copy_r_launch(cur, sub(w.Rpp, (size_t)(j0 / b) * b * b), Rout, nb, k, j0, std::min(b, J1 - j0), rows, e.cplx, e.stream);
        }
    }
    { ProfScope ps(e, Engine::CAT_MISC); set_identity_launch(w.Q, nb, rows, k, e.cplx, e.stream); }
    for (int J0 = ((k - 1) / B) * B; J0 >= 0; J0 -= B) {
        const int BW = std::min(B, k - J0);
        apply(w.Q, J0, J0, BW, sub(w.T, (size_t)(J0 / B) * B * B), J0, k - J0, false);
    }
    std::swap(cur, w.Q);
}

// Matrix-free form of M = R^T Rt with R = H0 H1, Rt = H2 H3 (the halves: each H is an enlarged corner or its
// transpose):  M = H1^T H0^T H2 H3 is never formed, the range finder applies the four factors to its k-column
// blocks.  At chi*D^2 = 16384 forming R, Rt and M costs 6 n^3 = 2.6e13 FLOP per site, the (2q+2) = 10 operator
// applications cost 10 * 3 extra n x n x k products = 0.8e13: the decomposition of SURVEY 7.1 item 4.
struct FactoredM {
    const void* H[4] = {nullptr, nullptr, nullptr, nullptr};   // physical n x n row-major buffers
    bool tr[4] = {false, false, false, false};                 // logical H = physical^T
    Tn view(int q, char row, char col, int n) const {
        const char lab[3] = {tr[q] ? col : row, tr[q] ? row : col, 0};
        return make_tn(const_cast<void*>(H[q]), lab, {n, n});
    }
};

// Which Ritz value stands for |lambda_{k+1}|, the edge of the interval the Chebyshev filter damps: the last ones of the block
// are the least converged and far BELOW the eigenvalues they approximate (an edge set too low turns the filter into plain
// power iteration); CTMB_CHEB_EDGE = fraction of the way from chi to k (default 0.75).
static int cheb_edge_index(int chi, int k) {
    static double f = -1.0;
    if (f < 0.0) { const char* ev = getenv("CTMB_CHEB_EDGE"); f = ev ? atof(ev) : 0.75; }
    return std::max(chi, std::min(k - 1, chi + (int)(f * (k - chi))));
}

static bool cheb_enabled() {
    static int on = -1;
    if (on < 0) { const char* ev = getenv("CTMB_CHEB"); on = ev ? atoi(ev) : 1; }
    return on != 0;
}

// M[b]: m x n row-major (or, with `fac`, given in factored form). eig_mode: M Hermitian (m == n), S receives signed eigenvalues.
//
// WARM START (`slots`: one tag per problem, or nullptr).  A CTM run decomposes, move after move, a slowly changing matrix
// per (direction, site): the leading kw = k - r Ritz vectors of the previous decomposition of the same slot (right singular
// vectors / eigenvectors, ordered) replace the first kw columns of the Gaussian sketch; the last r columns stay Gaussian, so
// that every direction keeps a generic component in the start block (a direction that is new to the dominant subspace is
// found at the rate of an r-column random sketch and shows up in the residual of the kept triplets until it has converged).
// Only the NUMBER of power iterations changes -- the residual test below accepts exactly the same results -- and near the
// fixed point of a CTM run it drops to zero.  ctmb_options.rsvd_stateless = 1 switches warm start (and the iteration-count
// memory) off; the piecewise entry points (ctmb_truncated_svd, ctmb_projectors, ...) never use it.
static Rsvd rsvd_batch(Engine& e, const std::vector<const void*>& M, int m, int n, int chi,
                       const ctmb_options& o, bool eig_mode, const std::vector<FactoredM>* fac = nullptr,
                       const std::vector<std::string>* slots = nullptr) {
    Rsvd r; r.nb = fac ? (int)fac->size() : (int)M.size(); r.m = m; r.n = n; r.chi = chi;
    const int nb = r.nb, k = sketch_width(m, n, chi, o);
    r.k = k;
    // warm start: kw ordered Ritz vectors are kept per slot (kw = chi when there is none: nothing below changes then)
    const bool warm = slots != nullptr && !o.rsvd_stateless && o.rsvd_tol > 0.0 && k < std::min(m, n) && k > chi + 1;
    const int kw = warm ? std::max(chi + 1, k - std::max(4, k / 8)) : chi;
    CTMB_CHECK(nb <= TC_MAX_BATCH, "too many problems in one batch");
    CTMB_CHECK(chi <= std::min(m, n), "chi exceeds matrix size");
    const size_t es = e.esize();
    // Gaussian sketch, generated once per (n,k,dtype,seed) and kept
    char key[128];
    snprintf(key, sizeof key, "omega:%d:%d:%d:%llu", n, k, (int)e.cplx, o.seed);
    void* omega = nullptr;
    if (!e.ws.dry()) {
        bool created = false;
        omega = e.persistent(key, (size_t)n * k * es, &created);
        if (created) { ProfScope ps(e, Engine::CAT_MISC); fill_gaussian_launch((double*)omega, (long long)n * k * (e.cplx ? 2 : 1), o.seed, e.stream); }
    }
    std::vector<Tn> Mt(nb);
    PtrBatch pY{}, pZ{}, pQ{}, pR{}, pNull{}, pW{}, pSig{}, pS{}, pUh{}, pWs{}, pG{}, pX{}, pTau{}, pF1{}, pF2{}, pC1{}, pC2{};
    std::vector<void*> R2(nb), W(nb), sig(nb), Uh(nb), Ws(nb);
    const int mx = std::max(m, n);
    const bool wy = qr_wy_supported(m, k, e.cplx) && qr_wy_supported(n, k, e.cplx);
    // neither the register / WY path nor one cluster holds the whole sketch: blocked factorisation
    int pw = std::min(std::min(qr_panel_width(m, k, e.cplx), qr_panel_width(n, k, e.cplx)), e.cplx ? 64 : 128);
    const bool blocked = !wy && pw < k;
    QrBlockedWs qbw;
    if (blocked) {
        // tall sketches (config 5: 16384 x 512): 32-column leaf panels factored with the rows split over the SMs
        const int tw = std::min(qr_tall_panel_width(nb, m, k, e.cplx), qr_tall_panel_width(nb, n, k, e.cplx));
        if (tw > pw || (tw > 0 && qr_tall_mode() == 2)) {
            pw = tw;
            if (!e.ws.dry()) qbw.tall = e.persistent("qrtall", qr_tall_scratch_bytes(TC_MAX_BATCH));
        }
    }
    CTMB_CHECK(!blocked || pw >= 4, "sketch too tall for the panel kernels");
    for (int b = 0; b < nb; ++b) {
        if (!fac) Mt[b] = make_tn(const_cast<void*>(M[b]), "ij", {m, n});
        // (the QR drivers swap these with their Q scratch, so all of them are sized for the taller side)
        pY.p[b] = e.ws.alloc((size_t)k * ((wy || blocked) ? mx : m) * es);       // column-major m x k
        pZ.p[b] = e.ws.alloc((size_t)k * ((wy || blocked) ? mx : n) * es);       // column-major n x k
        if (wy) {
            pQ.p[b] = e.ws.alloc((size_t)k * mx * es);  // scratch for the explicit Q of the WY form
            pG.p[b] = e.ws.alloc((size_t)k * k * es);
            pX.p[b] = e.ws.alloc((size_t)k * k * es);
            pTau.p[b] = e.ws.alloc((size_t)k * es);
        }
        if (blocked) qr_blocked_alloc(e, qbw, b, mx, k, pw);
        if (fac) { pF1.p[b] = e.ws.alloc((size_t)k * n * es); pF2.p[b] = e.ws.alloc((size_t)k * n * es); }
        if (eig_mode && k < n) {      // three-term recurrence of the Chebyshev filter (Hermitian branch)
            pC1.p[b] = e.ws.alloc((size_t)k * mx * es); pC2.p[b] = e.ws.alloc((size_t)k * mx * es);
        }
        R2[b] = e.ws.alloc((size_t)k * k * es);
        W[b] = e.ws.alloc((size_t)k * k * es);
        sig[b] = e.ws.alloc((size_t)k * 8);
        r.S.push_back(e.ws.alloc((size_t)k * 8));
        // (warm start: kw >= chi columns are gathered / formed on the side that is remembered; callers use the first chi)
        Uh[b] = e.ws.alloc((size_t)k * kw * es);
        Ws[b] = e.ws.alloc((size_t)k * kw * es);
        r.U.push_back(e.ws.alloc((size_t)m * (eig_mode ? kw : chi) * es));
        r.V.push_back(eig_mode ? nullptr : e.ws.alloc((size_t)n * kw * es));
        pR.p[b] = R2[b]; pW.p[b] = W[b]; pSig.p[b] = sig[b];
        pS.p[b] = r.S[b]; pUh.p[b] = Uh[b]; pWs.p[b] = Ws[b];
    }
    if (e.ws.dry()) return r;
    auto tnY = [&](int b, const char* lab) { return make_tn(pY.p[b], lab, {k, m}); };
    auto tnZ = [&](int b, const char* lab) { return make_tn(pZ.p[b], lab, {k, n}); };
    const double cf = e.cplx ? 4.0 : 1.0;
    // Y[s,i] = sum_j M[i,j] X[s,j]  (adjoint: Z[s,j] = sum_i conj(M[i,j]) Y[s,i]) for `cols` vectors per matrix
    auto apply_op = [&](const PtrBatch& in_full, const PtrBatch& out_full, int cols_full, bool adjoint) {
        // group mode: this rank applies the operator to its slab of columns (vectors are contiguous: column-major
        // storage), then the slabs are all-gathered in place; every member ends with the full, identical result
        const bool split = e.coll_active() && cols_full % e.coll_n == 0 && !e.ws.dry();
        const int cols = split ? cols_full / e.coll_n : cols_full;
        const int64_t in_len = adjoint ? m : n, out_len = adjoint ? n : m;
        PtrBatch in = in_full, out = out_full;
        if (split)
            for (int b = 0; b < nb; ++b) {
                in.p[b] = (char*)in_full.p[b] + (size_t)e.coll_rank * cols * in_len * es;
                out.p[b] = (char*)out_full.p[b] + (size_t)e.coll_rank * cols * out_len * es;
            }
        auto gather = [&]() {
            if (!split) return;
            for (int b = 0; b < nb; ++b) e.allgather(out_full.p[b], (size_t)cols * out_len * es);
        };
        if (!fac) {
            for (int b = 0; b < nb; ++b)
                e.contract(Mt[b], adjoint, make_tn(in.p[b], adjoint ? "si" : "sj", {cols, adjoint ? m : n}), false,
                           make_tn(out.p[b], adjoint ? "sj" : "si", {cols, adjoint ? n : m}));
            e.flush();
            gather();
            return;
        }
        // M = H1^T H0^T H2 H3 :  M X = H1^T (H0^T (H2 (H3 X))),   M^H Y = H3^H (H2^H (conj(H0) (conj(H1) Y)))
        for (int step = 0; step < 4; ++step) {
            for (int b = 0; b < nb; ++b) {
                const FactoredM& f = (*fac)[b];
                void* src = step == 0 ? in.p[b] : (step & 1 ? pF1.p[b] : pF2.p[b]);
                void* dst = step == 3 ? out.p[b] : (step & 1 ? pF2.p[b] : pF1.p[b]);
                Tn X = make_tn(src, "sx", {cols, n});
                Tn Yo = make_tn(dst, "sy", {cols, n});
                Tn H;
                if (!adjoint) {
                    // H3, H2 act as (y,x); H0^T, H1^T act as (x,y) of the stored factor
                    if (step == 0) H = f.view(3, 'y', 'x', n);
                    else if (step == 1) H = f.view(2, 'y', 'x', n);
                    else if (step == 2) H = f.view(0, 'x', 'y', n);
                    else H = f.view(1, 'x', 'y', n);
                } else {
                    if (step == 0) H = f.view(1, 'y', 'x', n);
                    else if (step == 1) H = f.view(0, 'y', 'x', n);
                    else if (step == 2) H = f.view(2, 'x', 'y', n);
                    else H = f.view(3, 'x', 'y', n);
                }
                e.contract(H, adjoint, X, false, Yo);
            }
            e.flush();
        }
        gather();
    };
    // orthonormalise the columns of the matrices in `cur` (rows x k, column-major); on return `cur`
    // holds the explicit thin Q (the buffers of `cur` and pQ are swapped in the WY form)
    auto qr = [&](PtrBatch& cur, const PtrBatch& Rout, int rows) {
        e.flush();
        const double fl = cf * nb * 4.0 * ((double)rows * k * k - (double)k * k * k / 3.0);
        if (blocked) { qr_blocked(e, cur, Rout, nb, rows, k, qbw); return; }
        if (!wy) {
            ProfScope ps(e, Engine::CAT_QR, fl, 2.0 * e.esize() * nb * (double)rows * k);
            qr_launch(cur, Rout, nb, rows, k, rows, e.cplx, e.stream);
            return;
        }
        {
            ProfScope ps(e, Engine::CAT_QR, fl / 2, 2.0 * e.esize() * nb * (double)rows * k);
            qr_wy_factor_launch(cur, Rout, pTau, nb, rows, k, rows, e.cplx, e.stream);
        }
        for (int b = 0; b < nb; ++b) {                  // G[t][s] = v_s^H v_t
            Tn V = make_tn(cur.p[b], "si", {k, rows});
            e.contract(V, true, relabel(V, "ti"), false, make_tn(pG.p[b], "ts", {k, k}));
        }
        e.flush();
        {
            ProfScope ps(e, Engine::CAT_QR, cf * nb * (double)k * k * k / 3.0, 3.0 * e.esize() * nb * (double)k * k);
            wy_tsolve_launch(pG, pTau, cur, pX, nb, k, rows, e.cplx, e.stream);
        }
        for (int b = 0; b < nb; ++b)                    // Q = -V X (+ E below)
            e.contract(make_tn(cur.p[b], "si", {k, rows}), false, make_tn(pX.p[b], "cs", {k, k}), false,
                       make_tn(pQ.p[b], "ci", {k, rows}), nullptr, -1.0);
        e.flush();
        {
            ProfScope ps(e, Engine::CAT_MISC);
            add_identity_launch(pQ, nb, k, rows, e.cplx, e.stream);
        }
        std::swap(cur, pQ);
    };
    if (eig_mode && k == n) {
        // small Hermitian problem decomposed exactly (config c1: n = 64; the 4 x 4 / 16 x 16 reduced density matrices):
        // no sketch at all, the shifted one-sided Jacobi runs on M itself.  Row-major M read as column-major is M^T =
        // conj(M): same eigenvalues, conjugated eigenvectors -- undone at the end in the complex case.
        e.flush();
        // PRECONDITIONED START (slots given, i.e. the C4v move of a CTM run: the same slowly changing matrix every call).  With
        // A the matrix as the kernel reads it and U the eigenvectors of the previous call, the Jacobi runs on G0 = (A + mu) U
        // instead of A + mu: G0 = V S W^H gives A + mu = V S (U W)^H, so V are the eigenvectors whatever U was; if A has hardly
        // changed the columns of G0 are nearly orthogonal already and the sweep count drops from ~11 to 2-3 (config 1 spends
        // 87 % of its move in this 64 x 64 problem).  The orthogonality of V comes from the sweeps, not from U: nothing drifts.
        std::vector<void*> prev(nb, nullptr);
        bool precond = slots != nullptr && !o.rsvd_stateless;
        if (precond)
            for (int b = 0; b < nb; ++b) {
                char sk[160];
                snprintf(sk, sizeof sk, "eigprev:%s:%d:%d", (*slots)[b].c_str(), n, (int)e.cplx);
                bool created = false;
                prev[b] = e.persistent(sk, (size_t)k * k * es, &created);
                if (created) precond = false;               // first call of the slot (the buffers are filled below)
            }
        if (precond) {
            unsigned long long* ssq = (unsigned long long*)e.persistent("eigmu", TC_MAX_BATCH * sizeof(unsigned long long));
            CTMB_CUDA(cudaMemsetAsync(ssq, 0, TC_MAX_BATCH * sizeof(unsigned long long), e.stream));
            ScaleBatch sb{};
            for (int b = 0; b < nb; ++b) { sb.p[b] = const_cast<void*>(M[b]); sb.count[b] = (long long)k * k; sb.amax[b] = ssq + b; }
            { ProfScope ps(e, Engine::CAT_MISC); sumsq_launch(sb, nb, e.cplx, e.stream); }
            for (int b = 0; b < nb; ++b)        // G0[c][i] = sum_j A[i,j] U[j,c] with A[i,j] = M[j,i] (row-major M read column-major)
                e.contract(make_tn(const_cast<void*>(M[b]), "ji", {k, k}), false, make_tn(prev[b], "cj", {k, k}), false,
                           make_tn(R2[b], "ci", {k, k}));
            e.flush();
            for (int b = 0; b < nb; ++b) {
                ProfScope ps(e, Engine::CAT_MISC);
                shift_axpy_launch((double*)R2[b], (const double*)prev[b], ssq + b, 1.0, (long long)k * k * (e.cplx ? 2 : 1), e.stream);
            }
            { ProfScope ps(e, Engine::CAT_JACOBI, 0, 2.0 * es * nb * (double)k * k); jacobi_launch(pR, pNull, pSig, nb, k, e.cplx, o.jacobi_max_sweeps, 0, 0, e.stream); }
            for (int b = 0; b < nb; ++b) {
                ProfScope ps(e, Engine::CAT_MISC);
                shift_axpy_launch((double*)sig[b], nullptr, ssq + b, -1.0, k, e.stream);
            }
        } else {
            for (int b = 0; b < nb; ++b)
                CTMB_CUDA(cudaMemcpyAsync(R2[b], M[b], (size_t)k * k * es, cudaMemcpyDeviceToDevice, e.stream));
            { ProfScope ps(e, Engine::CAT_JACOBI, 0, 2.0 * es * nb * (double)k * k); jacobi_launch(pR, pNull, pSig, nb, k, e.cplx, o.jacobi_max_sweeps, 1, 0, e.stream); }
        }
        if (slots != nullptr && !o.rsvd_stateless)
            for (int b = 0; b < nb; ++b)
                if (prev[b]) CTMB_CUDA(cudaMemcpyAsync(prev[b], R2[b], (size_t)k * k * es, cudaMemcpyDeviceToDevice, e.stream));
        { ProfScope ps(e, Engine::CAT_MISC); sortcols_launch(pR, pNull, pSig, pS, pUh, pNull, nb, k, chi, e.cplx, 1, e.stream); }
        for (int b = 0; b < nb; ++b) {
            CTMB_CUDA(cudaMemcpyAsync(r.U[b], Uh[b], (size_t)k * chi * es, cudaMemcpyDeviceToDevice, e.stream));
            if (e.cplx) { ProfScope ps(e, Engine::CAT_MISC); conj_inplace_launch(r.U[b], (long long)k * chi, e.stream); }
        }
        return r;
    }
    // Y = M * Omega  (Omega: the shared Gaussian sketch, or the slot's warm block [previous Ritz vectors | Gaussian columns])
    std::vector<void*> slot(nb, nullptr);
    bool all_warm = warm;
    if (warm)
        for (int b = 0; b < nb; ++b) {
            char sk[160];
            snprintf(sk, sizeof sk, "warm:%s:%d:%d:%d:%d:%d", (*slots)[b].c_str(), m, n, k, (int)eig_mode, (int)e.cplx);
            bool created = false;
            slot[b] = e.persistent(sk, (size_t)n * k * es, &created);
            if (created) e.drop_persistent("warm:" + (*slots)[b] + ":", sk);     // the slot's basis for another shape, if any
            if (created) {      // first decomposition of this slot: plain Gaussian sketch
                CTMB_CUDA(cudaMemcpyAsync(slot[b], omega, (size_t)n * k * es, cudaMemcpyDeviceToDevice, e.stream));
                all_warm = false;
            }
        }
    { PtrBatch pOm{}; for (int b = 0; b < nb; ++b) pOm.p[b] = warm ? slot[b] : omega; apply_op(pOm, pY, k, false); }
    qr(pY, pNull, m);
    const bool complete = (k == std::min(m, n));          // the sketch spans everything: exact, no iteration
    const bool adaptive = o.rsvd_tol > 0.0 && !complete;
    unsigned long long* dres = nullptr;
    unsigned long long* hres = nullptr;
    if (adaptive) {
        dres = (unsigned long long*)e.persistent("resid", 64 * sizeof(unsigned long long));    // [0..3]: mine, [4..]: the group's
        hres = e.pinned_words();                          // per-engine pinned read-back buffer (32 words)
    }
    int todo = complete ? 0 : o.rsvd_niter;               // full power iterations still to run
    // Adaptive mode remembers, per problem shape, how many iterations passed the residual test last time and starts
    // there (a CTM run decomposes a slowly changing matrix once per move): a failed first round costs a whole extra
    // Rayleigh-Ritz / Jacobi pass.  Every result still satisfies the same residual bound.
    char hkey[96];
    snprintf(hkey, sizeof hkey, "%d:%d:%d:%d:%d:%s", m, n, k, (int)eig_mode, (int)e.cplx, all_warm ? "warm" : "cold");
    const int q_min = all_warm ? 0 : 1;                   // a warm block may need no power iteration at all
    // residual bound: rsvd_tol * sqrt(n) relative to the largest singular value (the rounding floor of the residual
    // itself grows like eps * sqrt(n))
    const double tol_eff = o.rsvd_tol * std::sqrt((double)std::max(m, n));
    // (ctmb_options.rsvd_stateless: every call starts from rsvd_niter, so its result does not depend on the handle's history)
    // Orthogonalisation interval.  Between two QRs the component of a column along the dominant direction grows, relative to
    // the column's own direction j, by S_0 / S_j per operator application; rounding noise (1e-16) injected at the first of m
    // chained applications therefore reaches 1e-16 (S_0/S_j)^(m-1).  With the full dynamic range the projectors use
    // (S_0/S_chi up to 1e8, ctm_projectors.py:266-270) that allows two applications -- the shipped default: one QR per
    // M M^H in the SVD branch, per M M in the Hermitian branch.  Where the kept spectrum is flat (C4v corner of config 3:
    // lambda_0/lambda_chi = 25, so the iteration converges slowly AND needs few orthogonalisations) the measured range of
    // the previous decomposition of the same shape allows more: (m-1) log10(range) <= 12.
    int orth = 2;
    double cheb_a = 0.0;         // Hermitian branch: estimate of |lambda_{k+1}|, the edge of the interval the filter damps
    if (adaptive && !o.rsvd_stateless) {
        auto it = e.iter_hint.find(hkey);
        if (it != e.iter_hint.end()) {
            if (it->second.known) todo = it->second.q;
            if (it->second.range > 1.0) orth = std::max(2, std::min(8, 1 + (int)(12.0 / std::log10(it->second.range * 1.5))));
            if (eig_mode && it->second.known && it->second.tail > 0.0 && cheb_enabled()) {
                cheb_a = it->second.tail;
                // first filtered call of this shape: what is known about failing counts was learnt without the filter
                if (!it->second.cheb_seen) { it->second.cheb_seen = true; it->second.lo = -1; it->second.streak = 2; }
            }
        }
    }
    const int orth_iter = std::max(1, orth / 2);          // SVD branch: whole iterations (two applications each) per QR
    int used = 0;
    double prev_res = -1.0;
    for (int round = 0;; ++round) {
        used += todo;
        if (!eig_mode) {
            // power iterations Q <- orth(M (M^H Q)): ONE orthogonalisation per full iteration.  The
            // intermediate M^H Q is not re-orthogonalised: only directions with S/S0 > 1e-8 survive
            // the projector cut-off (ctm_projectors.py:266-270) and those keep >= 16-2*8 digits through
            // one unorthogonalised M M^H application; measured parity is identical to re-orthogonalising
            // at every half step ("range finder").
            for (int it = 0; it < todo; ++it) {
                apply_op(pY, pZ, k, true);                                                                       // Z = M^H Q
                apply_op(pZ, pY, k, false);                                                                      // Y = M Z
                if ((it + 1) % orth_iter == 0 || it + 1 == todo) qr(pY, pNull, m);
            }
            // Bt = M^H Q = Q2 R2  =>  M ~ Q R2^H Q2^H.  One-sided Jacobi on G = R2^H:  G W = Uh Sigma, so
            // U = Q Uh (normalised columns of the rotated G) and V = Q2 W (accumulated rotations); both come
            // out of the small problem with high relative accuracy down to the projector cut-off.
            apply_op(pY, pZ, k, true);
            qr(pZ, pR, n);
            { ProfScope ps(e, Engine::CAT_JACOBI, 0, 4.0 * e.esize() * nb * (double)k * k); jacobi_launch(pR, pW, pSig, nb, k, e.cplx, o.jacobi_max_sweeps, 0, 1, e.stream); }
            { ProfScope ps(e, Engine::CAT_MISC); sortcols_launch(pR, pW, pSig, pS, pUh, pWs, nb, k, kw, e.cplx, 0, e.stream); }
            for (int b = 0; b < nb; ++b)
                e.contract(tnY(b, "si"), false, make_tn(Uh[b], "cs", {chi, k}), false, make_tn(r.U[b], "ci", {chi, m}));
            e.flush();
            for (int b = 0; b < nb; ++b)         // (warm start: kw >= chi ordered right vectors, the first chi are the result)
                e.contract(tnZ(b, "sj"), false, make_tn(Ws[b], "cs", {kw, k}), false, make_tn(r.V[b], "cj", {kw, n}));
            e.flush();
        } else {
            // Hermitian: subspace iteration with M itself (4 applications per "iteration": the spectrum of
            // the corner decays half as fast as that of M = R^T Rt), one QR per two applications, then
            // Rayleigh-Ritz; the eigenvectors are the normalised columns of the rotated (T + mu).
            if (cheb_a > 0.0 && orth >= 2) {
                // Chebyshev-filtered subspace iteration (the flat spectrum of the C4v corner, lambda_0 / lambda_chi = 25 at config 3,
                // makes plain subspace iteration crawl: |lambda_{k+1} / lambda_chi| per application).  Between two QRs the block
                // is multiplied by T_d(x), x = 2 M^2 / a^2 - 1, a = |lambda_{k+1}| (the k-th Ritz value of the previous
                // decomposition of this shape): everything inside [-a, a] stays bounded by 1, |lambda| > a grows like
                // (x + sqrt(x^2 - 1))^d.  Same number of operator applications (2 d per block, blocks as long as the measured
                // range allows, see `orth`), T_{j+1} = (4 / a^2) M (M T_j) - 2 T_j - T_{j-1}.  The residual test below is the
                // judge; a failed round falls back to plain iteration.
                const long long cnt = (long long)n * k * (e.cplx ? 2 : 1);
                const double ia2 = 1.0 / (cheb_a * cheb_a);
                int napp = 4 * todo;
                const int blk_max = std::max(2, orth & ~1);
                while (napp > 0) {
                    const int deg = std::min(napp, blk_max) / 2;
                    PtrBatch cur = pY, prev{}, nxt = pC1, spare = pC2;
                    for (int j = 0; j < deg; ++j) {
                        for (int b = 0; b < nb; ++b)
                            e.contract(Mt[b], false, make_tn(cur.p[b], "sj", {k, n}), false, make_tn(pZ.p[b], "si", {k, n}));   // Z = M T_j
                        e.flush();
                        for (int b = 0; b < nb; ++b) {
                            ProfScope ps(e, Engine::CAT_MISC);
                            axpby_launch((double*)nxt.p[b], (const double*)cur.p[b], j == 0 ? -1.0 : -2.0,
                                         j == 0 ? nullptr : (const double*)prev.p[b], -1.0, cnt, e.stream);
                        }
                        for (int b = 0; b < nb; ++b)
                            e.contract(Mt[b], false, make_tn(pZ.p[b], "sj", {k, n}), false, make_tn(nxt.p[b], "si", {k, n}), nullptr,
                                       (j == 0 ? 2.0 : 4.0) * ia2, true);                                                          // += c M Z
                        e.flush();
                        // rotate: prev <- cur, cur <- nxt, nxt <- (old prev or the spare buffer)
                        PtrBatch old_prev = prev;
                        prev = cur; cur = nxt;
                        nxt = (j == 0) ? spare : old_prev;
                    }
                    qr(cur, pNull, n);          // (swaps `cur` with its own scratch where it forms Q out of place)
                    pY = cur; pC1 = prev; pC2 = nxt;       // T_{d-1} and the rotation's free buffer go back to the pool
                    napp -= 2 * deg;
                }
            } else
            for (int it = 0; it < 4 * todo; ++it) {
                for (int b = 0; b < nb; ++b) e.contract(Mt[b], false, tnY(b, "sj"), false, tnZ(b, "si"));
                if ((it + 1) % orth == 0 || it + 1 == 4 * todo) qr(pZ, pNull, n);
                else e.flush();
                std::swap(pY, pZ);
            }
            for (int b = 0; b < nb; ++b) e.contract(Mt[b], false, tnY(b, "sj"), false, tnZ(b, "si"));   // Z = M Q
            e.flush();
            for (int b = 0; b < nb; ++b)
                e.contract(tnY(b, "si"), true, tnZ(b, "ti"), false, make_tn(R2[b], "ts", {k, k}));       // Tm = Q^H Z
            e.flush();
            { ProfScope ps(e, Engine::CAT_JACOBI, 0, 2.0 * e.esize() * nb * (double)k * k); jacobi_launch(pR, pNull, pSig, nb, k, e.cplx, o.jacobi_max_sweeps, 1, 0, e.stream); }
            { ProfScope ps(e, Engine::CAT_MISC); sortcols_launch(pR, pNull, pSig, pS, pUh, pNull, nb, k, kw, e.cplx, 1, e.stream); }
            for (int b = 0; b < nb; ++b)
                e.contract(tnY(b, "si"), false, make_tn(Uh[b], "cs", {kw, k}), false, make_tn(r.U[b], "ci", {kw, m}));
            e.flush();
        }
        if (!adaptive) break;
        // residual of the kept triplets -- evaluated after EVERY round, the last one included, so that a result that misses
        // the bound is never returned silently (Engine::rsvd_status); pZ is free at this point and serves as scratch for M X
        PtrBatch pMX{}, pYv{};
        for (int b = 0; b < nb; ++b) {
            pMX.p[b] = pZ.p[b];
            pYv.p[b] = r.U[b];
        }
        { PtrBatch pXin{}; for (int b = 0; b < nb; ++b) pXin.p[b] = eig_mode ? r.U[b] : r.V[b]; apply_op(pXin, pMX, chi, false); }
        // [0] residual, [1] range S_0 / S_chi, [2] |S_{k-1}| (the smallest Ritz value of the block: edge of the Chebyshev filter)
        CTMB_CUDA(cudaMemsetAsync(dres, 0, 4 * sizeof(unsigned long long), e.stream));
        { ProfScope ps(e, Engine::CAT_MISC); resid_launch(pMX, pYv, pS, nb, m, chi, cheb_edge_index(chi, k), o.svd_reltol, dres, e.cplx, e.stream); }
        int nwords = 4;
        if (e.coll_active()) {
            // every member of the group must take the same decisions below: exchange residual and range (rounding may
            // differ between devices through non-deterministic reduction orders, and a split decision would dead-lock)
            CTMB_CHECK(e.coll_n <= 7, "group too large");
            CTMB_CUDA(cudaMemcpyAsync(dres + 4 + 4 * e.coll_rank, dres, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, e.stream));
            e.allgather(dres + 4, 4 * sizeof(unsigned long long));
            nwords = 4 * e.coll_n;
            CTMB_CUDA(cudaMemcpyAsync(hres, dres + 4, nwords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e.stream));
        } else
            CTMB_CUDA(cudaMemcpyAsync(hres, dres, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e.stream));
        CTMB_CUDA(cudaStreamSynchronize(e.stream));
        double res = 0.0, range = 0.0, tail = 0.0;
        for (int g = 0; g < nwords / 4; ++g) {
            double rg, sg, tg; memcpy(&rg, hres + 4 * g, sizeof rg); memcpy(&sg, hres + 4 * g + 1, sizeof sg); memcpy(&tg, hres + 4 * g + 2, sizeof tg);
            res = std::max(res, rg); range = std::max(range, sg); tail = std::max(tail, tg);
        }
        static int dbg = -1;
        if (dbg < 0) { const char* ev = getenv("CTMB_DEBUG_RESID"); dbg = ev ? atoi(ev) : 0; }
        if (dbg) fprintf(stderr, "[ctmb] rsvd %dx%d k=%d round %d iterations %d (QR every %d applications%s) residual %.3e (tol %.1e) range %.2e\n", m, n, k, round, used, orth, cheb_a > 0.0 ? ", Chebyshev filter" : "", res, tol_eff, range);
        Engine::IterHint& hint = e.iter_hint[hkey];
        hint.range = range;
        // a filtered extra round that makes little progress: the edge estimate may sit above wanted eigenvalues -- plain
        // iteration for the rest of this call and for the next calls of this shape
        if (cheb_a > 0.0 && res > tol_eff && round >= 1 && prev_res > 0.0 && res > 0.1 * prev_res) { cheb_a = 0.0; hint.cheb_pause = 16; }
        if (hint.cheb_pause > 0) { --hint.cheb_pause; hint.tail = 0.0; } else hint.tail = tail;
        ++e.rsvd_status.checks;
        if (res <= tol_eff) {
            if (round == 0) {
                // passed first time: probe fewer iterations next time -- two fewer with a margin of 64x, one with 8x.  The
                // residual has a rounding floor (eps sqrt(n) kappa: 0.3 - 0.5 of the bound at configs 3 and 5), so a margin
                // may never show although the count is far too high (a count doubled after a near miss stayed doubled: config
                // 3 ran 18 iterations where 9 - 10 pass).  Hence, after three such passes in a row, bisect towards the largest
                // count known to fail; a failed probe costs one short extra round (below) and is remembered in `lo`.
                if (++hint.age > 64) { hint.lo = -1; hint.age = 0; }
                int dec = res <= tol_eff / 64.0 ? 2 : (res <= 0.125 * tol_eff ? 1 : 0);
                if (dec == 0 && ++hint.streak >= 3 && used - hint.lo >= 2) { dec = std::max(1, (used - hint.lo) / 2); hint.streak = 0; }
                if (dec > 0) hint.streak = 0;
                while (dec > 0 && !(used - dec >= q_min && used - dec > hint.lo)) --dec;
                hint.q = used - dec;
            } else {
                // passed after an extra round: the decay rate measured between the last two rounds says how many of the
                // extra iterations were needed (a later call starts there instead of walking down one by one)
                int need = used;
                if (prev_res > res && res > 0.0 && todo > 0) {
                    const double per_iter = std::log(prev_res / res) / todo;             // decades (natural log) per iteration
                    const int spare = (int)std::floor(std::log(0.25 * tol_eff / res) / per_iter);
                    need = std::max(used - todo + 1, used - std::max(0, spare));
                }
                hint.q = need;
            }
            hint.known = true;
            break;
        }
        if (round == 0) { hint.lo = std::max(hint.lo, used); hint.age = 0; hint.streak = 0; }
        // rounding floor: more iterations do not help (less than 15 % gained per iteration of the last round)
        const bool floor_reached = prev_res > 0.0 && todo > 0 && res > prev_res * std::pow(0.85, (double)todo);
        const bool out_of_rounds = round + 1 >= std::max(1, o.rsvd_max_rounds);
        if (floor_reached || out_of_rounds) {
            // the result is returned although it misses the bound: say so (ctmb_get_rsvd_status), and remember the count
            // so that later calls do not walk through all the rounds again
            hint.q = floor_reached ? std::max(1, used - todo) : used;
            hint.known = true;
            ++e.rsvd_status.missed;
            e.rsvd_status.worst_ratio = std::max(e.rsvd_status.worst_ratio, res / tol_eff);
            if (dbg) fprintf(stderr, "[ctmb] rsvd %dx%d: residual %.3e stays above the bound %.1e (%s)\n", m, n, res, tol_eff,
                             floor_reached ? "rounding floor" : "rsvd_max_rounds");
            break;
        }
        // next round: from the second round on the measured decay rate predicts the missing iterations; until then double --
        // or, after a near miss (within 32x of the bound), add a third
        int next = res <= 32.0 * tol_eff ? std::max(1, (used + 2) / 3) : std::max(1, used);
        if (prev_res > res && res > 0.0 && todo > 0) {
            const double per_iter = std::log(prev_res / res) / todo;
            const int want = (int)std::ceil(std::log(res / (0.25 * tol_eff)) / per_iter);
            next = std::max(1, std::min(want, used));
        }
        prev_res = res;
        todo = next;
    }
    ++e.rsvd_status.calls;
    e.rsvd_status.iterations += used;
    if (warm)        // remember the ordered Ritz vectors (before the caller's sign fixing / scaling touches the first chi of them)
        for (int b = 0; b < nb; ++b)
            CTMB_CUDA(cudaMemcpyAsync(slot[b], eig_mode ? r.U[b] : r.V[b], (size_t)n * kw * es, cudaMemcpyDeviceToDevice, e.stream));
    return r;
}

// Thin Householder QR of one rows x k matrix (column-major, rows >= k) with its own scratch: A <- explicit Q, Rout <- R
// (k x k column-major, upper triangular, LAPACK sign convention: diagonal = -sign(alpha) ||x||).  Same drivers as the
// range finder above (register / cluster kernels, WY form, blocked factorisation), chosen by shape.
static void qr_thin(Engine& e, void* A, void* Rout, int rows, int k) {
    CTMB_CHECK(rows >= k && k >= 1, "qr: expects rows >= cols >= 1");
    const size_t es = e.esize();
    const size_t mark = e.ws.mark();
    const bool wy = qr_wy_supported(rows, k, e.cplx);
    int pw = std::min(qr_panel_width(rows, k, e.cplx), e.cplx ? 64 : 128);
    const bool blocked = !wy && pw < k;
    PtrBatch cur{}, pR{}, pQ{}, pG{}, pX{}, pTau{};
    cur.p[0] = A; pR.p[0] = Rout;
    QrBlockedWs qbw;
    if (blocked) {
        const int tw = qr_tall_panel_width(1, rows, k, e.cplx);
        if (tw > pw || (tw > 0 && qr_tall_mode() == 2)) {
            pw = tw;
            if (!e.ws.dry()) qbw.tall = e.persistent("qrtall", qr_tall_scratch_bytes(TC_MAX_BATCH));
        }
        CTMB_CHECK(pw >= 4, "matrix too tall for the panel kernels");
        qr_blocked_alloc(e, qbw, 0, rows, k, pw);
    }
    if (wy) {
        pQ.p[0] = e.ws.alloc((size_t)k * rows * es);
        pG.p[0] = e.ws.alloc((size_t)k * k * es);
        pX.p[0] = e.ws.alloc((size_t)k * k * es);
        pTau.p[0] = e.ws.alloc((size_t)k * es);
    }
    if (!e.ws.dry()) {
        e.flush();
        if (blocked) qr_blocked(e, cur, pR, 1, rows, k, qbw);
        else if (!wy) { ProfScope ps(e, Engine::CAT_QR); qr_launch(cur, pR, 1, rows, k, rows, e.cplx, e.stream); }
        else {
            { ProfScope ps(e, Engine::CAT_QR); qr_wy_factor_launch(cur, pR, pTau, 1, rows, k, rows, e.cplx, e.stream); }
            Tn V = make_tn(cur.p[0], "si", {k, rows});
            e.contract(V, true, relabel(V, "ti"), false, make_tn(pG.p[0], "ts", {k, k}));
            e.flush();
            { ProfScope ps(e, Engine::CAT_QR); wy_tsolve_launch(pG, pTau, cur, pX, 1, k, rows, e.cplx, e.stream); }
            e.contract(make_tn(cur.p[0], "si", {k, rows}), false, make_tn(pX.p[0], "cs", {k, k}), false,
                       make_tn(pQ.p[0], "ci", {k, rows}), nullptr, -1.0);
            e.flush();
            { ProfScope ps(e, Engine::CAT_MISC); add_identity_launch(pQ, 1, k, rows, e.cplx, e.stream); }
            std::swap(cur, pQ);
        }
        if (cur.p[0] != A)          // the drivers that form Q out of place leave it in their scratch
            CTMB_CUDA(cudaMemcpyAsync(A, cur.p[0], (size_t)rows * k * es, cudaMemcpyDeviceToDevice, e.stream));
    }
    e.ws.release(mark);
}

static ProjFinalizeArgs finalize_args(const Rsvd& r, const ctmb_options& o, bool conj_u, bool scale) {
    ProjFinalizeArgs a{};
    a.nb = r.nb; a.rowsU = r.m; a.rowsV = r.n; a.chi = r.chi; a.kavail = r.k;
    a.reltol = o.svd_reltol; a.eps_multiplet = o.eps_multiplet; a.abstol = o.multiplet_abstol;
    a.truncating = (r.chi < std::min(r.m, r.n)) && (r.k > r.chi);
    a.conj_u = conj_u; a.apply_scale = scale;
    a.v_div_sigma = 0;
    return a;
}

// ------------------------------------------------------------------------------------------
// projectors for a list of jobs
// ------------------------------------------------------------------------------------------
struct MoveCtx {
    Engine& e; int dir, nsites, chi; const ctmb_site* sites; const ctmb_options& o;
};

static double g_m_noise = 0.0;                              // ctmb_debug_set_m_noise
// R,Rt (n0 x n1 row-major) -> P, Pt (n0 x chi row-major)
static void projectors_from_matrices(Engine& e, const std::vector<const void*>& R, const std::vector<const void*>& Rt,
                                     int n0, int n1, int chi, const ctmb_options& o,
                                     const std::vector<void*>& P, const std::vector<void*>& Pt,
                                     const std::vector<double*>& Sout, bool trR = false, bool trRt = false,
                                     const std::vector<std::string>* slots = nullptr) {
    const int nb = (int)R.size();
    const size_t mark = e.ws.mark();
    std::vector<const void*> Mp(nb);
    std::vector<Tn> Rtn(nb), Rttn(nb);
    for (int b = 0; b < nb; ++b) {
        // (trR / trRt: the buffer holds the transpose -- the 4X2 method passes enlarged corners as they are stored)
        Rtn[b] = trR ? make_tn(const_cast<void*>(R[b]), "ik", {n1, n0}) : make_tn(const_cast<void*>(R[b]), "ki", {n0, n1});
        Rttn[b] = trRt ? make_tn(const_cast<void*>(Rt[b]), "jk", {n1, n0}) : make_tn(const_cast<void*>(Rt[b]), "kj", {n0, n1});
        Tn M = e.temp("ij", {n1, n1});
        Mp[b] = M.ptr;
        e.contract(Rtn[b], false, Rttn[b], false, M);           // M = R^T Rt  (plain transpose)
    }
    e.flush();
    if (g_m_noise > 0.0 && !e.ws.dry()) {
        // tests only (ctmb_debug_set_m_noise): unstructured noise of `g_m_noise` x max|M| on every entry of the explicit M,
        // i.e. what one more rounding of M would do -- measures how well the reference algorithm's own output is defined
        unsigned long long* slot = (unsigned long long*)e.persistent("mnoise", TC_MAX_BATCH * sizeof(unsigned long long));
        CTMB_CUDA(cudaMemsetAsync(slot, 0, TC_MAX_BATCH * sizeof(unsigned long long), e.stream));
        ScaleBatch sb{};
        for (int b = 0; b < nb; ++b) { sb.p[b] = const_cast<void*>(Mp[b]); sb.count[b] = (long long)n1 * n1; sb.amax[b] = slot + b; }
        absmax_launch(sb, nb, e.cplx, e.stream);
        for (int b = 0; b < nb; ++b)
            add_noise_launch((double*)const_cast<void*>(Mp[b]), (long long)n1 * n1 * (e.cplx ? 2 : 1), g_m_noise, slot + b,
                             0x1234567ull + b, e.stream);
    }
    Rsvd r = rsvd_batch(e, Mp, n1, n1, chi, o, false, nullptr, slots);
    if (!e.ws.dry()) {
        PtrBatch pU{}, pV{}, pS{}, pSo{};
        for (int b = 0; b < nb; ++b) { pU.p[b] = r.U[b]; pV.p[b] = r.V[b]; pS.p[b] = r.S[b]; pSo.p[b] = Sout.empty() ? nullptr : Sout[b]; }
        { ProfScope ps(e, Engine::CAT_MISC); proj_finalize_launch(pU, pV, pS, pSo, finalize_args(r, o, true, true), e.cplx, e.stream); }
        for (int b = 0; b < nb; ++b) {
            e.contract(relabel(Rtn[b], trR ? "ix" : "xi"), false, make_tn(r.U[b], "ci", {chi, n1}), false, make_tn(P[b], "xc", {n0, chi}));
            e.contract(relabel(Rttn[b], trRt ? "ix" : "xi"), false, make_tn(r.V[b], "ci", {chi, n1}), false, make_tn(Pt[b], "xc", {n0, chi}));
        }
        e.flush();
    }
    e.ws.release(mark);
}

static void halves_jobs(Engine& e, int dir, int chi, const std::vector<const ctmb_site*>& corners /*4 per job*/,
                        const std::vector<void*>& R, const std::vector<void*>& Rt, int64_t& n0, int64_t& n1) {
    const int nj = (int)R.size();
    const HalfSpec& hs = HALVES[dir];
    const size_t mark = e.ws.mark();
    std::vector<CornerReq> cj;
    std::vector<void*> cm(4 * nj);
    std::vector<int64_t> rows(4 * nj), cols(4 * nj);
    for (int j = 0; j < nj; ++j)
        for (int q = 0; q < 4; ++q) {
            const ctmb_site& s = *corners[4 * j + q];
            corner_shape(hs.kind[q], s, chi, rows[4 * j + q], cols[4 * j + q]);
            cm[4 * j + q] = e.ws.alloc((size_t)rows[4 * j + q] * cols[4 * j + q] * e.esize());
            cj.push_back(CornerReq{hs.kind[q], &s, cm[4 * j + q]});
        }
    corners_run(e, chi, cj);
    for (int j = 0; j < nj; ++j) {
        for (int h = 0; h < 2; ++h) {
            const int q0 = 4 * j + 2 * h, q1 = q0 + 1;
            Tn X = make_tn(cm[q0], hs.tr[2 * h] ? "ki" : "ik", {rows[q0], cols[q0]});
            Tn Yt = make_tn(cm[q1], hs.tr[2 * h + 1] ? "jk" : "kj", {rows[q1], cols[q1]});
            const int64_t r0 = hs.tr[2 * h] ? cols[q0] : rows[q0];
            const int64_t c1 = hs.tr[2 * h + 1] ? rows[q1] : cols[q1];
            if (j == 0 && h == 0) { n0 = r0; n1 = c1; }
            CTMB_CHECK(r0 == n0 && c1 == n1, "non-uniform bond dimensions across the unit cell are not supported");
            e.contract(X, false, Yt, false, make_tn(h == 0 ? R[j] : Rt[j], "ij", {r0, c1}));
        }
    }
    e.flush();
    e.ws.release(mark);
    // NOTE: R/Rt were allocated by the caller *before* this function took its mark.
}

static void half_shape(int dir, int chi, const ctmb_site* const* c4, int64_t& n0, int64_t& n1) {
    const HalfSpec& hs = HALVES[dir];
    int64_t r, c;
    corner_shape(hs.kind[0], *c4[0], chi, r, c); n0 = hs.tr[0] ? c : r;
    corner_shape(hs.kind[1], *c4[1], chi, r, c); n1 = hs.tr[1] ? r : c;
}

// matrix-free projectors: when D^2 > 20 the ten factored operator applications are cheaper than forming R, Rt, M
static int g_matrix_free_mode = -1;                         // CTMB_MATRIX_FREE = 0 never, 1 by the FLOP model (default), 2 always
static bool use_matrix_free(int64_t n0, int64_t n1, int chi) {
    int& mode = g_matrix_free_mode;
    if (mode < 0) { const char* ev = getenv("CTMB_MATRIX_FREE"); mode = ev ? atoi(ev) : 1; }
    if (mode == 0 || n0 != n1) return false;
    if (mode == 2) return true;
    return n0 >= 2048 && n0 > 20 * (int64_t)chi;
}

static void move_projectors(MoveCtx& mc, const int* corner_site, const std::vector<int>& jobs,
                            const std::vector<void*>& P, const std::vector<void*>& Pt) {
    Engine& e = mc.e;
    const int nj = (int)jobs.size();
    if (nj == 0) return;
    const size_t mark = e.ws.mark();
    std::vector<const ctmb_site*> corners(4 * nj);
    for (int j = 0; j < nj; ++j)
        for (int q = 0; q < 4; ++q) {
            int si = corner_site[4 * jobs[j] + q];
            CTMB_CHECK(si >= 0 && si < mc.nsites, "corner_site index out of range");
            corners[4 * j + q] = &mc.sites[si];
        }
    int64_t n0 = 0, n1 = 0;
    half_shape(mc.dir, mc.chi, &corners[0], n0, n1);
    CTMB_CHECK(mc.o.projector_method == 0 || mc.o.projector_method == 1, "Invalid Projector method");
    // warm-start slots of the decompositions: one per (direction, projector method, site job)
    std::vector<std::string> slots(nj);
    for (int j = 0; j < nj; ++j) {
        char t[48]; snprintf(t, sizeof t, "d%d:p%d:j%d", mc.dir, mc.o.projector_method, jobs[j]);
        slots[j] = t;
    }
    if (mc.o.projector_method == 1) {
        // '4X2' (ctm_projectors.py:66-136): R and Rt are the first enlarged corner of each half, transposed as there
        const HalfSpec& hs = HALVES[mc.dir];
        std::vector<CornerReq> cj;
        std::vector<const void*> Rc(nj), Rtc(nj);
        for (int j = 0; j < nj; ++j)
            for (int q = 0; q < 4; q += 2) {
                const ctmb_site& s = *corners[4 * j + q];
                int64_t rows, cols;
                corner_shape(hs.kind[q], s, mc.chi, rows, cols);
                CTMB_CHECK(rows == n0 && cols == n0, "the 4X2 projectors need equal bond dimensions");
                void* cm = e.ws.alloc((size_t)rows * cols * e.esize());
                cj.push_back(CornerReq{hs.kind[q], &s, cm});
                (q == 0 ? Rc[j] : Rtc[j]) = cm;
            }
        corners_run(e, mc.chi, cj);
        projectors_from_matrices(e, Rc, Rtc, (int)n0, (int)n0, mc.chi, mc.o, P, Pt, {}, hs.tr[0], hs.tr[2], &slots);
        e.ws.release(mark);
        return;
    }
    if (use_matrix_free(n0, n1, mc.chi)) {
        const HalfSpec& hs = HALVES[mc.dir];
        const int n = (int)n0, chi = mc.chi;
        std::vector<CornerReq> cj;
        std::vector<FactoredM> fac(nj);
        for (int j = 0; j < nj; ++j)
            for (int q = 0; q < 4; ++q) {
                const ctmb_site& s = *corners[4 * j + q];
                int64_t rows, cols;
                corner_shape(hs.kind[q], s, chi, rows, cols);
                CTMB_CHECK(rows == n && cols == n, "non-uniform bond dimensions across the unit cell are not supported");
                void* cm = e.ws.alloc((size_t)n * n * e.esize());
                cj.push_back(CornerReq{hs.kind[q], &s, cm});
                fac[j].H[q] = cm; fac[j].tr[q] = hs.tr[q];
            }
        corners_run(e, chi, cj);
        Rsvd r = rsvd_batch(e, {}, n, n, chi, mc.o, false, &fac, &slots);
        std::vector<void*> T1(nj), T2(nj);
        for (int j = 0; j < nj; ++j) { T1[j] = e.ws.alloc((size_t)chi * n * e.esize()); T2[j] = e.ws.alloc((size_t)chi * n * e.esize()); }
        if (!e.ws.dry()) {
            PtrBatch pU{}, pV{}, pS{}, pSo{};
            for (int b = 0; b < nj; ++b) { pU.p[b] = r.U[b]; pV.p[b] = r.V[b]; pS.p[b] = r.S[b]; pSo.p[b] = nullptr; }
            { ProfScope ps(e, Engine::CAT_MISC); proj_finalize_launch(pU, pV, pS, pSo, finalize_args(r, mc.o, true, true), e.cplx, e.stream); }
            // P = R conj(U) s = H0 (H1 Uf),  Pt = Rt V s = H2 (H3 Vf)
            for (int b = 0; b < nj; ++b) {
                e.contract(fac[b].view(1, 'k', 'i', n), false, make_tn(r.U[b], "ci", {chi, n}), false, make_tn(T1[b], "ck", {chi, n}));
                e.contract(fac[b].view(3, 'k', 'i', n), false, make_tn(r.V[b], "ci", {chi, n}), false, make_tn(T2[b], "ck", {chi, n}));
            }
            e.flush();
            for (int b = 0; b < nj; ++b) {
                e.contract(fac[b].view(0, 'x', 'k', n), false, make_tn(T1[b], "ck", {chi, n}), false, make_tn(P[b], "xc", {n, chi}));
                e.contract(fac[b].view(2, 'x', 'k', n), false, make_tn(T2[b], "ck", {chi, n}), false, make_tn(Pt[b], "xc", {n, chi}));
            }
            e.flush();
        }
        e.ws.release(mark);
        return;
    }
    std::vector<void*> R(nj), Rt(nj);
    for (int j = 0; j < nj; ++j) { R[j] = e.ws.alloc((size_t)n0 * n1 * e.esize()); Rt[j] = e.ws.alloc((size_t)n0 * n1 * e.esize()); }
    halves_jobs(e, mc.dir, mc.chi, corners, R, Rt, n0, n1);
    std::vector<const void*> Rc(R.begin(), R.end()), Rtc(Rt.begin(), Rt.end());
    projectors_from_matrices(e, Rc, Rtc, (int)n0, (int)n1, mc.chi, mc.o, P, Pt, {}, false, false, &slots);
    e.ws.release(mark);
}

static void move_absorb(MoveCtx& mc, const int* nb_site, const std::vector<int>& jobs,
                        const std::vector<const void*>& Pall, const std::vector<const void*>& Ptall,
                        void* const* nC1, void* const* nC2, void* const* nT) {
    Engine& e = mc.e;
    const int nj = (int)jobs.size();
    if (nj == 0) return;
    CTMB_CHECK(mc.o.norm_type == 0 || mc.o.norm_type == 1, "norm_type must be 0 ('inf') or 1 (2-norm)");
    const bool norm2 = mc.o.norm_type == 1;       // the slots then hold sum |x|^2, filled after the chains
    const AbsorbSpec& as = ABSORB[mc.dir];
    const int chi = mc.chi;
    unsigned long long* amax = nullptr;
    if (!e.ws.dry()) {
        amax = (unsigned long long*)e.persistent("amax", 3 * TC_MAX_BATCH * sizeof(unsigned long long));
        CTMB_CUDA(cudaMemsetAsync(amax, 0, 3 * TC_MAX_BATCH * sizeof(unsigned long long), e.stream));
    }
    CTMB_CHECK(3 * nj <= TC_MAX_BATCH, "too many sites for one absorb batch");
    std::vector<Engine::ChainJob> j1, j2, j3;
    ScaleBatch sb{};
    for (int j = 0; j < nj; ++j) {
        const int si = jobs[j], sn = nb_site[si];
        CTMB_CHECK(sn >= 0 && sn < mc.nsites, "nb_site index out of range");
        const ctmb_site& s = mc.sites[si];
        auto dimd = [&](int which) { int64_t d[3]; t_dims(s, which, chi, d); return which == 1 ? d[2] : which == 2 ? d[0] : d[1]; };
        const int64_t d1 = dimd(as.T1), d2 = dimd(as.T2);
        Tn Pt1 = make_tn(const_cast<void*>(Ptall[sn]), as.lPt1, {chi, d1, chi});
        Tn P2 = make_tn(const_cast<void*>(Pall[si]), as.lP2, {chi, d2, chi});
        // nC1
        Engine::ChainJob a1;
        a1.ops = {env_C(s, as.C1, chi, as.lC1), env_T(s, as.T1, chi, as.lT1), Pt1};
        a1.conj = {false, false, false};
        a1.out = make_tn(nC1[j], as.oC1, {chi, chi});
        a1.amax = (amax && !norm2) ? amax + 3 * j : nullptr;
        j1.push_back(a1);
        Engine::ChainJob a2;
        a2.ops = {env_C(s, as.C2, chi, as.lC2), env_T(s, as.T2, chi, as.lT2), P2};
        a2.conj = {false, false, false};
        a2.out = make_tn(nC2[j], as.oC2, {chi, chi});
        a2.amax = (amax && !norm2) ? amax + 3 * j + 1 : nullptr;
        j2.push_back(a2);
        // nT: the projector on the T1 side is P1 / Pt1 of the neighbour, on the T2 side P2 / Pt2 of this site
        const void* pa = as.pa_is_P1 ? Pall[sn] : Ptall[si];
        const void* pb = as.pa_is_P1 ? Ptall[si] : Pall[sn];
        const int64_t da = as.pa_is_P1 ? d1 : d2, db = as.pa_is_P1 ? d2 : d1;
        std::vector<Tn> ops = {env_T(s, as.T, chi, as.lT), make_tn(const_cast<void*>(pa), as.lPa, {chi, da, chi}),
                               make_tn(const_cast<void*>(pb), as.lPb, {chi, db, chi})};
        j3.push_back(sl_job(ops, 2, as.lA, s, as.oT, nT[j], (amax && !norm2) ? amax + 3 * j + 2 : nullptr));
        int64_t td[3]; t_dims(s, as.T, chi, td);
        if (!amax) continue;
        sb.p[3 * j] = nC1[j]; sb.count[3 * j] = (long long)chi * chi; sb.amax[3 * j] = amax + 3 * j;
        sb.p[3 * j + 1] = nC2[j]; sb.count[3 * j + 1] = (long long)chi * chi; sb.amax[3 * j + 1] = amax + 3 * j + 1;
        sb.p[3 * j + 2] = nT[j]; sb.count[3 * j + 2] = td[0] * td[1] * td[2]; sb.amax[3 * j + 2] = amax + 3 * j + 2;
    }
    // nC1 and nC2 chains have identical GEMM shapes: run them as one batch
    std::vector<Engine::ChainJob> j12 = j1;
    j12.insert(j12.end(), j2.begin(), j2.end());
    e.chain_multi(j12);
    e.chain_multi(j3);
    if (!e.ws.dry()) {
        if (norm2) { ProfScope ps(e, Engine::CAT_MISC); sumsq_launch(sb, 3 * nj, e.cplx, e.stream); }
        { ProfScope ps(e, Engine::CAT_MISC); scale_by_amax_launch(sb, 3 * nj, e.cplx, e.stream, norm2); }
    }
}

// ------------------------------------------------------------------------------------------
// reduced density matrix of a 2x2 plaquette (ctm/generic/rdm.py:1306-1592, strategy of rdm2x2_legacy):
// four enlarged corners -- with open physical legs at the sites kept open -- upper = LU.RU, lower = LD.RD, rho = upper.lower
// ------------------------------------------------------------------------------------------
static int g_rdm_block_rows = 0;                            // ctmb_debug_set_rdm_block_rows (tests: force the blocked trace)
static void rdm2x2_impl(Engine& e, int chi, const ctmb_site* const s4[4], int open_mask, void* rho) {
    static const int kinds[4] = {CTMB_LU, CTMB_RU, CTMB_LD, CTMB_RD};       // s0 s1 / s2 s3
    CTMB_CHECK(chi > 0 && s4 && open_mask > 0 && open_mask < 16, "bad arguments (at least one site must stay open)");
    const size_t mark = e.ws.mark();
    void* cm[4]; int64_t rows[4], cols[4], pd[4];
    std::vector<CornerReq> closed;
    std::vector<Engine::ChainJob> open;
    for (int q = 0; q < 4; ++q) {
        const ctmb_site& s = *s4[q];
        corner_shape(kinds[q], s, chi, rows[q], cols[q]);
        const bool op = (open_mask >> q) & 1;
        pd[q] = op ? s.dims[0] : 1;
        CTMB_CHECK(!op || s.dims[0] > 0, "open sites need the single-layer on-site tensor");
        cm[q] = e.ws.alloc((size_t)rows[q] * cols[q] * pd[q] * pd[q] * e.esize());
        if (!op) { closed.push_back(CornerReq{kinds[q], &s, cm[q]}); continue; }
        const CornerSpec& cs = CORNERS[kinds[q]];
        std::vector<Tn> ops = {env_C(s, cs.c, chi, cs.lc), env_T(s, cs.t1, chi, cs.l1), env_T(s, cs.t2, chi, cs.l2)};
        open.push_back(sl_job(ops, 3, cs.la, s, cs.out, cm[q], nullptr, true));
    }
    if (!closed.empty()) corners_run(e, chi, closed);
    if (!open.empty()) e.chain_multi(open);
    CTMB_CHECK(cols[0] == rows[1] && cols[2] == cols[3] && rows[0] == rows[2] && cols[1] == rows[3],
               "the four enlarged corners of the plaquette do not fit together");
    // matrix views [row, col, s, S]: LU "ab iI", RU "bc jJ", LD "ad kK", RD "cd lL"  (oracle rdm2x2: same strings)
    auto view = [&](int q, char r, char c, char k, char b) {
        std::string lab{r, c}; std::vector<int64_t> d{rows[q], cols[q]};
        if (pd[q] > 1 || ((open_mask >> q) & 1)) { lab.push_back(k); lab.push_back(b); d.push_back(pd[q]); d.push_back(pd[q]); }
        return make_tn(cm[q], lab, d);
    };
    const Tn LU = view(0, 'a', 'b', 'i', 'I'), RU = view(1, 'b', 'c', 'j', 'J');
    const Tn LD = view(2, 'a', 'd', 'k', 'K'), RD = view(3, 'c', 'd', 'l', 'L');
    auto half = [&](const Tn& X, const Tn& Y, char r, char c) {
        std::string lab{r, c}; std::vector<int64_t> d{X.dim[X.find(r)], Y.dim[Y.find(c)]};
        for (const Tn* t : {&X, &Y}) for (int q = 2; q < t->nd; ++q) { lab.push_back(t->idx[q]); d.push_back(t->dim[q]); }
        return e.temp(lab, d);
    };
    // rho[kets..., bras...] in the site order s0 s1 s2 s3
    std::string lab; std::vector<int64_t> d;
    for (int pass = 0; pass < 2; ++pass)
        for (int q = 0; q < 4; ++q)
            if ((open_mask >> q) & 1) { lab.push_back(pass ? "IJKL"[q] : "ijkl"[q]); d.push_back(pd[q]); }
    // The halves carry the open physical legs: rows x cols x p^4 elements each.  At config-5 size (n = 16384, p = 2: 4.3e9)
    // that exceeds the 32-bit offset tables of the contraction kernel, so the row index `a` is processed in blocks:
    // upper_blk = LU[a_blk] . RU, lower_blk = LD[a_blk] . RD, rho += upper_blk . lower_blk  (round 1 refused this size).
    const int64_t per_row = std::max(cols[1] * pd[0] * pd[0] * pd[1] * pd[1], rows[3] * pd[2] * pd[2] * pd[3] * pd[3]);   // elements of one row of upper / lower
    int64_t blk = rows[0];
    if (g_rdm_block_rows > 0) blk = std::min<int64_t>(blk, g_rdm_block_rows);
    while (blk > 1 && blk * per_row >= ((int64_t)1 << 31) / 2) blk = (blk + 1) / 2;
    if (blk >= rows[0]) {
        Tn upper = half(LU, RU, 'a', 'c'), lower = half(LD, RD, 'a', 'c');
        e.contract(LU, false, RU, false, upper);
        e.contract(LD, false, RD, false, lower);
        e.flush();
        e.contract(upper, false, lower, false, make_tn(rho, lab, d));
        e.flush();
    } else {
        const size_t es = e.esize();
        const size_t bmark = e.ws.mark();
        for (int64_t a0 = 0; a0 < rows[0]; a0 += blk) {
            const int64_t ab = std::min(blk, rows[0] - a0);
            auto rows_of = [&](const Tn& t, int q) {          // rows a0 .. a0+ab of the corner matrix q (row-major, phys legs minor)
                Tn v = t;
                v.ptr = e.ws.dry() ? nullptr : (void*)((char*)t.ptr + (size_t)a0 * cols[q] * pd[q] * pd[q] * es);
                v.dim[0] = ab;
                return v;
            };
            e.ws.release(bmark);
            const Tn LUb = rows_of(LU, 0), LDb = rows_of(LD, 2);
            Tn upper = half(LUb, RU, 'a', 'c'), lower = half(LDb, RD, 'a', 'c');
            e.contract(LUb, false, RU, false, upper);
            e.contract(LDb, false, RD, false, lower);
            e.flush();
            e.contract(upper, false, lower, false, make_tn(rho, lab, d), nullptr, 1.0, a0 > 0);
            e.flush();
        }
        e.ws.release(bmark);
    }
    e.ws.release(mark);
}

// One- and two-site density matrices (ctm/generic/rdm.py: rdm1x1_dl :114-258, rdm2x1_dl :352-500, rdm1x2_dl :672-826), built
// from OPEN enlarged corners closed by the remaining edge tensors; the strings are those of oracle/ctm_oracle.py
// (rdm1x1 / rdm2x1 / rdm1x2), which is pinned element-wise against the reference.   kind: 0 = 1x1, 1 = 2x1, 2 = 1x2.
static void rdm_small_impl(Engine& e, int kind, int chi, const ctmb_site* const s2[2], void* rho) {
    CTMB_CHECK(kind >= 0 && kind <= 2 && chi > 0 && s2 && s2[0] && (kind == 0 || s2[1]), "bad arguments");
    const size_t mark = e.ws.mark();
    const ctmb_site& s0 = *s2[0];
    const ctmb_site& s1 = kind == 0 ? s0 : *s2[1];
    CTMB_CHECK(s0.dims[0] > 0 && s1.dims[0] > 0, "density matrices need the single-layer on-site tensors");
    const int64_t p0 = s0.dims[0], p1 = s1.dims[0];
    auto open_corner = [&](int ck, const ctmb_site& s, int64_t& rows, int64_t& cols) {
        corner_shape(ck, s, chi, rows, cols);
        void* buf = e.ws.alloc((size_t)rows * cols * s.dims[0] * s.dims[0] * e.esize());
        const CornerSpec& cs = CORNERS[ck];
        std::vector<Tn> ops = {env_C(s, cs.c, chi, cs.lc), env_T(s, cs.t1, chi, cs.l1), env_T(s, cs.t2, chi, cs.l2)};
        std::vector<Engine::ChainJob> jobs = {sl_job(ops, 3, cs.la, s, cs.out, buf, nullptr, true)};
        e.chain_multi(jobs);
        return buf;
    };
    // edge pieces: chains of plain environment tensors into a contiguous buffer with the given output labels
    auto edge = [&](std::vector<Tn> ops, const char* out_lab) {
        std::map<char, int64_t> ext;
        for (const Tn& t : ops) for (int d = 0; d < t.nd; ++d) ext[t.idx[d]] = t.dim[d];
        std::vector<int64_t> od; size_t count = 1;
        for (const char* q = out_lab; *q; ++q) { od.push_back(ext[*q]); count *= (size_t)ext[*q]; }
        void* buf = e.ws.alloc(count * e.esize());
        Engine::ChainJob job;
        job.ops = ops; job.conj.assign(ops.size(), false);
        job.out = make_tn(buf, std::string(out_lab), od);
        std::vector<Engine::ChainJob> jobs = {job};
        e.chain_multi(jobs);
        return job.out;
    };
    int64_t r0, c0;
    void* LU = open_corner(CTMB_LU, s0, r0, c0);                              // [(l,D),(x,R),s,S]
    if (kind == 0) {
        Tn E = edge({env_C(s0, 1, chi, "xa"), env_T(s0, 3, chi, "aRb"), env_C(s0, 2, chi, "bc"), env_T(s0, 2, chi, "Dec"),
                     env_C(s0, 3, chi, "le")}, "lDxR");
        CTMB_CHECK(E.dim[0] * E.dim[1] == r0 && E.dim[2] * E.dim[3] == c0, "the edge of the site does not fit its enlarged corner");
        // (dummy mode w of extent 1: every output label comes from one operand, the second contributes none otherwise)
        e.contract(make_tn(LU, "yzsS", {r0, c0, p0, p0}), false, make_tn(E.ptr, "yzw", {r0, c0, 1}), false,
                   make_tn(rho, "sSw", {p0, p0, 1}));
    } else if (kind == 1) {
        int64_t r1, c1;
        void* RU = open_corner(CTMB_RU, s1, r1, c1);                          // [(e,l),(c,f),j,J]
        Tn B0 = edge({env_C(s0, 3, chi, "le"), env_T(s0, 2, chi, "Dec")}, "lDc");
        Tn B1 = edge({env_T(s1, 2, chi, "fEb"), env_C(s1, 2, chi, "cb")}, "cfE");
        CTMB_CHECK(B0.dim[0] * B0.dim[1] == r0 && B1.dim[0] * B1.dim[1] == c1 && c0 == r1 && B0.dim[2] == B1.dim[2],
                   "the two sites of the 2x1 patch do not fit together");
        Tn L = e.temp("czsS", {B0.dim[2], c0, p0, p0});
        Tn R = e.temp("tEjJ", {r1, B1.dim[2], p1, p1});
        e.contract(make_tn(LU, "yzsS", {r0, c0, p0, p0}), false, make_tn(B0.ptr, "yc", {r0, B0.dim[2]}), false, L);
        e.contract(make_tn(RU, "tqjJ", {r1, c1, p1, p1}), false, make_tn(B1.ptr, "qE", {c1, B1.dim[2]}), false, R);
        e.flush();
        e.contract(L, false, relabel(R, "zcjJ"), false, make_tn(rho, "sjSJ", {p0, p1, p0, p1}));
    } else {
        int64_t r1, c1;
        void* LD = open_corner(CTMB_LD, s1, r1, c1);                          // [(c,u),(e,r),k,K]
        Tn Rr = edge({env_C(s0, 1, chi, "xa"), env_T(s0, 3, chi, "aRb")}, "xRb");
        Tn Rb = edge({env_T(s1, 3, chi, "Brc"), env_C(s1, 2, chi, "ce")}, "erB");
        CTMB_CHECK(Rr.dim[0] * Rr.dim[1] == c0 && Rb.dim[0] * Rb.dim[1] == c1 && r0 == r1 && Rr.dim[2] == Rb.dim[2],
                   "the two sites of the 1x2 patch do not fit together");
        Tn Tp = e.temp("ybsS", {r0, Rr.dim[2], p0, p0});
        Tn Bt = e.temp("tBkK", {r1, Rb.dim[2], p1, p1});
        e.contract(make_tn(LU, "yzsS", {r0, c0, p0, p0}), false, make_tn(Rr.ptr, "zb", {c0, Rr.dim[2]}), false, Tp);
        e.contract(make_tn(LD, "tqkK", {r1, c1, p1, p1}), false, make_tn(Rb.ptr, "qB", {c1, Rb.dim[2]}), false, Bt);
        e.flush();
        e.contract(Tp, false, relabel(Bt, "ybkK"), false, make_tn(rho, "skSK", {p0, p1, p0, p1}));
    }
    e.flush();
    e.ws.release(mark);
}

// _sym_pos_def_matrix (ctm/generic/rdm.py:38-57)
static void sym_pos_def_impl(Engine& e, const void* raw, int n, int sym_pos_def, const ctmb_options& o, void* out) {
    CTMB_CHECK(n > 0, "bad arguments");
    if (!e.ws.dry()) { ProfScope ps(e, Engine::CAT_MISC); rdm_herm_launch(raw, out, n, 1, e.cplx, e.stream); }
    if (!sym_pos_def) return;
    CTMB_CHECK(n <= 160, "sym_pos_def=True needs the complete eigendecomposition: supported for matrices up to 160 x 160");
    const size_t mark = e.ws.mark();
    Rsvd r = rsvd_batch(e, {out}, n, n, n, o, true);
    if (!e.ws.dry()) { ProfScope ps(e, Engine::CAT_MISC); rdm_posdef_launch(out, r.U[0], (const double*)r.S[0], n, e.cplx, e.stream); }
    e.ws.release(mark);
}

static ctmb_options opts_or_default(const ctmb_options* o) {
    ctmb_options d; ctmb_default_options(&d);
    return o ? *o : d;
}

}  // namespace ctmb

// =============================================================================================
// C ABI
// =============================================================================================
using namespace ctmb;

struct ctmb_handle_s { Handle h; explicit ctmb_handle_s(int dev) : h(dev) {} };

#define CTMB_TRY try {
#define CTMB_CATCH(ret)                                                   \
    } catch (const std::exception& ex) { set_error(ex.what()); return ret; } \
      catch (...) { set_error("unknown error"); return ret; }

static void begin_call(ctmb_handle_t h, ctmb_dtype dt, void* ws, size_t ws_bytes, void* stream) {
    CTMB_CHECK(h != nullptr, "null handle");
    CTMB_CHECK(dt == CTMB_F64 || dt == CTMB_C128, "unsupported dtype");
    Engine& e = h->h.eng;
    CTMB_CHECK(e.device() >= 0, "planning-only handle (device -1): no CUDA device, only the *_workspace queries are available");
    CTMB_CUDA(cudaSetDevice(e.device()));
    e.cplx = (dt == CTMB_C128);
    e.stream = (cudaStream_t)stream;
    e.ws.reset(ws, ws_bytes);
    e.drop_pending();          // a previous call that threw between contract() and flush() must not leak its batch into this one
}
// dry run: measure the workspace the same call would use
static void begin_dry(ctmb_handle_t h, ctmb_dtype dt) {
    CTMB_CHECK(h != nullptr, "null handle");
    Engine& e = h->h.eng;
    e.cplx = (dt == CTMB_C128);
    e.ws.reset(nullptr, 0);
    e.drop_pending();
}

extern "C" {

int ctmb_version(void) { return 100; }
// diagnostics for tools/ (not part of include/ctmb.h): Jacobi sweeps and matrices since the last call
void ctmb_debug_jacobi_stats(unsigned long long* out) { ctmb::jacobi_stats(out); }
// tests: force (2) / forbid (0) / auto (1) the matrix-free projector path regardless of the problem size
void ctmb_debug_set_matrix_free(int mode) { ctmb::g_matrix_free_mode = mode; }
// tests: process the rows of the plaquette's halves in blocks of this many rows (0 = only where the offsets demand it)
void ctmb_debug_set_rdm_block_rows(int rows) { ctmb::g_rdm_block_rows = rows; }
// tests: relative amplitude of Gaussian noise added to the explicit M = R^T Rt before it is decomposed (0 = off)
void ctmb_debug_set_m_noise(double amp) { ctmb::g_m_noise = amp; }
// tests: 1 if square halves of extent n with environment dimension chi take the matrix-free projector path
int ctmb_debug_uses_matrix_free(int n, int chi) { return ctmb::use_matrix_free(n, n, chi) ? 1 : 0; }
const char* ctmb_last_error(void) { return get_error().c_str(); }

int ctmb_create(ctmb_handle_t* h, int device) {
    CTMB_TRY
    CTMB_CHECK(h != nullptr, "null out pointer");
    if (device == -1) {          // planning-only handle: the *_workspace queries work (no kernel is ever launched), compute calls fail
        *h = new ctmb_handle_s(-1);
        return 0;
    }
    int ndev = 0;
    CTMB_CUDA(cudaGetDeviceCount(&ndev));
    CTMB_CHECK(device >= 0 && device < ndev, "no such CUDA device");
    cudaDeviceProp prop;
    CTMB_CUDA(cudaGetDeviceProperties(&prop, device));
    CTMB_CHECK(prop.major >= 10, "libctmb is built for sm_100a (Blackwell) only");
    *h = new ctmb_handle_s(device);
    return 0;
    CTMB_CATCH(-1)
}

int ctmb_destroy(ctmb_handle_t h) {
    CTMB_TRY
    delete h;
    return 0;
    CTMB_CATCH(-1)
}

void ctmb_default_options(ctmb_options* o) {
    if (!o) return;
    o->svd_reltol = 1.0e-8; o->eps_multiplet = 1.0e-8; o->multiplet_abstol = 1.0e-14;
    o->rsvd_rank_factor = 0.0; o->rsvd_niter = 4; o->jacobi_max_sweeps = 40; o->norm_type = 0; o->rsvd_max_rounds = 5;
    o->seed = 0x5eed5eedull; o->rsvd_tol = 2.0e-15; o->projector_method = 0; o->rsvd_stateless = 0;
}

int ctmb_set_group(ctmb_handle_t h, int rank, int nranks, ctmb_allgather_fn fn, void* ctx) {
    CTMB_TRY
    CTMB_CHECK(h != nullptr, "null handle");
    CTMB_CHECK(nranks >= 1 && nranks <= 16 && rank >= 0 && rank < nranks, "bad group");
    CTMB_CHECK(nranks == 1 || fn != nullptr, "a group needs an all-gather callback");
    Engine& e = h->h.eng;
    e.coll_rank = rank; e.coll_n = nranks; e.coll_fn = nranks > 1 ? fn : nullptr; e.coll_ctx = ctx;
    return 0;
    CTMB_CATCH(-1)
}

int ctmb_get_rsvd_status(ctmb_handle_t h, long long* checks, long long* missed, double* worst_ratio, long long* calls,
                         long long* iterations, int reset) {
    CTMB_TRY
    CTMB_CHECK(h != nullptr, "null handle");
    Engine& e = h->h.eng;
    if (checks) *checks = e.rsvd_status.checks;
    if (missed) *missed = e.rsvd_status.missed;
    if (worst_ratio) *worst_ratio = e.rsvd_status.worst_ratio;
    if (calls) *calls = e.rsvd_status.calls;
    if (iterations) *iterations = e.rsvd_status.iterations;
    if (reset) e.rsvd_status = Engine::RsvdStatus{};
    return 0;
    CTMB_CATCH(-1)
}

int ctmb_get_counters(ctmb_handle_t h, long long* launches, double* flops) {
    CTMB_TRY
    CTMB_CHECK(h != nullptr, "null handle");
    if (launches) *launches = h->h.eng.launches;
    if (flops) *flops = h->h.eng.flops;
    return 0;
    CTMB_CATCH(-1)
}
int ctmb_reset_counters(ctmb_handle_t h) {
    CTMB_TRY
    CTMB_CHECK(h != nullptr, "null handle");
    h->h.eng.launches = 0; h->h.eng.flops = 0;
    h->h.eng.prof_reset();
    return 0;
    CTMB_CATCH(-1)
}
int ctmb_profile_enable(ctmb_handle_t h, int on) {
    CTMB_TRY
    CTMB_CHECK(h != nullptr, "null handle");
    h->h.eng.profiling = on != 0;
    return 0;
    CTMB_CATCH(-1)
}
int ctmb_profile_get(ctmb_handle_t h, double* ms, double* flops, double* bytes, long long* launches) {
    CTMB_TRY
    CTMB_CHECK(h != nullptr, "null handle");
    Engine::ProfTotals t[Engine::CAT_COUNT];
    h->h.eng.prof_collect(t);
    for (int i = 0; i < Engine::CAT_COUNT; ++i) {
        if (ms) ms[i] = t[i].ms;
        if (flops) flops[i] = t[i].flops;
        if (bytes) bytes[i] = t[i].bytes;
        if (launches) launches[i] = t[i].launches;
    }
    return 0;
    CTMB_CATCH(-1)
}

int ctmb_einsum2(ctmb_handle_t h, ctmb_dtype dt, const char* spec, const void* A, const long long* dimsA, int conjA,
                 const void* B, const long long* dimsB, int conjB, void* C, void* stream) {
    CTMB_TRY
    static char dummy[256];
    begin_call(h, dt, dummy, 0, stream);
    Engine& e = h->h.eng;
    std::string s(spec);
    size_t c1 = s.find(','), ar = s.find("->");
    CTMB_CHECK(c1 != std::string::npos && ar != std::string::npos && c1 < ar, "spec must look like 'ab,bc->ac'");
    std::string la = s.substr(0, c1), lb = s.substr(c1 + 1, ar - c1 - 1), lc = s.substr(ar + 2);
    std::vector<int64_t> da(dimsA, dimsA + la.size()), db(dimsB, dimsB + lb.size()), dc;
    for (char ch : lc) {
        size_t p = la.find(ch);
        if (p != std::string::npos) dc.push_back(da[p]);
        else { p = lb.find(ch); CTMB_CHECK(p != std::string::npos, "output label not in inputs"); dc.push_back(db[p]); }
    }
    e.contract(make_tn(const_cast<void*>(A), la, da), conjA != 0, make_tn(const_cast<void*>(B), lb, db), conjB != 0,
               make_tn(C, lc, dc));
    e.flush();
    return 0;
    CTMB_CATCH(-1)
}

static void c2x2_impl(ctmb_handle_t h, ctmb_corner kind, int chi, const ctmb_site* site, void* out) {
    CTMB_CHECK(kind >= 0 && kind < 4 && site != nullptr && chi > 0, "bad arguments");
    corners_run(h->h.eng, chi, {CornerReq{(int)kind, site, out}});
}
int ctmb_c2x2(ctmb_handle_t h, ctmb_dtype dt, ctmb_corner kind, int chi, const ctmb_site* site, void* out,
              void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    c2x2_impl(h, kind, chi, site, out);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_c2x2_workspace(ctmb_handle_t h, ctmb_dtype dt, ctmb_corner kind, int chi, const ctmb_site* site) {
    CTMB_TRY
    begin_dry(h, dt);
    c2x2_impl(h, kind, chi, site, nullptr);
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

static void halves_impl(ctmb_handle_t h, ctmb_direction dir, int chi, const ctmb_site* const corners[4], void* R, void* Rt) {
    CTMB_CHECK(dir >= 0 && dir < 4 && chi > 0, "bad arguments");
    std::vector<const ctmb_site*> c(corners, corners + 4);
    int64_t n0 = 0, n1 = 0;
    half_shape(dir, chi, corners, n0, n1);
    halves_jobs(h->h.eng, dir, chi, c, {R}, {Rt}, n0, n1);
}
int ctmb_halves(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int chi, const ctmb_site* const corners[4],
                void* R, void* Rt, void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    halves_impl(h, dir, chi, corners, R, Rt);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_halves_workspace(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int chi,
                             const ctmb_site* const corners[4]) {
    CTMB_TRY
    begin_dry(h, dt);
    halves_impl(h, dir, chi, corners, nullptr, nullptr);
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

int ctmb_projectors(ctmb_handle_t h, ctmb_dtype dt, const void* R, const void* Rt, int n0, int n1, int chi,
                    const ctmb_options* opt, void* P, void* Pt, double* S_out, void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    ctmb_options o = opts_or_default(opt);
    std::vector<double*> so; if (S_out) so.push_back(S_out);
    projectors_from_matrices(h->h.eng, {R}, {Rt}, n0, n1, chi, o, {P}, {Pt}, so);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_projectors_workspace(ctmb_handle_t h, ctmb_dtype dt, int n0, int n1, int chi, const ctmb_options* opt) {
    CTMB_TRY
    begin_dry(h, dt);
    ctmb_options o = opts_or_default(opt);
    projectors_from_matrices(h->h.eng, {nullptr}, {nullptr}, n0, n1, chi, o, {nullptr}, {nullptr}, {});
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

static void svd_impl(ctmb_handle_t h, const void* M, int m, int n, int chi, const ctmb_options& o, void* U, double* S, void* V) {
    Engine& e = h->h.eng;
    Rsvd r = rsvd_batch(e, {M}, m, n, chi, o, false);
    if (e.ws.dry()) return;
    PtrBatch pU{}, pV{}, pS{}, pSo{};
    pU.p[0] = r.U[0]; pV.p[0] = r.V[0]; pS.p[0] = r.S[0]; pSo.p[0] = S;
    { ProfScope ps(e, Engine::CAT_MISC); proj_finalize_launch(pU, pV, pS, pSo, finalize_args(r, o, false, false), e.cplx, e.stream); }
    CTMB_CUDA(cudaMemcpyAsync(U, r.U[0], (size_t)m * chi * e.esize(), cudaMemcpyDeviceToDevice, e.stream));
    CTMB_CUDA(cudaMemcpyAsync(V, r.V[0], (size_t)n * chi * e.esize(), cudaMemcpyDeviceToDevice, e.stream));
}
int ctmb_truncated_svd(ctmb_handle_t h, ctmb_dtype dt, const void* M, int m, int n, int chi, const ctmb_options* opt,
                       void* U, double* S, void* V, void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    svd_impl(h, M, m, n, chi, opts_or_default(opt), U, S, V);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_truncated_svd_workspace(ctmb_handle_t h, ctmb_dtype dt, int m, int n, int chi, const ctmb_options* opt) {
    CTMB_TRY
    begin_dry(h, dt);
    svd_impl(h, nullptr, m, n, chi, opts_or_default(opt), nullptr, nullptr, nullptr);
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

int ctmb_qr(ctmb_handle_t h, ctmb_dtype dt, void* A, int rows, int k, void* R, void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    qr_thin(h->h.eng, A, R, rows, k);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_qr_workspace(ctmb_handle_t h, ctmb_dtype dt, int rows, int k) {
    CTMB_TRY
    begin_dry(h, dt);
    qr_thin(h->h.eng, nullptr, nullptr, rows, k);
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

static Rsvd eig_impl(ctmb_handle_t h, const void* M, int n, int chi, const ctmb_options& o, double* D, void* U) {
    Engine& e = h->h.eng;
    Rsvd r = rsvd_batch(e, {M}, n, n, chi, o, true);
    if (e.ws.dry()) return r;
    PtrBatch pU{}, pV{}, pS{}, pSo{};
    pU.p[0] = r.U[0]; pV.p[0] = nullptr; pS.p[0] = r.S[0]; pSo.p[0] = D;
    { ProfScope ps(e, Engine::CAT_MISC); proj_finalize_launch(pU, pV, pS, pSo, finalize_args(r, o, false, false), e.cplx, e.stream); }
    if (U) CTMB_CUDA(cudaMemcpyAsync(U, r.U[0], (size_t)n * chi * e.esize(), cudaMemcpyDeviceToDevice, e.stream));
    return r;
}
int ctmb_truncated_eig_sym(ctmb_handle_t h, ctmb_dtype dt, const void* M, int n, int chi, const ctmb_options* opt,
                           double* D, void* U, void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    eig_impl(h, M, n, chi, opts_or_default(opt), D, U);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_truncated_eig_sym_workspace(ctmb_handle_t h, ctmb_dtype dt, int n, int chi, const ctmb_options* opt) {
    CTMB_TRY
    begin_dry(h, dt);
    eig_impl(h, nullptr, n, chi, opts_or_default(opt), nullptr, nullptr);
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

static void move_generic_impl(ctmb_handle_t h, ctmb_direction dir, int nsites, int chi, const ctmb_site* sites,
                              const int* corner_site, const int* nb_site, const ctmb_options& o,
                              void* const* nC1, void* const* nC2, void* const* nT) {
    CTMB_CHECK(dir >= 0 && dir < 4 && nsites > 0 && chi > 0 && sites && corner_site && nb_site, "bad arguments");
    Engine& e = h->h.eng;
    MoveCtx mc{e, (int)dir, nsites, chi, sites, o};
    std::vector<int> jobs(nsites);
    for (int i = 0; i < nsites; ++i) jobs[i] = i;
    // projector storage: n0 x chi per site, n0 from the first job's patch
    const ctmb_site* c4[4];
    for (int q = 0; q < 4; ++q) c4[q] = &sites[corner_site[q]];
    int64_t n0, n1; half_shape(dir, chi, c4, n0, n1);
    std::vector<void*> P(nsites), Pt(nsites);
    for (int i = 0; i < nsites; ++i) { P[i] = e.ws.alloc((size_t)n0 * chi * e.esize()); Pt[i] = e.ws.alloc((size_t)n0 * chi * e.esize()); }
    // process the projector jobs in groups that fit the batch limits
    const int group = std::max(1, TC_MAX_BATCH / 4);
    for (int j0 = 0; j0 < nsites; j0 += group) {
        std::vector<int> g(jobs.begin() + j0, jobs.begin() + std::min(nsites, j0 + group));
        std::vector<void*> Pg(P.begin() + j0, P.begin() + j0 + g.size()), Ptg(Pt.begin() + j0, Pt.begin() + j0 + g.size());
        move_projectors(mc, corner_site, g, Pg, Ptg);
    }
    std::vector<const void*> Pc(P.begin(), P.end()), Ptc(Pt.begin(), Pt.end());
    const int agroup = TC_MAX_BATCH / 3;
    for (int j0 = 0; j0 < nsites; j0 += agroup) {
        std::vector<int> g(jobs.begin() + j0, jobs.begin() + std::min(nsites, j0 + agroup));
        move_absorb(mc, nb_site, g, Pc, Ptc, nC1 ? nC1 + j0 : nullptr, nC2 ? nC2 + j0 : nullptr, nT ? nT + j0 : nullptr);
    }
}

int ctmb_move_generic(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int nsites, int chi, const ctmb_site* sites,
                      const int* corner_site, const int* nb_site, const ctmb_options* opt, void* const* nC1,
                      void* const* nC2, void* const* nT, void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    move_generic_impl(h, dir, nsites, chi, sites, corner_site, nb_site, opts_or_default(opt), nC1, nC2, nT);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_move_generic_workspace(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int nsites, int chi,
                                   const ctmb_site* sites, const int* corner_site, const int* nb_site,
                                   const ctmb_options* opt) {
    CTMB_TRY
    begin_dry(h, dt);
    std::vector<void*> dummy(nsites, nullptr);
    move_generic_impl(h, dir, nsites, chi, sites, corner_site, nb_site, opts_or_default(opt), dummy.data(), dummy.data(), dummy.data());
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

int ctmb_move_generic_projectors(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int nsites, int chi,
                                 const ctmb_site* sites, const int* corner_site, int njobs, const int* jobs,
                                 const ctmb_options* opt, void* const* P, void* const* Pt, void* ws, size_t ws_bytes,
                                 void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    ctmb_options o = opts_or_default(opt);
    MoveCtx mc{h->h.eng, (int)dir, nsites, chi, sites, o};
    const int group = std::max(1, TC_MAX_BATCH / 4);
    for (int j0 = 0; j0 < njobs; j0 += group) {
        int j1 = std::min(njobs, j0 + group);
        std::vector<int> g(jobs + j0, jobs + j1);
        std::vector<void*> Pg(P + j0, P + j1), Ptg(Pt + j0, Pt + j1);
        move_projectors(mc, corner_site, g, Pg, Ptg);
    }
    return 0;
    CTMB_CATCH(-1)
}

int ctmb_move_generic_absorb(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int nsites, int chi,
                             const ctmb_site* sites, const int* nb_site, int njobs, const int* jobs,
                             const ctmb_options* opt, const void* const* P, const void* const* Pt, void* const* nC1,
                             void* const* nC2, void* const* nT, void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    ctmb_options o = opts_or_default(opt);
    MoveCtx mc{h->h.eng, (int)dir, nsites, chi, sites, o};
    std::vector<const void*> Pc(P, P + nsites), Ptc(Pt, Pt + nsites);
    const int agroup = TC_MAX_BATCH / 3;
    for (int j0 = 0; j0 < njobs; j0 += agroup) {
        int j1 = std::min(njobs, j0 + agroup);
        std::vector<int> g(jobs + j0, jobs + j1);
        move_absorb(mc, nb_site, g, Pc, Ptc, nC1 + j0, nC2 + j0, nT + j0);
    }
    return 0;
    CTMB_CATCH(-1)
}

int ctmb_rdm2x2(ctmb_handle_t h, ctmb_dtype dt, int chi, const ctmb_site* const sites[4], int open_mask, void* rho,
                void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    rdm2x2_impl(h->h.eng, chi, sites, open_mask, rho);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_rdm2x2_workspace(ctmb_handle_t h, ctmb_dtype dt, int chi, const ctmb_site* const sites[4], int open_mask) {
    CTMB_TRY
    begin_dry(h, dt);
    rdm2x2_impl(h->h.eng, chi, sites, open_mask, nullptr);
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

int ctmb_rdm_small(ctmb_handle_t h, ctmb_dtype dt, int kind, int chi, const ctmb_site* const sites[2], void* rho,
                   void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    rdm_small_impl(h->h.eng, kind, chi, sites, rho);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_rdm_small_workspace(ctmb_handle_t h, ctmb_dtype dt, int kind, int chi, const ctmb_site* const sites[2]) {
    CTMB_TRY
    begin_dry(h, dt);
    rdm_small_impl(h->h.eng, kind, chi, sites, nullptr);
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

int ctmb_sym_pos_def(ctmb_handle_t h, ctmb_dtype dt, const void* rdm, int n, int sym_pos_def, void* out,
                     void* ws, size_t ws_bytes, void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    sym_pos_def_impl(h->h.eng, rdm, n, sym_pos_def, opts_or_default(nullptr), out);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_sym_pos_def_workspace(ctmb_handle_t h, ctmb_dtype dt, int n, int sym_pos_def) {
    CTMB_TRY
    begin_dry(h, dt);
    sym_pos_def_impl(h->h.eng, nullptr, n, sym_pos_def, opts_or_default(nullptr), nullptr);
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

static void move_c4v_impl(ctmb_handle_t h, const void* a, const int dims[5], const void* C, const void* T, int chi,
                          const ctmb_options& o, void* C_out, void* T_out, double* D_out) {
    Engine& e = h->h.eng;
    CTMB_CHECK(dims[1] == dims[2] && dims[2] == dims[3] && dims[3] == dims[4], "C4v needs equal bond dimensions");
    ctmb_site s{};
    s.a = a; for (int i = 0; i < 5; ++i) s.dims[i] = dims[i];
    const int64_t d = aux2(s, 0), n = (int64_t)chi * d;
    // enlarged corner 'ab,xbu,ael,@uldr->edxr'  (ctm_components_c4v.py:52-130)
    Tn Ct = make_tn(const_cast<void*>(C), "ab", {chi, chi});
    void* c2 = e.ws.alloc((size_t)n * n * e.esize());
    {
        std::vector<Tn> ops = {Ct, make_tn(const_cast<void*>(T), "xbu", {chi, chi, d}), make_tn(const_cast<void*>(T), "ael", {chi, chi, d})};
        std::vector<Engine::ChainJob> jobs = {sl_job(ops, 3, "uldr", s, "edxr", c2, nullptr)};
        e.chain_multi(jobs);
    }
    const std::vector<std::string> slots = {"c4v"};
    Rsvd r = rsvd_batch(e, {c2}, (int)n, (int)n, chi, o, true, nullptr, &slots);
    double* Dv = (double*)e.ws.alloc((size_t)chi * 8);
    void* nTraw = e.ws.alloc((size_t)chi * chi * d * e.esize());
    if (!e.ws.dry()) {
        PtrBatch pU{}, pV{}, pS{}, pSo{};
        pU.p[0] = r.U[0]; pS.p[0] = r.S[0]; pSo.p[0] = Dv;
        { ProfScope ps(e, Engine::CAT_MISC); proj_finalize_launch(pU, pV, pS, pSo, finalize_args(r, o, false, false), e.cplx, e.stream); }
    }
    // nT = 'acl,aux,@uldr,cdy->xyr' with (T, P, a, conj a, conj P), P = U viewed (chi, d, chi')
    {
        Tn P = make_tn(r.U[0], "xau", {chi, chi, d});        // column-major U: [x][(a,u)]
        Tn Pc = relabel(P, "ycd");
        std::vector<Tn> ops = {make_tn(const_cast<void*>(T), "acl", {chi, chi, d}), P, Pc};
        Engine::ChainJob job = sl_job(ops, 2, "uldr", s, "xyr", nTraw, nullptr);
        job.conj.back() = true;                               // conj(P) closes the chain
        std::vector<Engine::ChainJob> jobs = {job};
        e.chain_multi(jobs);
    }
    if (!e.ws.dry()) {
        CTMB_CHECK(o.norm_type == 0 || o.norm_type == 1, "norm_type must be 0 ('inf') or 1 (2-norm)");
        const bool norm2 = o.norm_type == 1;
        unsigned long long* amax = (unsigned long long*)e.persistent("amax", 3 * TC_MAX_BATCH * sizeof(unsigned long long));
        CTMB_CUDA(cudaMemsetAsync(amax, 0, 2 * sizeof(unsigned long long), e.stream));
        { ProfScope ps(e, Engine::CAT_MISC); c4v_sym_launch(nTraw, T_out, chi, (int)d, amax, e.cplx, e.stream); }
        ScaleBatch sb{}; sb.p[0] = T_out; sb.count[0] = (long long)chi * chi * d; sb.amax[0] = norm2 ? amax + 1 : amax;
        if (norm2) { ProfScope ps(e, Engine::CAT_MISC); sumsq_launch(sb, 1, e.cplx, e.stream); }
        { ProfScope ps(e, Engine::CAT_MISC); scale_by_amax_launch(sb, 1, e.cplx, e.stream, norm2); }
        { ProfScope ps(e, Engine::CAT_MISC); c4v_diag_launch(Dv, C_out, chi, e.cplx, e.stream); }
        if (D_out) CTMB_CUDA(cudaMemcpyAsync(D_out, Dv, (size_t)chi * 8, cudaMemcpyDeviceToDevice, e.stream));
    }
}

int ctmb_move_c4v(ctmb_handle_t h, ctmb_dtype dt, const void* a, const int dims[5], const void* C, const void* T,
                  int chi, const ctmb_options* opt, void* C_out, void* T_out, double* D_out, void* ws, size_t ws_bytes,
                  void* stream) {
    CTMB_TRY
    begin_call(h, dt, ws, ws_bytes, stream);
    ctmb_options o = opts_or_default(opt);
    if (!opt) { o.eps_multiplet = 1.0e-12; }   // truncated_eig_sym defaults (custom_eig.py:7-8)
    move_c4v_impl(h, a, dims, C, T, chi, o, C_out, T_out, D_out);
    return 0;
    CTMB_CATCH(-1)
}
size_t ctmb_move_c4v_workspace(ctmb_handle_t h, ctmb_dtype dt, const int dims[5], int chi, const ctmb_options* opt) {
    CTMB_TRY
    begin_dry(h, dt);
    move_c4v_impl(h, nullptr, dims, nullptr, nullptr, chi, opts_or_default(opt), nullptr, nullptr, nullptr);
    return h->h.eng.ws.peak() + 256;
    CTMB_CATCH(0)
}

}  // extern "C"
