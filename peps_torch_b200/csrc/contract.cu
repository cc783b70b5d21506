// Host-side contraction planner and batcher (see contract.h).
#include "contract.h"
#include <algorithm>
#include <cstring>
#include <sstream>

namespace ctmb {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const std::string& get_error() { return g_err; }

Tn make_tn(void* ptr, const std::string& idx, const std::vector<int64_t>& dims) {
    CTMB_CHECK(idx.size() == dims.size() && idx.size() <= (size_t)MAX_ND, "bad tensor spec");
    Tn t; t.ptr = ptr; t.nd = (int)idx.size();
    int64_t s = 1;
    for (int i = t.nd - 1; i >= 0; --i) { t.idx[i] = idx[i]; t.dim[i] = dims[i]; t.str[i] = s; s *= dims[i]; }
    t.idx[t.nd] = 0;
    return t;
}
Tn make_tn(void* ptr, const char* idx, std::initializer_list<int64_t> dims) {
    return make_tn(ptr, std::string(idx), std::vector<int64_t>(dims));
}

Tn make_strided(void* ptr, const char* idx, std::initializer_list<int64_t> dims, std::initializer_list<int64_t> strides) {
    Tn t = make_tn(ptr, std::string(idx), std::vector<int64_t>(dims));
    CTMB_CHECK(strides.size() == dims.size(), "bad strides");
    int i = 0;
    for (auto s : strides) t.str[i++] = s;
    return t;
}

Tn split_mode(const Tn& t, char c, char c1, char c2, int64_t d1, int64_t d2) {
    int p = t.find(c);
    CTMB_CHECK(p >= 0, "split_mode: label not found");
    CTMB_CHECK(t.dim[p] == d1 * d2, "split_mode: extent mismatch");
    CTMB_CHECK(t.nd + 1 <= MAX_ND, "too many modes");
    Tn r; r.ptr = t.ptr; r.nd = t.nd + 1;
    int o = 0;
    for (int i = 0; i < t.nd; ++i) {
        if (i == p) {
            r.idx[o] = c1; r.dim[o] = d1; r.str[o] = t.str[i] * d2; ++o;
            r.idx[o] = c2; r.dim[o] = d2; r.str[o] = t.str[i]; ++o;
        } else { r.idx[o] = t.idx[i]; r.dim[o] = t.dim[i]; r.str[o] = t.str[i]; ++o; }
    }
    r.idx[r.nd] = 0;
    return r;
}

Tn relabel(const Tn& t, const char* idx) {
    CTMB_CHECK((int)strlen(idx) == t.nd, "relabel: rank mismatch");
    Tn r = t;
    for (int i = 0; i < t.nd; ++i) r.idx[i] = idx[i];
    return r;
}

Tn transpose_view(const Tn& t, const char* order) {
    CTMB_CHECK((int)strlen(order) == t.nd, "transpose_view: rank mismatch");
    Tn r; r.ptr = t.ptr; r.nd = t.nd;
    for (int i = 0; i < t.nd; ++i) {
        int p = t.find(order[i]);
        CTMB_CHECK(p >= 0, "transpose_view: label not found");
        r.idx[i] = order[i]; r.dim[i] = t.dim[p]; r.str[i] = t.str[p];
    }
    r.idx[r.nd] = 0;
    return r;
}

void* Workspace::alloc(size_t bytes, bool back) {
    bytes = (bytes + 255) & ~(size_t)255;
    char* p;
    if (!back) { p = base_ ? base_ + lo_ : (char*)256 + lo_; lo_ += bytes; }
    else { hi_ += bytes; p = base_ ? base_ + cap_ - hi_ : (char*)256; }
    peak_ = std::max(peak_, lo_ + hi_);
    if (base_) CTMB_CHECK(lo_ + hi_ <= cap_, "workspace too small (query ctmb_*_workspace_bytes)");
    return p;
}

Engine::Engine(int device) : device_(device) {
    if (device >= 0) CTMB_CUDA(cudaSetDevice(device));      // device -1: planning-only engine (workspace queries)
}

Engine::~Engine() {
    for (auto& kv : plans_) if (kv.second.dev) cudaFree(kv.second.dev);
    for (auto& kv : persist_) if (kv.second.first) cudaFree(kv.second.first);
    prof_reset();
    for (auto e : ev_pool_) cudaEventDestroy(e);
    if (pinned_) cudaFreeHost(pinned_);
}

unsigned long long* Engine::pinned_words() {
    if (!pinned_) CTMB_CUDA(cudaMallocHost(&pinned_, 32 * sizeof(unsigned long long)));
    return pinned_;
}

void* Engine::persistent(const std::string& key, size_t bytes, bool* created) {
    auto it = persist_.find(key);
    if (it != persist_.end() && it->second.second >= bytes) { if (created) *created = false; return it->second.first; }
    if (it != persist_.end()) { cudaFree(it->second.first); persist_.erase(it); }
    void* p = nullptr;
    CTMB_CUDA(cudaMalloc(&p, bytes));
    persist_[key] = {p, bytes};
    if (created) *created = true;
    return p;
}

void Engine::drop_persistent(const std::string& prefix, const std::string& keep) {
    for (auto it = persist_.begin(); it != persist_.end();) {
        if (it->first != keep && it->first.compare(0, prefix.size(), prefix) == 0) { cudaFree(it->second.first); it = persist_.erase(it); }
        else ++it;
    }
}

Tn Engine::temp(const std::string& idx, const std::vector<int64_t>& dims, bool back) {
    int64_t n = 1; for (auto d : dims) n *= d;
    void* p = ws.alloc((size_t)n * esize(), back);
    return make_tn(p, idx, dims);
}

static void append_spec(std::ostringstream& os, const Tn& t) {
    os << t.idx << ':';
    for (int i = 0; i < t.nd; ++i) os << t.dim[i] << '/' << t.str[i] << ',';
    os << ';';
}

// The label / extent rules of a pairwise contraction, checked without touching the device: the workspace queries
// (dry runs) call this, so a planning-only handle validates every chain of an entry point on a machine without a GPU.
static void validate_contraction(const Tn& A, const Tn& B, const Tn& C) {
    int64_t M = 1, N = 1, K = 1;
    for (int i = 0; i < C.nd; ++i) {
        char c = C.idx[i];
        int pa = A.find(c), pb = B.find(c);
        CTMB_CHECK((pa >= 0) != (pb >= 0), "output label must come from exactly one operand");
        if (pa >= 0) { CTMB_CHECK(A.dim[pa] == C.dim[i], "extent mismatch (A,C)"); M *= C.dim[i]; }
        else { CTMB_CHECK(B.dim[pb] == C.dim[i], "extent mismatch (B,C)"); N *= C.dim[i]; }
        for (int q = 0; q < i; ++q) CTMB_CHECK(C.idx[q] != c, "repeated output label");
    }
    for (int i = 0; i < A.nd; ++i) {
        char c = A.idx[i];
        if (C.find(c) >= 0) continue;
        int pb = B.find(c);
        CTMB_CHECK(pb >= 0, "label of A neither in B nor in the output");
        CTMB_CHECK(A.dim[i] == B.dim[pb], "extent mismatch (A,B)");
        K *= A.dim[i];
    }
    for (int i = 0; i < B.nd; ++i) {
        char c = B.idx[i];
        CTMB_CHECK(C.find(c) >= 0 || A.find(c) >= 0, "label of B neither in A nor in the output");
    }
    CTMB_CHECK(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "GEMM extent overflow");
    CTMB_CHECK(A.numel() < (1ll << 31) && B.numel() < (1ll << 31) && C.numel() < (1ll << 31),
               "operand exceeds 2^31 elements (offset tables are 32-bit)");
}

const Plan& Engine::get_plan(const Tn& A, bool conjA, const Tn& B, bool conjB, const Tn& C) {
    std::ostringstream os;
    append_spec(os, A); append_spec(os, B); append_spec(os, C);
    os << (conjA ? 'c' : 'n') << (conjB ? 'c' : 'n');
    std::string key = os.str();
    auto it = plans_.find(key);
    if (it != plans_.end()) return it->second;

    // classify modes
    struct Mode { char c; int64_t dim, sa, sb, sc; };
    std::vector<Mode> mm, nn, kk;
    for (int i = 0; i < C.nd; ++i) {
        char c = C.idx[i];
        int pa = A.find(c), pb = B.find(c);
        CTMB_CHECK((pa >= 0) != (pb >= 0), "output label must come from exactly one operand");
        if (pa >= 0) { CTMB_CHECK(A.dim[pa] == C.dim[i], "extent mismatch (A,C)"); mm.push_back({c, C.dim[i], A.str[pa], 0, C.str[i]}); }
        else { CTMB_CHECK(B.dim[pb] == C.dim[i], "extent mismatch (B,C)"); nn.push_back({c, C.dim[i], 0, B.str[pb], C.str[i]}); }
    }
    for (int i = 0; i < A.nd; ++i) {
        char c = A.idx[i];
        if (C.find(c) >= 0) continue;
        int pb = B.find(c);
        CTMB_CHECK(pb >= 0, "label of A neither in B nor in the output");
        CTMB_CHECK(A.dim[i] == B.dim[pb], "extent mismatch (A,B)");
        kk.push_back({c, A.dim[i], A.str[i], B.str[pb], 0});
    }
    for (int i = 0; i < B.nd; ++i) {
        char c = B.idx[i];
        CTMB_CHECK(C.find(c) >= 0 || A.find(c) >= 0, "label of B neither in A nor in the output");
    }
    auto extent = [](const std::vector<Mode>& v) { int64_t n = 1; for (auto& m : v) n *= m.dim; return n; };
    int64_t M = extent(mm), N = extent(nn), K = extent(kk);
    CTMB_CHECK(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "GEMM extent overflow");

    auto table = [](const std::vector<Mode>& v, int64_t n, int which, std::vector<int>& out) {
        out.resize((size_t)n);
        std::vector<int64_t> cnt(v.size(), 0);
        int64_t off = 0;
        for (int64_t i = 0; i < n; ++i) {
            CTMB_CHECK(off >= 0 && off < (1ll << 31), "offset overflows int32");
            out[(size_t)i] = (int)off;
            for (int d = (int)v.size() - 1; d >= 0; --d) {
                int64_t s = which == 0 ? v[d].sa : which == 1 ? v[d].sb : v[d].sc;
                off += s;
                if (++cnt[d] < v[d].dim) break;
                off -= s * v[d].dim; cnt[d] = 0;
            }
        }
    };
    std::vector<int> t[6];
    table(mm, M, 0, t[0]); table(kk, K, 0, t[1]);
    table(kk, K, 1, t[2]); table(nn, N, 1, t[3]);
    table(mm, M, 2, t[4]); table(nn, N, 2, t[5]);

    Plan p; p.M = (int)M; p.N = (int)N; p.K = (int)K;
    auto last_stride = [](const std::vector<Mode>& v, int which) -> int64_t {
        // stride of the fastest mode with extent > 1 (huge if the group is trivial)
        for (int d = (int)v.size() - 1; d >= 0; --d)
            if (v[d].dim > 1) return which == 0 ? v[d].sa : v[d].sb;
        return (int64_t)1 << 62;
    };
    if (last_stride(kk, 0) <= last_stride(mm, 0)) p.flags |= TC_A_KFAST;
    if (last_stride(kk, 1) <= last_stride(nn, 1)) p.flags |= TC_B_KFAST;
    if (conjA) p.flags |= TC_CONJ_A;
    if (conjB) p.flags |= TC_CONJ_B;
    size_t total = 0;
    for (auto& v : t) total += v.size();
    std::vector<int> host; host.reserve(total);
    size_t offs[6];
    for (int i = 0; i < 6; ++i) { offs[i] = host.size(); host.insert(host.end(), t[i].begin(), t[i].end()); }
    CTMB_CUDA(cudaMalloc(&p.dev, total * sizeof(int)));
    // synchronous copy: plans are built once per shape, outside any timed/captured region
    CTMB_CUDA(cudaMemcpy(p.dev, host.data(), total * sizeof(int), cudaMemcpyHostToDevice));
    p.tab.a_m = p.dev + offs[0]; p.tab.a_k = p.dev + offs[1];
    p.tab.b_k = p.dev + offs[2]; p.tab.b_n = p.dev + offs[3];
    p.tab.c_m = p.dev + offs[4]; p.tab.c_n = p.dev + offs[5];
    // plain strided operands (every table an arithmetic progression): eligible for the TMA-fed kernel
    auto linear = [](const std::vector<int>& v, int& stride) {
        stride = v.size() > 1 ? v[1] : 0;
        for (size_t i = 0; i < v.size(); ++i) if ((long long)v[i] != (long long)i * stride) return false;
        return true;
    };
    p.tab.lin = (linear(t[0], p.tab.a_sm) && linear(t[1], p.tab.a_sk) && linear(t[2], p.tab.b_sk) && linear(t[3], p.tab.b_sn)) ? 1 : 0;
    p.tab.c_lin = (linear(t[4], p.tab.c_sm) && linear(t[5], p.tab.c_sn)) ? 1 : 0;
    auto res = plans_.emplace(key, p);
    return res.first->second;
}

cudaEvent_t Engine::get_event() {
    if (!ev_pool_.empty()) { cudaEvent_t e = ev_pool_.back(); ev_pool_.pop_back(); return e; }
    cudaEvent_t e; CTMB_CUDA(cudaEventCreate(&e)); return e;
}
void Engine::prof_begin(int) {
    prof_open_ = get_event();
    CTMB_CUDA(cudaEventRecord(prof_open_, stream));
}
void Engine::prof_end(int cat, double fl, double by) {
    cudaEvent_t b = get_event();
    CTMB_CUDA(cudaEventRecord(b, stream));
    prof_.push_back({cat, prof_open_, b, fl, by});
    prof_open_ = nullptr;
}
void Engine::prof_collect(ProfTotals out[CAT_COUNT]) {
    for (int i = 0; i < CAT_COUNT; ++i) out[i] = ProfTotals{};
    for (auto& r : prof_) {
        CTMB_CUDA(cudaEventSynchronize(r.b));
        float ms = 0; CTMB_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        out[r.cat].ms += ms; out[r.cat].flops += r.flops; out[r.cat].bytes += r.bytes; out[r.cat].launches += 1;
    }
}
void Engine::prof_reset() {
    for (auto& r : prof_) { ev_pool_.push_back(r.a); ev_pool_.push_back(r.b); }
    prof_.clear();
}

void Engine::flush() {
    if (!pend_active_) return;
    pend_active_ = false;
    {
        const double es = cplx ? 16.0 : 8.0;
        const double fl = 2.0 * pend_.M * (double)pend_.N * pend_.K * (cplx ? 4.0 : 1.0) * pend_.nbatch;
        const double by = es * pend_.nbatch * ((double)pend_.M * pend_.K + (double)pend_.K * pend_.N + (double)pend_.M * pend_.N);
        // Split-K for long reductions that would otherwise run on a handful of CTAs (the V^H A products of the
        // blocked QR at chi*D^2 = 16384: 8 x 504 outputs, K = 16384): slices write dense partials into an
        // engine-owned buffer, a second kernel sums them in a fixed order (deterministic, no atomics).
        pend_.ksplit = 1; pend_.kchunk = 0; pend_.partial = nullptr;
        int bm = 32, bn = 32;
        tc_tile_shape(pend_, cplx, bm, bn);
        const long long tiles = (long long)((pend_.M + bm - 1) / bm) * ((pend_.N + bn - 1) / bn) * pend_.nbatch;
        static int sk_mode = -1;
        if (sk_mode < 0) { const char* ev = getenv("CTMB_SPLITK"); sk_mode = ev ? atoi(ev) : 1; }
        // two regimes: long reductions on at most one wave of tiles (K >= 1024: 296 CTAs, chunks of >= 256), and the small
        // latency-bound products of configs 2 / 4 (128 <= K < 1024, a few hundred 32 x 32 tiles: the serial k loop of a CTA
        // is the latency of the launch; 1184 CTAs, chunks of >= 64: +2 % moves/s at config 2, +5 % at config 4)
        const bool long_k = pend_.K >= 1024 && tiles <= 148;
        const bool small_k = pend_.K >= 128 && pend_.K < 1024 && tiles <= 592;
        if (sk_mode && (long_k || small_k)) {
            int S = (int)std::min<long long>(std::min<long long>((long_k ? 296 : 1184) / tiles, pend_.K / (long_k ? 256 : 64)), 32);
            if (S >= 2) {
                int kchunk = ((pend_.K + S - 1) / S + 31) & ~31;
                S = (pend_.K + kchunk - 1) / kchunk;
                pend_.ksplit = S; pend_.kchunk = kchunk;
                pend_.partial = persistent("splitk", (size_t)pend_.nbatch * S * pend_.M * pend_.N * (cplx ? 16 : 8));
                tc_tile_shape(pend_, cplx, bm, bn);
            }
        }
        ProfScope ps(*this, CAT_GEMM, fl, by);
        tc_launch(pend_, cplx, stream);
        if (pend_.ksplit > 1) { ++launches; tc_splitk_reduce_launch(pend_, cplx, stream); }
    }
    pend_plans_.clear();
}

void Engine::contract(const Tn& A, bool conjA, const Tn& B, bool conjB, const Tn& C,
                      unsigned long long* amax, double alpha, bool accumulate) {
    if (ws.dry()) { validate_contraction(A, B, C); return; }
    const Plan& pl = get_plan(A, conjA && cplx, B, conjB && cplx, C);
    flops += 2.0 * pl.M * (double)pl.N * pl.K * (cplx ? 4.0 : 1.0);
    if (pend_active_) {
        bool ok = pend_.M == pl.M && pend_.N == pl.N && pend_.K == pl.K &&
                  pend_.alpha == alpha && pend_.nbatch < TC_MAX_BATCH;
        if (ok) {
            bool have = std::find(pend_plans_.begin(), pend_plans_.end(), &pl) != pend_plans_.end();
            if (!have && (int)pend_plans_.size() >= TC_MAX_TABS) ok = false;
        }
        if (!ok) flush();
    }
    if (!pend_active_) {
        pend_ = TcParams{};
        pend_.M = pl.M; pend_.N = pl.N; pend_.K = pl.K; pend_.alpha = alpha;
        pend_.nbatch = 0;
        pend_plans_.clear();
        pend_active_ = true;
    }
    int ti = -1;
    for (size_t i = 0; i < pend_plans_.size(); ++i) if (pend_plans_[i] == &pl) ti = (int)i;
    if (ti < 0) { ti = (int)pend_plans_.size(); pend_plans_.push_back(&pl); pend_.tab[ti] = pl.tab; }
    TcBatchEntry& e = pend_.batch[pend_.nbatch++];
    e.A = A.ptr; e.B = B.ptr; e.C = C.ptr; e.amax = amax; e.tab = ti; e.flags = pl.flags | (accumulate ? TC_ACCUM : 0);
}

void Engine::chain_multi(std::vector<ChainJob>& jobs, size_t temp_budget) {
    if (jobs.empty()) return;
    const size_t nops = jobs[0].ops.size();
    CTMB_CHECK(nops >= 2, "chain needs at least two operands");
    for (auto& j : jobs) CTMB_CHECK(j.ops.size() == nops && j.conj.size() == nops, "ragged chain jobs");
    flush();

    // per-step output label order of one job
    auto step_out = [&](const ChainJob& j, const Tn& cur, size_t i, std::string& idx, std::vector<int64_t>& dims) {
        std::string later;
        for (size_t q = i + 1; q < nops; ++q) later += j.ops[q].idx;
        later += j.out.idx;
        idx.clear(); dims.clear();
        auto add = [&](const Tn& t) {
            for (int d = 0; d < t.nd; ++d) {
                char c = t.idx[d];
                if (later.find(c) == std::string::npos) continue;
                if (idx.find(c) != std::string::npos) continue;
                idx.push_back(c); dims.push_back(t.dim[d]);
            }
        };
        add(cur); add(j.ops[i]);
    };
    // memory need of one job: max over steps of (input temp + output temp)
    size_t need = 0;
    {
        const ChainJob& j = jobs[0];
        Tn cur = j.ops[0];
        size_t prev = 0;
        for (size_t i = 1; i + 1 < nops; ++i) {
            std::string idx; std::vector<int64_t> dims;
            step_out(j, cur, i, idx, dims);
            size_t n = esize(); for (auto d : dims) n *= (size_t)d;
            n = (n + 255) & ~(size_t)255;
            need = std::max(need, prev + n);
            prev = n;
            cur = make_tn(nullptr, idx, dims);
        }
    }
    size_t chunk = jobs.size();
    if (need > 0) chunk = std::max<size_t>(1, std::min(jobs.size(), temp_budget / need));

    for (size_t j0 = 0; j0 < jobs.size(); j0 += chunk) {
        const size_t j1 = std::min(jobs.size(), j0 + chunk);
        const size_t mlo = ws.mark(false), mhi = ws.mark(true);
        std::vector<Tn> cur(j1 - j0);
        for (size_t j = j0; j < j1; ++j) cur[j - j0] = jobs[j].ops[0];
        for (size_t i = 1; i < nops; ++i) {
            const bool back = (i & 1) != 0;
            // the temps written two steps ago (same end of the stack) are dead now
            ws.release(back ? mhi : mlo, back);
            for (size_t j = j0; j < j1; ++j) {
                ChainJob& job = jobs[j];
                Tn out;
                if (i + 1 == nops) out = job.out;
                else {
                    std::string idx; std::vector<int64_t> dims;
                    step_out(job, cur[j - j0], i, idx, dims);
                    out = temp(idx, dims, back);
                }
                contract(cur[j - j0], i == 1 ? (bool)job.conj[0] : false, job.ops[i], job.conj[i], out,
                         i + 1 == nops ? job.amax : nullptr);
                cur[j - j0] = out;
            }
            flush();
        }
        ws.release(mlo, false); ws.release(mhi, true);
    }
}

}  // namespace ctmb
