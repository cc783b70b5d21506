// Host-side contraction planner: labelled tensor views -> offset tables -> batched launches.
#pragma once
#include "common.h"

namespace ctmb {

constexpr int MAX_ND = 12;

// A labelled, strided view of device memory. Strides in elements of the scalar type.
struct Tn {
    void* ptr = nullptr;
    int nd = 0;
    char idx[MAX_ND + 1] = {0};
    int64_t dim[MAX_ND] = {0};
    int64_t str[MAX_ND] = {0};

    int64_t numel() const { int64_t n = 1; for (int i = 0; i < nd; ++i) n *= dim[i]; return n; }
    int find(char c) const { for (int i = 0; i < nd; ++i) if (idx[i] == c) return i; return -1; }
};

// contiguous row-major view with the given labels
Tn make_tn(void* ptr, const char* idx, std::initializer_list<int64_t> dims);
Tn make_tn(void* ptr, const std::string& idx, const std::vector<int64_t>& dims);
// view with explicit element strides
Tn make_strided(void* ptr, const char* idx, std::initializer_list<int64_t> dims, std::initializer_list<int64_t> strides);
// split mode `c` (extent d1*d2) into two modes c1 (slow, extent d1) and c2 (fast, extent d2)
Tn split_mode(const Tn& t, char c, char c1, char c2, int64_t d1, int64_t d2);
// relabel / permute (no data movement)
Tn relabel(const Tn& t, const char* idx);
Tn transpose_view(const Tn& t, const char* new_order);

struct Plan {
    int M = 0, N = 0, K = 0;
    int flags = 0;
    int* dev = nullptr;        // one allocation holding the six tables
    TcTables tab{};
};

// Two-ended stack over one caller-provided device buffer. With base == nullptr it only
// measures (dry run used by the *_workspace_bytes queries; no kernel is launched then).
class Workspace {
public:
    void reset(void* base, size_t bytes) {
        // keep both ends 256-byte aligned (the back end grows down from base+cap)
        size_t mis = base ? (size_t)(256 - ((uintptr_t)base & 255)) & 255 : 0;
        if (mis > bytes) mis = bytes;
        base_ = base ? (char*)base + mis : nullptr; cap_ = (bytes - mis) & ~(size_t)255; lo_ = hi_ = 0; peak_ = 0;
    }
    void* alloc(size_t bytes, bool back = false);
    size_t mark(bool back = false) const { return back ? hi_ : lo_; }
    void release(size_t mark, bool back = false) { (back ? hi_ : lo_) = mark; }
    size_t peak() const { return peak_; }
    bool dry() const { return base_ == nullptr; }
private:
    char* base_ = nullptr; size_t cap_ = 0, lo_ = 0, hi_ = 0, peak_ = 0;
};

class Engine {
public:
    explicit Engine(int device);
    ~Engine();
    int device() const { return device_; }
    bool cplx = false;             // element type of the current call
    cudaStream_t stream = nullptr;
    Workspace ws;
    long long launches = 0;        // kernels launched (our own), for bench accounting
    struct IterHint { int q = 0; int lo = -1; int age = 0; double range = 0.0; bool known = false; double tail = 0.0; int cheb_pause = 0; int streak = 0; bool cheb_seen = false; };   // range: S_0 / S_chi seen last time (0: unknown)   // q: start count; lo: largest count that failed recently
    std::map<std::string, IterHint> iter_hint;   // adaptive range finder, per problem shape: power iterations that satisfied
                                                 // the residual test last time, and how long not to probe below them
    double flops = 0;              // algorithmic real flops enqueued
    // residual-checked range finder: how many checks ran, how many results were returned although they missed the bound
    // (rounding floor / rsvd_max_rounds), and the worst residual / bound ratio among those (ctmb_get_rsvd_status)
    struct RsvdStatus { long long checks = 0, missed = 0, calls = 0, iterations = 0; double worst_ratio = 0.0; };
    RsvdStatus rsvd_status;
    unsigned long long* pinned_words();     // 32 pinned host words owned by this engine (read-back of device scalars)
    void drop_pending() { pend_active_ = false; pend_plans_.clear(); }

    // optional per-kernel-class timing with CUDA events on the launching stream
    enum Cat { CAT_GEMM = 0, CAT_QR = 1, CAT_JACOBI = 2, CAT_MISC = 3, CAT_COUNT = 4 };
    bool profiling = false;
    void prof_begin(int cat);
    void prof_end(int cat, double flops, double bytes);
    struct ProfTotals { double ms = 0, flops = 0, bytes = 0; long long launches = 0; };
    void prof_collect(ProfTotals out[CAT_COUNT]);     // synchronises the recorded events
    void prof_reset();

    size_t esize() const { return cplx ? 16 : 8; }

    // Intra-site split of the range finder over a GROUP of GPUs ("G = 2N"): every member of the group runs the
    // same call on the same inputs; the n x n x k operator applications are split by sketch columns and the slabs are
    // exchanged by an in-place all-gather that the host provides (NCCL over NVLink through torch.distributed).
    // buf = [coll_n][bytes_per_rank], this rank's slot already filled; must be enqueued on / ordered with `stream`.
    typedef int (*allgather_fn)(void* ctx, void* buf, size_t bytes_per_rank, void* stream);
    int coll_rank = 0, coll_n = 1;
    allgather_fn coll_fn = nullptr;
    void* coll_ctx = nullptr;
    bool coll_active() const { return coll_n > 1 && coll_fn != nullptr; }
    void allgather(void* buf, size_t bytes_per_rank) {
        flush();
        CTMB_CHECK(coll_fn(coll_ctx, buf, bytes_per_rank, (void*)stream) == 0, "the host's all-gather callback failed");
    }

    // C = A * B over shared labels not in C; C's labels/strides define the output layout.
    // Enqueued into the pending batch; flushed when incompatible or on flush().
    void contract(const Tn& A, bool conjA, const Tn& B, bool conjB, const Tn& C,
                  unsigned long long* amax = nullptr, double alpha = 1.0, bool accumulate = false);
    // allocate a contiguous workspace tensor with the given labels
    Tn temp(const std::string& idx, const std::vector<int64_t>& dims, bool back = false);
    void flush();

    // Evaluate, for every job, the product of its operands left to right (pairwise; an
    // index survives a step iff a later operand or the output carries it) into job.out.
    // All jobs must have the same number of operands; step i of all jobs shares launches.
    struct ChainJob {
        std::vector<Tn> ops; std::vector<bool> conj; Tn out; unsigned long long* amax = nullptr;
    };
    void chain_multi(std::vector<ChainJob>& jobs, size_t temp_budget_bytes = (size_t)12 << 30);

    // persistent device scratch owned by the engine (random sketch matrices, amax slots ...)
    void* persistent(const std::string& key, size_t bytes, bool* created = nullptr);
    // frees every persistent buffer whose key starts with `prefix` except `keep` (a warm-start slot whose shape changed,
    // e.g. after env.extend(): the old basis is of no use and must not pile up)
    void drop_persistent(const std::string& prefix, const std::string& keep);

private:
    const Plan& get_plan(const Tn& A, bool conjA, const Tn& B, bool conjB, const Tn& C);
    int device_;
    std::map<std::string, Plan> plans_;
    std::map<std::string, std::pair<void*, size_t>> persist_;
    // pending batch
    TcParams pend_{};
    std::vector<const Plan*> pend_plans_;
    bool pend_active_ = false;
    struct ProfRec { int cat; cudaEvent_t a, b; double flops, bytes; };
    std::vector<ProfRec> prof_;
    std::vector<cudaEvent_t> ev_pool_;
    cudaEvent_t prof_open_ = nullptr;
    cudaEvent_t get_event();
    unsigned long long* pinned_ = nullptr;
};

// RAII helper: times one launch of class `cat` when profiling is on
struct ProfScope {
    Engine& e; int cat; double flops, bytes;
    ProfScope(Engine& eng, int c, double f = 0, double b = 0) : e(eng), cat(c), flops(f), bytes(b) { ++e.launches; if (e.profiling) e.prof_begin(c); }
    ~ProfScope() { if (e.profiling) e.prof_end(cat, flops, bytes); }
};

}  // namespace ctmb
