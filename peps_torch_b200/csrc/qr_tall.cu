// Tall-panel Householder factorisation (WY form) for the blocked QR of the range-finder sketches.
//
// The sketches of config 5 are 16384 x 512 (D = 8, chi = 256): 11 blocked QRs per move, each made of leaf panels.  Round 1
// factored a leaf inside the shared memory of one 8-CTA cluster, which holds only 8 columns at 16384 rows and parallelises
// over COLUMNS (one warp per column walks all 2048 local rows): 180 us per 8-column leaf, 64 leaves and 64 trailing-update
// GEMM chains per QR (profiles/r2_c5s_launches.md).  Here the rows are split in contiguous chunks over up to 148
// co-resident CTAs (cooperative launch, one thread per row, the chunk resident in shared memory), so a leaf is 32 columns
// wide and a column step costs one per-matrix barrier in global memory:
//
//   every CTA publishes  E_c = sum_{r>j} conj(x_r) a_{r,c}  (c >= j; c = j gives the tail norm) for its rows, CTA 0 also
//   row j; after the barrier every CTA sums the partials IN A FIXED ORDER (deterministic: the members of a multi-GPU group
//   factor the same sketch redundantly and must agree) and derives beta, tau and v^H a_c locally -- norm and reflector
//   application share one reduction, exactly as in qr_cluster_kernel (qr.cu), whose conventions are kept:
//   LAPACK geqr2 reflectors H = I - tau v v^H with v_j = 1; on exit A <- V (explicit unit diagonal, zeros above),
//   Rpp <- the b x b triangular factor, Tau <- tau.
//
// Replaces (with the GEMM-based block updates around it) the orthogonalisation inside the reference's randomised
// decompositions (linalg/svd_rsvd.py:6-117 torch.linalg.qr; torch.svd_lowrank at ctm_projectors.py:249-252).
#include "common.h"
#include "cx.h"
#include <cstdlib>

namespace ctmb {

constexpr int QT_THREADS = 256;
constexpr int QT_B = 32;                 // widest panel
constexpr int QT_NW = QT_THREADS / 32;

// folded butterfly: on entry every lane holds 32 partial values; on exit lane l holds the warp total of value l
// (31 shuffles instead of 160)
template <bool CPLX>
__device__ __forceinline__ typename Sc<CPLX>::T fold32(typename Sc<CPLX>::T (&acc)[QT_B], int lane) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    T v16[16], v8[8], v4[4], v2[2];
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0, h2 = (lane & 2) != 0, h1 = (lane & 1) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const T send = h16 ? acc[i] : acc[i + 16], keep = h16 ? acc[i + 16] : acc[i];
        v16[i] = S::add(keep, S::shfl_xor(send, 16));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const T send = h8 ? v16[i] : v16[i + 8], keep = h8 ? v16[i + 8] : v16[i];
        v8[i] = S::add(keep, S::shfl_xor(send, 8));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T send = h4 ? v8[i] : v8[i + 4], keep = h4 ? v8[i + 4] : v8[i];
        v4[i] = S::add(keep, S::shfl_xor(send, 4));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const T send = h2 ? v4[i] : v4[i + 2], keep = h2 ? v4[i + 2] : v4[i];
        v2[i] = S::add(keep, S::shfl_xor(send, 2));
    }
    const T send = h1 ? v2[0] : v2[1], keep = h1 ? v2[1] : v2[0];
    return S::add(keep, S::shfl_xor(send, 1));
}

__device__ __forceinline__ double qt_ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double2 qt_ldcg(const double2* p) { return __ldcg(p); }

// gpart: [nb][2][cpm + 1][QT_B] partial dots (slot cpm: row j of the panel); gbar: [nb] barrier counters (zeroed before launch)
template <bool CPLX>
__global__ void __launch_bounds__(QT_THREADS) qr_tall_panel_kernel(PtrBatch Ab, PtrBatch Rppb, PtrBatch Taub, int rows, int b,
                                                                   int ld, int cpm, int rpc, void* gpart_v, unsigned int* gbar) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const int mat = blockIdx.x / cpm, cta = blockIdx.x % cpm;
    T* __restrict__ A = reinterpret_cast<T*>(Ab.p[mat]);
    T* gpart = reinterpret_cast<T*>(gpart_v) + (size_t)mat * 2 * (cpm + 1) * QT_B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = cta * rpc;
    const int nloc = max(0, min(rpc, rows - r0));

    extern __shared__ __align__(16) unsigned char qt_smem[];
    T* slab = reinterpret_cast<T*>(qt_smem);                 // [b][rpc]   (column-major chunk)
    T* red = slab + (size_t)b * rpc;                         // [QT_NW][QT_B]
    T* fsm = red + QT_NW * QT_B;                             // [QT_B]     f_c of the current column
    T* taus = fsm + QT_B;                                    // [QT_B]
    T* scal = taus + QT_B;                                   // [2]: s, beta

    for (int c = 0; c < b; ++c)
        for (int lr = tid; lr < nloc; lr += QT_THREADS) slab[(size_t)c * rpc + lr] = A[(size_t)c * ld + r0 + lr];
    __syncthreads();

    volatile unsigned int* bar = gbar + mat;
    unsigned int epoch = 0;
    const int kmax = min(rows, b);
    const int cps = (cpm + QT_NW - 1) / QT_NW;               // CTAs per reduction slice
    for (int j = 0; j < kmax; ++j) {
        const int par = j & 1;
        T* gp = gpart + (size_t)par * (cpm + 1) * QT_B;
        // ---- partial dots of column j (rows below the diagonal) with the columns j .. b-1 ----
        T acc[QT_B];
#pragma unroll
        for (int c = 0; c < QT_B; ++c) acc[c] = S::zero();
        for (int lr = tid; lr < nloc; lr += QT_THREADS) {
            if (r0 + lr > j) {
                const T x = S::conj(slab[(size_t)j * rpc + lr]);
#pragma unroll
                for (int c = 0; c < QT_B; ++c)
                    if (c >= j && c < b) acc[c] = S::fma(x, slab[(size_t)c * rpc + lr], acc[c]);
            }
        }
        red[warp * QT_B + lane] = fold32<CPLX>(acc, lane);
        __syncthreads();
        if (warp == 0) {
            T s = red[lane];
#pragma unroll
            for (int w = 1; w < QT_NW; ++w) s = S::add(s, red[w * QT_B + lane]);
            gp[(size_t)cta * QT_B + lane] = s;
        } else if (warp == 1 && cta == 0) {                  // rows 0 .. rpc-1 >= b-1 live in CTA 0
            gp[(size_t)cpm * QT_B + lane] = lane < b ? slab[(size_t)lane * rpc + j] : S::zero();
        }
        // ---- per-matrix barrier ----
        __syncthreads();
        ++epoch;
        if (tid == 0) {
            __threadfence();
            atomicAdd(const_cast<unsigned int*>(bar), 1u);
            const unsigned int target = epoch * (unsigned int)cpm;
            while (*bar < target) { }
            __threadfence();
        }
        __syncthreads();
        // ---- totals in a fixed order: slice `warp` of the CTAs, value `lane` ----
        {
            // all loads of the slice in flight together (148 CTAs / 8 warps <= 19 partials per lane), then a fixed-order sum:
            // one L2 round trip instead of 19 dependent ones (first version: 5.7 us per column step, ncu smsp__cycles_active)
            constexpr int QT_MAXCPS = 20;
            T v[QT_MAXCPS];
            const int q0 = warp * cps;
#pragma unroll
            for (int i = 0; i < QT_MAXCPS; ++i) {
                const int q = q0 + i;
                v[i] = (i < cps && q < cpm) ? qt_ldcg(gp + (size_t)q * QT_B + lane) : S::zero();
            }
            T s = S::zero();
#pragma unroll
            for (int i = 0; i < QT_MAXCPS; ++i) s = S::add(s, v[i]);
            red[warp * QT_B + lane] = s;
        }
        __syncthreads();
        if (warp == 0) {
            T tot = red[lane];
#pragma unroll
            for (int w = 1; w < QT_NW; ++w) tot = S::add(tot, red[w * QT_B + lane]);
            const T rowv = qt_ldcg(gp + (size_t)cpm * QT_B + lane);
            const double tail = __shfl_sync(0xffffffffu, S::re(tot), j);
            const T alpha = S::make(__shfl_sync(0xffffffffu, S::re(rowv), j), __shfl_sync(0xffffffffu, S::im(rowv), j));
            T tj = S::zero(), sc = S::zero(); double beta;
            if (tail == 0.0 && S::im(alpha) == 0.0) { beta = S::re(alpha); }
            else {
                // ||x||^2 = |alpha|^2 + tail;  beta = -sign(Re alpha)||x||, tau = (beta-alpha)/beta, s = 1/(alpha-beta)
                const double nsq = S::abs2(alpha) + tail;
                const double rn = rsqrt(nsq);
                const double sg = copysign(1.0, S::re(alpha));
                beta = -sg * (nsq * rn);
                const double ib = -sg * rn;
                tj = S::make((beta - S::re(alpha)) * ib, -S::im(alpha) * ib);
                const T amb = S::sub(alpha, S::make(beta, 0.0));
                sc = S::scale(S::conj(amb), 1.0 / S::abs2(amb));
            }
            // v^H a_c = a_{j,c} + conj(s) E_c ;  f_c = conj(tau) v^H a_c
            fsm[lane] = (lane > j && lane < b) ? S::mul(S::conj(tj), S::add(rowv, S::mul(S::conj(sc), tot))) : S::zero();
            if (lane == 0) { taus[j] = tj; scal[0] = sc; scal[1] = S::make(beta, 0.0); }
        }
        __syncthreads();
        // ---- rank-1 update of the own rows; column j <- v, diagonal <- beta ----
        {
            const T sc = scal[0];
            for (int lr = tid; lr < nloc; lr += QT_THREADS) {
                const int gr = r0 + lr;
                if (gr > j) {
                    const T x = slab[(size_t)j * rpc + lr];
                    const T v = S::mul(x, sc);
#pragma unroll 4
                    for (int c = j + 1; c < b; ++c) slab[(size_t)c * rpc + lr] = S::sub(slab[(size_t)c * rpc + lr], S::mul(fsm[c], v));
                    slab[(size_t)j * rpc + lr] = v;
                } else if (gr == j) {
                    for (int c = j + 1; c < b; ++c) slab[(size_t)c * rpc + lr] = S::sub(slab[(size_t)c * rpc + lr], fsm[c]);
                    slab[(size_t)j * rpc + lr] = scal[1];
                }
            }
        }
        // (no barrier needed here: a row is read and written by one thread only; red / fsm are protected by the barriers above)
    }
    __syncthreads();
    // ---- R block, tau, and V (explicit unit diagonal, zeros above) ----
    if (cta == 0) {
        T* Rpp = reinterpret_cast<T*>(Rppb.p[mat]);
        T* tau_out = reinterpret_cast<T*>(Taub.p[mat]);
        if (Rpp != nullptr)
            for (int e = tid; e < b * b; e += QT_THREADS) {
                const int c = e / b, r = e % b;
                Rpp[e] = (r <= c && r < nloc) ? slab[(size_t)c * rpc + r] : S::zero();
            }
        for (int c = tid; c < b; c += QT_THREADS) tau_out[c] = c < kmax ? taus[c] : S::zero();
    }
    for (int c = 0; c < b; ++c)
        for (int lr = tid; lr < nloc; lr += QT_THREADS) {
            const int gr = r0 + lr;
            A[(size_t)c * ld + gr] = gr > c ? slab[(size_t)c * rpc + lr] : (gr == c ? S::one() : S::zero());
        }
}

// ---------------------------------------------------------------------------------------------
// CTMB_QR_TALL: 0 = off, 1 = where the leaf gets wider than the cluster kernel's (default), 2 = wherever it applies
static int qt_mode() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("CTMB_QR_TALL"); mode = e ? atoi(e) : 1; }
    return mode;
}
int qr_tall_mode() { return qt_mode(); }
static int qt_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;               // planning-only handle (no device): the B200 figure
    }
    return sms;
}
// layout of one launch: CTAs per matrix and rows per CTA (0: the tall kernel does not apply)
static bool qt_shape(int nb, int rows, int b, bool cplx, int& cpm, int& rpc, size_t& smem) {
    if (!qt_mode() || b > QT_B || b < 1 || rows < 1024 || nb < 1) return false;
    const size_t es = cplx ? 16 : 8;
    cpm = std::min(std::min(qt_sms() / nb, 160), rows / std::max(b, 32));   // rpc >= b: the diagonal rows of the panel live in CTA 0; <= 8 x 20 partials
    if (cpm < 2) return false;
    rpc = (rows + cpm - 1) / cpm;
    cpm = (rows + rpc - 1) / rpc;
    smem = ((size_t)b * rpc + (QT_NW + 2) * QT_B + 2) * es;
    return smem <= 200 * 1024;
}

int qr_tall_panel_width(int nb, int rows, int cols, bool cplx) {
    int cpm, rpc; size_t smem;
    const int b = std::min(QT_B, cols);
    return qt_shape(nb, rows, b, cplx, cpm, rpc, smem) ? b : 0;
}

size_t qr_tall_scratch_bytes(int nb) {
    return (size_t)nb * 2 * (qt_sms() + 1) * QT_B * 16 + 256 + (size_t)TC_MAX_BATCH * sizeof(unsigned int);
}

bool qr_tall_panel_launch(const PtrBatch& A, const PtrBatch& Rpp, const PtrBatch& Tau, int nb, int rows, int b, int ld, bool cplx,
                          void* scratch, cudaStream_t stream) {
    int cpm, rpc; size_t smem;
    if (scratch == nullptr || !qt_shape(nb, rows, b, cplx, cpm, rpc, smem)) return false;
    unsigned int* gbar = reinterpret_cast<unsigned int*>(scratch);
    void* gpart = reinterpret_cast<char*>(scratch) + 256;
    CTMB_CHECK(nb <= 64, "bad batch");
    CTMB_CUDA(cudaMemsetAsync(gbar, 0, 256, stream));
    const void* kern = cplx ? (const void*)qr_tall_panel_kernel<true> : (const void*)qr_tall_panel_kernel<false>;
    static size_t set[2] = {0, 0};
    if (smem > set[cplx ? 1 : 0]) {
        CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
        set[cplx ? 1 : 0] = 200 * 1024 + 1024;
    }
    PtrBatch a = A, r = Rpp, t = Tau;
    void* args[] = {&a, &r, &t, &rows, &b, &ld, &cpm, &rpc, &gpart, &gbar};
    CTMB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(nb * cpm), dim3(QT_THREADS), args, smem, stream));
    return true;
}

}  // namespace ctmb
