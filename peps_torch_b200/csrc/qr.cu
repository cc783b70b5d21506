// Batched Householder QR (explicit thin Q) of tall-skinny column-major matrices.
//
// Used by the randomised range finder that replaces the full LAPACK SVD / EVD of the
// reference (linalg/svd_gesdd.py:91, linalg/eig_sym.py:25; its own randomised variant:
// linalg/svd_rsvd.py:6-117).  Householder reflections are unconditionally backward stable,
// which matters here: the sketches Y = M*Q have singular values spanning >16 decades
// (rank-deficient M from zero-padded / product-like environments), where Gram-matrix based
// orthogonalisation (CholeskyQR) loses the small directions.
//
// One CTA per matrix; the matrix stays in L2/L1 (432x96 doubles = 332 KB at D=3, chi=48).
// Phase A: geqr2 (reflectors stored below the diagonal), Phase B: org2r (Q formed in place).
#include "common.h"
#include "cx.h"
#include <cooperative_groups.h>
#include <cstdlib>

namespace ctmb {

constexpr int QR_THREADS = 1024;

template <bool CPLX>
__global__ void __launch_bounds__(QR_THREADS) qr_kernel(PtrBatch Ab, PtrBatch Rb, int rows, int cols, int ld) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    T* __restrict__ A = reinterpret_cast<T*>(Ab.p[blockIdx.x]);
    T* __restrict__ Rout = reinterpret_cast<T*>(Rb.p[blockIdx.x]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = QR_THREADS / 32;
    extern __shared__ __align__(16) unsigned char qr_smem[];
    T* tau = reinterpret_cast<T*>(qr_smem);            // [cols]
    __shared__ double red[NW];
    __shared__ T sh_scale;
    __shared__ T sh_tau;

    const int kmax = min(rows, cols);
    // ---------------- Phase A: factorisation ----------------
    for (int j = 0; j < kmax; ++j) {
        T* cj = A + (size_t)j * ld;
        // ||x[1:]||^2
        double s = 0.0;
        for (int r = j + 1 + tid; r < rows; r += QR_THREADS) s += S::abs2(cj[r]);
        s = warp_sum(s);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (warp == 0) {
            double t = lane < NW ? red[lane] : 0.0;
            t = warp_sum(t);
            if (lane == 0) {
                T alpha = cj[j];
                T tj, sc;
                if (t == 0.0 && S::im(alpha) == 0.0) {
                    tj = S::zero(); sc = S::zero();          // H = I
                } else {
                    double beta = -copysign(sqrt(S::abs2(alpha) + t), S::re(alpha));
                    tj = S::make((beta - S::re(alpha)) / beta, -S::im(alpha) / beta);
                    sc = S::div(S::one(), S::sub(alpha, S::make(beta, 0.0)));
                    cj[j] = S::make(beta, 0.0);
                }
                tau[j] = tj; sh_tau = tj; sh_scale = sc;
            }
        }
        __syncthreads();
        const T sc = sh_scale;
        const T tj = sh_tau;
        for (int r = j + 1 + tid; r < rows; r += QR_THREADS) cj[r] = S::mul(cj[r], sc);
        __syncthreads();
        if (S::abs2(tj) != 0.0) {
            const T ctau = S::conj(tj);
            for (int c = j + 1 + warp; c < cols; c += NW) {
                T* cc = A + (size_t)c * ld;
                T dot = (lane == 0) ? cc[j] : S::zero();
                for (int r = j + 1 + lane; r < rows; r += 32) dot = S::fma(S::conj(cj[r]), cc[r], dot);
                dot = warp_sum_t<CPLX>(dot);
                const T f = S::mul(ctau, dot);
                if (lane == 0) cc[j] = S::sub(cc[j], f);
                for (int r = j + 1 + lane; r < rows; r += 32) cc[r] = S::sub(cc[r], S::mul(f, cj[r]));
            }
        }
        __syncthreads();
    }
    // ---------------- save R ----------------
    if (Rout != nullptr) {
        for (int e = tid; e < cols * cols; e += QR_THREADS) {
            int c = e / cols, r = e % cols;
            Rout[e] = (r <= c && r < rows) ? A[(size_t)c * ld + r] : S::zero();
        }
        __syncthreads();
    }
    // ---------------- Phase B: form Q in place ----------------
    for (int j = cols - 1; j >= kmax; --j) {   // cols > rows: not used by the engine, keep defined
        T* cj = A + (size_t)j * ld;
        for (int r = tid; r < rows; r += QR_THREADS) cj[r] = S::zero();
    }
    __syncthreads();
    for (int j = kmax - 1; j >= 0; --j) {
        T* cj = A + (size_t)j * ld;
        const T tj = tau[j];
        if (S::abs2(tj) != 0.0) {
            for (int c = j + 1 + warp; c < cols; c += NW) {
                T* cc = A + (size_t)c * ld;
                // row j of the trailing columns is zero at this point
                T dot = S::zero();
                for (int r = j + 1 + lane; r < rows; r += 32) dot = S::fma(S::conj(cj[r]), cc[r], dot);
                dot = warp_sum_t<CPLX>(dot);
                const T f = S::mul(tj, dot);
                if (lane == 0) cc[j] = S::sub(S::zero(), f);
                for (int r = j + 1 + lane; r < rows; r += 32) cc[r] = S::sub(cc[r], S::mul(f, cj[r]));
            }
        }
        __syncthreads();
        const T mt = S::sub(S::zero(), tj);
        for (int r = tid; r < rows; r += QR_THREADS) {
            T v;
            if (r < j) v = S::zero();
            else if (r == j) v = S::sub(S::one(), tj);
            else v = S::mul(mt, cj[r]);
            cj[r] = v;
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------
// Cluster-resident variant: the matrix lives in the shared memory of a thread-block cluster
// (rows dealt cyclically to the CTAs), so a column step costs ONE cluster barrier:
//   every CTA publishes its partial dot products  E_c = sum_{r>j} conj(x_r) a_{r,c}  (c >= j)
//   and the owner of row j publishes that row; after the barrier every CTA reads the partials
//   of its peers through distributed shared memory (DSMEM) and derives beta, tau and
//   v^H a_c = a_{j,c} + conj(s) E_c  locally (x = unscaled column, v = [1; s x], s = 1/(alpha-beta)),
//   i.e. norm and reflector application share one reduction.
// ---------------------------------------------------------------------------------------------
#ifdef QR_PROFILE
__device__ long long g_qr_prof[16];
#define QRP(i) if (blockIdx.x == 0 && threadIdx.x == 0) { long long t_ = clock64(); g_qr_prof[i] += t_ - qrp_t; qrp_t = t_; }
#else
#define QRP(i)
#endif
constexpr int QRC_THREADS = 512;
constexpr int QRC_CH = 8;          // columns handled together by one warp (independent FMA chains)

// dots of column `xj` (rows lr0..nloc) with the columns c0, c0+NW, ... of the local slab
template <bool CPLX>
__device__ __forceinline__ void slab_dots(const typename Sc<CPLX>::T* __restrict__ slab, int ldl, int nloc, int lr0,
                                          const typename Sc<CPLX>::T* __restrict__ xj, int cbeg, int cols,
                                          typename Sc<CPLX>::T* __restrict__ out) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    constexpr int NW = QRC_THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c0 = cbeg + warp; c0 < cols; c0 += NW * QRC_CH) {
        T acc[QRC_CH];
        const T* col[QRC_CH];
#pragma unroll
        for (int i = 0; i < QRC_CH; ++i) {
            acc[i] = S::zero();
            const int c = c0 + i * NW;
            col[i] = slab + (size_t)(c < cols ? c : c0) * ldl;
        }
        for (int lr = lr0 + lane; lr < nloc; lr += 32) {
            const T x = S::conj(xj[lr]);
#pragma unroll
            for (int i = 0; i < QRC_CH; ++i) acc[i] = S::fma(x, col[i][lr], acc[i]);
        }
        // folded butterfly: after the steps with offsets 16, 8, 4 lane l holds the sum (over the
        // lanes that agree with l in bits 0..1) of accumulator (l>>2)&7; two more steps finish it
        // (9 double shuffles instead of 40)
        {
            const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0, hi4 = (lane & 4) != 0;
            T b4[4], b2[2], b1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const T send = hi16 ? acc[i] : acc[i + 4];
                const T keep = hi16 ? acc[i + 4] : acc[i];
                b4[i] = S::add(keep, S::shfl_xor(send, 16));
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const T send = hi8 ? b4[i] : b4[i + 2];
                const T keep = hi8 ? b4[i + 2] : b4[i];
                b2[i] = S::add(keep, S::shfl_xor(send, 8));
            }
            {
                const T send = hi4 ? b2[0] : b2[1];
                const T keep = hi4 ? b2[1] : b2[0];
                b1 = S::add(keep, S::shfl_xor(send, 4));
            }
            b1 = S::add(b1, S::shfl_xor(b1, 2));
            b1 = S::add(b1, S::shfl_xor(b1, 1));
            // lane l (l & 3 == 0) holds accumulator index 4*bit4 + 2*bit3 + bit2
            const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            if ((lane & 3) == 0 && c0 + idx * NW < cols) out[c0 + idx * NW] = b1;
        }
    }
}

template <bool CPLX>
__global__ void __launch_bounds__(QRC_THREADS) qr_cluster_kernel(PtrBatch Ab, PtrBatch Rb, PtrBatch Taub, int rows, int cols,
                                                                  int ld, int CL, int wy_mode) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int mat = blockIdx.x / CL;
    T* __restrict__ A = reinterpret_cast<T*>(Ab.p[mat]);
    T* __restrict__ Rout = reinterpret_cast<T*>(Rb.p[mat]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = QRC_THREADS / 32;
    const int ldl = (rows + CL - 1) / CL;                 // local leading dimension
    const int nloc = (rows - rank + CL - 1) / CL;         // local rows: global r = lr*CL + rank

    extern __shared__ __align__(16) unsigned char qrc_smem[];
    T* slab = reinterpret_cast<T*>(qrc_smem);             // [cols][ldl]
    T* xbuf = slab + (size_t)cols * ldl;                  // [2][2*cols]  published partials + row j
    T* tot = xbuf + 4 * cols;                             // [2*cols]     gathered totals + row j
    T* tau = tot + 2 * cols;                              // [cols]

    for (int e = tid; e < cols * ldl; e += QRC_THREADS) {
        const int c = e / ldl, lr = e % ldl;
        slab[e] = lr < nloc ? A[(size_t)c * ld + (size_t)lr * CL + rank] : S::zero();
    }
    __syncthreads();
    cluster.sync();

#ifdef QR_PROFILE
    long long qrp_t = clock64();
#endif
    int par = 0;
    // after the cluster barrier: one DSMEM round trip brings the peers' partials (and the owner's
    // row j) into local shared memory; everything after that is CTA-local
    auto gather = [&](int cbeg, int owner, bool with_row) {
        const int off = par * 2 * cols;
        // one DSMEM load per thread: thread -> (column, peer rank); CL-lane groups reduce by shuffle
        const int ncol = cols - cbeg;
        const int per = QRC_THREADS / CL;                    // columns per pass
        for (int base = 0; base < ncol; base += per) {
            const int ci = base + tid / CL, r = tid % CL;
            T v = S::zero();
            if (ci < ncol) v = cluster.map_shared_rank(xbuf, r)[off + cbeg + ci];
            for (int o = 1; o < CL; o <<= 1) v = S::add(v, S::shfl_xor(v, o));
            if (ci < ncol && r == 0) tot[cbeg + ci] = v;
        }
        if (with_row) {
            const T* rowp = cluster.map_shared_rank(xbuf, owner) + off + cols;
            for (int c = cbeg + tid; c < cols; c += QRC_THREADS) tot[cols + c] = rowp[c];
        }
        __syncthreads();
    };

    const int kmax = min(rows, cols);
    // ---------------- Phase A ----------------
    for (int j = 0; j < kmax; ++j) {
        const int owner = j % CL, lj = j / CL;
        const int lr0 = (rank > owner) ? lj : lj + 1;
        const T* xj = slab + (size_t)j * ldl;
        T* xb = xbuf + par * 2 * cols;
        QRP(0)
        slab_dots<CPLX>(slab, ldl, nloc, lr0, xj, j, cols, xb);
        if (rank == owner)
            for (int c = j + tid; c < cols; c += QRC_THREADS) xb[cols + c] = slab[(size_t)c * ldl + lj];
        QRP(1)
        cluster.sync();
        QRP(2)
        gather(j, owner, true);
        QRP(3)
        // reflector parameters (computed redundantly by every thread from local shared memory)
        T tj = S::zero(), sc = S::zero(); double beta = 0.0;
        if (lane == 0) {
            // ||x||^2 = |alpha|^2 + tail; one rsqrt and one reciprocal:
            //   beta = -sign(Re alpha)||x||, 1/beta = -sign/||x||, tau = (beta-alpha)/beta, s = 1/(alpha-beta)
            const double tail = S::re(tot[j]);
            const T alpha = tot[cols + j];
            if (tail == 0.0 && S::im(alpha) == 0.0) { beta = S::re(alpha); }
            else {
                const double nsq = S::abs2(alpha) + tail;
                const double rn = rsqrt(nsq);
                const double sg = copysign(1.0, S::re(alpha));
                beta = -sg * (nsq * rn);
                const double ib = -sg * rn;
                tj = S::make((beta - S::re(alpha)) * ib, -S::im(alpha) * ib);
                const T amb = S::sub(alpha, S::make(beta, 0.0));
                sc = S::scale(S::conj(amb), 1.0 / S::abs2(amb));
            }
        }
        tj = S::make(__shfl_sync(0xffffffffu, S::re(tj), 0), __shfl_sync(0xffffffffu, S::im(tj), 0));
        sc = S::make(__shfl_sync(0xffffffffu, S::re(sc), 0), __shfl_sync(0xffffffffu, S::im(sc), 0));
        beta = __shfl_sync(0xffffffffu, beta, 0);
        QRP(4)
        if (S::abs2(tj) != 0.0) {
            const T ctau = S::conj(tj), csc = S::conj(sc);
            for (int c0 = j + 1 + warp; c0 < cols; c0 += NW * QRC_CH) {
                T fs[QRC_CH]; T* col[QRC_CH];
#pragma unroll
                for (int i = 0; i < QRC_CH; ++i) {
                    const int c = c0 + i * NW;
                    const bool ok = c < cols;
                    const int cc = ok ? c : c0;
                    col[i] = slab + (size_t)cc * ldl;
                    const T f = S::mul(ctau, S::add(tot[cols + cc], S::mul(csc, tot[cc])));
                    fs[i] = ok ? S::mul(f, sc) : S::zero();
                    if (ok && rank == owner && lane == 0) col[i][lj] = S::sub(col[i][lj], f);
                }
                for (int lr = lr0 + lane; lr < nloc; lr += 32) {
                    const T x = xj[lr];
#pragma unroll
                    for (int i = 0; i < QRC_CH; ++i)
                        if (c0 + i * NW < cols) col[i][lr] = S::sub(col[i][lr], S::mul(fs[i], x));
                }
            }
        }
        QRP(5)
        __syncthreads();
        QRP(6)
        {   // column j <- v (scaled), diagonal <- beta
            T* xw = slab + (size_t)j * ldl;
            for (int lr = lr0 + tid; lr < nloc; lr += QRC_THREADS) xw[lr] = S::mul(xw[lr], sc);
            if (rank == owner && tid == 0) xw[lj] = S::make(beta, 0.0);
            if (tid == 0) tau[j] = tj;
        }
        par ^= 1;
    }
    __syncthreads();
    // ---------------- save R ----------------
    if (Rout != nullptr) {
        for (int e = tid; e < cols * ldl; e += QRC_THREADS) {
            const int c = e / ldl, lr = e % ldl;
            const int r = lr * CL + rank;
            if (lr < nloc && r < cols) Rout[(size_t)c * cols + r] = (r <= c) ? slab[e] : S::zero();
        }
    }
    __syncthreads();
    if (wy_mode) {
        // WY mode (panel of the blocked factorisation): A <- V (explicit unit diagonal, zeros above), Taub <- tau
        T* tau_out = reinterpret_cast<T*>(Taub.p[mat]);
        if (rank == 0) for (int c = tid; c < cols; c += QRC_THREADS) tau_out[c] = c < kmax ? tau[c] : S::zero();
        for (int e = tid; e < cols * ldl; e += QRC_THREADS) {
            const int c = e / ldl, lr = e % ldl;
            const int r = lr * CL + rank;
            if (lr < nloc) A[(size_t)c * ld + r] = r > c ? slab[e] : (r == c ? S::one() : S::zero());
        }
        cluster.sync();
        return;
    }
    // ---------------- Phase B: explicit Q in place ----------------
    for (int j = cols - 1; j >= kmax; --j)
        for (int lr = tid; lr < nloc; lr += QRC_THREADS) slab[(size_t)j * ldl + lr] = S::zero();
    __syncthreads();
    for (int j = kmax - 1; j >= 0; --j) {
        const int owner = j % CL, lj = j / CL;
        const int lr0 = (rank > owner) ? lj : lj + 1;
        T* vj = slab + (size_t)j * ldl;
        const T tj = tau[j];
        if (j + 1 < cols && S::abs2(tj) != 0.0) {       // tau is identical on every CTA: uniform branch
            T* xb = xbuf + par * 2 * cols;
            QRP(7)
            slab_dots<CPLX>(slab, ldl, nloc, lr0, vj, j + 1, cols, xb);
            QRP(8)
            cluster.sync();
            QRP(9)
            gather(j + 1, owner, false);
            QRP(10)
            for (int c0 = j + 1 + warp; c0 < cols; c0 += NW * QRC_CH) {
                T fs[QRC_CH]; T* col[QRC_CH];
#pragma unroll
                for (int i = 0; i < QRC_CH; ++i) {
                    const int c = c0 + i * NW;
                    const bool ok = c < cols;
                    const int cc = ok ? c : c0;
                    col[i] = slab + (size_t)cc * ldl;
                    fs[i] = ok ? S::mul(tj, tot[cc]) : S::zero();
                    if (ok && rank == owner && lane == 0) col[i][lj] = S::sub(S::zero(), fs[i]);
                }
                for (int lr = lr0 + lane; lr < nloc; lr += 32) {
                    const T x = vj[lr];
#pragma unroll
                    for (int i = 0; i < QRC_CH; ++i)
                        if (c0 + i * NW < cols) col[i][lr] = S::sub(col[i][lr], S::mul(fs[i], x));
                }
            }
            QRP(11)
            __syncthreads();
            par ^= 1;
        }
        const T mt = S::sub(S::zero(), tj);
        for (int lr = tid; lr < nloc; lr += QRC_THREADS) {
            T v;
            if (rank == owner && lr == lj) v = S::sub(S::one(), tj);
            else if (lr >= lr0) v = S::mul(mt, vj[lr]);
            else v = S::zero();
            vj[lr] = v;
        }
        __syncthreads();
    }
    for (int e = tid; e < cols * ldl; e += QRC_THREADS) {
        const int c = e / ldl, lr = e % ldl;
        if (lr < nloc) A[(size_t)c * ld + (size_t)lr * CL + rank] = slab[e];
    }
    cluster.sync();      // no CTA may exit while peers can still read its shared memory
}


// ---------------------------------------------------------------------------------------------
// Register-resident cluster variant (the fast path for sketches up to 1024 x 128): the slab of a
// CTA lives in REGISTERS (warp w holds columns w, w+16, ..; lane l holds local rows l, l+32, ..),
// so the column step issues no shared-memory traffic for the matrix at all: the Householder
// vector is broadcast through a small shared array, the partial dots are folded with shuffles,
// published, and combined over the cluster exactly as above.  The ncu profile of the
// shared-memory variant showed it instruction-bound (830 warp instructions per step and warp,
// IPC 1.8, profiles/r1_qr_cluster_smem.md); here the step is several times shorter.
// ---------------------------------------------------------------------------------------------
template <bool CPLX, int RPL>
__global__ void __launch_bounds__(QRC_THREADS) qr_reg_kernel(PtrBatch Ab, PtrBatch Rb, PtrBatch Taub, int rows, int cols,
                                                              int ld, int CL, int wy_mode) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    namespace cg = cooperative_groups;
    constexpr int CPW = 8, NW = QRC_THREADS / 32;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int mat = blockIdx.x / CL;
    T* __restrict__ A = reinterpret_cast<T*>(Ab.p[mat]);
    T* __restrict__ Rout = reinterpret_cast<T*>(Rb.p[mat]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nloc = (rows - rank + CL - 1) / CL;

    extern __shared__ __align__(16) unsigned char qrr_smem[];
    T* xs = reinterpret_cast<T*>(qrr_smem);               // [32*RPL]   broadcast of x / v
    T* xbuf = xs + 32 * RPL;                              // [2][2*cols]
    T* tot = xbuf + 4 * cols;                             // [2*cols]
    T* tau = tot + 2 * cols;                              // [cols]

    T a[CPW][RPL];
    int grow[RPL];                                        // global row of register row r (huge if padding)
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int lr = lane + 32 * r;
        grow[r] = lr < nloc ? lr * CL + rank : 0x3fffffff;
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
            const int c = warp + NW * i;
            a[i][r] = (lr < nloc && c < cols) ? A[(size_t)c * ld + (size_t)lr * CL + rank] : S::zero();
        }
    }
    for (int c = tid; c < 4 * cols; c += QRC_THREADS) xbuf[c] = S::zero();
    __syncthreads();
    cluster.sync();

#ifdef QR_PROFILE
    long long qrp_t = clock64();
#endif
    int par = 0;
    auto gather = [&](int cbeg, int owner, bool with_row) {
        const int off = par * 2 * cols;
        const int ncol = cols - cbeg;
        const int per = QRC_THREADS / CL;
        // all remote (DSMEM) loads of the step are issued before anything consumes them: up to two column
        // passes (cols <= 128, per >= 64) and the owner's row j -- one remote round trip instead of three
        const int r = tid % CL;
        const int ci0 = tid / CL, ci1 = per + tid / CL;
        const T* rem = cluster.map_shared_rank(xbuf, r) + off + cbeg;
        T v0 = S::zero(), v1 = S::zero(), rw = S::zero();
        if (ci0 < ncol) v0 = rem[ci0];
        if (ci1 < ncol) v1 = rem[ci1];
        const int rc = cbeg + tid;
        if (with_row && rc < cols) rw = (cluster.map_shared_rank(xbuf, owner) + off + cols)[rc];
        for (int o = 1; o < CL; o <<= 1) { v0 = S::add(v0, S::shfl_xor(v0, o)); v1 = S::add(v1, S::shfl_xor(v1, o)); }
        if (r == 0) {
            if (ci0 < ncol) tot[cbeg + ci0] = v0;
            if (ci1 < ncol) tot[cbeg + ci1] = v1;
        }
        if (with_row && rc < cols) tot[cols + rc] = rw;
        for (int base = 2 * per; base < ncol; base += per) {      // (not reached for cols <= 128)
            const int ci = base + tid / CL;
            T v = S::zero();
            if (ci < ncol) v = rem[ci];
            for (int o = 1; o < CL; o <<= 1) v = S::add(v, S::shfl_xor(v, o));
            if (ci < ncol && r == 0) tot[cbeg + ci] = v;
        }
        __syncthreads();
    };
    // folded butterfly over the 8 per-column accumulators of a warp -> out[warp + NW*idx]
    auto fold_publish = [&](T (&acc)[CPW], T* out) {
        const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0, hi4 = (lane & 4) != 0;
        T b4[4], b2[2], b1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const T send = hi16 ? acc[i] : acc[i + 4];
            const T keep = hi16 ? acc[i + 4] : acc[i];
            b4[i] = S::add(keep, S::shfl_xor(send, 16));
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const T send = hi8 ? b4[i] : b4[i + 2];
            const T keep = hi8 ? b4[i + 2] : b4[i];
            b2[i] = S::add(keep, S::shfl_xor(send, 8));
        }
        {
            const T send = hi4 ? b2[0] : b2[1];
            const T keep = hi4 ? b2[1] : b2[0];
            b1 = S::add(keep, S::shfl_xor(send, 4));
        }
        b1 = S::add(b1, S::shfl_xor(b1, 2));
        b1 = S::add(b1, S::shfl_xor(b1, 1));
        const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        const int c = warp + NW * idx;
        if ((lane & 3) == 0 && c < cols) out[c] = b1;
    };

    const int kmax = min(rows, cols);
    // ---------------- Phase A ----------------
    for (int j = 0; j < kmax; ++j) {
        const int owner = j % CL, lj = j / CL;            // row j: CTA `owner`, local row lj
        const int wj = j % NW, ij = j / NW;               // column j: warp wj, register slot ij
        const bool mine_row = (rank == owner) && (lane == (lj & 31));
        const int rj = lj >> 5;
        T* xb = xbuf + par * 2 * cols;
        QRP(0)
        if (warp == wj) {
#pragma unroll
            for (int i = 0; i < CPW; ++i)
                if (i == ij) {
#pragma unroll
                    for (int r = 0; r < RPL; ++r) xs[lane + 32 * r] = a[i][r];
                }
        }
        if (mine_row) {                                   // publish row j of the columns of this warp
#pragma unroll
            for (int i = 0; i < CPW; ++i) {
                const int c = warp + NW * i;
#pragma unroll
                for (int r = 0; r < RPL; ++r)
                    if (r == rj && c < cols) xb[cols + c] = a[i][r];
            }
        }
        __syncthreads();
        QRP(1)
        T xr[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) xr[r] = (grow[r] > j && grow[r] < 0x3fffffff) ? xs[lane + 32 * r] : S::zero();
        {
            T acc[CPW];
#pragma unroll
            for (int i = 0; i < CPW; ++i) {
                acc[i] = S::zero();
#pragma unroll
                for (int r = 0; r < RPL; ++r) acc[i] = S::fma(S::conj(xr[r]), a[i][r], acc[i]);
            }
            fold_publish(acc, xb);
        }
        QRP(2)
        cluster.sync();
        QRP(3)
        gather(j, owner, true);
        QRP(4)
        T tj = S::zero(), sc = S::zero(); double beta = 0.0;
        if (lane == 0) {
            // ||x||^2 = |alpha|^2 + tail; one rsqrt and one reciprocal:
            //   beta = -sign(Re alpha)||x||, 1/beta = -sign/||x||, tau = (beta-alpha)/beta, s = 1/(alpha-beta)
            const double tail = S::re(tot[j]);
            const T alpha = tot[cols + j];
            if (tail == 0.0 && S::im(alpha) == 0.0) { beta = S::re(alpha); }
            else {
                const double nsq = S::abs2(alpha) + tail;
                const double rn = rsqrt(nsq);
                const double sg = copysign(1.0, S::re(alpha));
                beta = -sg * (nsq * rn);
                const double ib = -sg * rn;
                tj = S::make((beta - S::re(alpha)) * ib, -S::im(alpha) * ib);
                const T amb = S::sub(alpha, S::make(beta, 0.0));
                sc = S::scale(S::conj(amb), 1.0 / S::abs2(amb));
            }
        }
        tj = S::make(__shfl_sync(0xffffffffu, S::re(tj), 0), __shfl_sync(0xffffffffu, S::im(tj), 0));
        sc = S::make(__shfl_sync(0xffffffffu, S::re(sc), 0), __shfl_sync(0xffffffffu, S::im(sc), 0));
        beta = __shfl_sync(0xffffffffu, beta, 0);
        if (tid == 0) tau[j] = tj;
        QRP(5)
        const T ctau = S::conj(tj), csc = S::conj(sc);
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
            const int c = warp + NW * i;
            if (c > j && c < cols) {
                const T f = S::mul(ctau, S::add(tot[cols + c], S::mul(csc, tot[c])));
                const T fs = S::mul(f, sc);
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    a[i][r] = S::sub(a[i][r], S::mul(fs, xr[r]));
                    if (mine_row && r == rj) a[i][r] = S::sub(a[i][r], f);
                }
            } else if (c == j) {                           // column j <- v, diagonal <- beta
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    if (grow[r] > j) a[i][r] = S::mul(a[i][r], sc);
                    if (mine_row && r == rj) a[i][r] = S::make(beta, 0.0);
                }
            }
        }
        QRP(6)
        par ^= 1;
    }
    // ---------------- save R ----------------
    if (Rout != nullptr) {
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
            const int c = warp + NW * i;
#pragma unroll
            for (int r = 0; r < RPL; ++r)
                if (c < cols && grow[r] < cols) Rout[(size_t)c * cols + grow[r]] = (grow[r] <= c) ? a[i][r] : S::zero();
        }
    }
    __syncthreads();
    if (wy_mode) {
        // WY mode: leave the reflectors V (explicit unit diagonal, zeros above) in place of A and
        // tau in Taub; the explicit Q = E - V (T V1^H) is then formed by GEMMs (see qr_wy_* below)
        T* tau_out = reinterpret_cast<T*>(Taub.p[mat]);
        if (rank == 0) for (int c = tid; c < cols; c += QRC_THREADS) tau_out[c] = c < kmax ? tau[c] : S::zero();
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int lr = lane + 32 * r;
#pragma unroll
            for (int i = 0; i < CPW; ++i) {
                const int c = warp + NW * i;
                if (lr < nloc && c < cols)
                    A[(size_t)c * ld + (size_t)lr * CL + rank] = grow[r] > c ? a[i][r] : (grow[r] == c ? S::one() : S::zero());
            }
        }
        cluster.sync();
        return;
    }
    // ---------------- Phase B: explicit Q in place ----------------
    for (int j = kmax - 1; j >= 0; --j) {
        const int owner = j % CL, lj = j / CL;
        const int wj = j % NW, ij = j / NW;
        const bool mine_row = (rank == owner) && (lane == (lj & 31));
        const int rj = lj >> 5;
        const T tj = tau[j];
        T* xb = xbuf + par * 2 * cols;
        if (warp == wj) {
#pragma unroll
            for (int i = 0; i < CPW; ++i)
                if (i == ij) {
#pragma unroll
                    for (int r = 0; r < RPL; ++r) xs[lane + 32 * r] = a[i][r];
                }
        }
        __syncthreads();
        T vr[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) vr[r] = (grow[r] > j && grow[r] < 0x3fffffff) ? xs[lane + 32 * r] : S::zero();
        if (j + 1 < cols && S::abs2(tj) != 0.0) {          // uniform over the cluster
            T acc[CPW];
#pragma unroll
            for (int i = 0; i < CPW; ++i) {
                acc[i] = S::zero();
                const int c = warp + NW * i;
                if (c > j) {
#pragma unroll
                    for (int r = 0; r < RPL; ++r) acc[i] = S::fma(S::conj(vr[r]), a[i][r], acc[i]);
                }
            }
            fold_publish(acc, xb);
            cluster.sync();
            gather(j + 1, owner, false);
#pragma unroll
            for (int i = 0; i < CPW; ++i) {
                const int c = warp + NW * i;
                if (c > j && c < cols) {
                    const T f = S::mul(tj, tot[c]);
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        a[i][r] = S::sub(a[i][r], S::mul(f, vr[r]));
                        if (mine_row && r == rj) a[i][r] = S::sub(S::zero(), f);
                    }
                }
            }
            par ^= 1;
        } else {
            __syncthreads();                                // xs is rewritten by the next step
        }
        if (warp == wj) {                                  // column j <- H_j e_j
            const T mt = S::sub(S::zero(), tj);
#pragma unroll
            for (int i = 0; i < CPW; ++i)
                if (i == ij) {
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        T v = S::mul(mt, vr[r]);           // zero above the diagonal (vr masked)
                        if (mine_row && r == rj) v = S::sub(S::one(), tj);
                        a[i][r] = v;
                    }
                }
        }
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int lr = lane + 32 * r;
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
            const int c = warp + NW * i;
            if (lr < nloc && c < cols) A[(size_t)c * ld + (size_t)lr * CL + rank] = a[i][r];
        }
    }
    cluster.sync();
}

template <bool CPLX, int RPL>
static void qr_reg_run(const PtrBatch& A, const PtrBatch& Rout, const PtrBatch& Tau, int nb, int rows, int cols, int ld, int cl,
                       int wy_mode, cudaStream_t stream) {
    auto kern = qr_reg_kernel<CPLX, RPL>;
    const size_t smem = ((size_t)32 * RPL + 7 * (size_t)cols) * (CPLX ? 16 : 8);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nb * cl);
    cfg.blockDim = dim3(QRC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CTMB_CUDA(cudaLaunchKernelEx(&cfg, kern, A, Rout, Tau, rows, cols, ld, cl, wy_mode));
}

// register path: cols <= 128 and rows <= 8 * 32 * RPL
static bool qr_reg_shape(int rows, int cols, bool cplx, int& rpl_out, int& cl_out) {
    if (cols > 128) return false;
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("CTMB_QR_REG"); mode = e ? atoi(e) : 1; }
    if (!mode) return false;
    static int rpl_min = -1;                             // tuning knob: prefer fewer CTAs with more rows per lane
    if (rpl_min < 0) { const char* e = getenv("CTMB_QR_RPL_MIN"); rpl_min = e ? atoi(e) : 1; }
    for (int rpl : {1, 2, 4}) {
        if (cplx && rpl == 4) break;
        if (rpl < rpl_min && !(cplx && rpl == 2)) continue;
        for (int cl = 1; cl <= 8; cl <<= 1) {
            if ((rows + cl - 1) / cl > 32 * rpl) continue;
            rpl_out = rpl; cl_out = cl;
            return true;
        }
    }
    return false;
}

static bool qr_reg_try(const PtrBatch& A, const PtrBatch& Rout, const PtrBatch& Tau, int nb, int rows, int cols, int ld,
                       bool cplx, int wy_mode, cudaStream_t stream) {
    int rpl = 0, cl = 0;
    if (!qr_reg_shape(rows, cols, cplx, rpl, cl)) return false;
    if (rpl == 1) { if (cplx) qr_reg_run<true, 1>(A, Rout, Tau, nb, rows, cols, ld, cl, wy_mode, stream); else qr_reg_run<false, 1>(A, Rout, Tau, nb, rows, cols, ld, cl, wy_mode, stream); }
    else if (rpl == 2) { if (cplx) qr_reg_run<true, 2>(A, Rout, Tau, nb, rows, cols, ld, cl, wy_mode, stream); else qr_reg_run<false, 2>(A, Rout, Tau, nb, rows, cols, ld, cl, wy_mode, stream); }
    else { qr_reg_run<false, 4>(A, Rout, Tau, nb, rows, cols, ld, cl, wy_mode, stream); }
    return true;
}

// dynamic shared memory of wy_tsolve_kernel: X (k x k) and either all of G or the 16 columns of G one block step needs
constexpr size_t WY_TSOLVE_SMEM_MAX = 200 * 1024;
static size_t wy_tsolve_smem(int k, bool cplx, int* g_in_smem) {
    const size_t es = cplx ? 16 : 8;
    const int both = 2 * (size_t)k * k * es <= WY_TSOLVE_SMEM_MAX;
    if (g_in_smem) *g_in_smem = both;
    return both ? 2 * (size_t)k * k * es : ((size_t)k * k + 16 * (size_t)k) * es;
}

bool qr_wy_supported(int rows, int cols, bool cplx) {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("CTMB_QR_WY"); mode = e ? atoi(e) : 1; }
    int rpl, cl;
    // wy_tsolve keeps X and a block of G in shared memory (complex: k <= 105; wider sketches take the plain / blocked QR)
    if (cols > 128 || wy_tsolve_smem(cols, cplx, nullptr) > WY_TSOLVE_SMEM_MAX) return false;
    return mode && qr_reg_shape(rows, cols, cplx, rpl, cl);
}

// Phase A only: A <- V (explicit unit lower trapezoidal), Tau <- tau, Rout <- R (optional)
void qr_wy_factor_launch(const PtrBatch& A, const PtrBatch& Rout, const PtrBatch& Tau, int nb, int rows, int cols,
                         int ld, bool cplx, cudaStream_t stream) {
    CTMB_CHECK(qr_reg_try(A, Rout, Tau, nb, rows, cols, ld, cplx, 1, stream), "qr_wy_factor: unsupported shape");
}

// X = T V1^H (k x k, stored [c][s]) from the Gram matrix G = V^H V ([t][s] = v_s^H v_t), tau and
// the top k x k block of V:  T^-1 = striu(G) + diag(1/tau)  =>  back substitution
//   X[s,c] = tau_s ( V1^H[s,c] - sum_{t>s} G[s,t] X[t,c] ),   s = k-1 .. 0.
// Every column of X is an independent recurrence: ONE THREAD PER COLUMN walks it with no
// synchronisation at all (it only reads what it wrote itself).  G(s,t) is the same address for
// all threads (broadcast), X[t][c] is consecutive in c (conflict-free).  The first version split a
// column over four lanes and read V / tau from global memory inside the loop: 4-way bank conflicts
// plus an L2 round trip per step made it 122 us for k = 96 (profiles/r1_c2_qr_jacobi.md).
constexpr int TS_THREADS = 512;     // all of them stage the operands; the first k walk the recurrences
template <bool CPLX>
__global__ void __launch_bounds__(TS_THREADS) wy_tsolve_kernel(PtrBatch Gb, PtrBatch Taub, PtrBatch Vb, PtrBatch Xb, int k, int ldv,
                                                               int g_in_smem) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const T* __restrict__ G = reinterpret_cast<const T*>(Gb.p[blockIdx.x]);
    const T* __restrict__ tau = reinterpret_cast<const T*>(Taub.p[blockIdx.x]);
    const T* __restrict__ V = reinterpret_cast<const T*>(Vb.p[blockIdx.x]);
    T* __restrict__ X = reinterpret_cast<T*>(Xb.p[blockIdx.x]);
    extern __shared__ __align__(16) unsigned char wy_smem[];
    T* Xs = reinterpret_cast<T*>(wy_smem);                // [t][c]
    T* Gs = Xs + (size_t)k * k;                           // [s][t] (optional)
    __shared__ T taus[128];
    const int tid = threadIdx.x;
    // Staging: every load is independent, so keep many in flight per thread (batches of 8): with one
    // outstanding load per thread this kernel was nothing but 2 x 72 serial L2 round trips (~120 us).
    // Xs[s][c] <- V1^H[s,c]  (unit lower trapezoidal V: conj(V[c,s]) above the diagonal, 1 on it, 0 below)
    const int kk2 = k * k;
    for (int e0 = tid; e0 < kk2; e0 += TS_THREADS * 8) {
        T v[8], g[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * TS_THREADS;
            const int s2 = e / k, c2 = e % k;             // V[(size_t)s2 * ldv + c2] = V(row c2, column s2)
            v[u] = (e < kk2 && c2 > s2 && ldv > 0) ? V[(size_t)s2 * ldv + c2] : S::zero();
            g[u] = (e < kk2 && g_in_smem) ? G[e] : S::zero();
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * TS_THREADS;
            if (e < kk2) {
                const int s2 = e / k, c2 = e % k;
                Xs[e] = c2 > s2 ? S::conj(v[u]) : (c2 == s2 ? S::one() : S::zero());
                if (g_in_smem) Gs[c2 * k + s2] = g[u];    // G[e] = G[t = s2][s = c2]  ->  Gs[s][t]
            }
        }
    }
    for (int e = tid; e < k; e += TS_THREADS) taus[e] = tau[e];
    __syncthreads();
    {
        // Blocked back substitution (row blocks of NB, last block first).  Phase 1: inside the block every
        // column is an independent short recurrence (one thread per column, the block of X in registers, no
        // synchronisation).  Phase 2: the finished rows are eliminated from all rows above them.  G(r,t) is a
        // warp-wide broadcast (from shared memory, or through L1 when G and X do not both fit), X(t,c) is
        // consecutive in c.  A single thread walking a whole column (4560 dependent FMAs for k = 96) took 74 us.
        constexpr int NB = 16;
        constexpr int RG = TS_THREADS / 128;                 // row groups of the elimination phase
        // when G and X do not both fit (k = 128 real), only the 16 columns of G the current block needs are staged
        T* Gp = Gs;                                          // [NB][k], column t of G contiguous
        for (int s1 = k; s1 > 0; s1 -= NB) {
            const int s0 = s1 > NB ? s1 - NB : 0;
            const int nbk = s1 - s0;
            if (!g_in_smem) {
                for (int e = tid; e < nbk * k; e += TS_THREADS) Gp[e] = G[(size_t)(s0 + e / k) * k + e % k];
                __syncthreads();
            }
            auto Gat = [&](int row, int col) -> T { return g_in_smem ? Gs[(size_t)row * k + col] : Gp[(size_t)(col - s0) * k + row]; };
            // (rows below the diagonal of X are zero and come out as zero: no special casing)
            const int c = s0 + tid;
            if (c < k) {
                T x[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) x[i] = i < nbk ? Xs[(s0 + i) * k + c] : S::zero();
#pragma unroll
                for (int i = NB - 1; i >= 0; --i) {
                    if (i < nbk) {
                        T acc = S::zero();
#pragma unroll
                        for (int t = i + 1; t < NB; ++t) if (t < nbk) acc = S::fma(Gat(s0 + i, s0 + t), x[t], acc);
                        x[i] = S::mul(taus[s0 + i], S::sub(x[i], acc));
                    }
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) if (i < nbk) Xs[(s0 + i) * k + c] = x[i];
            }
            __syncthreads();
            // thread (g, cc) keeps the block of column cc in registers and walks rows g, g+RG, ...
            const int cc = s0 + (tid & 127), g = tid >> 7;
            if (cc < k && s0 > 0) {
                T x[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) x[i] = i < nbk ? Xs[(s0 + i) * k + cc] : S::zero();
                for (int r = g; r < s0; r += RG) {
                    T a0 = S::zero(), a1 = S::zero();
#pragma unroll
                    for (int t = 0; t < NB; t += 2) {
                        if (t < nbk) a0 = S::fma(Gat(r, s0 + t), x[t], a0);
                        if (t + 1 < nbk) a1 = S::fma(Gat(r, s0 + t + 1), x[t + 1], a1);
                    }
                    Xs[r * k + cc] = S::sub(Xs[r * k + cc], S::add(a0, a1));
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    for (int e = tid; e < k * k; e += TS_THREADS) { const int c2 = e / k, s2 = e % k; X[e] = Xs[s2 * k + c2]; }
}

void wy_tsolve_launch(const PtrBatch& G, const PtrBatch& Tau, const PtrBatch& V, const PtrBatch& X, int nb, int k,
                      int ldv, bool cplx, cudaStream_t stream) {
    int g_in_smem = 0;
    const size_t smem = wy_tsolve_smem(k, cplx, &g_in_smem);
    CTMB_CHECK(k <= 128 && smem <= WY_TSOLVE_SMEM_MAX, "wy_tsolve: k too large");
    if (cplx) {
        auto kern = wy_tsolve_kernel<true>;
        static bool set = false;
        if (!set) { CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WY_TSOLVE_SMEM_MAX)); set = true; }
        kern<<<nb, TS_THREADS, smem, stream>>>(G, Tau, V, X, k, ldv, g_in_smem);
    } else {
        auto kern = wy_tsolve_kernel<false>;
        static bool set = false;
        if (!set) { CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WY_TSOLVE_SMEM_MAX)); set = true; }
        kern<<<nb, TS_THREADS, smem, stream>>>(G, Tau, V, X, k, ldv, g_in_smem);
    }
    CTMB_CUDA(cudaGetLastError());
}

// Q[c][c] += 1 on the leading k x k diagonal of column-major n x k matrices
template <bool CPLX>
__global__ void add_identity_kernel(PtrBatch Qb, int k, int ld) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    T* Q = reinterpret_cast<T*>(Qb.p[blockIdx.x]);
    for (int c = threadIdx.x; c < k; c += blockDim.x) Q[(size_t)c * ld + c] = S::add(Q[(size_t)c * ld + c], S::one());
}
void add_identity_launch(const PtrBatch& Q, int nb, int k, int ld, bool cplx, cudaStream_t stream) {
    if (cplx) add_identity_kernel<true><<<nb, 128, 0, stream>>>(Q, k, ld);
    else add_identity_kernel<false><<<nb, 128, 0, stream>>>(Q, k, ld);
    CTMB_CUDA(cudaGetLastError());
}

static int g_qr_cluster_override = -1;

// smallest cluster size whose per-CTA slab fits in shared memory (0: does not fit at all)
static int qr_cluster_size(int rows, int cols, bool cplx, size_t& smem) {
    if (g_qr_cluster_override < 0) {
        const char* e = getenv("CTMB_QR_CLUSTER");
        g_qr_cluster_override = e ? atoi(e) : 0;
    }
    const size_t es = cplx ? 16 : 8;
    const size_t limit = 200 * 1024;
    for (int cl = 1; cl <= 8; cl <<= 1) {
        const int ldl = (rows + cl - 1) / cl;
        smem = ((size_t)cols * ldl + 7 * (size_t)cols) * es;
        const bool forced = g_qr_cluster_override > 0;
        if (smem <= limit && (!forced || cl >= g_qr_cluster_override)) return cl;
    }
    return 0;
}

template <bool CPLX>
static void qr_cluster_run(const PtrBatch& A, const PtrBatch& Rout, const PtrBatch& Tau, int nb, int rows, int cols, int ld, int cl,
                           size_t smem, int wy_mode, cudaStream_t stream) {
    auto kern = qr_cluster_kernel<CPLX>;
    static size_t set = 0;
    if (smem > set) {
        CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024 + 4096)));
        set = 200 * 1024 + 4096;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nb * cl);
    cfg.blockDim = dim3(QRC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CTMB_CUDA(cudaLaunchKernelEx(&cfg, kern, A, Rout, Tau, rows, cols, ld, cl, wy_mode));
}

void qr_launch(const PtrBatch& A, const PtrBatch& Rout, int nb, int rows, int cols, int ld,
               bool cplx, cudaStream_t stream) {
    CTMB_CHECK(nb >= 1 && nb <= TC_MAX_BATCH, "bad batch");
    CTMB_CHECK(rows >= cols, "qr: expects rows >= cols");
    { PtrBatch none{}; if (qr_reg_try(A, Rout, none, nb, rows, cols, ld, cplx, 0, stream)) return; }
    size_t csmem = 0;
    const int cl = qr_cluster_size(rows, cols, cplx, csmem);
    if (cl > 0) {
        PtrBatch none{};
        if (cplx) qr_cluster_run<true>(A, Rout, none, nb, rows, cols, ld, cl, csmem, 0, stream);
        else qr_cluster_run<false>(A, Rout, none, nb, rows, cols, ld, cl, csmem, 0, stream);
        return;
    }
    size_t smem = (size_t)cols * (cplx ? 16 : 8);
    if (cplx) qr_kernel<true><<<nb, QR_THREADS, smem, stream>>>(A, Rout, rows, cols, ld);
    else qr_kernel<false><<<nb, QR_THREADS, smem, stream>>>(A, Rout, rows, cols, ld);
    CTMB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// Panels of the blocked factorisation (sketches that fit neither the registers nor the shared
// memory of one cluster, e.g. 1536 x 192 complex at D=4 chi=96 or 16384 x 512 at D=8 chi=256).
// ---------------------------------------------------------------------------------------------
// widest panel (power of two, <= 128, <= cols) whose rows x b slab fits the registers or the shared memory of a cluster
int qr_panel_width(int rows, int cols, bool cplx) {
    for (int b = 128; b >= 4; b >>= 1) {
        if (b > cols && (b >> 1) >= cols) continue;
        const int bb = std::min(b, cols);
        int rpl, cl; size_t smem;
        if (qr_reg_shape(rows, bb, cplx, rpl, cl)) return bb;
        if (qr_cluster_size(rows, bb, cplx, smem) > 0) return bb;
    }
    return 0;
}
// factor one panel in WY form: A (rows x b, leading dimension ld) <- V, Tau <- tau, Rpp (b x b column-major) <- R
void qr_panel_launch(const PtrBatch& A, const PtrBatch& Rpp, const PtrBatch& Tau, int nb, int rows, int b, int ld,
                     bool cplx, cudaStream_t stream) {
    if (qr_reg_try(A, Rpp, Tau, nb, rows, b, ld, cplx, 1, stream)) return;
    size_t csmem = 0;
    const int cl = qr_cluster_size(rows, b, cplx, csmem);
    CTMB_CHECK(cl > 0, "qr_panel: panel does not fit a cluster");
    if (cplx) qr_cluster_run<true>(A, Rpp, Tau, nb, rows, b, ld, cl, csmem, 1, stream);
    else qr_cluster_run<false>(A, Rpp, Tau, nb, rows, b, ld, cl, csmem, 1, stream);
}

// rows j0 .. j0+b-1 of the k x k factor R (column-major, ld = k): the diagonal block from Rpp, the
// blocks to its right from the updated trailing matrix A (leading dimension ld), zeros to the left
template <bool CPLX>
__global__ void qr_copy_r_kernel(PtrBatch Ab, PtrBatch Rppb, PtrBatch Rb, int k, int j0, int b, int ld) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const T* A = reinterpret_cast<const T*>(Ab.p[blockIdx.y]);
    const T* Rpp = reinterpret_cast<const T*>(Rppb.p[blockIdx.y]);
    T* R = reinterpret_cast<T*>(Rb.p[blockIdx.y]);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < b * k; e += gridDim.x * blockDim.x) {
        const int c = e / b, r = e % b;
        T v = S::zero();
        if (c >= j0 + b) v = A[(size_t)c * ld + j0 + r];
        else if (c >= j0) v = Rpp[(size_t)(c - j0) * b + r];
        R[(size_t)c * k + j0 + r] = v;
    }
}
void qr_copy_r_launch(const PtrBatch& A, const PtrBatch& Rpp, const PtrBatch& R, int nb, int k, int j0, int b, int ld,
                      bool cplx, cudaStream_t stream) {
    dim3 grid((unsigned)std::min(64, (b * k + 255) / 256), nb);
    if (cplx) qr_copy_r_kernel<true><<<grid, 256, 0, stream>>>(A, Rpp, R, k, j0, b, ld);
    else qr_copy_r_kernel<false><<<grid, 256, 0, stream>>>(A, Rpp, R, k, j0, b, ld);
    CTMB_CUDA(cudaGetLastError());
}

// V[r, j0+c] (r >= J0) <- reflector entries of the leaf panel at columns j0..j0+bw of A: A below the leaf's
// diagonal, 1 on it, 0 above (rows J0 .. j0+c-1 of A hold R entries of the enclosing super panel)
template <bool CPLX>
__global__ void qr_copy_v_kernel(PtrBatch Ab, PtrBatch Vb, int rows, int J0, int j0, int bw) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const T* A = reinterpret_cast<const T*>(Ab.p[blockIdx.y]);
    T* V = reinterpret_cast<T*>(Vb.p[blockIdx.y]);
    const int prow = rows - J0;
    const long long n = (long long)prow * bw;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e / prow), r = J0 + (int)(e % prow);
        const size_t idx = (size_t)(j0 + c) * rows + r;
        V[idx] = r > j0 + c ? A[idx] : (r == j0 + c ? S::one() : S::zero());
    }
}
void qr_copy_v_launch(const PtrBatch& A, const PtrBatch& V, int nb, int rows, int J0, int j0, int bw, bool cplx,
                      cudaStream_t stream) {
    const long long n = (long long)(rows - J0) * bw;
    dim3 grid((unsigned)std::min<long long>(512, (n + 255) / 256), nb);
    if (cplx) qr_copy_v_kernel<true><<<grid, 256, 0, stream>>>(A, V, rows, J0, j0, bw);
    else qr_copy_v_kernel<false><<<grid, 256, 0, stream>>>(A, V, rows, J0, j0, bw);
    CTMB_CUDA(cudaGetLastError());
}

// Q <- first k columns of the identity (rows x k column-major, ld = rows)
template <bool CPLX>
__global__ void set_identity_kernel(PtrBatch Qb, int rows, int k) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    T* Q = reinterpret_cast<T*>(Qb.p[blockIdx.y]);
    const long long n = (long long)rows * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        Q[e] = (e / rows == e % rows) ? S::one() : S::zero();
}
void set_identity_launch(const PtrBatch& Q, int nb, int rows, int k, bool cplx, cudaStream_t stream) {
    dim3 grid((unsigned)std::min<long long>(1024, ((long long)rows * k + 255) / 256), nb);
    if (cplx) set_identity_kernel<true><<<grid, 256, 0, stream>>>(Q, rows, k);
    else set_identity_kernel<false><<<grid, 256, 0, stream>>>(Q, rows, k);
    CTMB_CUDA(cudaGetLastError());
}

#ifdef QR_PROFILE
// per-phase clock64 totals of CTA 0 (tools/micro/qr_bench.cu built with -DQR_PROFILE)
void qr_profile_dump(int steps) {
    long long h[16];
    cudaMemcpyFromSymbol(h, g_qr_prof, sizeof h);
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_qr_prof, z, sizeof z);
    printf("qr phases, cycles per column step:");
    for (int i = 0; i < 12; ++i) printf(" [%d] %.0f", i, (double)h[i] / steps);
    printf("\n");
}
#endif

}  // namespace ctmb
