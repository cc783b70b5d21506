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

void qr_launch(const PtrBatch& A, const PtrBatch& Rout, int nb, int rows, int cols, int ld,
               bool cplx, cudaStream_t stream) {
    CTMB_CHECK(nb >= 1 && nb <= TC_MAX_BATCH, "bad batch");
    CTMB_CHECK(rows >= cols, "qr: expects rows >= cols");
    size_t smem = (size_t)cols * (cplx ? 16 : 8);
    if (cplx) qr_kernel<true><<<nb, QR_THREADS, smem, stream>>>(A, Rout, rows, cols, ld);
    else qr_kernel<false><<<nb, QR_THREADS, smem, stream>>>(A, Rout, rows, cols, ld);
    CTMB_CUDA(cudaGetLastError());
}

}  // namespace ctmb
