// One-sided (Hestenes) Jacobi SVD of small k x k matrices, one CTA per matrix, the matrix and
// the accumulated rotations resident in shared memory (k <= ~118 real) or, above that, in
// global memory (L2).  This is the "small SVD" at the end of the randomised projector SVD that
// replaces the reference's full LAPACK gesdd (linalg/svd_gesdd.py:77-96) and, applied to the
// symmetric Rayleigh-Ritz matrix, the full eigh of the C4v path (linalg/eig_sym.py:14-34).
//
//   G_in * W = G_out,  columns of G_out mutually orthogonal,  sigma_j = ||G_out[:,j]||
//
// Jacobi is used because it delivers the small singular values of the graded triangular factor
// to high *relative* accuracy - the projectors need S^-1/2 down to S/S0 = 1e-8
// (ctm/generic/ctm_projectors.py:266-270).  Pairs of a round-robin tournament round are
// independent: each is rotated by a group of GS lanes using warp shuffles for the three dot
// products; rounds are separated by one __syncthreads().
#include "common.h"
#include "cx.h"

namespace ctmb {

constexpr int JAC_THREADS = 1024;

template <bool CPLX>
__global__ void __launch_bounds__(JAC_THREADS) jacobi_kernel(PtrBatch Gb, PtrBatch Wb, PtrBatch Sb, int k,
                                                              int gs, int use_smem, int max_sweeps, int shift) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    extern __shared__ __align__(16) unsigned char jac_smem[];
    T* Gg = reinterpret_cast<T*>(Gb.p[blockIdx.x]);
    T* Wg = reinterpret_cast<T*>(Wb.p[blockIdx.x]);
    double* sig = reinterpret_cast<double*>(Sb.p[blockIdx.x]);
    const int tid = threadIdx.x;
    T* G = Gg;
    T* W = Wg;
    if (use_smem) {
        G = reinterpret_cast<T*>(jac_smem);
        W = G + (size_t)k * k;
        for (int e = tid; e < k * k; e += JAC_THREADS) G[e] = Gg[e];
    }
    for (int e = tid; e < k * k; e += JAC_THREADS) W[e] = (e / k == e % k) ? S::one() : S::zero();
    __shared__ int rotated;
    __shared__ unsigned long long maxcos;     // bits of the largest |cos(angle)| rotated in this sweep
    if (tid == 0) { rotated = 0; maxcos = 0ull; }
    __syncthreads();

    const int kk = (k + 1) & ~1;          // even number of players (kk-1 >= k means a bye)
    const int npairs = kk / 2;
    const int ngroups = JAC_THREADS / gs;
    const int grp = tid / gs, gl = tid % gs;
    // lanes of one group shuffle among themselves only (groups of a warp may diverge)
    const unsigned gmask = (gs == 32) ? 0xffffffffu : (((1u << gs) - 1u) << ((tid & 31) & ~(gs - 1)));
    const double tol = 2.2e-16 * sqrt((double)k);   // as LAPACK xGESVJ: sqrt(m)*eps
    // optional spectral shift (eigen mode): G <- G + mu*I with mu = ||G||_F >= ||G||_2 makes a
    // Hermitian G positive semi-definite, so that its SVD is its eigen-decomposition
    // (lambda_j = sigma_j - mu, eigenvectors = W) without the +-lambda ambiguity.
    double mu = 0.0;
    if (shift) {
        __shared__ double redm[JAC_THREADS / 32];
        double a = 0.0;
        for (int e = tid; e < k * k; e += JAC_THREADS) a += S::abs2(G[e]);
        a = warp_sum(a);
        if ((tid & 31) == 0) redm[tid >> 5] = a;
        __syncthreads();
        a = 0.0;
        for (int w = 0; w < JAC_THREADS / 32; ++w) a += redm[w];
        mu = sqrt(a);
        __syncthreads();
        for (int e = tid; e < k; e += JAC_THREADS) G[(size_t)e * k + e] = S::add(G[(size_t)e * k + e], S::make(mu, 0.0));
        __syncthreads();
    }

    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        for (int round = 0; round < kk - 1; ++round) {
            for (int pi = grp; pi < npairs; pi += ngroups) {
                int p, q;
                if (pi == 0) { p = kk - 1; q = round; }
                else { p = (round + pi) % (kk - 1); q = (round - pi + (kk - 1)) % (kk - 1); }
                if (p >= k || q >= k) continue;      // bye (uniform within the group)
                if (p > q) { int t = p; p = q; q = t; }
                T* gp = G + (size_t)p * k;
                T* gq = G + (size_t)q * k;
                double a = 0.0, b = 0.0;
                T g = S::zero();
                for (int r = gl; r < k; r += gs) {
                    T x = gp[r], y = gq[r];
                    a += S::abs2(x); b += S::abs2(y);
                    g = S::fma(S::conj(x), y, g);
                }
                for (int o = gs >> 1; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(gmask, a, o);
                    b += __shfl_xor_sync(gmask, b, o);
                    g = S::add(g, S::shfl_xor_m(gmask, g, o));
                }
                const double ag = S::abs(g);
                const double lim = tol * sqrt(a) * sqrt(b);
                if (ag > lim && ag > 0.0) {
                    if (gl == 0) { rotated = 1; atomicMax(&maxcos, (unsigned long long)__double_as_longlong(ag / (sqrt(a) * sqrt(b)))); }
                    const double zeta = (b - a) / (2.0 * ag);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t);
                    const double s = c * t;
                    // y2 = y * conj(phase), phase = g/|g|  (so that x^H y2 = |g| is real)
                    const T cph = S::scale(S::conj(g), 1.0 / ag);
                    T* wp = W + (size_t)p * k;
                    T* wq = W + (size_t)q * k;
                    for (int r = gl; r < k; r += gs) {
                        T x = gp[r], y = S::mul(gq[r], cph);
                        gp[r] = S::sub(S::scale(x, c), S::scale(y, s));
                        gq[r] = S::add(S::scale(x, s), S::scale(y, c));
                        T u = wp[r], v = S::mul(wq[r], cph);
                        wp[r] = S::sub(S::scale(u, c), S::scale(v, s));
                        wq[r] = S::add(S::scale(u, s), S::scale(v, c));
                    }
                }
            }
            __syncthreads();
        }
        const int any = rotated;
        const double mc = __longlong_as_double((long long)maxcos);
        __syncthreads();
        if (tid == 0) { rotated = 0; maxcos = 0ull; }
        __syncthreads();
        // cyclic Jacobi converges quadratically: if the largest cosine met in this sweep was below
        // 1e-8 the columns are now orthogonal to ~1e-16 and the verification sweep can be skipped
        if (!any || mc < 1.0e-8) break;
    }
    // column norms, write back
    for (int c = grp; c < k; c += ngroups) {
        double a = 0.0;
        for (int r = gl; r < k; r += gs) a += S::abs2(G[(size_t)c * k + r]);
        for (int o = gs >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(gmask, a, o);
        if (gl == 0) sig[c] = sqrt(a) - mu;
    }
    if (use_smem) {
        __syncthreads();
        for (int e = tid; e < k * k; e += JAC_THREADS) { Gg[e] = G[e]; Wg[e] = W[e]; }
    }
}

static size_t g_jac_smem_limit = 0;
size_t jacobi_smem_limit() {
    if (g_jac_smem_limit == 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        g_jac_smem_limit = (size_t)v > 2048 ? (size_t)v - 1024 : 0;
    }
    return g_jac_smem_limit;
}

void jacobi_launch(const PtrBatch& G, const PtrBatch& W, const PtrBatch& sig, int nb, int k, bool cplx,
                   int max_sweeps, int shift, cudaStream_t stream) {
    CTMB_CHECK(nb >= 1 && nb <= TC_MAX_BATCH, "bad batch");
    const size_t need = 2 * (size_t)k * k * (cplx ? 16 : 8);
    const int use_smem = need <= jacobi_smem_limit();
    const size_t smem = use_smem ? need : 0;
    const int npairs = (k + 1) / 2;
    int gs = 32;
    while (gs > 8 && npairs * gs > JAC_THREADS) gs >>= 1;
    if (cplx) {
        auto kern = jacobi_kernel<true>;
        static size_t set = 0;
        if (smem > set) { CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jacobi_smem_limit())); set = jacobi_smem_limit(); }
        kern<<<nb, JAC_THREADS, smem, stream>>>(G, W, sig, k, gs, use_smem, max_sweeps, shift);
    } else {
        auto kern = jacobi_kernel<false>;
        static size_t set = 0;
        if (smem > set) { CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jacobi_smem_limit())); set = jacobi_smem_limit(); }
        kern<<<nb, JAC_THREADS, smem, stream>>>(G, W, sig, k, gs, use_smem, max_sweeps, shift);
    }
    CTMB_CUDA(cudaGetLastError());
}

}  // namespace ctmb
