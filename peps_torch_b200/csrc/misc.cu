// Small batched helper kernels of the projector / normalisation steps.
#include "common.h"
#include "cx.h"

namespace ctmb {

// ---------------------------------------------------------------------------------------------
// sort singular values (or eigenvalues by magnitude) descending and gather the leading columns
// ---------------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256) sortcols_kernel(PtrBatch Gb, PtrBatch Wb, PtrBatch sigb, PtrBatch Ssb,
                                                       PtrBatch Uhb, PtrBatch Wsb, int k, int ncol, int eig_mode) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const T* G = reinterpret_cast<const T*>(Gb.p[blockIdx.x]);
    const T* W = reinterpret_cast<const T*>(Wb.p[blockIdx.x]);
    const double* sig = reinterpret_cast<const double*>(sigb.p[blockIdx.x]);
    double* Ss = reinterpret_cast<double*>(Ssb.p[blockIdx.x]);
    T* Uh = reinterpret_cast<T*>(Uhb.p[blockIdx.x]);
    T* Ws = reinterpret_cast<T*>(Wsb.p[blockIdx.x]);
    extern __shared__ int perm[];   // perm[rank] = source column
    const int tid = threadIdx.x;
    // rank by counting: stable for ties (lower index first)
    for (int j = tid; j < k; j += blockDim.x) {
        const double vj = fabs(sig[j]);
        int rank = 0;
        for (int i = 0; i < k; ++i) {
            const double vi = fabs(sig[i]);
            rank += (vi > vj) || (vi == vj && i < j);
        }
        perm[rank] = j;
        Ss[rank] = eig_mode ? sig[j] : vj;
    }
    __syncthreads();
    for (int e = tid; e < k * ncol; e += blockDim.x) {
        const int c = e / k, r = e % k;
        const int src = perm[c];
        if (Ws != nullptr) Ws[e] = W[(size_t)src * k + r];
        if (Uh != nullptr) Uh[e] = G[(size_t)src * k + r];      // columns arrive normalised from the Jacobi kernel
    }
}

void sortcols_launch(const PtrBatch& G, const PtrBatch& W, const PtrBatch& sig, const PtrBatch& Ssorted,
                     const PtrBatch& Uhs, const PtrBatch& Ws, int nb, int k, int ncol, bool cplx, int eig_mode,
                     cudaStream_t stream) {
    size_t smem = (size_t)k * sizeof(int);
    if (cplx) sortcols_kernel<true><<<nb, 256, smem, stream>>>(G, W, sig, Ssorted, Uhs, Ws, k, ncol, eig_mode);
    else sortcols_kernel<false><<<nb, 256, smem, stream>>>(G, W, sig, Ssorted, Uhs, Ws, k, ncol, eig_mode);
    CTMB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// projector finalisation: multiplet-aware truncation (custom_svd.py:70-95), relative cut-off and
// S^-1/2 (ctm_projectors.py:266-270), phase fixing (svd_gesdd.py:18-26); one CTA per column.
// ---------------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256) proj_finalize_kernel(PtrBatch Ub, PtrBatch Vb, PtrBatch Sb, PtrBatch Soutb,
                                                            ProjFinalizeArgs a) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const int j = blockIdx.x, b = blockIdx.y;
    T* U = reinterpret_cast<T*>(Ub.p[b]);
    T* V = reinterpret_cast<T*>(Vb.p[b]);
    const double* Sv = reinterpret_cast<const double*>(Sb.p[b]);
    double* Sout = reinterpret_cast<double*>(Soutb.p[b]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chi = a.chi;

    __shared__ double sh_scale;
    __shared__ double sh_vscale;
    __shared__ int sh_keep;
    __shared__ long long red_amp[8];
    __shared__ int red_idx[8];
    if (tid == 0) {
        // multiplet rule on |S|
        int chi_new = chi;
        if (a.truncating) {
            auto gap = [&](int i) {
                double g = fabs(Sv[i]);
                if (g < a.abstol) g = 0.0;
                double v = (g - fabs(Sv[i + 1])) / (g + 1.0e-16);
                return v > 1.0 ? 0.0 : v;
            };
            if (gap(chi - 1) < a.eps_multiplet) {
                for (int i = chi - 1; i >= 0; --i)
                    if (gap(i) > a.eps_multiplet) { chi_new = i; break; }
            }
        }
        const bool kept = (j <= chi_new);     // St[chi_new+1:] = 0
        const double sj = kept ? Sv[j] : 0.0;
        // relative cut-off: S_sqrt[:count] = rsqrt(S[mask]) ; the mask is a prefix for sorted S
        const double s0 = (0 <= chi_new) ? fabs(Sv[0]) : 0.0;
        double sc = 0.0;
        if (kept && fabs(sj) / s0 > a.reltol) sc = rsqrt(fabs(sj));
        sh_keep = kept ? 1 : 0;
        sh_scale = a.apply_scale ? sc : (kept ? 1.0 : 0.0);
        // V = M^H U / S is formed only where the projector uses it (S/S0 above 1e-2 x the cut-off)
        sh_vscale = !a.v_div_sigma ? 1.0 : ((kept && fabs(sj) > 1.0e-2 * a.reltol * s0 && fabs(sj) > 0.0) ? 1.0 / fabs(sj) : 0.0);
        if (Sout != nullptr) Sout[j] = sj;
    }
    // first arg-max of int64(|U|*2^40) over the column
    T* uc = U + (size_t)j * a.rowsU;
    long long best = -1; int bidx = 0x7fffffff;
    for (int r = tid; r < a.rowsU; r += blockDim.x) {
        long long amp = (long long)(S::abs(uc[r]) * 1099511627776.0);
        if (amp > best) { best = amp; bidx = r; }     // r increases: keeps the first maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) { red_amp[warp] = best; red_idx[warp] = bidx; }
    __syncthreads();
    best = red_amp[0]; bidx = red_idx[0];
    for (int w = 1; w < 8; ++w)
        if (red_amp[w] > best || (red_amp[w] == best && red_idx[w] < bidx)) { best = red_amp[w]; bidx = red_idx[w]; }
    T ph = S::one();
    if (bidx < a.rowsU) {
        T x = uc[bidx];
        double ax = S::abs(x);
        if (ax > 0.0) ph = S::scale(x, 1.0 / ax);
    }
    __syncthreads();       // everyone has read uc[bidx] before the column is overwritten
    const double sc = sh_scale;
    const T cph = S::conj(ph);
    for (int r = tid; r < a.rowsU; r += blockDim.x) {
        T x = S::mul(uc[r], cph);             // sign-fixed U
        if (a.conj_u) x = S::conj(x);
        uc[r] = S::scale(x, sc);
    }
    if (V != nullptr) {
        T* vc = V + (size_t)j * a.rowsV;
        const double vs = sc * sh_vscale;
        for (int r = tid; r < a.rowsV; r += blockDim.x) vc[r] = S::scale(S::mul(vc[r], cph), vs);
    }
}

void proj_finalize_launch(const PtrBatch& U, const PtrBatch& V, const PtrBatch& S, const PtrBatch& Sout,
                          const ProjFinalizeArgs& a, bool cplx, cudaStream_t stream) {
    dim3 grid(a.chi, a.nb);
    if (cplx) proj_finalize_kernel<true><<<grid, 256, 0, stream>>>(U, V, S, Sout, a);
    else proj_finalize_kernel<false><<<grid, 256, 0, stream>>>(U, V, S, Sout, a);
    CTMB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// convergence check of the range finder: max_j || (M X)_j - s_j Y_j || / |s_0| over the columns with
// |s_j| / |s_0| > reltol  (eigen mode: X = Y = U, s = lambda; SVD mode: X = V, Y = U, s = sigma)
// ---------------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256) resid_kernel(PtrBatch MXb, PtrBatch Yb, PtrBatch Sb, int rows, int k, double reltol,
                                                    unsigned long long* out) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const int j = blockIdx.x, b = blockIdx.y;
    const T* mx = reinterpret_cast<const T*>(MXb.p[b]) + (size_t)j * rows;
    const T* y = reinterpret_cast<const T*>(Yb.p[b]) + (size_t)j * rows;
    const double* Sv = reinterpret_cast<const double*>(Sb.p[b]);
    const double s0 = fabs(Sv[0]), sj = Sv[j];
    // out[2]: the Ritz value of index k (edge of the Chebyshev filter of the Hermitian branch, move.cu: cheb_edge_index)
    if (j == 0 && threadIdx.x == 0) atomicMax(out + 2, (unsigned long long)__double_as_longlong(fabs(Sv[k])));
    if (!(fabs(sj) > reltol * s0) || s0 == 0.0) return;
    // out[1]: dynamic range S_0 / S_j of the kept triplets (decides how many operator applications the range finder may
    // chain between two orthogonalisations, move.cu)
    if (threadIdx.x == 0) atomicMax(out + 1, (unsigned long long)__double_as_longlong(s0 / fabs(sj)));
    double acc = 0.0;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) acc += S::abs2(S::sub(mx[r], S::scale(y[r], sj)));
    acc = warp_sum(acc);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) acc += red[w];
        atomicMax(out, (unsigned long long)__double_as_longlong(sqrt(acc) / s0));
    }
}

void resid_launch(const PtrBatch& MX, const PtrBatch& Y, const PtrBatch& S, int nb, int rows, int chi, int k, double reltol,
                  unsigned long long* out, bool cplx, cudaStream_t stream) {
    dim3 grid(chi, nb);
    if (cplx) resid_kernel<true><<<grid, 256, 0, stream>>>(MX, Y, S, rows, k, reltol, out);
    else resid_kernel<false><<<grid, 256, 0, stream>>>(MX, Y, S, rows, k, reltol, out);
    CTMB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// normalisation by the infinity norm (ctmrg.py:210-230)
// ---------------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256) absmax_kernel(ScaleBatch b) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const T* x = reinterpret_cast<const T*>(b.p[blockIdx.y]);
    const long long n = b.count[blockIdx.y];
    double m = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmax(m, S::abs(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
        atomicMax(const_cast<unsigned long long*>(b.amax[blockIdx.y]), (unsigned long long)__double_as_longlong(m));
    }
}

// sum |x|^2 into the slot (a double, zeroed by the caller): the 2-norm of ctm_absorb_normalization != 'inf'
// (ctmrg.py:212-214).  Block partials are combined with atomicAdd: the last bits depend on the block order.
template <bool CPLX>
__global__ void __launch_bounds__(256) sumsq_kernel(ScaleBatch b) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const T* x = reinterpret_cast<const T*>(b.p[blockIdx.y]);
    const long long n = b.count[blockIdx.y];
    double m = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m += S::abs2(x[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m += red[w];
        atomicAdd(reinterpret_cast<double*>(const_cast<unsigned long long*>(b.amax[blockIdx.y])), m);
    }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) scale_kernel(ScaleBatch b, int sqrt_mode) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    T* x = reinterpret_cast<T*>(b.p[blockIdx.y]);
    const long long n = b.count[blockIdx.y];
    double m = __longlong_as_double((long long)*b.amax[blockIdx.y]);
    if (sqrt_mode) m = sqrt(m);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = CPLX ? S::make(S::re(x[i]) / m, S::im(x[i]) / m) : S::make(S::re(x[i]) / m, 0.0);
}

static int blocks_for(const ScaleBatch& b, int nb) {
    long long mx = 1;
    for (int i = 0; i < nb; ++i) mx = std::max(mx, b.count[i]);
    long long g = (mx + 256 * 8 - 1) / (256 * 8);
    return (int)std::max(1ll, std::min(g, 1184ll));
}

void absmax_launch(const ScaleBatch& b, int nb, bool cplx, cudaStream_t stream) {
    dim3 grid(blocks_for(b, nb), nb);
    if (cplx) absmax_kernel<true><<<grid, 256, 0, stream>>>(b);
    else absmax_kernel<false><<<grid, 256, 0, stream>>>(b);
    CTMB_CUDA(cudaGetLastError());
}

void sumsq_launch(const ScaleBatch& b, int nb, bool cplx, cudaStream_t stream) {
    dim3 grid(blocks_for(b, nb), nb);
    if (cplx) sumsq_kernel<true><<<grid, 256, 0, stream>>>(b);
    else sumsq_kernel<false><<<grid, 256, 0, stream>>>(b);
    CTMB_CUDA(cudaGetLastError());
}

void scale_by_amax_launch(const ScaleBatch& b, int nb, bool cplx, cudaStream_t stream, bool sqrt_mode) {
    dim3 grid(blocks_for(b, nb), nb);
    if (cplx) scale_kernel<true><<<grid, 256, 0, stream>>>(b, sqrt_mode ? 1 : 0);
    else scale_kernel<false><<<grid, 256, 0, stream>>>(b, sqrt_mode ? 1 : 0);
    CTMB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// reduced density matrices: _sym_pos_def_matrix (ctm/generic/rdm.py:38-57)
//   out = (raw + raw^H) / 2 / Re tr(raw)                       (always)
//   if min(D) < 0:  out = U max(D,0) U^H / sum max(D,0)        (sym_pos_def=True; D, U = eigenpairs of `out`)
// ---------------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256) rdm_herm_kernel(const void* rawp, void* outp, int n, int normalize) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const T* raw = reinterpret_cast<const T*>(rawp);
    T* out = reinterpret_cast<T*>(outp);
    // every block sums the diagonal in the same order: identical trace everywhere
    double tr = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tr += S::re(raw[(size_t)i * n + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
    __shared__ double red[8];
    __shared__ double sh_tr;
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tr;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += red[w]; sh_tr = t; }
    __syncthreads();
    const double sc = normalize ? 0.5 / sh_tr : 0.5;
    const long long nn = (long long)n * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nn; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / n), j = (int)(e % n);
        out[e] = S::scale(S::add(raw[e], S::conj(raw[(size_t)j * n + i])), sc);
    }
}

__global__ void conj_inplace_kernel(double2* x, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i].y = -x[i].y;
}
void conj_inplace_launch(void* x, long long count, cudaStream_t stream) {
    int grid = (int)std::max(1ll, std::min((count + 255) / 256, 1184ll));
    conj_inplace_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<double2*>(x), count);
    CTMB_CUDA(cudaGetLastError());
}

void rdm_herm_launch(const void* raw, void* out, int n, int normalize, bool cplx, cudaStream_t stream) {
    long long nn = (long long)n * n;
    int grid = (int)std::max(1ll, std::min((nn + 255) / 256, 1184ll));
    if (cplx) rdm_herm_kernel<true><<<grid, 256, 0, stream>>>(raw, out, n, normalize);
    else rdm_herm_kernel<false><<<grid, 256, 0, stream>>>(raw, out, n, normalize);
    CTMB_CUDA(cudaGetLastError());
}

// U: n x n column-major (eigenvector k = U[k*n ..]), D: n eigenvalues in any order
template <bool CPLX>
__global__ void __launch_bounds__(256) rdm_posdef_kernel(void* outp, const void* Up, const double* D, int n) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    T* out = reinterpret_cast<T*>(outp);
    const T* U = reinterpret_cast<const T*>(Up);
    double dmin = 0.0, dsum = 0.0;
    for (int k = 0; k < n; ++k) { dmin = fmin(dmin, D[k]); dsum += fmax(D[k], 0.0); }
    if (!(dmin < 0.0)) return;                        // rdm.py:49: the matrix is replaced only if an eigenvalue is negative
    const long long nn = (long long)n * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nn; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / n), j = (int)(e % n);
        T acc = S::zero();
        for (int k = 0; k < n; ++k) {
            const double d = fmax(D[k], 0.0);
            if (d > 0.0) acc = S::add(acc, S::scale(S::mul(U[(size_t)k * n + i], S::conj(U[(size_t)k * n + j])), d));
        }
        out[e] = S::scale(acc, 1.0 / dsum);
    }
}

void rdm_posdef_launch(void* out, const void* U, const double* D, int n, bool cplx, cudaStream_t stream) {
    long long nn = (long long)n * n;
    int grid = (int)std::max(1ll, std::min((nn + 255) / 256, 1184ll));
    if (cplx) rdm_posdef_kernel<true><<<grid, 256, 0, stream>>>(out, U, D, n);
    else rdm_posdef_kernel<false><<<grid, 256, 0, stream>>>(out, U, D, n);
    CTMB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// C4v epilogue (ctmrg_c4v.py:374,446,182-197): nT <- (nT + conj(nT)^T01)/2 with |.|max, C' = diag(D)/|D0|
// ---------------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256) c4v_sym_kernel(const void* tin, void* tout, int chi, int d,
                                                      unsigned long long* amax) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const T* t = reinterpret_cast<const T*>(tin);
    T* o = reinterpret_cast<T*>(tout);
    const long long n = (long long)chi * chi * d;
    double m = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i % d);
        const long long xy = i / d;
        const int y = (int)(xy % chi), x = (int)(xy / chi);
        T v = S::scale(S::add(t[i], S::conj(t[((long long)y * chi + x) * d + r])), 0.5);
        o[i] = v;
        m = fmax(m, S::abs(v));
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o2));
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
        atomicMax(amax, (unsigned long long)__double_as_longlong(m));
    }
}

template <bool CPLX>
__global__ void c4v_diag_kernel(const double* D, void* cout, int chi) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    T* c = reinterpret_cast<T*>(cout);
    const double d0 = fabs(D[0]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chi * chi; i += gridDim.x * blockDim.x) {
        const int r = i / chi, q = i % chi;
        c[i] = (r == q) ? S::make(D[r] / d0, 0.0) : S::zero();
    }
}

void c4v_sym_launch(const void* tin, void* tout, int chi, int d, unsigned long long* amax, bool cplx,
                    cudaStream_t stream) {
    long long n = (long long)chi * chi * d;
    int grid = (int)std::max(1ll, std::min((n + 255) / 256, 1184ll));
    if (cplx) c4v_sym_kernel<true><<<grid, 256, 0, stream>>>(tin, tout, chi, d, amax);
    else c4v_sym_kernel<false><<<grid, 256, 0, stream>>>(tin, tout, chi, d, amax);
    CTMB_CUDA(cudaGetLastError());
}

void c4v_diag_launch(const double* D, void* cout, int chi, bool cplx, cudaStream_t stream) {
    int grid = std::max(1, std::min((chi * chi + 255) / 256, 1184));
    if (cplx) c4v_diag_kernel<true><<<grid, 256, 0, stream>>>(D, cout, chi);
    else c4v_diag_kernel<false><<<grid, 256, 0, stream>>>(D, cout, chi);
    CTMB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// deterministic standard-normal fill (counter based: splitmix64 + Box-Muller)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void gaussian_kernel(double* out, long long n, unsigned long long seed) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long a = splitmix64(seed ^ (2ull * i));
        unsigned long long b = splitmix64(seed ^ (2ull * i + 1ull));
        double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0);   // (0,1)
        double u2 = (b >> 11) * (1.0 / 9007199254740992.0);           // [0,1)
        out[i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
}

// x[i] += amp * (*scale) * N(0,1)  (tests: unstructured one-ulp noise on an intermediate, see ctmb_debug_set_m_noise)
__global__ void add_noise_kernel(double* x, long long n, double amp, const unsigned long long* scale, unsigned long long seed) {
    const double sc = amp * __longlong_as_double((long long)*scale);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long a = splitmix64(seed ^ (2ull * i));
        unsigned long long b = splitmix64(seed ^ (2ull * i + 1ull));
        double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        double u2 = (b >> 11) * (1.0 / 9007199254740992.0);
        x[i] += sc * sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
}

void add_noise_launch(double* x, long long count, double amp, const unsigned long long* scale, unsigned long long seed,
                      cudaStream_t stream) {
    int grid = (int)std::max(1ll, std::min((count + 255) / 256, 2368ll));
    add_noise_kernel<<<grid, 256, 0, stream>>>(x, count, amp, scale, seed);
    CTMB_CUDA(cudaGetLastError());
}

// y[i] += sign * sqrt(*sumsq) * (x ? x[i] : 1): the spectral shift mu = ||A||_F of the preconditioned Hermitian Jacobi (move.cu),
// with mu living on the device (the slot holds sum |a|^2 from sumsq_launch)
__global__ void shift_axpy_kernel(double* y, const double* x, const unsigned long long* sumsq, double sign, long long n) {
    const double mu = sign * sqrt(__longlong_as_double((long long)*sumsq));
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] += x ? mu * x[i] : mu;
}

void shift_axpy_launch(double* y, const double* x, const unsigned long long* sumsq, double sign, long long count, cudaStream_t stream) {
    int grid = (int)std::max(1ll, std::min((count + 255) / 256, 2368ll));
    shift_axpy_kernel<<<grid, 256, 0, stream>>>(y, x, sumsq, sign, count);
    CTMB_CUDA(cudaGetLastError());
}

// out = c1 * x + c2 * y over `count` doubles (three-term recurrence of the Chebyshev filter, move.cu); out may alias x or y
__global__ void axpby_kernel(double* out, const double* x, double c1, const double* y, double c2, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = c1 * x[i] + (y ? c2 * y[i] : 0.0);
}

void axpby_launch(double* out, const double* x, double c1, const double* y, double c2, long long count, cudaStream_t stream) {
    int grid = (int)std::max(1ll, std::min((count + 255) / 256, 2368ll));
    axpby_kernel<<<grid, 256, 0, stream>>>(out, x, c1, y, c2, count);
    CTMB_CUDA(cudaGetLastError());
}

void fill_gaussian_launch(double* out, long long count, unsigned long long seed, cudaStream_t stream) {
    int grid = (int)std::max(1ll, std::min((count + 255) / 256, 2368ll));
    gaussian_kernel<<<grid, 256, 0, stream>>>(out, count, seed);
    CTMB_CUDA(cudaGetLastError());
}

}  // namespace ctmb
