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
#include <cooperative_groups.h>
#include "cx.h"

namespace ctmb {

// When may the sweep loop stop without a verification sweep?  Cyclic Jacobi converges quadratically once every rotation
// of a sweep was by a small ANGLE: sin^2(theta) = |g|^2 r^2 rc^2 <= 1e-16.  (Round 1 tested the cosine between the two
// columns, |g|^2 <= 1e-16 a b; inside a tight cluster of singular values, a ~ b, a tiny cosine can still mean a large
// rotation angle, the quadratic argument does not apply and the cluster was left orthogonal to ~1e-11 only.)  A pair
// whose cosine is already below 2 tol (tol = eps sqrt(k), the threshold under which a pair is not rotated at all) is rotated
// but does not ask for another sweep either, whatever its angle (with 1e-13 here, a large-angle rotation inside a cluster
// late in the last sweep left the pair's partners of earlier rounds at cosines of a few 1e-13: test_clustered_spectrum_*): a rotation inside the pair preserves the magnitudes of its inner products with every other column, so exactly
// degenerate multiplets (angle 45 degrees at rounding-level cosines) cannot keep the loop alive.
// diagnostics: total sweeps executed / matrices processed since the last read (tools/ only)
__device__ unsigned long long g_jac_stats[2];

// Template parameters: SMEM = matrix (and W) resident in shared memory (compile-time, so that the column
// accesses are LDS/STS and not generic loads); GS = lanes per pair; THREADS = block size.  The first version
// ran 1024 threads (64 registers per thread): the register-cached columns spilled to local memory and 40 % of
// the stall samples sat on STL (profiles/r1_c2_qr_jacobi.md).  Now: real 512 threads x GS 8 (16 rows per lane,
// four pairs share the scalar rotation arithmetic of one warp instruction), complex 512 x GS 16.
template <bool CPLX, bool SMEM, int GS, int THREADS>
__global__ void __launch_bounds__(THREADS) jacobi_kernel(PtrBatch Gb, PtrBatch Wb, PtrBatch Sb, int k,
                                                          int max_sweeps, int shift, int transpose_in) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    constexpr int JAC_THREADS = THREADS;
    constexpr int gs = GS;
    constexpr bool use_smem = SMEM;
    extern __shared__ __align__(16) unsigned char jac_smem[];
    T* Gg = reinterpret_cast<T*>(Gb.p[blockIdx.x]);
    T* Wg = reinterpret_cast<T*>(Wb.p[blockIdx.x]);      // nullptr: right rotations are not accumulated
    double* sig = reinterpret_cast<double*>(Sb.p[blockIdx.x]);
    const int tid = threadIdx.x;
    const bool accw = Wg != nullptr;
    T* G = use_smem ? reinterpret_cast<T*>(jac_smem) : Gg;
    T* W = use_smem ? reinterpret_cast<T*>(jac_smem) + (size_t)k * k : Wg;
    if (use_smem) {
        if (transpose_in) for (int e = tid; e < k * k; e += JAC_THREADS) G[e] = S::conj(Gg[(size_t)(e % k) * k + e / k]);
        else for (int e = tid; e < k * k; e += JAC_THREADS) G[e] = Gg[e];
    } else if (transpose_in) {
        // in-place conjugate transpose in global memory
        for (int e = tid; e < k * k; e += JAC_THREADS) {
            const int c = e / k, r = e % k;
            if (r > c) { T x = Gg[e], y = Gg[(size_t)r * k + c]; Gg[e] = S::conj(y); Gg[(size_t)r * k + c] = S::conj(x); }
            else if (r == c) Gg[e] = S::conj(Gg[e]);
        }
    }
    if (accw) for (int e = tid; e < k * k; e += JAC_THREADS) W[e] = (e / k == e % k) ? S::one() : S::zero();
    __shared__ int rotated;
    __shared__ int notsmall;                  // a rotation with |cos(angle)| > 1e-8 happened in this sweep
    if (tid == 0) { rotated = 0; notsmall = 0; }
    __syncthreads();

    const int kk = (k + 1) & ~1;          // even number of players (kk-1 >= k means a bye)
    const int npairs = kk / 2;
    const int ngroups = JAC_THREADS / gs;
    const int grp = tid / gs, gl = tid % gs;
    // lanes of one group shuffle among themselves only (groups of a warp may diverge)
    const unsigned gmask = (gs == 32) ? 0xffffffffu : (((1u << gs) - 1u) << ((tid & 31) & ~(gs - 1)));
    const double tol = 2.2e-16 * sqrt((double)k);   // as LAPACK xGESVJ: sqrt(m)*eps
    const double tol2 = tol * tol;
    // optional spectral shift (eigen mode): G <- G + mu*I with mu = ||G||_F >= ||G||_2 makes a
    // Hermitian G positive semi-definite, so that its SVD is its eigen-decomposition
    // (lambda_j = sigma_j - mu, eigenvectors = normalised columns) without the +-lambda ambiguity.
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
    constexpr int NR = 128 / GS;          // register-cached rows per lane (k <= NR*gs = 128)
    const bool cached = k <= NR * gs;
    const bool vec2 = cached && (k & 1) == 0;   // even k: columns are 16-byte aligned

    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        for (int round = 0; round < kk - 1; ++round) {
            for (int pi = grp; pi < npairs; pi += ngroups) {
                int p, q;
                if (pi == 0) { p = kk - 1; q = round; }
                else { p = (round + pi) % (kk - 1); q = (round - pi + (kk - 1)) % (kk - 1); }
                if (p >= k || q >= k) continue;      // bye (uniform within the group)
                if (p > q) { int t = p; p = q; q = t; }
                if constexpr (!CPLX && SMEM) {
                    if (vec2) {
                        // 128-bit shared-memory accesses: lane owns the row PAIRS v = gl, gl+gs, ... (rows 2v, 2v+1).
                        // Halves the LDS/STS instruction count of a kernel whose rounds are bound by the
                        // shared-memory pipe (mio_throttle + barrier stalls, profiles/r1_c2_qr_jacobi.md).
                        constexpr int NV = NR / 2;
                        double2* gp2 = reinterpret_cast<double2*>(G + (size_t)p * k);
                        double2* gq2 = reinterpret_cast<double2*>(G + (size_t)q * k);
                        const int kv = k >> 1;
                        double2 xv[NV], yv[NV];
                        double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            const int v = gl + i * gs;
                            xv[i] = v < kv ? gp2[v] : make_double2(0.0, 0.0);
                            yv[i] = v < kv ? gq2[v] : make_double2(0.0, 0.0);
                            a = fma(xv[i].x, xv[i].x, fma(xv[i].y, xv[i].y, a));
                            b = fma(yv[i].x, yv[i].x, fma(yv[i].y, yv[i].y, b));
                            g = fma(xv[i].x, yv[i].x, fma(xv[i].y, yv[i].y, g));
                        }
                        for (int o = gs >> 1; o > 0; o >>= 1) {
                            a += __shfl_xor_sync(gmask, a, o);
                            b += __shfl_xor_sync(gmask, b, o);
                            g += __shfl_xor_sync(gmask, g, o);
                        }
                        const double ag2 = g * g;
                        if (ag2 > tol2 * a * b && ag2 > 0.0) {
                            const double d = b - a;
                            const double r = rsqrt(d * d + 4.0 * ag2);
                            const double c2 = 0.5 + 0.5 * fabs(d) * r;
                            const double rc = rsqrt(c2);
                            if (gl == 0) { rotated = 1; if (ag2 * r * r * rc * rc > 1.0e-16 && ag2 > 4.0 * tol2 * a * b) notsmall = 1; }
                            const double c = c2 * rc;
                            const double al = g * copysign(r * rc, d);
#pragma unroll
                            for (int i = 0; i < NV; ++i) {
                                const int v = gl + i * gs;
                                if (v < kv) {
                                    const double2 x = xv[i], y = yv[i];
                                    gp2[v] = make_double2(x.x * c - al * y.x, x.y * c - al * y.y);
                                    gq2[v] = make_double2(al * x.x + y.x * c, al * x.y + y.y * c);
                                }
                            }
                            if (accw) {
                                double2* wp2 = reinterpret_cast<double2*>(W + (size_t)p * k);
                                double2* wq2 = reinterpret_cast<double2*>(W + (size_t)q * k);
#pragma unroll
                                for (int i = 0; i < NV; ++i) {
                                    const int v = gl + i * gs;
                                    if (v < kv) {
                                        const double2 u = wp2[v], w = wq2[v];
                                        wp2[v] = make_double2(u.x * c - al * w.x, u.y * c - al * w.y);
                                        wq2[v] = make_double2(al * u.x + w.x * c, al * u.y + w.y * c);
                                    }
                                }
                            }
                        }
                        continue;
                    }
                }
                T* gp = G + (size_t)p * k;
                T* gq = G + (size_t)q * k;
                double a = 0.0, b = 0.0;
                T g = S::zero();
                T xc[NR], yc[NR];
                if (cached) {
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const int r = gl + i * gs;
                        xc[i] = r < k ? gp[r] : S::zero();
                        yc[i] = r < k ? gq[r] : S::zero();
                        a += S::abs2(xc[i]); b += S::abs2(yc[i]);
                        g = S::fma(S::conj(xc[i]), yc[i], g);
                    }
                } else {
                    for (int r = gl; r < k; r += gs) {
                        T x = gp[r], y = gq[r];
                        a += S::abs2(x); b += S::abs2(y);
                        g = S::fma(S::conj(x), y, g);
                    }
                }
                for (int o = gs >> 1; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(gmask, a, o);
                    b += __shfl_xor_sync(gmask, b, o);
                    g = S::add(g, S::shfl_xor_m(gmask, g, o));
                }
                const double ag2 = S::abs2(g);
                if (ag2 > tol2 * a * b && ag2 > 0.0) {
                    // Rotation [x y] <- [x y] [[c, conj(al)], [-al, c]] that makes the columns orthogonal:
                    //   tan(2 theta) = 2|g| / |b-a|,  c = cos(theta) >= 1/sqrt(2),  al = sign(b-a) conj(g) sin(theta)/|g|.
                    // Written with two reciprocal square roots and no division / square root:
                    //   r = 1/sqrt(d^2 + 4|g|^2), cos(2 theta) = |d| r, c^2 = (1 + cos(2 theta))/2, rc = 1/sqrt(c^2),
                    //   c = c^2 rc,  sin(theta)/|g| = r rc.
                    const double d = b - a;
                    const double r = rsqrt(d * d + 4.0 * ag2);
                    const double c2 = 0.5 + 0.5 * fabs(d) * r;
                    const double rc = rsqrt(c2);
                    if (gl == 0) { rotated = 1; if (ag2 * r * r * rc * rc > 1.0e-16 && ag2 > 4.0 * tol2 * a * b) notsmall = 1; }
                    const double c = c2 * rc;
                    const T al = S::scale(S::conj(g), copysign(r * rc, d));
                    const T cal = S::conj(al);
                    if (cached) {
#pragma unroll
                        for (int i = 0; i < NR; ++i) {
                            const int r = gl + i * gs;
                            if (r < k) {
                                const T x = xc[i], y = yc[i];
                                gp[r] = S::sub(S::scale(x, c), S::mul(al, y));
                                gq[r] = S::add(S::mul(cal, x), S::scale(y, c));
                            }
                        }
                    } else {
                        for (int r = gl; r < k; r += gs) {
                            T x = gp[r], y = gq[r];
                            gp[r] = S::sub(S::scale(x, c), S::mul(al, y));
                            gq[r] = S::add(S::mul(cal, x), S::scale(y, c));
                        }
                    }
                    if (accw) {
                        T* wp = W + (size_t)p * k;
                        T* wq = W + (size_t)q * k;
                        for (int r = gl; r < k; r += gs) {
                            T u = wp[r], v = wq[r];
                            wp[r] = S::sub(S::scale(u, c), S::mul(al, v));
                            wq[r] = S::add(S::mul(cal, u), S::scale(v, c));
                        }
                    }
                }
            }
            __syncthreads();
        }
        const int any = rotated;
        const int big = notsmall;
        __syncthreads();
        if (tid == 0) { rotated = 0; notsmall = 0; }
        __syncthreads();
        // cyclic Jacobi converges quadratically: if the largest cosine met in this sweep was below
        // 1e-8 the columns are now orthogonal to ~1e-16 and the verification sweep can be skipped
        if (!any || !big) { if (tid == 0) atomicAdd(&g_jac_stats[0], (unsigned long long)(sweep + 1)); break; }
        if (sweep + 1 == max_sweeps && tid == 0) atomicAdd(&g_jac_stats[0], (unsigned long long)max_sweeps);
    }
    if (tid == 0) atomicAdd(&g_jac_stats[1], 1ull);
    // column norms; the columns are written back NORMALISED (left singular vectors / eigenvectors)
    for (int c = grp; c < k; c += ngroups) {
        double a = 0.0;
        for (int r = gl; r < k; r += gs) a += S::abs2(G[(size_t)c * k + r]);
        for (int o = gs >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(gmask, a, o);
        const double nrm = sqrt(a);
        if (gl == 0) sig[c] = nrm - mu;
        const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
        for (int r = gl; r < k; r += gs) G[(size_t)c * k + r] = S::scale(G[(size_t)c * k + r], inv);
        if (accw) {
            // W is a product of ~sweeps * k plane rotations per column: its column norms drift by a few tens of ulps
            // (1.4e-14 measured at k = 96); renormalise, so that V = Q2 W has unit columns like U = Q Uhat
            double w2 = 0.0;
            for (int r = gl; r < k; r += gs) w2 += S::abs2(W[(size_t)c * k + r]);
            for (int o = gs >> 1; o > 0; o >>= 1) w2 += __shfl_xor_sync(gmask, w2, o);
            const double winv = w2 > 0.0 ? rsqrt(w2) : 0.0;
            for (int r = gl; r < k; r += gs) W[(size_t)c * k + r] = S::scale(W[(size_t)c * k + r], winv);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int e = tid; e < k * k; e += JAC_THREADS) { Gg[e] = G[e]; if (accw) Wg[e] = W[e]; }
    }
}


// ---------------------------------------------------------------------------------------------
// Multi-CTA variant for matrices that do not fit the shared memory of one SM (k = 192 complex at
// D=4 chi=96, k = 512 at D=8 chi=256): G and W stay in global memory (L2-resident), the pairs of a
// round are dealt to the warps of `cpm` CTAs per matrix, and the rounds are separated by a
// per-matrix barrier in global memory (all CTAs are co-resident: cooperative launch).  Loads
// bypass L1 (ld.global.cg): the columns are rewritten by other SMs between rounds.
// ---------------------------------------------------------------------------------------------
constexpr int JC_THREADS = 256;
__device__ unsigned int g_jc_bar[TC_MAX_BATCH];
__device__ int g_jc_flag[TC_MAX_BATCH][3][2];
__device__ double g_jc_mu[TC_MAX_BATCH];

__device__ __forceinline__ double ldcg_t(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ldcg_t(const double2* p) { return __ldcg(p); }

// prologue (one CTA per matrix): optional in-place conjugate transpose, optional spectral shift, W <- I,
// barrier / flag reset
template <bool CPLX>
__global__ void __launch_bounds__(1024) jacobi_prep_kernel(PtrBatch Gb, PtrBatch Wb, int k, int shift, int transpose_in) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    T* G = reinterpret_cast<T*>(Gb.p[blockIdx.x]);
    T* W = reinterpret_cast<T*>(Wb.p[blockIdx.x]);
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) {
        g_jc_bar[blockIdx.x] = 0u;
        for (int i = 0; i < 3; ++i) g_jc_flag[blockIdx.x][i][0] = g_jc_flag[blockIdx.x][i][1] = 0;
        g_jc_mu[blockIdx.x] = 0.0;
    }
    if (transpose_in) {
        for (int e = tid; e < k * k; e += nt) {
            const int c = e / k, r = e % k;
            if (r > c) { T x = G[e], y = G[(size_t)r * k + c]; G[e] = S::conj(y); G[(size_t)r * k + c] = S::conj(x); }
            else if (r == c) G[e] = S::conj(G[e]);
        }
        __syncthreads();
    }
    if (W != nullptr) for (int e = tid; e < k * k; e += nt) W[e] = (e / k == e % k) ? S::one() : S::zero();
    if (shift) {
        __shared__ double redm[32];
        double a = 0.0;
        for (int e = tid; e < k * k; e += nt) a += S::abs2(G[e]);
        a = warp_sum(a);
        if ((tid & 31) == 0) redm[tid >> 5] = a;
        __syncthreads();
        a = 0.0;
        for (int w = 0; w < nt / 32; ++w) a += redm[w];
        const double mu = sqrt(a);
        for (int e = tid; e < k; e += nt) G[(size_t)e * k + e] = S::add(G[(size_t)e * k + e], S::make(mu, 0.0));
        if (tid == 0) g_jc_mu[blockIdx.x] = mu;
    }
}

// NR = register-cached rows per lane (k <= 32 NR): the two columns of a pair are loaded ONCE, with all loads in flight
// together, rotated in registers and stored; W likewise.  (Round 1 walked the columns with a run-time loop, one L2 round
// trip per 32 rows and two passes: 17 us per round at k = 512, 52 ms per 512 x 512 factor -- profiles/r2_c5s_launches.md.)
template <bool CPLX, int NR>
__global__ void __launch_bounds__(JC_THREADS) jacobi_coop_kernel(PtrBatch Gb, PtrBatch Wb, PtrBatch Sb, int k, int cpm,
                                                                   int max_sweeps) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    const int mat = blockIdx.x / cpm, cta = blockIdx.x % cpm;
    T* G = reinterpret_cast<T*>(Gb.p[mat]);
    T* W = reinterpret_cast<T*>(Wb.p[mat]);
    double* sig = reinterpret_cast<double*>(Sb.p[mat]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = JC_THREADS / 32;
    const int gw = cta * NWARP + warp, nw = cpm * NWARP;       // this warp among the warps of the matrix
    const bool accw = W != nullptr;
    const int kk = (k + 1) & ~1, npairs = kk / 2;
    const double tol = 2.2e-16 * sqrt((double)k);
    const double tol2 = tol * tol;
    unsigned int epoch = 0;
    volatile unsigned int* bar = &g_jc_bar[mat];
    auto mat_barrier = [&]() {
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
    };
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        int* flag = g_jc_flag[mat][sweep % 3];
        if (cta == 0 && tid == 0) { int* nxt = g_jc_flag[mat][(sweep + 1) % 3]; nxt[0] = 0; nxt[1] = 0; }
        int rot = 0, big = 0;
        for (int round = 0; round < kk - 1; ++round) {
            for (int pi = gw; pi < npairs; pi += nw) {
                int p, q;
                if (pi == 0) { p = kk - 1; q = round; }
                else { p = (round + pi) % (kk - 1); q = (round - pi + (kk - 1)) % (kk - 1); }
                if (p >= k || q >= k) continue;
                if (p > q) { int t = p; p = q; q = t; }
                T* gp = G + (size_t)p * k;
                T* gq = G + (size_t)q * k;
                double a = 0.0, b = 0.0;
                T g = S::zero();
                T xr[NR], yr[NR];
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int r = lane + 32 * i;
                    xr[i] = r < k ? ldcg_t(gp + r) : S::zero();
                    yr[i] = r < k ? ldcg_t(gq + r) : S::zero();
                }
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    a += S::abs2(xr[i]); b += S::abs2(yr[i]);
                    g = S::fma(S::conj(xr[i]), yr[i], g);
                }
                a = warp_sum(a); b = warp_sum(b); g = warp_sum_t<CPLX>(g);
                const double ag2 = S::abs2(g);
                if (ag2 > tol2 * a * b && ag2 > 0.0) {
                    const double d = b - a;
                    const double r = rsqrt(d * d + 4.0 * ag2);
                    const double c2 = 0.5 + 0.5 * fabs(d) * r;
                    const double rc = rsqrt(c2);
                    rot = 1;
                    if (ag2 * r * r * rc * rc > 1.0e-16 && ag2 > 4.0 * tol2 * a * b) big = 1;
                    const double c = c2 * rc;
                    const T al = S::scale(S::conj(g), copysign(r * rc, d));
                    const T cal = S::conj(al);
                    if (accw) {                             // W columns: loads in flight while G is rotated
                        T* wp = W + (size_t)p * k;
                        T* wq = W + (size_t)q * k;
                        T ur[NR], vr[NR];
#pragma unroll
                        for (int i = 0; i < NR; ++i) {
                            const int rr = lane + 32 * i;
                            ur[i] = rr < k ? ldcg_t(wp + rr) : S::zero();
                            vr[i] = rr < k ? ldcg_t(wq + rr) : S::zero();
                        }
#pragma unroll
                        for (int i = 0; i < NR; ++i) {
                            const int rr = lane + 32 * i;
                            if (rr < k) {
                                gp[rr] = S::sub(S::scale(xr[i], c), S::mul(al, yr[i]));
                                gq[rr] = S::add(S::mul(cal, xr[i]), S::scale(yr[i], c));
                            }
                        }
#pragma unroll
                        for (int i = 0; i < NR; ++i) {
                            const int rr = lane + 32 * i;
                            if (rr < k) {
                                wp[rr] = S::sub(S::scale(ur[i], c), S::mul(al, vr[i]));
                                wq[rr] = S::add(S::mul(cal, ur[i]), S::scale(vr[i], c));
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < NR; ++i) {
                            const int rr = lane + 32 * i;
                            if (rr < k) {
                                gp[rr] = S::sub(S::scale(xr[i], c), S::mul(al, yr[i]));
                                gq[rr] = S::add(S::mul(cal, xr[i]), S::scale(yr[i], c));
                            }
                        }
                    }
                }
            }
            mat_barrier();
        }
        if (lane == 0 && rot) { atomicOr(&flag[0], 1); if (big) atomicOr(&flag[1], 1); }
        mat_barrier();
        const int any = *(volatile int*)&flag[0], anybig = *(volatile int*)&flag[1];
        if (!any || !anybig || sweep + 1 == max_sweeps) {
            if (cta == 0 && tid == 0) { atomicAdd(&g_jac_stats[0], (unsigned long long)(sweep + 1)); atomicAdd(&g_jac_stats[1], 1ull); }
            break;
        }
    }
    // column norms; the columns are written back NORMALISED
    const double mu = g_jc_mu[mat];
    for (int c = gw; c < k; c += nw) {
        double a = 0.0;
        for (int r = lane; r < k; r += 32) a += S::abs2(ldcg_t(G + (size_t)c * k + r));
        a = warp_sum(a);
        const double nrm = sqrt(a);
        if (lane == 0) sig[c] = nrm - mu;
        const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
        for (int r = lane; r < k; r += 32) G[(size_t)c * k + r] = S::scale(ldcg_t(G + (size_t)c * k + r), inv);
        if (accw) {                                     // renormalise the accumulated rotations (see jacobi_kernel)
            double w2 = 0.0;
            for (int r = lane; r < k; r += 32) w2 += S::abs2(ldcg_t(W + (size_t)c * k + r));
            w2 = warp_sum(w2);
            const double winv = w2 > 0.0 ? rsqrt(w2) : 0.0;
            for (int r = lane; r < k; r += 32) W[(size_t)c * k + r] = S::scale(ldcg_t(W + (size_t)c * k + r), winv);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Cluster-resident variant for the sizes in between (one matrix does not fit the shared memory of one SM, but fits the
// DISTRIBUTED shared memory of a thread-block cluster: k = 192 complex, the Rayleigh-Ritz problem of config 3; the full
// decompositions of the AD path): column c lives in the shared memory of CTA c % CL, every warp reads and writes the two
// columns of its pair through DSMEM, rounds are separated by the hardware cluster barrier (~0.25 us) instead of the
// atomic-and-spin barrier in global memory of the cooperative kernel (~2 us per round at 191 rounds x 9 sweeps).
// MEASURED [B200, config 3, k = 192 complex]: correct (same tests pass) but SLOWER than the cooperative kernel -- Jacobi
// 7.9 ms against 5.2 ms per move (24.7 against 21.8 ms per move): a warp's 16 dependent remote loads / stores per lane
// cost more through DSMEM than the in-flight ld.global.cg of the L2-resident version saves on the barrier.  Kept behind
// CTMB_JACOBI_CLUSTER=1 as a recorded experiment (profiles/r2_experiments.md); the cooperative kernel stays the default.
// ---------------------------------------------------------------------------------------------
constexpr int JCL_THREADS = 384;
template <bool CPLX, int NR>
__global__ void __launch_bounds__(JCL_THREADS) jacobi_cluster_kernel(PtrBatch Gb, PtrBatch Wb, PtrBatch Sb, int k, int CL,
                                                                      int max_sweeps) {
    using S = Sc<CPLX>;
    using T = typename S::T;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int mat = blockIdx.x / CL;
    T* G = reinterpret_cast<T*>(Gb.p[mat]);
    T* W = reinterpret_cast<T*>(Wb.p[mat]);
    double* sig = reinterpret_cast<double*>(Sb.p[mat]);
    const bool accw = W != nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = JCL_THREADS / 32;
    const int gw = rank * NWARP + warp, nw = CL * NWARP;
    const int cpc = (k + CL - 1) / CL;                    // column slots per CTA
    extern __shared__ __align__(16) unsigned char jcl_smem[];
    T* Gs = reinterpret_cast<T*>(jcl_smem);                // [cpc][k]
    T* Ws = Gs + (size_t)cpc * k;                          // [cpc][k] (only with accw)
    __shared__ int flags[3][2];                            // (the copy of CTA 0 is the one in use)
    // own columns: global -> shared
    for (int slot = 0; slot < cpc; ++slot) {
        const int c = slot * CL + rank;
        if (c >= k) break;
        for (int r = tid; r < k; r += JCL_THREADS) {
            Gs[(size_t)slot * k + r] = G[(size_t)c * k + r];
            if (accw) Ws[(size_t)slot * k + r] = W[(size_t)c * k + r];
        }
    }
    if (tid < 6) flags[tid / 2][tid % 2] = 0;
    cluster.sync();
    auto colG = [&](int c) { return cluster.map_shared_rank(Gs, c % CL) + (size_t)(c / CL) * k; };
    auto colW = [&](int c) { return cluster.map_shared_rank(Ws, c % CL) + (size_t)(c / CL) * k; };
    int (*flag0)[2] = reinterpret_cast<int (*)[2]>(cluster.map_shared_rank(&flags[0][0], 0));
    const int kk = (k + 1) & ~1, npairs = kk / 2;
    const double tol = 2.2e-16 * sqrt((double)k);
    const double tol2 = tol * tol;
    int sweeps_done = max_sweeps;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        if (rank == 0 && tid == 0) { flags[(sweep + 1) % 3][0] = 0; flags[(sweep + 1) % 3][1] = 0; }
        int rot = 0, big = 0;
        for (int round = 0; round < kk - 1; ++round) {
            for (int pi = gw; pi < npairs; pi += nw) {
                int p, q;
                if (pi == 0) { p = kk - 1; q = round; }
                else { p = (round + pi) % (kk - 1); q = (round - pi + (kk - 1)) % (kk - 1); }
                if (p >= k || q >= k) continue;
                if (p > q) { int t = p; p = q; q = t; }
                T* gp = colG(p);
                T* gq = colG(q);
                double a = 0.0, b = 0.0;
                T g = S::zero();
                T xr[NR], yr[NR];
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int r = lane + 32 * i;
                    xr[i] = r < k ? gp[r] : S::zero();
                    yr[i] = r < k ? gq[r] : S::zero();
                }
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    a += S::abs2(xr[i]); b += S::abs2(yr[i]);
                    g = S::fma(S::conj(xr[i]), yr[i], g);
                }
                a = warp_sum(a); b = warp_sum(b); g = warp_sum_t<CPLX>(g);
                const double ag2 = S::abs2(g);
                if (ag2 > tol2 * a * b && ag2 > 0.0) {
                    const double d = b - a;
                    const double r = rsqrt(d * d + 4.0 * ag2);
                    const double c2 = 0.5 + 0.5 * fabs(d) * r;
                    const double rc = rsqrt(c2);
                    rot = 1;
                    if (ag2 * r * r * rc * rc > 1.0e-16 && ag2 > 4.0 * tol2 * a * b) big = 1;
                    const double c = c2 * rc;
                    const T al = S::scale(S::conj(g), copysign(r * rc, d));
                    const T cal = S::conj(al);
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const int rr = lane + 32 * i;
                        if (rr < k) {
                            gp[rr] = S::sub(S::scale(xr[i], c), S::mul(al, yr[i]));
                            gq[rr] = S::add(S::mul(cal, xr[i]), S::scale(yr[i], c));
                        }
                    }
                    if (accw) {
                        T* wp = colW(p);
                        T* wq = colW(q);
#pragma unroll
                        for (int i = 0; i < NR; ++i) {
                            const int rr = lane + 32 * i;
                            if (rr < k) {
                                const T u = wp[rr], v = wq[rr];
                                wp[rr] = S::sub(S::scale(u, c), S::mul(al, v));
                                wq[rr] = S::add(S::mul(cal, u), S::scale(v, c));
                            }
                        }
                    }
                }
            }
            cluster.sync();
        }
        if (lane == 0 && rot) { atomicOr(&flag0[sweep % 3][0], 1); if (big) atomicOr(&flag0[sweep % 3][1], 1); }
        cluster.sync();
        const int any = *(volatile int*)&flag0[sweep % 3][0], anybig = *(volatile int*)&flag0[sweep % 3][1];
        cluster.sync();                                   // everybody has read the flags before CTA 0 may reuse the slot
        if (!any || !anybig) { sweeps_done = sweep + 1; break; }
    }
    if (rank == 0 && tid == 0) { atomicAdd(&g_jac_stats[0], (unsigned long long)sweeps_done); atomicAdd(&g_jac_stats[1], 1ull); }
    // column norms; own columns are written back NORMALISED
    const double mu = g_jc_mu[mat];
    for (int slot = warp; slot < cpc; slot += NWARP) {
        const int c = slot * CL + rank;
        if (c >= k) continue;
        const T* gc = Gs + (size_t)slot * k;
        double a = 0.0;
        for (int r = lane; r < k; r += 32) a += S::abs2(gc[r]);
        a = warp_sum(a);
        const double nrm = sqrt(a);
        if (lane == 0) sig[c] = nrm - mu;
        const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
        for (int r = lane; r < k; r += 32) G[(size_t)c * k + r] = S::scale(gc[r], inv);
        if (accw) {
            const T* wc = Ws + (size_t)slot * k;
            double w2 = 0.0;
            for (int r = lane; r < k; r += 32) w2 += S::abs2(wc[r]);
            w2 = warp_sum(w2);
            const double winv = w2 > 0.0 ? rsqrt(w2) : 0.0;
            for (int r = lane; r < k; r += 32) W[(size_t)c * k + r] = S::scale(wc[r], winv);
        }
    }
    cluster.sync();                                       // no CTA may exit while peers can still read its shared memory
}

template <bool CPLX, int NR>
static bool jacobi_cluster_run(const PtrBatch& G, const PtrBatch& W, const PtrBatch& sig, int nb, int k, int max_sweeps,
                               cudaStream_t stream) {
    const bool accw = W.p[0] != nullptr;
    const int CL = 8;
    const size_t es = CPLX ? 16 : 8;
    const size_t smem = (size_t)(accw ? 2 : 1) * ((k + CL - 1) / CL) * k * es;
    if (smem > 200 * 1024) return false;
    auto kern = jacobi_cluster_kernel<CPLX, NR>;
    static bool attr_set = false;
    if (!attr_set) {
        CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nb * CL);
    cfg.blockDim = dim3(JCL_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CTMB_CUDA(cudaLaunchKernelEx(&cfg, kern, G, W, sig, k, CL, max_sweeps));
    return true;
}

void jacobi_stats(unsigned long long out[2]) {
    cudaMemcpyFromSymbol(out, g_jac_stats, sizeof(unsigned long long) * 2);
    unsigned long long z[2] = {0, 0};
    cudaMemcpyToSymbol(g_jac_stats, z, sizeof z);
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
                   int max_sweeps, int shift, int transpose_in, cudaStream_t stream) {
    CTMB_CHECK(nb >= 1 && nb <= TC_MAX_BATCH, "bad batch");
    const bool accw = W.p[0] != nullptr;
    const size_t need = (accw ? 2 : 1) * (size_t)k * k * (cplx ? 16 : 8);
    const int use_smem = need <= jacobi_smem_limit();
    const size_t smem = use_smem ? need : 0;
    static int coop_mode = -1;
    if (coop_mode < 0) { const char* ev = getenv("CTMB_JACOBI_COOP"); coop_mode = ev ? atoi(ev) : 1; }
    static int cl_mode = -1;
    if (cl_mode < 0) { const char* ev = getenv("CTMB_JACOBI_CLUSTER"); cl_mode = ev ? atoi(ev) : 0; }   // off: measured slower
    if (!use_smem && cl_mode && k <= 256 && nb * 8 <= 128) {
        // in between: the matrix fits the distributed shared memory of a cluster of eight CTAs
        const size_t need8 = (accw ? 2 : 1) * (size_t)((k + 7) / 8) * k * (cplx ? 16 : 8);
        if (need8 <= 200 * 1024) {
            if (cplx) jacobi_prep_kernel<true><<<nb, 1024, 0, stream>>>(G, W, k, shift, transpose_in);
            else jacobi_prep_kernel<false><<<nb, 1024, 0, stream>>>(G, W, k, shift, transpose_in);
            CTMB_CUDA(cudaGetLastError());
            bool ok;
            if (k <= 128) ok = cplx ? jacobi_cluster_run<true, 4>(G, W, sig, nb, k, max_sweeps, stream) : jacobi_cluster_run<false, 4>(G, W, sig, nb, k, max_sweeps, stream);
            else ok = cplx ? jacobi_cluster_run<true, 8>(G, W, sig, nb, k, max_sweeps, stream) : jacobi_cluster_run<false, 8>(G, W, sig, nb, k, max_sweeps, stream);
            if (ok) return;
        }
    }
    if (!use_smem && coop_mode) {
        int dev = 0, nsm = 0;
        CTMB_CUDA(cudaGetDevice(&dev));
        CTMB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        int cpm = std::max(1, std::min(nsm / nb, ((k + 1) / 2 + JC_THREADS / 32 - 1) / (JC_THREADS / 32)));
        int mxs = max_sweeps, kk_ = k;
        if (cplx) jacobi_prep_kernel<true><<<nb, 1024, 0, stream>>>(G, W, k, shift, transpose_in);
        else jacobi_prep_kernel<false><<<nb, 1024, 0, stream>>>(G, W, k, shift, transpose_in);
        CTMB_CUDA(cudaGetLastError());
        PtrBatch g = G, w = W, sg = sig;
        void* args[] = {&g, &w, &sg, &kk_, &cpm, &mxs};
        CTMB_CHECK(k <= 512, "the multi-CTA Jacobi kernel caches columns of up to 512 rows in registers");
        const void* kern;
        if (k <= 128) kern = cplx ? (const void*)jacobi_coop_kernel<true, 4> : (const void*)jacobi_coop_kernel<false, 4>;
        else if (k <= 256) kern = cplx ? (const void*)jacobi_coop_kernel<true, 8> : (const void*)jacobi_coop_kernel<false, 8>;
        else kern = cplx ? (const void*)jacobi_coop_kernel<true, 16> : (const void*)jacobi_coop_kernel<false, 16>;
        CTMB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(nb * cpm), dim3(JC_THREADS), args, 0, stream));
        return;
    }
    auto launch = [&](auto kern, int threads) {
        static thread_local const void* set_for = nullptr;
        if (smem > 0 && set_for != (const void*)kern) {
            CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jacobi_smem_limit()));
            set_for = (const void*)kern;
        }
        kern<<<nb, threads, smem, stream>>>(G, W, sig, k, max_sweeps, shift, transpose_in);
    };
    if (use_smem) {
        if (cplx) launch(jacobi_kernel<true, true, 16, 512>, 512);
        else if ((k + 1) / 2 <= 48) launch(jacobi_kernel<false, true, 8, 384>, 384);   // 170 registers: no spills
        else launch(jacobi_kernel<false, true, 8, 512>, 512);
    } else {
        if (cplx) launch(jacobi_kernel<true, false, 32, 1024>, 1024);
        else launch(jacobi_kernel<false, false, 32, 1024>, 1024);
    }
    CTMB_CUDA(cudaGetLastError());
}

}  // namespace ctmb
