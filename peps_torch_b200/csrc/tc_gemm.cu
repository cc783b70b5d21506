// Batched tensor-contraction GEMM on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64) for sm_100a.
//
//   C[c_m[m] + c_n[n]] = alpha * sum_k opA(A[a_m[m] + a_k[k]]) * opB(B[b_k[k] + b_n[n]])
//
// Every tensor contraction of the CTM move (reference: the tensordot/einsum chains of
// ctm/generic/ctm_components.py:372-434, ctm/generic/ctmrg.py:342-438,
// ctm/one_site_c4v/ctmrg_c4v.py:376-443) is a GEMM whose row/column/reduction indices are
// groups of tensor modes.  Because free and contracted modes are disjoint, the address of an
// element is separable into a row part and a column part; both are tabulated once per plan
// (int32 element offsets).  The kernel therefore contracts permuted / strided / fused-leg views
// in place: no permute+contiguous copy is ever materialised (the reference makes one per step).
//
// tcgen05 has no f64 kind (ptxas rejects .kind::f64 for sm_100a), so the FP64 tensor path on
// B200 is the warp-level DMMA; operands are staged global->shared with cp.async (LDGSTS) in a
// multi-stage ring, fragments are read conflict-free from a k-contiguous, (BK+4)-padded layout.
#include "common.h"
#include <type_traits>

namespace ctmb {

__device__ __forceinline__ void cp_async_8(void* smem, const void* g, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(g), "r"(sz));
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* g, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(g), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool CPLX>
__global__ void __launch_bounds__(256) tc_kernel(const __grid_constant__ TcParams p) {
    using T = typename std::conditional<CPLX, double2, double>::type;
    constexpr int LDS = BK + 4;
    constexpr int WARPS_M = BM / WM;
    constexpr int TM = WM / 8, TN = WN / 8;
    constexpr int A_PER = BM * BK / 256, B_PER = BN * BK / 256;
    static_assert((BM / WM) * (BN / WN) == 8, "8 warps");
    static_assert(A_PER >= 1 && B_PER >= 1, "tile too small");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* As = reinterpret_cast<T*>(smem_raw);
    T* Bs = As + STAGES * BM * LDS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TcBatchEntry be = p.batch[blockIdx.z];
    const TcTables tb = p.tab[be.tab];
    // tiles are linearised on grid.x (n-tiles fastest): grid.y is limited to 65535, the row count of a
    // corner intermediate at D=8, chi=256 (4.2M) is not
    const int ntn = (p.N + BN - 1) / BN;
    const int m0 = (blockIdx.x / ntn) * BM, n0 = (blockIdx.x % ntn) * BN;
    const T* __restrict__ A = reinterpret_cast<const T*>(be.A);
    const T* __restrict__ B = reinterpret_cast<const T*>(be.B);
    const int M = p.M, N = p.N, K = p.K;
    const int flags = be.flags;
    const bool a_kfast = (flags & TC_A_KFAST) != 0, b_kfast = (flags & TC_B_KFAST) != 0;

    // loader maps: element e = tid + i*256 of a BMxBK (BNxBK) tile
    int a_row[A_PER], a_col[A_PER], a_off[A_PER];
    int b_row[B_PER], b_col[B_PER], b_off[B_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        int e = tid + i * 256;
        int r = a_kfast ? e / BK : e % BM;
        int c = a_kfast ? e % BK : e / BM;
        a_row[i] = r; a_col[i] = c;
        a_off[i] = (m0 + r < M) ? tb.a_m[m0 + r] : -1;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
        int e = tid + i * 256;
        int r = b_kfast ? e / BK : e % BN;
        int c = b_kfast ? e % BK : e / BN;
        b_row[i] = r; b_col[i] = c;
        b_off[i] = (n0 + r < N) ? tb.b_n[n0 + r] : -1;
    }

    auto load_tile = [&](int kt, int stage) {
        const int k0 = kt * BK;
        T* as = As + stage * BM * LDS;
        T* bs = Bs + stage * BN * LDS;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int k = k0 + a_col[i];
            bool v = (a_off[i] >= 0) && (k < K);
            const T* src = v ? (A + a_off[i] + tb.a_k[k]) : A;
            if (CPLX) cp_async_16(as + a_row[i] * LDS + a_col[i], src, v);
            else cp_async_8(as + a_row[i] * LDS + a_col[i], src, v);
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int k = k0 + b_col[i];
            bool v = (b_off[i] >= 0) && (k < K);
            const T* src = v ? (B + b_off[i] + tb.b_k[k]) : B;
            if (CPLX) cp_async_16(bs + b_row[i] * LDS + b_col[i], src, v);
            else cp_async_8(bs + b_row[i] * LDS + b_col[i], src, v);
        }
    };

    const int wm0 = (warp % WARPS_M) * WM, wn0 = (warp / WARPS_M) * WN;
    double acc[TM][TN][2];
    double acci[CPLX ? TM : 1][CPLX ? TN : 1][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            acc[i][j][0] = acc[i][j][1] = 0.0;
            if constexpr (CPLX) acci[i][j][0] = acci[i][j][1] = 0.0;
        }

    const int ktiles = (K + BK - 1) / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ktiles) load_tile(s, s);
        cp_async_commit();
    }
    const double sa = (flags & TC_CONJ_A) ? -1.0 : 1.0;
    const double sb = (flags & TC_CONJ_B) ? -1.0 : 1.0;

    for (int kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < ktiles) load_tile(nk, nk % STAGES);
            cp_async_commit();
        }
        const T* as = As + (kt % STAGES) * BM * LDS + (wm0 + (lane >> 2)) * LDS + (lane & 3);
        const T* bs = Bs + (kt % STAGES) * BN * LDS + (wn0 + (lane >> 2)) * LDS + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            T af[TM], bf[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) af[i] = as[i * 8 * LDS + kk];
#pragma unroll
            for (int j = 0; j < TN; ++j) bf[j] = bs[j * 8 * LDS + kk];
            if constexpr (!CPLX) {
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const double ar = af[i].x, ai = sa * af[i].y;
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        const double br = bf[j].x, bi = sb * bf[j].y;
                        dmma(acc[i][j][0], acc[i][j][1], ar, br);
                        dmma(acc[i][j][0], acc[i][j][1], -ai, bi);
                        dmma(acci[i][j][0], acci[i][j][1], ar, bi);
                        dmma(acci[i][j][0], acci[i][j][1], ai, br);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: registers -> global through the C offset tables
    T* __restrict__ C = reinterpret_cast<T*>(be.C);
    const double alpha = p.alpha;
    double lmax = 0.0;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = m0 + wm0 + i * 8 + (lane >> 2);
        if (r >= M) continue;
        const int ro = tb.c_m[r];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = n0 + wn0 + j * 8 + (lane & 3) * 2 + h;
                if (c >= N) continue;
                if constexpr (!CPLX) {
                    double v = alpha * acc[i][j][h];
                    C[ro + tb.c_n[c]] = v;
                    lmax = fmax(lmax, fabs(v));
                } else {
                    double2 v = make_double2(alpha * acc[i][j][h], alpha * acci[i][j][h]);
                    C[ro + tb.c_n[c]] = v;
                    lmax = fmax(lmax, hypot(v.x, v.y));
                }
            }
        }
    }
    if (be.amax != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        __shared__ double red[8];
        if (lane == 0) red[warp] = lmax;
        __syncthreads();
        if (tid == 0) {
            double m = red[0];
#pragma unroll
            for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
            atomicMax(be.amax, (unsigned long long)__double_as_longlong(m));
        }
    }
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool CPLX>
static size_t tc_smem() {
    return (size_t)STAGES * (BM + BN) * (BK + 4) * (CPLX ? 16 : 8);
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool CPLX>
static void tc_run(const TcParams& p, cudaStream_t stream) {
    auto kern = tc_kernel<BM, BN, BK, WM, WN, STAGES, CPLX>;
    static bool attr_set = false;
    size_t smem = tc_smem<BM, BN, BK, WM, WN, STAGES, CPLX>();
    if (!attr_set) {
        CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long tiles = (long long)((p.N + BN - 1) / BN) * ((p.M + BM - 1) / BM);
    CTMB_CHECK(tiles < (1ll << 31), "too many tiles for one launch");
    dim3 grid((unsigned)tiles, 1, p.nbatch);
    kern<<<grid, 256, smem, stream>>>(p);
    CTMB_CUDA(cudaGetLastError());
}

void tc_launch(const TcParams& p, bool cplx, cudaStream_t stream) {
    CTMB_CHECK(p.nbatch >= 1 && p.nbatch <= TC_MAX_BATCH, "bad batch count");
    if (p.M == 0 || p.N == 0) return;
    auto ntiles = [&](int bm, int bn) { return (long long)((p.M + bm - 1) / bm) * ((p.N + bn - 1) / bn) * p.nbatch; };
    // One SM retires only ~64 FP64 FMA per clock, so the serial K loop of a CTA is the latency of
    // a small contraction: prefer the tile that spreads the work over at least ~2 waves of CTAs.
    if (!cplx) {
        if (p.N <= 16) tc_run<128, 32, 16, 32, 16, 3, false>(p, stream);
        else if (ntiles(128, 128) >= 240 && p.M > 64 && p.N > 64) tc_run<128, 128, 16, 64, 32, 3, false>(p, stream);
        else if (ntiles(64, 64) >= 296) tc_run<64, 64, 16, 32, 16, 4, false>(p, stream);
        else tc_run<32, 32, 16, 16, 8, 4, false>(p, stream);
    } else {
        if (p.N <= 16) tc_run<128, 32, 16, 32, 16, 3, true>(p, stream);
        else if (ntiles(64, 64) >= 296) tc_run<64, 64, 16, 32, 16, 3, true>(p, stream);
        else tc_run<32, 32, 16, 16, 8, 4, true>(p, stream);
    }
}

}  // namespace ctmb
