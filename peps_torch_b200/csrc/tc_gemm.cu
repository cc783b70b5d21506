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

    // split-K: this CTA reduces k in [kbeg, kend) only and writes a dense partial tile
    const int kbeg = p.ksplit > 1 ? blockIdx.y * p.kchunk : 0;
    const int kend = p.ksplit > 1 ? min(K, kbeg + p.kchunk) : K;
    auto load_tile = [&](int kt, int stage) {
        const int k0 = kbeg + kt * BK;
        const int K = kend;                    // shadows the full extent: loads beyond the slice are zero-filled
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

    const int ktiles = (kend - kbeg + BK - 1) / BK;
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

    if (p.ksplit > 1) {
        // split-K epilogue: dense partial tile, no alpha / accumulate / amax (the reduction kernel applies them)
        T* __restrict__ P = reinterpret_cast<T*>(p.partial) + ((size_t)(blockIdx.z * p.ksplit + blockIdx.y) * M) * N;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int r = m0 + wm0 + i * 8 + (lane >> 2);
            if (r >= M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = n0 + wn0 + j * 8 + (lane & 3) * 2 + h;
                    if (c >= N) continue;
                    if constexpr (!CPLX) P[(size_t)r * N + c] = acc[i][j][h];
                    else P[(size_t)r * N + c] = make_double2(acc[i][j][h], acci[i][j][h]);
                }
        }
        return;
    }
    // epilogue: registers -> global through the C offset tables
    T* __restrict__ C = reinterpret_cast<T*>(be.C);
    const double alpha = p.alpha;
    const bool accum = (flags & TC_ACCUM) != 0;
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
                    if (accum) v += C[ro + tb.c_n[c]];
                    C[ro + tb.c_n[c]] = v;
                    lmax = fmax(lmax, fabs(v));
                } else {
                    double2 v = make_double2(alpha * acc[i][j][h], alpha * acci[i][j][h]);
                    if (accum) { const double2 o = C[ro + tb.c_n[c]]; v.x += o.x; v.y += o.y; }
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


// ---------------------------------------------------------------------------------------------
// Warp-specialised variant for large contractions (128 x 128 x 16 tiles, double).
// The ncu source view of tc_kernel<128,128,..> on a 4096^3 product (profiles/r1_gemm4096.md) shows
// the DMMA pipe 59 % busy: every warp alternates between ~300 gather/address instructions (one
// of them waiting for the offset-table load) and its 128 DMMAs, and all eight warps meet at a
// __syncthreads per k-tile, so the tensor pipe idles during every load phase.  Here the roles are
// split: 4 PRODUCER warps run the table look-ups and cp.async gathers and signal "full" mbarriers
// (cp.async.mbarrier.arrive.noinc), 8 CONSUMER warps only issue LDS + DMMA and hand stages back
// through "empty" mbarriers.  No block-wide barrier inside the k loop.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_cp_async_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int WS_BM = 128, WS_BN = 128, WS_BK = 16, WS_STAGES = 5;
constexpr int WS_CONSUMERS = 256, WS_PRODUCERS = 128, WS_THREADS = WS_CONSUMERS + WS_PRODUCERS;
// Two shared-memory layouts per operand tile, chosen by the operand's fast direction in global memory so
// that the producers' cp.async writes AND the consumers' fragment reads are both bank-conflict free:
//   k-fast operand:  tile[row][k], row stride LDK = 20  (a warp writes two rows of 16 consecutive k)
//   m-fast operand:  tile[k][row], k stride   LDM = 132 (a warp writes 32 consecutive rows of one k)
// Fragment read of lane l: row = l>>2, k = l&3: half-warp addresses r*20+k resp. k*132+r hit 16 distinct
// 8-byte bank pairs (20 = 4 mod 16, 132 = 4 mod 16).  With the first version's single [row][k] layout the
// m-fast gathers wrote with an 8-way conflict (5.8e8 conflict wavefronts in the corner's C.T1.T2 product,
// shared-memory pipe 64 % busy, DMMA 66 %: profiles/r1_c5_corner.md).
constexpr int WS_LDK = WS_BK + 4, WS_LDM = WS_BM + 4;
constexpr int WS_TILE = (WS_BM * WS_LDK > WS_BK * WS_LDM) ? WS_BM * WS_LDK : WS_BK * WS_LDM;   // doubles per operand stage

template <bool A_KF, bool B_KF>
__global__ void __launch_bounds__(WS_THREADS, 1) tc_kernel_ws(const __grid_constant__ TcParams p) {
    constexpr int BM = WS_BM, BN = WS_BN, BK = WS_BK, STAGES = WS_STAGES;
    constexpr int WM = 64, WN = 32, TM = WM / 8, TN = WN / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + STAGES * WS_TILE;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(Bs + STAGES * WS_TILE);
    unsigned long long* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TcBatchEntry be = p.batch[blockIdx.z];
    const TcTables tb = p.tab[be.tab];
    const int ntn = (p.N + BN - 1) / BN;
    const int m0 = (blockIdx.x / ntn) * BM, n0 = (blockIdx.x % ntn) * BN;
    const int M = p.M, N = p.N, K = p.K;
    const int ktiles = (K + BK - 1) / BK;
    constexpr bool a_kfast = A_KF, b_kfast = B_KF;          // uniform over the batch (checked by the launcher)
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], WS_PRODUCERS); mbar_init(&empty[s], WS_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (tid >= WS_CONSUMERS) {
        // ------------------------------ producers ------------------------------
        const int pt = tid - WS_CONSUMERS;
        const double* __restrict__ A = reinterpret_cast<const double*>(be.A);
        const double* __restrict__ B = reinterpret_cast<const double*>(be.B);
        constexpr int PER = BM * BK / WS_PRODUCERS;          // 16 elements of A and of B per thread and k-tile
        // element e = pt + i*128 of a 128 x 16 tile: k-fast operands -> (row pt/16 + 8i, k pt%16), m-fast -> (row pt, k i);
        // everything but the row offsets is an arithmetic progression in i (128 % 16 == 0)
        const int a_c0 = a_kfast ? (pt & 15) : 0, b_c0 = b_kfast ? (pt & 15) : 0;
        constexpr int a_cs = a_kfast ? 0 : 1, b_cs = b_kfast ? 0 : 1;
        const int a_s0 = a_kfast ? (pt >> 4) * WS_LDK + (pt & 15) : pt, b_s0 = b_kfast ? (pt >> 4) * WS_LDK + (pt & 15) : pt;
        constexpr int a_ss = a_kfast ? 8 * WS_LDK : WS_LDM, b_ss = b_kfast ? 8 * WS_LDK : WS_LDM;
        int a_off[PER], b_off[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int ar = a_kfast ? (pt >> 4) + 8 * i : pt, br = b_kfast ? (pt >> 4) + 8 * i : pt;
            a_off[i] = (m0 + ar < M) ? tb.a_m[m0 + ar] : -1;
            b_off[i] = (n0 + br < N) ? tb.b_n[n0 + br] : -1;
        }
        for (int kt = 0; kt < ktiles; ++kt) {
            const int s = kt % STAGES;
            mbar_wait(&empty[s], ((kt / STAGES) & 1) ^ 1);
            double* as = As + s * WS_TILE + a_s0;
            double* bs = Bs + s * WS_TILE + b_s0;
            const int k0 = kt * BK;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int ka = k0 + a_c0 + i * a_cs, kb = k0 + b_c0 + i * b_cs;
                const bool va = (a_off[i] >= 0) && (ka < K);
                const bool vb = (b_off[i] >= 0) && (kb < K);
                cp_async_8(as + i * a_ss, va ? (A + a_off[i] + tb.a_k[ka]) : A, va);
                cp_async_8(bs + i * b_ss, vb ? (B + b_off[i] + tb.b_k[kb]) : B, vb);
            }
            mbar_cp_async_arrive(&full[s]);
        }
        return;
    }
    // ------------------------------ consumers ------------------------------
    constexpr int WARPS_M = BM / WM;
    const int wm0 = (warp % WARPS_M) * WM, wn0 = (warp / WARPS_M) * WN;
    const int fr = lane >> 2, fk = lane & 3;
    // fragment addressing: base + i * (8 rows) + kk * (k stride)
    const int a_base = a_kfast ? (wm0 + fr) * WS_LDK + fk : fk * WS_LDM + wm0 + fr;
    const int b_base = b_kfast ? (wn0 + fr) * WS_LDK + fk : fk * WS_LDM + wn0 + fr;
    constexpr int a_si = a_kfast ? 8 * WS_LDK : 8, a_sk = a_kfast ? 1 : WS_LDM;
    constexpr int b_si = b_kfast ? 8 * WS_LDK : 8, b_sk = b_kfast ? 1 : WS_LDM;
    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < ktiles; ++kt) {
        const int s = kt % STAGES;
        mbar_wait(&full[s], (kt / STAGES) & 1);
        const double* as = As + s * WS_TILE + a_base;
        const double* bs = Bs + s * WS_TILE + b_base;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[TM], bf[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) af[i] = as[i * a_si + kk * a_sk];
#pragma unroll
            for (int j = 0; j < TN; ++j) bf[j] = bs[j * b_si + kk * b_sk];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    double* __restrict__ C = reinterpret_cast<double*>(be.C);
    const double alpha = p.alpha;
    const bool accum = (be.flags & TC_ACCUM) != 0;
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
                double v = alpha * acc[i][j][h];
                if (accum) v += C[ro + tb.c_n[c]];
                C[ro + tb.c_n[c]] = v;
                lmax = fmax(lmax, fabs(v));
            }
        }
    }
    if (be.amax != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if (lane == 0) atomicMax(be.amax, (unsigned long long)__double_as_longlong(lmax));
    }
}

template <bool A_KF, bool B_KF>
static void tc_run_ws_t(const TcParams& p, cudaStream_t stream) {
    auto kern = tc_kernel_ws<A_KF, B_KF>;
    static bool attr_set = false;
    const size_t smem = (size_t)WS_STAGES * 2 * WS_TILE * 8 + 2 * WS_STAGES * 8;
    if (!attr_set) {
        CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long tiles = (long long)((p.N + WS_BN - 1) / WS_BN) * ((p.M + WS_BM - 1) / WS_BM);
    CTMB_CHECK(tiles < (1ll << 31), "too many tiles for one launch");
    dim3 grid((unsigned)tiles, 1, p.nbatch);
    kern<<<grid, WS_THREADS, smem, stream>>>(p);
    CTMB_CUDA(cudaGetLastError());
}
// the kernel is specialised on the operands' fast directions: a batch with mixed loader flags (the eight
// corner x corner products of the halves have different transposes) is launched as one sub-batch per flag pair
static bool tc_run_ws(const TcParams& p, cudaStream_t stream) {
    for (int lf = 0; lf < 4; ++lf) {
        const bool ak = (lf & 1) != 0, bk = (lf & 2) != 0;
        TcParams q = p;
        q.nbatch = 0;
        for (int i = 0; i < p.nbatch; ++i) {
            const int f = p.batch[i].flags;
            if (((f & TC_A_KFAST) != 0) == ak && ((f & TC_B_KFAST) != 0) == bk) q.batch[q.nbatch++] = p.batch[i];
        }
        if (q.nbatch == 0) continue;
        if (tc_run_tma(q, ak, bk, stream)) continue;          // plain strided operands: fed by the TMA unit
        if (ak && bk) tc_run_ws_t<true, true>(q, stream);
        else if (ak) tc_run_ws_t<true, false>(q, stream);
        else if (bk) tc_run_ws_t<false, true>(q, stream);
        else tc_run_ws_t<false, false>(q, stream);
    }
    return true;
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
    dim3 grid((unsigned)tiles, p.ksplit > 1 ? p.ksplit : 1, p.nbatch);
    kern<<<grid, 256, smem, stream>>>(p);
    CTMB_CUDA(cudaGetLastError());
}

template <bool CPLX>
__global__ void tc_splitk_reduce_kernel(const __grid_constant__ TcParams p) {
    using T = typename std::conditional<CPLX, double2, double>::type;
    const TcBatchEntry be = p.batch[blockIdx.y];
    const TcTables tb = p.tab[be.tab];
    const long long mn = (long long)p.M * p.N;
    const T* __restrict__ P = reinterpret_cast<const T*>(p.partial) + (size_t)blockIdx.y * p.ksplit * mn;
    T* __restrict__ C = reinterpret_cast<T*>(be.C);
    const bool accum = (be.flags & TC_ACCUM) != 0;
    double lmax = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < mn; e += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(e / p.N), n = (int)(e % p.N);
        const int off = tb.c_m[m] + tb.c_n[n];
        if constexpr (!CPLX) {
            double v = 0.0;
            for (int s = 0; s < p.ksplit; ++s) v += P[(size_t)s * mn + e];
            v *= p.alpha;
            if (accum) v += C[off];
            C[off] = v;
            lmax = fmax(lmax, fabs(v));
        } else {
            double2 v = make_double2(0.0, 0.0);
            for (int s = 0; s < p.ksplit; ++s) { const double2 q = P[(size_t)s * mn + e]; v.x += q.x; v.y += q.y; }
            v.x *= p.alpha; v.y *= p.alpha;
            if (accum) { const double2 o = C[off]; v.x += o.x; v.y += o.y; }
            C[off] = v;
            lmax = fmax(lmax, hypot(v.x, v.y));
        }
    }
    if (be.amax != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if ((threadIdx.x & 31) == 0) atomicMax(be.amax, (unsigned long long)__double_as_longlong(lmax));
    }
}
void tc_splitk_reduce_launch(const TcParams& p, bool cplx, cudaStream_t stream) {
    const long long mn = (long long)p.M * p.N;
    dim3 grid((unsigned)std::min<long long>(256, (mn + 255) / 256), p.nbatch);
    if (cplx) tc_splitk_reduce_kernel<true><<<grid, 256, 0, stream>>>(p);
    else tc_splitk_reduce_kernel<false><<<grid, 256, 0, stream>>>(p);
    CTMB_CUDA(cudaGetLastError());
}

// 0: 128x32, 1: 128x128 (warp-specialised when possible), 2: 64x64, 3: 32x32
static int tc_config(const TcParams& p, bool cplx) {
    auto ntiles = [&](int bm, int bn) { return (long long)((p.M + bm - 1) / bm) * ((p.N + bn - 1) / bn) * p.nbatch; };
    if (p.N <= 16) return 0;
    if (!cplx && p.ksplit <= 1 && ntiles(128, 128) >= 240 && p.M > 64 && p.N > 64) return 1;
    if (ntiles(64, 64) >= 296) return 2;
    return 3;
}
void tc_tile_shape(const TcParams& p, bool cplx, int& bm, int& bn) {
    switch (tc_config(p, cplx)) {
        case 0: bm = 128; bn = 32; break;
        case 1: bm = 128; bn = 128; break;
        case 2: bm = 64; bn = 64; break;
        default: bm = 32; bn = 32; break;
    }
}

void tc_launch(const TcParams& p, bool cplx, cudaStream_t stream) {
    CTMB_CHECK(p.nbatch >= 1 && p.nbatch <= TC_MAX_BATCH, "bad batch count");
    if (p.M == 0 || p.N == 0) return;
    // One SM retires only ~64 FP64 FMA per clock, so the serial K loop of a CTA is the latency of
    // a small contraction: prefer the tile that spreads the work over at least ~2 waves of CTAs.
    const int cfg = tc_config(p, cplx);
    if (!cplx) {
        if (cfg == 0) tc_run<128, 32, 16, 32, 16, 3, false>(p, stream);
        else if (cfg == 1) {
            static int ws_mode = -1;
            if (ws_mode < 0) { const char* ev = getenv("CTMB_GEMM_WS"); ws_mode = ev ? atoi(ev) : 1; }
            if (!ws_mode || !tc_run_ws(p, stream)) tc_run<128, 128, 16, 64, 32, 3, false>(p, stream);
        }
        else if (cfg == 2) tc_run<64, 64, 16, 32, 16, 4, false>(p, stream);
        else tc_run<32, 32, 32, 16, 8, 3, false>(p, stream);
    } else {
        if (cfg == 0) tc_run<128, 32, 16, 32, 16, 3, true>(p, stream);
        else if (cfg == 2) tc_run<64, 64, 16, 32, 16, 3, true>(p, stream);
        else tc_run<32, 32, 16, 16, 8, 4, true>(p, stream);
    }
}

}  // namespace ctmb
