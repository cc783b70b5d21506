// TMA-fed FP64 tensor-contraction GEMM for operands that are plain strided matrices (sm_100a).
//
// At chi D^2 = 16384 (config 5) 96 % of a CTM move is spent in n x n x k products whose operands are dense matrices:
// the enlarged corners H (row-major n x n, used as H or H^T), the k-column blocks of the range finder and the
// projector products.  For those the per-element offset-table gathers of tc_kernel_ws (4 producer warps, 32 cp.async
// and two table look-ups per thread and k-tile) are replaced by the Tensor Memory Accelerator: ONE elected thread issues
// two `cp.async.bulk.tensor` per k-tile (SASS: UTMALDG), completion is counted in bytes on the stage's mbarrier
// (`mbarrier.arrive.expect_tx` / `complete_tx::bytes`), the eight consumer warps are unchanged in structure (LDS.64 +
// DMMA.8x8x4 -- tcgen05 has no f64 kind) and hand stages back through `empty` mbarriers.
//
// Shared-memory layouts written by the TMA unit (both conflict-free for the DMMA fragment pattern row = lane>>2,
// k = lane&3, i.e. eight rows x four k per warp load):
//   k-fast operand (k contiguous in global memory): 2-D map, box {16 k, 128 rows}, SWIZZLE_128B: element (row, k) at
//       row*128 B + ((k>>1) ^ (row&7))*16 B + (k&1)*8 B  -- the eight rows of a fragment hit eight different 16-byte chunks;
//   m-fast operand (rows contiguous in global memory): 3-D map {8 m_lo, K, M/8}, box {8, 16 k, 16 m_hi}, no swizzle:
//       element (k, m) at ((m>>3)*16 + k)*64 B + (m&7)*8 B  -- consecutive k are 64 B apart, so the four k of a fragment
//       alternate between the two halves of the 128-byte bank window (two wavefronts for 256 B: the minimum).
// In global memory the 3-D box is still one contiguous kilobyte per k (m_hi has stride 64 B).
//
// Scheduling: one CTA per output tile, or -- where the tile count leaves the last wave of CTAs under-filled -- stream-K:
// the k-tiles of all tiles are dealt evenly to one persistent CTA per SM (see the kernel and tc_run_tma).
#include "common.h"
#include <cuda.h>
#include <cstring>
#include <cmath>

namespace ctmb {

namespace {

constexpr int TM_BM = 128, TM_BN = 128, TM_BK = 16, TM_STAGES = 6;
constexpr int TM_CONSUMERS = 256, TM_THREADS = TM_CONSUMERS + 32;
constexpr int TM_TILE = TM_BM * TM_BK;                       // doubles per operand stage (16 KB)
constexpr unsigned TM_STAGE_BYTES = 2u * TM_TILE * 8u;

struct TmaMaps {
    CUtensorMap a[TC_MAX_BATCH];
    CUtensorMap b[TC_MAX_BATCH];
};

__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TMA_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TMA_WAIT_DONE;\n"
        "bra TMA_WAIT_LOOP;\n"
        "TMA_WAIT_DONE:\n"
        "}\n" ::"r"(s_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(s_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 ::"r"(s_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma8(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <bool A_KF, bool B_KF>
__global__ void __launch_bounds__(TM_THREADS, 1) tc_kernel_tma(const __grid_constant__ TcParams p,
                                                                const __grid_constant__ TmaMaps maps) {
    constexpr int BM = TM_BM, BN = TM_BN, BK = TM_BK, STAGES = TM_STAGES;
    constexpr int WM = 64, WN = 32, TM = WM / 8, TN = WN / 8;
    extern __shared__ unsigned char tma_smem_raw[];
    // SWIZZLE_128B tiles must start on a 1024-byte boundary of the shared window
    unsigned char* base = tma_smem_raw + ((1024u - (s_u32(tma_smem_raw) & 1023u)) & 1023u);
    double* As = reinterpret_cast<double*>(base);
    double* Bs = As + STAGES * TM_TILE;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(Bs + STAGES * TM_TILE);
    unsigned long long* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntn = (p.N + BN - 1) / BN;
    const int M = p.M, N = p.N;
    // Work = units of one k-tile, ordered (batch entry, output tile, k-tile).  A CTA owns the contiguous range [u0, u1) and
    // walks it in SEGMENTS (the part of its range inside one output tile):
    //   standard launch  grid (tiles, 1, nbatch): one segment = one whole tile;
    //   stream-K launch  grid (G, 1, 1), p.pad = 1: the units are dealt evenly to G = #SM CTAs, so no wave is left
    //   under-filled; a tile cut between two CTAs is completed in C by red.global.add.f64 (C zeroed by the launcher).  The
    //   launcher only chooses it when a CTA's share is at least one tile long, so a tile has at most TWO contributors and the
    //   sum does not depend on their order (0 + a + b == 0 + b + a): deterministic.
    const long long KT = p.K / BK;                            // the launcher only takes K % 16 == 0
    const long long per_entry = (long long)ntn * ((M + BM - 1) / BM) * KT;
    long long u0, u1;
    if (p.pad) {
        const long long total = per_entry * p.nbatch;
        u0 = total * blockIdx.x / gridDim.x;
        u1 = total * (blockIdx.x + 1) / gridDim.x;
    } else {
        u0 = ((long long)blockIdx.z * (per_entry / KT) + blockIdx.x) * KT;
        u1 = u0 + KT;
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], TM_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == TM_CONSUMERS / 32) {
        // ------------------------------ producer: one elected lane drives the TMA unit ------------------------------
        if (lane == 0) {
            int it = 0;                                       // k-tiles issued so far: stage = it % STAGES
            for (long long u = u0; u < u1;) {
                const int b = (int)(u / per_entry);
                const long long rem = u - (long long)b * per_entry;
                const int tile = (int)(rem / KT), kbeg = (int)(rem % KT);
                const int kend = (int)((u1 - u < KT - kbeg) ? kbeg + (u1 - u) : KT);
                const int m0 = (tile / ntn) * BM, n0 = (tile % ntn) * BN;
                const CUtensorMap* ma = &maps.a[b];
                const CUtensorMap* mb = &maps.b[b];
                for (int kt = kbeg; kt < kend; ++kt, ++it) {
                    const int s = it % STAGES;
                    mb_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    mb_expect_tx(&full[s], TM_STAGE_BYTES);
                    const int k0 = kt * BK;
                    if (A_KF) tma_load_2d(As + s * TM_TILE, ma, k0, m0, &full[s]);
                    else tma_load_3d(As + s * TM_TILE, ma, 0, k0, m0 >> 3, &full[s]);
                    if (B_KF) tma_load_2d(Bs + s * TM_TILE, mb, k0, n0, &full[s]);
                    else tma_load_3d(Bs + s * TM_TILE, mb, 0, k0, n0 >> 3, &full[s]);
                }
                u += kend - kbeg;
            }
        }
        return;
    }
    // ------------------------------ consumers ------------------------------
    constexpr int WARPS_M = BM / WM;
    const int wm0 = (warp % WARPS_M) * WM, wn0 = (warp / WARPS_M) * WN;
    const int fr = lane >> 2, fk = lane & 3;
    // element offsets (doubles) of this lane's fragment entries: base + i * (eight rows) + koff[kk / 4]
    int a_koff[BK / 4], b_koff[BK / 4];
#pragma unroll
    for (int q = 0; q < BK / 4; ++q) {
        const int k = 4 * q + fk;
        const int kf = ((((k >> 1) ^ fr) << 1) | (k & 1));   // k-fast, 128-byte swizzle (row & 7 == fr)
        const int mf = k * 8;                                 // m-fast: k stride 64 B
        a_koff[q] = A_KF ? kf : mf;
        b_koff[q] = B_KF ? kf : mf;
    }
    const int a_base = A_KF ? (wm0 + fr) * BK : (wm0 >> 3) * (BK * 8) + fr;
    const int b_base = B_KF ? (wn0 + fr) * BK : (wn0 >> 3) * (BK * 8) + fr;
    constexpr int a_si = A_KF ? 8 * BK : BK * 8;              // eight rows further: 8 x 128 B resp. one m_hi block (16 x 64 B)
    constexpr int b_si = B_KF ? 8 * BK : BK * 8;
    const double alpha = p.alpha;
    int it = 0;
    for (long long u = u0; u < u1;) {
        const int b = (int)(u / per_entry);
        const long long rem = u - (long long)b * per_entry;
        const int tile = (int)(rem / KT), kbeg = (int)(rem % KT);
        const int kend = (int)((u1 - u < KT - kbeg) ? kbeg + (u1 - u) : KT);
        const int m0 = (tile / ntn) * BM, n0 = (tile % ntn) * BN;
        double acc[TM][TN][2];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kt = kbeg; kt < kend; ++kt, ++it) {
            const int s = it % STAGES;
            mb_wait(&full[s], (it / STAGES) & 1);
            const double* as = As + s * TM_TILE + a_base;
            const double* bs = Bs + s * TM_TILE + b_base;
#pragma unroll
            for (int q = 0; q < BK / 4; ++q) {
                double af[TM], bf[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) af[i] = as[i * a_si + a_koff[q]];
#pragma unroll
                for (int j = 0; j < TN; ++j) bf[j] = bs[j * b_si + b_koff[q]];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma8(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
            if (lane == 0) mb_arrive(&empty[s]);
        }
        u += kend - kbeg;
        // ---- epilogue of the segment ----
        const TcBatchEntry be = p.batch[b];
        const TcTables tb = p.tab[be.tab];
        double* __restrict__ C = reinterpret_cast<double*>(be.C);
        const bool accum = (be.flags & TC_ACCUM) != 0;
        const bool partial = !(kbeg == 0 && kend == (int)KT);
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
                    if (partial) { atomicAdd(&C[ro + tb.c_n[c]], v); continue; }
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
}

int tma_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0; cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    }
    return fn;
}

// matrix with `rows` rows and K columns: element (row, k) at base + row * s_row + k * s_k (element strides)
bool encode_operand(CUtensorMap* map, const void* base, int rows, int K, long long s_row, long long s_k, bool kfast) {
    EncodeFn enc = encode_fn();
    if (!enc) return false;
    if (((uintptr_t)base & 15) != 0) return false;
    const cuuint32_t ones[3] = {1, 1, 1};
    if (kfast) {
        if (s_k != 1 || (s_row & 1) != 0 || s_row <= 0) return false;
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)s_row * 8};
        const cuuint32_t box[2] = {(cuuint32_t)TM_BK, (cuuint32_t)TM_BM};
        return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims, strides, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (s_row != 1 || (s_k & 1) != 0 || s_k <= 0 || (rows & 7) != 0) return false;
    const cuuint64_t dims[3] = {8, (cuuint64_t)K, (cuuint64_t)(rows / 8)};
    const cuuint64_t strides[2] = {(cuuint64_t)s_k * 8, 64};
    const cuuint32_t box[3] = {8, (cuuint32_t)TM_BK, (cuuint32_t)(TM_BM / 8)};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool A_KF, bool B_KF>
void run_tma_t(const TcParams& p0, const TmaMaps& maps, bool streamk, cudaStream_t stream) {
    auto kern = tc_kernel_tma<A_KF, B_KF>;
    static bool attr_set = false;
    const size_t smem = (size_t)TM_STAGES * TM_STAGE_BYTES + 2 * TM_STAGES * 8 + 1024;
    if (!attr_set) {
        CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long tiles = (long long)((p0.N + TM_BN - 1) / TM_BN) * ((p0.M + TM_BM - 1) / TM_BM);
    CTMB_CHECK(tiles < (1ll << 31), "too many tiles for one launch");
    TcParams p = p0;
    p.pad = streamk ? 1 : 0;
    dim3 grid = streamk ? dim3((unsigned)tma_sm_count(), 1, 1) : dim3((unsigned)tiles, 1, p.nbatch);
    kern<<<grid, TM_THREADS, smem, stream>>>(p, maps);
    CTMB_CUDA(cudaGetLastError());
}

}  // namespace

// Launches the batch through the TMA-fed kernel if every entry qualifies (plain strided operands with the alignment the
// tensor maps need, K a multiple of 16, uniform fast directions); returns false otherwise and launches nothing.
bool tc_run_tma(const TcParams& p, bool a_kfast, bool b_kfast, cudaStream_t stream) {
    static int mode = -1;
    if (mode < 0) { const char* ev = getenv("CTMB_GEMM_TMA"); mode = ev ? atoi(ev) : 1; }
    if (!mode || (p.K % TM_BK) != 0 || p.ksplit > 1) return false;
    TmaMaps maps;
    for (int i = 0; i < p.nbatch; ++i) {
        const TcBatchEntry& be = p.batch[i];
        const TcTables& tb = p.tab[be.tab];
        if (!tb.lin) return false;
        if (!encode_operand(&maps.a[i], be.A, p.M, p.K, tb.a_sm, tb.a_sk, a_kfast)) return false;
        if (!encode_operand(&maps.b[i], be.B, p.N, p.K, tb.b_sn, tb.b_sk, b_kfast)) return false;
    }
    // Wave quantisation: one CTA per SM, so 512 tiles (a 16384 x 512 block of the range finder) take four waves of 148 of
    // which the last is 46 % full -- 13 % of the tensor pipe idle; 256 tiles (the half block of a group member) 1.73 waves.
    // Where that costs more than 7 %, the launch is scheduled stream-K (see the kernel): the k-tiles of all output tiles are
    // dealt evenly to one CTA per SM and tiles cut between two CTAs are completed with red.global.add.f64 into a zeroed C.
    // Chosen only when a CTA's share is at least one tile long (tiles >= #SM), so every element has at most two addends.
    bool streamk = false;
    {
        static int mode2 = -1;
        if (mode2 < 0) { const char* ev = getenv("CTMB_GEMM_TMA_SPLIT"); mode2 = ev ? atoi(ev) : 1; }
        const long long units = (long long)((p.N + TM_BN - 1) / TM_BN) * ((p.M + TM_BM - 1) / TM_BM) * p.nbatch;
        const int sms = tma_sm_count();
        const double w = (double)units / sms;
        bool ok = mode2 && units >= sms && p.K >= 1024 && w / std::ceil(w) < 0.93;
        for (int i = 0; ok && i < p.nbatch; ++i) {
            const TcBatchEntry& be = p.batch[i];
            const TcTables& tb = p.tab[be.tab];
            const bool dense = tb.c_lin && ((tb.c_sm == 1 && tb.c_sn == p.M) || (tb.c_sn == 1 && tb.c_sm == p.N));
            if (!dense || (be.flags & TC_ACCUM) || be.amax != nullptr) ok = false;
        }
        if (ok) {
            streamk = true;
            for (int i = 0; i < p.nbatch; ++i)
                CTMB_CUDA(cudaMemsetAsync(p.batch[i].C, 0, (size_t)p.M * p.N * 8, stream));
        }
    }
    if (a_kfast && b_kfast) run_tma_t<true, true>(p, maps, streamk, stream);
    else if (a_kfast) run_tma_t<true, false>(p, maps, streamk, stream);
    else if (b_kfast) run_tma_t<false, true>(p, maps, streamk, stream);
    else run_tma_t<false, false>(p, maps, streamk, stream);
    return true;
}

}  // namespace ctmb
