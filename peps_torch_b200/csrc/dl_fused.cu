// Fused double-layer absorption for the enlarged corner:  (C.T1.T2) . a . conj(a)  ->  n x n corner.
//
// Reference: the last two tensordots + permute/contiguous of c2x2_{LU,RU,RD,LD}_sl_c
// (ctm/generic/ctm_components.py:404-434,563-586,712-733,862-885).  With X = C.T1.T2 viewed per pair
// of environment indices (x1,x2) as a D^2 x D^2 matrix X[(K1,K2),(k1,k2)] (bra legs x ket legs of the
// two contracted auxiliary bonds), the corner block of that pair is
//     Y[(K1,K2),(s,o1,o2)]  = sum_{k1,k2} X[(K1,K2),(k1,k2)] a[s,k1,k2,o1,o2]           (step 3)
//     Out[(o1,o2),(O1,O2)]  = sum_{s,K1,K2} Y[(K1,K2),(s,o1,o2)] conj(a)[s,K1,K2,O1,O2]  (step 4)
// The reference (and the unfused chain here) writes Y to memory: [chi,chi,D^2,p,D^2] = 4.3 GB at D=8,
// chi=256, read back once.  This kernel keeps Y in shared memory: per pair it streams 32 KB of X in and
// 32 KB of the corner out and does 2 x 1 MFLOP on the FP64 tensor pipe (DMMA), i.e. 32 FLOP/B against
// HBM -- compute bound.  One persistent CTA per SM; a (one copy, both steps read it with different row
// groupings), Y and a double-buffered X tile live in 204 KB of shared memory; X tiles arrive by 16-byte
// cp.async while the previous pair is being multiplied.
#include "common.h"

namespace ctmb {

__device__ __forceinline__ void dl_cp_async_16(void* smem, const void* g) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g));
}
__device__ __forceinline__ void dl_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void dl_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void dl_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int DK, int DO, int P>
__global__ void __launch_bounds__(256, 1) dl_corner_kernel(const __grid_constant__ DlParams p) {
    constexpr int DK2 = DK * DK, DO2 = DO * DO, N3 = P * DO2, LDK = DK2 + 4;
    static_assert(DK2 % 32 == 0 && DO2 % 32 == 0 && N3 % 128 == 0 && DO2 % 64 == 0, "tile shape");
    extern __shared__ __align__(16) double dl_smem[];
    double* As = dl_smem;                       // [N3][LDK]   a[(s,o1,o2)][(k1,k2)]
    double* Ys = As + N3 * LDK;                 // [N3][LDK]   Y[(s,o1,o2)][(K1,K2)]
    double* Xs = Ys + N3 * LDK;                 // [2][DK2][LDK]  X[(K1,K2)][(k1,k2)]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int e = tid; e < N3 * DK2; e += 256) {
        const int n = e / DK2, k = e % DK2;
        const int s = n / DO2, o1 = (n % DO2) / DO, o2 = n % DO;
        const int k1 = k / DK, k2 = k % DK;
        As[n * LDK + k] = p.a[s * p.as_s + k1 * p.as_k1 + k2 * p.as_k2 + o1 * p.as_o1 + o2 * p.as_o2];
    }
    auto load_x = [&](int pair, int buf) {
        const double* src = p.X + (size_t)pair * DK2 * DK2;
        double* dst = Xs + buf * DK2 * LDK;
        for (int ch = tid; ch < DK2 * DK2 / 2; ch += 256) {
            const int row = ch / (DK2 / 2), col = (ch % (DK2 / 2)) * 2;
            dl_cp_async_16(dst + row * LDK + col, src + row * DK2 + col);
        }
    };
    int pair = blockIdx.x, buf = 0;
    if (pair < p.npairs) load_x(pair, 0);
    dl_commit();

    const int wm = warp & 1, wn = warp >> 1;    // 2 x 4 warps
    for (; pair < p.npairs; pair += gridDim.x, buf ^= 1) {
        const int nxt = pair + gridDim.x;
        if (nxt < p.npairs) load_x(nxt, buf ^ 1);
        dl_commit();
        dl_wait<1>();
        __syncthreads();                        // X tile visible; everybody is done with Ys of the previous pair
        // ---- step 3: Y[m][n] = sum_k Xs[m][k] As[n][k],  m = (K1,K2) 0..DK2-1,  n = (s,o1,o2) 0..N3-1
        {
            constexpr int WM = DK2 / 2, WN = N3 / 4, TM = WM / 8, TN = WN / 8;
            double acc[TM][TN][2];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            const double* xa = Xs + buf * DK2 * LDK + (wm * WM + (lane >> 2)) * LDK + (lane & 3);
            const double* ab = As + (wn * WN + (lane >> 2)) * LDK + (lane & 3);
#pragma unroll 4
            for (int kk = 0; kk < DK2; kk += 4) {
                double af[TM], bf[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) af[i] = xa[i * 8 * LDK + kk];
#pragma unroll
                for (int j = 0; j < TN; ++j) bf[j] = ab[j * 8 * LDK + kk];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dl_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const int m = wm * WM + i * 8 + (lane >> 2);
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const int n = wn * WN + j * 8 + (lane & 3) * 2;
                    Ys[n * LDK + m] = acc[i][j][0];
                    Ys[(n + 1) * LDK + m] = acc[i][j][1];
                }
            }
        }
        __syncthreads();
        // ---- step 4: Out[m][n] = sum_s sum_kk Ys[(s,m)][kk] As[(s,n)][kk],  m = (o1,o2), n = (O1,O2)
        {
            constexpr int WM = DO2 / 2, WN = DO2 / 4, TM = WM / 8, TN = WN / 8;
            double acc[TM][TN][2];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
            for (int s = 0; s < P; ++s) {
                const double* ya = Ys + (s * DO2 + wm * WM + (lane >> 2)) * LDK + (lane & 3);
                const double* ab = As + (s * DO2 + wn * WN + (lane >> 2)) * LDK + (lane & 3);
#pragma unroll 4
                for (int kk = 0; kk < DK2; kk += 4) {
                    double af[TM], bf[TN];
#pragma unroll
                    for (int i = 0; i < TM; ++i) af[i] = ya[i * 8 * LDK + kk];
#pragma unroll
                    for (int j = 0; j < TN; ++j) bf[j] = ab[j * 8 * LDK + kk];
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < TN; ++j) dl_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
            }
            double* o = p.out + (size_t)(pair / p.n2) * p.st1 + (size_t)(pair % p.n2) * p.st2;
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const int m = wm * WM + i * 8 + (lane >> 2);
                const int rom = p.ro[m];
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const int n = wn * WN + j * 8 + (lane & 3) * 2;
                    o[rom + p.co[n]] = acc[i][j][0];
                    o[rom + p.co[n + 1]] = acc[i][j][1];
                }
            }
        }
    }
    dl_wait<0>();
}

bool dl_corner_supported(int dk1, int dk2, int do1, int do2, int pdim, bool cplx) {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("CTMB_DL_FUSED"); mode = e ? atoi(e) : 1; }
    return mode && !cplx && dk1 == 8 && dk2 == 8 && do1 == 8 && do2 == 8 && pdim == 2;
}

void dl_corner_launch(const DlParams& p, int dk, int dopen, int pdim, cudaStream_t stream) {
    CTMB_CHECK(dk == 8 && dopen == 8 && pdim == 2, "dl_corner: unsupported shape");
    auto kern = dl_corner_kernel<8, 8, 2>;
    constexpr size_t smem = (size_t)(2 * 128 * 68 + 2 * 64 * 68) * 8;
    static bool set = false;
    if (!set) { CTMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
    int dev = 0, nsm = 0;
    CTMB_CUDA(cudaGetDevice(&dev));
    CTMB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    kern<<<std::min(nsm, p.npairs), 256, smem, stream>>>(p);
    CTMB_CUDA(cudaGetLastError());
}

}  // namespace ctmb
