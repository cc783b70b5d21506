// Microbenchmark: cost of the synchronisation primitives the QR / Jacobi kernels are built from.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <int MODE>
__global__ void k(double* out, int iters, int CL) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double buf[512];
    __shared__ double tot[256];
    double acc = threadIdx.x;
    buf[threadIdx.x % 512] = acc;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { __syncthreads(); }
        if (MODE == 1) { cluster.sync(); }
        if (MODE == 2) {   // cluster.sync + DSMEM gather + syncthreads
            cluster.sync();
            if (threadIdx.x < 96) {
                double v = 0;
                for (int r = 0; r < CL; ++r) v += cluster.map_shared_rank(buf, r)[threadIdx.x + (i & 1) * 128];
                tot[threadIdx.x] = v;
            }
            __syncthreads();
            acc += tot[i % 96];
        }
        if (MODE == 3) {   // scalar reflector math: sqrt + 2 divisions, dependent
            double a = acc + 1.0, t = a * 0.5 + i;
            double beta = -copysign(sqrt(a * a + t), a);
            double tj = (beta - a) / beta;
            double sc = 1.0 / (a - beta);
            acc = tj + sc;
        }
        if (MODE == 4) {   // warp reduction of 8 doubles (5 steps)
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = acc + q;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
            acc = v[0] + v[3] + v[7];
        }
        if (MODE == 5) {   // arrive/wait split
            cluster.barrier_arrive(); cluster.barrier_wait();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[MODE] = double(t1 - t0) / iters;
    if (acc == 12345.678) out[10] = acc;
}

template <int MODE>
void run(double* d, int threads, int CL, const char* name) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(4 * CL); cfg.blockDim = dim3(threads);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k<MODE>, d, 2000, CL);
    cudaDeviceSynchronize();
    double h[16]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-44s threads=%4d CL=%d : %8.1f cycles/iter  (%s)\n", name, threads, CL, h[MODE], cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double* d; cudaMalloc(&d, 16 * sizeof(double));
    for (int threads : {256, 512, 1024}) {
        run<0>(d, threads, 1, "__syncthreads");
        for (int CL : {1, 2, 4, 8}) {
            run<1>(d, threads, CL, "cluster.sync");
            run<5>(d, threads, CL, "cluster arrive+wait");
            run<2>(d, threads, CL, "cluster.sync + DSMEM gather + syncthreads");
        }
        run<3>(d, threads, 1, "reflector scalar math (sqrt, 2 div)");
        run<4>(d, threads, 1, "warp reduce 8 doubles");
    }
    return 0;
}
