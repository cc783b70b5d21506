// Stand-alone timing of the QR kernels on 4 matrices (default 432x96, config c2).
#include "../../peps_torch_b200/csrc/common.h"
#include <vector>
#include <cstdlib>
using namespace ctmb;
#ifdef QR_PROFILE
namespace ctmb { void qr_profile_dump(int steps); }
#endif
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DQR_PROFILE] -o tools/micro/qr_bench tools/micro/qr_bench.cu peps_torch_b200/csrc/qr.cu
int main(int argc, char** argv) {
    int rows = argc > 1 ? atoi(argv[1]) : 432, cols = argc > 2 ? atoi(argv[2]) : 96, nb = 4;
    std::vector<double> h((size_t)rows * cols);
    PtrBatch A{}, R{}, Tau{}, G{}, X{};
    for (int b = 0; b < nb; ++b) {
        for (auto& x : h) x = rand() / (double)RAND_MAX - 0.5;
        cudaMalloc(&A.p[b], h.size() * 8); cudaMalloc(&Tau.p[b], cols * 8);
        cudaMalloc(&G.p[b], cols * cols * 8); cudaMalloc(&X.p[b], cols * cols * 8);
        cudaMemset(G.p[b], 0, cols * cols * 8);
        cudaMemcpy(A.p[b], h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    }
    cudaEvent_t e0, e1, e2, e3; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    for (int it = 0; it < 3; ++it) {
        for (int b = 0; b < nb; ++b) cudaMemcpy(A.p[b], h.data(), h.size() * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
#ifndef QR_PROFILE
        qr_launch(A, R, nb, rows, cols, rows, false, 0);
#endif
        cudaEventRecord(e1);
        for (int b = 0; b < nb; ++b) cudaMemcpyAsync(A.p[b], h.data(), h.size() * 8, cudaMemcpyHostToDevice, 0);
        cudaEventRecord(e2);
        float ms_f = 0, ms_t = 0;
        if (qr_wy_supported(rows, cols, false)) {
            qr_wy_factor_launch(A, R, Tau, nb, rows, cols, rows, false, 0);
            cudaEventRecord(e3);
            cudaEvent_t e4; cudaEventCreate(&e4);
            wy_tsolve_launch(G, Tau, A, X, nb, cols, rows, false, 0);
            cudaEventRecord(e4); cudaEventSynchronize(e4);
            cudaEventElapsedTime(&ms_f, e2, e3); cudaEventElapsedTime(&ms_t, e3, e4);
        }
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
#ifdef QR_PROFILE
        qr_profile_dump(cols);
#endif
        printf("qr %dx%d x%d: full (explicit Q) %.1f us | WY factor %.1f us | tsolve %.1f us  (%s)\n", rows, cols, nb, ms * 1e3,
               ms_f * 1e3, ms_t * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
