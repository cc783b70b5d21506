// Stand-alone timing of qr_launch on 4 matrices of 432x96 (config c2) with per-phase cycle counters.
#include "../../peps_torch_b200/csrc/common.h"
#include <vector>
#include <cstdlib>
namespace ctmb { extern __device__ long long g_qr_prof[16]; }
using namespace ctmb;
int main(int argc, char** argv) {
    int rows = argc > 1 ? atoi(argv[1]) : 432, cols = argc > 2 ? atoi(argv[2]) : 96, nb = 4;
    std::vector<double> h((size_t)rows * cols);
    PtrBatch A{}, R{};
    for (int b = 0; b < nb; ++b) {
        for (auto& x : h) x = rand() / (double)RAND_MAX - 0.5;
        cudaMalloc(&A.p[b], h.size() * 8);
        cudaMemcpy(A.p[b], h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        for (int b = 0; b < nb; ++b) cudaMemcpy(A.p[b], h.data(), h.size() * 8, cudaMemcpyHostToDevice);
        long long z[16] = {0};
        cudaMemcpyToSymbol(g_qr_prof, z, sizeof z);
        cudaEventRecord(e0);
        qr_launch(A, R, nb, rows, cols, rows, false, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpyFromSymbol(z, g_qr_prof, sizeof z);
        printf("qr %dx%d x%d: %.1f us | per-step cycles: bcast+sync %lld dots+fold %lld csync %lld gather %lld scalar %lld update %lld other %lld\n", rows, cols, nb, ms * 1e3,
               z[1] / cols, z[2] / cols, z[3] / cols, z[4] / cols, z[5] / cols, z[6] / cols, z[0] / cols);
    }
    return 0;
}
