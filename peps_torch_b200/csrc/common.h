// Shared declarations of libctmb (B200 / sm_100a CTMRG engine).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>
#include <algorithm>

namespace ctmb {

void set_error(const std::string& msg);

#define CTMB_CHECK(cond, msg)                                                       \
    do {                                                                            \
        if (!(cond)) throw std::runtime_error(std::string(msg) + " [" #cond "] at " \
                                              __FILE__ ":" + std::to_string(__LINE__)); \
    } while (0)

#define CTMB_CUDA(call)                                                             \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) throw std::runtime_error(std::string("CUDA error ") + \
            cudaGetErrorString(e__) + " in " #call " at " __FILE__ ":" + std::to_string(__LINE__)); \
    } while (0)

// ---------------------------------------------------------------------------------
// tensor-contraction GEMM  C[m,n] = alpha * sum_k opA(A)[m,k] * opB(B)[k,n]
// element (m,k) of A lives at A + a_m[m] + a_k[k] (element offsets), same for B and C.
// ---------------------------------------------------------------------------------
struct TcTables {
    const int* a_m; const int* a_k;
    const int* b_k; const int* b_n;
    const int* c_m; const int* c_n;
    // lin != 0: A and B are plain strided matrices, a_m[m] = m * a_sm, a_k[k] = k * a_sk, b_k[k] = k * b_sk, b_n[n] = n * b_sn
    // (element strides) -- what the TMA-fed kernel needs to describe them by tensor maps (tc_gemm_tma.cu)
    int lin, a_sm, a_sk, b_sk, b_sn;
    int c_lin, c_sm, c_sn;      // c_lin != 0: C is a plain strided matrix as well, c_m[m] = m * c_sm, c_n[n] = n * c_sn
};

struct TcBatchEntry {
    const void* A; const void* B; void* C;
    unsigned long long* amax;   // optional: atomicMax of |C| (bits of a non-negative double)
    int tab;                    // which table set
    int flags;                  // bit0 conjA, bit1 conjB, bit2 A loader k-fast, bit3 B loader k-fast
};

constexpr int TC_MAX_BATCH = 32;
constexpr int TC_MAX_TABS = 8;

struct TcParams {
    int M, N, K;
    int nbatch;
    int ksplit;                 // > 1: split-K, slice y of the grid reduces k in [y*kchunk, (y+1)*kchunk) into `partial`
    int kchunk;
    int pad;
    void* partial;              // [nbatch][ksplit][M][N] dense partial products (summed by tc_splitk_reduce_launch)
    double alpha;
    TcTables tab[TC_MAX_TABS];
    TcBatchEntry batch[TC_MAX_BATCH];
};

enum { TC_CONJ_A = 1, TC_CONJ_B = 2, TC_A_KFAST = 4, TC_B_KFAST = 8, TC_ACCUM = 16 /* C += alpha A B */ };

// launches one batched contraction; cplx=false: double, cplx=true: double2 (re,im)
void tc_launch(const TcParams& p, bool cplx, cudaStream_t stream);
// tile shape tc_launch will use for this problem (for the split-K decision of the caller)
void tc_tile_shape(const TcParams& p, bool cplx, int& bm, int& bn);
// C[c_m[m] + c_n[n]] (+)= alpha * sum_s partial[z][s][m][n], optional max|.|
void tc_splitk_reduce_launch(const TcParams& p, bool cplx, cudaStream_t stream);
void tc_init_attributes();
// TMA-fed variant of the 128 x 128 x 16 warp-specialised kernel; false (nothing launched) if the batch does not qualify
bool tc_run_tma(const TcParams& p, bool a_kfast, bool b_kfast, cudaStream_t stream);

// ---------------------------------------------------------------------------------
// dense helpers (qr.cu, jacobi.cu, misc.cu) -- all batched over `nb` problems
// matrices are column-major with leading dimension ld unless stated otherwise
// ---------------------------------------------------------------------------------
struct PtrBatch { void* p[TC_MAX_BATCH]; };

// in-place Householder QR of (rows x cols) column-major matrices; on exit the matrix holds
// the explicit thin Q; if Rout.p[i] != nullptr the cols x cols upper-triangular R is
// written there (column-major, ld = cols).
void qr_launch(const PtrBatch& A, const PtrBatch& Rout, int nb, int rows, int cols, int ld,
               bool cplx, cudaStream_t stream);

// WY form (fast path, sketches up to 1024 x 128): factor leaves V / tau / R; the explicit Q is
// then  Q = E - V X,  X = T V1^H  (wy_tsolve) with the Gram matrix G = V^H V from a GEMM.
bool qr_wy_supported(int rows, int cols, bool cplx);
void qr_wy_factor_launch(const PtrBatch& A, const PtrBatch& Rout, const PtrBatch& Tau, int nb, int rows, int cols,
                         int ld, bool cplx, cudaStream_t stream);
void wy_tsolve_launch(const PtrBatch& G, const PtrBatch& Tau, const PtrBatch& V, const PtrBatch& X, int nb, int k,
                      int ldv, bool cplx, cudaStream_t stream);
void add_identity_launch(const PtrBatch& Q, int nb, int k, int ld, bool cplx, cudaStream_t stream);
// blocked factorisation of sketches too large for one cluster (see qr.cu): panel width, panel factorisation in
// WY form, extraction of the rows of R belonging to a panel, Q <- E.  wy_tsolve_launch with ldv <= 0 returns T
// itself (X = T, stored [c][s]) instead of T V1^H.
int qr_panel_width(int rows, int cols, bool cplx);
void qr_panel_launch(const PtrBatch& A, const PtrBatch& Rpp, const PtrBatch& Tau, int nb, int rows, int b, int ld,
                     bool cplx, cudaStream_t stream);
// tall panels (qr_tall.cu): rows split over up to 148 co-resident CTAs, 32 columns per leaf; width 0 = does not apply
int qr_tall_panel_width(int nb, int rows, int cols, bool cplx);
size_t qr_tall_scratch_bytes(int nb);
int qr_tall_mode();
bool qr_tall_panel_launch(const PtrBatch& A, const PtrBatch& Rpp, const PtrBatch& Tau, int nb, int rows, int b, int ld, bool cplx,
                          void* scratch, cudaStream_t stream);
void qr_copy_r_launch(const PtrBatch& A, const PtrBatch& Rpp, const PtrBatch& R, int nb, int k, int j0, int b, int ld,
                      bool cplx, cudaStream_t stream);
void set_identity_launch(const PtrBatch& Q, int nb, int rows, int k, bool cplx, cudaStream_t stream);
void qr_copy_v_launch(const PtrBatch& A, const PtrBatch& V, int nb, int rows, int J0, int j0, int bw, bool cplx,
                      cudaStream_t stream);

// one-sided Jacobi SVD of k x k column-major G (ld=k): on exit G = Uhat*Sigma (columns
// orthogonal), W accumulates the right rotations (G_in * W = G_out), sig[k] the column norms.
void jacobi_launch(const PtrBatch& G, const PtrBatch& W, const PtrBatch& sig, int nb, int k,
                   bool cplx, int max_sweeps, int shift, int transpose_in, cudaStream_t stream);
size_t jacobi_smem_limit();

// sort sig descending -> Ssorted[k]; Uhs = normalised sorted columns of G (first ncol),
// Ws = sorted columns of W (first ncol); both k x ncol column-major.
void sortcols_launch(const PtrBatch& G, const PtrBatch& W, const PtrBatch& sig, const PtrBatch& Ssorted,
                     const PtrBatch& Uhs, const PtrBatch& Ws, int nb, int k, int ncol, bool cplx,
                     int eig_mode, cudaStream_t stream);

struct ProjFinalizeArgs {
    int nb, rowsU, rowsV, chi, kavail;   // kavail: number of valid entries of S (>= chi+1 if truncating)
    double reltol, eps_multiplet, abstol;
    int truncating;                       // chi < min(m,n): multiplet rule applies
    int conj_u;                           // write conj(U)*phase*s into Uout (projector form)
    int apply_scale;                      // multiply by S^-1/2 (else plain sign-fixed U,V)
    int v_div_sigma;                      // V holds M^H U: divide column j by S_j first (0 if S_j is below the cut)
};
// U (rowsU x chi), V (rowsV x chi) column-major in place -> Uout/Vout; Sout[chi] truncated spectrum
void proj_finalize_launch(const PtrBatch& U, const PtrBatch& V, const PtrBatch& S, const PtrBatch& Sout,
                          const ProjFinalizeArgs& a, bool cplx, cudaStream_t stream);

// x /= amax (amax given as bits); count elements each
struct ScaleBatch { void* p[TC_MAX_BATCH]; const unsigned long long* amax[TC_MAX_BATCH]; long long count[TC_MAX_BATCH]; };
// sqrt_mode: the slot holds sum |x|^2 (sumsq_launch) instead of max |x|
void scale_by_amax_launch(const ScaleBatch& b, int nb, bool cplx, cudaStream_t stream, bool sqrt_mode = false);
// slot (a double, zeroed by the caller) += sum |x|^2
void sumsq_launch(const ScaleBatch& b, int nb, bool cplx, cudaStream_t stream);
// reduced density matrices (ctm/generic/rdm.py:38-57): out = (raw + raw^H)/2 [/ Re tr raw]; positive part from eigenpairs
void conj_inplace_launch(void* x, long long count, cudaStream_t stream);     // complex only: x <- conj(x)
void rdm_herm_launch(const void* raw, void* out, int n, int normalize, bool cplx, cudaStream_t stream);
void rdm_posdef_launch(void* out, const void* U, const double* D, int n, bool cplx, cudaStream_t stream);
// amax of |x|
void absmax_launch(const ScaleBatch& b, int nb, bool cplx, cudaStream_t stream);

void resid_launch(const PtrBatch& MX, const PtrBatch& Y, const PtrBatch& S, int nb, int rows, int chi, int k, double reltol,
                  unsigned long long* out, bool cplx, cudaStream_t stream);
void c4v_sym_launch(const void* tin, void* tout, int chi, int d, unsigned long long* amax, bool cplx,
                    cudaStream_t stream);
void c4v_diag_launch(const double* D, void* cout, int chi, bool cplx, cudaStream_t stream);

void fill_gaussian_launch(double* out, long long count, unsigned long long seed, cudaStream_t stream);
void add_noise_launch(double* x, long long count, double amp, const unsigned long long* scale, unsigned long long seed,
                      cudaStream_t stream);
void shift_axpy_launch(double* y, const double* x, const unsigned long long* sumsq, double sign, long long count, cudaStream_t stream);
void axpby_launch(double* out, const double* x, double c1, const double* y, double c2, long long count, cudaStream_t stream);

// fused double-layer absorption of the enlarged corner (dl_fused.cu)
struct DlParams {
    const double* X;            // [npairs][(K1,K2)][(k1,k2)]: C.T1.T2 per pair of environment indices, bra-major
    const double* a;            // on-site tensor a[s,u,l,d,r]
    double* out;                // corner matrix
    int npairs, n2;             // pair = i1 * n2 + i2
    long long st1, st2;         // element strides of the two environment indices in `out`
    int as_s, as_k1, as_k2, as_o1, as_o2;   // element strides of a for (s, contracted legs, open legs)
    int ro[64], co[64];         // offsets of the ket (o1,o2) and bra (O1,O2) open pairs in `out`
};
bool dl_corner_supported(int dk1, int dk2, int do1, int do2, int pdim, bool cplx);
void dl_corner_launch(const DlParams& p, int dk, int dopen, int pdim, cudaStream_t stream);

}  // namespace ctmb
