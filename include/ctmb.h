/* libctmb -- C ABI of the B200-native CTMRG move engine.
 *
 * Drop-in boundary for the hot path of jurajHasik/peps-torch.  The
 * reference has no FFI: its "plugin interface" is the pair of Python modules
 * ctm/generic/ctmrg.py and ctm/one_site_c4v/ctmrg_c4v.py.  Every entry point below replaces the
 * reference function cited next to it; the Python host code (peps_torch_b200/) binds them with
 * ctypes and passes torch tensors as containers only (tensor.data_ptr(), current CUDA stream).
 *
 * Conventions
 *   - all data pointers are DEVICE pointers to contiguous row-major tensors of dtype
 *     CTMB_F64 (double) or CTMB_C128 (interleaved re,im doubles);
 *   - all outputs and the workspace are allocated by the caller; query the workspace size with
 *     the matching *_workspace function (same arguments, data pointers may be NULL);
 *   - `stream` is a cudaStream_t (0 = legacy default stream); calls are asynchronous w.r.t. the host
 *     except for the first call with a new shape, which builds and uploads contraction plans;
 *   - every function returns 0 on success; on failure a negative code, and ctmb_last_error() gives
 *     the message (no exception ever crosses the boundary);
 *   - index conventions are the reference's (ctm/generic/env.py:57-77): on-site a[s,u,l,d,r];
 *     C(-1,-1)[down,right] C(1,-1)[left,down] C(1,1)[up,left] C(-1,1)[up,right];
 *     T(0,-1)[chi,d,chi] T(-1,0)[chi,chi,d] T(0,1)[d,chi,chi] T(1,0)[chi,d,chi];
 *     fused double-layer legs are (ket,bra), ket-major.
 */
#ifndef CTMB_H
#define CTMB_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ctmb_handle_s* ctmb_handle_t;

typedef enum { CTMB_F64 = 0, CTMB_C128 = 1 } ctmb_dtype;
/* directions in the order of CTMARGS.ctm_move_sequence (config.py:392) */
typedef enum { CTMB_UP = 0, CTMB_LEFT = 1, CTMB_DOWN = 2, CTMB_RIGHT = 3 } ctmb_direction;
/* enlarged corners (ctm/generic/ctm_components.py:314-885) */
typedef enum { CTMB_LU = 0, CTMB_RU = 1, CTMB_RD = 2, CTMB_LD = 3 } ctmb_corner;

/* Hot-path knobs of CTMARGS (config.py:369-409) plus the parameters of the randomised
 * range finder that replaces the full LAPACK decomposition. */
typedef struct {
    double svd_reltol;         /* projector_svd_reltol        (1e-8)  ctm_projectors.py:266 */
    double eps_multiplet;      /* projector_eps_multiplet     (1e-8)  custom_svd.py:70-95; C4v: 1e-12 custom_eig.py:7-8 */
    double multiplet_abstol;   /* projector_multiplet_abstol  (1e-14) */
    double rsvd_rank_factor;   /* sketch width k = min(n, ceil(factor*chi)); default 0 = automatic: 1.75 for n <= 1024 (latency-bound
                                  QR / Jacobi steps), 2.0 above (bound by the n x n x k products); measured, profiles/r1_sketch_width.md */
    int rsvd_niter;            /* power iterations q of the first call with a new shape (later calls start from the count that
                                  satisfied the residual test last time); default 4 */
    int jacobi_max_sweeps;     /* default 40 */
    int norm_type;             /* ctm_absorb_normalization (ctmrg.py:210-230): 0 = 'inf' (max |x|, default), 1 = any other value of
                                  the reference = vector 2-norm (the C4v move scales C by |C[0,0]| either way, ctmrg_c4v.py:182-197) */
    int rsvd_max_rounds;       /* adaptive mode: at most this many rounds (default 5): a round that misses the bound adds a third of
                                  the count (near miss), the count again, or what the measured decay rate predicts; a round that
                                  gains less than 15 % per iteration ends the iteration (rounding floor).  A result that still
                                  misses the bound is returned AND counted: ctmb_get_rsvd_status */
    unsigned long long seed;   /* seed of the Gaussian sketch (deterministic) */
    double rsvd_tol;           /* > 0 (default 2e-15): residual-checked range finder.  After the power iterations
                                  max_j ||M v_j - s_j u_j|| / s_0 <= rsvd_tol * sqrt(n) is checked on the kept triplets (one host
                                  synchronisation per call) and further rounds are run if needed: the bound is a few hundred
                                  ulps, i.e. LAPACK-grade residuals, because the projectors amplify the error of a triplet by
                                  s_0 / s_j.  0 = fixed iteration count, no synchronisation. */
    int projector_method;      /* CTMARGS.projector_method: 0 = '4X4' (halves of the 4x4 network, ctm_projectors.py:14-64; default),
                                  1 = '4X2' (R, Rt = the two enlarged corners next to the bond, ctm_projectors.py:66-136) */
    int rsvd_stateless;        /* 0 (default): the handle remembers, per problem shape, the iteration count that passed the residual
                                  test last time (bisecting towards the largest count known to fail) and, per (direction, site)
                                  slot of the moves, the ordered Ritz vectors of the previous decomposition, which replace most of
                                  the Gaussian sketch (warm start: n x k elements per slot; a CTM run decomposes a slowly changing
                                  matrix once per move) -- only the NUMBER of iterations depends on it, every result passes the same
                                  residual test;  1: every call starts from a Gaussian sketch and rsvd_niter, so its result does not
                                  depend on the handle's history */
} ctmb_options;

/* One unit-cell site: on-site tensor and its eight environment tensors. */
typedef struct {
    const void* a;             /* a[p,Du,Dl,Dd,Dr]  -- or, with dims[0] == 0, the DOUBLE-LAYER tensor A[du,dl,dd,dr] whose legs
                                  are the fused (ket,bra) pairs: what ctmrg.run builds under ctm_force_dl (ctmrg.py:51-61),
                                  run_overlap (:137-147) and ctm_MOVE_dl (ctmrg_c4v.py:229-233) */
    int dims[5];               /* p, Du, Dl, Dd, Dr      (double-layer: 0, du, dl, dd, dr = the fused extents) */
    int pad;
    const void* C[4];          /* C(-1,-1), C(1,-1), C(1,1), C(-1,1)   -- all chi x chi */
    const void* T[4];          /* T(0,-1), T(-1,0), T(0,1), T(1,0) */
} ctmb_site;

int ctmb_version(void);
const char* ctmb_last_error(void);
/* device >= 0: a CUDA device of compute capability >= 10.0.  device == -1: planning-only handle -- the *_workspace
 * queries run (they validate every contraction of the call and size its workspace without launching anything),
 * every compute entry point fails with an error.  There is no CPU compute path. */
int ctmb_create(ctmb_handle_t* h, int device);
int ctmb_destroy(ctmb_handle_t h);
void ctmb_default_options(ctmb_options* opt);
/* Intra-site split over a group of GPUs (the "G = 2N" layout; yastn's device farm, _env_ctm_dist_mp.py:305-394,
 * is the reference's counterpart).  Every member of the group makes the SAME ctmb_move_generic_projectors call on the
 * same inputs; libctmb splits the n x n x k operator applications of the range finder by sketch columns and exchanges
 * the slabs through `fn`: an IN-PLACE all-gather over the group, buf = [nranks][bytes_per_rank] with this rank's slot
 * filled, ordered with `stream` (the host implements it with NCCL over NVLink, torch.distributed.all_gather_into_tensor).
 * QR / Jacobi / truncation run redundantly on every member, which therefore all return identical projectors.
 * nranks = 1 switches the group off. */
typedef int (*ctmb_allgather_fn)(void* ctx, void* buf, size_t bytes_per_rank, void* stream);
int ctmb_set_group(ctmb_handle_t h, int rank, int nranks, ctmb_allgather_fn fn, void* ctx);

/* Status of the residual-checked range finder since the last reset: number of residual checks, number of decompositions
 * that were RETURNED ALTHOUGH they missed the bound rsvd_tol * sqrt(n) (the residual stopped improving -- rounding floor --
 * or rsvd_max_rounds was reached) and the worst residual / bound ratio among those.  The host wrappers turn missed > 0 into
 * a Python warning; calls / iterations: iterative decompositions run and power iterations they took in total.
 * (The reference's LAPACK SVD has no such diagnostic: svd_gesdd.py:77-96.) */
int ctmb_get_rsvd_status(ctmb_handle_t h, long long* checks, long long* missed, double* worst_ratio, long long* calls,
                         long long* iterations, int reset);

/* kernels launched / algorithmic real flops enqueued by this handle since the last reset */
int ctmb_get_counters(ctmb_handle_t h, long long* launches, double* flops);
int ctmb_reset_counters(ctmb_handle_t h);
/* Per-kernel-class device timing with CUDA events on the launching stream (used by bench.py for
 * the roofline entry).  Classes: 0 tensor-contraction GEMM, 1 Householder QR, 2 Jacobi, 3 misc.
 * ctmb_profile_get synchronises the recorded events and fills four-element arrays with the
 * totals since the last ctmb_reset_counters: milliseconds, algorithmic flops, algorithmic bytes,
 * launches. */
int ctmb_profile_enable(ctmb_handle_t h, int on);
int ctmb_profile_get(ctmb_handle_t h, double* ms, double* flops, double* bytes, long long* launches);

/* Pairwise tensor contraction "ab,buc->auc" of contiguous tensors (tn_interface.py:3-10
 * contract/einsum/mm).  Labels: single letters; output labels come from exactly one operand. */
int ctmb_einsum2(ctmb_handle_t h, ctmb_dtype dt, const char* spec,
                 const void* A, const long long* dimsA, int conjA,
                 const void* B, const long long* dimsB, int conjB, void* C, void* stream);

/* Enlarged corner c2x2_{LU,RU,RD,LD}_sl_c (ctm_components.py:372-434,532-586,683-733,832-885):
 * out is the (chi*D^2) x (chi*D^2) matrix of SURVEY Appendix A for that corner of `site`. */
int ctmb_c2x2(ctmb_handle_t h, ctmb_dtype dt, ctmb_corner kind, int chi, const ctmb_site* site,
              void* out, void* ws, size_t ws_bytes, void* stream);
size_t ctmb_c2x2_workspace(ctmb_handle_t h, ctmb_dtype dt, ctmb_corner kind, int chi, const ctmb_site* site);

/* halves_of_4x4_CTM_MOVE_{UP,LEFT,DOWN,RIGHT}_c (ctm_components.py:55-75,123-139,186-201,249-265):
 * corners[4] are the four sites of the 2x2 patch in the order the reference lists them
 * (UP: RU,RD,LU,LD  LEFT: LU,RU,LD,RD  DOWN: LD,LU,RD,RU  RIGHT: RD,LD,RU,LU). R, Rt: n x n. */
int ctmb_halves(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int chi, const ctmb_site* const corners[4],
                void* R, void* Rt, void* ws, size_t ws_bytes, void* stream);
size_t ctmb_halves_workspace(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int chi,
                             const ctmb_site* const corners[4]);

/* ctm_get_projectors_from_matrices (ctm_projectors.py:142-293): R, Rt are n0 x n1 row-major;
 * P, Pt are n0 x chi row-major; S_out (may be NULL) receives the chi truncated singular values. */
int ctmb_projectors(ctmb_handle_t h, ctmb_dtype dt, const void* R, const void* Rt, int n0, int n1, int chi,
                    const ctmb_options* opt, void* P, void* Pt, double* S_out,
                    void* ws, size_t ws_bytes, void* stream);
size_t ctmb_projectors_workspace(ctmb_handle_t h, ctmb_dtype dt, int n0, int n1, int chi, const ctmb_options* opt);

/* truncated_svd_gesdd with keep_multiplets (custom_svd.py:38-101) + fix_svd_signs
 * (svd_gesdd.py:18-26): M is m x n row-major; U (m x chi) and V (n x chi) are written
 * COLUMN-major (each singular vector contiguous); S has chi entries. */
int ctmb_truncated_svd(ctmb_handle_t h, ctmb_dtype dt, const void* M, int m, int n, int chi,
                       const ctmb_options* opt, void* U, double* S, void* V,
                       void* ws, size_t ws_bytes, void* stream);
size_t ctmb_truncated_svd_workspace(ctmb_handle_t h, ctmb_dtype dt, int m, int n, int chi, const ctmb_options* opt);

/* truncated_eig_sym with keep_multiplets (custom_eig.py:7-67): M Hermitian n x n; D: chi
 * eigenvalues of largest magnitude (descending |.|); U (n x chi) column-major. */
int ctmb_truncated_eig_sym(ctmb_handle_t h, ctmb_dtype dt, const void* M, int n, int chi,
                           const ctmb_options* opt, double* D, void* U,
                           void* ws, size_t ws_bytes, void* stream);
size_t ctmb_truncated_eig_sym_workspace(ctmb_handle_t h, ctmb_dtype dt, int n, int chi, const ctmb_options* opt);

/* Thin Householder QR = torch.linalg.qr(M) as used by ctm_MOVE_QR_sl (ctm/one_site_c4v/ctmrg_c4v.py:513-516, the
 * projector of the QR variant of the C4v move).  A: rows x k, COLUMN-major (= the transpose of a row-major torch matrix),
 * rows >= k; on return A holds the explicit thin Q, R (k x k column-major) the upper-triangular factor.  Signs follow
 * LAPACK's Householder convention (R_jj = -sign(alpha) ||x||), i.e. Q and R equal torch's up to rounding.  One difference:
 * for a SQUARE complex matrix LAPACK also rotates the last diagonal entry of R to the real axis; here that entry keeps its
 * phase (the last column has nothing below the diagonal and gets no reflector).  Tall matrices -- every use on the path -- are
 * not affected. */
int ctmb_qr(ctmb_handle_t h, ctmb_dtype dt, void* A, int rows, int k, void* R, void* ws, size_t ws_bytes, void* stream);
size_t ctmb_qr_workspace(ctmb_handle_t h, ctmb_dtype dt, int rows, int k);

/* One directional move over all sites = ctm_MOVE (ctm/generic/ctmrg.py:179-319) including
 * projectors (4X4), absorb_truncate_CTM_MOVE_* (ctmrg.py:324-804) and move_normalize_c.
 *   sites[nsites]        the unit cell with its current environment
 *   corner_site[4*i+j]   index into sites[] of the j-th corner of the 2x2 patch of job i
 *                        (order as in ctmb_halves)
 *   nb_site[i]           index of the site whose projectors P1,Pt1 job i uses (coord+shift)
 *   nC1/nC2/nT[i]        outputs of job i (chi x chi, chi x chi, T-shaped for `dir`); the host
 *                        stores them at site coord-direction (ctmrg.py:313-319). */
int ctmb_move_generic(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int nsites, int chi,
                      const ctmb_site* sites, const int* corner_site, const int* nb_site,
                      const ctmb_options* opt, void* const* nC1, void* const* nC2, void* const* nT,
                      void* ws, size_t ws_bytes, void* stream);
size_t ctmb_move_generic_workspace(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int nsites, int chi,
                                   const ctmb_site* sites, const int* corner_site, const int* nb_site,
                                   const ctmb_options* opt);

/* Restricted forms used by the multi-GPU shard (each rank owns a subset of the jobs):
 * projectors of the listed jobs only, and absorption of the listed jobs with given projectors. */
int ctmb_move_generic_projectors(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int nsites, int chi,
                                 const ctmb_site* sites, const int* corner_site, int njobs, const int* jobs,
                                 const ctmb_options* opt, void* const* P, void* const* Pt,
                                 void* ws, size_t ws_bytes, void* stream);
int ctmb_move_generic_absorb(ctmb_handle_t h, ctmb_dtype dt, ctmb_direction dir, int nsites, int chi,
                             const ctmb_site* sites, const int* nb_site, int njobs, const int* jobs,
                             const ctmb_options* opt, const void* const* P, const void* const* Pt,
                             void* const* nC1, void* const* nC2, void* const* nT,
                             void* ws, size_t ws_bytes, void* stream);

/* One C4v move = ctm_MOVE_sl (ctm/one_site_c4v/ctmrg_c4v.py:325-463): a[p,D,D,D,D], C[chi,chi],
 * T[chi,chi,D^2] -> C_out, T_out (same shapes); D_out (may be NULL): the chi kept eigenvalues.
 * With dims[0] == 0 `a` is the double-layer tensor A[D^2,D^2,D^2,D^2]: ctm_MOVE_dl (:200-322). */
int ctmb_move_c4v(ctmb_handle_t h, ctmb_dtype dt, const void* a, const int dims[5], const void* C,
                  const void* T, int chi, const ctmb_options* opt, void* C_out, void* T_out, double* D_out,
                  void* ws, size_t ws_bytes, void* stream);
size_t ctmb_move_c4v_workspace(ctmb_handle_t h, ctmb_dtype dt, const int dims[5], int chi, const ctmb_options* opt);

/* ---- observables on the converged environment --------------------------------------------------
 * rdm2x2 (ctm/generic/rdm.py:1306-1592; contraction strategy of rdm2x2_legacy :1362): UN-normalised reduced density
 * matrix of the 2x2 plaquette  s0 s1 / s2 s3.  sites[4] = the unit-cell sites (with their environment tensors) at
 * coord, coord+(1,0), coord+(0,1), coord+(1,1); bit q of open_mask keeps site q open (`open_sites` of the reference),
 * the others are traced.  rho: row-major [kets of the open sites in site order..., bras in the same order...]. */
int ctmb_rdm2x2(ctmb_handle_t h, ctmb_dtype dt, int chi, const ctmb_site* const sites[4], int open_mask, void* rho,
                void* ws, size_t ws_bytes, void* stream);
size_t ctmb_rdm2x2_workspace(ctmb_handle_t h, ctmb_dtype dt, int chi, const ctmb_site* const sites[4], int open_mask);

/* rdm1x1 / rdm2x1 / rdm1x2 (ctm/generic/rdm.py:114-258, 352-500, 672-826; C4v: rdm1x1(_sl), rdm2x1(_sl) of
 * ctm/one_site_c4v/rdm_c4v.py:168-392, 394-665 on the rotated environment): UN-normalised one- and two-site density
 * matrices.  kind 0: sites[0] = site at coord, rho[s;s'];  kind 1: sites = coord, coord+(1,0), rho[s0,s1;s0',s1'];
 * kind 2: sites = coord, coord+(0,1), rho[s0,s1;s0',s1']. */
int ctmb_rdm_small(ctmb_handle_t h, ctmb_dtype dt, int kind, int chi, const ctmb_site* const sites[2], void* rho,
                   void* ws, size_t ws_bytes, void* stream);
size_t ctmb_rdm_small_workspace(ctmb_handle_t h, ctmb_dtype dt, int kind, int chi, const ctmb_site* const sites[2]);

/* _sym_pos_def_matrix (ctm/generic/rdm.py:38-57): out = (rdm + rdm^H)/2; with sym_pos_def != 0 and a negative
 * eigenvalue, out = U max(D,0) U^H; finally out /= Re tr(out).  rdm, out: n x n row-major (may not alias). */
int ctmb_sym_pos_def(ctmb_handle_t h, ctmb_dtype dt, const void* rdm, int n, int sym_pos_def, void* out,
                     void* ws, size_t ws_bytes, void* stream);
size_t ctmb_sym_pos_def_workspace(ctmb_handle_t h, ctmb_dtype dt, int n, int sym_pos_def);

#ifdef __cplusplus
}
#endif
#endif /* CTMB_H */
