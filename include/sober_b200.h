/*
 * sober_b200 -- C ABI of the B200 (sm_100a) implementation of SOBER's RCHQ batch-selection hot path.
 *
 * Plain C: raw DEVICE pointers, sizes and a cudaStream_t (passed as void*).  No torch types, no
 * allocation, no exceptions, no Python.  Every function returns SOBER_OK (0) or an error code and only
 * ENQUEUES work on `stream` (no host synchronisation) unless stated otherwise.
 *
 * The reference is pure Python/PyTorch (no FFI of its own), so each entry point cites the lines of the
 * reference it replaces; INTEGRATION.md shows the ctypes binding and how `SOBER/_rchq.py` is rebound.
 *
 * All floating-point data is IEEE binary64 ("f64"); alive-lists are int32 row ids (N < 2^31).
 * Matrices are row-major.
 */
#ifndef SOBER_B200_H
#define SOBER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOBER_B200_ABI_VERSION 2   /* 2: car_panel, gp_rows, popc_probe added; sober_group_args grew kx / ldkx / aw / n_obs / transform at its end */

enum sober_status {
    SOBER_OK = 0,
    SOBER_ERR_ARG = 1,         /* inconsistent sizes / null pointers */
    SOBER_ERR_CUDA = 2,        /* a CUDA runtime call failed (see sober_last_cuda_error) */
    SOBER_ERR_UNSUPPORTED = 3, /* shape or family outside what the kernels cover */
    SOBER_ERR_WORKSPACE = 4    /* workspace too small */
};

/* Kernel families: the base kernels reachable through SOBER/_kernel.py:16-30 in the examples --
 * gpytorch ScaleKernel(RBFKernel | MaternKernel(nu)) and SOBER/_drug_modelling.py:15-25,86-101. */
enum sober_family {
    SOBER_RBF = 0,      /* exp(-d2/2)                                  d2 = |u-v|^2, u=(x-c)/l            */
    SOBER_MATERN12 = 1, /* exp(-r)                                     r  = sqrt(max(d2,1e-30))           */
    SOBER_MATERN32 = 2, /* (1+sqrt3 r) exp(-sqrt3 r)                                                      */
    SOBER_MATERN52 = 3, /* (1+sqrt5 r+5/3 r^2) exp(-sqrt5 r)                                              */
    SOBER_TANIMOTO = 4, /* max(0,(<x,z>+1e-6)/(1e-6+|x|^2+|z|^2-<x,z>))                                   */
    SOBER_TANIMOTO_BITS = 5, /* the same on bit-packed {0,1} rows: <x,z> = popcount(x & z)  (sober_pack_bits)  */
    SOBER_HAMMING_LUT = 6    /* any stationary kernel with one lengthscale on bit-packed {0,1} rows: the squared
                                distance is the Hamming distance h = popcount(x ^ z) times a constant, so
                                k = lut[h], lut (d + 1 doubles) supplied by the caller (examples/ising.py:
                                ScaleKernel(RBFKernel) on {0,1}^24)                                              */
};

int sober_abi_version(void);
/* Text of the last CUDA error seen by this library on the calling thread ("" if none). */
const char* sober_last_cuda_error(void);
/* Number of SMs of the current device (grid sizing on the host side). */
int sober_sm_count(int* out);

/* ---------------------------------------------------------------------------------------------------
 * Streaming preparation passes (HBM-bound).
 * ------------------------------------------------------------------------------------------------- */

/* Stationary families.  P[i, k] = (X[i, k] - center[k]) * inv_ls[k] for k < d and P[i, d] = sum_k P[i,k]^2.
 * `ldp` >= d + 1 (the pad, if any, is zeroed).  Replaces the per-call `x1.div(lengthscale)` / centring of
 * gpytorch's RBF/Matern forward reached from SOBER/_rchq.py:124.  X: n x d, row stride ldx. */
int sober_prepare_points(const double* X, int64_t ldx, int64_t n, int32_t d, const double* center,
                         const double* inv_ls, double* P, int64_t ldp, void* stream);

/* Small-d record layout for the register kernel of K1 (d <= 8, stationary families): one row per ALIVE point,
 * in alive-list order,   rec[j] = [ (X[idx[j]] - center) * inv_ls  (d) | its squared norm | mu[j] | zero pad ],
 * row stride ldr = d + 2 rounded up to even (16-byte aligned rows: the kernel stages them with 1-D TMA bulk
 * copies).  idx NULL = identity, mu NULL = 1.  Fuses the gather `samp[idx]`, the lengthscale division and the
 * norm of gpytorch's distance (SOBER/_rchq.py:124 -> covar_module.forward). */
int sober_make_records(const double* X, int64_t ldx, int32_t d, const double* center, const double* inv_ls,
                       const int32_t* idx, const double* mu, int64_t m, double* rec, int64_t ldr, void* stream);

/* Bit-packing of {0,1}-valued rows (fingerprints, SOBER/_drug_modelling.py; the reference stores them as floats):
 * words[i, w] bit b = (X[i, 64 w + b] != 0), ldw >= ceil(d / 64) words per row (padding words zeroed; K1 wants ldw in
 * {1,2,4,8,16,32}); popc[i] = number of set bits (= |x|^2);
 * *not_binary (device int32, must be zeroed by the caller) is set to 1 if any entry is neither 0 nor 1.
 * One streaming pass over X; afterwards family SOBER_TANIMOTO_BITS evaluates <x,z> as popcount(x & z). */
int sober_pack_bits(const double* X, int64_t ldx, int64_t n, int32_t d, uint64_t* words, int32_t ldw, double* popc,
                    int32_t* not_binary, void* stream);

/* out[i] = sum_k X[i,k]^2  (Tanimoto |x|^2, SOBER/_drug_modelling.py:21-22). */
int sober_row_sqnorm(const double* X, int64_t ldx, int64_t n, int32_t d, double* out, void* stream);

/* Stable compaction of the non-zero weights: `idx_story = arange(N)[mu != 0]` (SOBER/_rchq.py:63-65).
 * Writes ascending row ids to idx_out, their weights to mu_out and the count to *count_out (device).
 * workspace: at least sober_compact_workspace(n) bytes. */
int64_t sober_compact_workspace(int64_t n);
int sober_compact_nonzero(const double* mu, int64_t n, int32_t* idx_out, double* mu_out, int64_t* count_out,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K1: fused cross-kernel + weighted strided group sums  (SOBER/_rchq.py:116-136,152; also :35,:78 as a
 * plain Gram when S = number of points and a single row).
 *
 *   At[g, l]  = outputscale * sum_{rows e} k(z_l, x_{e*S+g}) * mu_{e*S+g}       over ALL valid positions
 *   totw[g]   =               sum_{rows e, e*S+g < ES} mu_{e*S+g}
 *
 * A "position" p is the rank of a point in the global ascending alive-list; this device owns positions
 * [pos0, pos0 + n_local).  Position p belongs to group p mod S.  Positions >= ES (the remainder of
 * SOBER/_rchq.py:128-136) are accumulated into At (column p - ES, "the quirk") but not into totw.
 * Never materialises the (E, L, S) Gram of SOBER/_rchq.py:124.
 *
 * Two candidate layouts:
 *  - indexed (any d): X / ldx / xn / xn_stride = candidate rows and their squared norms (stationary families: the
 *    output of sober_prepare_points, xn = P + d, xn_stride = ldp; Tanimoto: the raw rows and sober_row_sqnorm);
 *    idx = local alive-list (row ids into X), NULL = identity; mu = weights aligned with idx, NULL = 1.
 *  - records (d <= 8): rec / ldr from sober_make_records, already in alive-list order (X, xn, idx, mu unused).
 * Stationary families expect coordinates pre-multiplied by the family constant (1/sqrt2 RBF, 1 Matern-1/2,
 * sqrt3 Matern-3/2, sqrt5 Matern-5/2): fold it into inv_ls.
 * Zt (L x d): landmark table -- stationary: -2 (z - c) * inv_ls ; Tanimoto: z.   zn (L): |.|^2 of the same.
 * SOBER_TANIMOTO_BITS / SOBER_HAMMING_LUT (indexed layout only): X and Zt point to uint64 word rows from
 *   sober_pack_bits (ldx = words per row, d = number of BITS), xn / zn are the popcounts as doubles (unused by the
 *   LUT family), lut as described at the enum.
 * At: S x L (transposed on purpose: coalesced stores and it is the left operand of the projection).
 * workspace holds the per-split partial sums (deterministic two-stage reduction, no atomics).
 * variant: 0 = automatic (records -> register kernel, else tiled; SOBER_TANIMOTO_BITS with 256..1024-bit rows and at least
 *          2^20 pairs -> the tcgen05 kernel of csrc/group_bits_mma.cu), 1 = force the generic tiled kernel,
 *          4 = force the popcount kernel for SOBER_TANIMOTO_BITS, 5 = the first tcgen05 kernel (landmark tile in shared
 *          memory instead of TMEM; kept for comparison).
 * ------------------------------------------------------------------------------------------------- */
typedef struct sober_group_args {
    const double* X;
    int64_t ldx;
    const double* xn;
    int64_t xn_stride;
    const int32_t* idx;
    const double* mu;
    int64_t n_local;
    int64_t pos0;
    int64_t n_global;
    int64_t ES;
    int32_t S;
    int32_t L;
    int32_t d;
    int32_t family;
    double outputscale;
    const double* Zt;
    const double* zn;
    double* At;
    double* totw;
    int32_t variant;
    int32_t unit_weights; /* record layout only: ignore the weight stored in the records (plain Gram) */
    const double* rec;    /* record layout (sober_make_records), row j = local position j; NULL = indexed layout */
    int64_t ldr;
    const double* lut;    /* SOBER_HAMMING_LUT: d + 1 kernel values indexed by the Hamming distance */
    /* Non-linear posterior mode (indexed layout only; all NULL / 0 otherwise) -- SOBER/BASQ/_scale_mmlt.py:256-275:
     *   At[g, l] = sum k_post(z_l, x) mu,  k_post = expm1( outputscale * k(z_l, x) - <aw[l, :], kx[row(x), :]> )
     * kx: (N x n_obs, ldkx) rows k(x, X_obs) addressed by row id like X;  aw: (L x n_obs) rows k(z_l, X_obs) W. */
    const double* kx;
    int64_t ldkx;
    const double* aw;
    int32_t n_obs;
    int32_t transform;    /* reserved: 1 = expm1 */
} sober_group_args;

int64_t sober_group_accumulate_workspace(const sober_group_args* args);
int sober_group_accumulate(const sober_group_args* args, void* workspace, int64_t workspace_bytes, void* stream);

/* Same reduction when the Gram tile has been produced by an opaque callable (generic-kernel path):
 *   At[g, l] += sum_{j : (pos_begin + j) mod S == g} G[l, j] * mu[j],   totw likewise for positions < ES.
 * G: L x m row-major (ldg).  (SOBER/_rchq.py:124-126 with `kernel` a black box.) */
int sober_group_accumulate_gram(const double* G, int64_t ldg, int32_t L, int64_t m, const double* mu,
                                int64_t pos_begin, int64_t ES, int32_t S, double* At, double* totw,
                                void* stream);

/* ---------------------------------------------------------------------------------------------------
 * CAR elimination (SOBER/_rchq.py:237-266) on a given null-space basis.
 *   basis: k x S row-major, row c = column c of Phi (i.e. `Vh[-(N-n):, :]` as torch returns it).
 *          DESTROYED (used as the publish buffer).
 *   mu:    S weights, updated in place to the reduced measure (zeros at eliminated positions).
 *   pivots_out (k int32, may be NULL): eliminated position per step, -1 after an early stop.
 *   steps_out (1 int32, may be NULL): number of steps taken.
 * exact != 0: arithmetic order is the reference's (unfused mul / div / sub), so given the same basis the pivots
 * are bit-identical.  exact == 0: one division per row (t_i = v_i / v_j) and an FMA per element (4x faster at
 * S = 2000, differs from the reference only in rounding).  Persistent cooperative kernel: sync_ws needs
 * sober_car_workspace(k) bytes (zeroed by the call).
 * ------------------------------------------------------------------------------------------------- */
int64_t sober_car_workspace(int32_t k);
int sober_car_eliminate(double* basis, int32_t k, int32_t S, double* mu, int32_t exact, int32_t* pivots_out,
                        int32_t* steps_out, void* sync_ws, int64_t sync_ws_bytes, void* stream);

/* The whole Caratheodory reduction (SOBER/_rchq.py:224-270) in ONE kernel on one thread-block cluster, state
 * resident in distributed shared memory (sizes up to roughly S <= 440 on 8 CTAs, S <= 630 on 16):
 *   basis == NULL: Householder QR of design (S x np, row-major, = [1 | X]) -> null-space basis -> elimination
 *                  (LAPACK reflector convention: the basis equals the trailing columns of the complete Q);
 *   basis != NULL: elimination only, on the given k x S rows (k = S - np), design ignored.
 * exact != 0 keeps the reference's unfused (phi_j * v_i) / v_j arithmetic in the elimination (bit-identical pivots
 * for the same basis); exact == 0 uses one division per row and an FMA per element.
 * mu (S) is updated in place; info (2 int32, may be NULL): [0] elimination steps taken, [1] 1 if the QR ran.
 * sober_car_cluster_fits returns the cluster size that would be used (0: does not fit -> use the two-step path). */
int sober_car_cluster_fits(int32_t S, int32_t np, int32_t have_basis);
int sober_car_cluster(const double* design, const double* basis, int32_t S, int32_t np, double* mu, int32_t exact,
                      int32_t* info, void* stream);
/* Same, with 12 int64 cycle counters of CTA 0 / thread 0 written to prof (diagnostics: where a step's time goes). */
int sober_car_cluster_profiled(const double* design, const double* basis, int32_t S, int32_t np, double* mu,
                               int32_t exact, int32_t* info, int64_t* prof, void* stream);

/* Elimination only (SOBER/_rchq.py:237-266) on a given basis (k x S rows, DESTROYED is not -- it is only read),
 * column-distributed over an 8-CTA cluster with the matrix in registers: no cross-CTA reduction per step, one
 * broadcast of the pivot column.  Covers k <= 256, S <= 448; sober_car_cluster_cols_fits returns 0 otherwise.
 * exact as in sober_car_cluster.  mu (S) updated in place; info[0] (may be NULL) = steps taken. */
int sober_car_cluster_cols_fits(int32_t S, int32_t k);
int sober_car_cluster_cols(double* basis, int32_t k, int32_t S, double* mu, int32_t exact, int32_t* info, void* stream);
/* Same with 8 int64 cycle counters (CTA 0 / thread 0) written to prof (diagnostics). */
int sober_car_cluster_cols_profiled(double* basis, int32_t k, int32_t S, double* mu, int32_t exact, int32_t* info,
                                    int64_t* prof, void* stream);

/* Elimination only (SOBER/_rchq.py:237-266) on a given basis as a BLOCKED factorisation (csrc/car_panel.cu): panels of
 * up to 64 pivot columns are eliminated on one 8-CTA cluster (rows distributed over the CTAs, one st.async all-to-all
 * exchange of pivot-row candidates per step, ~0.5 us), the trailing columns then receive the rank-nb update as one
 * whole-GPU FP64 GEMM.  If the whole basis fits the cluster's shared memory it is a single panel.  S <= 4096.
 * Fused arithmetic (the exact == 0 variant of the kernels above): pivots equal the reference's up to rounding.
 *   basis: k x S rows, DESTROYED.  mu (S) updated in place.  nb_hint: 0 = automatic, else the panel width (<= 64).
 *   info (2 int32, device, may be NULL): [0] 1 if the elimination stopped early (no positive entry), [1] steps taken.
 *   workspace: sober_car_panel_workspace(S, k) bytes.  sober_car_panel_fits: 0 = unsupported, 1 = single panel,
 *   2 = blocked. */
int sober_car_panel_fits(int32_t S, int32_t k);
int64_t sober_car_panel_workspace(int32_t S, int32_t k);
int sober_car_panel(double* basis, int32_t k, int32_t S, double* mu, int32_t nb_hint, int32_t* info, void* workspace,
                    int64_t workspace_bytes, void* stream);
/* Same with 8 int64 cycle counters (CTA 0 / thread 0 of the panel kernel, accumulated: zero them first). */
int sober_car_panel_profiled(double* basis, int32_t k, int32_t S, double* mu, int32_t nb_hint, int32_t* info,
                             void* workspace, int64_t workspace_bytes, int64_t* prof, void* stream);

/* The second count of the remainder (SOBER/_rchq.py:153-164: the points beyond E*S are ALSO added to the last group)
 * applied to the group sums of one iteration in one launch:  at_last_row (= At + (S-1)*L', L' entries) += tail_at;
 * totw_out[g] = totw_in[g] (+ tail_tw[0] for g = S-1).  tail_at = tail_tw = NULL: no remainder, totw is just copied. */
int sober_apply_tail(double* at_last_row, const double* tail_at, int32_t Lp, const double* totw_in,
                     const double* tail_tw, int32_t S, double* totw_out, void* stream);

/* Fused helpers of one Caratheodory step (csrc/car_helpers.cu), each replacing several library launches:
 *  sober_car_prepare: out (S x (n+1), ldo) = column-normalised design matrix [1 | F / div] -- barycentres
 *    (SOBER/_rchq.py:166; div may be NULL), ones column (:229) and the column scaling of the projector null space.
 *  sober_car_summary: after the elimination.  If |delta|_F >= defect_limit (delta: ndelta doubles, the orthogonality
 *    defect of the one-pass Cholesky-QR) or a weight is not finite, every weight becomes NaN; then
 *    summary[i] = kept groups among 0..i, summary[S] = 1 if everything was fine, rank[i] = kept groups below i
 *    (the bookkeeping of SOBER/_rchq.py:198-221 as the host reads it with one copy).  S <= 4096. */
int sober_car_prepare(const double* F, int64_t ldf, const double* div, int32_t S, int32_t n, double* out, int64_t ldo,
                      void* stream);
int sober_car_summary(double* w, int32_t S, const double* delta, int64_t ndelta, double defect_limit, int32_t* summary,
                      int32_t* rank, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Weight update + compaction of the alive-list (SOBER/_rchq.py:198-221).
 *   For local position j (global p = pos0 + j):
 *     p <  ES: g = p mod S; kept iff wstar[g] > 0; new global position q = (p / S) * K + rank[g]
 *     p >= ES: kept iff tail_keep;                  q = (ES / S) * K + (p - ES)
 *     kept:    mu_out[q - new_pos0] = (mu_in[j] * wstar[g]) / totw[g]      (g = S-1 for the tail)
 *              idx_out[q - new_pos0] = idx_in[j]
 *   rank[g] = number of kept groups below g, K = number of kept groups (both computed by the caller).
 *   rec_in / rec_out (may be NULL): the record rows move with their points, weight slot (column d + 1) updated.
 * ------------------------------------------------------------------------------------------------- */
int sober_update_compact(const int32_t* idx_in, const double* mu_in, int64_t n_local, int64_t pos0, int64_t ES,
                         int32_t S, const double* wstar, const double* totw, const int32_t* rank, int32_t K,
                         int32_t tail_keep, int64_t new_pos0, int32_t* idx_out, double* mu_out,
                         const double* rec_in, double* rec_out, int64_t ldr, int32_t d, void* stream);

/* The same with K, tail_keep and new_pos0 taken on the DEVICE from ``summary`` (the inclusive cumulative kept-count of
 * sober_car_summary): K = summary[S-1], tail_keep = K > summary[S-2], new_pos0 = number of survivors below pos0 (the
 * closed form above).  Nothing here waits for the host, so the call can be enqueued right behind the Caratheodory step;
 * the outputs must hold n_local entries (the number of survivors is not known to the host yet). */
int sober_update_compact_dev(const int32_t* idx_in, const double* mu_in, int64_t n_local, int64_t pos0, int64_t ES,
                             int32_t S, const double* wstar, const double* totw, const int32_t* rank,
                             const int32_t* summary, int32_t* idx_out, double* mu_out, const double* rec_in,
                             double* rec_out, int64_t ldr, int32_t d, void* stream);

/* dst[:] = 0 ; dst[idx[j]] = w[j]   -- the in-place sparse result of SOBER/_rchq.py:109-110. */
int sober_scatter_result(double* dst, int64_t n, const int64_t* idx, const double* w, int64_t m, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Projection onto the Nystrom basis + barycentres + design matrix (SOBER/_rchq.py:148-166, :229), FP64 tensor pipe:
 *   design[g, 0] = 1;  design[g, 1 + j] = (sum_l (At[g,l] + [g == S-1] tail[l]) Uext[j,l]) / totw_out[g]
 *   totw_out[g] = totw[g] + [g == S-1] tail_tw[0]
 * At: S x Lp (lda), Uext: n x Lp (ldu) = [U | -U K_zX W], design: S x (n + 1) (ldd); tail (Lp) / tail_tw (1) may be
 * NULL (no remainder); totw_out may be NULL.  mma.sync m8n8k4 f64 (DMMA) with the epilogue fused.
 * ------------------------------------------------------------------------------------------------- */
int sober_project_design(const double* At, int64_t lda, const double* tail, const double* totw, const double* tail_tw,
                         const double* Uext, int64_t ldu, int32_t S, int32_t Lp, int32_t n, double* design, int64_t ldd,
                         double* totw_out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Small dense helper of the Cholesky-QR steps (Nystrom range finder, projector null space):
 *   solve X * R = Y for upper-triangular R (q x q row-major, q <= 256), Y and X m x q row-major (X may alias Y).
 * R must be 16-byte aligned with ldr <= 384 (its 16-row tiles are TMA bulk copies), else SOBER_ERR_UNSUPPORTED.
 * One warp per two rows, the rows in registers.  Replaces torch.linalg.solve_triangular (cuBLAS trsm: ~0.1 ms per call at
 * q = 200, called ~40 times per recombination).
 * ------------------------------------------------------------------------------------------------- */
int sober_trsm_right_upper(const double* Y, int64_t ldy, const double* R, int64_t ldr, int32_t m, int32_t q, double* X,
                           int64_t ldx, void* stream);

/* Cholesky factorisation G = R^T R of a symmetric positive definite q x q matrix (row-major, q <= 224; the upper
 * triangle of G is read), R upper triangular with zeros below the diagonal.  One 2-CTA cluster, the triangle in
 * registers, columns handed over through a shared-memory ring / bulk DSMEM copies (csrc/chol_pair.cu).  Replaces
 * torch.linalg.cholesky_ex (cuSOLVER potrf: ~0.13 ms at q = 200).  *info (device) = 0, or 1 + the first column whose
 * pivot was not positive (R is then NaN from that column on), as LAPACK reports it. */
int sober_cholesky_upper_fits(int32_t q);
int sober_cholesky_upper(const double* G, int64_t ldg, int32_t q, double* R, int64_t ldr, int32_t* info, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Assignment step of Lloyd's k-means (SOBER/_weights.py:100-126, the KMeans that produces the Nystrom landmarks,
 * SOBER/_sampler.py:316-317):  labels[i] = argmin_k sum_j (X[i,j] - C[k,j])^2, first minimum on ties, first NaN wins
 * (torch.argmin).  X: n x d (ldx), C: K x d contiguous, d <= 16; labels: n int64.  Nothing of size n*K is stored.
 * ------------------------------------------------------------------------------------------------- */
int sober_kmeans_assign(const double* X, int64_t ldx, int64_t n, int32_t d, const double* C, int32_t K, int64_t* labels,
                        void* stream);

/* ---------------------------------------------------------------------------------------------------
 * pi evaluation over the candidate set (SOBER/_gp.py:212-238 `predict`, SOBER/_pi.py:20-38
 * `PI.lfi`): the per-candidate epilogue of the GP posterior.  K (m x n_obs, ldk) = k(x_i, Xobs_j) from
 * sober_group_accumulate in Gram mode, T (m x n_obs, ldt) = K W with W = (K_obs + noise I)^-1 (may be NULL: mean only),
 * alpha (n_obs) = W (y - c):
 *   mean[i] = mean_const + sum_j K[i,j] alpha[j]
 *   var[i]  = max(min_var, (kxx ? kxx[i] : kxx_const) - sum_j K[i,j] T[i,j] + noise)
 *   pi[i]   = Phi((mean[i] - eta) / sqrt(var[i]))
 * Any of mean / var / pi may be NULL.  One warp per row, one streaming pass over the two tiles.
 * ------------------------------------------------------------------------------------------------- */
int sober_gp_rows(const double* K, int64_t ldk, const double* T, int64_t ldt, const double* alpha, int64_t m,
                  int32_t n_obs, double mean_const, const double* kxx, double kxx_const, double noise, double min_var,
                  double eta, double* mean, double* var, double* pi, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Stream confined to all SMs of the current device but `reserve_sms` (a CUDA green context; created on first use,
 * cached, never destroyed).  Work launched on it leaves the reserved SMs to the other streams: the first K1 pass runs
 * there beside the one-CTA kernels of the Nystrom range finder.  *stream = NULL when the driver cannot partition the
 * device (still SOBER_OK): do not overlap then.  *sm_count = SMs of the partition.
 * ------------------------------------------------------------------------------------------------- */
int sober_partition_stream(int32_t reserve_sms, void** stream, int32_t* sm_count);

/* ---------------------------------------------------------------------------------------------------
 * Diagnostics: FP64 FMA throughput probe (the roofline denominator for K1, which is FP64-pipe bound).
 * Launches `blocks` x 256 threads, each doing iters * 8 dependent-chain DFMAs; flops = blocks*256*iters*16.
 * ------------------------------------------------------------------------------------------------- */
int sober_fp64_probe(int32_t blocks, int64_t iters, double* sink, void* stream);
/* Integer-pipe probe (the roofline denominator of the bit-packed K1 kernels): blocks x 256 threads, each
 * iters * 4 word-ops, one word-op = popcount(x & z) on 64-bit words accumulated into an int32. */
int sober_popc_probe(int32_t blocks, int64_t iters, int32_t* sink, void* stream);
/* FP64 tensor-pipe probe: blocks x 8 warps, each iters * 8 DMMA m8n8k4; flops = blocks * 8 * iters * 8 * 512. */
int sober_dmma_probe(int32_t blocks, int64_t iters, double* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SOBER_B200_H */
