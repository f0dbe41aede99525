/*
 * visco_b200.h — C ABI of libvisco_b200.so: the B200 (sm_100a) replacement for VISCO's data-parallel hot path.
 *
 * The reference has no FFI: its "operator API" for this path is three Python callables
 *     apply_svd(visdata, decorrelation, compressionrank)      visco/compress_ms.py:322-363
 *     find_n_decorrelation(singular_values, decorrelation)    visco/compress_ms.py:295-319
 *     reconstruct_vis(U, S, Vt)                               visco/decompress_ms.py:107-131
 * invoked once per (baseline, correlation) matrix through dask.delayed (compress_ms.py:610,640,671 and
 * decompress_ms.py:196). This library is what a ctypes binding behind those names calls; the batched entry
 * points take every matrix of a dask batch at once (compress_ms.py:571-697, decompress_ms.py:207-213).
 *
 * Conventions
 *   - every function returns an int status (VK_OK == 0); vk_last_error(h) gives a human-readable message.
 *   - complex64 is interleaved (re, im) float pairs, C order, as numpy stores it.
 *   - "dev" pointers are CUDA device pointers on the handle's device, 16-byte aligned; the caller owns them.
 *     "host" entry points take ordinary (ideally pinned) host pointers and do the copies themselves.
 *   - all device work is enqueued on the handle's stream (vk_set_stream); device entry points are asynchronous
 *     unless stated otherwise; vk_sync() waits for the stream.
 *   - one handle per GPU and per host thread; different handles may be used concurrently.
 *
 * Layouts (B matrices, each m rows (time) x n columns (channel); r = min(m, n); kmax = padded rank)
 *   A    [B][m][n]      complex64
 *   U    [B][m][kmax]   complex64   columns >= ranks[b] are zero
 *   S    [B][kmax]      float32     entries >= ranks[b] are zero
 *   Vt   [B][kmax][n]   complex64   rows >= ranks[b] are zero      (the reference calls it Vt / WT)
 *   ranks[B]            int32
 *   stats[B][4]         float32     { ||A||_F^2 (sum of all sigma^2), retained energy (sum of kept sigma^2),
 *                                     eigensolver iterations (QL iterations, or Jacobi sweeps; 0 when only the
 *                                     leading eigenpairs were computed), converged flag (1/0) }
 */
#ifndef VISCO_B200_H
#define VISCO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VK_OK 0
#define VK_EINVAL 1  /* bad argument (shape, NULL pointer, kmax too small, ...)  -> Python ValueError      */
#define VK_ENOMEM 2  /* device or host allocation failed                         -> Python MemoryError     */
#define VK_ECUDA 3   /* a CUDA call or kernel failed                             -> Python RuntimeError    */
#define VK_ENOCONV 4 /* the eigensolver did not converge for some matrix (numpy raises LinAlgError here)        */
#define VK_ENONFINITE 5 /* NaN/Inf in the input (flagged rows NaN-masked by ds.where, compress_ms.py:470)  */

typedef struct vk_context* vk_handle;

/* ---- life cycle ------------------------------------------------------------------------------------------- */
int vk_create(vk_handle* out, int device);
int vk_destroy(vk_handle h);
const char* vk_last_error(vk_handle h);
const char* vk_version(void);
/* cudaStream_t passed as void*; NULL = the legacy default stream. */
int vk_set_stream(vk_handle h, void* cuda_stream);
int vk_sync(vk_handle h);
/* options: "jacobi_tol" (relative off-diagonal stop level, default 1e-4: a sweep that met nothing larger ends the
 *          iteration; its own rotations leave the vectors orthogonal at the float32 noise floor), "max_sweeps" (default 30),
 *          "gram_impl" (0 = auto: tcgen05 where the shape allows, 1 = force SIMT fp32, 2 = force tcgen05),
 *          "check_finite" (default 1), "check_every" (host convergence poll period in sweeps, default 1),
 *          "jacobi_bsz" (vectors per block, 0 = auto), "jacobi_groups" (concurrent matrix groups, 0 = auto), "chunk" (matrices per internal pass, 0 = auto),
 *          "stage_timing" (0/1, see vk_last_stage_ms), "topk" (shortcuts for small ranks: 0 = auto, 1 = none, 2 = extended; what they
 *          are depends on "eig_impl", see there; with the Jacobi solver: blocked subspace iteration for fixed rank <= 4 / <= 8
 *          with per-matrix fallback to the full solver), "gemm_impl" (0 = tcgen05 GEMM for k > 8, 1 = SIMT),
 *          "small_reg" (1 = register-resident recursive-tournament kernel on the small path where the shape allows
 *          (default), 0 = shared-memory round-robin kernel only),
 *          "eig_impl" (Hermitian eigensolver of the Gram path: 0 = auto = 2 where min(m, n) <= 1024, 1 = one-sided cyclic
 *          Jacobi, 2 = Householder tridiagonalisation + implicit QL; with 0/2 a fixed rank <= 32 takes only the leading
 *          eigenpairs: Sturm bisection + twisted factorisation, unless "topk" = 1; "topk" = 2 extends that to the energy
 *          rule: all eigenvalues by bisection, the rank estimated on the device, leading pairs per matrix where that
 *          rank is <= 30 - pays on high-SNR data where most matrices qualify),
 *          "illcond_thr" (default 0.005: a matrix whose smallest retained singular value is below that * sigma_1 cannot be
 *          resolved through the float32 Gram matrix and is done again by one-sided Jacobi on the matrix itself; 0 = never),
 *          "ql_maxit" (QL iterations per eigenvalue before the pass is handed to the Jacobi solver, default 60),
 *          "eigvec_impl" (full spectrum with eig_impl 0/2: 0 = bisection + twisted factorisation + Newton-Schulz + tcgen05
 *          GEMMs, 1 = implicit QL with recorded rotations),
 *          "tridiag_impl" (0 = lower triangle with deferred updates for 128 < min(m,n) <= 512 and the warp-level kernel for
 *          <= 64, 1 = round 1's undeferred full-storage kernels, 2 = full storage with deferred updates),
 *          "tridiag_variant" (launch shape of the lower-triangle kernel: 0 = two matrices per SM when the batch exceeds the
 *          SM count or three or more host threads are inside vk_compress_batched on this device, else one; 1 / 2 force
 *          one / two), "tridiag_nts" (rows of the trailing block that finish in shared memory, default 64),
 *          "tail_split" (0 = the remainder of a batch that is not a whole number of waves runs as its own sub-batch on a
 *          second stream, 1 = off), "split_variant" (launch shape of the main sub-batch under that split, values as
 *          "tridiag_variant"), "tridiag_pf" (min(m,n) = 512, two matrices per SM: tiles the L2 is asked for ahead of the
 *          loads, default 1; 0 = none),
 *          "small_impl" (one-sided Jacobi on the matrix itself: 0 = for min(m,n) <= 32 when it fits one CTA, 1 = for every
 *          shape that fits (min(m,n) <= 64: BASELINE configs[3] as named), 2 = never),
 *          "tridiag_small_rs" (33 < min(m,n) <= 64: groups of two warps that share the rows of a matrix in the warp-level
 *          tridiagonalisation, 1 / 2 / 4; 0 = two from 60 on),
 *          "bisect_impl" (leading eigenvalues at a fixed small rank: 0 = batches of at least four matrices per SM with
 *          min(m,n) <= 128 pack several matrices into a warp, one lane per eigenvalue; 1 = always one CTA per matrix with
 *          eight lanes per eigenvalue),
 *          "gram_small" (min(m,n) <= 64 on the Gram path: 0 = Gram product and trace normalisation in one CTA per matrix,
 *          1 = the SIMT GEMM followed by the normalisation pass),
 *          "factors_impl" (small ranks, wide matrices: 0 = one fused cluster kernel, 1 = the separate kernels),
 *          "recon_tc_impl" (8 < k <= 32: 0 = persistent tcgen05 kernel with bulk tensor stores, 1 = the older kernels),
 *          "recon_generic" / "jacobi_generic" (comparison runs: 1 = the generic SIMT GEMM for every reconstruction and factor
 *          formation / the global-memory Jacobi kernels for every size; "jacobi_generic" = 2 also takes the round-1
 *          full-storage tridiagonalisation for 256 < min(m,n) <= 512).
 *          Every choice of these leaves the results within the tolerances of DESIGN.md section 2. */
int vk_set_option(vk_handle h, const char* key, double value);
/* bytes of device workspace vk_compress_batched needs for this problem (it allocates/grows the handle's own
 * workspace when ws == NULL). */
size_t vk_workspace_bytes(vk_handle h, int B, int m, int n, int kmax);

/* ---- the hot path ----------------------------------------------------------------------------------------- */
/* Replaces apply_svd (compress_ms.py:322-363) for a batch: economy SVD of every A[b], rank choice
 *   fixed_rank > 0              -> k = min(fixed_rank, r)                     (wins, compress_ms.py:352-353)
 *   else decorrelation > 0      -> k = find_n_decorrelation(S, decorrelation) (compress_ms.py:295-319, float32)
 *   else                        -> k = r
 * and truncated factors. kmax must be >= the largest possible k (fixed_rank, or r in energy/full mode).
 * Synchronous with respect to `ranks`/`stats` only if the caller syncs; status reflects errors found so far
 * (convergence / non-finite input are reported by this call because it polls the device while iterating). */
int vk_compress_batched(vk_handle h, const void* A_dev, int B, int m, int n, int fixed_rank, double decorrelation,
                        int kmax, void* U_dev, float* S_dev, void* Vt_dev, int32_t* ranks_dev, float* stats_dev,
                        void* ws_dev, size_t ws_bytes);

/* Replaces reconstruct_vis (decompress_ms.py:107-131) for a batch: out[b] = (U[b] * S[b][None, :]) @ Vt[b],
 * using the first ranks[b] modes (ranks_dev == NULL -> all kmax modes). out is [B][m][n] complex64. */
int vk_reconstruct_batched(vk_handle h, const void* U_dev, const float* S_dev, const void* Vt_dev,
                           const int32_t* ranks_dev, int B, int m, int n, int kmax, void* out_dev);

/* Same two operations with HOST buffers (numpy arrays): H2D copy, compute, D2H copy, synchronous.
 * These are what the single-matrix Python drop-ins apply_svd / reconstruct_vis call. */
int vk_compress_host(vk_handle h, const void* A_host, int B, int m, int n, int fixed_rank, double decorrelation,
                     int kmax, void* U_host, float* S_host, void* Vt_host, int32_t* ranks_host, float* stats_host);
int vk_reconstruct_host(vk_handle h, const void* U_host, const float* S_host, const void* Vt_host,
                        const int32_t* ranks_host, int B, int m, int n, int kmax, void* out_host);

/* Replaces find_n_decorrelation (compress_ms.py:295-319) on device: S_dev [B][r] float32 descending. */
int vk_find_n_decorrelation_batched(vk_handle h, const float* S_dev, int B, int r, double decorrelation,
                                    int32_t* ranks_dev);

/* ---- stage-level entry points (tests, ncu) ---------------------------------------------------------------- */
/* Gram product on the smaller side. side 0: W[b][i][t] = sum_v A[t][v] conj(A[i][v])   (r = m, G = A A^H, W = G^T)
 *                                   side 1: W[b][i][j] = sum_t conj(A[t][j]) A[t][i]   (r = n, G = A^H A, W = G^T)
 * W is [B][r][r] complex64: row i of W is column i of the Hermitian Gram matrix. impl as "gram_impl". */
int vk_gram_batched(vk_handle h, const void* A_dev, int B, int m, int n, int side, int impl, void* W_dev);
/* One-sided cyclic Jacobi on the columns of a Hermitian PSD matrix given as W (above), in place: on return the rows
 * of W are mutually orthogonal, unnormalised eigenvectors of the Gram matrix (row i = (lambda_i * r / trace) * v_i).
 * lambda_dev [B][r] receives the eigenvalues sorted descending, info_dev [B][2] = {sweeps, converged}. */
int vk_eigh_jacobi_batched(vk_handle h, void* W_dev, int B, int r, float* lambda_dev, int32_t* info_dev);
/* One-sided (Hestenes) Jacobi SVD of small matrices held entirely in shared memory (min(m,n) <= 64 and
 * min(m,n) * (max(m,n) + min(m,n)) * 8 bytes <= 200 KiB). Full factors, sorted: U [B][m][r], S [B][r], Vt [B][r][n]. */
int vk_svd_jacobi_small_batched(vk_handle h, const void* A_dev, int B, int m, int n, void* U_dev, float* S_dev,
                                void* Vt_dev, int32_t* info_dev);
/* 1 if (m, n) takes the one-sided Jacobi small-matrix path inside vk_compress_batched with default options, else 0
 * (Gram path). Default: eligible shapes with min(m,n) <= 32; option "small_impl" = 1 takes it for every eligible shape
 * (min(m,n) <= 64, e.g. BASELINE configs[3]), 2 never. */
int vk_uses_small_path(int m, int n);
/* 1 if the tcgen05 Gram kernel handles (m, n, side), else 0 (SIMT Gram). */
int vk_gram_uses_tcgen05(int m, int n, int side);

/* ---- layout on either side of the path ---------------------------------------------------- */
/* Gather per-(baseline, correlation) matrices out of an MS column data[row][chan][corr] (complex64, corr fastest):
 *   stack == 1:  A[bl * ncs + j][t][v]                    = data[row_idx[bl][t]][v][corr_sel[bl][j]]
 *   stack == 2:  A[bl * (ncs/2) + j/2][(j%2) * m + t][v]  = ...   (--correlation-optimized vstack of two correlations)
 * row_idx is [nbl][m], corr_sel is [nbl][ncs] (correlation planes per baseline entry).
 * Replaces the boolean-mask isel and [:, :, ci] slicing of compress_ms.py:591-592,604-608,664. row_idx < 0 = padding.
 * vk_scatter_baselines is the inverse (decompress_ms.py:216-232 incl. unstack_vis :95-104). */
int vk_gather_baselines(vk_handle h, const void* data_dev, int nchan, int ncorr, const int32_t* row_idx_dev, int nbl,
                        int m, const int32_t* corr_sel_dev, int ncs, int stack, void* A_dev);
int vk_scatter_baselines(vk_handle h, const void* cube_dev, int nchan, int ncorr, const int32_t* row_idx_dev, int nbl,
                         int m, const int32_t* corr_sel_dev, int ncs, int stack, void* data_dev);

/* Validates the index arrays of the two calls above before they are used: entries of row_idx outside [-1, nrow) and
 * entries of corr_sel outside [0, ncorr) are counted into *bad_host (synchronous). Returns VK_EINVAL when any is found
 * (the reference's numpy indexing raises IndexError there, decompress_ms.py:216-232). */
int vk_check_layout_indices(vk_handle h, const int32_t* row_idx_dev, size_t nrow_idx, int nrow,
                            const int32_t* corr_sel_dev, size_t ncorr_sel, int ncorr, int32_t* bad_host);

/* ---- flags ---------------------------------------------------------------------------------- */
/* np.packbits(flags, axis=None) / np.unpackbits(packed, count=n) on device, big bit order (compress_ms.py:478-483,
 * decompress_ms.py:240-246). flags are one byte per element (numpy bool). packed has (n + 7) / 8 bytes. */
int vk_packbits(vk_handle h, const uint8_t* flags_dev, size_t n, uint8_t* packed_dev);
int vk_unpackbits(vk_handle h, const uint8_t* packed_dev, size_t n, uint8_t* flags_dev);
/* In-place da.where(FLAG, replacement, data) over n complex64 visibilities: replacement = model_dev[e] when model_dev is
 * not NULL (use_model_data, compress_ms.py:531-542), else the constant value_re + i value_im (flagvalue, :549-562). */
int vk_flag_replace(vk_handle h, void* data_dev, const uint8_t* flags_dev, const void* model_dev, float value_re,
                    float value_im, size_t n);

/* ---- benchmark generator (SURVEY.md section 8d model), on device ------------------------------------------ */
/* A[b] for b = baseline * ncorr + corr; global baseline index = bl_offset + baseline out of nbl_total. */
int vk_synth_fill(vk_handle h, void* A_dev, int nbl_local, int ncorr, int m, int n, int bl_offset, int nbl_total,
                  uint64_t seed);

/* number of kernel launches issued through this handle since creation (bench.py's gpu_launches claim) */
int64_t vk_launch_count(vk_handle h);
/* elapsed device milliseconds of the stages of the LAST vk_compress_batched call, measured with CUDA events on the
 * handle's stream: t[0] gram, t[1] jacobi (all sweeps), t[2] select+truncate, t[3] factor formation (U/Vt),
 * t[4] small-path kernel, t[5] total. Requires option "stage_timing" = 1 (adds event records + one sync). */
int vk_last_stage_ms(vk_handle h, float* t6);
/* the eigen-solve slot t[1] of the last vk_compress_batched call split by kernel of the direct solver (option
 * "eig_impl" != 1): t[0] tridiagonalisation, t[1] leading eigenpairs (fixed small rank), t[2] implicit QL,
 * t[3] reflector accumulation, t[4] rotation application. Requires "stage_timing" = 1. */
int vk_last_eig_ms(vk_handle h, float* t5);

#ifdef __cplusplus
}
#endif
#endif /* VISCO_B200_H */
