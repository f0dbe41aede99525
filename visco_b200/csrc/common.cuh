// Shared declarations for libvisco_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/visco_b200.h"

#define VK_SMEM_BUDGET (200 * 1024)  // dynamic shared memory we allow one CTA to ask for
#define VK_MAX_R 2048                // largest min(m, n) the selection kernel sorts
#define VK_HOST_CHUNKS 4             // sub-batches the *_host entry points pipeline copies and compute over
#define VK_MAX_GROUPS 4              // independent matrix groups (streams) the Jacobi driver overlaps

struct vk_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    void* ws = nullptr;  // grow-only device workspace owned by the handle
    size_t ws_bytes = 0;
    void* stage = nullptr;  // grow-only device staging for the *_host entry points
    size_t stage_bytes = 0;
    int32_t* h_poll = nullptr;  // pinned host words the convergence loop copies into
    void* d_scratch = nullptr;  // 256 bytes of device memory for counters of the small utility kernels
    int64_t launches = 0;
    int num_sms = 148;
    // options
    float jacobi_tol = 1e-4f;  // sweep-level stop: every off-diagonal met in the sweep was below this (then rotated)
    int max_sweeps = 30;
    int gram_impl = 0;
    int check_finite = 1;
    int check_every = 1;
    int jacobi_bsz = 0;  // 0 = auto
    int jacobi_groups = 0;  // 0 = auto
    cudaStream_t sub[VK_MAX_GROUPS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t sub_ev[VK_MAX_GROUPS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork_ev = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t copy_stream2 = nullptr;  // host-to-device direction of vk_reconstruct_host
    cudaEvent_t host_ev[2 * VK_HOST_CHUNKS + 1] = {};
    int stage_timing = 0;
    int chunk = 0;  // matrices per internal pass, 0 = auto
    int topk = 0;            // 0 = auto (subspace iteration for compressionrank <= 4), 1 = full Jacobi only, 2 = up to rank 8
    int gemm_impl = 0;       // 0 = auto (tcgen05 GEMM for k > 8 where the shape allows), 1 = SIMT only
    int tridiag_variant = 0;  // tridiag_sym.cu: 0 = by batch size, 1 = one matrix per SM, 2 = two per SM
    int small_impl = 0;       // one-sided Jacobi path: 0 = for r <= 32, 1 = whenever the matrix fits one CTA, 2 = never
    int factors_impl = 0;     // small ranks on the wide Gram path: 0 = one fused cluster kernel, 1 = the separate kernels
    int tail_split = 0;       // direct eigensolver: 0 = remainder sub-batch on a second stream (tridiag.cu), 1 = off
    bool in_split = false;
    cudaStream_t tail_stream = nullptr;
    cudaEvent_t tail_ev[2] = {nullptr, nullptr};
    int tridiag_nts = -1;     // tridiag_sym.cu: largest trailing block that moves to shared memory (-1: whatever fits)
    int tridiag_small_rs = 0; // tridiag_small.cu: row groups per matrix for 33 < r <= 64 (1, 2 or 4; 0 = two from r = 60 on)
    int bisect_impl = 0;      // leading eigenvalues of many small problems: 0 = packed kernel (several matrices per warp), 1 = one CTA each
    int gram_small = 0;       // min(m,n) <= 64: 0 = fused Gram + normalisation kernel, 1 = SIMT GEMM + normalisation pass
    int tridiag_pf = 1;       // tridiag_sym.cu: L2 prefetch distance in tiles (r = 512, two matrices per SM; 0 = none)
    int split_variant = 0;    // tridiag.cu: launch shape of the main sub-batch under the remainder split (as tridiag_variant)
    int32_t* bad = nullptr;   // per-matrix flags of the current vk_compress_batched call: Gram trace outside the safe range
    size_t bad_bytes = 0;     // (grow-only; bad[B] is the count)
    int recon_tc_impl = 0;   // 0 = persistent tcgen05 kernel for 8 < k <= 32 (recon_tc.cu), 1 = the older kernels
    int recon_generic = 0;   // 1 = always use the generic GEMM reconstruction kernel (debug / comparison)
    int small_reg = 1;       // 1 = register-resident recursive tournament for power-of-two small problems
    int jacobi_generic = 0;  // 1 = never use the register-resident cross kernel (debug / comparison)
    int eig_impl = 0;        // 0 = auto, 1 = cyclic Jacobi, 2 = tridiagonalisation + implicit QL (tridiag.cu)
    int tridiag_impl = 0;    // 0 = auto (deferred updates for 384 < r <= 512), 1 = update every step, 2 = deferred updates also for 128 < r <= 256
    int eigvec_impl = 0;     // eigenvectors of T on the full path: 0 = twisted factorisation + Newton-Schulz + GEMMs, 1 = implicit QL
    int ql_maxit = 60;       // QL iterations allowed per eigenvalue (tests lower it to exercise the Jacobi fallback)
    int64_t eig_fallbacks = 0;  // internal passes the direct solver handed back to the Jacobi solver
    // retained sigma_k / sigma_1 below this: redo the matrix without a Gram product (0 = never). Measured error of
    // sigma_k through the float32 Gram path (tools/illcond_probe.py): 5e-6 at a ratio of 0.008, 2e-5 at 0.005, 9e-5 at
    // 0.004, 4e-4 at 0.0024, 2e-2 at 0.0007
    float illcond_thr = 0.005f;
    int64_t illcond_redone = 0; // matrices redone that way since the handle was created
    void* ws2 = nullptr;        // grow-only device workspace of that path
    size_t ws2_bytes = 0;
    float stage_ms[6] = {0, 0, 0, 0, 0, 0};
    float eig_ms[5] = {0, 0, 0, 0, 0};  // direct eigensolver: tridiag, leading pairs, QL, reflector accumulation, rotations
    cudaEvent_t eig_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

#define VK_CUDA(h, call)                                                                                   \
    do {                                                                                                   \
        cudaError_t _e = (call);                                                                           \
        if (_e != cudaSuccess) {                                                                           \
            (h)->err = std::string(#call) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                       std::to_string(__LINE__) + ")";                                                     \
            return VK_ECUDA;                                                                               \
        }                                                                                                  \
    } while (0)

#define VK_LAUNCH_CHECK(h)                  \
    do {                                    \
        (h)->launches++;                    \
        VK_CUDA(h, cudaPeekAtLastError()); \
    } while (0)

static inline int vk_fail(vk_context* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

// ---- small device helpers ---------------------------------------------------------------------------------
// acc += a * b
__device__ __forceinline__ void cfma(float2& acc, float2 a, float2 b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- stage launchers (defined in the .cu files) -------------------------------------------------------------
// Layout of the "vector workspace" W used by the eigensolver / small SVD: [B][r][ld] complex64, vector i contiguous.
struct JacobiPlan {
    int r;      // number of vectors
    int ldot;   // leading entries that enter the inner products
    int ltot;   // entries that are rotated (ltot - ldot trailing entries carry the accumulated rotations)
    int ld;     // stride between vectors (>= ltot)
    int bsz;    // vectors per block (= warps per CTA)
    int nb;     // number of blocks (even)
    size_t smem;
};
JacobiPlan vk_jacobi_plan(const vk_context* h, int r, int ldot, int ltot);

// W in/out; offmax/done/sweeps are per-matrix device arrays of length B (int32 / float bits), active is 1 int.
// preset_done: done_dev / sweeps_dev already hold the verdict of the fixed-rank fast path (1 = solved, skip)
int vk_launch_jacobi(vk_context* h, float2* W, int B, const JacobiPlan& p, int32_t* sweeps_dev, int32_t* done_dev,
                     unsigned* offmax_dev, int32_t* active_dev, bool preset_done = false);
// direct eigensolver (tridiag.cu): same W in/out contract as vk_launch_jacobi for the Gram path (ldot = ltot = r)
bool vk_eigqr_supported(int r);
size_t vk_eigqr_scratch_bytes(int B, int r);
int vk_launch_eigqr(vk_context* h, float2* W, int B, int r, int ld, void* scratch, int32_t* sweeps_dev,
                    int32_t* done_dev, int fixed_rank = 0, double decorrelation = 0.0);
int vk_concurrent_compress(int device);   // host threads inside vk_compress_batched on this device (api.cu)
// lower-triangle, deferred-update Householder tridiagonalisation (tridiag_sym.cu); outputs as the kernels of tridiag.cu
bool vk_tridiag_symdefer_supported(int r);
int vk_launch_tridiag_symdefer(vk_context* h, cudaStream_t st, float2* W, int B, int r, int ld, size_t wstride, float* d,
                               float* e, float* tau, float2* ph);
// r <= 64: one warp per matrix, matrix in shared memory (tridiag_small.cu)
bool vk_tridiag_small_supported(int r);
int vk_launch_tridiag_small(vk_context* h, cudaStream_t st, float2* W, int B, int r, int ld, size_t wstride, float* d, float* e,
                            float* tau, float2* ph);
bool vk_topk_supported(int r, int fixed_rank, bool force);
int vk_launch_topk(vk_context* h, float2* W, int B, int r, int fixed_rank, int32_t* done_dev, int32_t* sweeps_dev);

int vk_launch_gram_simt(vk_context* h, const float2* A, int B, int m, int n, int side, float2* W);
int vk_launch_gram_tc(vk_context* h, const float2* A, int B, int m, int n, float2* W);
// min(m, n) <= 64: Gram product and trace normalisation fused, one CTA per matrix (stages.cu)
bool vk_gram_small_supported(int m, int n);
int vk_launch_gram_small(vk_context* h, const float2* A, int B, int m, int n, float2* W, float* gscale_dev,
                         int32_t* nonfinite_dev, int32_t* bad_dev, int32_t* nbad_dev);
bool vk_gram_tc_supported(int m, int n, int side);

// scale[b] = r / trace(W[b]) applied in place; gscale_dev[b] = trace/r ; nonfinite_dev[0] |= 1 when trace is NaN/Inf
// bad_dev / nbad_dev (optional): per-matrix flag and count of traces outside the float32-safe range (see stages.cu)
int vk_launch_gram_normalise(vk_context* h, float2* W, int B, int r, float* gscale_dev, int32_t* nonfinite_dev,
                             int32_t* bad_dev = nullptr, int32_t* nbad_dev = nullptr);

// small path: build [B][r][ld] vectors (rows of A, or columns when m > n) followed by an r x r identity
int vk_launch_pack_small(vk_context* h, const float2* A, int B, int m, int n, float2* W, int ld, float* gscale_dev,
                         int32_t* nonfinite_dev);

// norms of the vectors, sort, sigma, rank choice. mode_gram: sigma = sqrt(norm * gscale) else sigma = norm * gscale
int vk_launch_select(vk_context* h, const float2* W, int B, int r, int ldot, int ld, const float* gscale_dev,
                     int mode_gram, int fixed_rank, double decorrelation, int kmax, int32_t* perm_dev, float* inv_dev,
                     float* S_dev, int32_t* ranks_dev, float* stats_dev, const int32_t* sweeps_dev,
                     const int32_t* done_dev);
int vk_launch_pack_info(vk_context* h, const int32_t* sweeps, const int32_t* done, int B, int32_t* info);
int vk_launch_count_not_done(vk_context* h, const int32_t* done, int B, int32_t* out);
int vk_launch_flag_illcond(vk_context* h, const float* S, const int32_t* ranks, int B, int kmax, float thr, int32_t* flags,
                           int32_t* count);
int vk_launch_find_n(vk_context* h, const float* S, int B, int r, double decorrelation, int32_t* ranks);

// factor formation
int vk_launch_factors_gram(vk_context* h, const float2* A, const float2* W, int B, int m, int n, int side, int kmax,
                           const int32_t* perm_dev, const float* inv_dev, const int32_t* ranks_dev, float* norm2_dev,
                           float2* U, float* S, float2* Vt, float* stats_dev, float2* xbuf);
int vk_launch_factors_small(vk_context* h, const float2* W, int ld, int B, int m, int n, int kmax,
                            const int32_t* perm_dev, const float* inv_dev, const int32_t* ranks_dev, float2* U,
                            float2* Vt);
bool vk_cgemm_tc_supported(int m, int n, int kmax);
int vk_launch_formv_tc(vk_context* h, const float2* X, const float2* A, const int32_t* ranks, float2* Vt, float* norm2,
                       int B, int m, int n, int kmax);
int vk_launch_recon_tc(vk_context* h, const float2* U, const float* S, const float2* Vt, const int32_t* ranks, float2* out,
                       int B, int m, int n, int kmax);
int vk_launch_cgemm_tc_plain(vk_context* h, const float2* P, const float2* Q, float2* D, const float* rowscale, int B,
                             int M, int N, int K);
bool vk_recon_tc_supported(int m, int n, int kmax);
int vk_launch_recon_tc_smallk(vk_context* h, const float2* U, const float* S, const float2* Vt, const int32_t* ranks,
                              float2* out, int B, int m, int n, int kmax);
int vk_launch_reconstruct(vk_context* h, const float2* U, const float* S, const float2* Vt, const int32_t* ranks, int B,
                          int m, int n, int kmax, float2* out);
int vk_launch_synth(vk_context* h, float2* A, int nbl_local, int ncorr, int m, int n, int bl_offset, int nbl_total,
                    uint64_t seed);
