// C ABI of libvisco_b200.so (see include/visco_b200.h). Orchestrates the stages; no arithmetic here.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

// host threads currently inside vk_compress_batched, per device: with several handles feeding one GPU the kernels that
// would take a whole SM per matrix choose the launch shape that lets two of them share one (tridiag_sym.cu)
static std::atomic<int> g_in_compress[64];
int vk_concurrent_compress(int device) { return (device >= 0 && device < 64) ? g_in_compress[device].load(std::memory_order_relaxed) : 1; }

namespace {

struct InCompress {
    int d;
    explicit InCompress(int dev) : d(dev) {
        if (d >= 0 && d < 64) g_in_compress[d].fetch_add(1, std::memory_order_relaxed);
    }
    ~InCompress() {
        if (d >= 0 && d < 64) g_in_compress[d].fetch_sub(1, std::memory_order_relaxed);
    }
};

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct WsLayout {
    size_t W, perm, inv, gscale, sweeps, done, offmax, active, nonfinite, norm2, xbuf, eig, total;
};

// the one-sided (Hestenes) Jacobi kernels hold the whole problem in one CTA's shared memory
bool small_eligible(int m, int n) {
    const int r = m < n ? m : n;
    const int L = m < n ? n : m;
    if (r > 64) return false;
    return (size_t)(r + (r & 1)) * (size_t)(L + r) * sizeof(float2) + 1024 <= VK_SMEM_BUDGET;
}
// ... and are taken ("small_impl" 0, the default) up to r = 32; above that the Gram path with the warp-level
// tridiagonalisation (tridiag_small.cu) is 5x faster (BASELINE configs[3], 8320 matrices of 64 x 64: 4.2 vs 24.7 ms) and
// the ill-conditioned-set safeguard still sends what needs it through the Jacobi kernels. "small_impl" 1: always when
// eligible (the configuration BASELINE configs[3] names), 2: never.
bool small_path(const vk_context* h, int m, int n) {
    if (!small_eligible(m, n)) return false;
    const int impl = h ? h->small_impl : 0;
    if (impl == 1) return true;
    if (impl == 2) return false;
    return (m < n ? m : n) <= 32;
}

// eigensolver of the Gram path: 1 = cyclic Jacobi (with the blocked subspace iteration for small fixed ranks),
// 0 / 2 = tridiagonalisation + implicit QL where the size is supported (leading pairs only for small fixed ranks)
bool use_qr(const vk_context* h, int m, int n, int fixed_rank = 0) {
    (void)fixed_rank;
    const int r = m < n ? m : n;
    if (!h || small_path(h, m, n) || !vk_eigqr_supported(r) || h->eig_impl == 1) return false;
    return true;
}

WsLayout ws_layout(const vk_context* h, int chunk, int m, int n, int kmax, int gchunk = 0, bool qr = false, bool direct = false) {
    if (gchunk < chunk) gchunk = chunk;
    const int r = m < n ? m : n;
    const int L = m < n ? n : m;
    WsLayout w;
    size_t off = 0;
    const size_t wbytes = (small_path(h, m, n) || direct) ? (size_t)gchunk * r * (L + r) * sizeof(float2)
                                                       : (size_t)gchunk * r * r * sizeof(float2);
    w.W = off, off += align_up(wbytes);
    w.perm = off, off += align_up((size_t)chunk * r * 4);
    w.inv = off, off += align_up((size_t)chunk * r * 4);
    w.gscale = off, off += align_up((size_t)gchunk * 4);
    w.sweeps = off, off += align_up((size_t)chunk * 4);
    w.done = off, off += align_up((size_t)chunk * 4);
    w.offmax = off, off += align_up((size_t)chunk * 4);
    w.active = off, off += 256;
    w.nonfinite = off, off += 256;
    w.norm2 = off, off += align_up((size_t)chunk * kmax * 4);
    // K-major copy of conj(U_k)^T for the tcgen05 V-formation (wide Gram path, k > 8)
    w.xbuf = off;
    if (!small_path(h, m, n) && !direct && m <= n && vk_cgemm_tc_supported(m, n, kmax))
        off += align_up((size_t)chunk * kmax * m * 8);
    w.eig = off;
    if (qr) off += align_up(vk_eigqr_scratch_bytes(chunk, r));
    w.total = off;
    return w;
}

int auto_chunk(const vk_context* h, int B, int m, int n, bool qr) {
    const int r = m < n ? m : n;
    const int L = m < n ? n : m;
    if (qr) {
        // the direct solver has a latency-bound stage (the scalar QL iteration, one lane per matrix) whose duration
        // does not depend on the number of matrices: take as many per pass as 14 GB of scratch allow
        const size_t per = vk_eigqr_scratch_bytes(1, r) + (size_t)r * r * 8;
        size_t c = ((size_t)14 << 30) / per;
        if (c < 1) c = 1;
        if (c > (size_t)B) c = B;
        return (int)c;
    }
    const size_t per = small_path(h, m, n) ? (size_t)r * (L + r) * 8 : (size_t)r * r * 8;
    // keep the Jacobi working set of one internal pass inside ~half of the 126 MB L2, but never below a few waves
    size_t c = (64u << 20) / (per ? per : 1);
    if (c < 32) c = 32;
    if (c > (size_t)B) c = B;
    return (int)c;
}

// The Gram product is launched over a super-chunk of several Jacobi chunks: one launch then covers many waves of
// tiles (a 32-matrix chunk of 512 x 4096 matrices is only 2.2 waves of 148 CTAs, a 28 % quantisation loss).
int gram_chunk(const vk_context* h, int B, int chunk, int m, int n) {
    if (small_path(h, m, n)) return chunk;
    const int r = m < n ? m : n;
    size_t cap = (512u << 20) / ((size_t)r * r * 8);  // at most 512 MB of Gram matrices at a time
    if (cap < 1) cap = 1;
    int g = chunk;
    while (g < 128 && (size_t)(g + chunk) <= cap) g += chunk;
    if (g > B) g = B < chunk ? chunk : ((B + chunk - 1) / chunk) * chunk;
    return g;
}

int ensure(vk_context* h, void** p, size_t* have, size_t need) {
    if (*have >= need) return VK_OK;
    if (*p) {
        // every stream of this handle may still reference the old block: the Jacobi driver's group streams and the copy
        // streams of the *_host entry points as well as the main stream
        cudaStreamSynchronize(h->stream);
        for (int i = 0; i < VK_MAX_GROUPS; ++i)
            if (h->sub[i]) cudaStreamSynchronize(h->sub[i]);
        if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
        if (h->copy_stream2) cudaStreamSynchronize(h->copy_stream2);
        if (h->tail_stream) cudaStreamSynchronize(h->tail_stream);
        cudaFree(*p);
        *p = nullptr;
        *have = 0;
    }
    need = align_up(need, 1 << 20);
    if (cudaMalloc(p, need) != cudaSuccess) {
        cudaGetLastError();
        return vk_fail(h, VK_ENOMEM, "cudaMalloc of " + std::to_string(need) + " bytes failed");
    }
    *have = need;
    return VK_OK;
}

struct StageTimer {
    vk_context* h;
    bool on;
    explicit StageTimer(vk_context* hh) : h(hh), on(hh->stage_timing != 0) {}
    void mark(int i) {
        if (on) cudaEventRecord(h->ev[i], h->stream);
    }
    void collect(int slot, int i0, int i1) {
        if (!on) return;
        float ms = 0.f;
        cudaEventSynchronize(h->ev[i1]);
        if (cudaEventElapsedTime(&ms, h->ev[i0], h->ev[i1]) == cudaSuccess) h->stage_ms[slot] += ms;
    }
};

int gram_stage(vk_context* h, const float2* A, int B, int m, int n, float2* W, float* gscale, int32_t* nonfinite,
               int32_t* bad = nullptr, int32_t* nbad = nullptr) {
    const int r = m < n ? m : n;
    const int side = m <= n ? 0 : 1;
    int rc;
    // (tensor maps need a 16-byte aligned base: a misaligned device pointer takes the SIMT kernels)
    const bool aligned16 = (reinterpret_cast<uintptr_t>(A) % 16) == 0;
    const bool tc = (h->gram_impl == 2) || (h->gram_impl == 0 && vk_gram_tc_supported(m, n, side) && aligned16);
    if (tc) {
        if (!vk_gram_tc_supported(m, n, side) || !aligned16)
            return vk_fail(h, VK_EINVAL, "gram_impl=2 (tcgen05) does not support this shape or alignment");
        if ((rc = vk_launch_gram_tc(h, A, B, m, n, W))) return rc;
    } else if (h->gram_small == 0 && vk_gram_small_supported(m, n)) {
        return vk_launch_gram_small(h, A, B, m, n, W, gscale, nonfinite, bad, nbad);   // normalisation included
    } else {
        if ((rc = vk_launch_gram_simt(h, A, B, m, n, side, W))) return rc;
    }
    return vk_launch_gram_normalise(h, W, B, r, gscale, nonfinite, bad, nbad);
}

int compress_chunk(vk_context* h, const float2* A, int B, int m, int n, int fixed_rank, double decorrelation, int kmax,
                   float2* U, float* S, float2* Vt, int32_t* ranks, float* stats, unsigned char* ws, const WsLayout& L,
                   int sub0 = 0, bool gram_done = false, bool force_jacobi = false, bool direct = false) {
    // sub0: index of this chunk's first matrix inside the Gram super-chunk (W and gscale are laid out per super-chunk)
    const int r = m < n ? m : n;
    const int side = m <= n ? 0 : 1;
    const bool no_gram = small_path(h, m, n) || direct;  // one-sided Jacobi on the matrix itself
    float2* W = reinterpret_cast<float2*>(ws + L.W) + (no_gram ? 0 : (size_t)sub0 * r * r);
    int32_t* perm = reinterpret_cast<int32_t*>(ws + L.perm);
    float* inv = reinterpret_cast<float*>(ws + L.inv);
    float* gscale = reinterpret_cast<float*>(ws + L.gscale) + sub0;
    int32_t* sweeps = reinterpret_cast<int32_t*>(ws + L.sweeps);
    int32_t* done = reinterpret_cast<int32_t*>(ws + L.done);
    unsigned* offmax = reinterpret_cast<unsigned*>(ws + L.offmax);
    int32_t* active = reinterpret_cast<int32_t*>(ws + L.active);
    int32_t* nonfinite = reinterpret_cast<int32_t*>(ws + L.nonfinite);
    float* norm2 = reinterpret_cast<float*>(ws + L.norm2);
    int rc;
    StageTimer tm(h);
    if (!gram_done) VK_CUDA(h, cudaMemsetAsync(nonfinite, 0, 4, h->stream));
    if (no_gram) {
        const int Llong = m < n ? n : m;
        const JacobiPlan p = vk_jacobi_plan(h, r, Llong, Llong + r);
        tm.mark(0);
        if ((rc = vk_launch_pack_small(h, A, B, m, n, W, p.ld, gscale, nonfinite))) return rc;
        if ((rc = vk_launch_jacobi(h, W, B, p, sweeps, done, offmax, active))) return rc;
        tm.mark(1);
        if ((rc = vk_launch_select(h, W, B, r, p.ldot, p.ld, gscale, 0, fixed_rank, decorrelation, kmax, perm, inv, S,
                                   ranks, stats, sweeps, done)))
            return rc;
        tm.mark(2);
        if ((rc = vk_launch_factors_small(h, W, p.ld, B, m, n, kmax, perm, inv, ranks, U, Vt))) return rc;
        tm.mark(3);
        tm.collect(4, 0, 1);
        tm.collect(2, 1, 2);
        tm.collect(3, 2, 3);
    } else {
        const JacobiPlan p = vk_jacobi_plan(h, r, r, r);
        tm.mark(0);
        if (!gram_done && (rc = gram_stage(h, A, B, m, n, W, gscale, nonfinite))) return rc;
        tm.mark(1);
        // fixed small rank: blocked subspace iteration first; the full Jacobi solver only sees what it left unsolved
        const bool qr = use_qr(h, m, n, fixed_rank) && !force_jacobi;
        if (qr) {
            if ((rc = vk_launch_eigqr(h, W, B, r, p.ld, ws + L.eig, sweeps, done, fixed_rank, decorrelation))) return rc;
        } else {
            const bool fast = h->topk != 1 && vk_topk_supported(r, fixed_rank, h->topk == 2);
            if (fast && (rc = vk_launch_topk(h, W, B, r, fixed_rank, done, sweeps))) return rc;
            if ((rc = vk_launch_jacobi(h, W, B, p, sweeps, done, offmax, active, fast))) return rc;
        }
        tm.mark(2);
        if ((rc = vk_launch_select(h, W, B, r, p.ldot, p.ld, gscale, 1, fixed_rank, decorrelation, kmax, perm, inv, S,
                                   ranks, stats, sweeps, done)))
            return rc;
        tm.mark(3);
        float2* xbuf = (side == 0 && vk_cgemm_tc_supported(m, n, kmax)) ? reinterpret_cast<float2*>(ws + L.xbuf) : nullptr;
        if ((rc = vk_launch_factors_gram(h, A, W, B, m, n, side, kmax, perm, inv, ranks, norm2, U, S, Vt, stats, xbuf)))
            return rc;
        tm.mark(4);
        tm.collect(0, 0, 1);
        tm.collect(1, 1, 2);
        tm.collect(2, 2, 3);
        tm.collect(3, 3, 4);
    }
    // poll: non-finite input, and whether the direct eigensolver gave a matrix up (QL iteration limit, rotation store)
    const bool qr_used = !no_gram && use_qr(h, m, n, fixed_rank) && !force_jacobi;
    if (h->check_finite || qr_used) {
        h->h_poll[2] = 0;
        if (qr_used) {
            if ((rc = vk_launch_count_not_done(h, done, B, active))) return rc;
            VK_CUDA(h, cudaMemcpyAsync(h->h_poll + 2, active, 4, cudaMemcpyDeviceToHost, h->stream));
        }
        VK_CUDA(h, cudaMemcpyAsync(h->h_poll + 1, nonfinite, 4, cudaMemcpyDeviceToHost, h->stream));
        VK_CUDA(h, cudaStreamSynchronize(h->stream));
        if (h->check_finite && h->h_poll[1]) return vk_fail(h, VK_ENONFINITE, "input contains NaN or Inf");
        if (h->h_poll[2] > 0) {
            // rare: redo this pass with the cyclic Jacobi solver (the Gram matrices were overwritten: recompute them)
            h->eig_fallbacks++;
            return compress_chunk(h, A, B, m, n, fixed_rank, decorrelation, kmax, U, S, Vt, ranks, stats, ws, L, sub0, false,
                                  true);
        }
    }
    return VK_OK;
}

// Matrices whose retained singular values reach below illcond_thr * sigma_1 cannot be resolved through a float32 Gram
// matrix (lambda is only known to ~1e-7 lambda_max). They are rare in this application (the truncation normally stops
// far above that), so they are simply done again by the small-matrix method - one-sided Jacobi on the matrix itself,
// vectors streamed from global memory - in sub-batches gathered into a second workspace, and their results replace the
// first ones. One 4-byte poll per call when nothing is flagged.
int redo_ill_conditioned(vk_context* h, const float2* A, int B, int m, int n, int fixed_rank, double decorrelation,
                         int kmax, float2* U, float* S, float2* Vt, int32_t* ranks, float* stats) {
    int rc;
    // the flags (and everything else of this path) live in a second workspace: the first is sized per internal pass
    const size_t fbytes = align_up((size_t)B * 4) + 256;
    if ((rc = ensure(h, &h->ws2, &h->ws2_bytes, fbytes))) return rc;
    int32_t* flags = static_cast<int32_t*>(h->ws2);
    int32_t* count = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(h->ws2) + align_up((size_t)B * 4));
    h->h_poll[3] = 0;
    if (h->illcond_thr > 0.f) {
        if ((rc = vk_launch_flag_illcond(h, S, ranks, B, kmax, h->illcond_thr, flags, count))) return rc;
        VK_CUDA(h, cudaMemcpyAsync(h->h_poll + 3, count, 4, cudaMemcpyDeviceToHost, h->stream));
    }
    // ... plus the matrices whose Gram trace left the float32-safe range (h->bad, set by the normalisation kernel)
    h->h_poll[4] = 0;
    if (h->bad) VK_CUDA(h, cudaMemcpyAsync(h->h_poll + 4, h->bad + B, 4, cudaMemcpyDeviceToHost, h->stream));
    VK_CUDA(h, cudaStreamSynchronize(h->stream));
    const int nflag = h->h_poll[3], nbad = h->h_poll[4];
    if (nflag <= 0 && nbad <= 0) return VK_OK;
    std::vector<int32_t> hf((size_t)B, 0), hb((size_t)B, 0);
    if (nflag > 0) VK_CUDA(h, cudaMemcpy(hf.data(), flags, (size_t)B * 4, cudaMemcpyDeviceToHost));
    if (nbad > 0) VK_CUDA(h, cudaMemcpy(hb.data(), h->bad, (size_t)B * 4, cudaMemcpyDeviceToHost));
    std::vector<int> idx;
    for (int b = 0; b < B; ++b)
        if (hf[b] || hb[b]) idx.push_back(b);
    if (idx.empty()) return VK_OK;
    const size_t bA = (size_t)m * n * 8, bU = (size_t)m * kmax * 8, bS = (size_t)kmax * 4, bV = (size_t)kmax * n * 8;
    const size_t per = align_up(bA) + align_up(bU) + align_up(bS) + align_up(bV) + 512 +
                       ws_layout(h, 1, m, n, kmax, 1, false, true).total;
    int cap = (int)(((size_t)2 << 30) / per);
    if (cap < 1) cap = 1;
    if (cap > (int)idx.size()) cap = (int)idx.size();
    const WsLayout Ld = ws_layout(h, cap, m, n, kmax, cap, false, true);
    const size_t oA = 0, oU = oA + align_up(bA * cap), oS = oU + align_up(bU * cap), oV = oS + align_up(bS * cap),
                 oR = oV + align_up(bV * cap), oT = oR + align_up((size_t)cap * 4), oW = oT + align_up((size_t)cap * 16);
    if ((rc = ensure(h, &h->ws2, &h->ws2_bytes, oW + Ld.total))) return rc;
    unsigned char* p = static_cast<unsigned char*>(h->ws2);
    float2* A2 = reinterpret_cast<float2*>(p + oA);
    float2* U2 = reinterpret_cast<float2*>(p + oU);
    float* S2 = reinterpret_cast<float*>(p + oS);
    float2* V2 = reinterpret_cast<float2*>(p + oV);
    int32_t* R2 = reinterpret_cast<int32_t*>(p + oR);
    float* T2 = reinterpret_cast<float*>(p + oT);
    cudaStream_t st = h->stream;
    for (size_t s0 = 0; s0 < idx.size(); s0 += cap) {
        const int nb = (int)std::min((size_t)cap, idx.size() - s0);
        for (int i = 0; i < nb; ++i)
            VK_CUDA(h, cudaMemcpyAsync(A2 + (size_t)i * m * n, A + (size_t)idx[s0 + i] * m * n, bA, cudaMemcpyDeviceToDevice, st));
        if ((rc = compress_chunk(h, A2, nb, m, n, fixed_rank, decorrelation, kmax, U2, S2, V2, R2, T2, p + oW, Ld, 0, false,
                                 false, true)))
            return rc;
        for (int i = 0; i < nb; ++i) {
            const size_t b = (size_t)idx[s0 + i];
            VK_CUDA(h, cudaMemcpyAsync(U + b * m * kmax, U2 + (size_t)i * m * kmax, bU, cudaMemcpyDeviceToDevice, st));
            VK_CUDA(h, cudaMemcpyAsync(S + b * kmax, S2 + (size_t)i * kmax, bS, cudaMemcpyDeviceToDevice, st));
            VK_CUDA(h, cudaMemcpyAsync(Vt + b * kmax * n, V2 + (size_t)i * kmax * n, bV, cudaMemcpyDeviceToDevice, st));
            VK_CUDA(h, cudaMemcpyAsync(ranks + b, R2 + i, 4, cudaMemcpyDeviceToDevice, st));
            VK_CUDA(h, cudaMemcpyAsync(stats + b * 4, T2 + (size_t)i * 4, 16, cudaMemcpyDeviceToDevice, st));
        }
        h->illcond_redone += nb;
    }
    return VK_OK;
}

int check_common(vk_context* h, int B, int m, int n, int kmax) {
    if (!h) return VK_EINVAL;
    if (B < 0 || m < 1 || n < 1 || kmax < 1) return vk_fail(h, VK_EINVAL, "bad shape: need B >= 0, m, n, kmax >= 1");
    const int r = m < n ? m : n;
    if (r > VK_MAX_R) return vk_fail(h, VK_EINVAL, "min(m, n) > 2048 is not supported");
    return VK_OK;
}

}  // namespace

extern "C" {

const char* vk_version(void) { return "visco_b200 0.1.0 (sm_100a)"; }

int vk_create(vk_handle* out, int device) {
    if (!out) return VK_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return VK_ECUDA;
    vk_context* h = new (std::nothrow) vk_context();
    if (!h) return VK_ENOMEM;
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete h;
        return VK_ECUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        h->num_sms = prop.multiProcessorCount;
        if (prop.major != 10) {
            delete h;
            return VK_ECUDA;  // sm_100a only: no other architecture is built into this library
        }
    }
    if (cudaMallocHost(reinterpret_cast<void**>(&h->h_poll), 64) != cudaSuccess) {
        delete h;
        return VK_ENOMEM;
    }
    std::memset(h->h_poll, 0, 64);
    if (cudaMalloc(&h->d_scratch, 256) != cudaSuccess) {
        cudaFreeHost(h->h_poll);
        delete h;
        return VK_ENOMEM;
    }
    for (auto& e : h->ev) cudaEventCreate(&e);
    for (auto& e : h->eig_ev) cudaEventCreate(&e);
    *out = h;
    return VK_OK;
}

int vk_destroy(vk_handle h) {
    if (!h) return VK_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->ws) cudaFree(h->ws);
    if (h->stage) cudaFree(h->stage);
    if (h->ws2) cudaFree(h->ws2);
    if (h->bad) cudaFree(h->bad);
    if (h->h_poll) cudaFreeHost(h->h_poll);
    if (h->d_scratch) cudaFree(h->d_scratch);
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : h->eig_ev)
        if (e) cudaEventDestroy(e);
    for (auto& s : h->sub)
        if (s) cudaStreamDestroy(s);
    for (auto& e : h->sub_ev)
        if (e) cudaEventDestroy(e);
    if (h->fork_ev) cudaEventDestroy(h->fork_ev);
    if (h->tail_stream) cudaStreamDestroy(h->tail_stream);
    for (int i = 0; i < 2; ++i)
        if (h->tail_ev[i]) cudaEventDestroy(h->tail_ev[i]);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->copy_stream2) cudaStreamDestroy(h->copy_stream2);
    for (auto& e : h->host_ev)
        if (e) cudaEventDestroy(e);
    delete h;
    return VK_OK;
}

const char* vk_last_error(vk_handle h) { return h ? h->err.c_str() : "null handle"; }

int vk_set_stream(vk_handle h, void* s) {
    if (!h) return VK_EINVAL;
    h->stream = reinterpret_cast<cudaStream_t>(s);
    return VK_OK;
}

int vk_sync(vk_handle h) {
    if (!h) return VK_EINVAL;
    VK_CUDA(h, cudaSetDevice(h->device));
    VK_CUDA(h, cudaStreamSynchronize(h->stream));
    return VK_OK;
}

int vk_set_option(vk_handle h, const char* key, double v) {
    if (!h || !key) return VK_EINVAL;
    const std::string k(key);
    if (k == "jacobi_tol")
        h->jacobi_tol = (float)v;
    else if (k == "max_sweeps")
        h->max_sweeps = (int)v;
    else if (k == "gram_impl")
        h->gram_impl = (int)v;
    else if (k == "check_finite")
        h->check_finite = (int)v;
    else if (k == "check_every")
        h->check_every = (int)v;
    else if (k == "jacobi_groups")
        h->jacobi_groups = (int)v;
    else if (k == "jacobi_bsz")
        h->jacobi_bsz = (int)v;
    else if (k == "stage_timing")
        h->stage_timing = (int)v;
    else if (k == "topk")
        h->topk = (int)v;
    else if (k == "gemm_impl")
        h->gemm_impl = (int)v;
    else if (k == "recon_tc_impl")
        h->recon_tc_impl = (int)v;
    else if (k == "recon_generic")
        h->recon_generic = (int)v;
    else if (k == "small_reg")
        h->small_reg = (int)v;
    else if (k == "jacobi_generic")
        h->jacobi_generic = (int)v;
    else if (k == "eig_impl")
        h->eig_impl = (int)v;
    else if (k == "tridiag_impl")
        h->tridiag_impl = (int)v;
    else if (k == "tridiag_nts")
        h->tridiag_nts = (int)v;
    else if (k == "tridiag_small_rs")
        h->tridiag_small_rs = (int)v;
    else if (k == "bisect_impl")
        h->bisect_impl = (int)v;
    else if (k == "gram_small")
        h->gram_small = (int)v;
    else if (k == "tridiag_pf")
        h->tridiag_pf = (int)v;
    else if (k == "split_variant")
        h->split_variant = (int)v;
    else if (k == "tail_split")
        h->tail_split = (int)v;
    else if (k == "factors_impl")
        h->factors_impl = (int)v;
    else if (k == "small_impl")
        h->small_impl = (int)v;
    else if (k == "tridiag_variant")
        h->tridiag_variant = (int)v;
    else if (k == "eigvec_impl")
        h->eigvec_impl = (int)v;
    else if (k == "ql_maxit")
        h->ql_maxit = (int)v;
    else if (k == "illcond_thr")
        h->illcond_thr = (float)v;
    else if (k == "chunk")
        h->chunk = (int)v;
    else
        return vk_fail(h, VK_EINVAL, "unknown option " + k);
    return VK_OK;
}

size_t vk_workspace_bytes(vk_handle h, int B, int m, int n, int kmax) {
    if (B <= 0 || m < 1 || n < 1 || kmax < 1) return 0;
    // the rank rule is not known here: report the larger of the two eigensolver configurations
    size_t need = 0;
    for (int q = 0; q < 2; ++q) {
        const bool qr = q == 1;
        if (qr && !use_qr(h, m, n, 0)) continue;
        int chunk = (h && h->chunk > 0) ? h->chunk : auto_chunk(h, B, m, n, qr);
        if (chunk > B) chunk = B;
        const size_t t = ws_layout(h, chunk, m, n, kmax, gram_chunk(h, B, chunk, m, n), qr).total;
        if (t > need) need = t;
    }
    return need;
}

int vk_uses_small_path(int m, int n) { return (m >= 1 && n >= 1 && small_path(nullptr, m, n)) ? 1 : 0; }
int vk_gram_uses_tcgen05(int m, int n, int side) { return vk_gram_tc_supported(m, n, side) ? 1 : 0; }

int vk_compress_batched(vk_handle h, const void* A, int B, int m, int n, int fixed_rank, double decorrelation, int kmax,
                        void* U, float* S, void* Vt, int32_t* ranks, float* stats, void* ws, size_t ws_bytes) {
    int rc = check_common(h, B, m, n, kmax);
    if (rc) return rc;
    if (B == 0) return VK_OK;
    if (!A || !U || !S || !Vt || !ranks || !stats) return vk_fail(h, VK_EINVAL, "null buffer");
    const int r = m < n ? m : n;
    const int need_k = fixed_rank > 0 ? (fixed_rank < r ? fixed_rank : r) : r;
    if (kmax < need_k) return vk_fail(h, VK_EINVAL, "kmax smaller than the largest possible rank");
    if (kmax > r) return vk_fail(h, VK_EINVAL, "kmax larger than min(m, n)");
    if (decorrelation < 0.0 || !(decorrelation == decorrelation))
        return vk_fail(h, VK_EINVAL, "decorrelation must be >= 0");
    VK_CUDA(h, cudaSetDevice(h->device));
    InCompress in_compress(h->device);
    const bool qr = use_qr(h, m, n, fixed_rank);
    int chunk = h->chunk > 0 ? h->chunk : auto_chunk(h, B, m, n, qr);
    if (chunk > B) chunk = B;
    const int gchunk = gram_chunk(h, B, chunk, m, n);
    const WsLayout L = ws_layout(h, chunk, m, n, kmax, gchunk, qr);
    unsigned char* wsp = static_cast<unsigned char*>(ws);
    if (!wsp) {
        if ((rc = ensure(h, &h->ws, &h->ws_bytes, L.total))) return rc;
        wsp = static_cast<unsigned char*>(h->ws);
    } else if (ws_bytes < L.total) {
        return vk_fail(h, VK_EINVAL, "workspace too small: need " + std::to_string(L.total) + " bytes");
    }
    for (int i = 0; i < 6; ++i) h->stage_ms[i] = 0.f;
    for (int i = 0; i < 5; ++i) h->eig_ms[i] = 0.f;
    cudaEvent_t e0 = h->ev[6], e1 = h->ev[7];
    if (h->stage_timing) cudaEventRecord(e0, h->stream);
    const float2* Ap = static_cast<const float2*>(A);
    float2* Up = static_cast<float2*>(U);
    float2* Vp = static_cast<float2*>(Vt);
    const bool gram_path = !small_path(h, m, n);
    if (gram_path) {
        // per-matrix flags of Gram traces outside the float32-safe range (input beyond ~1e19 or below ~1e-19): those
        // matrices are done again without a Gram product after the main loop
        void* pb = h->bad;
        if ((rc = ensure(h, &pb, &h->bad_bytes, ((size_t)B + 1) * 4))) return rc;
        h->bad = static_cast<int32_t*>(pb);
        VK_CUDA(h, cudaMemsetAsync(h->bad, 0, ((size_t)B + 1) * 4, h->stream));
    }
    for (int g0 = 0; g0 < B; g0 += gchunk) {
        const int ng = (B - g0) < gchunk ? (B - g0) : gchunk;
        if (gram_path) {
            // Gram + normalisation for the whole super-chunk in one launch each
            int32_t* nonfinite = reinterpret_cast<int32_t*>(wsp + L.nonfinite);
            VK_CUDA(h, cudaMemsetAsync(nonfinite, 0, 4, h->stream));
            StageTimer tg(h);
            tg.mark(0);
            rc = gram_stage(h, Ap + (size_t)g0 * m * n, ng, m, n, reinterpret_cast<float2*>(wsp + L.W),
                            reinterpret_cast<float*>(wsp + L.gscale), nonfinite, h->bad + g0, h->bad + B);
            if (rc) return rc;
            tg.mark(1);
            tg.collect(0, 0, 1);
        }
        for (int s0 = 0; s0 < ng; s0 += chunk) {
            const int nb = (ng - s0) < chunk ? (ng - s0) : chunk;
            const int b0 = g0 + s0;
            rc = compress_chunk(h, Ap + (size_t)b0 * m * n, nb, m, n, fixed_rank, decorrelation, kmax,
                                Up + (size_t)b0 * m * kmax, S + (size_t)b0 * kmax, Vp + (size_t)b0 * kmax * n, ranks + b0,
                                stats + (size_t)b0 * 4, wsp, L, s0, gram_path);
            if (rc) return rc;
        }
    }
    if (gram_path) {
        rc = redo_ill_conditioned(h, Ap, B, m, n, fixed_rank, decorrelation, kmax, Up, S, Vp, ranks, stats);
        if (rc) return rc;
    }
    if (h->stage_timing) {
        cudaEventRecord(e1, h->stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&h->stage_ms[5], e0, e1);
    }
    return VK_OK;
}

int vk_reconstruct_batched(vk_handle h, const void* U, const float* S, const void* Vt, const int32_t* ranks, int B,
                           int m, int n, int kmax, void* out) {
    if (!h) return VK_EINVAL;
    if (B < 0 || m < 1 || n < 1 || kmax < 1) return vk_fail(h, VK_EINVAL, "bad shape: need B >= 0, m, n, kmax >= 1");
    if (B == 0) return VK_OK;
    if (!U || !S || !Vt || !out) return vk_fail(h, VK_EINVAL, "null buffer");
    VK_CUDA(h, cudaSetDevice(h->device));
    return vk_launch_reconstruct(h, static_cast<const float2*>(U), S, static_cast<const float2*>(Vt), ranks, B, m, n,
                                 kmax, static_cast<float2*>(out));
}

int vk_find_n_decorrelation_batched(vk_handle h, const float* S, int B, int r, double decorrelation, int32_t* ranks) {
    if (!h) return VK_EINVAL;
    if (B < 0 || r < 1 || !S || !ranks) return vk_fail(h, VK_EINVAL, "bad argument");
    if (B == 0) return VK_OK;
    VK_CUDA(h, cudaSetDevice(h->device));
    return vk_launch_find_n(h, S, B, r, decorrelation, ranks);
}

int vk_compress_host(vk_handle h, const void* A, int B, int m, int n, int fixed_rank, double decorrelation, int kmax,
                     void* U, float* S, void* Vt, int32_t* ranks, float* stats) {
    int rc = check_common(h, B, m, n, kmax);
    if (rc) return rc;
    if (B == 0) return VK_OK;
    if (!A || !U || !S || !Vt || !ranks || !stats) return vk_fail(h, VK_EINVAL, "null buffer");
    VK_CUDA(h, cudaSetDevice(h->device));
    const size_t bA = align_up((size_t)B * m * n * 8), bU = align_up((size_t)B * m * kmax * 8),
                 bS = align_up((size_t)B * kmax * 4), bV = align_up((size_t)B * kmax * n * 8),
                 bR = align_up((size_t)B * 4), bT = align_up((size_t)B * 16);
    if ((rc = ensure(h, &h->stage, &h->stage_bytes, bA + bU + bS + bV + bR + bT))) return rc;
    unsigned char* p = static_cast<unsigned char*>(h->stage);
    void* dA = p;
    void* dU = p + bA;
    float* dS = reinterpret_cast<float*>(p + bA + bU);
    void* dV = p + bA + bU + bS;
    int32_t* dR = reinterpret_cast<int32_t*>(p + bA + bU + bS + bV);
    float* dT = reinterpret_cast<float*>(p + bA + bU + bS + bV + bR);
    // Pipelined in up to VK_HOST_CHUNKS sub-batches: all H2D copies are queued back to back on a copy stream, each
    // followed by an event; the compute stream waits for sub-batch i's event only, so the upload of sub-batch i+1 runs
    // under the factorisation of sub-batch i (the factorisation polls the device, i.e. blocks this host thread, which is
    // why the copies are queued up front). Factors return on the copy stream as soon as their sub-batch is done.
    if (!h->copy_stream) VK_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->copy_stream2) VK_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream2, cudaStreamNonBlocking));
    for (auto& e : h->host_ev)
        if (!e) VK_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // only large batches are split: a sub-batch must still fill the GPU for the whole Jacobi iteration (measured: four
    // sub-batches of 28 matrices of 256 x 1024 take 2.3x the time of one batch of 112)
    int nch = (int)(((size_t)B * m * n * 8) / ((size_t)512 << 20));
    if (nch > VK_HOST_CHUNKS) nch = VK_HOST_CHUNKS;
    if (nch > B) nch = B;
    if (nch < 1) nch = 1;
    VK_CUDA(h, cudaEventRecord(h->host_ev[2 * VK_HOST_CHUNKS], h->stream));        // staging buffers are free again
    VK_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->host_ev[2 * VK_HOST_CHUNKS], 0));
    const char* Ah = static_cast<const char*>(A);
    for (int i = 0; i < nch; ++i) {
        const int b0 = (int)((long long)B * i / nch), b1 = (int)((long long)B * (i + 1) / nch);
        VK_CUDA(h, cudaMemcpyAsync(static_cast<char*>(dA) + (size_t)b0 * m * n * 8, Ah + (size_t)b0 * m * n * 8,
                                   (size_t)(b1 - b0) * m * n * 8, cudaMemcpyHostToDevice, h->copy_stream));
        VK_CUDA(h, cudaEventRecord(h->host_ev[i], h->copy_stream));
    }
    for (int i = 0; i < nch; ++i) {
        const int b0 = (int)((long long)B * i / nch), b1 = (int)((long long)B * (i + 1) / nch), nb = b1 - b0;
        VK_CUDA(h, cudaStreamWaitEvent(h->stream, h->host_ev[i], 0));
        rc = vk_compress_batched(h, static_cast<char*>(dA) + (size_t)b0 * m * n * 8, nb, m, n, fixed_rank, decorrelation,
                                 kmax, static_cast<char*>(dU) + (size_t)b0 * m * kmax * 8, dS + (size_t)b0 * kmax,
                                 static_cast<char*>(dV) + (size_t)b0 * kmax * n * 8, dR + b0, dT + (size_t)b0 * 4, nullptr, 0);
        if (rc) return rc;
        VK_CUDA(h, cudaEventRecord(h->host_ev[VK_HOST_CHUNKS + i], h->stream));
        // factors come down on a second copy stream: under the uploads still queued on the first one (full duplex)
        VK_CUDA(h, cudaStreamWaitEvent(h->copy_stream2, h->host_ev[VK_HOST_CHUNKS + i], 0));
        cudaStream_t cs = h->copy_stream2;
        VK_CUDA(h, cudaMemcpyAsync(static_cast<char*>(U) + (size_t)b0 * m * kmax * 8, static_cast<char*>(dU) + (size_t)b0 * m * kmax * 8,
                                   (size_t)nb * m * kmax * 8, cudaMemcpyDeviceToHost, cs));
        VK_CUDA(h, cudaMemcpyAsync(S + (size_t)b0 * kmax, dS + (size_t)b0 * kmax, (size_t)nb * kmax * 4, cudaMemcpyDeviceToHost, cs));
        VK_CUDA(h, cudaMemcpyAsync(static_cast<char*>(Vt) + (size_t)b0 * kmax * n * 8, static_cast<char*>(dV) + (size_t)b0 * kmax * n * 8,
                                   (size_t)nb * kmax * n * 8, cudaMemcpyDeviceToHost, cs));
        VK_CUDA(h, cudaMemcpyAsync(ranks + b0, dR + b0, (size_t)nb * 4, cudaMemcpyDeviceToHost, cs));
        VK_CUDA(h, cudaMemcpyAsync(stats + (size_t)b0 * 4, dT + (size_t)b0 * 4, (size_t)nb * 16, cudaMemcpyDeviceToHost, cs));
    }
    VK_CUDA(h, cudaStreamSynchronize(h->copy_stream));
    VK_CUDA(h, cudaStreamSynchronize(h->copy_stream2));
    VK_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int b = 0; b < B; ++b)
        if (stats[4 * b + 3] == 0.f)
            return vk_fail(h, VK_ENOCONV, "the eigensolver did not converge for matrix " + std::to_string(b));
    return VK_OK;
}

int vk_reconstruct_host(vk_handle h, const void* U, const float* S, const void* Vt, const int32_t* ranks, int B, int m,
                        int n, int kmax, void* out) {
    if (!h) return VK_EINVAL;
    if (B < 0 || m < 1 || n < 1 || kmax < 1) return vk_fail(h, VK_EINVAL, "bad shape: need B >= 0, m, n, kmax >= 1");
    if (B == 0) return VK_OK;
    if (!U || !S || !Vt || !out) return vk_fail(h, VK_EINVAL, "null buffer");
    VK_CUDA(h, cudaSetDevice(h->device));
    int rc;
    const size_t bO = align_up((size_t)B * m * n * 8), bU = align_up((size_t)B * m * kmax * 8),
                 bS = align_up((size_t)B * kmax * 4), bV = align_up((size_t)B * kmax * n * 8),
                 bR = align_up((size_t)B * 4);
    if ((rc = ensure(h, &h->stage, &h->stage_bytes, bO + bU + bS + bV + bR))) return rc;
    unsigned char* p = static_cast<unsigned char*>(h->stage);
    void* dO = p;
    void* dU = p + bO;
    float* dS = reinterpret_cast<float*>(p + bO + bU);
    void* dV = p + bO + bU + bS;
    int32_t* dR = reinterpret_cast<int32_t*>(p + bO + bU + bS + bV);
    // sub-batches with the two copy directions on their own streams: the factors of sub-batch i+1 go up (copy stream 2)
    // while sub-batch i is reconstructed and the matrices of sub-batch i-1 come down (copy stream 1) - PCIe is full
    // duplex, and with large ranks the factors are as big as the matrices (MeerKAT shard: 15.8 GB up, 17.4 GB down)
    if (!h->copy_stream) VK_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->copy_stream2) VK_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream2, cudaStreamNonBlocking));
    for (auto& e : h->host_ev)
        if (!e) VK_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const int nch = B < VK_HOST_CHUNKS ? B : VK_HOST_CHUNKS;
    VK_CUDA(h, cudaEventRecord(h->host_ev[2 * VK_HOST_CHUNKS], h->stream));  // staging buffers are free again
    VK_CUDA(h, cudaStreamWaitEvent(h->copy_stream2, h->host_ev[2 * VK_HOST_CHUNKS], 0));
    VK_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->host_ev[2 * VK_HOST_CHUNKS], 0));
    for (int i = 0; i < nch; ++i) {
        const int b0 = (int)((long long)B * i / nch), b1 = (int)((long long)B * (i + 1) / nch), nb = b1 - b0;
        cudaStream_t up = h->copy_stream2;
        VK_CUDA(h, cudaMemcpyAsync(static_cast<char*>(dU) + (size_t)b0 * m * kmax * 8, static_cast<const char*>(U) + (size_t)b0 * m * kmax * 8,
                                   (size_t)nb * m * kmax * 8, cudaMemcpyHostToDevice, up));
        VK_CUDA(h, cudaMemcpyAsync(dS + (size_t)b0 * kmax, S + (size_t)b0 * kmax, (size_t)nb * kmax * 4, cudaMemcpyHostToDevice, up));
        VK_CUDA(h, cudaMemcpyAsync(static_cast<char*>(dV) + (size_t)b0 * kmax * n * 8, static_cast<const char*>(Vt) + (size_t)b0 * kmax * n * 8,
                                   (size_t)nb * kmax * n * 8, cudaMemcpyHostToDevice, up));
        if (ranks) VK_CUDA(h, cudaMemcpyAsync(dR + b0, ranks + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, up));
        VK_CUDA(h, cudaEventRecord(h->host_ev[VK_HOST_CHUNKS + i], up));
    }
    for (int i = 0; i < nch; ++i) {
        const int b0 = (int)((long long)B * i / nch), b1 = (int)((long long)B * (i + 1) / nch), nb = b1 - b0;
        VK_CUDA(h, cudaStreamWaitEvent(h->stream, h->host_ev[VK_HOST_CHUNKS + i], 0));
        rc = vk_reconstruct_batched(h, static_cast<char*>(dU) + (size_t)b0 * m * kmax * 8, dS + (size_t)b0 * kmax,
                                    static_cast<char*>(dV) + (size_t)b0 * kmax * n * 8, ranks ? dR + b0 : nullptr, nb, m, n, kmax,
                                    static_cast<char*>(dO) + (size_t)b0 * m * n * 8);
        if (rc) return rc;
        VK_CUDA(h, cudaEventRecord(h->host_ev[i], h->stream));
        VK_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->host_ev[i], 0));
        VK_CUDA(h, cudaMemcpyAsync(static_cast<char*>(out) + (size_t)b0 * m * n * 8, static_cast<char*>(dO) + (size_t)b0 * m * n * 8,
                                   (size_t)nb * m * n * 8, cudaMemcpyDeviceToHost, h->copy_stream));
    }
    VK_CUDA(h, cudaStreamSynchronize(h->copy_stream2));
    VK_CUDA(h, cudaStreamSynchronize(h->copy_stream));
    VK_CUDA(h, cudaStreamSynchronize(h->stream));
    return VK_OK;
}

int vk_gram_batched(vk_handle h, const void* A, int B, int m, int n, int side, int impl, void* W) {
    if (!h) return VK_EINVAL;
    if (B < 0 || m < 1 || n < 1 || (side != 0 && side != 1) || !A || !W) return vk_fail(h, VK_EINVAL, "bad argument");
    if (B == 0) return VK_OK;
    VK_CUDA(h, cudaSetDevice(h->device));
    const bool tc = impl == 2 || (impl == 0 && vk_gram_tc_supported(m, n, side));
    if (tc) {
        if (!vk_gram_tc_supported(m, n, side)) return vk_fail(h, VK_EINVAL, "tcgen05 Gram does not support this shape");
        return vk_launch_gram_tc(h, static_cast<const float2*>(A), B, m, n, static_cast<float2*>(W));
    }
    return vk_launch_gram_simt(h, static_cast<const float2*>(A), B, m, n, side, static_cast<float2*>(W));
}

int vk_eigh_jacobi_batched(vk_handle h, void* W, int B, int r, float* lambda, int32_t* info) {
    if (!h) return VK_EINVAL;
    if (B < 0 || r < 1 || r > VK_MAX_R || !W || !lambda || !info) return vk_fail(h, VK_EINVAL, "bad argument");
    if (B == 0) return VK_OK;
    VK_CUDA(h, cudaSetDevice(h->device));
    int rc;
    // bookkeeping scratch only (the caller owns W)
    size_t off = 0;
    const size_t o_perm = off; off += align_up((size_t)B * r * 4);
    const size_t o_inv = off; off += align_up((size_t)B * r * 4);
    const size_t o_g = off; off += align_up((size_t)B * 4);
    const size_t o_sw = off; off += align_up((size_t)B * 4);
    const size_t o_dn = off; off += align_up((size_t)B * 4);
    const size_t o_om = off; off += align_up((size_t)B * 4);
    const size_t o_rk = off; off += align_up((size_t)B * 4);
    const size_t o_st = off; off += align_up((size_t)B * 16);
    const size_t o_ac = off; off += 256;
    const size_t o_nf = off; off += 256;
    const bool qr = h->eig_impl != 1 && vk_eigqr_supported(r);
    const size_t o_eig = off;
    if (qr) off += align_up(vk_eigqr_scratch_bytes(B, r));
    if ((rc = ensure(h, &h->ws, &h->ws_bytes, off))) return rc;
    unsigned char* ws = static_cast<unsigned char*>(h->ws);
    int32_t* perm = reinterpret_cast<int32_t*>(ws + o_perm);
    float* inv = reinterpret_cast<float*>(ws + o_inv);
    float* gscale = reinterpret_cast<float*>(ws + o_g);
    int32_t* sweeps = reinterpret_cast<int32_t*>(ws + o_sw);
    int32_t* done = reinterpret_cast<int32_t*>(ws + o_dn);
    unsigned* offmax = reinterpret_cast<unsigned*>(ws + o_om);
    int32_t* ranks = reinterpret_cast<int32_t*>(ws + o_rk);
    float* stats = reinterpret_cast<float*>(ws + o_st);
    int32_t* active = reinterpret_cast<int32_t*>(ws + o_ac);
    int32_t* nonfinite = reinterpret_cast<int32_t*>(ws + o_nf);
    float2* Wp = static_cast<float2*>(W);
    VK_CUDA(h, cudaMemsetAsync(nonfinite, 0, 4, h->stream));
    if ((rc = vk_launch_gram_normalise(h, Wp, B, r, gscale, nonfinite))) return rc;
    if (qr) {
        if ((rc = vk_launch_eigqr(h, Wp, B, r, r, ws + o_eig, sweeps, done))) return rc;
    } else {
        const JacobiPlan p = vk_jacobi_plan(h, r, r, r);
        if ((rc = vk_launch_jacobi(h, Wp, B, p, sweeps, done, offmax, active))) return rc;
    }
    // mode 2: report the eigenvalues themselves (norm * trace scale), sorted descending
    if ((rc = vk_launch_select(h, Wp, B, r, r, r, gscale, 2, 0, 0.f, r, perm, inv, lambda, ranks, stats, sweeps, done)))
        return rc;
    return vk_launch_pack_info(h, sweeps, done, B, info);
}

int vk_svd_jacobi_small_batched(vk_handle h, const void* A, int B, int m, int n, void* U, float* S, void* Vt,
                                int32_t* info) {
    if (!h) return VK_EINVAL;
    if (B < 0 || m < 1 || n < 1 || !A || !U || !S || !Vt) return vk_fail(h, VK_EINVAL, "bad argument");
    if (!small_eligible(m, n)) return vk_fail(h, VK_EINVAL, "shape is not eligible for the small-matrix path");
    if (B == 0) return VK_OK;
    VK_CUDA(h, cudaSetDevice(h->device));
    const int r = m < n ? m : n;
    int rc;
    // this entry point IS the one-sided Jacobi path, whatever the routing option says
    const int save_impl = h->small_impl;
    h->small_impl = 1;
    struct Restore {
        vk_context* h;
        int v;
        ~Restore() { h->small_impl = v; }
    } restore{h, save_impl};
    // ranks + stats scratch live behind the regular workspace
    const WsLayout L = ws_layout(h, B, m, n, r);
    const size_t extra = align_up((size_t)B * 4) + align_up((size_t)B * 16);
    if ((rc = ensure(h, &h->ws, &h->ws_bytes, L.total + extra))) return rc;
    unsigned char* ws = static_cast<unsigned char*>(h->ws);
    int32_t* ranks = reinterpret_cast<int32_t*>(ws + L.total);
    float* stats = reinterpret_cast<float*>(ws + L.total + align_up((size_t)B * 4));
    const int save = h->check_finite;
    h->check_finite = 0;
    rc = compress_chunk(h, static_cast<const float2*>(A), B, m, n, 0, 0.f, r, static_cast<float2*>(U), S,
                        static_cast<float2*>(Vt), ranks, stats, ws, L);
    h->check_finite = save;
    if (rc) return rc;
    if (info) {
        const int32_t* sweeps = reinterpret_cast<const int32_t*>(ws + L.sweeps);
        const int32_t* done = reinterpret_cast<const int32_t*>(ws + L.done);
        if ((rc = vk_launch_pack_info(h, sweeps, done, B, info))) return rc;
    }
    return VK_OK;
}

int vk_synth_fill(vk_handle h, void* A, int nbl_local, int ncorr, int m, int n, int bl_offset, int nbl_total,
                  uint64_t seed) {
    if (!h) return VK_EINVAL;
    if (!A || nbl_local < 0 || ncorr < 1 || m < 1 || n < 1 || nbl_total < 1) return vk_fail(h, VK_EINVAL, "bad argument");
    if (nbl_local == 0) return VK_OK;
    VK_CUDA(h, cudaSetDevice(h->device));
    return vk_launch_synth(h, static_cast<float2*>(A), nbl_local, ncorr, m, n, bl_offset, nbl_total, seed);
}

int64_t vk_launch_count(vk_handle h) { return h ? h->launches : 0; }

int vk_last_eig_ms(vk_handle h, float* t5) {
    if (!h || !t5) return VK_EINVAL;
    for (int i = 0; i < 5; ++i) t5[i] = h->eig_ms[i];
    return VK_OK;
}

int vk_last_stage_ms(vk_handle h, float* t6) {
    if (!h || !t6) return VK_EINVAL;
    for (int i = 0; i < 6; ++i) t6[i] = h->stage_ms[i];
    return VK_OK;
}

}  // extern "C"
