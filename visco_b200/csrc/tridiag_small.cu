// Householder tridiagonalisation of small Hermitian matrices (r <= 64): one or two WARPS per matrix (a 32- or 64-thread
// CTA), the matrix in shared memory. Same contract as the kernels of tridiag.cu (conventions in its header): on exit d, e,
// tau, ph hold the real tridiagonal, the reflector scales and the accumulated sub-diagonal phases, and row j of M right
// of the diagonal holds reflector j. Part of the replacement of the LAPACK cgesdd call behind np.linalg.svd (reference
// compress_ms.py:350) for BASELINE configs[3] (2080 x 4 matrices of 64 x 64) when it runs through the Gram path.
//
// The CTA-per-matrix kernels of tridiag.cu spend 62 Householder steps x 4 barriers x 16 warps on 32 KB of data (4.9 ms
// for 8320 matrices). Here thread t owns column j+1+t of the trailing block; the matrix is Hermitian and kept in full, so
// both the product and the update run down the rows with the thread's column fixed:
//     p_k = tau sum_i conj(a_ik) v_i            a_ik <- a_ik - v_i conj(w_k) - w_i conj(v_k)
// (row-major reads with consecutive lanes: no bank conflicts, v_i / w_i are broadcasts, no reduction per row). Two
// reductions per step (|x|^2 and v^H p) are the only cross-lane traffic. Shared memory (r^2 + 3 r complex numbers) lets
// six 64 x 64 matrices share an SM; one warp each ran at 2.2 ms for the 8320 matrices (1.5 warps per scheduler, latency
// bound), two warps each hide more of it.
#include <cmath>

#include "common.cuh"

namespace {

__device__ __forceinline__ float2 cmul_s(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// NW = warps per column set (1: r <= 33, 2: r <= 64); RS = row splits: RS groups of NW warps share a matrix, group g takes
// the rows j + 1 + g, j + 1 + g + RS, ... of every column loop (the partial products meet in shared memory). One group ran
// at 3 us per Householder step of a 64 x 64 matrix - twelve warps per SM (shared memory allows six matrices), each a chain
// of ~700 dependent instructions per step at 37 % issue utilisation; two groups: 1.75 -> 1.57 ms per 8320 matrices.
template <int NW, int RS>
__global__ void __launch_bounds__(32 * NW * RS)
tridiag_small_kernel(float2* __restrict__ Wall, int r, int ld, size_t wstride, float* __restrict__ dall,
                     float* __restrict__ eall, float* __restrict__ tauall, float2* __restrict__ phall) {
    extern __shared__ float2 ts_sm[];
    __shared__ float red[2 * NW * RS];
    __shared__ float2 pp[RS > 1 ? RS : 1][32 * NW];       // partial products of the row groups
    constexpr int NT = 32 * NW * RS, NC = 32 * NW, NWT = NW * RS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ct = RS > 1 ? tid % NC : tid, grp = RS > 1 ? tid / NC : 0;   // column slot, row group
    const int b = blockIdx.x;
    float2* A = ts_sm;
    float2* vs = A + r * r;
    float2* ws = vs + 64;
    float2* M = Wall + (size_t)b * wstride;
    float* d = dall + (size_t)b * r;
    float* e = eall + (size_t)b * r;
    float* taus = tauall + (size_t)b * r;
    float2* ph = phall + (size_t)b * r;
    auto sync = [&]() {
        if (NWT == 1) __syncwarp();
        else __syncthreads();
    };
    auto block_sum = [&](float v, int slot) {              // (only row group 0 contributes: the others pass 0)
        v = warp_sum(v);
        if (NWT == 1) return v;
        if (lane == 0) red[slot * NWT + warp] = v;
        __syncthreads();
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += red[slot * NWT + w];
        return t;
    };
    for (int idx = tid; idx < r * r; idx += NT) A[idx] = M[(size_t)(idx / r) * ld + idx % r];
    float2 phase = make_float2(1.f, 0.f);
    if (tid == 0) ph[0] = phase;
    sync();
    for (int j = 0; j + 2 < r; ++j) {
        const int k = j + 1 + ct;
        const bool live = k < r;
        const int c = live ? k : r - 1;                    // dead threads read a valid address
        // row j right of the diagonal; a = its conjugate = the column below the diagonal
        float2 a = make_float2(0.f, 0.f);
        if (live) {
            const float2 x = A[j * r + k];
            a = make_float2(x.x, -x.y);
            if (tid == 0) vs[j + 1] = a;                   // alpha, for everybody
        }
        const float tot = block_sum(grp == 0 ? a.x * a.x + a.y * a.y : 0.f, 0);   // (the barrier inside also publishes alpha)
        if (NWT == 1) __syncwarp();
        const float2 alpha = vs[j + 1];
        float tau = 0.f, ej = 0.f;
        float2 v0 = alpha;
        if (tot > 1e-30f) {
            const float xn = sqrtf(tot);
            const float aa = sqrtf(alpha.x * alpha.x + alpha.y * alpha.y);
            float2 p1 = make_float2(1.f, 0.f);
            if (aa > 0.f) p1 = make_float2(alpha.x / aa, alpha.y / aa);
            v0 = make_float2(alpha.x + p1.x * xn, alpha.y + p1.y * xn);
            tau = 1.f / (xn * (xn + aa));
            ej = xn;
            phase = cmul_s(phase, make_float2(-p1.x, -p1.y));  // sub-diagonal element is -p1 * xn
        }
        sync();                                            // alpha has been read
        if (ct == 0) a = v0;
        if (tid == 0) {
            d[j] = A[j * r + j].x;
            taus[j] = tau;
            e[j] = ej;
            ph[j + 1] = phase;
        }
        // the reflector: to shared memory for the broadcasts, and to row j of M right of the diagonal
        if (live && grp == 0) vs[k] = a, M[(size_t)j * ld + k] = a;
        sync();
        if (tau == 0.f) continue;                          // (uniform)
        // p_k = tau sum_{i > j} conj(a_ik) v_i for this thread's column; four independent partial sums
        float2 p;
        {
            float2 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = make_float2(0.f, 0.f);
            int i = j + 1 + grp;
            for (; i + 3 * RS < r; i += 4 * RS) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 vi = vs[i + u * RS];
                    const float2 x = A[(i + u * RS) * r + c];
                    q[u].x = fmaf(x.x, vi.x, fmaf(x.y, vi.y, q[u].x));
                    q[u].y = fmaf(x.x, vi.y, fmaf(-x.y, vi.x, q[u].y));
                }
            }
            for (; i < r; i += RS) {
                const float2 vi = vs[i];
                const float2 x = A[i * r + c];
                q[0].x = fmaf(x.x, vi.x, fmaf(x.y, vi.y, q[0].x));
                q[0].y = fmaf(x.x, vi.y, fmaf(-x.y, vi.x, q[0].y));
            }
            p = make_float2(tau * ((q[0].x + q[1].x) + (q[2].x + q[3].x)), tau * ((q[0].y + q[1].y) + (q[2].y + q[3].y)));
        }
        if (RS > 1) {
            // the row groups' parts of p_k, added in group order by everybody (every group needs w_k for its rows)
            pp[grp][ct] = p;
            __syncthreads();
            p = pp[0][ct];
#pragma unroll
            for (int g = 1; g < RS; ++g) p.x += pp[g][ct].x, p.y += pp[g][ct].y;
        }
        if (!live) p = make_float2(0.f, 0.f);
        // K = tau/2 v^H p (real for a Hermitian block) ; w = p - K v
        const float kk = block_sum(grp == 0 ? a.x * p.x + a.y * p.y : 0.f, 1);
        const float K = 0.5f * tau * kk;
        const float2 w = make_float2(p.x - K * a.x, p.y - K * a.y);
        if (live && grp == 0) ws[k] = w;
        sync();
        // a_ik -= v_i conj(w_k) + w_i conj(v_k)
        if (live) {
#pragma unroll 4
            for (int i = j + 1 + grp; i < r; i += RS) {
                const float2 vi = vs[i], wi = ws[i];
                float2 x = A[i * r + k];
                x.x = fmaf(-vi.x, w.x, fmaf(-vi.y, w.y, fmaf(-wi.x, a.x, fmaf(-wi.y, a.y, x.x))));
                x.y = fmaf(-vi.y, w.x, fmaf(vi.x, w.y, fmaf(-wi.y, a.x, fmaf(wi.x, a.y, x.y))));
                A[i * r + k] = x;
            }
        }
        sync();
    }
    if (tid == 0) {
        if (r == 1) {
            d[0] = A[0].x;
        } else {
            const int j = r - 2;
            const float2 x00 = A[j * r + j], x01 = A[j * r + j + 1], x11 = A[(j + 1) * r + j + 1];
            d[j] = x00.x;
            d[j + 1] = x11.x;
            const float ea = sqrtf(x01.x * x01.x + x01.y * x01.y);  // sub-diagonal element is conj(x01)
            e[j] = ea;
            if (ea > 0.f) phase = cmul_s(phase, make_float2(x01.x / ea, -x01.y / ea));
            ph[j + 1] = phase;
            taus[j] = 0.f;
        }
        e[r - 1] = 0.f;
        taus[r - 1] = 0.f;
    }
}

}  // namespace

bool vk_tridiag_small_supported(int r) { return r >= 1 && r <= 64; }

int vk_launch_tridiag_small(vk_context* h, cudaStream_t st, float2* W, int B, int r, int ld, size_t wstride, float* d, float* e,
                            float* tau, float2* ph) {
    const size_t smem = ((size_t)r * r + 128) * sizeof(float2);
    // row groups per matrix (r > 33), measured on 8320 / 2000 matrices: r = 64: 1.748 (one) / 1.565 (two) / 1.956 ms (four);
    // r = 57: 0.316 / 0.324 / 0.435 ms; r = 48: 0.236 / 0.236 / 0.338 ms - two from r = 60 on; "tridiag_small_rs" overrides
    const int rs = h->tridiag_small_rs > 0 ? h->tridiag_small_rs : (r >= 60 ? 2 : 1);
    if (r > 33 && rs == 4) {
        VK_CUDA(h, cudaFuncSetAttribute(tridiag_small_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tridiag_small_kernel<2, 4><<<B, 256, smem, st>>>(W, r, ld, wstride, d, e, tau, ph);
    } else if (r > 33 && rs != 1) {
        VK_CUDA(h, cudaFuncSetAttribute(tridiag_small_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tridiag_small_kernel<2, 2><<<B, 128, smem, st>>>(W, r, ld, wstride, d, e, tau, ph);
    } else if (r > 33) {
        VK_CUDA(h, cudaFuncSetAttribute(tridiag_small_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tridiag_small_kernel<2, 1><<<B, 64, smem, st>>>(W, r, ld, wstride, d, e, tau, ph);
    } else {
        VK_CUDA(h, cudaFuncSetAttribute(tridiag_small_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tridiag_small_kernel<1, 1><<<B, 32, smem, st>>>(W, r, ld, wstride, d, e, tau, ph);
    }
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}
