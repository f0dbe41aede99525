// Hermitian eigensolver, direct variant ("eig_impl" = 2): Householder tridiagonalisation -> implicit QL on the real
// tridiagonal with every plane rotation recorded and level-scheduled -> rotations applied to the accumulated
// reflectors. About 10x fewer flops than nine cyclic Jacobi sweeps; float32 throughout (prototype with the measured
// accuracy: tools/proto_tridiag.py). Replaces the LAPACK cgesdd call behind np.linalg.svd (reference compress_ms.py:
// apply_svd) together with the Gram product and the factor formation; parity is checked through the same tests as the
// Jacobi solver.
//
// Conventions: M = W[b] is the r x r row-major Hermitian matrix conj(G) (row i of W = column i of the Gram matrix G).
//   M = Q T Q^H,  Q = H_0 H_1 ... H_{r-3},  H_j = I - tau_j v_j v_j^H   (v_j lives on indices j+1 .. r-1)
//   T = D T_real D^H with unit phases D, T_real = Z Lambda Z^T  =>  eigenvectors of M are the columns of Q D Z.
// Xt holds (Q D Z)^T: row i is eigenvector i, so a plane rotation of columns (i, i+1) of Z acts on rows i, i+1 of Xt.
// On exit W[b][i][:] = lambda_i * conj(Xt[i][:]) - the layout the Jacobi solver leaves (vectors of norm lambda_i).
#include <cmath>
#include <cstdio>

#include "common.cuh"

namespace {

constexpr int TD_THREADS = 512;
constexpr int TD_WARPS = TD_THREADS / 32;
constexpr int FQ_THREADS = 256;
constexpr int FQ_WARPS = FQ_THREADS / 32;
constexpr int FQ_TJ = 8;  // reflectors staged per shared-memory tile
constexpr int RA_C = 8;     // lanes per application slot (two columns of the 16-column slab each)
constexpr int RA_NS_MAX = 128;  // application slots per CTA = sweeps in flight: 64 (r <= 256) or 128
constexpr int QL_MAXIT = 60;

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// ---- 1. tridiagonalisation: one CTA per matrix, matrix in global memory (L2). The two-sided rank-2 update of step
//         j-1 is fused with the matrix-vector product of step j: one read + one write of the trailing block per step.
//         Lane l of a warp owns the absolute columns k = 32 e + l; a warp stages RB whole rows in registers before it
//         touches them, so RB * (r - j) / 32 loads per lane are in flight. Four barriers per step: every thread sums the
//         per-warp partials and repeats the scalar work, v^H p is accumulated inside the pass. Once the trailing block
//         fits next to the vectors it moves to shared memory for good (r <= 384). All shared vectors are indexed by
//         absolute column. Used for r <= 256 and r > 512. ---------------------------------------------------------
template <int EPL, int RB>
__global__ void __launch_bounds__(TD_THREADS, (EPL * RB <= 8) ? 2 : 1)
    tridiag_kernel(float2* __restrict__ Wall, int r, int ld, size_t wstride, float* __restrict__ dall,
                   float* __restrict__ eall, float* __restrict__ tauall, float2* __restrict__ phall, int nts) {
    constexpr int WD = EPL * 32;
    extern __shared__ float2 td_sm[];
    float2* T = td_sm + 5 * WD;  // [ts][ts] trailing block once it fits (ts <= nts): rows/cols j0 .. r-1
    int j0 = -1, ts = 0;         // j0 >= 0: resident
    float2* vprev = td_sm;
    float2* vnew = td_sm + WD;
    float2* wv = td_sm + 2 * WD;
    float2* pv = td_sm + 3 * WD;
    float2* nrow = td_sm + 4 * WD;  // row j+1 as the pass of step j left it (saves a global round trip per step)
    __shared__ float s_part[TD_WARPS];    // per-warp partial of |a|^2
    __shared__ float2 s_kpart[TD_WARPS];  // per-warp partial of v^H p
    __shared__ float2 s_alpha;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* M = Wall + (size_t)b * wstride;
    float* d = dall + (size_t)b * r;
    float* e = eall + (size_t)b * r;
    float* taus = tauall + (size_t)b * r;
    float2* ph = phall + (size_t)b * r;
    float2 phase = make_float2(1.f, 0.f);
    if (tid == 0) ph[0] = phase;
    for (int k = tid; k < 5 * WD; k += TD_THREADS) td_sm[k] = make_float2(0.f, 0.f);
    __syncthreads();

    for (int j = 0; j + 2 < r; ++j) {
        const int e0 = (j + 1) >> 5;
        if (j0 < 0 && r - j <= nts) {
            // from here on the trailing block lives in shared memory: the remaining steps never wait for L2 again
            j0 = j, ts = r - j;
            for (int idx = tid; idx < ts * ts; idx += TD_THREADS) {
                const int i = idx / ts, k = idx - i * ts;
                T[idx] = M[(size_t)(j0 + i) * ld + j0 + k];
            }
            __syncthreads();
        }
        // row j with the pending update of step j-1 applied: diagonal d_j and the column below it (a = conj(row))
        float ss = 0.f;
        {
            const float2 v0 = vprev[j], w0 = wv[j];  // zero while j == 0
            for (int k = tid; k < WD; k += TD_THREADS) {
                float2 a = make_float2(0.f, 0.f);
                if (k >= j && k < r) {
                    float2 x = (j == 0) ? M[k] : nrow[k];
                    const float2 wk = wv[k], vk = vprev[k];
                    x.x -= v0.x * wk.x + v0.y * wk.y + w0.x * vk.x + w0.y * vk.y;
                    x.y -= v0.y * wk.x - v0.x * wk.y + w0.y * vk.x - w0.x * vk.y;
                    if (k == j) {
                        d[j] = x.x;
                    } else {
                        a = make_float2(x.x, -x.y);
                        ss = fmaf(x.x, x.x, fmaf(x.y, x.y, ss));
                        if (k == j + 1) s_alpha = a;
                    }
                }
                vnew[k] = a;
            }
        }
        // four barriers per step: every thread sums the per-warp partials and repeats the scalar work itself
        ss = warp_sum(ss);
        if (lane == 0) s_part[warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < TD_WARPS; ++w) tot += s_part[w];
        float tau = 0.f;
        {
            float ej = 0.f;
            float2 v0 = s_alpha;
            if (tot > 1e-30f) {
                const float xn = sqrtf(tot);
                const float2 alpha = v0;
                const float aa = sqrtf(alpha.x * alpha.x + alpha.y * alpha.y);
                float2 p1 = make_float2(1.f, 0.f);
                if (aa > 0.f) p1 = make_float2(alpha.x / aa, alpha.y / aa);
                v0 = make_float2(alpha.x + p1.x * xn, alpha.y + p1.y * xn);
                tau = 1.f / (xn * (xn + aa));
                ej = xn;
                phase = cmulf(phase, make_float2(-p1.x, -p1.y));  // sub-diagonal element is -p1 * xn
            }
            if (tid == 0) {
                vnew[j + 1] = v0;
                taus[j] = tau;
                e[j] = ej;
                ph[j + 1] = phase;
            }
        }
        __syncthreads();
        // the reflector replaces the (now dead) part of row j right of the diagonal
        {
            float2* row = M + (size_t)j * ld;
            for (int k = j + 1 + tid; k < r; k += TD_THREADS) row[k] = vnew[k];
        }
        // fused pass over the trailing block
        float2 kacc = make_float2(0.f, 0.f);  // this warp's part of v^H p (identical on all its lanes)
        if (j0 >= 0) {
            for (int i = j + 1 + warp; i < r; i += TD_WARPS) {
                float2* row = T + (size_t)(i - j0) * ts - j0;
                const float2 vi = vprev[i], wi = wv[i];
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll 2
                for (int k = j + 1 + lane; k < r; k += 32) {
                    float2 t = row[k];
                    const float2 wk = wv[k], vk = vprev[k];
                    t.x -= vi.x * wk.x + vi.y * wk.y + wi.x * vk.x + wi.y * vk.y;
                    t.y -= vi.y * wk.x - vi.x * wk.y + wi.y * vk.x - wi.x * vk.y;
                    row[k] = t;
                    if (i == j + 1) nrow[k] = t;
                    cfma(acc, t, vnew[k]);
                }
                acc.x = tau * warp_sum(acc.x);
                acc.y = tau * warp_sum(acc.y);
                const float2 v = vnew[i];
                kacc.x += v.x * acc.x + v.y * acc.y;
                kacc.y += v.x * acc.y - v.y * acc.x;
                if (lane == 0) pv[i] = acc;
            }
        } else
        for (int ib = j + 1 + warp; ib < r; ib += TD_WARPS * RB) {
            float2 x[RB][EPL];
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                const int i = ib + q * TD_WARPS;
                const float2* row = M + (size_t)i * ld;
#pragma unroll
                for (int ee = 0; ee < EPL; ++ee) {
                    const int k = ee * 32 + lane;
                    x[q][ee] = (ee >= e0 && k > j && k < r && i < r) ? row[k] : make_float2(0.f, 0.f);
                }
            }
            float2 vi[RB], wi[RB], acc[RB];
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                const int i = ib + q * TD_WARPS;
                vi[q] = i < r ? vprev[i] : make_float2(0.f, 0.f);
                wi[q] = i < r ? wv[i] : make_float2(0.f, 0.f);
                acc[q] = make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int ee = 0; ee < EPL; ++ee) {
                const int k = ee * 32 + lane;
                if (ee >= e0 && k > j && k < r) {
                    const float2 wk = wv[k], vk = vprev[k], vn = vnew[k];
#pragma unroll
                    for (int q = 0; q < RB; ++q) {
                        float2 t = x[q][ee];
                        t.x -= vi[q].x * wk.x + vi[q].y * wk.y + wi[q].x * vk.x + wi[q].y * vk.y;
                        t.y -= vi[q].y * wk.x - vi[q].x * wk.y + wi[q].y * vk.x - wi[q].x * vk.y;
                        x[q][ee] = t;
                        cfma(acc[q], t, vn);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                const int i = ib + q * TD_WARPS;
                if (i < r) {
                    float2* row = M + (size_t)i * ld;
                    if (j > 0) {
#pragma unroll
                        for (int ee = 0; ee < EPL; ++ee) {
                            const int k = ee * 32 + lane;
                            if (ee >= e0 && k > j && k < r) row[k] = x[q][ee];
                        }
                    }
                    if (i == j + 1) {
#pragma unroll
                        for (int ee = 0; ee < EPL; ++ee) {
                            const int k = ee * 32 + lane;
                            if (ee >= e0 && k > j && k < r) nrow[k] = x[q][ee];
                        }
                    }
                    const float ax = tau * warp_sum(acc[q].x), ay = tau * warp_sum(acc[q].y);
                    const float2 v = vnew[i];
                    kacc.x += v.x * ax + v.y * ay;
                    kacc.y += v.x * ay - v.y * ax;
                    if (lane == 0) pv[i] = make_float2(ax, ay);
                }
            }
        }
        if (lane == 0) s_kpart[warp] = kacc;
        __syncthreads();
        // K = tau/2 * v^H p ;  w = p - K v
        float2 kk = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < TD_WARPS; ++w) kk.x += s_kpart[w].x, kk.y += s_kpart[w].y;
        const float2 K = make_float2(0.5f * tau * kk.x, 0.5f * tau * kk.y);
        for (int k = j + 1 + tid; k < r; k += TD_THREADS) {
            const float2 v = vnew[k], p = pv[k];
            wv[k] = make_float2(p.x - (K.x * v.x - K.y * v.y), p.y - (K.x * v.y + K.y * v.x));
        }
        float2* t = vprev;
        vprev = vnew;
        vnew = t;
        __syncthreads();
    }
    if (tid == 0) {
        if (r == 1) {
            d[0] = M[0].x;
        } else {
            const int j = r - 2;
            float2 x00, x01, x11;
            if (j0 >= 0) {
                x00 = T[(size_t)(j - j0) * ts + j - j0], x01 = T[(size_t)(j - j0) * ts + j + 1 - j0];
                x11 = T[(size_t)(j + 1 - j0) * ts + j + 1 - j0];
            } else {
                x00 = M[(size_t)j * ld + j], x01 = M[(size_t)j * ld + j + 1], x11 = M[(size_t)(j + 1) * ld + j + 1];
            }
            if (r >= 3) {
                const float2 v0 = vprev[j], v1 = vprev[j + 1], w0 = wv[j], w1 = wv[j + 1];
                x00.x -= 2.f * (v0.x * w0.x + v0.y * w0.y);
                x11.x -= 2.f * (v1.x * w1.x + v1.y * w1.y);
                x01.x -= v0.x * w1.x + v0.y * w1.y + w0.x * v1.x + w0.y * v1.y;
                x01.y -= v0.y * w1.x - v0.x * w1.y + w0.y * v1.x - w0.x * v1.y;
            }
            d[j] = x00.x;
            d[j + 1] = x11.x;
            const float ea = sqrtf(x01.x * x01.x + x01.y * x01.y);  // sub-diagonal element is conj(x01)
            e[j] = ea;
            if (ea > 0.f) phase = cmulf(phase, make_float2(x01.x / ea, -x01.y / ea));
            ph[j + 1] = phase;
            taus[j] = 0.f;
        }
        e[r - 1] = 0.f;
        taus[r - 1] = 0.f;
    }
}

// ---- 1b. the same reduction touching only the lower triangle (used for 256 < r <= 512): a pass reads and writes the elements
//          (i, k), j < k <= i, once, which halves the traffic and two thirds of the arithmetic. Element x = A(i, k)
//          contributes x v_k to (A v)_i (row part, reduced over the lanes of the warp that owns row i) and, for k < i,
//          conj(x) v_i to (A v)_k (column part, accumulated per lane - a lane owns the columns k = 32 e + lane - and
//          summed over the warps through shared memory). v^H A v = 2 Re sum_i conj(v_i) rowpart_i - sum_i |v_i|^2 A(i, i)
//          needs the row parts only. The column below the next pivot and the next diagonal element are collected by the
//          pass; the reflectors go to the (otherwise unused) upper triangle where formq / backtr expect them.
//          Once the trailing block fits it is expanded to a full square in shared memory (resident steps). -------------
template <int EPL, int RB>
__global__ void __launch_bounds__(TD_THREADS, 1)
    tridiag_sym_kernel(float2* __restrict__ Wall, int r, int ld, size_t wstride, float* __restrict__ dall,
                       float* __restrict__ eall, float* __restrict__ tauall, float2* __restrict__ phall, int nts) {
    constexpr int WD = EPL * 32;
    extern __shared__ float2 td_sm[];
    float2* vprev = td_sm;
    float2* vnew = td_sm + WD;
    float2* wv = td_sm + 2 * WD;
    float2* pv = td_sm + 3 * WD;      // row parts of A v (unscaled)
    float2* acol = td_sm + 4 * WD;    // column j+1 below the diagonal as the pass of step j left it
    float2* pc = td_sm + 5 * WD;      // [TD_WARPS][WD] per-warp column parts of A v
    float2* T = pc + TD_WARPS * WD;   // [ts][ts] resident trailing block (full square), rows/cols j0 .. r-1
    int j0 = -1, ts = 0;
    __shared__ float s_part[TD_WARPS];
    __shared__ float4 s_kpart[TD_WARPS];  // (Re, Im of sum conj(v_i) rowpart_i, sum |v_i|^2 A_ii, -)
    __shared__ float2 s_alpha;
    __shared__ float s_diag;              // A(j+1, j+1) as the pass of step j left it
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* M = Wall + (size_t)b * wstride;
    float* d = dall + (size_t)b * r;
    float* e = eall + (size_t)b * r;
    float* taus = tauall + (size_t)b * r;
    float2* ph = phall + (size_t)b * r;
    float2 phase = make_float2(1.f, 0.f);
    if (tid == 0) ph[0] = phase;
    for (int k = tid; k < (5 + TD_WARPS) * WD; k += TD_THREADS) td_sm[k] = make_float2(0.f, 0.f);
    __syncthreads();

    for (int j = 0; j + 2 < r; ++j) {
        const int e0 = (j + 1) >> 5;
        if (j0 < 0 && r - j <= nts) {
            j0 = j, ts = r - j;
            for (int idx = tid; idx < ts * ts; idx += TD_THREADS) {
                const int i = idx / ts, k = idx - i * ts;
                float2 v;
                if (k <= i) {
                    v = M[(size_t)(j0 + i) * ld + j0 + k];
                } else {
                    v = M[(size_t)(j0 + k) * ld + j0 + i];
                    v.y = -v.y;
                }
                T[idx] = v;
            }
            __syncthreads();
        }
        const bool resident = j0 >= 0;
        // column j below the diagonal and the diagonal element, with the pending update of step j-1 applied
        float ss = 0.f;
        {
            const float2 vj = vprev[j], wj = wv[j];  // zero while j == 0
            for (int k = tid; k < WD; k += TD_THREADS) {
                float2 a = make_float2(0.f, 0.f);
                if (k > j && k < r) {
                    float2 x = (j == 0) ? M[(size_t)k * ld] : acol[k];
                    const float2 vk = vprev[k], wk = wv[k];
                    x.x -= vk.x * wj.x + vk.y * wj.y + wk.x * vj.x + wk.y * vj.y;
                    x.y -= vk.y * wj.x - vk.x * wj.y + wk.y * vj.x - wk.x * vj.y;
                    a = x;
                    ss = fmaf(x.x, x.x, fmaf(x.y, x.y, ss));
                    if (k == j + 1) s_alpha = a;
                } else if (k == j) {
                    float xd = (j == 0) ? M[0].x : s_diag;
                    xd -= 2.f * (vj.x * wj.x + vj.y * wj.y);
                    d[j] = xd;
                }
                vnew[k] = a;
            }
        }
        ss = warp_sum(ss);
        if (lane == 0) s_part[warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < TD_WARPS; ++w) tot += s_part[w];
        float tau = 0.f;
        {
            float ej = 0.f;
            float2 v0 = s_alpha;
            if (tot > 1e-30f) {
                const float xn = sqrtf(tot);
                const float2 alpha = v0;
                const float aa = sqrtf(alpha.x * alpha.x + alpha.y * alpha.y);
                float2 p1 = make_float2(1.f, 0.f);
                if (aa > 0.f) p1 = make_float2(alpha.x / aa, alpha.y / aa);
                v0 = make_float2(alpha.x + p1.x * xn, alpha.y + p1.y * xn);
                tau = 1.f / (xn * (xn + aa));
                ej = xn;
                phase = cmulf(phase, make_float2(-p1.x, -p1.y));  // sub-diagonal element is -p1 * xn
            }
            if (tid == 0) {
                vnew[j + 1] = v0;
                taus[j] = tau;
                e[j] = ej;
                ph[j + 1] = phase;
            }
        }
        __syncthreads();
        // the reflector goes to row j right of the diagonal (upper triangle: not part of the working storage)
        {
            float2* row = M + (size_t)j * ld;
            for (int k = j + 1 + tid; k < r; k += TD_THREADS) row[k] = vnew[k];
        }
        float2 kacc = make_float2(0.f, 0.f);  // sum conj(v_i) rowpart_i over this warp's rows (same on all lanes)
        float kd = 0.f;                       // sum |v_i|^2 A(i, i) (per lane, reduced below)
        if (resident) {
            for (int i = j + 1 + warp; i < r; i += TD_WARPS) {
                float2* row = T + (size_t)(i - j0) * ts - j0;
                const float2 vi = vprev[i], wi = wv[i];
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll 2
                for (int k = j + 1 + lane; k < r; k += 32) {
                    float2 t = row[k];
                    const float2 wk = wv[k], vk = vprev[k];
                    t.x -= vi.x * wk.x + vi.y * wk.y + wi.x * vk.x + wi.y * vk.y;
                    t.y -= vi.y * wk.x - vi.x * wk.y + wi.y * vk.x - wi.x * vk.y;
                    row[k] = t;
                    if (i == j + 1) {
                        if (k == j + 1) s_diag = t.x;
                        else acol[k] = make_float2(t.x, -t.y);  // column j+1 = conj(row j+1)
                    }
                    cfma(acc, t, vnew[k]);
                }
                acc.x = warp_sum(acc.x);
                acc.y = warp_sum(acc.y);
                const float2 v = vnew[i];
                kacc.x += v.x * acc.x + v.y * acc.y;
                kacc.y += v.x * acc.y - v.y * acc.x;
                if (lane == 0) pv[i] = acc;
            }
        } else {
            float2 pcr[EPL];
#pragma unroll
            for (int ee = 0; ee < EPL; ++ee) pcr[ee] = make_float2(0.f, 0.f);
            for (int ib = j + 1 + warp; ib < r; ib += TD_WARPS * RB) {
                float2 x[RB][EPL];
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const int i = ib + q * TD_WARPS;
                    const float2* row = M + (size_t)i * ld;
#pragma unroll
                    for (int ee = 0; ee < EPL; ++ee) {
                        const int k = ee * 32 + lane;
                        x[q][ee] = (ee >= e0 && k > j && k <= i && i < r) ? row[k] : make_float2(0.f, 0.f);
                    }
                }
                float2 vi[RB], wi[RB], vni[RB], acc[RB];
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const int i = ib + q * TD_WARPS;
                    const bool ok = i < r;
                    vi[q] = ok ? vprev[i] : make_float2(0.f, 0.f);
                    wi[q] = ok ? wv[i] : make_float2(0.f, 0.f);
                    vni[q] = ok ? vnew[i] : make_float2(0.f, 0.f);
                    acc[q] = make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int ee = 0; ee < EPL; ++ee) {
                    const int k = ee * 32 + lane;
                    if (ee >= e0 && k > j && k < r) {
                        const float2 wk = wv[k], vk = vprev[k], vn = vnew[k];
#pragma unroll
                        for (int q = 0; q < RB; ++q) {
                            const int i = ib + q * TD_WARPS;
                            if (k <= i && i < r) {
                                float2 t = x[q][ee];
                                t.x -= vi[q].x * wk.x + vi[q].y * wk.y + wi[q].x * vk.x + wi[q].y * vk.y;
                                t.y -= vi[q].y * wk.x - vi[q].x * wk.y + wi[q].y * vk.x - wi[q].x * vk.y;
                                x[q][ee] = t;
                                cfma(acc[q], t, vn);
                                if (k < i) {
                                    // column part: conj(t) * v_i
                                    pcr[ee].x = fmaf(t.x, vni[q].x, fmaf(t.y, vni[q].y, pcr[ee].x));
                                    pcr[ee].y = fmaf(t.x, vni[q].y, fmaf(-t.y, vni[q].x, pcr[ee].y));
                                    if (k == j + 1) acol[i] = t;
                                } else {
                                    kd = fmaf(vni[q].x * vni[q].x + vni[q].y * vni[q].y, t.x, kd);
                                    if (i == j + 1) s_diag = t.x;
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const int i = ib + q * TD_WARPS;
                    if (i < r) {
                        if (j > 0) {
                            float2* row = M + (size_t)i * ld;
#pragma unroll
                            for (int ee = 0; ee < EPL; ++ee) {
                                const int k = ee * 32 + lane;
                                if (ee >= e0 && k > j && k <= i) row[k] = x[q][ee];
                            }
                        }
                        const float ax = warp_sum(acc[q].x), ay = warp_sum(acc[q].y);
                        kacc.x += vni[q].x * ax + vni[q].y * ay;
                        kacc.y += vni[q].x * ay - vni[q].y * ax;
                        if (lane == 0) pv[i] = make_float2(ax, ay);
                    }
                }
            }
#pragma unroll
            for (int ee = 0; ee < EPL; ++ee) pc[warp * WD + ee * 32 + lane] = pcr[ee];
            kd = warp_sum(kd);
        }
        if (lane == 0) s_kpart[warp] = make_float4(kacc.x, kacc.y, kd, 0.f);
        __syncthreads();
        // v^H A v (real), K = tau^2/2 * v^H A v, p = tau * (row part + column parts), w = p - K v
        float kre = 0.f, kdd = 0.f;
#pragma unroll
        for (int w = 0; w < TD_WARPS; ++w) kre += s_kpart[w].x, kdd += s_kpart[w].z;
        const float vav = resident ? kre : 2.f * kre - kdd;
        const float K = 0.5f * tau * tau * vav;
        for (int k = j + 1 + tid; k < r; k += TD_THREADS) {
            float2 p = pv[k];
            if (!resident) {
#pragma unroll
                for (int w = 0; w < TD_WARPS; ++w) {
                    const float2 c = pc[w * WD + k];
                    p.x += c.x, p.y += c.y;
                }
            }
            const float2 v = vnew[k];
            wv[k] = make_float2(tau * p.x - K * v.x, tau * p.y - K * v.y);
        }
        float2* t = vprev;
        vprev = vnew;
        vnew = t;
        __syncthreads();
    }
    if (tid == 0) {
        if (r == 1) {
            d[0] = M[0].x;
        } else {
            const int j = r - 2;
            // (j, j), (j+1, j) and (j+1, j+1) with the pending update of step j-1 = r-3 applied
            float x00, x11;
            float2 x10;
            if (r >= 3) {
                x00 = s_diag;
                x10 = acol[j + 1];
                x11 = j0 >= 0 ? T[(size_t)(j + 1 - j0) * ts + j + 1 - j0].x : M[(size_t)(j + 1) * ld + j + 1].x;
                const float2 v0 = vprev[j], v1 = vprev[j + 1], w0 = wv[j], w1 = wv[j + 1];
                x00 -= 2.f * (v0.x * w0.x + v0.y * w0.y);
                x11 -= 2.f * (v1.x * w1.x + v1.y * w1.y);
                x10.x -= v1.x * w0.x + v1.y * w0.y + w1.x * v0.x + w1.y * v0.y;
                x10.y -= v1.y * w0.x - v1.x * w0.y + w1.y * v0.x - w1.x * v0.y;
            } else {
                x00 = M[0].x, x10 = M[(size_t)ld], x11 = M[(size_t)ld + 1].x;
            }
            d[j] = x00;
            d[j + 1] = x11;
            const float ea = sqrtf(x10.x * x10.x + x10.y * x10.y);  // sub-diagonal element
            e[j] = ea;
            if (ea > 0.f) phase = cmulf(phase, make_float2(x10.x / ea, x10.y / ea));
            ph[j + 1] = phase;
            taus[j] = 0.f;
        }
        e[r - 1] = 0.f;
        taus[r - 1] = 0.f;
    }
}


// ---- 1d. full-storage reduction with DEFERRED updates while the trailing block lives in global memory (L2 / HBM), then
//          the shared-memory resident steps of kernel 1. The streaming phase of kernel 1 reads AND writes the trailing
//          block every step and is bound by the L2 throughput an SM gets (r = 256) or by HBM (r = 512); here up to NB
//          reflector pairs (v_s, w_s) stay pending in shared memory (LAPACK latrd): a step only READS the block as the
//          last update pass left it (A0) and corrects the product,
//              A v = A0 v - sum_s [ v_s (w_s^H v) + w_s (v_s^H v) ],
//          and the row that defines the next reflector is corrected the same way; when NB pairs are pending the pass of
//          the next step applies them all (rank-2NB update, read + write) before its product. Whole rows, whole columns:
//          no triangle bookkeeping, the reflector's entries for a lane's columns stay in registers for the pass, and RB
//          rows share every shared-memory operand of the update. Bytes per step: (1 + 2/NB)/2 of kernel 1. -------------
template <int EPL, int RB, int NB>
__global__ void __launch_bounds__(TD_THREADS, 1)
    tridiag_defer_kernel(float2* __restrict__ Wall, int r, int ld, size_t wstride, float* __restrict__ dall,
                         float* __restrict__ eall, float* __restrict__ tauall, float2* __restrict__ phall, int nts) {
    constexpr int WD = EPL * 32;
    static_assert(NB <= TD_WARPS, "one warp per pending pair computes its two inner products");
    extern __shared__ float2 td_sm[];
    float2* vprev = td_sm;            // resident phase only: pending single pair (vprev, wv)
    float2* vnew = td_sm + WD;
    float2* wv = td_sm + 2 * WD;
    float2* pv = td_sm + 3 * WD;
    float2* nrow = td_sm + 4 * WD;    // row j+1 as memory (or T) holds it after the pass of step j
    float2* Vp = td_sm + 5 * WD;      // [NB][WD] pending reflectors (streaming phase)
    float2* Wp = Vp + NB * WD;        // [NB][WD] pending w vectors
    float2* T = Wp + NB * WD;         // [ts][ts] trailing block once it fits (ts <= nts): rows/cols j0 .. r-1
    int j0 = -1, ts = 0;
    __shared__ float s_part[TD_WARPS];
    __shared__ float2 s_kpart[TD_WARPS];
    __shared__ float2 s_alpha;
    __shared__ float2 s_ab[2 * NB];   // alpha_s = w_s^H v, beta_s = v_s^H v
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* M = Wall + (size_t)b * wstride;
    float* d = dall + (size_t)b * r;
    float* e = eall + (size_t)b * r;
    float* taus = tauall + (size_t)b * r;
    float2* ph = phall + (size_t)b * r;
    float2 phase = make_float2(1.f, 0.f);
    if (tid == 0) ph[0] = phase;
    for (int k = tid; k < (5 + 2 * NB) * WD; k += TD_THREADS) td_sm[k] = make_float2(0.f, 0.f);
    __syncthreads();
    int P = 0;  // pending pairs of the streaming phase (uniform)

    for (int j = 0; j + 2 < r; ++j) {
        const int e0 = (j + 1) >> 5;
        if (j0 < 0 && r - j <= nts) {
            // from here on the trailing block lives in shared memory, with every pending pair applied on the way in
            j0 = j, ts = r - j;
            for (int idx = tid; idx < ts * ts; idx += TD_THREADS) {
                const int i = j0 + idx / ts, k = j0 + idx % ts;
                float2 x = M[(size_t)i * ld + k];
                for (int s2 = 0; s2 < P; ++s2) {
                    const float2 vi = Vp[s2 * WD + i], wi = Wp[s2 * WD + i], vk = Vp[s2 * WD + k], wk = Wp[s2 * WD + k];
                    x.x -= vi.x * wk.x + vi.y * wk.y + wi.x * vk.x + wi.y * vk.y;
                    x.y -= vi.y * wk.x - vi.x * wk.y + wi.y * vk.x - wi.x * vk.y;
                }
                T[idx] = x;
            }
            __syncthreads();
            for (int k = j + tid; k < r; k += TD_THREADS) nrow[k] = T[k - j0];   // row j of the updated block
            for (int k = tid; k < WD; k += TD_THREADS) vprev[k] = wv[k] = make_float2(0.f, 0.f);
            P = 0;
            __syncthreads();
        }
        const bool resident = j0 >= 0;
        // row j with everything pending applied: diagonal d_j and the column below it (a = conj(row))
        float ss = 0.f;
        {
            const float2 v0 = vprev[j], w0 = wv[j];  // zero outside the resident phase
            for (int k = tid; k < WD; k += TD_THREADS) {
                float2 a = make_float2(0.f, 0.f);
                if (k >= j && k < r) {
                    float2 x = (j == 0) ? M[k] : nrow[k];
                    if (resident) {
                        const float2 wk = wv[k], vk = vprev[k];
                        x.x -= v0.x * wk.x + v0.y * wk.y + w0.x * vk.x + w0.y * vk.y;
                        x.y -= v0.y * wk.x - v0.x * wk.y + w0.y * vk.x - w0.x * vk.y;
                    } else {
                        for (int s2 = 0; s2 < P; ++s2) {
                            const float2 vj = Vp[s2 * WD + j], wj = Wp[s2 * WD + j], vk = Vp[s2 * WD + k], wk = Wp[s2 * WD + k];
                            x.x -= vj.x * wk.x + vj.y * wk.y + wj.x * vk.x + wj.y * vk.y;
                            x.y -= vj.y * wk.x - vj.x * wk.y + wj.y * vk.x - wj.x * vk.y;
                        }
                    }
                    if (k == j) {
                        d[j] = x.x;
                    } else {
                        a = make_float2(x.x, -x.y);
                        ss = fmaf(x.x, x.x, fmaf(x.y, x.y, ss));
                        if (k == j + 1) s_alpha = a;
                    }
                }
                vnew[k] = a;
            }
        }
        ss = warp_sum(ss);
        if (lane == 0) s_part[warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < TD_WARPS; ++w) tot += s_part[w];
        float tau = 0.f;
        {
            float ej = 0.f;
            float2 v0 = s_alpha;
            if (tot > 1e-30f) {
                const float xn = sqrtf(tot);
                const float2 alpha = v0;
                const float aa = sqrtf(alpha.x * alpha.x + alpha.y * alpha.y);
                float2 p1 = make_float2(1.f, 0.f);
                if (aa > 0.f) p1 = make_float2(alpha.x / aa, alpha.y / aa);
                v0 = make_float2(alpha.x + p1.x * xn, alpha.y + p1.y * xn);
                tau = 1.f / (xn * (xn + aa));
                ej = xn;
                phase = cmulf(phase, make_float2(-p1.x, -p1.y));  // sub-diagonal element is -p1 * xn
            }
            if (tid == 0) {
                vnew[j + 1] = v0;
                taus[j] = tau;
                e[j] = ej;
                ph[j + 1] = phase;
            }
        }
        __syncthreads();
        // the reflector replaces the (now dead) part of row j right of the diagonal
        {
            float2* row = M + (size_t)j * ld;
            for (int k = j + 1 + tid; k < r; k += TD_THREADS) row[k] = vnew[k];
        }
        float2 kacc = make_float2(0.f, 0.f);  // this warp's part of v^H (A0 v) resp. v^H p (identical on all its lanes)
        const bool upd = !resident && (P == NB);
        if (resident) {
            for (int i = j + 1 + warp; i < r; i += TD_WARPS) {
                float2* row = T + (size_t)(i - j0) * ts - j0;
                const float2 vi = vprev[i], wi = wv[i];
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll 2
                for (int k = j + 1 + lane; k < r; k += 32) {
                    float2 t = row[k];
                    const float2 wk = wv[k], vk = vprev[k];
                    t.x -= vi.x * wk.x + vi.y * wk.y + wi.x * vk.x + wi.y * vk.y;
                    t.y -= vi.y * wk.x - vi.x * wk.y + wi.y * vk.x - wi.x * vk.y;
                    row[k] = t;
                    if (i == j + 1) nrow[k] = t;
                    cfma(acc, t, vnew[k]);
                }
                acc.x = tau * warp_sum(acc.x);
                acc.y = tau * warp_sum(acc.y);
                const float2 v = vnew[i];
                kacc.x += v.x * acc.x + v.y * acc.y;
                kacc.y += v.x * acc.y - v.y * acc.x;
                if (lane == 0) pv[i] = acc;
            }
        } else {
            if (!upd && warp < P) {
                // alpha_s = w_s^H v, beta_s = v_s^H v for the pending pair s = warp
                float2 aa = make_float2(0.f, 0.f), bb = make_float2(0.f, 0.f);
                for (int k = j + 1 + lane; k < r; k += 32) {
                    const float2 v = vnew[k], wk = Wp[warp * WD + k], vk = Vp[warp * WD + k];
                    aa.x = fmaf(wk.x, v.x, fmaf(wk.y, v.y, aa.x));
                    aa.y = fmaf(wk.x, v.y, fmaf(-wk.y, v.x, aa.y));
                    bb.x = fmaf(vk.x, v.x, fmaf(vk.y, v.y, bb.x));
                    bb.y = fmaf(vk.x, v.y, fmaf(-vk.y, v.x, bb.y));
                }
                aa.x = warp_sum(aa.x), aa.y = warp_sum(aa.y), bb.x = warp_sum(bb.x), bb.y = warp_sum(bb.y);
                if (lane == 0) s_ab[warp] = aa, s_ab[NB + warp] = bb;
            }
            // the reflector's entries for this lane's columns stay in registers for the whole pass
            float2 vn[EPL];
#pragma unroll
            for (int ee = 0; ee < EPL; ++ee) {
                const int k = ee * 32 + lane;
                vn[ee] = (k > j && k < r) ? vnew[k] : make_float2(0.f, 0.f);
            }
            for (int ib = j + 1 + warp; ib < r; ib += TD_WARPS * RB) {
                float2 x[RB][EPL];
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const int i = ib + q * TD_WARPS;
                    const float2* row = M + (size_t)i * ld;
#pragma unroll
                    for (int ee = 0; ee < EPL; ++ee) {
                        const int k = ee * 32 + lane;
                        x[q][ee] = (ee >= e0 && k > j && k < r && i < r) ? row[k] : make_float2(0.f, 0.f);
                    }
                }
                if (upd) {
#pragma unroll 1
                    for (int s2 = 0; s2 < NB; ++s2) {
                        float2 vi[RB], wi[RB];
#pragma unroll
                        for (int q = 0; q < RB; ++q) {
                            const int i = ib + q * TD_WARPS;
                            vi[q] = i < r ? Vp[s2 * WD + i] : make_float2(0.f, 0.f);
                            wi[q] = i < r ? Wp[s2 * WD + i] : make_float2(0.f, 0.f);
                        }
#pragma unroll
                        for (int ee = 0; ee < EPL; ++ee) {
                            if (ee >= e0) {
                                const int k = ee * 32 + lane;
                                const float2 wk = Wp[s2 * WD + k], vk = Vp[s2 * WD + k];
#pragma unroll
                                for (int q = 0; q < RB; ++q) {
                                    x[q][ee].x -= vi[q].x * wk.x + vi[q].y * wk.y + wi[q].x * vk.x + wi[q].y * vk.y;
                                    x[q][ee].y -= vi[q].y * wk.x - vi[q].x * wk.y + wi[q].y * vk.x - wi[q].x * vk.y;
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const int i = ib + q * TD_WARPS;
                    if (i < r) {
                        float2* row = M + (size_t)i * ld;
                        if (upd) {
#pragma unroll
                            for (int ee = 0; ee < EPL; ++ee) {
                                const int k = ee * 32 + lane;
                                if (ee >= e0 && k > j && k < r) row[k] = x[q][ee];
                            }
                        }
                        if (i == j + 1) {
#pragma unroll
                            for (int ee = 0; ee < EPL; ++ee) {
                                const int k = ee * 32 + lane;
                                if (ee >= e0 && k > j && k < r) nrow[k] = x[q][ee];
                            }
                        }
                        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                        for (int ee = 0; ee < EPL; ++ee)
                            if (ee >= e0) cfma(acc, x[q][ee], vn[ee]);   // vn is zero at and left of column j
                        const float ax = warp_sum(acc.x), ay = warp_sum(acc.y);
                        const float2 v = vnew[i];
                        kacc.x += v.x * ax + v.y * ay;
                        kacc.y += v.x * ay - v.y * ax;
                        if (lane == 0) pv[i] = make_float2(ax, ay);   // unscaled, uncorrected
                    }
                }
            }
        }
        if (lane == 0) s_kpart[warp] = kacc;
        __syncthreads();
        float2 kk = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < TD_WARPS; ++w) kk.x += s_kpart[w].x, kk.y += s_kpart[w].y;
        if (resident) {
            // K = tau/2 * v^H p ;  w = p - K v
            const float2 K = make_float2(0.5f * tau * kk.x, 0.5f * tau * kk.y);
            for (int k = j + 1 + tid; k < r; k += TD_THREADS) {
                const float2 v = vnew[k], p = pv[k];
                wv[k] = make_float2(p.x - (K.x * v.x - K.y * v.y), p.y - (K.x * v.y + K.y * v.x));
            }
            float2* t = vprev;
            vprev = vnew;
            vnew = t;
        } else {
            // p = tau (A0 v - sum_s [v_s alpha_s + w_s beta_s]) ; v^H p = tau (v^H A0 v - 2 Re sum_s conj(beta_s) alpha_s)
            const int np = upd ? 0 : P;
            for (int s2 = 0; s2 < np; ++s2) {
                const float2 al = s_ab[s2], be = s_ab[NB + s2];
                kk.x -= 2.f * (be.x * al.x + be.y * al.y);
            }
            const float2 K = make_float2(0.5f * tau * tau * kk.x, 0.5f * tau * tau * kk.y);
            const int slot = upd ? 0 : P;
            for (int k = j + 1 + tid; k < r; k += TD_THREADS) {
                float2 p = pv[k];
                for (int s2 = 0; s2 < np; ++s2) {
                    const float2 al = s_ab[s2], be = s_ab[NB + s2];
                    const float2 vk = Vp[s2 * WD + k], wk = Wp[s2 * WD + k];
                    p.x -= vk.x * al.x - vk.y * al.y + wk.x * be.x - wk.y * be.y;
                    p.y -= vk.x * al.y + vk.y * al.x + wk.x * be.y + wk.y * be.x;
                }
                const float2 v = vnew[k];
                Wp[slot * WD + k] = make_float2(tau * p.x - (K.x * v.x - K.y * v.y), tau * p.y - (K.x * v.y + K.y * v.x));
                Vp[slot * WD + k] = v;
            }
            P = upd ? 1 : P + 1;
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (r == 1) {
            d[0] = M[0].x;
        } else {
            const int j = r - 2;
            float2 x00, x01, x11;
            if (j0 >= 0) {
                x00 = T[(size_t)(j - j0) * ts + j - j0], x01 = T[(size_t)(j - j0) * ts + j + 1 - j0];
                x11 = T[(size_t)(j + 1 - j0) * ts + j + 1 - j0];
            } else {
                x00 = M[(size_t)j * ld + j], x01 = M[(size_t)j * ld + j + 1], x11 = M[(size_t)(j + 1) * ld + j + 1];
            }
            if (r >= 3) {
                if (j0 >= 0) {
                    const float2 v0 = vprev[j], v1 = vprev[j + 1], w0 = wv[j], w1 = wv[j + 1];
                    x00.x -= 2.f * (v0.x * w0.x + v0.y * w0.y);
                    x11.x -= 2.f * (v1.x * w1.x + v1.y * w1.y);
                    x01.x -= v0.x * w1.x + v0.y * w1.y + w0.x * v1.x + w0.y * v1.y;
                    x01.y -= v0.y * w1.x - v0.x * w1.y + w0.y * v1.x - w0.x * v1.y;
                } else {
                    for (int s2 = 0; s2 < P; ++s2) {
                        const float2 v0 = Vp[s2 * WD + j], v1 = Vp[s2 * WD + j + 1], w0 = Wp[s2 * WD + j], w1 = Wp[s2 * WD + j + 1];
                        x00.x -= 2.f * (v0.x * w0.x + v0.y * w0.y);
                        x11.x -= 2.f * (v1.x * w1.x + v1.y * w1.y);
                        x01.x -= v0.x * w1.x + v0.y * w1.y + w0.x * v1.x + w0.y * v1.y;
                        x01.y -= v0.y * w1.x - v0.x * w1.y + w0.y * v1.x - w0.x * v1.y;
                    }
                }
            }
            d[j] = x00.x;
            d[j + 1] = x11.x;
            const float ea = sqrtf(x01.x * x01.x + x01.y * x01.y);  // sub-diagonal element is conj(x01)
            e[j] = ea;
            if (ea > 0.f) phase = cmulf(phase, make_float2(x01.x / ea, -x01.y / ea));
            ph[j + 1] = phase;
            taus[j] = 0.f;
        }
        e[r - 1] = 0.f;
        taus[r - 1] = 0.f;
    }
}

// One reflector applied to the RPW rows a warp holds, y <- (I - tau v v^H) y, with the first live column chunk E0 known at
// compile time: a run-time `if (e >= e0)` inside the unrolled loops only predicates the instructions (ncu / SASS: 258 of
// 320 FFMAs predicated, all of them issued), so the shrinking support of the reflectors saved nothing.
template <int EPL, int RPW, int E0>
__device__ __forceinline__ void fq_apply(float2 (&y)[RPW][EPL], const float2* __restrict__ sv, float tau, int lane, int qmin) {
    float2 v[EPL];
#pragma unroll
    for (int e = E0; e < EPL; ++e) v[e] = sv[e * 32 + lane];
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
        if (q < qmin) continue;
        float2 u = make_float2(0.f, 0.f);
#pragma unroll
        for (int e = E0; e < EPL; ++e) {
            u.x = fmaf(v[e].x, y[q][e].x, fmaf(v[e].y, y[q][e].y, u.x));
            u.y = fmaf(v[e].x, y[q][e].y, fmaf(-v[e].y, y[q][e].x, u.y));
        }
        u.x = tau * warp_sum(u.x);
        u.y = tau * warp_sum(u.y);
#pragma unroll
        for (int e = E0; e < EPL; ++e) {
            y[q][e].x = fmaf(-u.x, v[e].x, fmaf(u.y, v[e].y, y[q][e].x));
            y[q][e].y = fmaf(-u.x, v[e].y, fmaf(-u.y, v[e].x, y[q][e].y));
        }
    }
}
template <int EPL, int RPW, int E0 = 0>
__device__ __forceinline__ void fq_dispatch(int e0, float2 (&y)[RPW][EPL], const float2* __restrict__ sv, float tau, int lane,
                                            int qmin) {
    if constexpr (E0 < EPL) {
        if (e0 == E0) fq_apply<EPL, RPW, E0>(y, sv, tau, lane, qmin);
        else fq_dispatch<EPL, RPW, E0 + 1>(e0, y, sv, tau, lane, qmin);
    }
}

// ---- 2. Xt0 = (Q D)^T: row i = H_0 ... H_{r-3} e_i, kept in the registers of one warp for all reflectors -------------
template <int EPL, int RPW>
__global__ void __launch_bounds__(FQ_THREADS, (EPL <= 16) ? 2 : 1)
    formq_kernel(const float2* __restrict__ Wall, int r, int ld, size_t wstride, const float* __restrict__ tauall,
                 const float2* __restrict__ phall, float2* __restrict__ Xall, const int32_t* __restrict__ skip) {
    constexpr int WIDTH = EPL * 32;
    extern __shared__ float2 fq_sv[];  // [FQ_TJ][WIDTH]
    __shared__ float stau[FQ_TJ];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (skip && skip[b]) return;
    const float2* M = Wall + (size_t)b * wstride;
    const float* taus = tauall + (size_t)b * r;
    const int rbase = blockIdx.x * (FQ_WARPS * RPW);
    const int i0 = rbase + warp * RPW;
    int imax = rbase + FQ_WARPS * RPW - 1;
    if (imax > r - 1) imax = r - 1;

    float2 y[RPW][EPL];
#pragma unroll
    for (int q = 0; q < RPW; ++q)
#pragma unroll
        for (int e = 0; e < EPL; ++e) y[q][e] = make_float2((e * 32 + lane == i0 + q) ? 1.f : 0.f, 0.f);

    int jstart = imax - 1;
    if (jstart > r - 3) jstart = r - 3;
    for (int jt = jstart; jt >= 0; jt -= FQ_TJ) {
        __syncthreads();
        for (int idx = tid; idx < FQ_TJ * WIDTH; idx += FQ_THREADS) {
            const int t = idx / WIDTH, k = idx - t * WIDTH, j = jt - t;
            float2 v = make_float2(0.f, 0.f);
            if (j >= 0 && k > j && k < r) v = M[(size_t)j * ld + k];
            fq_sv[idx] = v;
        }
        if (tid < FQ_TJ) stau[tid] = (jt - tid >= 0) ? taus[jt - tid] : 0.f;
        __syncthreads();
#pragma unroll 1
        for (int t = 0; t < FQ_TJ; ++t) {
            const int j = jt - t;
            if (j < 0) break;
            const float tau = stau[t];
            if (tau == 0.f || i0 + RPW - 1 <= j) continue;
            // rows i <= j are not touched by reflector j (rows ascend with q)
            fq_dispatch<EPL, RPW>((j + 1) >> 5, y, fq_sv + t * WIDTH, tau, lane, j + 1 - i0);
        }
    }
    float2* X = Xall + (size_t)b * r * r;
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
        const int i = i0 + q;
        if (i >= r) continue;
        const float2 p = phall[(size_t)b * r + i];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int k = e * 32 + lane;
            if (k < r) X[(size_t)i * r + k] = cmulf(y[q][e], p);
        }
    }
}

__device__ __forceinline__ float2 lds_v2(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(unsigned a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ float rsqrt_ftz(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- 3. implicit QL on the real tridiagonal, one warp per matrix: lane 0 runs the (inherently serial) chase and
//         records each rotation. The loop is latency bound (one active lane, in-order issue), so (d_i, e_i) are
//         interleaved and fetched with one 64-bit load one rotation ahead, the hypot-underflow test and the Newton
//         correction of the reciprocal square root sit off the dependent chain, and nothing but the recurrence and
//         one store of (c, s) happens per rotation.
//         Scheduling: the rotation of sweep s on rows (i, i+1) gets level key_s - i with
//         key_s = max(key_{s-1} + 2, top_s + 1, end level of sweep s - NS + top_s): consecutive sweeps trail each
//         other by two rows, so the rotations of one level touch disjoint row pairs, every rotation comes after all
//         earlier ones on its rows, and sweep s can reuse the application slot of sweep s - NS. --------------------
struct SweepRec {
    int q0, top, cnt, base;  // first rotation, first row i, rotations, level of the first rotation
    float2 p[8];             // the first eight (c, s): saves the consumer a dependent load
};

__global__ void __launch_bounds__(32) tql_kernel(int r, const float* __restrict__ dall, const float* __restrict__ eall,
                                                 float* __restrict__ lamall, float2* __restrict__ csall,
                                                 SweepRec* __restrict__ swall, int32_t* __restrict__ metaall, int cap,
                                                 int scap, int lcap, const int32_t* __restrict__ skip, int maxit, int nslots) {
    extern __shared__ float ql_sm[];
    float2* de = reinterpret_cast<float2*>(ql_sm) + 1;  // de[i] = (d_i, e_i), i = -1 .. r-1 (de[-1] is a pad)
    int* endlv = reinterpret_cast<int*>(ql_sm + 2 * (r + 2));  // [nslots] end level of the last sweep of each slot
    const int b = blockIdx.x, lane = threadIdx.x;
    if (skip && skip[b]) return;
    float2* cs = csall + (size_t)b * cap;
    SweepRec* sw = swall + (size_t)b * scap;
    // QL deflates from the top and needs the large entries at the bottom (LAPACK steqr picks QL or QR by the same
    // test); the tridiagonalisation above leaves them at the top, so the iteration usually runs on the index-reversed
    // matrix: d'[i] = d[r-1-i], e'[i] = e[r-2-i]. The consumer then loads the rows of Xt in reversed order.
    const bool rev = r > 1 && fabsf(dall[(size_t)b * r]) > fabsf(dall[(size_t)b * r + r - 1]);
    for (int i = lane; i < r; i += 32) {
        const int src = rev ? r - 1 - i : i;
        const float ee = rev ? (i < r - 1 ? eall[(size_t)b * r + r - 2 - i] : 0.f) : eall[(size_t)b * r + i];
        de[i] = make_float2(dall[(size_t)b * r + src], ee);
    }
    for (int i = lane; i < nslots; i += 32) endlv[i] = 0;
    if (lane == 0) de[-1] = make_float2(0.f, 0.f);
    __syncwarp();
    const unsigned de_sa = (unsigned)__cvta_generic_to_shared(de);
    int nrot = 0, ns = 0, key = -1, nlev = 0, iters = 0, status = 0;
    for (int l = 0; l < r && status == 0; ++l) {
        int it = 0;
        while (true) {
            // smallest m >= l with a negligible e[m] (warp-cooperative scan)
            int m = r - 1;
            for (int base = l; base < r - 1; base += 32) {
                const int idx = base + lane;
                bool ok = false;
                if (idx < r - 1) {
                    const float2 a = de[idx];
                    const float dd = fabsf(a.x) + fabsf(de[idx + 1].x);
                    ok = (fabsf(a.y) + dd == dd);
                }
                const unsigned mask = __ballot_sync(0xffffffffu, ok);
                if (mask) {
                    m = base + __ffs(mask) - 1;
                    break;
                }
            }
            if (m == l) break;
            if (++it > maxit) {
                status = 2;
                break;
            }
            ++iters;
            int st0 = 0;
            if (lane == 0) {
                if (nrot + (m - l) > cap || ns >= scap) {
                    st0 = 1;
                } else {
                    const float dl = de[l].x, el = de[l].y;
                    float g = (de[l + 1].x - dl) / (2.f * el);
                    float rr = sqrtf(fmaf(g, g, 1.f));
                    g = de[m].x - dl + el / (g + copysignf(rr, g));
                    float s = 1.f, c = 1.f, p = 0.f;
                    bool early = false;
                    int i = m - 1;
                    const int top = i;
                    float d_hi = de[m].x, d_lo = de[i].x, e_i = de[i].y;
                    unsigned a = de_sa + 8u * (unsigned)i;  // shared address of de[i]
                    float2* csp = cs + nrot;
#pragma unroll 2
                    for (; i >= l; --i) {
                        const float2 nx = lds_v2(a - 8u);  // (d, e)[i-1], used by the next rotation
                        const float f = s * e_i, bb2 = 2.f * c * e_i, ff = f * f;
                        const float x = fmaf(g, g, ff);
                        const float y0 = rsqrt_ftz(x);
                        // y = y0 (1 + hh), hh = (1 - x y0^2) / 2: one Newton step, applied to the products
                        const float hh = fmaf(-0.5f * x * y0, y0, 0.5f);
                        const float c0 = g * y0, s0 = f * y0, r0 = x * y0;
                        const float cn = fmaf(c0, hh, c0), sn = fmaf(s0, hh, s0);
                        const float g1 = d_hi - p;
                        rr = fmaf(d_lo - g1, sn, cn * bb2);
                        const float pn = sn * rr;
                        if (x < 1e-36f) {  // hypot underflow: the reference algorithm's r == 0 recovery
                            sts_f32(a + 12u, 0.f);
                            sts_f32(a + 8u, d_hi - p);
                            de[m].y = 0.f;
                            early = true;
                            break;
                        }
                        sts_f32(a + 12u, fmaf(r0, hh, r0));  // e[i+1]
                        sts_f32(a + 8u, g1 + pn);           // d[i+1]
                        g = fmaf(cn, rr, -0.5f * bb2);
                        s = sn, c = cn, p = pn;
                        *csp++ = make_float2(cn, sn);
                        d_hi = d_lo, d_lo = nx.x, e_i = nx.y;
                        a -= 8u;
                    }
                    if (!early) {
                        de[l].x = d_hi - p;
                        de[l].y = g;
                        de[m].y = 0.f;
                    }
                    const int applied = top - i;
                    if (applied > 0) {
                        const int slot = ns % nslots;
                        key = max(max(key + 2, top + 1), endlv[slot] + top);
                        const int base = key - top;
                        endlv[slot] = base + applied;
                        nlev = max(nlev, base + applied - 1);
                        SweepRec* rec = sw + ns;
                        rec->q0 = nrot, rec->top = top, rec->cnt = applied, rec->base = base;
                        ++ns;
                        nrot += applied;
                    }
                }
            }
            __syncwarp();
            if (__shfl_sync(0xffffffffu, st0, 0)) {
                status = 1;
                break;
            }
        }
    }
    nrot = __shfl_sync(0xffffffffu, nrot, 0);
    ns = __shfl_sync(0xffffffffu, ns, 0);
    nlev = __shfl_sync(0xffffffffu, nlev, 0);
    if (status == 0 && nlev > lcap) status = 1;
    for (int i = lane; i < r; i += 32) lamall[(size_t)b * r + i] = de[i].x;
    if (lane == 0) {
        int32_t* meta = metaall + (size_t)b * 4;
        meta[0] = ns;
        meta[1] = nlev;
        meta[2] = iters;
        meta[3] = status | (rev ? 0x100 : 0);
    }
    if (status != 0) return;
    // copy the first eight (c, s) of every sweep into its record
    for (int q = lane; q < ns; q += 32) {
        SweepRec* rec = sw + q;
        const int q0 = rec->q0, cnt = rec->cnt;
#pragma unroll
        for (int u = 0; u < 8; ++u) rec->p[u] = u < cnt ? cs[q0 + u] : make_float2(1.f, 0.f);
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- 4. apply the rotations to a slab of RA_C columns of Xt held in shared memory, level by level (one barrier per
//         level). Slot q (RA_C lanes) follows the sweeps q, q + NS, q + 2 NS, ...: at level base + u it rotates rows
//         (top - u, top - u + 1); the row shared by consecutive rotations stays in a register, so a rotation costs one
//         shared-memory load and one store per column. Each lane owns two adjacent columns (one float4): a slot's row is
//         128 bytes = all 32 banks, so the four slots of a warp serialise without bank conflicts (one column per lane
//         put them two rows = 128 bytes apart on the same banks). NS = 64 slots per CTA for r <= 256, 128 above.
//         (c, s) come from a 16-entry FIFO per slot that cp.async refills one block of eight ahead; the record of the
//         slot's next sweep and its first block arrive the same way in private cells. Nothing in the loop waits on a
//         global load. Then W[i][:] = lambda_i conj(Xt[i][:]). -------------------------------------------------------
template <int RA_NS>
__global__ void __launch_bounds__(RA_C * RA_NS, RA_NS == 64 ? 3 : 2) rotapply_kernel(const float2* __restrict__ Xall, int r,
                                                                 const float2* __restrict__ csall,
                                                                 const SweepRec* __restrict__ swall,
                                                                 const int32_t* __restrict__ metaall,
                                                                 const float* __restrict__ lamall,
                                                                 float2* __restrict__ Wall, int ld, size_t wstride,
                                                                 int cap, int scap, int32_t* __restrict__ done,
                                                                 int32_t* __restrict__ sweeps,
                                                                 const int32_t* __restrict__ skip) {
    extern __shared__ float4 ra_sm[];
    constexpr int RA_THREADS = RA_C * RA_NS;
    if (skip && skip[blockIdx.y]) return;
    constexpr int C = RA_C;        // lanes per slot
    constexpr int CW = 2 * RA_C;   // columns per slab
    static_assert(C == 8, "parameter blocks of eight");
    int4* hdrs = reinterpret_cast<int4*>(ra_sm);                    // [RA_THREADS] next sweep's record
    float2* nfirst = reinterpret_cast<float2*>(hdrs + RA_THREADS);  // [RA_THREADS] its first block, per lane
    float2* prm = nfirst + RA_THREADS;                              // [RA_NS][16] parameter FIFO per slot
    float4* S = reinterpret_cast<float4*>(prm + RA_NS * 16);        // slab [r][C] float4 = two complex columns
    const int b = blockIdx.y, tid = threadIdx.x;
    const int32_t* meta = metaall + (size_t)b * 4;
    const int ns = meta[0], nlev = meta[1], status = meta[3] & 0xff;
    const bool rev = (meta[3] & 0x100) != 0;
    if (blockIdx.x == 0 && tid == 0) {
        done[b] = status == 0 ? 1 : 0;
        sweeps[b] = meta[2];
    }
    if (status != 0) return;
    const int col0 = blockIdx.x * CW;
    const float2* X = Xall + (size_t)b * r * r;
    float2* S2 = reinterpret_cast<float2*>(S);
    for (int idx = tid; idx < r * CW; idx += RA_THREADS) {
        const int i = idx / CW, c2 = idx - i * CW;
        const int src = rev ? r - 1 - i : i;  // the QL iteration ran on the index-reversed tridiagonal
        S2[idx] = (col0 + c2 < r) ? X[(size_t)src * r + col0 + c2] : make_float2(0.f, 0.f);
    }
    const float2* cs = csall + (size_t)b * cap;
    const SweepRec* sw = swall + (size_t)b * scap;
    const int slot = tid / C, cc = tid % C;
    float2* myprm = prm + slot * 16;
    float4* Scc = S + cc;
    int s = slot;
    int q0 = 0, top = 0, cnt = 0, u = -0x40000000;  // u = level - base of the current sweep
    if (s < ns) {
        const SweepRec* rec = sw + s;
        q0 = rec->q0, top = rec->top, cnt = rec->cnt, u = -rec->base;
        myprm[cc] = rec->p[cc];
        if (8 + cc < cnt) cp_async8(&myprm[8 + cc], cs + q0 + 8 + cc);
    }
    if (s + RA_NS < ns) {
        const SweepRec* rec = sw + s + RA_NS;
        cp_async16(&hdrs[tid], rec);
        cp_async8(&nfirst[tid], &rec->p[cc]);
    }
    float4 hi = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int lv = 1; lv <= nlev; ++lv) {
        ++u;
        if ((unsigned)u < (unsigned)cnt) {
            const float2 p = myprm[u & 15];  // (c, s), broadcast within the slot
            const int i = top - u;
            if (u == 0) hi = Scc[(top + 1) * C];
            const float4 a = Scc[i * C];
            Scc[(i + 1) * C] = make_float4(fmaf(p.y, a.x, p.x * hi.x), fmaf(p.y, a.y, p.x * hi.y),
                                           fmaf(p.y, a.z, p.x * hi.z), fmaf(p.y, a.w, p.x * hi.w));
            hi = make_float4(fmaf(p.x, a.x, -p.y * hi.x), fmaf(p.x, a.y, -p.y * hi.y), fmaf(p.x, a.z, -p.y * hi.z),
                             fmaf(p.x, a.w, -p.y * hi.w));
            const int k8 = u & 7;
            if (u == cnt - 1) {
                Scc[i * C] = hi;
                // the slot's next sweep: its record has been in the cells since the previous switch
                s += RA_NS;
                if (s < ns) {
                    cp_async_wait_all();
                    const int4 hd = hdrs[tid];
                    q0 = hd.x, top = hd.y, cnt = hd.z, u = lv - hd.w;
                    myprm[cc] = nfirst[tid];
                    if (8 + cc < cnt) cp_async8(&myprm[8 + cc], cs + q0 + 8 + cc);
                    if (s + RA_NS < ns) {
                        const SweepRec* rec = sw + s + RA_NS;
                        cp_async16(&hdrs[tid], rec);
                        cp_async8(&nfirst[tid], &rec->p[cc]);
                    }
                } else {
                    cnt = 0;
                }
            } else if (k8 == 7) {
                cp_async_wait_all();  // block (u + 1) / 8 is complete; the level barrier publishes it to the slot
            } else if (k8 == 0 && u >= 8) {
                // block u / 8 is in use: the other half of the FIFO is free for block u / 8 + 1
                const int nb = u + 8 + cc;
                if (nb < cnt) cp_async8(&myprm[nb & 15], cs + q0 + nb);
            }
        }
        __syncthreads();
    }
    float2* Wm = Wall + (size_t)b * wstride;
    const float* lam = lamall + (size_t)b * r;
    for (int idx = tid; idx < r * CW; idx += RA_THREADS) {
        const int i = idx / CW, c2 = idx - i * CW;
        if (col0 + c2 < r) {
            const float l = lam[i];
            const float2 v = S2[idx];
            Wm[(size_t)i * ld + col0 + c2] = make_float2(l * v.x, -l * v.y);
        }
    }
}

// =====================================================================================================================
// Fixed small rank (compressionrank <= TK_MAXK): only the k leading eigenpairs are needed, so after the
// tridiagonalisation the QL iteration, the reflector accumulation and the rotation application are replaced by
//   (a) Sturm bisection for the k+1 largest eigenvalues of T (eight lanes per eigenvalue, three bits per round),
//   (b) one twisted factorisation per eigenvalue for its eigenvector of T, then modified Gram-Schmidt,
//   (c) the reflectors applied to those k vectors only.
// Close eigenvalues only mix their own vectors (angle ~ eps / gap), which Gram-Schmidt keeps orthonormal and which
// changes neither the retained subspace nor, to second order, the refined singular values; what must be avoided is two
// lanes converging to the same vector (numerically multiple eigenvalues). That is detected directly: a vector that
// loses more than 3/4 of its squared length in the Gram-Schmidt step sends the matrix to the full path, as does a
// non-positive leading eigenvalue or overflowing element growth. flag[b] = 0 makes (b)/(c) skip the matrix and the
// full-path kernels take it; flag[b] = 1 makes those skip it.
constexpr int TK_MAXK = 32;
constexpr float TK_GAP = 0.f;  // eigenvalue-gap pre-test (relative to lambda_max); 0: rely on the duplicate test

__device__ __forceinline__ int sturm_count(const float* d, const float* e2, int r, float x, float pivmin) {
    // number of eigenvalues of T smaller than x (fast division: only the signs of the pivots matter)
    int c = 0;
    float q = d[0] - x;
    if (fabsf(q) < pivmin) q = -pivmin;
    c += q < 0.f;
    for (int i = 1; i < r; ++i) {
        q = d[i] - x - __fdividef(e2[i - 1], q);
        if (fabsf(q) < pivmin) q = -pivmin;
        c += q < 0.f;
    }
    return c;
}

// eight lanes per eigenvalue: every round evaluates seven interior points of the bracket (three bits per round)
constexpr int BS_ROUNDS = 10;

__global__ void __launch_bounds__(8 * (TK_MAXK + 1) + 24) bisect_kernel(int r, int nev, const float* __restrict__ dall,
                                                                      const float* __restrict__ eall,
                                                                      float* __restrict__ lamtop,
                                                                      int32_t* __restrict__ flag, float gap) {
    extern __shared__ float bs_sm[];
    float* d = bs_sm;
    float* e2 = bs_sm + r;
    __shared__ float s_lo[9], s_hi[9], s_lam[TK_MAXK + 1];
    const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    float lo = 3.4e38f, hi = -3.4e38f;
    for (int i = tid; i < r; i += nthr) {
        const float di = dall[(size_t)b * r + i];
        const float ei = i < r - 1 ? eall[(size_t)b * r + i] : 0.f;
        const float ep = i > 0 ? eall[(size_t)b * r + i - 1] : 0.f;
        d[i] = di;
        e2[i] = ei * ei;
        const float rad = fabsf(ei) + fabsf(ep);
        lo = fminf(lo, di - rad);
        hi = fmaxf(hi, di + rad);
    }
    lo = -warp_max(-lo);
    hi = warp_max(hi);
    if (lane == 0) s_lo[tid >> 5] = lo, s_hi[tid >> 5] = hi;
    __syncthreads();
    for (int w = 0; w < (nthr + 31) / 32; ++w) lo = fminf(lo, s_lo[w]), hi = fmaxf(hi, s_hi[w]);
    const float scale = fmaxf(fabsf(lo), fabsf(hi));
    const float pivmin = fmaxf(1e-30f, 1e-14f * scale * scale);
    // eigenvalue number idx in ascending order: the smallest x with count(x) > idx
    const int g = tid >> 3, j = tid & 7;
    const int idx = r - 1 - g;
    float a = lo - 1e-6f * scale - 1e-30f, c = hi + 1e-6f * scale + 1e-30f;
    for (int it = 0; it < BS_ROUNDS; ++it) {
        const float h = 0.125f * (c - a);
        bool above = true;  // lane 7 stands for the upper end of the bracket
        if (g < nev && j < 7) above = sturm_count(d, e2, r, a + (float)(j + 1) * h, pivmin) > idx;
        const unsigned bits = (__ballot_sync(0xffffffffu, above) >> (lane & 24)) & 0xffu;
        const int js = __ffs(bits | 0x80u) - 1;
        const float a0 = a;
        a = a0 + (float)js * h;
        c = js == 7 ? c : a0 + (float)(js + 1) * h;
    }
    if (g < nev && j == 0) {
        const float lam = 0.5f * (a + c);
        s_lam[g] = lam;
        lamtop[(size_t)b * (TK_MAXK + 1) + g] = lam;
    }
    __syncthreads();
    if (tid == 0) {
        bool ok = s_lam[0] > 0.f;
        const float thr = gap * fabsf(s_lam[0]);
        for (int t = 0; t + 1 < nev; ++t) ok = ok && (s_lam[t] - s_lam[t + 1] >= thr);
        flag[b] = ok ? 1 : 0;
    }
}

// The same for MANY small problems (BASELINE configs[3]: 8320 matrices of r = 64): eight lanes per eigenvalue buy latency
// with 7/3 of the Sturm counts, and a CTA per matrix leaves a quarter of its lanes idle - 8320 CTAs of three warps were bound
// by the instructions they issue (236 us; the three leading-pair kernels 0.80 -> 0.69 ms with this one). Here a warp takes 32 / nev matrices, one lane per eigenvalue, plain bisection until
// the bracket cannot shrink; d and e^2 of the CTA's matrices sit in shared memory.
constexpr int BP_WARPS = 4;
// matrices per warp: as many as fit its lanes, at most eight (32 matrices per CTA: 32 KiB of shared memory at r = 128)
__host__ __device__ inline int bp_group(int nev) { return 32 / nev < 8 ? 32 / nev : 8; }
__global__ void __launch_bounds__(32 * BP_WARPS) bisect_packed_kernel(int B, int r, int nev, const float* __restrict__ dall,
                                                                      const float* __restrict__ eall,
                                                                      float* __restrict__ lamtop, int32_t* __restrict__ flag,
                                                                      float gap) {
    extern __shared__ float bs_sm[];   // [matrices of the CTA][2 r]: d, e^2
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = bp_group(nev), mpc = G * BP_WARPS;
    const int b0 = blockIdx.x * mpc;
    for (int idx = tid; idx < mpc * r; idx += 32 * BP_WARPS) {
        const int ml = idx / r, i = idx - ml * r, b = b0 + ml;
        float di = 0.f, ei = 0.f;
        if (b < B) {
            di = dall[(size_t)b * r + i];
            ei = i < r - 1 ? eall[(size_t)b * r + i] : 0.f;
        }
        bs_sm[ml * 2 * r + i] = di;
        bs_sm[ml * 2 * r + r + i] = ei * ei;
    }
    __syncthreads();
    const int g = lane / nev, t = lane - g * nev;
    const int ml = warp * G + g, b = b0 + ml;
    const bool act = g < G && b < B;
    float lam = 0.f;
    if (act) {
        const float* d = bs_sm + ml * 2 * r;
        const float* e2 = d + r;
        float lo = 3.4e38f, hi = -3.4e38f, ep = 0.f;
        for (int i = 0; i < r; ++i) {
            const float ei = sqrtf(e2[i]);
            const float rad = ei + ep;
            lo = fminf(lo, d[i] - rad);
            hi = fmaxf(hi, d[i] + rad);
            ep = ei;
        }
        const float scale = fmaxf(fabsf(lo), fabsf(hi));
        const float pivmin = fmaxf(1e-30f, 1e-14f * scale * scale);
        const int idx = r - 1 - t;  // t-th largest: the smallest x with count(x) > idx
        float a = lo - 1e-6f * scale - 1e-30f, c = hi + 1e-6f * scale + 1e-30f;
        for (int it = 0; it < 48; ++it) {
            const float mid = 0.5f * (a + c);
            if (!(mid > a && mid < c)) break;
            if (sturm_count(d, e2, r, mid, pivmin) > idx) c = mid;
            else a = mid;
        }
        lam = 0.5f * (a + c);
        lamtop[(size_t)b * (TK_MAXK + 1) + t] = lam;
    }
    // flag[b]: leading eigenvalue positive, gaps at least gap * |lambda_0| (as bisect_kernel)
    const int base = (g < G ? g : 0) * nev;
    const float l0 = __shfl_sync(0xffffffffu, lam, base);
    bool ok = l0 > 0.f;
    const float thr = gap * fabsf(l0);
    float prev = l0;
    for (int q = 1; q < nev; ++q) {
        const float lq = __shfl_sync(0xffffffffu, lam, base + q);
        ok = ok && (prev - lq >= thr);
        prev = lq;
    }
    if (act && t == 0) flag[b] = ok ? 1 : 0;
}

// Energy rule with a small resulting rank: every eigenvalue by plain bisection (one thread each), then the rank the
// selection stage will find, estimated from the same eigenvalues: kvec[b] = that rank + 2 (ties may move it by one),
// flag[b] = 1 when the leading-pair path can deliver that many vectors.
__global__ void __launch_bounds__(1024) bisect_all_kernel(int r, double energy, int klimit, const float* __restrict__ dall,
                                                          const float* __restrict__ eall, float* __restrict__ lamall,
                                                          float* __restrict__ lamtop, int32_t* __restrict__ kvec,
                                                          int32_t* __restrict__ flag) {
    extern __shared__ float bs_sm[];
    float* d = bs_sm;
    float* e2 = bs_sm + r;
    float* lam = bs_sm + 2 * r;
    __shared__ float s_lo[32], s_hi[32];
    const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    float lo = 3.4e38f, hi = -3.4e38f;
    for (int i = tid; i < r; i += nthr) {
        const float di = dall[(size_t)b * r + i];
        const float ei = i < r - 1 ? eall[(size_t)b * r + i] : 0.f;
        const float ep = i > 0 ? eall[(size_t)b * r + i - 1] : 0.f;
        d[i] = di;
        e2[i] = ei * ei;
        const float rad = fabsf(ei) + fabsf(ep);
        lo = fminf(lo, di - rad);
        hi = fmaxf(hi, di + rad);
    }
    lo = -warp_max(-lo);
    hi = warp_max(hi);
    if (lane == 0) s_lo[tid >> 5] = lo, s_hi[tid >> 5] = hi;
    __syncthreads();
    for (int w = 0; w < (nthr + 31) / 32; ++w) lo = fminf(lo, s_lo[w]), hi = fmaxf(hi, s_hi[w]);
    const float scale = fmaxf(fabsf(lo), fabsf(hi));
    const float pivmin = fmaxf(1e-30f, 1e-14f * scale * scale);
    for (int t = tid; t < r; t += nthr) {
        const int idx = r - 1 - t;  // t-th largest
        float a = lo - 1e-6f * scale - 1e-30f, c = hi + 1e-6f * scale + 1e-30f;
        for (int it = 0; it < 48; ++it) {
            const float mid = 0.5f * (a + c);
            if (!(mid > a && mid < c)) break;
            if (sturm_count(d, e2, r, mid, pivmin) > idx) c = mid;
            else a = mid;
        }
        const float l = 0.5f * (a + c);
        lam[t] = l;
        lamall[(size_t)b * r + t] = l;
        if (t <= TK_MAXK) lamtop[(size_t)b * (TK_MAXK + 1) + t] = l;
    }
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int t = 0; t < r; ++t) tot += fmax((double)lam[t], 0.0);
        const double thr = energy * tot;
        double cum = 0.0;
        int k = r;
        for (int t = 0; t < r; ++t) {
            cum += fmax((double)lam[t], 0.0);
            if (cum >= thr) {
                k = t + 1;
                break;
            }
        }
        const int kv = k + 2;
        kvec[b] = kv < r ? kv : r;
        flag[b] = (lam[0] > 0.f && kv <= klimit && kv <= r - 2) ? 1 : 0;
    }
}

// eigenvectors of T by twisted factorisation (one lane per eigenvalue), then modified Gram-Schmidt over the k vectors
__global__ void __launch_bounds__(32) twisted_kernel(int r, int kuni, const int32_t* __restrict__ kvec,
                                                     const float* __restrict__ dall, const float* __restrict__ eall,
                                                     const float* __restrict__ lamtop, int32_t* __restrict__ flag,
                                                     float* __restrict__ zall, float* __restrict__ dmall) {
    const int b = blockIdx.x, lane = threadIdx.x;
    if (flag[b] == 0) return;
    const int k = kvec ? kvec[b] : kuni;  // energy rule: per matrix
    const float* d = dall + (size_t)b * r;
    const float* e = eall + (size_t)b * r;
    float* zb = zall + (size_t)b * TK_MAXK * r;
    int bad = 0;
    if (lane < k) {
        const float lam = lamtop[(size_t)b * (TK_MAXK + 1) + lane];
        float* z = zb + (size_t)lane * r;
        float* dm = dmall + ((size_t)b * TK_MAXK + lane) * r;
        const float pivmin = fmaxf(1e-30f, 1e-14f * lam * lam);
        // backward pivots dm[i] of U D- U^T = T - lam
        float q = d[r - 1] - lam;
        if (fabsf(q) < pivmin) q = -pivmin;
        dm[r - 1] = q;
        for (int i = r - 2; i >= 0; --i) {
            const float ei = e[i];
            q = d[i] - lam - ei * ei / q;
            if (fabsf(q) < pivmin) q = -pivmin;
            dm[i] = q;
        }
        // forward pivots dp[i] of L D+ L^T (kept in z for now) and the twist index: argmin |gamma_i|,
        // gamma_i = dp[i] + dm[i] - (d[i] - lam)
        float p = d[0] - lam;
        if (fabsf(p) < pivmin) p = -pivmin;
        z[0] = p;
        float best = fabsf(dm[0]);  // gamma_0 = dm[0]
        int kt = 0;
        for (int i = 1; i < r; ++i) {
            const float ei = e[i - 1];
            p = d[i] - lam - ei * ei / p;
            if (fabsf(p) < pivmin) p = -pivmin;
            z[i] = p;
            const float gam = fabsf(p + dm[i] - (d[i] - lam));
            if (gam < best) best = gam, kt = i;
        }
        // z[kt] = 1; upward with the forward pivots, downward with the backward ones
        float nrm = 1.f, zi = 1.f;
        for (int i = kt - 1; i >= 0; --i) {
            zi = -(e[i] / z[i]) * zi;  // z[i] still holds dp[i]
            z[i] = zi;
            nrm = fmaf(zi, zi, nrm);
        }
        zi = 1.f;
        for (int i = kt + 1; i < r; ++i) {
            zi = -(e[i - 1] / dm[i]) * zi;
            z[i] = zi;
            nrm = fmaf(zi, zi, nrm);
        }
        z[kt] = 1.f;
        const float sc = rsqrtf(nrm);
        for (int i = 0; i < r; ++i) z[i] *= sc;
        if (!(nrm < 3e38f)) bad = 1;  // element growth overflowed: leave the matrix to the full path
    }
    if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0) flag[b] = 0;
        return;
    }
    // modified Gram-Schmidt in order of decreasing eigenvalue (all lanes cooperate, entries strided over lanes)
    for (int t = 0; t < k; ++t) {
        float* zt = zb + (size_t)t * r;
        for (int s = 0; s < t; ++s) {
            const float* zs = zb + (size_t)s * r;
            float dot = 0.f;
            for (int i = lane; i < r; i += 32) dot = fmaf(zs[i], zt[i], dot);
            dot = warp_sum(dot);
            for (int i = lane; i < r; i += 32) zt[i] = fmaf(-dot, zs[i], zt[i]);
            __syncwarp();
        }
        float nn = 0.f;
        for (int i = lane; i < r; i += 32) nn = fmaf(zt[i], zt[i], nn);
        nn = warp_sum(nn);
        if (!(nn >= 0.25f)) {
            // this vector (unit length before) was mostly a copy of earlier ones: numerically multiple eigenvalue,
            // the twisted factorisations converged to the same vector. The full path handles clusters.
            if (lane == 0) flag[b] = 0;
            return;
        }
        const float sc = nn > 0.f ? rsqrtf(nn) : 0.f;
        for (int i = lane; i < r; i += 32) zt[i] *= sc;
        __syncwarp();
    }
}

// v_t = Q D z_t for the k vectors of one matrix (one CTA; warp w takes vectors w*RPW ..), then
// W[t][:] = lambda_t conj(v_t) for t < k and zero rows below: what the selection and factor stages expect.
template <int EPL, int RPW>
__global__ void __launch_bounds__(FQ_THREADS, (EPL * RPW <= 32) ? 2 : 1)
    backtr_kernel(float2* __restrict__ Wall, int r, int ld, size_t wstride, int kuni, const int32_t* __restrict__ kvec,
                  const float* __restrict__ lamall, const float* __restrict__ tauall, const float2* __restrict__ phall,
                  const float* __restrict__ zall, const float* __restrict__ lamtop, const int32_t* __restrict__ flag,
                  int32_t* __restrict__ done, int32_t* __restrict__ sweeps) {
    constexpr int WIDTH = EPL * 32;
    extern __shared__ float2 fq_sv[];  // [FQ_TJ][WIDTH]
    __shared__ float stau[FQ_TJ];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (flag[b] == 0) return;
    const int k = kvec ? kvec[b] : kuni;
    float2* M = Wall + (size_t)b * wstride;
    const float* taus = tauall + (size_t)b * r;
    const int i0 = warp * RPW;
    float2 y[RPW][EPL];
#pragma unroll
    for (int q = 0; q < RPW; ++q)
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int kk = e * 32 + lane;
            float2 v = make_float2(0.f, 0.f);
            if (i0 + q < k && kk < r) {
                const float z = zall[((size_t)b * TK_MAXK + i0 + q) * r + kk];
                const float2 p = phall[(size_t)b * r + kk];
                v = make_float2(z * p.x, z * p.y);
            }
            y[q][e] = v;
        }
    for (int jt = r - 3; jt >= 0; jt -= FQ_TJ) {
        __syncthreads();
        for (int idx = tid; idx < FQ_TJ * WIDTH; idx += FQ_THREADS) {
            const int t = idx / WIDTH, kk = idx - t * WIDTH, j = jt - t;
            float2 v = make_float2(0.f, 0.f);
            if (j >= 0 && kk > j && kk < r) v = M[(size_t)j * ld + kk];
            fq_sv[idx] = v;
        }
        if (tid < FQ_TJ) stau[tid] = (jt - tid >= 0) ? taus[jt - tid] : 0.f;
        __syncthreads();
        if (i0 >= k) continue;
#pragma unroll 1
        for (int t = 0; t < FQ_TJ; ++t) {
            const int j = jt - t;
            if (j < 0) break;
            const float tau = stau[t];
            if (tau == 0.f) continue;
            fq_dispatch<EPL, RPW>((j + 1) >> 5, y, fq_sv + t * WIDTH, tau, lane, 0);
        }
    }
    __syncthreads();  // every reflector has been read: the rows of M can be overwritten
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
        const int t = i0 + q;
        if (t >= k) continue;
        const float lam = lamtop[(size_t)b * (TK_MAXK + 1) + t];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int kk = e * 32 + lane;
            if (kk < r) M[(size_t)t * ld + kk] = make_float2(lam * y[q][e].x, -lam * y[q][e].y);
        }
    }
    // rows below: zero, or (energy rule: the selection stage needs every eigenvalue) a vector of norm lambda_row
    for (int idx = tid; idx < (r - k) * r; idx += FQ_THREADS) {
        const int row = k + idx / r, kk = idx % r;
        M[(size_t)row * ld + kk] = make_float2((lamall && kk == 0) ? lamall[(size_t)b * r + row] : 0.f, 0.f);
    }
    if (tid == 0) {
        done[b] = 1;
        sweeps[b] = 0;
    }
}


// =====================================================================================================================
// Full spectrum without the QL iteration ("eigvec_impl" = 0, r a multiple of 16, r >= 128): after the tridiagonalisation
//   (a) every eigenvalue of T by Sturm bisection (one thread per eigenvalue),
//   (b) one twisted factorisation per eigenvalue for its eigenvector of T (one thread per eigenvector, no
//       reorthogonalisation): Zt[t][:] = z_t, stored as complex numbers with zero imaginary part,
//   (c) E = Zt Zt^T - I on the tensor cores (gram_tc.cu); max |E| per matrix decides: below EV_ORTHO_LIMIT one
//       Newton-Schulz step Zt <- (I - E/2) Zt (tcgen05 GEMM) restores orthonormality to ~1e-6, above it (numerically
//       multiple eigenvalues: exactly rank-deficient input, repeated singular values) the matrix is left to the QL path,
//   (d) Xt = Zt (Q D)^T as one tcgen05 GEMM against the accumulated reflectors, written as lambda_t conj(Xt[t][:]).
// Measured in float32 (tools/proto_mrrr.py) on signal + noise and noise-only Gram matrices of r = 256 / 512: max |E| =
// 6e-5 .. 3e-4 before and 6e-7 .. 1e-6 after the Newton-Schulz step; residual |T z - lambda z| <= 6e-7 |T|. Close
// eigenvalues only mix their own vectors, which changes neither a retained subspace nor (to second order) the singular
// values that the factor stage refines from the matrix itself. Replaces the serial QL chase (tql_kernel) and the
// barrier-bound rotation application (rotapply_kernel) by work that is either embarrassingly parallel or a GEMM.
constexpr float EV_ORTHO_LIMIT = 0.05f;
constexpr int EV_THREADS = 128;

__global__ void __launch_bounds__(EV_THREADS) bisect_full_kernel(int r, const float* __restrict__ dall,
                                                                 const float* __restrict__ eall,
                                                                 float* __restrict__ lamall) {
    extern __shared__ float bs_sm[];
    float* d = bs_sm;
    float* e2 = bs_sm + r;
    __shared__ float s_lo[EV_THREADS / 32], s_hi[EV_THREADS / 32];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    float lo = 3.4e38f, hi = -3.4e38f;
    for (int i = tid; i < r; i += EV_THREADS) {
        const float di = dall[(size_t)b * r + i];
        const float ei = i < r - 1 ? eall[(size_t)b * r + i] : 0.f;
        const float ep = i > 0 ? eall[(size_t)b * r + i - 1] : 0.f;
        d[i] = di;
        e2[i] = ei * ei;
        const float rad = fabsf(ei) + fabsf(ep);
        lo = fminf(lo, di - rad);
        hi = fmaxf(hi, di + rad);
    }
    lo = -warp_max(-lo);
    hi = warp_max(hi);
    if (lane == 0) s_lo[tid >> 5] = lo, s_hi[tid >> 5] = hi;
    __syncthreads();
    for (int w = 0; w < EV_THREADS / 32; ++w) lo = fminf(lo, s_lo[w]), hi = fmaxf(hi, s_hi[w]);
    const float scale = fmaxf(fabsf(lo), fabsf(hi));
    const float pivmin = fmaxf(1e-30f, 1e-14f * scale * scale);
    const int t = blockIdx.x * EV_THREADS + tid;  // t-th largest eigenvalue
    float a = lo - 1e-6f * scale - 1e-30f, c = hi + 1e-6f * scale + 1e-30f;
    // first level shared by the block: Sturm counts at EV_THREADS equispaced points, one per thread; every thread then starts
    // from the grid cell that holds its eigenvalue (saves log2(EV_THREADS) of its own bisection steps)
    __shared__ int s_cnt[EV_THREADS + 1];
    const float g0 = a, gh = (c - a) * (1.f / EV_THREADS);
    s_cnt[tid + 1] = tid + 1 < EV_THREADS ? sturm_count(d, e2, r, g0 + gh * (float)(tid + 1), pivmin) : r;
    if (tid == 0) s_cnt[0] = 0;
    __syncthreads();
    if (t >= r) return;
    const int idx = r - 1 - t;
    {
        int q0 = 0, q1 = EV_THREADS;          // invariant: s_cnt[q0] <= idx < s_cnt[q1]
        while (q1 - q0 > 1) {
            const int qm = (q0 + q1) >> 1;
            if (s_cnt[qm] > idx) q1 = qm;
            else q0 = qm;
        }
        if (q1 < EV_THREADS) c = g0 + gh * (float)q1;
        if (q0 > 0) a = g0 + gh * (float)q0;
    }
    for (int it = 0; it < 48; ++it) {
        const float mid = 0.5f * (a + c);
        if (!(mid > a && mid < c)) break;
        if (sturm_count(d, e2, r, mid, pivmin) > idx) c = mid;
        else a = mid;
    }
    lamall[(size_t)b * r + t] = 0.5f * (a + c);
}

// One thread per eigenvector. The pivots of both factorisations live in global scratch laid out [i][t] (coalesced over
// the eigenvalue index t); the finished vector is transposed through shared memory so that Zt[t][:] is written in
// 128-byte rows. flag[b] is cleared when a vector overflows.
__global__ void __launch_bounds__(EV_THREADS) twisted_full_kernel(int r, const float* __restrict__ dall,
                                                                  const float* __restrict__ eall,
                                                                  const float* __restrict__ lamall,
                                                                  float* __restrict__ dpall, float* __restrict__ dmall,
                                                                  float2* __restrict__ Ztall, int32_t* __restrict__ flag) {
    extern __shared__ float tw_sm[];
    float* d = tw_sm;
    float* e = tw_sm + r;
    float* tile = tw_sm + 2 * r;  // [EV_THREADS / 32][32][33]
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < r; i += EV_THREADS) {
        d[i] = dall[(size_t)b * r + i];
        e[i] = i < r - 1 ? eall[(size_t)b * r + i] : 0.f;
    }
    __syncthreads();
    const int t0 = blockIdx.x * EV_THREADS;
    const int t = t0 + tid;
    const bool live = t < r;
    float* dp = dpall + (size_t)b * r * r + t;  // element i at dp[i * r]
    float* dm = dmall + (size_t)b * r * r + t;
    float sc = 0.f;
    int bad = 0;
    if (live) {
        const float lam = lamall[(size_t)b * r + t];
        const float pivmin = fmaxf(1e-30f, 1e-14f * lam * lam);
        float q = d[r - 1] - lam;
        if (fabsf(q) < pivmin) q = -pivmin;
        dm[(size_t)(r - 1) * r] = q;
        for (int i = r - 2; i >= 0; --i) {
            const float ei = e[i];
            q = d[i] - lam - ei * ei / q;
            if (fabsf(q) < pivmin) q = -pivmin;
            dm[(size_t)i * r] = q;
        }
        float p = d[0] - lam;
        if (fabsf(p) < pivmin) p = -pivmin;
        dp[0] = p;
        float best = fabsf(q);  // gamma_0 = dm[0]
        int kt = 0;
        // (the pivots written above come back from L2: they are fetched eight steps ahead of the recurrences that use them -
        // fetched inside the dependent loops the kernel ran at 12 % issue utilisation, latency bound)
        constexpr int PF = 8;
        for (int i0 = 1; i0 < r; i0 += PF) {
            float dmv[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) dmv[u] = (i0 + u < r) ? dm[(size_t)(i0 + u) * r] : 0.f;
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int i = i0 + u;
                if (i < r) {
                    const float ei = e[i - 1];
                    p = d[i] - lam - ei * ei / p;
                    if (fabsf(p) < pivmin) p = -pivmin;
                    dp[(size_t)i * r] = p;
                    const float gam = fabsf(p + dmv[u] - (d[i] - lam));
                    if (gam < best) best = gam, kt = i;
                }
            }
        }
        // z[kt] = 1; upward with the forward pivots, downward with the backward ones (z overwrites dp)
        float nrm = 1.f, zi = 1.f;
        for (int i0 = kt - 1; i0 >= 0; i0 -= PF) {
            float dpv[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) dpv[u] = (i0 - u >= 0) ? dp[(size_t)(i0 - u) * r] : 1.f;
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int i = i0 - u;
                if (i >= 0) {
                    zi = -(e[i] / dpv[u]) * zi;
                    dp[(size_t)i * r] = zi;
                    nrm = fmaf(zi, zi, nrm);
                }
            }
        }
        zi = 1.f;
        for (int i0 = kt + 1; i0 < r; i0 += PF) {
            float dmv[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) dmv[u] = (i0 + u < r) ? dm[(size_t)(i0 + u) * r] : 1.f;
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int i = i0 + u;
                if (i < r) {
                    zi = -(e[i - 1] / dmv[u]) * zi;
                    dp[(size_t)i * r] = zi;
                    nrm = fmaf(zi, zi, nrm);
                }
            }
        }
        dp[(size_t)kt * r] = 1.f;
        sc = rsqrtf(nrm);
        if (!(nrm < 3e38f)) bad = 1, sc = 0.f;
    }
    if (bad) flag[b] = 0;
    // transpose: this warp's 32 vectors, 32 entries at a time
    float* tl = tile + warp * 32 * 33;
    float2* Zt = Ztall + (size_t)b * r * r;
    const int tw0 = t0 + warp * 32;
    for (int i0 = 0; i0 < r; i0 += 32) {
        __syncwarp();
#pragma unroll 4
        for (int ii = 0; ii < 32; ++ii) {
            const int i = i0 + ii;
            tl[ii * 33 + lane] = (live && i < r) ? dp[(size_t)i * r] * sc : 0.f;
        }
        __syncwarp();
#pragma unroll 4
        for (int tt = 0; tt < 32; ++tt) {
            const int i = i0 + lane;
            if (tw0 + tt < r && i < r) Zt[(size_t)(tw0 + tt) * r + i] = make_float2(tl[lane * 33 + tt], 0.f);
        }
    }
}

// W holds Zt Zt^T (real part; from gram_tc). In place: P = 1.5 I - 0.5 Re W (imaginary part zero), the left factor of the
// Newton-Schulz step. flag[b] &= (max |Re W - I| <= EV_ORTHO_LIMIT); done / sweeps are preset for the matrices that stay.
__global__ void __launch_bounds__(256) nsprep_kernel(float2* __restrict__ Wall, int r, int32_t* __restrict__ flag,
                                                     unsigned* __restrict__ emax) {
    const int b = blockIdx.y;
    float2* W = Wall + (size_t)b * r * r;
    float mx = 0.f;
    const size_t n = (size_t)r * r;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / r), j = (int)(idx - (size_t)i * r);
        const float g = W[idx].x;
        const float ee = g - (i == j ? 1.f : 0.f);
        const float ae = fabsf(ee);
        mx = (ae > mx || !(ae == ae)) ? (ae == ae ? ae : 3e38f) : mx;
        W[idx] = make_float2((i == j ? 1.f : 0.f) - 0.5f * ee, 0.f);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) atomicMax(emax + b, __float_as_uint(mx));  // non-negative floats order like unsigned
}
__global__ void __launch_bounds__(128) nsflag_kernel(const unsigned* __restrict__ emax, int B, int32_t* __restrict__ flag,
                                                     int32_t* __restrict__ done, int32_t* __restrict__ sweeps) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float mx = __uint_as_float(emax[b]);
    const int ok = (flag[b] != 0 && mx <= EV_ORTHO_LIMIT) ? 1 : 0;
    flag[b] = ok;
    if (ok) {
        done[b] = 1;
        sweeps[b] = 0;
    }
}
__global__ void __launch_bounds__(128) fill_i32_kernel(int32_t* __restrict__ p, int n, int32_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

inline size_t al(size_t x) { return (x + 255) / 256 * 256; }

struct EigScratch {
    size_t X, d, e, tau, ph, lam, cs, sw, meta, lamtop, flag, kvec, z, dm, zt, zt2, emax, total;
    int cap, scap, lcap;
};

EigScratch eig_layout(int B, int r) {
    EigScratch s;
    s.cap = (int)(((size_t)3 * r * r) / 2 + 256);
    s.scap = 8 * r + 64;
    s.lcap = 1 << 30;
    size_t off = 0;
    s.X = off, off += al((size_t)B * r * r * 8);
    s.d = off, off += al((size_t)B * r * 4);
    s.e = off, off += al((size_t)B * r * 4);
    s.tau = off, off += al((size_t)B * r * 4);
    s.ph = off, off += al((size_t)B * r * 8);
    s.lam = off, off += al((size_t)B * r * 4);
    s.cs = off, off += al((size_t)B * s.cap * 8);
    s.sw = off, off += al((size_t)B * s.scap * sizeof(SweepRec));
    s.meta = off, off += al((size_t)B * 16);
    s.lamtop = off, off += al((size_t)B * (TK_MAXK + 1) * 4);
    s.flag = off, off += al((size_t)B * 4);
    s.kvec = off, off += al((size_t)B * 4);
    s.z = off, off += al((size_t)B * TK_MAXK * r * 4);
    s.dm = off, off += al((size_t)B * TK_MAXK * r * 4);
    // full-spectrum eigenvector path: Zt and its Newton-Schulz update (complex r x r each); the pivots of the twisted
    // factorisations (2 r^2 floats) reuse the rotation store `cs` (1.5 r^2 + 256 float2), which only the QL path writes
    s.zt = off, off += al((size_t)B * r * r * 8);
    s.zt2 = off, off += al((size_t)B * r * r * 8);
    s.emax = off, off += al((size_t)B * 4);
    s.total = off;
    return s;
}

template <int EPL, int RB>
int launch_tridiag(vk_context* h, cudaStream_t st, float2* W, int B, int r, int ld, size_t wstride, float* d, float* e,
                   float* tau, float2* ph) {
    // vectors + the largest trailing block that fits next to them (kept in shared memory for the last nts steps)
    const size_t vec = (size_t)5 * EPL * 32 * sizeof(float2);
    int nts = (int)sqrt((double)(VK_SMEM_BUDGET - vec) / sizeof(float2));
    if (nts > r) nts = r;
    // measured: 0.82 -> 0.69 ms at r = 160, 2.11 -> 2.04 ms at r = 256 (112 matrices); at r = 512 the 200 KB of shared
    // memory cost the global phase its L1 (35.5 -> 40.9 ms per 296 matrices), so large matrices stay all-global
    if (r > 384) nts = 0;
    const size_t smem = vec + (size_t)nts * nts * sizeof(float2);
    VK_CUDA(h, cudaFuncSetAttribute(tridiag_kernel<EPL, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tridiag_kernel<EPL, RB><<<B, TD_THREADS, smem, st>>>(W, r, ld, wstride, d, e, tau, ph, nts);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

template <int EPL, int RB>
int launch_tridiag_sym(vk_context* h, cudaStream_t st, float2* W, int B, int r, int ld, size_t wstride, float* d, float* e,
                       float* tau, float2* ph) {
    const size_t vec = (size_t)(5 + TD_WARPS) * EPL * 32 * sizeof(float2);
    int nts = (int)sqrt((double)(VK_SMEM_BUDGET - vec) / sizeof(float2));
    if (nts > r) nts = r;
    if (r > 384) nts = 0;  // see launch_tridiag
    const size_t smem = vec + (size_t)nts * nts * sizeof(float2);
    VK_CUDA(h, cudaFuncSetAttribute(tridiag_sym_kernel<EPL, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tridiag_sym_kernel<EPL, RB><<<B, TD_THREADS, smem, st>>>(W, r, ld, wstride, d, e, tau, ph, nts);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}


template <int EPL, int RB, int NB>
int launch_tridiag_defer(vk_context* h, cudaStream_t st, float2* W, int B, int r, int ld, size_t wstride, float* d, float* e,
                         float* tau, float2* ph) {
    const size_t vec = (size_t)(5 + 2 * NB) * EPL * 32 * sizeof(float2);
    int nts = (int)sqrt((double)(VK_SMEM_BUDGET - vec) / sizeof(float2));
    if (nts > r) nts = r;
    if (r > 384) nts = 0;  // see launch_tridiag
    const size_t smem = vec + (size_t)nts * nts * sizeof(float2);
    VK_CUDA(h, cudaFuncSetAttribute(tridiag_defer_kernel<EPL, RB, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tridiag_defer_kernel<EPL, RB, NB><<<B, TD_THREADS, smem, st>>>(W, r, ld, wstride, d, e, tau, ph, nts);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

template <int EPL, int RPW>
int launch_formq(vk_context* h, cudaStream_t st, const float2* W, int B, int r, int ld, size_t wstride, const float* tau,
                 const float2* ph, float2* X, const int32_t* skip) {
    const size_t smem = (size_t)FQ_TJ * EPL * 32 * sizeof(float2);
    VK_CUDA(h, cudaFuncSetAttribute(formq_kernel<EPL, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((r + FQ_WARPS * RPW - 1) / (FQ_WARPS * RPW), B);
    formq_kernel<EPL, RPW><<<grid, FQ_THREADS, smem, st>>>(W, r, ld, wstride, tau, ph, X, skip);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

template <int NS>
int launch_rotapply(vk_context* h, cudaStream_t st, const float2* X, int B, int r, const EigScratch& L,
                    unsigned char* sc, float2* W, int ld, size_t wstride, int32_t* done, int32_t* sweeps,
                    const int32_t* skip) {
    constexpr int THREADS = RA_C * NS;
    const size_t smem = (size_t)THREADS * 24 + (size_t)NS * 16 * 8 + (size_t)r * RA_C * sizeof(float4);
    VK_CUDA(h, cudaFuncSetAttribute(rotapply_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((r + 2 * RA_C - 1) / (2 * RA_C), B);
    rotapply_kernel<NS><<<grid, THREADS, smem, st>>>(X, r, reinterpret_cast<const float2*>(sc + L.cs),
                                                    reinterpret_cast<const SweepRec*>(sc + L.sw),
                                                    reinterpret_cast<const int32_t*>(sc + L.meta),
                                                    reinterpret_cast<const float*>(sc + L.lam), W, ld, wstride, L.cap,
                                                    L.scap, done, sweeps, skip);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

template <int EPL, int RPW>
int launch_backtr(vk_context* h, cudaStream_t st, float2* W, int B, int r, int ld, size_t wstride, int k, const float* tau,
                  const float2* ph, const EigScratch& L, unsigned char* sc, int32_t* done, int32_t* sweeps,
                  const int32_t* kvec = nullptr, const float* lamall = nullptr) {
    const size_t smem = (size_t)FQ_TJ * EPL * 32 * sizeof(float2);
    VK_CUDA(h, cudaFuncSetAttribute(backtr_kernel<EPL, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    backtr_kernel<EPL, RPW><<<B, FQ_THREADS, smem, st>>>(W, r, ld, wstride, k, kvec, lamall, tau, ph,
                                                         reinterpret_cast<const float*>(sc + L.z),
                                                         reinterpret_cast<const float*>(sc + L.lamtop),
                                                         reinterpret_cast<const int32_t*>(sc + L.flag), done, sweeps);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

// largest fixed rank the leading-eigenpair path takes at this size (0 = none)
int topk_qr_limit(int r) {
    if (r < 4) return 0;
    const int lim = r <= 512 ? TK_MAXK : TK_MAXK / 2;  // vectors per CTA of the back-transformation
    return lim < r - 1 ? lim : r - 2;
}

}  // namespace

bool vk_eigqr_supported(int r) { return r >= 2 && r <= 1024; }

// (+ slack: the remainder split below lays two sub-batches out one after the other)
size_t vk_eigqr_scratch_bytes(int B, int r) { return eig_layout(B, r).total + (64u << 10); }

// W [B][r][ld] in/out (see the header comment); scratch: vk_eigqr_scratch_bytes(B, r) bytes of device memory.
// done_dev[b] = 1 / sweeps_dev[b] = QL iterations on success; done_dev[b] = 0 when the rotation store overflowed or
// the QL iteration did not converge (W[b] is then left tridiagonalised, i.e. unusable).
// fixed_rank > 0 (and small enough): only the leading fixed_rank vectors are produced for matrices whose leading
// eigenvalues are well separated (rows below are zero); the others take the full path.
int vk_launch_eigqr(vk_context* h, float2* W, int B, int r, int ld, void* scratch, int32_t* sweeps_dev,
                    int32_t* done_dev, int fixed_rank, double decorrelation) {
    if (B <= 0) return VK_OK;
    if (!vk_eigqr_supported(r)) return vk_fail(h, VK_EINVAL, "eig_impl=2 does not support this size");
    // Remainder split ("tail_split" 0 = on): the tridiagonalisation runs one matrix per SM, so B = q * SMs + rem matrices take
    // q + 1 waves and the last one holds only rem matrices (the MeerKAT shard: 1040 = 7 x 148 + 4, 12 % of the kernel). The
    // remainder is a sub-batch of its own on a second stream that starts when the main sub-batch has left the
    // tridiagonalisation: its few CTAs run under the main sub-batch's later stages. Same arithmetic per matrix either way.
    {
        const int nsm = h->num_sms, rem = B % nsm;
        if (h->tail_split == 0 && !h->in_split && h->stage_timing == 0 && h->tridiag_impl == 0 && vk_tridiag_symdefer_supported(r) &&
            B > nsm && rem > 0 && rem * 4 <= nsm) {
            const int B1 = B - rem;
            if (!h->tail_stream) VK_CUDA(h, cudaStreamCreateWithFlags(&h->tail_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i)
                if (!h->tail_ev[i]) VK_CUDA(h, cudaEventCreateWithFlags(&h->tail_ev[i], cudaEventDisableTiming));
            cudaStream_t main_st = h->stream;
            const int save_variant = h->tridiag_variant;
            h->in_split = true;
            // (round 2: one matrix per SM, B1 being a whole number of waves; with the L2 prefetch of the two-per-SM shape the
            //  automatic choice wins: 183.3 -> 178.7 ms per compress of the MeerKAT shard; "split_variant" = 1 restores it)
            h->tridiag_variant = h->split_variant;
            int rc = vk_launch_eigqr(h, W, B1, r, ld, scratch, sweeps_dev, done_dev, fixed_rank, decorrelation);
            if (!rc) {
                unsigned char* sc2 = static_cast<unsigned char*>(scratch) + al(eig_layout(B1, r).total);
                cudaStreamWaitEvent(h->tail_stream, h->tail_ev[0], 0);   // recorded behind the main sub-batch's tridiagonalisation
                h->stream = h->tail_stream;
                rc = vk_launch_eigqr(h, W + (size_t)B1 * r * ld, rem, r, ld, sc2, sweeps_dev + B1, done_dev + B1, fixed_rank,
                                     decorrelation);
                h->stream = main_st;
                cudaEventRecord(h->tail_ev[1], h->tail_stream);
                cudaStreamWaitEvent(main_st, h->tail_ev[1], 0);
            }
            h->tridiag_variant = save_variant;
            h->in_split = false;
            return rc;
        }
    }
    const EigScratch L = eig_layout(B, r);
    unsigned char* sc = static_cast<unsigned char*>(scratch);
    cudaStream_t st = h->stream;
    const size_t wstride = (size_t)r * ld;
    float* d = reinterpret_cast<float*>(sc + L.d);
    float* e = reinterpret_cast<float*>(sc + L.e);
    float* tau = reinterpret_cast<float*>(sc + L.tau);
    float2* ph = reinterpret_cast<float2*>(sc + L.ph);
    float2* X = reinterpret_cast<float2*>(sc + L.X);
    int rc;
    const bool dbg = h->stage_timing >= 1;  // per-kernel event times (vk_last_eig_ms; also on stderr at level 2)
    cudaEvent_t* ev = h->eig_ev;
    if (dbg) cudaEventRecord(ev[0], st);
    // lower-triangle variant where it wins: 256 < r <= 512 (32.0 vs 35.3 ms per 296 matrices at r = 512; at r <= 256 the
    // full-storage kernel is faster, 1.98 vs 2.07 ms for 112 matrices of r = 256, above 512 it would spill)
    // deferred updates (kernel 1d): 29.4 vs 30.7 ms per 256 matrices of r = 512 against the lower-triangle kernel; at r = 256
    // it is slower than kernel 1 (2.11 vs 1.98 ms per 112 matrices: both are bound by instruction issue - ncu: 1.0e9 warp
    // instructions, issue slots 54 % busy with 4 warps per scheduler, FMA pipe 25 % - not by the bytes they move), so it is
    // only taken there on request ("tridiag_impl" = 2). Also measured and dropped at r = 256: the same kernel with packed
    // fma.rn.f32x2 arithmetic and (x, x, y, y) operand vectors (6 packed FMAs per element instead of 12 scalar ones, two rows
    // per warp to stay inside 128 registers): 2.09 ms
    // "tridiag_impl": 0 = lower triangle + deferred updates (tridiag_sym.cu) where it applies, 1 = the undeferred kernels,
    // 2 = the full-storage deferred kernel 1d
    if (h->tridiag_impl == 0 && vk_tridiag_symdefer_supported(r)) rc = vk_launch_tridiag_symdefer(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    else if (r > 128 && r <= 256 && h->tridiag_impl == 2) rc = launch_tridiag_defer<8, 4, 8>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    else if (r > 384 && r <= 512 && h->tridiag_impl != 1) rc = launch_tridiag_defer<16, 2, 8>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    else if (r > 256 && r <= 512 && h->jacobi_generic != 2) rc = launch_tridiag_sym<16, 1>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    else if (r <= 64 && h->tridiag_impl == 0) rc = vk_launch_tridiag_small(h, st, W, B, r, ld, wstride, d, e, tau, ph);  // one warp per matrix
    else if (r <= 64) rc = launch_tridiag<2, 4>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    else if (r <= 128) rc = launch_tridiag<4, 4>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    else if (r <= 256) rc = launch_tridiag<8, 4>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    else if (r <= 512) rc = launch_tridiag<16, 2>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    else rc = launch_tridiag<32, 1>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    if (rc) return rc;
    if (h->in_split && st != h->tail_stream) cudaEventRecord(h->tail_ev[0], st);   // the remainder sub-batch may start
    if (dbg) cudaEventRecord(ev[1], st);
    const int32_t* skip = nullptr;
    // sweeps in flight in the rotation application (and in the level assignment of the QL kernel): measured 1.85 vs 2.31 ms
    // at r = 256 with 64 vs 128 slots, 40.9 vs 35.8 ms per 296 matrices at r = 512
    const int nslots = r <= 256 ? 64 : 128;
    if (h->topk != 1 && fixed_rank > 0 && fixed_rank <= topk_qr_limit(r)) {
        const int k = fixed_rank;
        float* lamtop = reinterpret_cast<float*>(sc + L.lamtop);
        int32_t* flag = reinterpret_cast<int32_t*>(sc + L.flag);
        if (h->bisect_impl != 1 && B >= 4 * h->num_sms && k + 1 <= 16 && r <= 128) {
            // many small problems: several matrices per warp, one lane per eigenvalue
            const int mpc = bp_group(k + 1) * BP_WARPS;
            bisect_packed_kernel<<<(B + mpc - 1) / mpc, 32 * BP_WARPS, (size_t)mpc * 2 * r * 4, st>>>(B, r, k + 1, d, e, lamtop,
                                                                                                   flag, TK_GAP);
        } else {
            bisect_kernel<<<B, (8 * (k + 1) + 31) / 32 * 32, (size_t)2 * r * 4, st>>>(r, k + 1, d, e, lamtop, flag,
                                                                                    TK_GAP);
        }
        VK_LAUNCH_CHECK(h);
        twisted_kernel<<<B, 32, 0, st>>>(r, k, nullptr, d, e, lamtop, flag, reinterpret_cast<float*>(sc + L.z),
                                         reinterpret_cast<float*>(sc + L.dm));
        VK_LAUNCH_CHECK(h);
        // vectors per warp: as few as the eight warps of the CTA allow (k <= 8: one each)
#define VK_BACKTR(EPL)                                                                                                  \
    (k <= 8    ? launch_backtr<EPL, 1>(h, st, W, B, r, ld, wstride, k, tau, ph, L, sc, done_dev, sweeps_dev)              \
     : k <= 16 ? launch_backtr<EPL, 2>(h, st, W, B, r, ld, wstride, k, tau, ph, L, sc, done_dev, sweeps_dev)              \
               : launch_backtr<EPL, 4>(h, st, W, B, r, ld, wstride, k, tau, ph, L, sc, done_dev, sweeps_dev))
        // (r <= 64, 8320 matrices, k = 8: four vectors per warp instead of one lose, 0.90 against 0.69 ms for the three kernels)
        if (r <= 64) rc = VK_BACKTR(2);
        else if (r <= 128) rc = VK_BACKTR(4);
        else if (r <= 256) rc = VK_BACKTR(8);
        else if (r <= 512) rc = VK_BACKTR(16);
        else rc = k <= 8 ? launch_backtr<32, 1>(h, st, W, B, r, ld, wstride, k, tau, ph, L, sc, done_dev, sweeps_dev)
                         : launch_backtr<32, 2>(h, st, W, B, r, ld, wstride, k, tau, ph, L, sc, done_dev, sweeps_dev);
#undef VK_BACKTR
        if (rc) return rc;
        skip = flag;  // the full path below takes what is left (flag 0)
    } else if (h->topk == 2 && fixed_rank <= 0 && decorrelation > 0.0 && decorrelation < 1.0 && r >= 8) {
        // (opt-in: pays when most matrices of the batch end up with a small rank - high SNR data; one matrix that needs the
        // full path brings the latency of the QL kernel back, and the extra bisection costs 1-5 % when none qualifies)
        // energy rule: all eigenvalues by bisection; matrices whose rank (+2) is within reach take the leading-pair
        // path with their own k and leave the remaining eigenvalues as dummy rows for the selection stage
        float* lamtop = reinterpret_cast<float*>(sc + L.lamtop);
        float* lamall = reinterpret_cast<float*>(sc + L.lam);
        int32_t* flag = reinterpret_cast<int32_t*>(sc + L.flag);
        int32_t* kvec = reinterpret_cast<int32_t*>(sc + L.kvec);
        const int klimit = topk_qr_limit(r);
        const int nthr = r >= 1024 ? 1024 : (r + 31) / 32 * 32;
        bisect_all_kernel<<<B, nthr, (size_t)3 * r * 4, st>>>(r, decorrelation * decorrelation, klimit, d, e, lamall, lamtop,
                                                             kvec, flag);
        VK_LAUNCH_CHECK(h);
        twisted_kernel<<<B, 32, 0, st>>>(r, 0, kvec, d, e, lamtop, flag, reinterpret_cast<float*>(sc + L.z),
                                         reinterpret_cast<float*>(sc + L.dm));
        VK_LAUNCH_CHECK(h);
        if (r <= 64) rc = launch_backtr<2, 4>(h, st, W, B, r, ld, wstride, 0, tau, ph, L, sc, done_dev, sweeps_dev, kvec, lamall);
        else if (r <= 128) rc = launch_backtr<4, 4>(h, st, W, B, r, ld, wstride, 0, tau, ph, L, sc, done_dev, sweeps_dev, kvec, lamall);
        else if (r <= 256) rc = launch_backtr<8, 4>(h, st, W, B, r, ld, wstride, 0, tau, ph, L, sc, done_dev, sweeps_dev, kvec, lamall);
        else if (r <= 512) rc = launch_backtr<16, 4>(h, st, W, B, r, ld, wstride, 0, tau, ph, L, sc, done_dev, sweeps_dev, kvec, lamall);
        else rc = launch_backtr<32, 2>(h, st, W, B, r, ld, wstride, 0, tau, ph, L, sc, done_dev, sweeps_dev, kvec, lamall);
        if (rc) return rc;
        skip = flag;
    }
    // full spectrum: eigenvectors of T by twisted factorisation + Newton-Schulz + GEMM back-transformation (see above);
    // the QL kernels below then only see the matrices that path gave up (flag 0), normally none
    const bool mrrr = skip == nullptr && h->eigvec_impl != 1 && r >= 128 && (r % 16) == 0 && ld == r;
    if (mrrr) {
        float* lam = reinterpret_cast<float*>(sc + L.lam);
        int32_t* flag = reinterpret_cast<int32_t*>(sc + L.flag);
        unsigned* emax = reinterpret_cast<unsigned*>(sc + L.emax);
        float* dp = reinterpret_cast<float*>(sc + L.cs);
        float* dm = dp + (size_t)B * r * r;
        float2* Zt = reinterpret_cast<float2*>(sc + L.zt);
        float2* Zt2 = reinterpret_cast<float2*>(sc + L.zt2);
        fill_i32_kernel<<<(B + 127) / 128, 128, 0, st>>>(flag, B, 1);
        VK_LAUNCH_CHECK(h);
        VK_CUDA(h, cudaMemsetAsync(emax, 0, (size_t)B * 4, st));
        const dim3 egrid((r + EV_THREADS - 1) / EV_THREADS, B);
        bisect_full_kernel<<<egrid, EV_THREADS, (size_t)2 * r * 4, st>>>(r, d, e, lam);
        VK_LAUNCH_CHECK(h);
        twisted_full_kernel<<<egrid, EV_THREADS, ((size_t)2 * r + (EV_THREADS / 32) * 32 * 33) * 4, st>>>(r, d, e, lam, dp, dm,
                                                                                                     Zt, flag);
        VK_LAUNCH_CHECK(h);
        if (dbg) cudaEventRecord(ev[2], st);
        if (dbg) cudaEventRecord(ev[3], st);
        // reflectors of every matrix -> X = (Q D)^T; W is free from here on
        if (r <= 128) rc = launch_formq<4, 4>(h, st, W, B, r, ld, wstride, tau, ph, X, nullptr);
        else if (r <= 256) rc = launch_formq<8, 4>(h, st, W, B, r, ld, wstride, tau, ph, X, nullptr);
        else if (r <= 512) rc = launch_formq<16, 2>(h, st, W, B, r, ld, wstride, tau, ph, X, nullptr);
        else rc = launch_formq<32, 1>(h, st, W, B, r, ld, wstride, tau, ph, X, nullptr);
        if (rc) return rc;
        if (dbg) cudaEventRecord(ev[4], st);
        // W <- Zt Zt^H (= Zt Zt^T: zero imaginary parts), then P = I - E/2 in place and the verdict per matrix
        if ((rc = vk_launch_gram_tc(h, Zt, B, r, r, W))) return rc;
        {
            int gx = (int)(((size_t)r * r + 255) / 256);
            if (gx > 64) gx = 64;
            nsprep_kernel<<<dim3(gx, B), 256, 0, st>>>(W, r, flag, emax);
            VK_LAUNCH_CHECK(h);
            nsflag_kernel<<<(B + 127) / 128, 128, 0, st>>>(emax, B, flag, done_dev, sweeps_dev);
            VK_LAUNCH_CHECK(h);
        }
        // Zt2 = P Zt ; W[t][:] = lambda_t conj((Zt2 X)[t][:])
        if ((rc = vk_launch_cgemm_tc_plain(h, W, Zt, Zt2, nullptr, B, r, r, r))) return rc;
        if ((rc = vk_launch_cgemm_tc_plain(h, Zt2, X, W, lam, B, r, r, r))) return rc;
        skip = flag;
        // the matrices that were given up: QL on T, rotations applied to X (already formed), W rewritten
        {
            const size_t smem = (size_t)2 * (r + 2) * 4 + (size_t)RA_NS_MAX * 4;
            VK_CUDA(h, cudaFuncSetAttribute(tql_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            tql_kernel<<<B, 32, smem, st>>>(r, d, e, lam, reinterpret_cast<float2*>(sc + L.cs),
                                            reinterpret_cast<SweepRec*>(sc + L.sw), reinterpret_cast<int32_t*>(sc + L.meta),
                                            L.cap, L.scap, L.lcap, skip, h->ql_maxit > 0 ? h->ql_maxit : QL_MAXIT, nslots);
            VK_LAUNCH_CHECK(h);
        }
        rc = nslots == 64 ? launch_rotapply<64>(h, st, X, B, r, L, sc, W, ld, wstride, done_dev, sweeps_dev, skip)
                          : launch_rotapply<128>(h, st, X, B, r, L, sc, W, ld, wstride, done_dev, sweeps_dev, skip);
        if (dbg) {
            cudaEventRecord(ev[5], st);
            cudaEventSynchronize(ev[5]);
            float t[5];
            for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
            for (int i = 0; i < 5; ++i) h->eig_ms[i] += t[i];
            if (h->stage_timing >= 2)
                fprintf(stderr, "[eigqr B=%d r=%d full] tridiag %.3f  bisect+twisted %.3f  formq %.3f  gram+NS+back %.3f ms\n", B, r,
                        t[0], t[1], t[3], t[4]);
        }
        return rc;
    }
    if (dbg) cudaEventRecord(ev[2], st);
    // the scalar QL iteration is latency bound (one lane per matrix): it runs alone - sharing the SMs with another
    // kernel slows its dependent chain by more than the overlap wins (measured: 10.3 ms serial, 18.6 ms overlapped)
    {
        const size_t smem = (size_t)2 * (r + 2) * 4 + (size_t)RA_NS_MAX * 4;
        VK_CUDA(h, cudaFuncSetAttribute(tql_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tql_kernel<<<B, 32, smem, st>>>(r, d, e, reinterpret_cast<float*>(sc + L.lam),
                                        reinterpret_cast<float2*>(sc + L.cs), reinterpret_cast<SweepRec*>(sc + L.sw),
                                        reinterpret_cast<int32_t*>(sc + L.meta), L.cap, L.scap, L.lcap, skip,
                                        h->ql_maxit > 0 ? h->ql_maxit : QL_MAXIT, nslots);
        VK_LAUNCH_CHECK(h);
    }
    if (dbg) cudaEventRecord(ev[3], st);
    if (r <= 64) rc = launch_formq<2, 4>(h, st, W, B, r, ld, wstride, tau, ph, X, skip);
    else if (r <= 128) rc = launch_formq<4, 4>(h, st, W, B, r, ld, wstride, tau, ph, X, skip);
    else if (r <= 256) rc = launch_formq<8, 4>(h, st, W, B, r, ld, wstride, tau, ph, X, skip);
    else if (r <= 512) rc = launch_formq<16, 2>(h, st, W, B, r, ld, wstride, tau, ph, X, skip);
    else rc = launch_formq<32, 1>(h, st, W, B, r, ld, wstride, tau, ph, X, skip);
    if (rc) return rc;
    if (dbg) cudaEventRecord(ev[4], st);
    rc = nslots == 64 ? launch_rotapply<64>(h, st, X, B, r, L, sc, W, ld, wstride, done_dev, sweeps_dev, skip)
                      : launch_rotapply<128>(h, st, X, B, r, L, sc, W, ld, wstride, done_dev, sweeps_dev, skip);
    if (dbg) {
        cudaEventRecord(ev[5], st);
        cudaEventSynchronize(ev[5]);
        float t[5];
        for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
        for (int i = 0; i < 5; ++i) h->eig_ms[i] += t[i];
        if (h->stage_timing >= 2)
            fprintf(stderr, "[eigqr B=%d r=%d k=%d] tridiag %.3f  leading-pairs %.3f  tql %.3f  formq %.3f  rotapply %.3f ms\n",
                    B, r, fixed_rank, t[0], t[1], t[2], t[3], t[4]);
    }
    return rc;
}
