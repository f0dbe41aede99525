// SIMT stages of the hot path (sm_100a): Gram product for shapes the tcgen05 kernel does not take, Gram
// normalisation, small-matrix packing, singular-value selection / rank truncation, factor formation, rank-k
// reconstruction, and the benchmark generator. Reference call sites are cited at each launcher.
#include <cooperative_groups.h>

#include "cgemm.cuh"

namespace {

// =================================================================================================================
// Gram product (SIMT)                                    reference: implicit in np.linalg.svd, compress_ms.py:350
// =================================================================================================================
struct GramOp0 {  // W[i][t] = sum_v conj(A[i][v]) * A[t][v]          (r = m)
    const float2* A;
    float2* W;
    int m, n;
    static constexpr bool A_K_CONTIG = true, B_K_CONTIG = true;
    static constexpr int REDUCE = 0;
    __device__ int M(int) const { return m; }
    __device__ int N(int) const { return m; }
    __device__ int K(int) const { return n; }
    __host__ __device__ int Mfill() const { return m; }
    __host__ __device__ int Nfill() const { return m; }
    __device__ float2 loadA(int b, int i, int k) const {
        const float2 v = A[((size_t)b * m + i) * n + k];
        return make_float2(v.x, -v.y);
    }
    __device__ float2 loadB(int b, int k, int j) const { return A[((size_t)b * m + j) * n + k]; }
    __device__ void store(int b, int i, int j, float2 v) const { W[((size_t)b * m + i) * m + j] = v; }
    __device__ void reduce_add(int, int, float) const {}
};
struct GramOp1 {  // W[i][j] = sum_t A[t][i] * conj(A[t][j])          (r = n)
    const float2* A;
    float2* W;
    int m, n;
    static constexpr bool A_K_CONTIG = false, B_K_CONTIG = false;
    static constexpr int REDUCE = 0;
    __device__ int M(int) const { return n; }
    __device__ int N(int) const { return n; }
    __device__ int K(int) const { return m; }
    __host__ __device__ int Mfill() const { return n; }
    __host__ __device__ int Nfill() const { return n; }
    __device__ float2 loadA(int b, int i, int k) const { return A[((size_t)b * m + k) * n + i]; }
    __device__ float2 loadB(int b, int k, int j) const {
        const float2 v = A[((size_t)b * m + k) * n + j];
        return make_float2(v.x, -v.y);
    }
    __device__ void store(int b, int i, int j, float2 v) const { W[((size_t)b * n + i) * n + j] = v; }
    __device__ void reduce_add(int, int, float) const {}
};

// scale W[b] so that trace == r; gscale[b] = trace / r. One CTA per matrix.
// A trace that is not finite, or outside [1e-30, 1e30] * r, means the squares of the input left the float32 range (|a| beyond
// ~1e19 or below ~1e-19; LAPACK scales such input): the matrix is replaced by the identity so that the stages behind stay
// harmless, bad[b] is set, and the driver does it again without a Gram product (api.cu). NaN / Inf input ends up there
// too and is reported by that path.
__global__ void __launch_bounds__(1024) gram_normalise_kernel(float2* __restrict__ W, int r, float* __restrict__ gscale,
                                                             int32_t* __restrict__ nonfinite, int32_t* __restrict__ bad,
                                                             int32_t* __restrict__ nbad) {
    __shared__ float part[32];
    __shared__ float sc;
    const int b = blockIdx.x;
    float2* Wb = W + (size_t)b * r * r;
    float s = 0.f;
    for (int i = threadIdx.x; i < r; i += blockDim.x) s += Wb[(size_t)i * r + i].x;
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tr = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tr += part[w];
        float g = tr / (float)r;
        bool isbad = false;
        // (a trace of exactly zero is either a zero matrix or input whose squares vanished: the second pass tells)
        if (!isfinite(g) || fabsf(g) < 1e-30f || fabsf(g) > 1e30f) {
            if (bad) {
                isbad = true;
                bad[b] = 1;
                atomicAdd(nbad, 1);
            } else if (!isfinite(g)) {
                atomicOr(nonfinite, 1);
            }
            g = 1.f;
        }
        if (!(g > 0.f)) g = 1.f;  // all-zero matrix: leave it alone
        gscale[b] = g;
        sc = isbad ? -1.f : 1.f / g;
    }
    __syncthreads();
    const float f = sc;
    const size_t tot = (size_t)r * r;
    if (f < 0.f) {   // out-of-range matrix: identity
        for (size_t e = threadIdx.x; e < tot; e += blockDim.x)
            Wb[e] = make_float2((e / r == e % r) ? 1.f : 0.f, 0.f);
        return;
    }
    if ((tot & 1) == 0) {  // two complex numbers per access (the matrix base is 16-byte aligned when r*r is even)
        float4* W4 = reinterpret_cast<float4*>(Wb);
        const size_t n4 = tot >> 1;
#pragma unroll 4
        for (size_t e = threadIdx.x; e < n4; e += blockDim.x) {
            float4 v = W4[e];
            v.x *= f, v.y *= f, v.z *= f, v.w *= f;
            W4[e] = v;
        }
    } else {
        for (size_t e = threadIdx.x; e < tot; e += blockDim.x) {
            float2 v = Wb[e];
            v.x *= f;
            v.y *= f;
            Wb[e] = v;
        }
    }
}

// Gram product AND trace normalisation of a matrix with min(m, n) <= 64 in one CTA (BASELINE configs[3]: 8320 matrices of
// 64 x 64). The general SIMT GEMM above spends its time in guarded element-wise operand loads and two barriers per 16
// contraction steps, and the normalisation is a second pass over W. Here the r vectors (rows of A when m <= n, else columns)
// sit in shared memory 64 contraction steps at a time (row stride 66: 128-bit reads of two steps, conflict-free over the
// sixteen row classes), thread (ib, jb), ib <= jb, owns the 4 x 4 entries (ib + 16 p, jb + 16 q) - only the upper triangle of
// blocks is computed, the mirror image is written as conjugates - and the finished matrix is staged in the same shared
// memory for the trace and a coalesced, already scaled store. Same contract as vk_launch_gram_simt followed by
// vk_launch_gram_normalise (reference: implicit in np.linalg.svd, compress_ms.py:350).
template <bool SIDE1>
__global__ void __launch_bounds__(160) gram_small_kernel(const float2* __restrict__ A, int m, int n, float2* __restrict__ W,
                                                        float* __restrict__ gscale, int32_t* __restrict__ nonfinite,
                                                        int32_t* __restrict__ bad, int32_t* __restrict__ nbad) {
    constexpr int S = 66, NT = 160;
    __shared__ __align__(16) float2 As[64 * S];
    __shared__ float s_tr[2];
    __shared__ float s_f;
    const int r = SIDE1 ? n : m, K = SIDE1 ? m : n;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float2* Ab = A + (size_t)b * m * n;
    int ib = 0, jb = 0;
    {
        int rem = tid;
        while (ib < 16 && rem >= 16 - ib) rem -= 16 - ib, ++ib;
        jb = ib + rem;
    }
    const bool worker = tid < 136 && ib < r && jb < r;   // (classes beyond r hold zero rows only)
    float2 acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = make_float2(0.f, 0.f);
    // (rows of an even length from a 16-byte aligned base: two contraction steps per load)
    const bool vec2 = !SIDE1 && (n & 1) == 0 && (reinterpret_cast<uintptr_t>(Ab) & 15) == 0;
    for (int k0 = 0; k0 < K; k0 += 64) {
        if (vec2) {
            for (int idx = tid; idx < 64 * 32; idx += NT) {
                const int i = idx >> 5, kk = (idx & 31) * 2;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < m && k0 + kk < n) v = *reinterpret_cast<const float4*>(Ab + (size_t)i * n + k0 + kk);
                *reinterpret_cast<float4*>(As + i * S + kk) = v;
            }
        } else
        for (int idx = tid; idx < 64 * 64; idx += NT) {
            float2 v = make_float2(0.f, 0.f);
            int i, kk;
            if (!SIDE1) {
                i = idx >> 6, kk = idx & 63;
                if (i < m && k0 + kk < n) v = Ab[(size_t)i * n + k0 + kk];
            } else {
                kk = idx >> 6, i = idx & 63;
                if (k0 + kk < m && i < n) v = Ab[(size_t)(k0 + kk) * n + i];
            }
            As[i * S + kk] = v;
        }
        __syncthreads();
        if (worker) {
            const float4* ap = reinterpret_cast<const float4*>(As + ib * S);
            const float4* bp = reinterpret_cast<const float4*>(As + jb * S);
#pragma unroll 4
            for (int k2 = 0; k2 < 32; ++k2) {
                float4 a[4], c[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) a[p] = ap[p * 8 * S + k2], c[p] = bp[p * 8 * S + k2];   // (16 rows = 8 S float4)
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        // conj(a) * c for both contraction steps
                        acc[p][q].x = fmaf(a[p].x, c[q].x, fmaf(a[p].y, c[q].y, fmaf(a[p].z, c[q].z, fmaf(a[p].w, c[q].w, acc[p][q].x))));
                        acc[p][q].y = fmaf(a[p].x, c[q].y, fmaf(-a[p].y, c[q].x, fmaf(a[p].z, c[q].w, fmaf(-a[p].w, c[q].z, acc[p][q].y))));
                    }
            }
        }
        __syncthreads();
    }
    // the finished matrix, full storage, in shared memory: G[i][t] (side 1: the conjugate)
    if (worker) {
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = ib + 16 * p, t = jb + 16 * q;
                if (ib == jb && p > q) continue;         // (a diagonal class holds both (i, t) and (t, i): keep one)
                float2 v = acc[p][q];
                if (SIDE1) v.y = -v.y;
                if (i == t) v.y = 0.f;
                As[i * S + t] = v;
                if (i != t) As[t * S + i] = make_float2(v.x, -v.y);   // exactly Hermitian
            }
    }
    __syncthreads();
    if (tid < 64) {
        float d = tid < r ? As[tid * S + tid].x : 0.f;
        d = warp_sum(d);
        if ((tid & 31) == 0) s_tr[tid >> 5] = d;
    }
    __syncthreads();
    if (tid == 0) {
        // (as gram_normalise_kernel)
        const float tr = s_tr[0] + s_tr[1];
        float g = tr / (float)r;
        bool isbad = false;
        if (!isfinite(g) || fabsf(g) < 1e-30f || fabsf(g) > 1e30f) {
            if (bad) {
                isbad = true;
                bad[b] = 1;
                atomicAdd(nbad, 1);
            } else if (!isfinite(g)) {
                atomicOr(nonfinite, 1);
            }
            g = 1.f;
        }
        if (!(g > 0.f)) g = 1.f;
        gscale[b] = g;
        s_f = isbad ? -1.f : 1.f / g;
    }
    __syncthreads();
    const float f = s_f;
    float2* Wb = W + (size_t)b * r * r;
    if ((r & 1) == 0 && f >= 0.f && (reinterpret_cast<uintptr_t>(Wb) & 15) == 0) {
        const int h2 = r >> 1;
        for (int idx = tid; idx < r * h2; idx += NT) {
            const int i = idx / h2, t = (idx - i * h2) * 2;
            float4 v = *reinterpret_cast<const float4*>(As + i * S + t);
            v.x *= f, v.y *= f, v.z *= f, v.w *= f;
            *reinterpret_cast<float4*>(Wb + (size_t)i * r + t) = v;
        }
        return;
    }
    for (int idx = tid; idx < r * r; idx += NT) {
        const int i = idx / r, t = idx - i * r;
        float2 v = As[i * S + t];
        if (f < 0.f) v = make_float2(i == t ? 1.f : 0.f, 0.f);
        else v.x *= f, v.y *= f;
        Wb[idx] = v;
    }
}

// Small path: vectors = rows of A (m <= n) or columns of A (m > n), normalised to unit rms norm, followed by e_i.
__global__ void __launch_bounds__(256) pack_small_kernel(const float2* __restrict__ A, int m, int n,
                                                         float2* __restrict__ W, int ld, float* __restrict__ gscale,
                                                         int32_t* __restrict__ nonfinite) {
    __shared__ float part[8];
    __shared__ float sc;
    const int b = blockIdx.x;
    const float2* Ab = A + (size_t)b * m * n;
    const int r = m <= n ? m : n;
    const int L = m <= n ? n : m;
    // largest magnitude first: the squares are summed on data scaled by a power of two into [1, 2) at the top, so inputs
    // of any float32 magnitude neither overflow nor vanish here (LAPACK's cgesdd scales the same way)
    __shared__ int s_exp;
    float amax = 0.f;
    for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
        const float2 v = Ab[e];
        amax = fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y)));
        if (!isfinite(v.x) || !isfinite(v.y)) amax = __int_as_float(0x7f800000);
    }
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, part[w]);
        int ex = 0;
        if (!isfinite(mx)) atomicOr(nonfinite, 1);
        else if (mx > 0.f) ex = ilogbf(mx);
        s_exp = ex;
    }
    __syncthreads();
    const int ex = s_exp;
    float s = 0.f;
    for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
        const float2 v = Ab[e];
        const float x = scalbnf(v.x, -ex), y = scalbnf(v.y, -ex);
        s = fmaf(x, x, fmaf(y, y, s));
    }
    s = warp_sum(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
        float g = sqrtf(tot / (float)r);
        if (!isfinite(g)) g = 1.f;
        if (!(g > 0.f)) g = 1.f;
        gscale[b] = scalbnf(g, ex);
        sc = 1.f / g;
    }
    __syncthreads();
    const float f = sc;
    float2* Wb = W + (size_t)b * r * ld;
    if (m <= n) {
        for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
            const int i = e / n, x = e - i * n;
            float2 v = Ab[e];
            v.x = scalbnf(v.x, -ex) * f;
            v.y = scalbnf(v.y, -ex) * f;
            Wb[(size_t)i * ld + x] = v;
        }
    } else {
        // vector j = column j of A; read A coalesced, scattered (small) writes
        for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
            const int t = e / n, j = e - t * n;
            float2 v = Ab[e];
            v.x = scalbnf(v.x, -ex) * f;
            v.y = scalbnf(v.y, -ex) * f;
            Wb[(size_t)j * ld + t] = v;
        }
    }
    for (int e = threadIdx.x; e < r * r; e += blockDim.x) {
        const int i = e / r, j = e - i * r;
        Wb[(size_t)i * ld + L + j] = make_float2(i == j ? 1.f : 0.f, 0.f);
    }
}

// =================================================================================================================
// Selection: norms -> sort -> sigma -> rank              reference: compress_ms.py:295-319 (energy rule, float32),
//                                                                   compress_ms.py:352-361 (precedence, slicing)
// =================================================================================================================
// numpy's float32 add-reduce (np.sum of a contiguous array): pairwise summation with an 8-way unrolled leaf of at most
// 128 elements (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum). Restated so that the energy rule sees
// the same float32 total the reference sees — at decorrelation = 1.0 the reference's answer hinges on whether the
// sequential cumsum reaches this pairwise total (it returns rank 1 when it does not).
__device__ float np_pairwise_sum_sq(const float* a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, __fmul_rn(a[i], a[i]));
        return res;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = __fmul_rn(a[j], a[j]);
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], __fmul_rn(a[i + j], a[i + j]));
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, __fmul_rn(a[i], a[i]));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum_sq(a, n2), np_pairwise_sum_sq(a + n2, n - n2));
}

__device__ int energy_rank(const float* sig, int r, double decorrelation) {
    // total = np.sum(S**2) ; threshold = float32(dec**2) * total ; cumulative = np.cumsum(S**2) ;
    // n = argmax(cumulative >= threshold) + 1          (reference compress_ms.py:311-315, all float32)
    const float total = np_pairwise_sum_sq(sig, r);
    const float thr = __fmul_rn((float)(decorrelation * decorrelation), total);
    float cum = 0.f;
    for (int c = 0; c < r; ++c) {
        cum = __fadd_rn(cum, __fmul_rn(sig[c], sig[c]));  // no FMA contraction: numpy squares, then adds
        if (cum >= thr) return c + 1;
    }
    return 1;  // argmax of an all-False array is 0
}

__global__ void __launch_bounds__(256)
select_kernel(const float2* __restrict__ W, int r, int ldot, int ld, const float* __restrict__ gscale, int mode_gram,
              int fixed_rank, double decorrelation, int kmax, int32_t* __restrict__ perm, float* __restrict__ inv,
              float* __restrict__ S, int32_t* __restrict__ ranks, float* __restrict__ stats,
              const int32_t* __restrict__ sweeps, const int32_t* __restrict__ done) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int P = 1;
    while (P < r) P <<= 1;
    float* key = reinterpret_cast<float*>(smem_raw);  // [P]
    int* idx = reinterpret_cast<int*>(key + P);       // [P]
    float* sig = reinterpret_cast<float*>(idx + P);   // [P]
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const float2* Wb = W + (size_t)b * r * ld;
    for (int i = warp; i < P; i += nwarps) {
        float s = -1.f;
        if (i < r) {
            const float2* v = Wb + (size_t)i * ld;
            float acc = 0.f;
            for (int t = lane; t < ldot; t += 32) {
                const float2 x = v[t];
                acc = fmaf(x.x, x.x, fmaf(x.y, x.y, acc));
            }
            acc = warp_sum(acc);
            s = sqrtf(acc);
            if (!(s >= 0.f)) s = 0.f;  // NaN -> 0 (flagged elsewhere)
        }
        if (lane == 0) {
            key[i] = s;
            idx[i] = i;
        }
    }
    __syncthreads();
    // bitonic sort, descending by key (ties: lower index first)
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < P; t += blockDim.x) {
                const int p = t ^ j;
                if (p > t) {
                    const bool desc = ((t & k) == 0);
                    const float kt = key[t], kp = key[p];
                    const int it = idx[t], ip = idx[p];
                    const bool t_before_p = (kt > kp) || (kt == kp && it < ip);
                    if (desc ? !t_before_p : t_before_p) {
                        key[t] = kp;
                        key[p] = kt;
                        idx[t] = ip;
                        idx[p] = it;
                    }
                }
            }
            __syncthreads();
        }
    }
    const float g = gscale[b];
    for (int c = threadIdx.x; c < r; c += blockDim.x) {
        const float nv = key[c];
        sig[c] = (mode_gram == 1) ? sqrtf(nv * g) : nv * g;
    }
    __syncthreads();
    __shared__ int k_sh;
    if (threadIdx.x == 0) {
        int k;
        if (fixed_rank > 0)
            k = fixed_rank < r ? fixed_rank : r;
        else if (decorrelation > 0.0)
            k = energy_rank(sig, r, decorrelation);
        else
            k = r;
        if (k > kmax) k = kmax;
        k_sh = k;
        ranks[b] = k;
        double tot = 0.0, kept = 0.0;
        for (int c = 0; c < r; ++c) {
            const double e = (double)sig[c] * (double)sig[c];
            tot += e;
            if (c < k) kept += e;
        }
        // Gram path: ||A||_F^2 is the trace of the Gram matrix (exact even when only the leading vectors were computed)
        stats[4 * b + 0] = (mode_gram == 1) ? g * (float)r : (float)tot;
        stats[4 * b + 1] = (float)kept;
        stats[4 * b + 2] = (float)sweeps[b];
        stats[4 * b + 3] = (float)done[b];
    }
    __syncthreads();
    const int k = k_sh;
    for (int c = threadIdx.x; c < r; c += blockDim.x) {
        perm[(size_t)b * r + c] = idx[c];
        const float nv = key[c];
        inv[(size_t)b * r + c] = nv > 0.f ? 1.f / nv : 0.f;
    }
    for (int c = threadIdx.x; c < kmax; c += blockDim.x) S[(size_t)b * kmax + c] = (c < k) ? sig[c] : 0.f;
}

__global__ void pack_info_kernel(const int32_t* __restrict__ sweeps, const int32_t* __restrict__ done, int B,
                                 int32_t* __restrict__ info) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        info[2 * b] = sweeps[b];
        info[2 * b + 1] = done[b];
    }
}

__global__ void find_n_kernel(const float* __restrict__ S, int B, int r, double decorrelation,
                              int32_t* __restrict__ ranks) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) ranks[b] = energy_rank(S + (size_t)b * r, r, decorrelation);
}

// =================================================================================================================
// Factor formation                                        reference: U[:, :n], S[:n], Vt[:n, :] compress_ms.py:359-361
// =================================================================================================================
// dst[b][c][x] = f(W[b][perm[c]][off + x]) * (scaled ? inv[c] : 1), zero for c >= rank.   x < len, c < kmax
__global__ void __launch_bounds__(256)
rows_from_vectors_kernel(const float2* __restrict__ W, int r, int ld, int off, int len, int kmax,
                         const int32_t* __restrict__ perm, const float* __restrict__ inv,
                         const int32_t* __restrict__ ranks, int conj, int scaled, float2* __restrict__ dst) {
    const int b = blockIdx.y, c = blockIdx.x;
    float2* d = dst + ((size_t)b * kmax + c) * len;
    if (c >= ranks[b]) {
        for (int x = threadIdx.x; x < len; x += blockDim.x) d[x] = make_float2(0.f, 0.f);
        return;
    }
    const int p = perm[(size_t)b * r + c];
    const float f = scaled ? inv[(size_t)b * r + c] : 1.f;
    const float fi = conj ? -f : f;
    const float2* src = W + ((size_t)b * r + p) * ld + off;
    for (int x = threadIdx.x; x < len; x += blockDim.x) {
        const float2 v = src[x];
        d[x] = make_float2(v.x * f, v.y * fi);
    }
}
// dst[b][t][c] = f(W[b][perm[c]][off + t]) * (scaled ? inv[c] : 1), zero for c >= rank.   t < len, c < kmax
__global__ void __launch_bounds__(256)
cols_from_vectors_kernel(const float2* __restrict__ W, int r, int ld, int off, int len, int kmax,
                         const int32_t* __restrict__ perm, const float* __restrict__ inv,
                         const int32_t* __restrict__ ranks, int conj, int scaled, float2* __restrict__ dst) {
    __shared__ float2 tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 warps
    const int k = ranks[b];
    for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc, t = t0 + tx;
        float2 v = make_float2(0.f, 0.f);
        if (c < k && t < len) {
            const int p = perm[(size_t)b * r + c];
            const float f = scaled ? inv[(size_t)b * r + c] : 1.f;
            const float2 s = W[((size_t)b * r + p) * ld + off + t];
            v = make_float2(s.x * f, conj ? -s.y * f : s.y * f);
        }
        tile[cc][tx] = v;
    }
    __syncthreads();
    for (int tt = ty; tt < 32; tt += 8) {
        const int t = t0 + tt, c = c0 + tx;
        if (t < len && c < kmax) dst[((size_t)b * len + t) * kmax + c] = tile[tx][tt];
    }
}

struct FormVOp {  // Vt[c][v] = sum_t conj(W[perm c][t]) inv[c] * A[t][v]            (wide, r = m)
    const float2* A;
    const float2* W;
    const int32_t* perm;
    const float* inv;
    const int32_t* ranks;
    float2* Vt;
    float* norm2;
    int m, n, kmax;
    static constexpr bool A_K_CONTIG = true, B_K_CONTIG = false;
    static constexpr int REDUCE = 1;
    __device__ int M(int b) const { return ranks[b]; }
    __device__ int N(int) const { return n; }
    __device__ int K(int) const { return m; }
    __host__ __device__ int Mfill() const { return kmax; }
    __host__ __device__ int Nfill() const { return n; }
    __device__ float2 loadA(int b, int c, int t) const {
        const int p = perm[(size_t)b * m + c];
        const float f = inv[(size_t)b * m + c];
        const float2 v = W[((size_t)b * m + p) * m + t];
        return make_float2(v.x * f, -v.y * f);
    }
    __device__ float2 loadB(int b, int t, int j) const { return A[((size_t)b * m + t) * n + j]; }
    __device__ void store(int b, int c, int j, float2 v) const { Vt[((size_t)b * kmax + c) * n + j] = v; }
    __device__ void reduce_add(int b, int c, float v) const { atomicAdd(&norm2[(size_t)b * kmax + c], v); }
};
// V-formation for small k (north_star item (c), wide matrices): Vt[c][v] = sum_t conj(u_c[t]) A[t][v] with k <= KC <= 8.
// HBM-bound by the single read of A: a thread owns two adjacent channels, streams its float4 of every row of A
// (128-bit loads, 8 rows in flight), and keeps KC x 2 complex accumulators in registers; the coefficients
// conj(W[perm c][t]) * inv[c] sit in shared memory pre-expanded as (xr, xr, -xi, xi) so a complex MAC is two FFMA2.
// The epilogue writes the unnormalised rows and accumulates their squared norms (the refined singular values).
template <int KC>
__global__ void __launch_bounds__(128)
formv_smallk_kernel(const float2* __restrict__ A, const float2* __restrict__ W, const int32_t* __restrict__ perm,
                    const float* __restrict__ inv, const int32_t* __restrict__ ranks, float2* __restrict__ Vt,
                    float* __restrict__ norm2, int m, int n, int kmax, int strips) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* xs = reinterpret_cast<float4*>(smem_raw);  // [m][KC]
    __shared__ float red[KC];
    const int strip = blockIdx.x % strips;
    const int b = blockIdx.x / strips;
    const int k = min(ranks[b], kmax);
    if (threadIdx.x < KC) red[threadIdx.x] = 0.f;
    for (int e = threadIdx.x; e < m * KC; e += 128) {
        const int c = e / m, t = e - c * m;  // t fastest: coalesced reads of the eigenvector rows
        float2 x = make_float2(0.f, 0.f);
        if (c < k) {
            const int p = perm[(size_t)b * m + c];
            const float f = inv[(size_t)b * m + c];
            const float2 w = W[((size_t)b * m + p) * m + t];
            x = make_float2(w.x * f, -w.y * f);  // conj(u_c[t])
        }
        xs[t * KC + c] = make_float4(x.x, x.x, -x.y, x.y);
    }
    __syncthreads();
    const int v0 = strip * 256 + threadIdx.x * 2;
    const bool live = v0 < n;
    float2 acc0[KC], acc1[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) acc0[c] = acc1[c] = make_float2(0.f, 0.f);
    if (live) {
        const float2* a = A + (size_t)b * m * n + v0;
#pragma unroll 8
        for (int t = 0; t < m; ++t) {
            const float4 av = __ldcs(reinterpret_cast<const float4*>(a + (size_t)t * n));
            const float2 a0 = make_float2(av.x, av.y), a0s = make_float2(av.y, av.x);
            const float2 a1 = make_float2(av.z, av.w), a1s = make_float2(av.w, av.z);
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                const float4 x = xs[t * KC + c];
                const float2 xa = make_float2(x.x, x.y), xb = make_float2(x.z, x.w);
                acc0[c] = __ffma2_rn(xa, a0, acc0[c]);
                acc0[c] = __ffma2_rn(xb, a0s, acc0[c]);
                acc1[c] = __ffma2_rn(xa, a1, acc1[c]);
                acc1[c] = __ffma2_rn(xb, a1s, acc1[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        if (c < kmax) {
            const bool valid = c < k;
            const float4 o = valid ? make_float4(acc0[c].x, acc0[c].y, acc1[c].x, acc1[c].y) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) *reinterpret_cast<float4*>(Vt + ((size_t)b * kmax + c) * n + v0) = o;
            float s = (valid && live) ? (o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w) : 0.f;
            s = warp_sum(s);
            if ((threadIdx.x & 31) == 0 && valid) atomicAdd(&red[c], s);
        }
    }
    __syncthreads();
    if (threadIdx.x < KC && threadIdx.x < k) atomicAdd(&norm2[(size_t)b * kmax + threadIdx.x], red[threadIdx.x]);
}

template <int KC>
static int launch_formv_smallk(vk_context* h, const float2* A, const float2* W, const int32_t* perm, const float* inv,
                               const int32_t* ranks, float2* Vt, float* norm2, int B, int m, int n, int kmax) {
    const int strips = (n + 255) / 256;
    const long long nblocks = (long long)B * strips;
    if (nblocks > 0x7fffffffLL) return vk_fail(h, VK_EINVAL, "formV: grid too large");
    const size_t smem = (size_t)m * KC * sizeof(float4);
    if (smem > 48 * 1024)
        VK_CUDA(h, cudaFuncSetAttribute(formv_smallk_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    formv_smallk_kernel<KC><<<(unsigned)nblocks, 128, smem, h->stream>>>(A, W, perm, inv, ranks, Vt, norm2, m, n, kmax, strips);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

// Fused factor formation for small ranks on the wide Gram path (north_star item (c)): ONE launch produces U_k, the refined
// singular values, the normalised Vt rows and the retained energy. The channels of a matrix are shared by a thread-block
// cluster (one CTA per 256-channel strip, up to 8); every CTA streams its strip of A once (as formv_smallk_kernel), the
// squared row norms are exchanged through distributed shared memory, and the rows leave the registers already divided by
// sigma - no atomics, no second pass over Vt, no memset, deterministic summation order.
// RS: the rows of A are split over RS groups of 128 threads (more loads in flight: with one group the kernel ran at
// 1.9 TB/s, 12 warps per SM, latency bound); the groups' partial sums meet in shared memory.
template <int KC, int RS>
__global__ void __launch_bounds__(128 * RS)
factors_fused_kernel(const float2* __restrict__ A, const float2* __restrict__ W, const int32_t* __restrict__ perm,
                     const float* __restrict__ inv, const int32_t* __restrict__ ranks, float2* __restrict__ U,
                     float* __restrict__ S, float2* __restrict__ Vt, float* __restrict__ stats, int m, int n, int kmax,
                     int strips) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int NT = 128 * RS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* xs = reinterpret_cast<float4*>(smem_raw);                     // [m][KC]
    float4* xch = xs + (size_t)m * KC;                                    // [RS - 1][KC][128] partial sums of the other groups
    __shared__ float red[4][KC];                       // per warp: squared norms of this CTA's part of every row
    __shared__ float sig[KC];
    const int strip = (int)cluster.block_rank();
    const int b = blockIdx.x / strips;
    const int k = min(ranks[b], kmax);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = threadIdx.x & 127, rs = threadIdx.x >> 7;
    __shared__ int sp[KC];
    __shared__ float sf[KC];
    const int v0 = strip * 256 + ch * 2;
    const bool live = v0 < n;
    const int mq = (m + RS - 1) / RS;
    const int t0 = rs * mq, t1 = min(m, t0 + mq);
    const float2* a = A + (size_t)b * m * n + v0;
    constexpr int G = 8;
    float4 cur[G], nxt[G];
    // the first row group of A is on its way before anything else happens
#pragma unroll
    for (int g = 0; g < G; ++g)
        cur[g] = (live && t0 + g < t1) ? __ldcs(reinterpret_cast<const float4*>(a + (size_t)(t0 + g) * n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < KC) {
        const bool on = (int)threadIdx.x < k;
        sp[threadIdx.x] = on ? perm[(size_t)b * m + threadIdx.x] : 0;
        sf[threadIdx.x] = on ? inv[(size_t)b * m + threadIdx.x] : 0.f;
    }
    __syncthreads();
    // coefficients conj(u_c[t]) = conj(W[perm c][t]) inv[c], expanded for the packed FMAs (independent loads: all in flight)
#pragma unroll 4
    for (int e = threadIdx.x; e < m * KC; e += NT) {
        const int c = e / m, t = e - c * m;  // t fastest: coalesced reads of the eigenvector rows
        const float f = sf[c];
        const float2 w = W[((size_t)b * m + sp[c]) * m + t];
        const float2 x = make_float2(w.x * f, -w.y * f);
        xs[t * KC + c] = make_float4(x.x, x.x, -x.y, x.y);
    }
    __syncthreads();
    float2 acc0[KC], acc1[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) acc0[c] = acc1[c] = make_float2(0.f, 0.f);
    if (live) {
        // register double buffer: the eight loads of the next row group are in flight while this one is consumed (left
        // to the compiler, the loads trailed their uses one by one: every first use stalled, 1.9 TB/s)
        for (int tb = t0; tb < t1; tb += G) {
#pragma unroll
            for (int g = 0; g < G; ++g)
                nxt[g] = (tb + G + g < t1) ? __ldcs(reinterpret_cast<const float4*>(a + (size_t)(tb + G + g) * n))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int t = min(tb + g, m - 1);         // rows past the end carry zeros
                const float4 av = cur[g];
                const float2 a0 = make_float2(av.x, av.y), a0s = make_float2(av.y, av.x);
                const float2 a1 = make_float2(av.z, av.w), a1s = make_float2(av.w, av.z);
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    const float4 x = xs[t * KC + c];
                    const float2 xa = make_float2(x.x, x.y), xb = make_float2(x.z, x.w);
                    acc0[c] = __ffma2_rn(xa, a0, acc0[c]);
                    acc0[c] = __ffma2_rn(xb, a0s, acc0[c]);
                    acc1[c] = __ffma2_rn(xa, a1, acc1[c]);
                    acc1[c] = __ffma2_rn(xb, a1s, acc1[c]);
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) cur[g] = nxt[g];
        }
    }
    if (RS > 1) {
        if (rs > 0) {
#pragma unroll
            for (int c = 0; c < KC; ++c)
                xch[((size_t)(rs - 1) * KC + c) * 128 + ch] = make_float4(acc0[c].x, acc0[c].y, acc1[c].x, acc1[c].y);
        }
        __syncthreads();
        if (rs == 0) {
#pragma unroll
            for (int g = 0; g < RS - 1; ++g)
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    const float4 o = xch[((size_t)g * KC + c) * 128 + ch];
                    acc0[c].x += o.x, acc0[c].y += o.y, acc1[c].x += o.z, acc1[c].y += o.w;
                }
        }
    }
    if (rs == 0) {
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            float s2 = (c < k && live) ? (acc0[c].x * acc0[c].x + acc0[c].y * acc0[c].y + acc1[c].x * acc1[c].x + acc1[c].y * acc1[c].y) : 0.f;
            s2 = warp_sum(s2);
            if (lane == 0) red[warp][c] = s2;
        }
    }
    cluster.sync();
    // sigma_c = || A^H u_c ||: the strips' parts in rank order, the warps' parts in warp order
    if (threadIdx.x < KC) {
        float tot = 0.f;
        for (int rk = 0; rk < strips; ++rk) {
            const float* peer = cluster.map_shared_rank(&red[0][0], rk);
#pragma unroll
            for (int w = 0; w < 4; ++w) tot += peer[w * KC + threadIdx.x];
        }
        sig[threadIdx.x] = sqrtf(tot);
    }
    __syncthreads();
    if (rs == 0) {
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            if (c < kmax && live) {
                const float sg = sig[c];
                const float f = (c < k && sg > 0.f) ? 1.f / sg : 0.f;
                *reinterpret_cast<float4*>(Vt + ((size_t)b * kmax + c) * n + v0) =
                    make_float4(acc0[c].x * f, acc0[c].y * f, acc1[c].x * f, acc1[c].y * f);
            }
        }
    }
    if (strip == 0) {
        if (threadIdx.x < k) S[(size_t)b * kmax + threadIdx.x] = sig[threadIdx.x];
        if (threadIdx.x == 0) {
            double e = 0.0;
            for (int c = 0; c < k; ++c) e += (double)sig[c] * (double)sig[c];
            stats[4 * b + 1] = (float)e;
        }
        // U[t][c] = u_c[t] = conj of what xs holds; zero beyond the rank
        float2* Ub = U + (size_t)b * m * kmax;
        for (int e = threadIdx.x; e < m * kmax; e += NT) {
            const int t = e / kmax, c = e - t * kmax;
            const float4 x = xs[t * KC + c];
            Ub[e] = c < k ? make_float2(x.x, -x.w) : make_float2(0.f, 0.f);
        }
    }
    cluster.sync();   // nobody leaves while a peer may still read its partial sums
}

template <int KC>
static int launch_factors_fused(vk_context* h, const float2* A, const float2* W, const int32_t* perm, const float* inv,
                                const int32_t* ranks, float2* U, float* S, float2* Vt, float* stats, int B, int m, int n,
                                int kmax) {
    const int strips = (n + 255) / 256;
    const long long nblocks = (long long)B * strips;
    if (nblocks > 0x7fffffffLL) return vk_fail(h, VK_EINVAL, "factors: grid too large");
    constexpr int RS = 1;
    const size_t smem = (size_t)m * KC * sizeof(float4) + (size_t)(RS - 1) * KC * 128 * sizeof(float4);
    if (smem > 48 * 1024)
        VK_CUDA(h, cudaFuncSetAttribute(factors_fused_kernel<KC, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nblocks, 1, 1);
    cfg.blockDim = dim3(128 * RS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)strips;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VK_CUDA(h, cudaLaunchKernelEx(&cfg, factors_fused_kernel<KC, RS>, A, W, perm, inv, ranks, U, S, Vt, stats, m, n, kmax, strips));
    h->launches += 1;
    return VK_OK;
}

struct FormUOp {  // U[t][c] = sum_j A[t][j] * W[perm c][j] inv[c]                    (tall, r = n)
    const float2* A;
    const float2* W;
    const int32_t* perm;
    const float* inv;
    const int32_t* ranks;
    float2* U;
    float* norm2;
    int m, n, kmax;
    static constexpr bool A_K_CONTIG = true, B_K_CONTIG = true;
    static constexpr int REDUCE = 2;
    __device__ int M(int) const { return m; }
    __device__ int N(int b) const { return ranks[b]; }
    __device__ int K(int) const { return n; }
    __host__ __device__ int Mfill() const { return m; }
    __host__ __device__ int Nfill() const { return kmax; }
    __device__ float2 loadA(int b, int t, int j) const { return A[((size_t)b * m + t) * n + j]; }
    __device__ float2 loadB(int b, int j, int c) const {
        const int p = perm[(size_t)b * n + c];
        const float f = inv[(size_t)b * n + c];
        const float2 v = W[((size_t)b * n + p) * n + j];
        return make_float2(v.x * f, v.y * f);
    }
    __device__ void store(int b, int t, int c, float2 v) const { U[((size_t)b * m + t) * kmax + c] = v; }
    __device__ void reduce_add(int b, int c, float v) const { atomicAdd(&norm2[(size_t)b * kmax + c], v); }
};

// S[b][c] = sqrt(norm2) ; Vt[b][c][:] /= S           (refined singular value = || A^H u_c ||)
__global__ void __launch_bounds__(256) scale_rows_kernel(float2* __restrict__ Vt, int n, int kmax,
                                                         const float* __restrict__ norm2,
                                                         const int32_t* __restrict__ ranks, float* __restrict__ S) {
    const int b = blockIdx.y, c = blockIdx.x;
    if (c >= ranks[b]) return;
    const float s = sqrtf(norm2[(size_t)b * kmax + c]);
    if (threadIdx.x == 0) S[(size_t)b * kmax + c] = s;
    const float f = s > 0.f ? 1.f / s : 0.f;
    float2* d = Vt + ((size_t)b * kmax + c) * n;
    for (int x = threadIdx.x; x < n; x += blockDim.x) {
        float2 v = d[x];
        v.x *= f;
        v.y *= f;
        d[x] = v;
    }
}
__global__ void __launch_bounds__(256) scale_cols_kernel(float2* __restrict__ U, int m, int kmax,
                                                         const float* __restrict__ norm2,
                                                         const int32_t* __restrict__ ranks, float* __restrict__ S) {
    const int b = blockIdx.y;
    const int k = ranks[b];
    const size_t tot = (size_t)m * kmax;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % kmax);
        if (c < k) {
            const float s = sqrtf(norm2[(size_t)b * kmax + c]);
            const float f = s > 0.f ? 1.f / s : 0.f;
            float2 v = U[(size_t)b * tot + e];
            v.x *= f;
            v.y *= f;
            U[(size_t)b * tot + e] = v;
            if (e < (size_t)kmax) S[(size_t)b * kmax + c] = s;
        }
    }
}
__global__ void retained_energy_kernel(const float* __restrict__ S, int kmax, int B, float* __restrict__ stats) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        double e = 0.0;
        for (int c = 0; c < kmax; ++c) e += (double)S[(size_t)b * kmax + c] * (double)S[(size_t)b * kmax + c];
        stats[4 * b + 1] = (float)e;
    }
}

// =================================================================================================================
// Reconstruction                                         reference: reconstruct_vis, decompress_ms.py:107-131
// =================================================================================================================
struct ReconOp {  // out[t][v] = sum_c (U[t][c] S[c]) Vt[c][v]
    const float2* U;
    const float* S;
    const float2* Vt;
    const int32_t* ranks;
    float2* out;
    int m, n, kmax;
    static constexpr bool A_K_CONTIG = true, B_K_CONTIG = false;
    static constexpr int REDUCE = 0;
    __device__ int M(int) const { return m; }
    __device__ int N(int) const { return n; }
    __device__ int K(int b) const {
        if (!ranks) return kmax;
        const int k = ranks[b];
        return k < kmax ? (k < 0 ? 0 : k) : kmax;
    }
    __host__ __device__ int Mfill() const { return m; }
    __host__ __device__ int Nfill() const { return n; }
    __device__ float2 loadA(int b, int t, int c) const {
        const float2 u = U[((size_t)b * m + t) * kmax + c];
        const float s = S[(size_t)b * kmax + c];
        return make_float2(u.x * s, u.y * s);
    }
    __device__ float2 loadB(int b, int c, int j) const { return Vt[((size_t)b * kmax + c) * n + j]; }
    __device__ void store(int b, int t, int j, float2 v) const { out[((size_t)b * m + t) * n + j] = v; }
    __device__ void reduce_add(int, int, float) const {}
};

// Rank-k reconstruction for small k (north_star item (d)): out[t][v] = sum_c (U[t][c] S[c]) Vt[c][v], k <= KC <= 8.
// HBM-bound by the output stream, so everything else is arranged around 128-bit coalesced streaming stores:
//   * a thread owns two adjacent channels (one float4 of output per row) and keeps its Vt[c][v0..v0+1] values, plus
//     their (im, re) swapped twins, in registers for the whole row loop;
//   * U*S for the CTA's rows sits in shared memory pre-expanded as (ar, ar, -ai, ai), so one complex MAC is two packed
//     fma.rn.f32x2 (FFMA2): acc += (ar, ar) * (br, bi) + (-ai, ai) * (bi, br);
//   * rows are split across CTAs (ROWS per CTA) so a small batch still fills 148 SMs.
template <int KC, int ROWS>
__global__ void __launch_bounds__(128)
recon_smallk_kernel(const float2* __restrict__ U, const float* __restrict__ S, const float2* __restrict__ Vt,
                    const int32_t* __restrict__ ranks, float2* __restrict__ out, int m, int n, int kmax, int strips,
                    int rsplit) {
    __shared__ float4 us[ROWS * KC];
    int bid = blockIdx.x;
    const int rs = bid % rsplit;
    bid /= rsplit;
    const int strip = bid % strips;
    const int b = bid / strips;
    const int t0 = rs * ROWS;
    const int rows = min(ROWS, m - t0);
    // only the first ranks[b] modes count (include/visco_b200.h): whatever sits beyond them is never read
    int kr = kmax;
    if (ranks) kr = min(max(ranks[b], 0), kmax);
    for (int e = threadIdx.x; e < rows * KC; e += 128) {
        const int t = e / KC, c = e - t * KC;
        float2 u = make_float2(0.f, 0.f);
        float s = 0.f;
        if (c < kr) {
            u = U[((size_t)b * m + t0 + t) * kmax + c];
            s = S[(size_t)b * kmax + c];
        }
        us[e] = make_float4(u.x * s, u.x * s, -u.y * s, u.y * s);
    }
    const int v0 = strip * 256 + threadIdx.x * 2;
    const bool live = v0 < n;  // n is even on this path, so v0 + 1 < n too
    float2 va[KC], vas[KC], vb[KC], vbs[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && c < kr) v = *reinterpret_cast<const float4*>(Vt + ((size_t)b * kmax + c) * n + v0);
        va[c] = make_float2(v.x, v.y);
        vas[c] = make_float2(v.y, v.x);
        vb[c] = make_float2(v.z, v.w);
        vbs[c] = make_float2(v.w, v.z);
    }
    __syncthreads();
    if (!live) return;
    float2* o = out + ((size_t)b * m + t0) * n + v0;
#pragma unroll 4
    for (int t = 0; t < rows; ++t) {
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            const float4 a = us[t * KC + c];
            const float2 aa = make_float2(a.x, a.y), ab = make_float2(a.z, a.w);
            acc0 = __ffma2_rn(aa, va[c], acc0);
            acc0 = __ffma2_rn(ab, vas[c], acc0);
            acc1 = __ffma2_rn(aa, vb[c], acc1);
            acc1 = __ffma2_rn(ab, vbs[c], acc1);
        }
        __stcs(reinterpret_cast<float4*>(o + (size_t)t * n), make_float4(acc0.x, acc0.y, acc1.x, acc1.y));
    }
}

template <int KC>
static int launch_recon_smallk(vk_context* h, const float2* U, const float* S, const float2* Vt, const int32_t* ranks,
                               int B, int m, int n, int kmax, float2* out) {
    constexpr int ROWS = 64;
    const int strips = (n + 255) / 256;
    const int rsplit = (m + ROWS - 1) / ROWS;
    const long long nblocks = (long long)B * strips * rsplit;
    if (nblocks > 0x7fffffffLL) return vk_fail(h, VK_EINVAL, "reconstruct: grid too large");
    recon_smallk_kernel<KC, ROWS><<<(unsigned)nblocks, 128, 0, h->stream>>>(U, S, Vt, ranks, out, m, n, kmax, strips, rsplit);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

// =================================================================================================================
// Synthetic MeerKAT-like visibilities
// =================================================================================================================
__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ float u01(uint64_t h) { return ((h >> 40) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(256)
synth_kernel(float2* __restrict__ A, int nbl_local, int ncorr, int m, int n, int bl_offset, int nbl_total,
             uint64_t seed) {
    constexpr int NSRC = 10;
    __shared__ float rho[NSRC], phi[NSRC];
    const int bc = blockIdx.y;  // local matrix index = bl * ncorr + corr
    const int bl = bc / ncorr, corr = bc - bl * ncorr;
    const int gbl = bl_offset + bl;
    if (threadIdx.x < NSRC) {
        const uint64_t k = mix64(seed ^ mix64(0x51ull + (uint64_t)gbl * 64 + threadIdx.x));
        const float R = 30.f * (float)(gbl + 1) / (float)nbl_total;
        rho[threadIdx.x] = (2.f * u01(k) - 1.f) * R;
        phi[threadIdx.x] = 6.2831853f * u01(mix64(k));
    }
    __syncthreads();
    const float gain = (corr == 0 || corr == ncorr - 1) ? 1.f : 0.01f;
    const size_t per = (size_t)m * n;
    float2* Ab = A + (size_t)bc * per;
    const uint64_t key = mix64(seed ^ mix64(0xABCDull + (uint64_t)gbl * 16 + corr));
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(e / n), v = (int)(e - (size_t)t * n);
        const float ft = (float)t / (float)m;
        const float fv = 1.f + 0.2f * (float)v / (float)n;
        float re = 0.f, im = 0.f;
#pragma unroll
        for (int s = 0; s < NSRC; ++s) {
            float sn, cs;
            sincosf(6.2831853f * rho[s] * ft * fv + phi[s], &sn, &cs);
            re += cs;
            im += sn;
        }
        const uint64_t hh = mix64(key ^ (uint64_t)e);
        const float u1 = u01(hh), u2 = u01(mix64(hh));
        const float rad = sqrtf(-2.f * logf(u1)) * 0.70710678f;
        float sn, cs;
        sincosf(6.2831853f * u2, &sn, &cs);
        Ab[e] = make_float2(gain * re + rad * cs, gain * im + rad * sn);
    }
}

}  // namespace

// =================================================================================================================
// launchers
// =================================================================================================================
int vk_launch_gram_simt(vk_context* h, const float2* A, int B, int m, int n, int side, float2* W) {
    if (side == 0) {
        GramOp0 op{A, W, m, n};
        return cgemm_launch<64, 64, 4, 4, 16>(h, op, B);
    }
    GramOp1 op{A, W, m, n};
    return cgemm_launch<64, 64, 4, 4, 16>(h, op, B);
}

bool vk_gram_small_supported(int m, int n) { return (m < n ? m : n) <= 64; }

int vk_launch_gram_small(vk_context* h, const float2* A, int B, int m, int n, float2* W, float* gscale_dev,
                         int32_t* nonfinite_dev, int32_t* bad_dev, int32_t* nbad_dev) {
    if (B <= 0) return VK_OK;
    if (m <= n) gram_small_kernel<false><<<B, 160, 0, h->stream>>>(A, m, n, W, gscale_dev, nonfinite_dev, bad_dev, nbad_dev);
    else gram_small_kernel<true><<<B, 160, 0, h->stream>>>(A, m, n, W, gscale_dev, nonfinite_dev, bad_dev, nbad_dev);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int vk_launch_gram_normalise(vk_context* h, float2* W, int B, int r, float* gscale_dev, int32_t* nonfinite_dev,
                             int32_t* bad_dev, int32_t* nbad_dev) {
    gram_normalise_kernel<<<B, r >= 128 ? 1024 : 256, 0, h->stream>>>(W, r, gscale_dev, nonfinite_dev, bad_dev, nbad_dev);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int vk_launch_pack_small(vk_context* h, const float2* A, int B, int m, int n, float2* W, int ld, float* gscale_dev,
                         int32_t* nonfinite_dev) {
    pack_small_kernel<<<B, 256, 0, h->stream>>>(A, m, n, W, ld, gscale_dev, nonfinite_dev);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int vk_launch_select(vk_context* h, const float2* W, int B, int r, int ldot, int ld, const float* gscale_dev,
                     int mode_gram, int fixed_rank, double decorrelation, int kmax, int32_t* perm_dev, float* inv_dev,
                     float* S_dev, int32_t* ranks_dev, float* stats_dev, const int32_t* sweeps_dev,
                     const int32_t* done_dev) {
    int P = 1;
    while (P < r) P <<= 1;
    const size_t smem = (size_t)P * 12;
    // (r <= 64: 64 threads - the kernel is a latency chain per matrix, smaller CTAs put four times as many on an SM)
    select_kernel<<<B, r <= 64 ? 64 : 256, smem, h->stream>>>(W, r, ldot, ld, gscale_dev, mode_gram, fixed_rank, decorrelation, kmax,
                                               perm_dev, inv_dev, S_dev, ranks_dev, stats_dev, sweeps_dev, done_dev);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

// out[0] = number of matrices whose eigensolver reported failure (done == 0)
__global__ void count_not_done_kernel(const int32_t* __restrict__ done, int B, int32_t* __restrict__ out) {
    int c = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) c += done[b] == 0;
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
        out[0] = t;
    }
}

// flags[b] = 1 when the smallest retained singular value of matrix b is below thr * the largest: the Gram matrix (float32)
// resolves lambda only down to ~1e-7 lambda_max, i.e. sigma down to a few 1e-2 sigma_max at the 1e-4 relative level
__global__ void flag_illcond_kernel(const float* __restrict__ S, const int32_t* __restrict__ ranks, int B, int kmax,
                                    float thr, int32_t* __restrict__ flags, int32_t* __restrict__ count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int k = ranks[b];
    int f = 0;
    if (k > 1) {
        const float s0 = S[(size_t)b * kmax], sk = S[(size_t)b * kmax + k - 1];
        f = (s0 > 0.f && sk < thr * s0) ? 1 : 0;
    }
    flags[b] = f;
    if (f) atomicAdd(count, 1);
}

int vk_launch_flag_illcond(vk_context* h, const float* S, const int32_t* ranks, int B, int kmax, float thr, int32_t* flags,
                           int32_t* count) {
    VK_CUDA(h, cudaMemsetAsync(count, 0, sizeof(int32_t), h->stream));
    flag_illcond_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(S, ranks, B, kmax, thr, flags, count);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int vk_launch_count_not_done(vk_context* h, const int32_t* done, int B, int32_t* out) {
    count_not_done_kernel<<<1, 256, 0, h->stream>>>(done, B, out);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int vk_launch_pack_info(vk_context* h, const int32_t* sweeps, const int32_t* done, int B, int32_t* info) {
    pack_info_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(sweeps, done, B, info);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int vk_launch_find_n(vk_context* h, const float* S, int B, int r, double decorrelation, int32_t* ranks) {
    find_n_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(S, B, r, decorrelation, ranks);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

static int launch_rows(vk_context* h, const float2* W, int r, int ld, int off, int len, int kmax, const int32_t* perm,
                       const float* inv, const int32_t* ranks, int conj, int scaled, float2* dst, int B) {
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
        rows_from_vectors_kernel<<<dim3(kmax, nb), 256, 0, h->stream>>>(
            W + (size_t)b0 * r * ld, r, ld, off, len, kmax, perm + (size_t)b0 * r, inv + (size_t)b0 * r, ranks + b0,
            conj, scaled, dst + (size_t)b0 * kmax * len);
        VK_LAUNCH_CHECK(h);
    }
    return VK_OK;
}
static int launch_cols(vk_context* h, const float2* W, int r, int ld, int off, int len, int kmax, const int32_t* perm,
                       const float* inv, const int32_t* ranks, int conj, int scaled, float2* dst, int B) {
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
        cols_from_vectors_kernel<<<dim3((len + 31) / 32, (kmax + 31) / 32, nb), 256, 0, h->stream>>>(
            W + (size_t)b0 * r * ld, r, ld, off, len, kmax, perm + (size_t)b0 * r, inv + (size_t)b0 * r, ranks + b0,
            conj, scaled, dst + (size_t)b0 * len * kmax);
        VK_LAUNCH_CHECK(h);
    }
    return VK_OK;
}

int vk_launch_factors_small(vk_context* h, const float2* W, int ld, int B, int m, int n, int kmax,
                            const int32_t* perm_dev, const float* inv_dev, const int32_t* ranks_dev, float2* U,
                            float2* Vt) {
    const int r = m <= n ? m : n;
    int rc;
    if (m <= n) {
        // vectors = rotated rows of A (length n) then accumulated rotations R (length m): Vt = row / sigma, U = R^H
        if ((rc = launch_rows(h, W, r, ld, 0, n, kmax, perm_dev, inv_dev, ranks_dev, 0, 1, Vt, B))) return rc;
        if ((rc = launch_cols(h, W, r, ld, n, m, kmax, perm_dev, inv_dev, ranks_dev, 1, 0, U, B))) return rc;
    } else {
        // vectors = rotated columns of A (length m) then accumulated rotations (length n): U = col / sigma, Vt = M^H
        if ((rc = launch_cols(h, W, r, ld, 0, m, kmax, perm_dev, inv_dev, ranks_dev, 0, 1, U, B))) return rc;
        if ((rc = launch_rows(h, W, r, ld, m, n, kmax, perm_dev, inv_dev, ranks_dev, 1, 0, Vt, B))) return rc;
    }
    return VK_OK;
}

int vk_launch_factors_gram(vk_context* h, const float2* A, const float2* W, int B, int m, int n, int side, int kmax,
                           const int32_t* perm_dev, const float* inv_dev, const int32_t* ranks_dev, float* norm2_dev,
                           float2* U, float* S, float2* Vt, float* stats_dev, float2* xbuf) {
    int rc;
    if (side == 0 && kmax <= 16 && (n % 2) == 0 && (n + 255) / 256 <= 8 && h->factors_impl == 0 && !h->recon_generic &&
        ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Vt)) % 16) == 0 &&
        (size_t)(m + 3 * 128) * (kmax <= 8 ? 8 : 16) * sizeof(float4) <= VK_SMEM_BUDGET) {
        // small rank, wide matrix: everything in one launch (a cluster of up to eight 256-channel strips per matrix)
        if (kmax <= 2) return launch_factors_fused<2>(h, A, W, perm_dev, inv_dev, ranks_dev, U, S, Vt, stats_dev, B, m, n, kmax);
        if (kmax <= 4) return launch_factors_fused<4>(h, A, W, perm_dev, inv_dev, ranks_dev, U, S, Vt, stats_dev, B, m, n, kmax);
        if (kmax <= 8) return launch_factors_fused<8>(h, A, W, perm_dev, inv_dev, ranks_dev, U, S, Vt, stats_dev, B, m, n, kmax);
        return launch_factors_fused<16>(h, A, W, perm_dev, inv_dev, ranks_dev, U, S, Vt, stats_dev, B, m, n, kmax);
    }
    VK_CUDA(h, cudaMemsetAsync(norm2_dev, 0, sizeof(float) * (size_t)B * kmax, h->stream));
    if (side == 0) {
        const int r = m;
        if ((rc = launch_cols(h, W, r, r, 0, m, kmax, perm_dev, inv_dev, ranks_dev, 0, 1, U, B))) return rc;
        FormVOp op{A, W, perm_dev, inv_dev, ranks_dev, Vt, norm2_dev, m, n, kmax};
        const bool aligned = (n % 2 == 0) && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Vt)) % 16 == 0);
        // (the K-major operand X has rows of m complex numbers: its tensor map needs 16-byte strides, i.e. an even m)
        if (xbuf && h->gemm_impl == 0 && kmax > 16 && (m % 2) == 0 && vk_cgemm_tc_supported(m, n, kmax) && aligned &&
            (reinterpret_cast<uintptr_t>(xbuf) % 16) == 0) {
            // large rank: X = conj(U_k)^T / lambda materialised K-major, then the tcgen05 complex GEMM
            if ((rc = launch_rows(h, W, r, r, 0, m, kmax, perm_dev, inv_dev, ranks_dev, 1, 1, xbuf, B))) return rc;
            rc = vk_launch_formv_tc(h, xbuf, A, ranks_dev, Vt, norm2_dev, B, m, n, kmax);
        } else if (aligned && kmax <= 16 && (size_t)m * (kmax <= 8 ? 8 : 16) * sizeof(float4) <= VK_SMEM_BUDGET &&
                   !h->recon_generic) {
            if (kmax <= 2)
                rc = launch_formv_smallk<2>(h, A, W, perm_dev, inv_dev, ranks_dev, Vt, norm2_dev, B, m, n, kmax);
            else if (kmax <= 4)
                rc = launch_formv_smallk<4>(h, A, W, perm_dev, inv_dev, ranks_dev, Vt, norm2_dev, B, m, n, kmax);
            else if (kmax <= 8)
                rc = launch_formv_smallk<8>(h, A, W, perm_dev, inv_dev, ranks_dev, Vt, norm2_dev, B, m, n, kmax);
            else
                rc = launch_formv_smallk<16>(h, A, W, perm_dev, inv_dev, ranks_dev, Vt, norm2_dev, B, m, n, kmax);
        } else if (kmax <= 8)
            rc = cgemm_launch<8, 256, 8, 2, 16>(h, op, B);
        else if (kmax <= 32)
            rc = cgemm_launch<32, 128, 4, 4, 16>(h, op, B);
        else
            rc = cgemm_launch<64, 64, 4, 4, 16>(h, op, B);
        if (rc) return rc;
        for (int b0 = 0; b0 < B; b0 += 65535) {
            const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
            scale_rows_kernel<<<dim3(kmax, nb), 256, 0, h->stream>>>(Vt + (size_t)b0 * kmax * n, n, kmax,
                                                                      norm2_dev + (size_t)b0 * kmax, ranks_dev + b0,
                                                                      S + (size_t)b0 * kmax);
            VK_LAUNCH_CHECK(h);
        }
    } else {
        const int r = n;
        if ((rc = launch_rows(h, W, r, r, 0, n, kmax, perm_dev, inv_dev, ranks_dev, 1, 1, Vt, B))) return rc;
        FormUOp op{A, W, perm_dev, inv_dev, ranks_dev, U, norm2_dev, m, n, kmax};
        if (kmax <= 8)
            rc = cgemm_launch<256, 8, 2, 8, 16>(h, op, B);
        else if (kmax <= 32)
            rc = cgemm_launch<128, 32, 4, 4, 16>(h, op, B);
        else
            rc = cgemm_launch<64, 64, 4, 4, 16>(h, op, B);
        if (rc) return rc;
        for (int b0 = 0; b0 < B; b0 += 65535) {
            const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
            int gx = (int)(((size_t)m * kmax + 255) / 256);
            if (gx > 64) gx = 64;
            scale_cols_kernel<<<dim3(gx, nb), 256, 0, h->stream>>>(U + (size_t)b0 * m * kmax, m, kmax,
                                                                    norm2_dev + (size_t)b0 * kmax, ranks_dev + b0,
                                                                    S + (size_t)b0 * kmax);
            VK_LAUNCH_CHECK(h);
        }
    }
    retained_energy_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(S, kmax, B, stats_dev);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int vk_launch_reconstruct(vk_context* h, const float2* U, const float* S, const float2* Vt, const int32_t* ranks, int B,
                          int m, int n, int kmax, float2* out) {
    // small rank: dedicated streaming kernel (modes >= ranks[b] are skipped, as on the other two paths)
    const bool aligned = (n % 2 == 0) && ((reinterpret_cast<uintptr_t>(Vt) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
    // (k <= 16 stays below the FFMA2 limit of ~0.7 x HBM; the tcgen05 GEMM only wins from k ~ 20 on, measured)
    if (aligned && h->recon_tc_impl == 0 && !h->recon_generic && vk_recon_tc_supported(m, n, kmax) &&
        (reinterpret_cast<uintptr_t>(U) % 8) == 0)
        return vk_launch_recon_tc_smallk(h, U, S, Vt, ranks, out, B, m, n, kmax);   // 8 < k <= 32: persistent tcgen05 kernel
    if (aligned && kmax <= 16 && !h->recon_generic) {
        if (kmax <= 2) return launch_recon_smallk<2>(h, U, S, Vt, ranks, B, m, n, kmax, out);
        if (kmax <= 4) return launch_recon_smallk<4>(h, U, S, Vt, ranks, B, m, n, kmax, out);
        if (kmax <= 8) return launch_recon_smallk<8>(h, U, S, Vt, ranks, B, m, n, kmax, out);
        return launch_recon_smallk<16>(h, U, S, Vt, ranks, B, m, n, kmax, out);
    }
    if (h->gemm_impl == 0 && kmax > 16 && vk_cgemm_tc_supported(m, n, kmax) &&
        ((reinterpret_cast<uintptr_t>(U) | reinterpret_cast<uintptr_t>(Vt) | reinterpret_cast<uintptr_t>(out)) % 16 == 0))
        return vk_launch_recon_tc(h, U, S, Vt, ranks, out, B, m, n, kmax);
    ReconOp op{U, S, Vt, ranks, out, m, n, kmax};
    return cgemm_launch<64, 64, 4, 4, 8>(h, op, B);
}

int vk_launch_synth(vk_context* h, float2* A, int nbl_local, int ncorr, int m, int n, int bl_offset, int nbl_total,
                    uint64_t seed) {
    const int nmat = nbl_local * ncorr;
    int gx = (int)(((size_t)m * n + 255) / 256);
    if (gx > 128) gx = 128;
    for (int b0 = 0; b0 < nmat; b0 += 65532) {  // multiple of 4 keeps (bl, corr) decoding intact for ncorr | 65532
        int nb = nmat - b0;
        if (nb > 65532) nb = 65532;
        if (b0 % ncorr != 0) return vk_fail(h, VK_EINVAL, "synth: ncorr must divide 65532");
        synth_kernel<<<dim3(gx, nb), 256, 0, h->stream>>>(A + (size_t)b0 * m * n, nbl_local, ncorr, m, n,
                                                          bl_offset + b0 / ncorr, nbl_total, seed);
        VK_LAUNCH_CHECK(h);
    }
    return VK_OK;
}
