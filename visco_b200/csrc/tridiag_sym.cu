// Householder tridiagonalisation, lower triangle only and with deferred updates ("tridiag_impl" = 0, the default for
// 128 < r <= 512). Same contract as the kernels of tridiag.cu (conventions in its header): on exit d, e, tau, ph hold the
// real tridiagonal, the reflector scales and the accumulated sub-diagonal phases, and row j of M right of the diagonal holds
// reflector j. Part of the replacement of the LAPACK cgesdd call behind np.linalg.svd (reference compress_ms.py:350).
//
// Why another kernel: the full-storage kernels read (and, undeferred, write) the whole trailing block every Householder
// step - 16 r^3 / 3 bytes per matrix undeferred, 8 r^3 / 3 (1 + 2/NB) deferred - and, worse, execute 50-100 instructions
// per matrix element (ncu: issue slots 55 % busy at four warps per scheduler, i.e. instruction bound, DRAM at 12 %).
// A Hermitian matrix needs only its lower triangle: element (i, k), k < i, serves both (A v)_i += a_ik v_k and
// (A v)_k += conj(a_ik) v_i. Together with LAPACK-latrd style deferral (up to NB reflector pairs (v_s, w_s) pending in
// shared memory; a step only READS the triangle as the last update pass left it and corrects the product,
//     A v = A0 v - sum_s [ v_s (w_s^H v) + w_s (v_s^H v) ] )
// the bytes per matrix drop to 4 r^3 / 3 (1 + 2/NB), and the pass is organised so that an element costs ~0.4 warp
// instructions:
//
//   * one CTA (16 warps) per matrix. The live part of the triangle is cut into tiles of TR rows x 32 columns; a warp owns
//     whole row blocks (dealt out serpentine-wise, so that long and short blocks pair up) and walks along a block from the
//     first live column chunk to the chunk that holds the diagonal. Lane l owns column 32 K + l of tile K.
//   * row part (A v)_i, k <= i: per-lane partial sums racc[TR] carried along the whole block and reduced across the lanes
//     once per block by a butterfly reduce-scatter; column part, k < i: summed over the TR rows in registers and added to
//     the warp's own slice part[warp][k] in shared memory, combined by the vector update after the pass;
//   * no predicates inside a block: only the tile with the diagonal is masked. Rows and columns at or left of j that
//     share a tile with live ones are harmless - their v entries are zero and what they produce is never read;
//   * the next tile's TR loads are issued before the current tile is used (register double buffer);
//   * v^H A v = 2 Re sum_i conj(v_i) rowpart_i - sum_i a_ii |v_i|^2 comes out of the row parts alone;
//   * column j+1 of the triangle, which defines the next reflector, is captured by the pass itself (the lane that holds
//     it stores its conjugate into nrow[]), so no step ever issues a strided column read;
//   * every NB-th pass first applies the NB pending pairs to its tiles (read + write) - (v, w) are interleaved as float4
//     so one 128-bit shared-memory load serves a row or a column of the rank-2 update;
//   * once the trailing block fits into shared memory (next to the vectors; it overlays part[]) it moves there with
//     everything pending applied and undeferred resident steps (as in kernel 1d of tridiag.cu) finish the reduction.
#include <cmath>
#include <cstdio>

#include "common.cuh"

namespace {


__device__ __forceinline__ float2 cmulf2(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// x -= vi conj(wk) + wi conj(vk)
__device__ __forceinline__ void rank2(float2& x, float2 vi, float2 wi, float2 vk, float2 wk) {
    x.x = fmaf(-vi.x, wk.x, fmaf(-vi.y, wk.y, fmaf(-wi.x, vk.x, fmaf(-wi.y, vk.y, x.x))));
    x.y = fmaf(-vi.y, wk.x, fmaf(vi.x, wk.y, fmaf(-wi.y, vk.x, fmaf(wi.x, vk.y, x.y))));
}
// pending pair s at index i: (v.x, v.y, w.x, w.y)
__device__ __forceinline__ void rank2p(float2& x, float4 a, float4 c) {
    rank2(x, make_float2(a.x, a.y), make_float2(a.z, a.w), make_float2(c.x, c.y), make_float2(c.z, c.w));
}

// sum t[q] over the 32 lanes; on return t[0] of lane l holds the total of row (l >> (5 - log2 N))
template <int N>
__device__ __forceinline__ void reduce_scatter(float2 (&t)[N], int lane) {
    int bit = 16;
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int q = 0; q < h; ++q) {
            const float2 send = up ? t[q] : t[q + h];
            const float2 keep = up ? t[q + h] : t[q];
            t[q].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, bit);
            t[q].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, bit);
        }
        bit >>= 1;
    }
#pragma unroll
    for (int o = 16 / N; o >= 1; o >>= 1) {   // the lanes that still share a row
        t[0].x += __shfl_xor_sync(0xffffffffu, t[0].x, o);
        t[0].y += __shfl_xor_sync(0xffffffffu, t[0].y, o);
    }
}

// one 128-byte line into L2 (no register, no scoreboard entry): see the PF template parameter below
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct TileIt {
    int rho, I, K, Klast;   // serpentine round, row block, column chunk, chunk with the diagonal of the block
    bool valid;
};

// NT threads per CTA (512: one CTA per SM, the whole SM on one matrix - few matrices; 256: two CTAs, i.e. two matrices, per
// SM so that one's barriers and scalar phases hide behind the other's pass - many matrices); DB: register double buffer
// LDC: the row stride when it is known at compile time (0: use ld) - the eight loads of a tile then share one base
// register pair and immediate offsets instead of a 64-bit add each
// PF (r = 512, two matrices per SM, no register double buffer): the triangles of a wave do not fit the L2 (296 MB at the
// start, and the first third of the steps requests 70 % of the bytes), so a tile load is a DRAM access the warp then waits
// for. A warp asks the L2 for the tile it will load pfd tiles later - one prefetch instruction per tile, a lane per
// 128-byte line - and, when its walk of step j is over, for the first tiles of its walk of step j + 1. Measured on the
// MeerKAT shard (1040 matrices): 61.5 -> 54.5 ms with pfd = 1; 2: 55-59, 4: 55-60, 8: 67, 12: 70 ms (further ahead the
// lines are gone again before they are used). The 512-thread shape already loads one tile ahead into registers and loses
// with the extra instructions (63 -> 67-70 ms). Also measured, without effect on this kernel: L2 eviction-priority hints
// (evict_last on the bottom rows of every triangle / on a fraction of the lines, evict_first on the rest) and a tile-major
// copy of the triangle (a tile as TR * 256 contiguous bytes instead of TR pieces 4 KiB apart).
template <int EPL, int TR, int NB, int NT, bool DB, bool RAGGED, int LDC, bool PF = false>
__global__ void __launch_bounds__(NT, NT == 512 ? 1 : 2)
    tridiag_symdefer_kernel(float2* __restrict__ Wall, int r, int ld_, size_t wstride, float* __restrict__ dall,
                            float* __restrict__ eall, float* __restrict__ tauall, float2* __restrict__ phall, int nts,
                            int pfd) {
    constexpr int WD = EPL * 32;
    const int ld = LDC ? LDC : ld_;
    constexpr int SD_THREADS = NT, SD_WARPS = NT / 32;
    static_assert(NB <= SD_WARPS, "one warp per pending pair computes its two inner products");
    static_assert(TR == 8 || TR == 16, "row block height");
    extern __shared__ float2 sd_sm[];
    float2* vprev = sd_sm;            // resident phase only: pending single pair (vprev, wv)
    float2* vnew = sd_sm + WD;
    float2* wv = sd_sm + 2 * WD;
    float2* pv = sd_sm + 3 * WD;      // row parts of the product
    float2* nrow = sd_sm + 4 * WD;    // row j+1 (= conj of column j+1 of the triangle) as the pass of step j left it
    float4* VW = reinterpret_cast<float4*>(sd_sm + 5 * WD);   // [NB][WD] pending (v, w) pairs (streaming phase)
    float2* part = sd_sm + (5 + 2 * NB) * WD;                  // [SD_WARPS][WD] column parts (streaming phase)
    float2* T = part;                 // [ts][ts] trailing block once it fits (ts <= nts): rows/cols j0 .. r-1
    int j0 = -1, ts = 0;
    __shared__ float s_part[SD_WARPS];
    __shared__ float s_kpart[SD_WARPS];
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
    for (int k = tid; k < (5 + 2 * NB) * WD; k += SD_THREADS) sd_sm[k] = make_float2(0.f, 0.f);
    __syncthreads();
    int P = 0;  // pending pairs of the streaming phase (uniform)
    const int nblk = (r + TR - 1) / TR;

#ifdef VK_TRIDIAG_CLOCKS
    long long ck[6] = {0, 0, 0, 0, 0, 0}, c0 = clock64(), c1;
#define VK_CK(i) do { c1 = clock64(); ck[i] += c1 - c0; c0 = c1; } while (0)
#else
#define VK_CK(i)
#endif
    for (int j = 0; j + 2 < r; ++j) {
        const int e0 = (j + 1) >> 5;
        if (j0 < 0 && r - j <= nts) {
            // from here on the trailing block lives in shared memory (full storage, rebuilt from the lower triangle), with
            // every pending pair applied on the way in
            j0 = j, ts = r - j;
            for (int idx = tid; idx < ts * ts; idx += SD_THREADS) {
                const int i = j0 + idx / ts, k = j0 + idx % ts;
                float2 x;
                if (k <= i) x = M[(size_t)i * ld + k];
                else {
                    x = M[(size_t)k * ld + i];
                    x.y = -x.y;
                }
                for (int s2 = 0; s2 < P; ++s2) rank2p(x, VW[s2 * WD + i], VW[s2 * WD + k]);
                T[idx] = x;
            }
            __syncthreads();
            for (int k = j + tid; k < r; k += SD_THREADS) nrow[k] = T[k - j0];   // row j of the updated block
            for (int k = tid; k < WD; k += SD_THREADS) vprev[k] = wv[k] = make_float2(0.f, 0.f);
            P = 0;
            __syncthreads();
        }
        const bool resident = j0 >= 0;
        // row j with everything pending applied: diagonal d_j and the column below it (a = conj(row))
        float ss = 0.f;
        {
            const float2 v0 = vprev[j], w0 = wv[j];  // zero outside the resident phase
            for (int k = tid; k < WD; k += SD_THREADS) {
                float2 a = make_float2(0.f, 0.f);
                if (k >= j && k < r) {
                    float2 x = (j == 0) ? M[k] : nrow[k];
                    if (resident) {
                        rank2(x, v0, w0, vprev[k], wv[k]);
                    } else {
                        for (int s2 = 0; s2 < P; ++s2) rank2p(x, VW[s2 * WD + j], VW[s2 * WD + k]);
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
        VK_CK(0);
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < SD_WARPS; ++w) tot += s_part[w];
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
                phase = cmulf2(phase, make_float2(-p1.x, -p1.y));  // sub-diagonal element is -p1 * xn
            }
            if (tid == 0) {
                vnew[j + 1] = v0;
                taus[j] = tau;
                e[j] = ej;
                ph[j + 1] = phase;
            }
        }
        __syncthreads();
        VK_CK(1);
        // the reflector replaces row j right of the diagonal (upper triangle: never read by the streaming passes)
        {
            float2* row = M + (size_t)j * ld;
            for (int k = j + 1 + tid; k < r; k += SD_THREADS) row[k] = vnew[k];
        }
        float kacc = 0.f;  // this warp's part of v^H (A0 v) resp. v^H p (identical on all its lanes)
        const bool upd = !resident && (P == NB);
        if (resident) {
            for (int i = j + 1 + warp; i < r; i += SD_WARPS) {
                float2* row = T + (size_t)(i - j0) * ts - j0;
                const float2 vi = vprev[i], wi = wv[i];
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll 2
                for (int k = j + 1 + lane; k < r; k += 32) {
                    float2 t = row[k];
                    rank2(t, vi, wi, vprev[k], wv[k]);
                    row[k] = t;
                    if (i == j + 1) nrow[k] = t;
                    cfma(acc, t, vnew[k]);
                }
                acc.x = tau * warp_sum(acc.x);
                acc.y = tau * warp_sum(acc.y);
                const float2 v = vnew[i];
                kacc += v.x * acc.x + v.y * acc.y;
                if (lane == 0) pv[i] = acc;
            }
        } else {
            // alpha_s = w_s^H v, beta_s = v_s^H v of pending pair s: one warp each, taken from the end of the warp order that
            // gets no block in the last (incomplete) round of the walk below (same arithmetic whichever warp does it)
            const int ps = ((((nblk - (j + 1) / TR) - 1) / SD_WARPS) & 1) ? warp : SD_WARPS - 1 - warp;
            if (!upd && ps < P) {
                float2 aa = make_float2(0.f, 0.f), bb = make_float2(0.f, 0.f);
                for (int k = j + 1 + lane; k < r; k += 32) {
                    const float2 v = vnew[k];
                    const float4 c = VW[ps * WD + k];
                    aa.x = fmaf(c.z, v.x, fmaf(c.w, v.y, aa.x));
                    aa.y = fmaf(c.z, v.y, fmaf(-c.w, v.x, aa.y));
                    bb.x = fmaf(c.x, v.x, fmaf(c.y, v.y, bb.x));
                    bb.y = fmaf(c.x, v.y, fmaf(-c.y, v.x, bb.y));
                }
                aa.x = warp_sum(aa.x), aa.y = warp_sum(aa.y), bb.x = warp_sum(bb.x), bb.y = warp_sum(bb.y);
                if (lane == 0) s_ab[ps] = aa, s_ab[NB + ps] = bb;
            }
            const int jl = (j + 1) & 31;                 // lane that holds column j+1 (in chunk e0)
            const int b0 = (j + 1) / TR;                 // first row block with a live row
            const int L = nblk - b0;
            float2* mypart = part + warp * WD;
            for (int k = e0 * 32 + lane; k < WD; k += 32) mypart[k] = make_float2(0.f, 0.f);
            float kl = 0.f;                              // per lane: 2 Re conj(v_i) rowpart_i - a_ii |v_i|^2 of the rows it met

            // a warp's walk over the tiles of a step whose first live row block / column chunk are wb0 / we0 (wL blocks).
            // The blocks are dealt out from the LONGEST (bottom) one, serpentine-wise, so that the incomplete last round holds
            // the shortest blocks: dealt from the top, the last round gave a few warps the longest blocks on top of their
            // share and the others waited at the barrier (tiles of the busiest warp summed over the steps, r = 256: 1526 ->
            // 1174 against 935 for a perfect split; r = 512: 8541 -> 6941 against 6462)
            auto first_of = [&](int wb0, int we0, int wL) {
                TileIt t;
                t.rho = 0, t.I = wb0 + wL - 1 - warp, t.K = we0, t.valid = warp < wL;
                t.Klast = (t.I * TR + TR - 1) >> 5;
                return t;
            };
            auto advance_of = [&](TileIt t, int wb0, int we0, int wL) {
                if (t.K < t.Klast) {
                    ++t.K;
                    return t;
                }
                ++t.rho;
                const int u = t.rho * SD_WARPS + ((t.rho & 1) ? SD_WARPS - 1 - warp : warp);
                // (u grows with the round, so the first block past the live range ends this warp's walk)
                t.valid = u < wL;
                t.I = wb0 + wL - 1 - u;
                t.K = we0;
                t.Klast = (t.I * TR + TR - 1) >> 5;
                return t;
            };
            auto first = [&]() { return first_of(b0, e0, L); };
            auto advance = [&](TileIt t) { return advance_of(t, b0, e0, L); };
            auto prefetch = [&](const TileIt& t) {
                if (lane < 2 * TR) prefetch_l2(M + (size_t)(t.I * TR + (lane >> 1)) * ld + t.K * 32 + (lane & 1) * 16);
            };
            auto load = [&](const TileIt& t, float2 (&x)[TR]) {
                const int k = t.K * 32 + lane;
                const float2* src = M + (size_t)(t.I * TR) * ld + k;
#pragma unroll
                for (int rr = 0; rr < TR; ++rr) {
                    if (RAGGED) x[rr] = (t.I * TR + rr < r && k < r) ? src[(size_t)rr * ld] : make_float2(0.f, 0.f);
                    else if (LDC) x[rr] = src[rr * LDC];
                    else x[rr] = src[(size_t)rr * ld];
                }
            };
            constexpr bool CACHE_V = TR == 8;            // (16 cached rows spill at 128 registers)
            float2 racc[TR], vrow[CACHE_V ? TR : 1];     // row parts and v_i of the current block
#pragma unroll
            for (int rr = 0; rr < TR; ++rr) racc[rr] = make_float2(0.f, 0.f);
            int vblock = -1;
            auto process = [&](const TileIt& t, float2 (&x)[TR]) {
                const int i0 = t.I * TR;
                const int k = t.K * 32 + lane;
                const bool last = t.K == t.Klast;
                if (CACHE_V && t.I != vblock) {
                    vblock = t.I;
#pragma unroll
                    for (int rr = 0; rr < TR; ++rr) vrow[rr] = vnew[i0 + rr];   // i0 + rr < WD; zero at and left of j, beyond r
                }
                if (upd) {
#pragma unroll 1
                    for (int s2 = 0; s2 < NB; ++s2) {
                        const float4 c = VW[s2 * WD + k];
#pragma unroll
                        for (int rr = 0; rr < TR; ++rr) rank2p(x[rr], VW[s2 * WD + i0 + rr], c);
                    }
                    float2* dst = M + (size_t)i0 * ld + k;
#pragma unroll
                    for (int rr = 0; rr < TR; ++rr) {
                        bool ok = !last || k <= i0 + rr;
                        if (RAGGED) ok = ok && (i0 + rr < r) && (k < r);
                        if (ok) dst[LDC ? (size_t)(rr * LDC) : (size_t)rr * ld] = x[rr];
                    }
                }
                const float2 vk = vnew[k];               // zero at and left of column j
                float2 yk = make_float2(0.f, 0.f);
                if (!last) {
                    if (t.K == e0 && lane == jl) {
#pragma unroll
                        for (int rr = 0; rr < TR; ++rr) nrow[i0 + rr] = make_float2(x[rr].x, -x[rr].y);   // column j+1
                    }
#pragma unroll
                    for (int rr = 0; rr < TR; ++rr) {
                        const float2 vi = CACHE_V ? vrow[CACHE_V ? rr : 0] : vnew[i0 + rr];
                        cfma(racc[rr], x[rr], vk);
                        // strictly below the diagonal: (A v)_k += conj(a_ik) v_i
                        yk.x = fmaf(x[rr].x, vi.x, fmaf(x[rr].y, vi.y, yk.x));
                        yk.y = fmaf(x[rr].x, vi.y, fmaf(-x[rr].y, vi.x, yk.y));
                    }
                } else {
                    // the tile with the diagonal of the block: k <= i only; k == i feeds the diagonal sum
#pragma unroll
                    for (int rr = 0; rr < TR; ++rr) {
                        const int i = i0 + rr;
                        const float2 vi = CACHE_V ? vrow[CACHE_V ? rr : 0] : vnew[i];
                        float2 xx = k <= i ? x[rr] : make_float2(0.f, 0.f);
                        if (RAGGED && i >= r) xx = make_float2(0.f, 0.f);
                        if (t.K == e0 && lane == jl) nrow[i] = make_float2(xx.x, -xx.y);
                        cfma(racc[rr], xx, vk);
                        if (k == i) {
                            kl = fmaf(-xx.x, vi.x * vi.x + vi.y * vi.y, kl);
                            xx = make_float2(0.f, 0.f);
                        }
                        yk.x = fmaf(xx.x, vi.x, fmaf(xx.y, vi.y, yk.x));
                        yk.y = fmaf(xx.x, vi.y, fmaf(-xx.y, vi.x, yk.y));
                    }
                }
                {
                    float2 pk = mypart[k];
                    pk.x += yk.x, pk.y += yk.y;
                    mypart[k] = pk;
                }
                if (last) {
                    // the block is complete: reduce its row parts over the lanes
                    reduce_scatter<TR>(racc, lane);
                    const int i = i0 + (lane >> (TR == 16 ? 1 : 2));
                    if ((lane & (TR == 16 ? 1 : 3)) == 0 && i < r) {
                        const float2 a = racc[0];
                        pv[i] = a;                       // unscaled, uncorrected row part (unused for i <= j)
                        const float2 v = vnew[i];
                        kl = fmaf(2.f, v.x * a.x + v.y * a.y, kl);
                    }
#pragma unroll
                    for (int rr = 0; rr < TR; ++rr) racc[rr] = make_float2(0.f, 0.f);
                }
            };
            TileIt tp = first();                         // PF: the tile pfd loads ahead
            if (PF)
                for (int q = 0; q < pfd && tp.valid; ++q) tp = advance(tp);
            auto ahead = [&]() {
                if (PF && tp.valid) {
                    prefetch(tp);
                    tp = advance(tp);
                }
            };
            if (DB) {
                float2 xa[TR], xb[TR];
                TileIt t0 = first(), t1;
                if (t0.valid) {
                    load(t0, xa);
                    while (true) {
                        t1 = advance(t0);
                        if (t1.valid) load(t1, xb);
                        process(t0, xa);
                        if (!t1.valid) break;
                        t0 = advance(t1);
                        if (t0.valid) load(t0, xa);
                        process(t1, xb);
                        if (!t0.valid) break;
                    }
                }
            } else {
                float2 xa[TR];
                for (TileIt t0 = first(); t0.valid; t0 = advance(t0)) {
                    load(t0, xa);
                    ahead();
                    process(t0, xa);
                }
            }
            if (PF && j + 3 < r) {
                // the first tiles of this warp's walk of step j + 1
                const int nb0 = (j + 2) / TR, ne0 = (j + 2) >> 5, nL = nblk - nb0;
                TileIt tn = first_of(nb0, ne0, nL);
                for (int q = 0; q < pfd && tn.valid; ++q) {
                    prefetch(tn);
                    tn = advance_of(tn, nb0, ne0, nL);
                }
            }
            kacc = warp_sum(kl);
        }
        VK_CK(2);
        if (lane == 0) s_kpart[warp] = kacc;
        __syncthreads();
        VK_CK(3);
        float kk = 0.f;
#pragma unroll
        for (int w = 0; w < SD_WARPS; ++w) kk += s_kpart[w];
        if (resident) {
            // K = tau/2 * v^H p ;  w = p - K v      (v^H p is real for a Hermitian block)
            const float K = 0.5f * tau * kk;
            for (int k = j + 1 + tid; k < r; k += SD_THREADS) {
                const float2 v = vnew[k], p = pv[k];
                wv[k] = make_float2(p.x - K * v.x, p.y - K * v.y);
            }
            float2* t = vprev;
            vprev = vnew;
            vnew = t;
        } else {
            // p = tau (A0 v - sum_s [v_s alpha_s + w_s beta_s]) ; v^H p = tau (v^H A0 v - 2 Re sum_s conj(beta_s) alpha_s)
            const int np = upd ? 0 : P;
            for (int s2 = 0; s2 < np; ++s2) {
                const float2 al = s_ab[s2], be = s_ab[NB + s2];
                kk -= 2.f * (be.x * al.x + be.y * al.y);
            }
            const float K = 0.5f * tau * tau * kk;
            const int slot = upd ? 0 : P;
            for (int k = j + 1 + tid; k < r; k += SD_THREADS) {
                float2 p = pv[k];
#pragma unroll
                for (int w = 0; w < SD_WARPS; ++w) {
                    const float2 c = part[w * WD + k];
                    p.x += c.x, p.y += c.y;
                }
                for (int s2 = 0; s2 < np; ++s2) {
                    const float2 al = s_ab[s2], be = s_ab[NB + s2];
                    const float4 c = VW[s2 * WD + k];
                    p.x -= c.x * al.x - c.y * al.y + c.z * be.x - c.w * be.y;
                    p.y -= c.x * al.y + c.y * al.x + c.z * be.y + c.w * be.x;
                }
                const float2 v = vnew[k];
                VW[slot * WD + k] = make_float4(v.x, v.y, tau * p.x - K * v.x, tau * p.y - K * v.y);
            }
            P = upd ? 1 : P + 1;
        }
        __syncthreads();
        VK_CK(4);
    }
#ifdef VK_TRIDIAG_CLOCKS
    if (tid % 32 == 0 && b == 0)
        printf("[tridiag clocks] warp %d: row+norm %lld  scalar %lld  own pass %lld  barrier wait %lld  update %lld\n", warp, ck[0], ck[1],
               ck[2], ck[3], ck[4]);
#endif
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
                x00 = M[(size_t)j * ld + j], x01 = M[(size_t)(j + 1) * ld + j], x11 = M[(size_t)(j + 1) * ld + j + 1];
                x01.y = -x01.y;
            }
            if (r >= 3) {
                if (j0 >= 0) {
                    const float2 v0 = vprev[j], v1 = vprev[j + 1], w0 = wv[j], w1 = wv[j + 1];
                    x00.x -= 2.f * (v0.x * w0.x + v0.y * w0.y);
                    x11.x -= 2.f * (v1.x * w1.x + v1.y * w1.y);
                    rank2(x01, v0, w0, v1, w1);
                } else {
                    for (int s2 = 0; s2 < P; ++s2) {
                        const float4 c0 = VW[s2 * WD + j], c1 = VW[s2 * WD + j + 1];
                        x00.x -= 2.f * (c0.x * c0.z + c0.y * c0.w);
                        x11.x -= 2.f * (c1.x * c1.z + c1.y * c1.w);
                        rank2p(x01, c0, c1);
                    }
                }
            }
            d[j] = x00.x;
            d[j + 1] = x11.x;
            const float ea = sqrtf(x01.x * x01.x + x01.y * x01.y);  // sub-diagonal element is conj(x01)
            e[j] = ea;
            if (ea > 0.f) phase = cmulf2(phase, make_float2(x01.x / ea, -x01.y / ea));
            ph[j + 1] = phase;
            taus[j] = 0.f;
        }
        e[r - 1] = 0.f;
        taus[r - 1] = 0.f;
    }
}

template <int EPL, int TR, int NB, int NT, bool DB, bool RAGGED, int LDC, bool PF = false>
int launch_symdefer(vk_context* h, cudaStream_t st, float2* W, int B, int r, int ld, size_t wstride, float* d, float* e,
                    float* tau, float2* ph) {
    constexpr int WD = EPL * 32;
    const size_t budget = NT == 512 ? (size_t)VK_SMEM_BUDGET : (size_t)110 * 1024;   // two CTAs per SM
    const size_t vec = (size_t)(5 + 2 * NB) * WD * sizeof(float2);
    const size_t partb = (size_t)(NT / 32) * WD * sizeof(float2);
    int nts = budget > vec ? (int)sqrt((double)(budget - vec) / sizeof(float2)) : 0;
    if (nts > r) nts = r;
    // the resident steps only pay for the last few dozen rows (measured at r = 256: 155 rows 1.50, 128: 1.51, 96: 1.47,
    // 64: 1.46, 32: 1.48, none: 1.52 ms; no difference at r = 512); "tridiag_nts" overrides the cap
    const int cap = h->tridiag_nts >= 0 ? h->tridiag_nts : 64;
    if (cap < nts) nts = cap;
    size_t tb = (size_t)nts * nts * sizeof(float2);
    if (tb < partb) tb = partb;
    const size_t smem = vec + tb;
    auto kern = tridiag_symdefer_kernel<EPL, TR, NB, NT, DB, RAGGED, LDC, PF>;
    VK_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, NT, smem, st>>>(W, r, ld, wstride, d, e, tau, ph, nts, h->tridiag_pf);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

// variant: 0 = one matrix per SM (512 threads, 8-row tiles, double buffer), 1 = two per SM (256 threads, 16-row tiles)
template <int EPL>
int launch_symdefer_r(vk_context* h, cudaStream_t st, float2* W, int B, int r, int ld, size_t wstride, float* d, float* e,
                      float* tau, float2* ph, int variant) {
    const bool ragged = (r % 32) != 0;
    const bool full = !ragged && ld == EPL * 32;   // the usual case: r = ld = 256 / 384 / 512
    if (variant == 1) {
        if (ragged) return launch_symdefer<EPL, 16, 6, 256, false, true, 0>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
        if (full && EPL == 16 && h->tridiag_pf > 0 && B > 2 * h->num_sms)   // more triangles than the L2 holds
            return launch_symdefer<EPL, 16, 6, 256, false, false, EPL * 32, true>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
        if (full) return launch_symdefer<EPL, 16, 6, 256, false, false, EPL * 32>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
        return launch_symdefer<EPL, 16, 6, 256, false, false, 0>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    }
    if (ragged) return launch_symdefer<EPL, 8, 8, 512, true, true, 0>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    if (full) return launch_symdefer<EPL, 8, 8, 512, true, false, EPL * 32>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
    return launch_symdefer<EPL, 8, 8, 512, true, false, 0>(h, st, W, B, r, ld, wstride, d, e, tau, ph);
}

}  // namespace

bool vk_tridiag_symdefer_supported(int r) { return r > 128 && r <= 512; }

int vk_launch_tridiag_symdefer(vk_context* h, cudaStream_t st, float2* W, int B, int r, int ld, size_t wstride, float* d,
                               float* e, float* tau, float2* ph) {
    // (three or four CTAs of 256 threads per SM with 8-row tiles, 80 / 64 registers: 1.35 / 1.7-1.8 ms per KAT-7 cube with
    // 6 to 12 concurrent handles against 1.26 ms; two per SM with 8-row tiles, which balance better (94 % against 80 %), with
    // or without the register double buffer: 1.25-1.27 against 1.18 ms, MeerKAT shard compress 179-186 against 171 ms)
    // (prefetch.global.L1 of the next tile at r = 256, two per SM: 1.22-1.24 against 1.18 ms; deferral depth 4 / 6 / 8 of
    // that shape: 1.19-1.22 / 1.18-1.19 / 1.18-1.19 ms per cube with six handles, 2.17 / 2.18 / 2.19 ms alone)
    // two matrices per SM once there are more matrices than SMs, or when three or more host threads are feeding this GPU
    // through their own handles (KAT-7 cube, 112 matrices: alone 1.58 vs 2.05 ms, but three concurrent handles reach 1.57
    // instead of 1.67 ms per cube because the cubes' kernels can share SMs); "tridiag_variant": 1 / 2 force one / two per SM
    int variant = (B > h->num_sms || vk_concurrent_compress(h->device) >= 3) ? 1 : 0;
    if (h->tridiag_variant == 1) variant = 0;
    if (h->tridiag_variant == 2) variant = 1;
    if (r <= 256) return launch_symdefer_r<8>(h, st, W, B, r, ld, wstride, d, e, tau, ph, variant);
    if (r <= 384) return launch_symdefer_r<12>(h, st, W, B, r, ld, wstride, d, e, tau, ph, variant);
    return launch_symdefer_r<16>(h, st, W, B, r, ld, wstride, d, e, tau, ph, variant);
}
