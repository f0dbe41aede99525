// Batched complex64 SIMT GEMM core (fp32 FMA pipe), parameterised by an "Op" that says how operand elements are
// fetched and where results go. Used for the shapes the tcgen05 kernels do not take: Gram products of small/odd
// matrices, the factor formation V_k = A^H U_k S_k^-1 / U_k = A V_k S_k^-1, and rank-k reconstruction.
//
//   C[b](i, j) = sum_k  opA(b, i, k) * opB(b, k, j)       i < M(b), j < N(b), k < K(b)
//
// One CTA computes a TILE_M x TILE_N tile of one matrix b; a thread owns TR x TC outputs in registers. Operand tiles
// are staged through shared memory as As[kk][i], Bs[kk][j] so the inner loop reads TR + TC contiguous values.
#pragma once
#include "common.cuh"

// Op concept:
//   static constexpr bool A_K_CONTIG / B_K_CONTIG : is k the contiguous index of that operand in global memory
//   static constexpr int REDUCE : 0 none, 1 accumulate sum_j |C(i,j)|^2 per row i, 2 accumulate sum_i |C(i,j)|^2 per col j
//   int M(b), N(b), K(b)      valid extents for matrix b
//   int Mfill(), Nfill()      extents that must be WRITTEN (zeros outside the valid region), also size the grid
//   float2 loadA(b, i, k), loadB(b, k, j)
//   void store(b, i, j, v)
//   void reduce_add(b, idx, v)   (REDUCE != 0)

template <int TILE_M, int TILE_N, int TR, int TC, int KC, class Op>
__global__ void __launch_bounds__((TILE_M / TR) * (TILE_N / TC)) cgemm_kernel(const Op op, int tiles_m, int tiles_n) {
    constexpr int TXN = TILE_N / TC;
    constexpr int TYN = TILE_M / TR;
    constexpr int NT = TXN * TYN;
    constexpr int APAD = Op::A_K_CONTIG ? 1 : 0;
    constexpr int BPAD = Op::B_K_CONTIG ? 1 : 0;
    __shared__ float2 As[KC][TILE_M + APAD];
    __shared__ float2 Bs[KC][TILE_N + BPAD];
    __shared__ float red[(Op::REDUCE == 1) ? TILE_M : ((Op::REDUCE == 2) ? TILE_N : 1)];

    const int tiles = tiles_m * tiles_n;
    const int b = blockIdx.x / tiles;
    const int tile = blockIdx.x - b * tiles;
    const int m0 = (tile / tiles_n) * TILE_M;
    const int n0 = (tile % tiles_n) * TILE_N;
    const int M = op.M(b), N = op.N(b), K = op.K(b);
    const int Mf = op.Mfill(), Nf = op.Nfill();
    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;

    float2 acc[TR][TC];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = make_float2(0.f, 0.f);

    const bool live = (m0 < M) && (n0 < N);
    if (live) {
        for (int k0 = 0; k0 < K; k0 += KC) {
            for (int e = tid; e < KC * TILE_M; e += NT) {
                int kk, i;
                if (Op::A_K_CONTIG) {
                    kk = e % KC;
                    i = e / KC;
                } else {
                    i = e % TILE_M;
                    kk = e / TILE_M;
                }
                float2 v = make_float2(0.f, 0.f);
                if (k0 + kk < K && m0 + i < M) v = op.loadA(b, m0 + i, k0 + kk);
                As[kk][i] = v;
            }
            for (int e = tid; e < KC * TILE_N; e += NT) {
                int kk, j;
                if (Op::B_K_CONTIG) {
                    kk = e % KC;
                    j = e / KC;
                } else {
                    j = e % TILE_N;
                    kk = e / TILE_N;
                }
                float2 v = make_float2(0.f, 0.f);
                if (k0 + kk < K && n0 + j < N) v = op.loadB(b, k0 + kk, n0 + j);
                Bs[kk][j] = v;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                float2 a[TR], bb[TC];
#pragma unroll
                for (int i = 0; i < TR; ++i) a[i] = As[kk][ty * TR + i];
#pragma unroll
                for (int j = 0; j < TC; ++j) bb[j] = Bs[kk][tx * TC + j];
#pragma unroll
                for (int i = 0; i < TR; ++i)
#pragma unroll
                    for (int j = 0; j < TC; ++j) cfma(acc[i][j], a[i], bb[j]);
            }
            __syncthreads();
        }
    }

    if (Op::REDUCE != 0) {
        constexpr int NRED = (Op::REDUCE == 1) ? TILE_M : TILE_N;
        for (int e = tid; e < NRED; e += NT) red[e] = 0.f;
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TR; ++i) {
        const int gi = m0 + ty * TR + i;
#pragma unroll
        for (int j = 0; j < TC; ++j) {
            const int gj = n0 + tx * TC + j;
            if (gi < Mf && gj < Nf) {
                const bool valid = (gi < M) && (gj < N);
                const float2 v = valid ? acc[i][j] : make_float2(0.f, 0.f);
                op.store(b, gi, gj, v);
                if (Op::REDUCE == 1 && valid) atomicAdd(&red[ty * TR + i], v.x * v.x + v.y * v.y);
                if (Op::REDUCE == 2 && valid) atomicAdd(&red[tx * TC + j], v.x * v.x + v.y * v.y);
            }
        }
    }
    if (Op::REDUCE != 0) {
        __syncthreads();
        if (Op::REDUCE == 1) {
            for (int e = tid; e < TILE_M; e += NT)
                if (m0 + e < M && live) op.reduce_add(b, m0 + e, red[e]);
        } else {
            for (int e = tid; e < TILE_N; e += NT)
                if (n0 + e < N && live) op.reduce_add(b, n0 + e, red[e]);
        }
    }
}

template <int TILE_M, int TILE_N, int TR, int TC, int KC, class Op>
static int cgemm_launch(vk_context* h, const Op& op, int B) {
    const int tiles_m = (op.Mfill() + TILE_M - 1) / TILE_M;
    const int tiles_n = (op.Nfill() + TILE_N - 1) / TILE_N;
    const long long nblocks = (long long)B * tiles_m * tiles_n;
    if (nblocks <= 0) return VK_OK;
    if (nblocks > 0x7fffffffLL) return vk_fail(h, VK_EINVAL, "cgemm: grid too large");
    constexpr int NT = (TILE_M / TR) * (TILE_N / TC);
    cgemm_kernel<TILE_M, TILE_N, TR, TC, KC, Op><<<(unsigned)nblocks, NT, 0, h->stream>>>(op, tiles_m, tiles_n);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}
