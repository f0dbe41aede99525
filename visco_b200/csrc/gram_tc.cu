// Batched complex Gram product W = (A A^H)^T on the 5th-generation tensor cores (north_star item (a)).
//
//   W[b][i][t] = sum_v A[b][t][v] * conj(A[b][i][v])          i, t < m ; v < n ; complex64, row-major
//
// replaces the O(m^2 n) part of the LAPACK cgesdd call under da.linalg.svd (reference visco/compress_ms.py:350).
//
// Formulation. View A[b] as a REAL m x 2n matrix R (re, im interleaved — exactly how complex64 sits in memory):
//   Re W[i][t] = R_i . R_t                       Im W[i][t] = R_i . Q_t ,   Q_t[2v] = R_t[2v+1], Q_t[2v+1] = -R_t[2v]
// so one real MMA with the stacked B operand [R_J ; Q_J] (N = 256) yields the real part in TMEM columns 0..127 and
// the imaginary part in columns 128..255 of a 128 x 128 complex output tile. No tensor-core format carries 24
// mantissa bits, so each fp32 operand is split hi + lo (both TF32, round-to-nearest) and the product is formed as
// hi*hi + hi*lo + lo*hi ("3xTF32"). The tensor core adds into its fp32 accumulator with truncation (measured: a
// relative bias of ~3e-8 per MMA, 2.5e-5 after the 768 MMAs of n = 1024), so accumulation is two-level: MMAs chain for
// only CHUNK_KB K-blocks (48 MMAs) into one of two ping-pong TMEM accumulators, and the finished chunk is promoted
// into fp32 registers with round-to-nearest adds while the tensor core fills the other buffer.
//
// Kernel: one CTA per upper-triangular 128 x 128 tile pair (I <= J) of one matrix; the mirrored tile is written as the
// conjugate transpose. Warp roles: warp 0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_64B boxes of 128 rows x 16
// floats), warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator, warps 4..11 = converters (split raw fp32 tiles into
// hi/lo and the swapped copy, in place, swizzle-agnostic because the split is element-wise within 16-byte chunks)
// and promoters (tcgen05.ld of finished chunks -> register accumulators -> global at the end). The converter warps
// raise their register budget with setmaxnreg (128 accumulators per thread); the other warps give theirs up.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int TILE = 128;             // rows of a tile (both I and J)
// K-blocks of 16 floats (one SWIZZLE_64B row) in a four-stage ring: three loads stay in flight behind the block being
// multiplied (with 32-float blocks only two 96 KiB stages fit, and a stage goes load -> split -> MMA strictly in turn)
constexpr int KB_FLOATS = 16;
constexpr int NSTAGE = 4;
constexpr uint32_t A_TILE_BYTES = TILE * KB_FLOATS * 4;          // 8 KiB
constexpr uint32_t B_TILE_BYTES = 2 * TILE * KB_FLOATS * 4;      // 16 KiB (R_J rows then Q_J rows)
constexpr uint32_t OFF_A_HI = 0, OFF_A_LO = A_TILE_BYTES, OFF_B_HI = 2 * A_TILE_BYTES,
                   OFF_B_LO = 2 * A_TILE_BYTES + B_TILE_BYTES;
constexpr uint32_t STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;  // 48 KiB
constexpr uint32_t OFF_BARS = NSTAGE * STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BARS + 256 + 1024;  // + barriers + alignment slack
constexpr int NUM_CONVERTERS = 256;
constexpr int NUM_THREADS = 128 + NUM_CONVERTERS;
constexpr uint32_t TMEM_COLS = 512;  // two 256-column accumulators (re | im), ping-pong
constexpr int CHUNK_KB = 8;           // K-blocks per TMEM accumulation chain (8 x 6 = 48 MMAs)
constexpr int DRAIN_LAG = 3;          // a chunk is promoted this many K-blocks after its last block was converted

// (r0, i0, r1, i1) -> (i0, -r0, i1, -r1)
__device__ __forceinline__ float4 swap_neg(float4 v) { return make_float4(v.y, -v.x, v.w, -v.z); }

// Promote one finished accumulation chunk: TMEM (this warp's 32 lanes x 64 re + 64 im columns) -> += registers.
__device__ __forceinline__ void drain_chunk(int c, uint32_t bar_accf, uint32_t bar_acce, uint32_t tmem_base, int quad,
                                            int chalf, float (&acc_re)[64], float (&acc_im)[64]) {
    const int p = c & 1;
    mbar_wait(bar_accf + 8 * p, ((uint32_t)c >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(p * 256 + chalf * 64);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t re[16], im[16];
        tmem_ld16(taddr + g * 16, re);
        tmem_ld16(taddr + 128 + g * 16, im);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            acc_re[g * 16 + j] += __uint_as_float(re[j]);
            acc_im[g * 16 + j] += __uint_as_float(im[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(bar_acce + 8 * p);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap mapW, int tma_store,
               float2* __restrict__ W, int m, int n2, int tiles_per_mat, int T) {
    extern __shared__ unsigned char smem_raw[];
    // swizzled operands need 1024-byte alignment
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + OFF_BARS;
    // barrier slots (8 bytes each): raw_full[s] @0,8 ; conv_full[s] @16,24 ; empty[s] @32,40 ; acc_full[p] @48,56 ;
    // acc_empty[p] @64,72 ; tmem ptr @96
    const uint32_t bar_raw = bars, bar_conv = bars + 8 * NSTAGE, bar_empty = bars + 16 * NSTAGE, bar_accf = bars + 24 * NSTAGE,
                   bar_acce = bars + 24 * NSTAGE + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BARS + 24 * NSTAGE + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / tiles_per_mat;
    int tile = blockIdx.x - b * tiles_per_mat;
    // decode upper-triangular tile index -> (I, J), I <= J
    int I = 0;
    while (tile >= T - I) {
        tile -= T - I;
        ++I;
    }
    const int J = I + tile;
    const int KB = (n2 + KB_FLOATS - 1) / KB_FLOATS;
    // diagonal tile: the A operand IS the first half of the B operand (R_J) - it is neither loaded nor converted a second
    // time (64 of the 304 KB of shared-memory traffic per K-block, which is what bounds this kernel)
    const bool diag = I == J;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(bar_raw + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, NUM_CONVERTERS);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int p = 0; p < 2; ++p) {
            mbar_init(bar_accf + 8 * p, 1);
            mbar_init(bar_acce + 8 * p, NUM_CONVERTERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int NC = (KB + CHUNK_KB - 1) / CHUNK_KB;  // accumulation chunks

    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
      if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % NSTAGE;
                const uint32_t use = kb / NSTAGE;
                mbar_wait(bar_empty + 8 * s, (use & 1) ^ 1);
                const uint32_t st = sbase + s * STAGE_BYTES;
                mbar_arrive_expect_tx(bar_raw + 8 * s, diag ? A_TILE_BYTES : 2 * A_TILE_BYTES);
                if (!diag) tma_load_3d(st + OFF_A_HI, &tmap, bar_raw + 8 * s, kb * KB_FLOATS, I * TILE, b);
                tma_load_3d(st + OFF_B_HI, &tmap, bar_raw + 8 * s, kb * KB_FLOATS, J * TILE, b);
            }
        }
      } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = tf32, K-major both, N = 256, M = 128
            const uint32_t idesc = idesc_tf32(128, 256, false);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % NSTAGE;
                const uint32_t use = kb / NSTAGE;
                const int c = kb / CHUNK_KB, p = c & 1;
                const bool first = (kb % CHUNK_KB) == 0;
                if (first) {  // the promoters must have drained this accumulator (chunk c - 2)
                    mbar_wait(bar_acce + 8 * p, (((uint32_t)c >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(bar_conv + 8 * s, use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = sbase + s * STAGE_BYTES;
                const uint64_t a_hi = desc_kmajor_sw64(st + (diag ? OFF_B_HI : OFF_A_HI)),
                               a_lo = desc_kmajor_sw64(st + (diag ? OFF_B_LO : OFF_A_LO));
                const uint64_t b_hi = desc_kmajor_sw64(st + OFF_B_HI), b_lo = desc_kmajor_sw64(st + OFF_B_LO);
                const uint32_t d = tmem_base + (uint32_t)p * 256u;
#pragma unroll
                for (int k = 0; k < KB_FLOATS / 8; ++k) {
                    const uint64_t adv = (uint64_t)(k * 32 >> 4);  // 8 tf32 = 32 bytes per MMA along K
                    umma_tf32(d, a_lo + adv, b_hi + adv, idesc, !(first && k == 0));
                    umma_tf32(d, a_hi + adv, b_lo + adv, idesc, 1);
                    umma_tf32(d, a_hi + adv, b_hi + adv, idesc, 1);
                }
                umma_commit(bar_empty + 8 * s);  // frees the stage when these MMAs have read it
                if ((kb % CHUNK_KB) == CHUNK_KB - 1 || kb == KB - 1) umma_commit(bar_accf + 8 * p);  // chunk complete
            }
        }
      }
    } else {
        // ===================== converters =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int ct = threadIdx.x - 128;
        const int quad = warp & 3;            // TMEM lane quadrant this warp may access
        const int chalf = (warp - 4) >> 2;    // which 64-column half of the tile this warp owns
        float acc_re[64], acc_im[64];         // this thread's row x 64 complex columns, fp32 RN accumulation
#pragma unroll
        for (int j = 0; j < 64; ++j) acc_re[j] = acc_im[j] = 0.f;
        int next_drain = 0;

        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % NSTAGE;
            const uint32_t use = kb / NSTAGE;
            mbar_wait(bar_raw + 8 * s, use & 1);
            const uint32_t st = sbase + s * STAGE_BYTES;
            constexpr int CH = A_TILE_BYTES / 16;  // 512 chunks per 128-row tile
#pragma unroll
            for (int i = 0; i < CH / NUM_CONVERTERS; ++i) {
                const uint32_t o = (uint32_t)(ct + i * NUM_CONVERTERS) * 16;
                if (!diag) {
                    const Split4 sa = split4(lds128(st + OFF_A_HI + o));
                    sts128(st + OFF_A_HI + o, sa.hi);
                    sts128(st + OFF_A_LO + o, sa.lo);
                }
                const Split4 sb = split4(lds128(st + OFF_B_HI + o));
                sts128(st + OFF_B_HI + o, sb.hi);
                sts128(st + OFF_B_LO + o, sb.lo);
                sts128(st + OFF_B_HI + A_TILE_BYTES + o, swap_neg(sb.hi));
                sts128(st + OFF_B_LO + A_TILE_BYTES + o, swap_neg(sb.lo));
            }
            // make the generic-proxy writes visible to the tensor core (async proxy), then signal
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_conv + 8 * s);
            // promote a finished chunk DRAIN_LAG K-blocks after its last block was converted: its MMAs have mostly retired
            // by then, so this wait does not stall the conversion stream
            if (kb >= CHUNK_KB + DRAIN_LAG && ((kb - DRAIN_LAG) % CHUNK_KB) == 0) drain_chunk(next_drain++, bar_accf, bar_acce, tmem_base, quad, chalf, acc_re, acc_im);
        }
        while (next_drain < NC) drain_chunk(next_drain++, bar_accf, bar_acce, tmem_base, quad, chalf, acc_re, acc_im);

        // ===================== epilogue: registers -> global =====================
        const int i_loc = quad * 32 + lane;   // row of the tile = TMEM lane
        const int gi = I * TILE + i_loc;
        float2* Wb = W + (size_t)b * m * m;
        if (tma_store) {
            // the tile itself: staged per warp in the (now idle) operand stages - four boxes of 32 rows x 128 bytes,
            // SWIZZLE_128B - and stored by the TMA unit, clipped at the ragged edge of m (from the registers a warp
            // instruction touched 32 rows with 8 bytes each)
            const uint32_t stg = sbase + (uint32_t)(warp - 4) * 16384u + (uint32_t)lane * 128u;
#pragma unroll
            for (int j = 0; j < 64; j += 2) {
                const int box = j >> 4, cc = (j & 15) >> 1;
                sts128(stg + (uint32_t)box * 4096u + (uint32_t)((cc ^ (lane & 7)) << 4),
                       make_float4(acc_re[j], acc_im[j], acc_re[j + 1], acc_im[j + 1]));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t wbase = sbase + (uint32_t)(warp - 4) * 16384u;
                const int c0 = 2 * (J * TILE + chalf * 64), c1 = I * TILE + quad * 32;
#pragma unroll
                for (int box = 0; box < 4; ++box) tma_store_3d(&mapW, wbase + box * 4096u, c0 + box * 32, c1, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (gi < m) {
            if (!tma_store) {
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const int gt = J * TILE + chalf * 64 + j;
                    if (gt < m) Wb[(size_t)gi * m + gt] = make_float2(acc_re[j], acc_im[j]);
                }
            }
            if (I != J) {
                // mirrored tile: W[t][i] = conj(W[i][t]); lanes run along i -> coalesced
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const int gt = J * TILE + chalf * 64 + j;
                    if (gt < m) Wb[(size_t)gt * m + gi] = make_float2(acc_re[j], -acc_im[j]);
                }
            }
        }
        if (tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem outlives the reads
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace

bool vk_gram_tc_supported(int m, int n, int side) {
    // wide matrices only (G = A A^H, contraction along the contiguous channel axis); TMA needs 16-byte row strides
    return side == 0 && m <= n && m > 64 && (n % 2) == 0 && m <= VK_MAX_R;
}

int vk_launch_gram_tc(vk_context* h, const float2* A, int B, int m, int n, float2* W) {
    PFN_encodeTiled encode = get_encode_tiled();
    if (!encode) return vk_fail(h, VK_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    const int T = (m + TILE - 1) / TILE;
    const int tiles = T * (T + 1) / 2;
    VK_CUDA(h, cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    // the tensor map's outermost dimension is the batch; keep it within what one map may hold and one grid may cover
    const int maxB = 32768;
    for (int b0 = 0; b0 < B; b0 += maxB) {
        const int nb = (B - b0) < maxB ? (B - b0) : maxB;
        CUtensorMap tmap;
        const cuuint64_t dims[3] = {(cuuint64_t)2 * n, (cuuint64_t)m, (cuuint64_t)nb};
        const cuuint64_t strides[2] = {(cuuint64_t)2 * n * sizeof(float), (cuuint64_t)m * 2 * n * sizeof(float)};
        const cuuint32_t box[3] = {KB_FLOATS, TILE, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                                  const_cast<float2*>(A + (size_t)b0 * m * n), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return vk_fail(h, VK_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
        // output as a 3-D tensor (2m floats, m rows, nb matrices) for the bulk stores of the epilogue (needs 16-byte rows)
        CUtensorMap mapW = tmap;
        const int tma_store = (m % 2 == 0) && (reinterpret_cast<uintptr_t>(W) % 16 == 0);
        if (tma_store) {
            const cuuint64_t wd[3] = {(cuuint64_t)2 * m, (cuuint64_t)m, (cuuint64_t)nb};
            const cuuint64_t ws[2] = {(cuuint64_t)m * 8, (cuuint64_t)m * m * 8};
            const cuuint32_t wb[3] = {32, 32, 1};
            const CUresult rw = encode(&mapW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, W + (size_t)b0 * m * m, wd, ws, wb, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rw != CUDA_SUCCESS)
                return vk_fail(h, VK_ECUDA, "cuTensorMapEncodeTiled (Gram output) failed with code " + std::to_string((int)rw));
        }
        gram_tc_kernel<<<(unsigned)(nb * tiles), NUM_THREADS, SMEM_BYTES, h->stream>>>(
            tmap, mapW, tma_store, W + (size_t)b0 * m * m, m, 2 * n, tiles, T);
        VK_LAUNCH_CHECK(h);
    }
    return VK_OK;
}
