// Batched complex GEMM on tcgen05 for the two large-rank products of the path (north_star items (c) and (d) at k > 8):
//
//   MODE_FORMV:  Vt[c][v]  = sum_t  X[c][t] * A[t][v]        X = conj(U_k)^T / lambda (materialised, [B][kmax][m])
//   MODE_RECON:  out[t][v] = sum_c  U[t][c] * (S[c] Vt[c][v])
//
// i.e. D[M x N] = P[M x K] * Q[K x N] with P K-major (its contraction index is contiguous in memory) and Q N-major
// (its output index is contiguous). Replaces the BLAS cgemm behind `U_mult_S @ Vt` (reference decompress_ms.py:131)
// and the right-vector part of LAPACK cgesdd (reference compress_ms.py:350).
//
// Complex arithmetic on a real tensor core: the A operand is the real view of P (row i = [pr(i,0), pi(i,0), pr(i,1), ...],
// 2K floats, K-major). The B operand is a 2K x 2N real matrix built on the fly in shared memory from the rows of Q:
// contraction row 2k = the raw row (qr, qi interleaved along N), row 2k+1 = (-qi, qr), so that D's real view comes out
// directly as interleaved complex64:  D[i][2v] = sum pr*qr - pi*qi ,  D[i][2v+1] = sum pr*qi + pi*qr.
// Both operands are split hi + lo (TF32, round-to-nearest) and multiplied as hi*hi + hi*lo + lo*hi (3xTF32); MMA chains are
// CHUNK_KB K-blocks long into ping-pong TMEM accumulators that the converter warps promote into fp32 registers with
// round-to-nearest adds (see gram_tc.cu for why).
//
// One CTA computes a 128 x 128 complex tile. Warp roles: warp 0 TMA producer (A operand: 3-D map, SWIZZLE_64B boxes of
// 128 rows x 16 floats; B operand: 4-D map over (32 floats, contraction row, 32-float group, batch) so that the box
// {32, 8, 8, 1} lands as 8 groups of 8 rows x 128 bytes), warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..11
// converters (A: split in place; B: un-swizzle the raw row, optionally scale it by S[c], write rows 2k / 2k+1 of the
// MN-major SWIZZLE_128B_BASE32B operand, hi and lo) and promoters / epilogue.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int TILE_M = 128;
constexpr int TILE_NC = 128;              // complex output columns per tile = 256 floats = MMA N
// K-blocks of 8 complex contraction elements in a FOUR-stage ring: with blocks of 16 only two 112 KiB stages fit, and a
// stage then goes TMA (1.5-2 us from L2 / HBM) -> conversion -> MMA strictly in turn - the tensor pipe was 38 % busy
// (profiles/r02_ncu_full_c3_gram_cgemm.txt); half-size blocks keep three loads in flight behind the one being multiplied
// (46 %). What bounds the kernel from here on is SHARED-MEMORY BANDWIDTH: per K-block TMA writes 16 KiB, the converters
// read 16 and write 48 KiB, and the six MMAs read 6 x (4 + 8) KiB of operands - 152 KiB = 1190 cycles at 128 B/clk against
// 786 cycles of MMA time; switching loads, conversions, MMAs or stores off one by one shortened a run additively (each
// one's share of that traffic). Measured and dropped on top of this: a ring of eight raw slots feeding two converted slots
// (10.0 vs 10.3 ms on 296 equal-rank reconstructions, but 40.5 vs 37.2-38.3 ms on the MeerKAT shard's ragged ranks), one
// barrier arrival per warp instead of per thread (no change), two converter groups alternating K-blocks (10.7 ms), TMA
// multicast of the A tile inside clusters of 2 / 4 column neighbours (11.2 / 12.5-13.6 ms). The A operand rows are 16
// floats = 64 bytes: SWIZZLE_64B.
constexpr int KB_C = 8;                   // complex contraction elements per K-block (16 floats of A, 16 rows of B)
constexpr int NSTAGE = 4;
constexpr uint32_t A_BYTES = TILE_M * 2 * KB_C * 4;   // 8 KiB  (128 rows x 16 floats)
constexpr uint32_t BRAW_BYTES = 8 * KB_C * 128;   // 8 KiB  (8 groups x 8 rows x 128 B)
constexpr uint32_t BCONV_BYTES = 8 * 2 * KB_C * 128;  // 16 KiB (8 groups x 16 rows x 128 B)
constexpr uint32_t OFF_A_HI = 0, OFF_A_LO = A_BYTES, OFF_B_RAW = 2 * A_BYTES, OFF_B_HI = 2 * A_BYTES + BRAW_BYTES,
                   OFF_B_LO = 2 * A_BYTES + BRAW_BYTES + BCONV_BYTES;
constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + BRAW_BYTES + 2 * BCONV_BYTES;  // 56 KiB
constexpr uint32_t OFF_BARS = NSTAGE * STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BARS + 256 + 1024;
constexpr int NUM_CONVERTERS = 256;
constexpr int NUM_THREADS = 128 + NUM_CONVERTERS;
constexpr uint32_t TMEM_COLS = 512;
constexpr int CHUNK_KB = 8;               // K-blocks per TMEM accumulation chain (8 x 6 = 48 MMAs)
constexpr int DRAIN_LAG = 3;              // a chunk is promoted this many K-blocks after its last block was converted

enum { MODE_FORMV = 0, MODE_RECON = 1, MODE_PLAIN = 2 };

// (r0, i0, r1, i1) -> (-i0, r0, -i1, r1) : the row that multiplies the imaginary part of the A operand
__device__ __forceinline__ float4 rot90(float4 v) { return make_float4(-v.y, v.x, -v.w, v.z); }

// promote one finished chunk: this thread's row x 128 consecutive accumulator columns (64 complex outputs)
__device__ __forceinline__ void drain_chunk(int c, uint32_t bar_accf, uint32_t bar_acce, uint32_t tmem_base, int quad,
                                            int chalf, float (&acc)[128]) {
    const int p = c & 1;
    mbar_wait(bar_accf + 8 * p, ((uint32_t)c >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(p * 256 + chalf * 128);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t a[16], b[16];
        tmem_ld16(taddr + g * 32, a);
        tmem_ld16(taddr + g * 32 + 16, b);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            acc[g * 32 + j] += __uint_as_float(a[j]);
            acc[g * 32 + 16 + j] += __uint_as_float(b[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(bar_acce + 8 * p);
}

struct GemmArgs {
    float2* D;             // output [B][Mtot][ldn] complex
    const int32_t* ranks;  // [B]
    const float* S;        // MODE_RECON: [B][kmax] scale of contraction row c
    float* norm2;          // MODE_FORMV: [B][kmax] accumulates sum_v |D[c][v]|^2
    const float* rowscale; // MODE_PLAIN (optional): [B][Mtot], D[i][:] = rowscale[i] * conj(P Q)[i][:]
    int Mtot;              // rows of D that must be written (kmax for FORMV, m for RECON)
    int Ntot;              // complex columns (n)
    int Ktot;              // complex contraction length available (m for FORMV, kmax for RECON)
    int kmax;
    int tiles_m, tiles_n;
};

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
cgemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapD, const GemmArgs g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + OFF_BARS;
    const uint32_t bar_raw = bars, bar_conv = bars + 8 * NSTAGE, bar_empty = bars + 16 * NSTAGE, bar_accf = bars + 24 * NSTAGE,
                   bar_acce = bars + 24 * NSTAGE + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BARS + 24 * NSTAGE + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = g.tiles_m * g.tiles_n;
    const int b = blockIdx.x / tiles;
    const int tile = blockIdx.x - b * tiles;
    const int m0 = (tile / g.tiles_n) * TILE_M;
    const int n0c = (tile % g.tiles_n) * TILE_NC;  // first complex column
    const int rank = MODE == MODE_PLAIN ? 0 : min(g.ranks ? g.ranks[b] : g.kmax, g.kmax);
    // valid extents for this matrix
    const int Mvalid = MODE == MODE_FORMV ? rank : g.Mtot;
    const int Kvalid = MODE == MODE_RECON ? rank : g.Ktot;
    const int KB = (m0 < Mvalid) ? (Kvalid + KB_C - 1) / KB_C : 0;  // nothing to multiply for all-padding tiles
    const int NC = (KB + CHUNK_KB - 1) / CHUNK_KB;

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
                mbar_arrive_expect_tx(bar_raw + 8 * s, A_BYTES + BRAW_BYTES);
                tma_load_3d(st + OFF_A_HI, &mapA, bar_raw + 8 * s, kb * 2 * KB_C, m0, b);
                tma_load_4d(st + OFF_B_RAW, &mapB, bar_raw + 8 * s, 0, kb * KB_C, n0c * 2 / 32, b);
            }
        }
      } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // D = f32, A = B = tf32, A K-major, B MN-major (bit 16), N = 256, M = 128
            const uint32_t idesc = idesc_tf32(128, 256, true);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % NSTAGE;
                const uint32_t use = kb / NSTAGE;
                const int c = kb / CHUNK_KB, p = c & 1;
                const bool first = (kb % CHUNK_KB) == 0;
                if (first) {
                    mbar_wait(bar_acce + 8 * p, (((uint32_t)c >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(bar_conv + 8 * s, use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = sbase + s * STAGE_BYTES;
                const uint64_t a_hi = desc_kmajor_sw64(st + OFF_A_HI), a_lo = desc_kmajor_sw64(st + OFF_A_LO);
                const uint32_t d = tmem_base + (uint32_t)p * 256u;
#pragma unroll
                for (int k = 0; k < 2 * KB_C / 8; ++k) {  // MMAs of K = 8 floats per K-block of 2 KB_C
                    const uint64_t adv = (uint64_t)(k * 32 >> 4);
                    // B: contraction rows 8k .. 8k+7 of every 32-float group: atom k of each group
                    const uint64_t b_hi = desc_mnmajor_sw128_32b(st + OFF_B_HI + k * 1024, 2 * KB_C * 128, 512);
                    const uint64_t b_lo = desc_mnmajor_sw128_32b(st + OFF_B_LO + k * 1024, 2 * KB_C * 128, 512);
                    umma_tf32(d, a_lo + adv, b_hi, idesc, !(first && k == 0));
                    umma_tf32(d, a_hi + adv, b_lo, idesc, 1);
                    umma_tf32(d, a_hi + adv, b_hi, idesc, 1);
                }
                umma_commit(bar_empty + 8 * s);
                if ((kb % CHUNK_KB) == CHUNK_KB - 1 || kb == KB - 1) umma_commit(bar_accf + 8 * p);
            }
        }
      }
    } else {
        // ===================== converters / promoters / epilogue =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int ct = threadIdx.x - 128;
        const int quad = warp & 3;
        const int chalf = (warp - 4) >> 2;  // which 128-float half of the 256 accumulator columns
        float acc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) acc[j] = 0.f;
        int next_drain = 0;

        // what this thread converts is the same in every K-block: two 16-byte chunks of the A tile, two of the raw B tile
        constexpr int NA = (int)(A_BYTES / 16) / NUM_CONVERTERS, NB = (int)(BRAW_BYTES / 16) / NUM_CONVERTERS;
        int colA[NA], tB[NB];
        uint32_t offA[NA], offBraw[NB], o0B[NB], o1B[NB];
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int c = ct + i * NUM_CONVERTERS;
            offA[i] = (uint32_t)c * 16;
            // four 16-byte chunks per 64-byte row, logical chunk = physical chunk XOR ((row >> 1) & 3) under SWIZZLE_64B
            colA[i] = 2 * ((c & 3) ^ ((c >> 3) & 3));
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int c = ct + i * NUM_CONVERTERS;      // physical 16-byte chunk of the raw tile
            const int grp = c / (KB_C * 8);              // 32-float group (KB_C rows x 8 chunks each)
            const int t = (c >> 3) & (KB_C - 1);         // contraction row inside the K-block
            const int lc = (c & 7) ^ (t & 7);            // logical chunk: TMA swizzled it with the row index
            const int r0 = 2 * t, r1 = 2 * t + 1;        // rows of the 2K x 2N real operand
            tB[i] = t;
            offBraw[i] = (uint32_t)c * 16;
            // destination swizzle (BASE32B): 32-byte chunk index (lc >> 1) XOR (row & 3); 16-byte halves keep their order
            o0B[i] = (uint32_t)(grp * 2 * KB_C + r0) * 128 + (uint32_t)(((((lc >> 1) ^ (r0 & 3)) << 1) | (lc & 1)) << 4);
            o1B[i] = (uint32_t)(grp * 2 * KB_C + r1) * 128 + (uint32_t)(((((lc >> 1) ^ (r1 & 3)) << 1) | (lc & 1)) << 4);
        }
        const float* Sb = MODE == MODE_RECON ? g.S + (size_t)b * g.kmax : nullptr;

        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % NSTAGE;
            const uint32_t use = kb / NSTAGE;
            float sc[NB];
            bool on[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {   // (the scale factors are on their way before the wait)
                const int cc = kb * KB_C + tB[i];
                on[i] = cc < Kvalid;
                sc[i] = (MODE == MODE_RECON && on[i]) ? __ldg(Sb + cc) : 1.f;
            }
            mbar_wait(bar_raw + 8 * s, use & 1);
            const uint32_t st = sbase + s * STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                float4 av = lds128(st + OFF_A_HI + offA[i]);
                if (MODE == MODE_RECON) {
                    // columns of U at or beyond ranks[b] are not part of the product, whatever they hold (the last K-block
                    // may reach past the rank)
                    const int col = kb * KB_C + colA[i];
                    if (col >= Kvalid) av.x = av.y = 0.f;
                    if (col + 1 >= Kvalid) av.z = av.w = 0.f;
                }
                const Split4 sa = split4(av);
                sts128(st + OFF_A_HI + offA[i], sa.hi);
                sts128(st + OFF_A_LO + offA[i], sa.lo);
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                float4 v = lds128(st + OFF_B_RAW + offBraw[i]);
                if (MODE == MODE_RECON) {
                    // modes beyond ranks[b] are never read as data, whatever they hold
                    v = on[i] ? make_float4(v.x * sc[i], v.y * sc[i], v.z * sc[i], v.w * sc[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                // the (-qi, qr) twin row is the same numbers permuted and negated: so are its hi and lo parts (rounding to
                // nearest is symmetric under negation)
                const Split4 s0 = split4(v);
                sts128(st + OFF_B_HI + o0B[i], s0.hi);
                sts128(st + OFF_B_LO + o0B[i], s0.lo);
                sts128(st + OFF_B_HI + o1B[i], rot90(s0.hi));
                sts128(st + OFF_B_LO + o1B[i], rot90(s0.lo));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_conv + 8 * s);
            if (kb >= CHUNK_KB + DRAIN_LAG && ((kb - DRAIN_LAG) % CHUNK_KB) == 0)
                drain_chunk(next_drain++, bar_accf, bar_acce, tmem_base, quad, chalf, acc);
        }
        while (next_drain < NC) drain_chunk(next_drain++, bar_accf, bar_acce, tmem_base, quad, chalf, acc);

        // ---- epilogue: this thread holds row (m0 + quad*32 + lane), complex columns n0c + chalf*64 .. +63 = 512 bytes. Stored
        // straight from the registers every warp instruction touches 32 rows with 16 bytes each (half sectors); instead the
        // warp stages its 32 x 64 block in the (now idle) operand stages - four TMA boxes of 32 rows x 128 bytes,
        // SWIZZLE_128B: a lane writes its own row, the 16-byte chunk index XOR-ed with the row keeps the stores conflict
        // free - and one lane hands them to the TMA unit, which also clips at the ragged ends of M and N ----
        const int gi = m0 + quad * 32 + lane;
        const bool valid_row = gi < Mvalid;
        float nsum = 0.f;
        float rs = 1.f, rsi = 1.f;
        if (MODE == MODE_PLAIN && g.rowscale && gi < g.Mtot) {
            rs = g.rowscale[(size_t)b * g.Mtot + gi];
            rsi = -rs;
        }
        const uint32_t stg = sbase + (uint32_t)(warp - 4) * 16384u + (uint32_t)lane * 128u;
#pragma unroll
        for (int j = 0; j < 64; j += 2) {
            const int v = n0c + chalf * 64 + j;
            float4 o = valid_row ? make_float4(acc[2 * j], acc[2 * j + 1], acc[2 * j + 2], acc[2 * j + 3])
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == MODE_PLAIN) o = make_float4(rs * o.x, rsi * o.y, rs * o.z, rsi * o.w);
            if (MODE == MODE_FORMV) {
                if (v + 1 < g.Ntot) nsum += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
                else if (v < g.Ntot) nsum += o.x * o.x + o.y * o.y;
            }
            const int box = j >> 4, cc = (j & 15) >> 1;      // 16 complex columns per box, two per 16-byte chunk
            sts128(stg + (uint32_t)box * 4096u + (uint32_t)((cc ^ (lane & 7)) << 4), o);
        }
        if (MODE == MODE_FORMV && valid_row && gi < g.Mtot) atomicAdd(&g.norm2[(size_t)b * g.kmax + gi], nsum);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const uint32_t wbase = sbase + (uint32_t)(warp - 4) * 16384u;
            const int c0 = 2 * (n0c + chalf * 64), c1 = m0 + quad * 32;
#pragma unroll
            for (int box = 0; box < 4; ++box) tma_store_3d(&mapD, wbase + box * 4096u, c0 + box * 32, c1, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory outlives the bulk reads
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// A operand: P[b][rows][kc] complex, K-major: real view dims (2*kc, rows, B)
int make_map_a(vk_context* h, CUtensorMap* map, const float2* P, int rows, int kc, int nb) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return vk_fail(h, VK_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    const cuuint64_t dims[3] = {(cuuint64_t)2 * kc, (cuuint64_t)rows, (cuuint64_t)nb};
    const cuuint64_t strides[2] = {(cuuint64_t)2 * kc * 4, (cuuint64_t)rows * 2 * kc * 4};
    const cuuint32_t box[3] = {2 * KB_C, TILE_M, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float2*>(P), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return vk_fail(h, VK_ECUDA, "tensor map (A operand) failed: " + std::to_string((int)r));
    return VK_OK;
}
// output D[b][rows][nc] complex: real view dims (2*nc, rows, B); boxes of 32 rows x 128 bytes
int make_map_d(vk_context* h, CUtensorMap* map, float2* D, int rows, int nc, int nb) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return vk_fail(h, VK_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    const cuuint64_t dims[3] = {(cuuint64_t)2 * nc, (cuuint64_t)rows, (cuuint64_t)nb};
    const cuuint64_t strides[2] = {(cuuint64_t)nc * 8, (cuuint64_t)rows * nc * 8};
    const cuuint32_t box[3] = {32, 32, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, D, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return vk_fail(h, VK_ECUDA, "tensor map (output) failed: " + std::to_string((int)r));
    return VK_OK;
}
// B operand: Q[b][krows][nc] complex, N contiguous: view (32 floats, krows, 2*nc/32 groups, B); box {32, KB_C, 8, 1}
int make_map_b(vk_context* h, CUtensorMap* map, const float2* Q, int krows, int nc, int nb) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return vk_fail(h, VK_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    const cuuint64_t dims[4] = {32, (cuuint64_t)krows, (cuuint64_t)(2 * nc / 32), (cuuint64_t)nb};
    const cuuint64_t strides[3] = {(cuuint64_t)2 * nc * 4, 128, (cuuint64_t)krows * 2 * nc * 4};
    const cuuint32_t box[4] = {32, KB_C, 8, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float2*>(Q), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return vk_fail(h, VK_ECUDA, "tensor map (B operand) failed: " + std::to_string((int)r));
    return VK_OK;
}

}  // namespace

bool vk_cgemm_tc_supported(int m, int n, int kmax) {
    // tensor maps need 16-byte strides (kmax even) and whole 32-float groups along the channel axis (n % 16 == 0)
    return kmax > 8 && (kmax % 2) == 0 && (n % 16) == 0 && m >= 1;
}

// Vt[b][c][:] = sum_t X[b][c][t] A[b][t][:]   (rows c >= ranks[b] are written as zeros); norm2[b][c] += |Vt[b][c]|^2
int vk_launch_formv_tc(vk_context* h, const float2* X, const float2* A, const int32_t* ranks, float2* Vt, float* norm2,
                       int B, int m, int n, int kmax) {
    VK_CUDA(h, cudaFuncSetAttribute(cgemm_tc_kernel<MODE_FORMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const int maxB = 16384;
    for (int b0 = 0; b0 < B; b0 += maxB) {
        const int nb = (B - b0) < maxB ? (B - b0) : maxB;
        CUtensorMap ma, mb, md;
        int rc;
        if ((rc = make_map_a(h, &ma, X + (size_t)b0 * kmax * m, kmax, m, nb))) return rc;
        if ((rc = make_map_b(h, &mb, A + (size_t)b0 * m * n, m, n, nb))) return rc;
        if ((rc = make_map_d(h, &md, Vt + (size_t)b0 * kmax * n, kmax, n, nb))) return rc;
        GemmArgs g;
        g.D = Vt + (size_t)b0 * kmax * n;
        g.ranks = ranks + b0;
        g.S = nullptr;
        g.norm2 = norm2 + (size_t)b0 * kmax;
        g.rowscale = nullptr;
        g.Mtot = kmax;
        g.Ntot = n;
        g.Ktot = m;
        g.kmax = kmax;
        g.tiles_m = (kmax + TILE_M - 1) / TILE_M;
        g.tiles_n = (n + TILE_NC - 1) / TILE_NC;
        cgemm_tc_kernel<MODE_FORMV><<<(unsigned)((long long)nb * g.tiles_m * g.tiles_n), NUM_THREADS, SMEM_BYTES, h->stream>>>(ma, mb, md, g);
        VK_LAUNCH_CHECK(h);
    }
    return VK_OK;
}

// Plain batched product D[b] = P[b] Q[b] (P [B][M][K] K-major, Q [B][K][N], D [B][M][N], all complex64, contiguous), or
// with rowscale != NULL: D[b][i][:] = rowscale[b][i] * conj((P Q)[i][:]). Needs K even and N a multiple of 16.
// Used by the eigenvector stage (tridiag.cu): Newton-Schulz step and back-transformation of the eigenvectors of T.
int vk_launch_cgemm_tc_plain(vk_context* h, const float2* P, const float2* Q, float2* D, const float* rowscale, int B,
                             int M, int N, int K) {
    if ((K % 2) || (N % 16)) return vk_fail(h, VK_EINVAL, "cgemm_tc_plain: needs even K and N % 16 == 0");
    VK_CUDA(h, cudaFuncSetAttribute(cgemm_tc_kernel<MODE_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const int maxB = 16384;
    for (int b0 = 0; b0 < B; b0 += maxB) {
        const int nb = (B - b0) < maxB ? (B - b0) : maxB;
        CUtensorMap ma, mb, md;
        int rc;
        if ((rc = make_map_a(h, &ma, P + (size_t)b0 * M * K, M, K, nb))) return rc;
        if ((rc = make_map_b(h, &mb, Q + (size_t)b0 * K * N, K, N, nb))) return rc;
        if ((rc = make_map_d(h, &md, D + (size_t)b0 * M * N, M, N, nb))) return rc;
        GemmArgs g;
        g.D = D + (size_t)b0 * M * N;
        g.ranks = nullptr;
        g.S = nullptr;
        g.norm2 = nullptr;
        g.rowscale = rowscale ? rowscale + (size_t)b0 * M : nullptr;
        g.Mtot = M;
        g.Ntot = N;
        g.Ktot = K;
        g.kmax = K;
        g.tiles_m = (M + TILE_M - 1) / TILE_M;
        g.tiles_n = (N + TILE_NC - 1) / TILE_NC;
        cgemm_tc_kernel<MODE_PLAIN><<<(unsigned)((long long)nb * g.tiles_m * g.tiles_n), NUM_THREADS, SMEM_BYTES, h->stream>>>(ma, mb, md, g);
        VK_LAUNCH_CHECK(h);
    }
    return VK_OK;
}

// out[b][t][:] = sum_{c < ranks[b]} U[b][t][c] S[b][c] Vt[b][c][:]
int vk_launch_recon_tc(vk_context* h, const float2* U, const float* S, const float2* Vt, const int32_t* ranks, float2* out,
                       int B, int m, int n, int kmax) {
    VK_CUDA(h, cudaFuncSetAttribute(cgemm_tc_kernel<MODE_RECON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const int maxB = 16384;
    for (int b0 = 0; b0 < B; b0 += maxB) {
        const int nb = (B - b0) < maxB ? (B - b0) : maxB;
        CUtensorMap ma, mb, md;
        int rc;
        if ((rc = make_map_a(h, &ma, U + (size_t)b0 * m * kmax, m, kmax, nb))) return rc;
        if ((rc = make_map_b(h, &mb, Vt + (size_t)b0 * kmax * n, kmax, n, nb))) return rc;
        if ((rc = make_map_d(h, &md, out + (size_t)b0 * m * n, m, n, nb))) return rc;
        GemmArgs g;
        g.D = out + (size_t)b0 * m * n;
        g.ranks = ranks ? ranks + b0 : nullptr;
        g.S = S + (size_t)b0 * kmax;
        g.norm2 = nullptr;
        g.rowscale = nullptr;
        g.Mtot = m;
        g.Ntot = n;
        g.Ktot = kmax;
        g.kmax = kmax;
        g.tiles_m = (m + TILE_M - 1) / TILE_M;
        g.tiles_n = (n + TILE_NC - 1) / TILE_NC;
        cgemm_tc_kernel<MODE_RECON><<<(unsigned)((long long)nb * g.tiles_m * g.tiles_n), NUM_THREADS, SMEM_BYTES, h->stream>>>(ma, mb, md, g);
        VK_LAUNCH_CHECK(h);
    }
    return VK_OK;
}
