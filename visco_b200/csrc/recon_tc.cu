// Rank-k reconstruction for 8 < k <= 32 on tcgen05 (north_star item (d), BASELINE configs[4] sweep):
//
//     out[b][t][v] = sum_{c < ranks[b]} U[b][t][c] S[b][c] Vt[b][c][v]          (reference decompress_ms.py:128-131)
//
// At these ranks the product needs 8k flops per 8-byte output: beyond k ~ 12 the FP32 pipe cannot keep up with HBM
// (stages.cu:recon_smallk_kernel is FFMA bound from k = 10 on), and cgemm_tc.cu - built for long contractions - spends a
// whole CTA life (barrier set-up, TMEM allocation, two 112 KiB stages, TMA latency, epilogue) on every 128 x 128 tile
// with nothing overlapped. This kernel is persistent and software-pipelined across tiles instead:
//
//   * one CTA per SM walks over row panels (128 rows of one matrix) and, inside a panel, over column tiles of 64 complex
//     channels. The A operand of a panel - U S split into real and imaginary parts, each split again into TF32 hi + lo -
//     is built once in shared memory and reused by every column tile of the panel.
//   * complex arithmetic with TWO real accumulators per tile and the RAW rows of Vt as the only B operand:
//         D1 = Re(US) * Vt_raw      D2 = Im(US) * Vt_raw          (Vt_raw row c = (qr, qi) interleaved along the channel)
//         out_re[v] = D1[2v] - D2[2v+1]        out_im[v] = D1[2v+1] + D2[2v]
//     so no (-qi, qr) twin rows are written (half the operand conversion of cgemm_tc) and D's real view is interleaved
//     complex64 after one add per value in the epilogue. 3xTF32 (hi*hi + hi*lo + lo*hi), fp32 accumulation in TMEM; the
//     contraction is at most 32 long, i.e. at most 24 MMAs per accumulator chain, so no second accumulation level.
//   * 8 converter/epilogue warps: load the raw Vt tile of tile t+1 into registers (plain coalesced 128-bit loads: 64 bytes
//     per thread, no TMA needed at this size), split + store tile t's operand in the MN-major SWIZZLE_128B_BASE32B layout,
//     then drain tile t-1 from TMEM into a per-warp staging block and hand it to the TMA unit (bulk tensor stores) while
//     the tensor core works on tile t. Two operand buffers and two TMEM accumulator pairs (2 x 256 columns) carry the
//     overlap; one elected thread issues the MMAs.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int RT_M = 128;                 // rows per panel = MMA M
constexpr int RT_NC = 64;                 // complex columns per tile
constexpr int RT_NF = 2 * RT_NC;          // floats per tile row = MMA N
constexpr int RT_KMAX = 32;               // contraction rows per operand buffer
constexpr uint32_t RT_A_BYTES = RT_M * 128;                 // one K-major operand: 128 rows x 32 floats
constexpr uint32_t RT_B_BYTES = (RT_NF / 32) * RT_KMAX * 128;  // one MN-major operand: 4 groups x 32 rows x 128 B = 16 KiB
constexpr uint32_t RT_OFF_A = 0;                            // Pr_hi, Pr_lo, Pi_hi, Pi_lo
constexpr uint32_t RT_OFF_B = 4 * RT_A_BYTES;               // 2 buffers x (hi, lo)
constexpr uint32_t RT_OFF_STAGE = RT_OFF_B + 4 * RT_B_BYTES;    // 128 rows x 512 B staging tile of the epilogue
constexpr uint32_t RT_OFF_BARS = RT_OFF_STAGE + RT_M * 512;
constexpr uint32_t RT_SMEM = RT_OFF_BARS + 128 + 1024;
constexpr int RT_CONV = 256;
constexpr int RT_THREADS = 128 + RT_CONV;
constexpr uint32_t RT_TMEM_COLS = 512;

struct ReconArgs {
    const float2* U;
    const float* S;
    const float2* Vt;
    const int32_t* ranks;
    float2* out;
    int B, m, n, kmax;
    int tiles_m, tiles_n;   // row panels per matrix, column tiles per matrix
    int nsplit;             // column-tile groups per panel (extra parallelism for small batches)
    int items;              // B * tiles_m * nsplit
};

struct TileRef {
    int b, m0, n0c, t;
};

// Drain one finished tile: TMEM -> registers (one accumulator row per thread) -> complex values -> the warp's own 8 KiB
// staging block in shared memory (two TMA boxes of 32 rows x 128 bytes, SWIZZLE_128B: a lane writes its own row, the
// 16-byte chunk index XOR-ed with the row keeps the 128-bit stores conflict free) -> HBM by two bulk tensor stores that one
// lane issues: whole 128-byte lines, clipped by the tensor map at the ragged ends of m and n, and asynchronous - the warp
// goes on to the next tile while the TMA unit reads the block; it only waits for that read before writing the block again.
// (Earlier variants: rows stored straight from the accumulator registers, 16 bytes per lane and lanes a row apart - a
// memory transaction per lane, 14 GB/s per SM; staged block read back and stored by the warp itself - 0.64 of HBM, the
// eight warps were latency bound on the shared-memory round trip.) No CTA-wide barrier: the eight warps drift freely.
__device__ __forceinline__ void rt_epilogue(const CUtensorMap* mapO, const TileRef& tr, uint32_t bar_accfull,
                                            uint32_t bar_accfree, uint32_t tmem_base, uint32_t stage, int quad, int chalf,
                                            int lane, int cwarp) {
    const int s = tr.t & 1;
    mbar_wait(bar_accfull + 8 * s, ((uint32_t)tr.t >> 1) & 1);
    fence_after_sync();
    const uint32_t st = stage + cwarp * (32 * 256);   // this warp's rows quad*32 .. +31, complex columns chalf*32 .. +31
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 256 + chalf * 64);
    // the bulk stores of the previous tile must have read the block
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t d1a[16], d1b[16], d2a[16], d2b[16];
        tmem_ld16(taddr + h * 32, d1a);
        tmem_ld16(taddr + h * 32 + 16, d1b);
        tmem_ld16(taddr + 128 + h * 32, d2a);
        tmem_ld16(taddr + 128 + h * 32 + 16, d2b);
        tmem_ld_wait();
        if (h == 1) {
            // everything this thread needs from the accumulator pair is in registers: hand it back to the MMA issuer
            fence_before_sync();
            mbar_arrive(bar_accfree + 8 * s);
        }
        // 32 accumulator columns = 16 complex outputs = 8 chunks of 16 bytes = box h of this thread's row
        const uint32_t rowaddr = st + h * 4096 + lane * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t* p1 = q < 4 ? d1a : d1b;
            const uint32_t* p2 = q < 4 ? d2a : d2b;
            const int o = (q & 3) * 4;
            const float4 r = make_float4(__uint_as_float(p1[o]) - __uint_as_float(p2[o + 1]),
                                         __uint_as_float(p1[o + 1]) + __uint_as_float(p2[o]),
                                         __uint_as_float(p1[o + 2]) - __uint_as_float(p2[o + 3]),
                                         __uint_as_float(p1[o + 3]) + __uint_as_float(p2[o + 2]));
            sts128(rowaddr + ((q ^ (lane & 7)) << 4), r);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();  // the block is staged and visible to the async proxy
    if (lane == 0) {
        const int c0 = 2 * (tr.n0c + chalf * 32), c1 = tr.m0 + quad * 32;
        tma_store_3d(mapO, st, c0, c1, tr.b);
        tma_store_3d(mapO, st + 4096, c0 + 32, c1, tr.b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}

__global__ void __launch_bounds__(RT_THREADS, 1) recon_tc_kernel(const __grid_constant__ CUtensorMap mapO, const ReconArgs g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + RT_OFF_BARS;
    const uint32_t bar_bfull = bars, bar_bfree = bars + 16, bar_accfull = bars + 32, bar_accfree = bars + 48, bar_afull = bars + 64;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + RT_OFF_BARS + 96);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_bfull + 8 * s, RT_CONV);
            mbar_init(bar_bfree + 8 * s, 1);
            mbar_init(bar_accfull + 8 * s, 1);
            mbar_init(bar_accfree + 8 * s, RT_CONV);
        }
        mbar_init(bar_afull, RT_CONV);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), RT_TMEM_COLS);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int ksteps = (g.kmax + 7) >> 3;            // MMA K = 8 contraction rows per instruction

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        if (warp == 1 && lane == 0) {
            // ===================== MMA issuer =====================
            const uint32_t idesc = idesc_tf32(RT_M, RT_NF, true);
            const uint64_t pr_hi = desc_kmajor_sw128(sbase + RT_OFF_A), pr_lo = desc_kmajor_sw128(sbase + RT_OFF_A + RT_A_BYTES);
            const uint64_t pi_hi = desc_kmajor_sw128(sbase + RT_OFF_A + 2 * RT_A_BYTES),
                           pi_lo = desc_kmajor_sw128(sbase + RT_OFF_A + 3 * RT_A_BYTES);
            int t = 0, pc = 0;
            for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
                const int part = item % g.nsplit;
                const int nt0 = part * g.tiles_n / g.nsplit, nt1 = (part + 1) * g.tiles_n / g.nsplit;  // never empty: nsplit <= tiles_n
                mbar_wait(bar_afull, pc & 1);
                fence_after_sync();
                for (int nt = nt0; nt < nt1; ++nt, ++t) {
                    const int s = t & 1;
                    const uint32_t use = (uint32_t)t >> 1;
                    mbar_wait(bar_accfree + 8 * s, (use & 1) ^ 1);
                    mbar_wait(bar_bfull + 8 * s, use & 1);
                    fence_after_sync();
                    const uint32_t bb = sbase + RT_OFF_B + s * 2 * RT_B_BYTES;
                    const uint32_t d1 = tmem_base + (uint32_t)s * 256u, d2 = d1 + 128u;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);
                        const uint64_t b_hi = desc_mnmajor_sw128_32b(bb + k * 1024, RT_KMAX * 128, 512);
                        const uint64_t b_lo = desc_mnmajor_sw128_32b(bb + RT_B_BYTES + k * 1024, RT_KMAX * 128, 512);
                        umma_tf32(d1, pr_lo + adv, b_hi, idesc, k != 0);
                        umma_tf32(d1, pr_hi + adv, b_lo, idesc, 1);
                        umma_tf32(d1, pr_hi + adv, b_hi, idesc, 1);
                        umma_tf32(d2, pi_lo + adv, b_hi, idesc, k != 0);
                        umma_tf32(d2, pi_hi + adv, b_lo, idesc, 1);
                        umma_tf32(d2, pi_hi + adv, b_hi, idesc, 1);
                    }
                    umma_commit(bar_bfree + 8 * s);
                    umma_commit(bar_accfull + 8 * s);
                }
                ++pc;
            }
        }
    } else {
        // ===================== converters / epilogue =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int ct = threadIdx.x - 128;
        const int quad = warp & 3;
        const int chalf = (warp - 4) >> 2;
        const int nq = ksteps;                       // float4 per thread of the raw Vt tile: kp * 32 / 256, kp = 8 ksteps
        float4 raw[RT_KMAX / 8];
        bool pending = false;
        TileRef prev = {0, 0, 0, 0};
        int t = 0;
        for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
            const int part = item % g.nsplit;
            const int pm = item / g.nsplit;
            const int b = pm / g.tiles_m, m0 = (pm % g.tiles_m) * RT_M;
            const int nt0 = part * g.tiles_n / g.nsplit, nt1 = (part + 1) * g.tiles_n / g.nsplit;
            const int rank = g.ranks ? min(max(g.ranks[b], 0), g.kmax) : g.kmax;
            const float2* Vb = g.Vt + (size_t)b * g.kmax * g.n;
            // raw Vt values of the first tile of the panel: thread ct, slot i -> float4 index f = ct + 256 i of the
            // [kp rows][32 float4] tile (row c = f / 32, 16-byte chunk q = f % 32)
            auto load_raw = [&](int nt) {
#pragma unroll
                for (int i = 0; i < RT_KMAX / 8; ++i) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < nq) {
                        const int f = ct + RT_CONV * i;
                        const int c = f >> 5, q = f & 31;
                        const int vc = nt * RT_NC + 2 * q;       // first complex column of this chunk
                        if (c < rank) {
                            const float2* src = Vb + (size_t)c * g.n + vc;
                            if (vc + 1 < g.n) v = __ldg(reinterpret_cast<const float4*>(src));
                            else if (vc < g.n) {
                                const float2 one = __ldg(src);
                                v = make_float4(one.x, one.y, 0.f, 0.f);
                            }
                        }
                    }
                    raw[i] = v;
                }
            };
            load_raw(nt0);
            // the previous panel's last tile must have left the tensor core before its A operand is replaced
            if (pending) {
                rt_epilogue(&mapO, prev, bar_accfull, bar_accfree, tmem_base, sbase + RT_OFF_STAGE, quad, chalf, lane, warp - 4);
                pending = false;
            }
            // ---- A operand of the panel: Pr = Re(U S), Pi = Im(U S), hi / lo, K-major SWIZZLE_128B ----
            {
                const int row = ct >> 1, half = ct & 1;          // two threads per row, 16 contraction columns each
                const int gi = m0 + row;
                const float2* urow = g.U + ((size_t)b * g.m + (gi < g.m ? gi : 0)) * g.kmax;
                const float* sb = g.S + (size_t)b * g.kmax;
#pragma unroll
                for (int cq = 0; cq < 4; ++cq) {                 // 16-byte chunk = 4 contraction columns
                    const int chunk = half * 4 + cq;
                    float pr[4], pi[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = chunk * 4 + e;
                        float2 u = make_float2(0.f, 0.f);
                        if (gi < g.m && c < rank) {
                            u = __ldg(urow + c);
                            const float sc = __ldg(sb + c);
                            u.x *= sc, u.y *= sc;
                        }
                        pr[e] = u.x, pi[e] = u.y;
                    }
                    const Split4 sr = split4(make_float4(pr[0], pr[1], pr[2], pr[3]));
                    const Split4 si = split4(make_float4(pi[0], pi[1], pi[2], pi[3]));
                    const uint32_t off = (uint32_t)row * 128 + (uint32_t)((chunk ^ (row & 7)) << 4);
                    sts128(sbase + RT_OFF_A + off, sr.hi);
                    sts128(sbase + RT_OFF_A + RT_A_BYTES + off, sr.lo);
                    sts128(sbase + RT_OFF_A + 2 * RT_A_BYTES + off, si.hi);
                    sts128(sbase + RT_OFF_A + 3 * RT_A_BYTES + off, si.lo);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(bar_afull);
            }
            for (int nt = nt0; nt < nt1; ++nt, ++t) {
                const int s = t & 1;
                const uint32_t use = (uint32_t)t >> 1;
                mbar_wait(bar_bfree + 8 * s, (use & 1) ^ 1);
                const uint32_t b_hi = sbase + RT_OFF_B + s * 2 * RT_B_BYTES;
                const uint32_t b_lo = b_hi + RT_B_BYTES;
#pragma unroll
                for (int i = 0; i < RT_KMAX / 8; ++i) {
                    if (i < nq) {
                        const int f = ct + RT_CONV * i;
                        const int c = f >> 5, q = f & 31;
                        const int grp = q >> 3, lc = q & 7;
                        const Split4 sp = split4(raw[i]);
                        // MN-major SWIZZLE_128B_BASE32B: 32-byte chunk index (lc >> 1) XOR (row & 3); 16-byte halves keep order
                        const uint32_t o = (uint32_t)(grp * RT_KMAX + c) * 128 + (uint32_t)(((((lc >> 1) ^ (c & 3)) << 1) | (lc & 1)) << 4);
                        sts128(b_hi + o, sp.hi);
                        sts128(b_lo + o, sp.lo);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(bar_bfull + 8 * s);
                if (nt + 1 < nt1) load_raw(nt + 1);
                if (pending) rt_epilogue(&mapO, prev, bar_accfull, bar_accfree, tmem_base, sbase + RT_OFF_STAGE, quad, chalf, lane, warp - 4);
                pending = true;
                prev = {b, m0, nt * RT_NC, t};
            }
        }
        if (pending) rt_epilogue(&mapO, prev, bar_accfull, bar_accfree, tmem_base, sbase + RT_OFF_STAGE, quad, chalf, lane, warp - 4);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory outlives the bulk stores
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        fence_after_sync();
        tmem_dealloc(tmem_base, RT_TMEM_COLS);
    }
}

}  // namespace

bool vk_recon_tc_supported(int m, int n, int kmax) { return kmax > 8 && kmax <= RT_KMAX && (n % 2) == 0 && m >= 1; }

int vk_launch_recon_tc_smallk(vk_context* h, const float2* U, const float* S, const float2* Vt, const int32_t* ranks,
                              float2* out, int B, int m, int n, int kmax) {
    VK_CUDA(h, cudaFuncSetAttribute(recon_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RT_SMEM));
    ReconArgs g;
    g.U = U, g.S = S, g.Vt = Vt, g.ranks = ranks, g.out = out;
    g.B = B, g.m = m, g.n = n, g.kmax = kmax;
    g.tiles_m = (m + RT_M - 1) / RT_M;
    g.tiles_n = (n + RT_NC - 1) / RT_NC;
    const long long panels = (long long)B * g.tiles_m;
    int nsplit = 1;
    const int want = 2 * h->num_sms;
    while (panels * nsplit < want && nsplit * 2 <= g.tiles_n) nsplit *= 2;
    g.nsplit = nsplit;
    if (panels * nsplit > 0x7fffffffLL) return vk_fail(h, VK_EINVAL, "reconstruct: too many tiles");
    g.items = (int)(panels * nsplit);
    const int grid = g.items < h->num_sms ? g.items : h->num_sms;
    // output as a 3-D tensor (2n floats, m rows, B matrices): boxes of 32 rows x 128 bytes, clipped at m and 2n
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return vk_fail(h, VK_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    CUtensorMap mapO;
    const cuuint64_t dims[3] = {(cuuint64_t)2 * n, (cuuint64_t)m, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)n * 8, (cuuint64_t)m * n * 8};
    const cuuint32_t box[3] = {32, 32, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(&mapO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return vk_fail(h, VK_ECUDA, "tensor map (reconstruction output) failed: " + std::to_string((int)r));
    recon_tc_kernel<<<grid, RT_THREADS, RT_SMEM, h->stream>>>(mapO, g);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}
