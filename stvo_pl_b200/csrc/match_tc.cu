// match_tc.cu — K1 on the 5th-generation tensor cores: all-pairs 256-bit Hamming 2-NN as a +-1 contraction.
//
// Replaces the arithmetic of StVO::matchNNR / StVO::match (src/matching.cpp:41-91), i.e. OpenCV's
// cv::BFMatcher(NORM_HAMMING)::knnMatch(desc1, desc2, ., 2) per direction.  With descriptor bits mapped to +-1,
//     a . b = (#equal bits) - (#different bits) = 256 - 2 d(a, b)          (exact, |a . b| <= 256)
// so the N1 x N2 distance matrix of one matching problem is a 256-deep GEMM.  Three kernels:
//   tc_expand_kernel   32-byte descriptor rows -> 256 e4m3 bytes (+1.0 = 0x38, -1.0 = 0xB8), written directly in the
//                      128-byte-swizzled K-major operand layout tcgen05.mma reads (tiles of 128 rows = 32 KB contiguous),
//                      so the tiles move with plain bulk copies (UBLKCP) and need no tensor map;
//   tc_hamming_kernel  persistent, warp-specialised: producer warp (bulk copies + mbarriers), one MMA-issuing thread
//                      (tcgen05.mma kind::f8f6f4, 128 x 128 x 32 per instruction, f16 accumulators in TMEM: every partial
//                      sum is an integer of magnitude <= 256, exact in f16), eight epilogue warps reading TMEM with
//                      tcgen05.ld ... .pack::16b: two columns per register, so the top-2 update of BOTH directions
//                      costs 3 packed min/max per two distances and direction:
//                        row direction   (query -> trains): thread = query row, fold over the columns in registers;
//                        column direction (train -> queries): element-wise running top-2 over the row tiles in
//                                         registers, one cross-lane reduction per work item (not per tile).
//                      Keys carry no indices in the hot loop: only (best, second) VALUES and a small candidate tag
//                      (which 16-column group / which lane the best came from);
//   tc_resolve_kernel  merges the per-block partials, converts to distances and recovers the best neighbour's index by
//                      re-evaluating the <= 16 tagged candidates with XOR + POPC (exact; ties resolved to the lowest index
//                      as OpenCV does), emitting the same packed keys (dist << 16 | index) K1's popcount form emits, so
//                      everything downstream (ratio test, mutual filter, K2) is unchanged.
// Integer-exact end to end: the bit-exact parity tests of the matcher are the acceptance test.
#include <cuda_fp16.h>

#include "common.cuh"
#include "match_tc.cuh"

namespace plstvo {

// ---------------------------------------------------------------------------------------------------------------
// expand: bits -> e4m3 +-1 bytes in the swizzled operand layout
// tile (128 rows) = [slab 0..1][row group 0..15][row 0..7][16-byte chunk position 0..7]; slab = 128 K-bytes;
// logical chunk c of row r sits at position c ^ (r & 7)  (the 128B swizzle: address bits [4,7) ^= bits [7,10))
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_expand_kernel(const TcSide* __restrict__ sides) {
    const TcSide s = sides[blockIdx.y];
    const int tile = blockIdx.x;
    if (tile * TC_ROWS >= s.n) return;
    __shared__ uint32_t lut[16];
    if (threadIdx.x < 16) {
        uint32_t w = 0;
        for (int b = 0; b < 4; ++b) w |= ((threadIdx.x >> b) & 1 ? 0x38u : 0xB8u) << (8 * b);
        lut[threadIdx.x] = w;
    }
    __syncthreads();
    uint8_t* dst_tile = s.dst + (size_t)tile * TC_TILE_BYTES;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int u = it * 256 + threadIdx.x;
        const int c = u & 7, slab = (u >> 3) & 1, r = u >> 4;
        const int row = tile * TC_ROWS + r;
        uint4 out = make_uint4(0, 0, 0, 0);   // rows past the end: zeros (their results are masked, never used)
        if (row < s.n) {
            const uint32_t hw = *reinterpret_cast<const uint16_t*>(s.src + (size_t)row * 32 + slab * 16 + c * 2);
            out = make_uint4(lut[hw & 15], lut[(hw >> 4) & 15], lut[(hw >> 8) & 15], lut[hw >> 12]);
        }
        *reinterpret_cast<uint4*>(dst_tile + slab * 16384 + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) * 16)) = out;
    }
}

cudaError_t launch_tc_expand(const TcSide* sides, int n_sides, int max_tiles, cudaStream_t stream) {
    if (n_sides <= 0 || max_tiles <= 0) return cudaSuccess;
    tc_expand_kernel<<<dim3(max_tiles, n_sides), 256, 0, stream>>>(sides);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// PTX helpers (tcgen05 / TMEM)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// 32 lanes x 32 columns of f16 accumulators -> 16 registers, two adjacent columns per register (low half = even column)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t hmax2u(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hmin2u(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hne2mask(uint32_t a, uint32_t b) {   // 0xFFFF per half where a != b
    uint32_t d;
    asm("set.ne.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hgt2mask(uint32_t a, uint32_t b) {   // 0xFFFF per half where a > b
    uint32_t d;
    asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ float h_lo(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p & 0xFFFFu))); }
__device__ __forceinline__ float h_hi(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p >> 16))); }
__device__ __forceinline__ uint32_t f2h_bits(float f) { return (uint32_t)__half_as_ushort(__float2half_rn(f)); }

constexpr uint32_t NEG2 = 0xFC00FC00u;   // (-inf, -inf)

// shared-memory operand descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart (SBO), version 1
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)0x40004040u << 32);
}
// instruction descriptor: D = f16, A = B = e4m3, both K-major, N = 128, M = 128
constexpr uint32_t TC_IDESC = (uint32_t)((TC_ROWS >> 3) << 17) | (uint32_t)((TC_ROWS >> 4) << 24);

// ---------------------------------------------------------------------------------------------------------------
// the matcher
// ---------------------------------------------------------------------------------------------------------------
constexpr int TC_OFF_X = 0;
constexpr int TC_OFF_Y = TC_XSTAGES * TC_TILE_BYTES;
constexpr int TC_OFF_SCR = TC_OFF_Y + 2 * TC_TILE_BYTES;                  // per epilogue warp: 16 x 33 words
constexpr int TC_SCR_WARP = 16 * 33 * 4;
constexpr int TC_OFF_MRG = TC_OFF_SCR + 8 * TC_SCR_WARP;                  // [2 groups][4 quarters][128 columns] uint2
constexpr int TC_OFF_BAR = TC_OFF_MRG + 2 * 4 * 128 * 8;
constexpr int TC_NBARS = 2 * TC_XSTAGES + 2 + 2 * TC_ASTAGES;
constexpr int TC_OFF_SLOT = TC_OFF_BAR + TC_NBARS * 8;
constexpr int TC_SMEM_USED = TC_OFF_SLOT + 16;

size_t tc_smem_bytes() { return (size_t)TC_SMEM_USED + 1024; }

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_hamming_kernel(const TcProblem* __restrict__ problems, const TcItem* __restrict__ items, int n_items,
                  __half* __restrict__ debug_tile) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TC_OFF_BAR);
    uint64_t* x_full = bars;
    uint64_t* x_empty = bars + TC_XSTAGES;
    uint64_t* y_full = bars + 2 * TC_XSTAGES;
    uint64_t* y_empty = y_full + 1;
    uint64_t* t_full = y_empty + 1;
    uint64_t* t_empty = t_full + TC_ASTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TC_OFF_SLOT);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < TC_XSTAGES; ++i) {
            mbar_init(&x_full[i], 1);
            mbar_init(&x_empty[i], 1);
        }
        mbar_init(y_full, 1);
        mbar_init(y_empty, 1);
        for (int i = 0; i < TC_ASTAGES; ++i) {
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], 8);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TC_ASTAGES * 256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer: bulk copies of operand tiles =====
        if (lane == 0) {
            uint32_t xs = 0, xph = 0, yit = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++yit) {
                const TcItem item = items[it];
                const TcProblem pr = problems[item.problem];
                const int nyt = (pr.n2 + TC_ROWS - 1) / TC_ROWS, nxt = (pr.n1 + TC_ROWS - 1) / TC_ROWS;
                const int ytiles = min(2, nyt - 2 * item.yblk);
                mbar_wait(y_empty, (yit & 1) ^ 1);
                mbar_arrive_expect_tx(y_full, (uint32_t)ytiles * TC_TILE_BYTES);
                for (int h = 0; h < ytiles; ++h)
                    bulk_g2s(sm + TC_OFF_Y + h * TC_TILE_BYTES, pr.ye + (size_t)(2 * item.yblk + h) * TC_TILE_BYTES,
                             TC_TILE_BYTES, y_full);
                for (int t = 0; t < nxt; ++t) {
                    mbar_wait(&x_empty[xs], xph ^ 1);
                    mbar_arrive_expect_tx(&x_full[xs], TC_TILE_BYTES);
                    bulk_g2s(sm + TC_OFF_X + xs * TC_TILE_BYTES, pr.xe + (size_t)t * TC_TILE_BYTES, TC_TILE_BYTES,
                             &x_full[xs]);
                    if (++xs == TC_XSTAGES) { xs = 0; xph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread =====
        if (lane == 0) {
            uint32_t xs = 0, xph = 0, as = 0, aph = 0, yit = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++yit) {
                const TcItem item = items[it];
                const TcProblem pr = problems[item.problem];
                const int nxt = (pr.n1 + TC_ROWS - 1) / TC_ROWS;
                mbar_wait(y_full, yit & 1);
                for (int t = 0; t < nxt; ++t) {
                    mbar_wait(&t_empty[as], aph ^ 1);
                    mbar_wait(&x_full[xs], xph);
                    tc_fence_after();
                    const uint32_t xa = sbase + TC_OFF_X + xs * TC_TILE_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t ya = sbase + TC_OFF_Y + h * TC_TILE_BYTES;
                        const uint32_t d = tmem_base + as * 256 + h * 128;
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint32_t ko = (kk >> 2) * 16384 + (kk & 3) * 32;
                            tc_mma_f8(d, tc_desc(xa + ko), tc_desc(ya + ko), TC_IDESC, kk > 0);
                        }
                    }
                    tc_commit(&x_empty[xs]);
                    tc_commit(&t_full[as]);
                    if (++xs == TC_XSTAGES) { xs = 0; xph ^= 1; }
                    if (++as == TC_ASTAGES) { as = 0; aph ^= 1; }
                }
                tc_commit(y_empty);
            }
        }
    } else {
        // ===== epilogue: 8 warps.  quarter q = TMEM lanes 32q..32q+31 (hardware: warp id % 4), group gq = which 128 columns
        const int ew = warp - 2, q = warp & 3, gq = ew >> 2;
        uint32_t* scr = reinterpret_cast<uint32_t*>(sm + TC_OFF_SCR + ew * TC_SCR_WARP);
        uint2* mrg = reinterpret_cast<uint2*>(sm + TC_OFF_MRG) + gq * 4 * 128;
        uint32_t as = 0, aph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const TcItem item = items[it];
            const TcProblem pr = problems[item.problem];
            const int nxt = (pr.n1 + TC_ROWS - 1) / TC_ROWS;
            const int half = item.yblk * 2 + gq;              // 128-column block index of this group
            const int ycol0 = half * TC_ROWS;
            const int nvalid = min(TC_ROWS, pr.n2 - ycol0);   // <= 0: nothing for this group
            uint32_t c1[64], c2[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) c1[i] = c2[i] = NEG2;

            for (int t = 0; t < nxt; ++t) {
                mbar_wait(&t_full[as], aph);
                tc_fence_after();
                const int x = t * TC_ROWS + q * 32 + lane;
                const bool xvalid = x < pr.n1;
                const bool clean = (nvalid == TC_ROWS) && ((t + 1) * TC_ROWS <= pr.n1);   // warp-uniform
                uint32_t r1 = NEG2, r2 = NEG2, rc = 0;
                const uint32_t tad = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256 + gq * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[16];
                    tc_ld32(tad + c * 32, v);
                    tc_wait_ld();
                    if (debug_tile && it == 0 && t == 0) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            *reinterpret_cast<uint32_t*>(debug_tile + (size_t)(q * 32 + lane) * 256 + gq * 128 + c * 32 + 2 * i) = v[i];
                    }
                    if (!clean) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int col = c * 32 + 2 * i;
                            if (!xvalid || col >= nvalid) v[i] = NEG2;
                            else if (col + 1 >= nvalid) v[i] = (v[i] & 0xFFFFu) | 0xFC000000u;
                        }
                    }
                    const uint32_t old = r1;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t tmin = hmin2u(r1, v[i]);
                        r1 = hmax2u(r1, v[i]);
                        r2 = hmax2u(r2, tmin);
                    }
                    const uint32_t chg = hne2mask(r1, old);
                    rc = (chg & (uint32_t)(c * 0x00010001u)) | (~chg & rc);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t tmin = hmin2u(c1[c * 16 + i], v[i]);
                        c1[c * 16 + i] = hmax2u(c1[c * 16 + i], v[i]);
                        c2[c * 16 + i] = hmax2u(c2[c * 16 + i], tmin);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[as]);
                if (++as == TC_ASTAGES) { as = 0; aph ^= 1; }
                if (xvalid && nvalid > 0) {
                    // combine the even-column and odd-column streams of this 128-column block
                    const float a1 = h_lo(r1), b1 = h_hi(r1), a2 = h_lo(r2), b2 = h_hi(r2);
                    const float v1 = fmaxf(a1, b1), v2 = fmaxf(fminf(a1, b1), fmaxf(a2, b2));
                    const uint32_t par = b1 > a1 ? 1u : 0u;
                    const uint32_t chunk = par ? (rc >> 16) : (rc & 0xFFFFu);
                    pr.rowp[(size_t)half * pr.n1 + x] = make_uint2(f2h_bits(v1) | (f2h_bits(v2) << 16), chunk | (par << 2));
                }
            }

            // ---- column direction: reduce the element-wise state over the 128 threads of the group ----
            // per warp: transpose 16 registers at a time through shared memory; lane L folds register (L & 15) over
            // source lanes 16 (L >> 4) .. +15, the two halves meet by shuffle; then the four quarters merge per column.
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int j = lane & 15, hl = lane >> 4;
#pragma unroll
                for (int i = 0; i < 16; ++i) scr[i * 33 + lane] = c1[p * 16 + i];
                __syncwarp();
                uint32_t m1 = NEG2, m2 = NEG2, id = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t v = scr[j * 33 + hl * 16 + k];
                    const uint32_t tmin = hmin2u(m1, v);
                    const uint32_t nm = hmax2u(m1, v);
                    m2 = hmax2u(m2, tmin);
                    const uint32_t chg = hne2mask(nm, m1);
                    id = (chg & (uint32_t)((hl * 16 + k) * 0x00010001u)) | (~chg & id);
                    m1 = nm;
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 16; ++i) scr[i * 33 + lane] = c2[p * 16 + i];
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 16; ++k) m2 = hmax2u(m2, scr[j * 33 + hl * 16 + k]);
                __syncwarp();
                // halves: lanes L and L ^ 16 hold the same register over the other 16 source lanes
                const uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, m1, 16), o2 = __shfl_xor_sync(0xFFFFFFFFu, m2, 16),
                               oid = __shfl_xor_sync(0xFFFFFFFFu, id, 16);
                const uint32_t gt = hgt2mask(o1, m1);
                const uint32_t n2v = hmax2u(hmin2u(m1, o1), hmax2u(m2, o2));
                const uint32_t n1v = hmax2u(m1, o1);
                const uint32_t nid = (gt & oid) | (~gt & id);
                if (hl == 0) {
                    const int col = (p * 16 + j) * 2;
                    mrg[q * 128 + col] = make_uint2((n1v & 0xFFFFu) | (n2v << 16), nid & 0xFFFFu);
                    mrg[q * 128 + col + 1] = make_uint2((n1v >> 16) | (n2v & 0xFFFF0000u), nid >> 16);
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + gq) : "memory");
            {
                const int col = (ew & 3) * 32 + lane;
                float b1 = -1e30f, b2 = -1e30f;
                uint32_t bt = 0;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const uint2 e = mrg[qq * 128 + col];
                    const float a1 = h_lo(e.x), a2 = h_hi(e.x);
                    b2 = fmaxf(fminf(b1, a1), fmaxf(b2, a2));
                    if (a1 > b1) { b1 = a1; bt = (uint32_t)qq * 32u + e.y; }
                }
                if (col < nvalid) pr.colp[ycol0 + col] = make_uint2(f2h_bits(b1) | (f2h_bits(b2) << 16), bt);
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + gq) : "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_ASTAGES * 256)
                     : "memory");
    }
}

cudaError_t launch_tc_hamming(const TcProblem* problems, const TcItem* items, int n_items, int grid, __half* debug_tile,
                              cudaStream_t stream) {
    if (n_items <= 0) return cudaSuccess;
    static size_t configured[64] = {};
    const size_t smem = tc_smem_bytes();
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(tc_hamming_kernel), smem, configured);
    if (e != cudaSuccess) return e;
    if (grid > n_items) grid = n_items;
    tc_hamming_kernel<<<grid, TC_THREADS, smem, stream>>>(problems, items, n_items, debug_tile);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// resolve: values -> distances, candidate tags -> exact indices, packed keys out
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint8_t* p) {
    const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(p)), b1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
           __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(256) tc_resolve_kernel(const MatchProblem* __restrict__ mps, const TcProblem* __restrict__ tps) {
    const MatchProblem mp = mps[blockIdx.x];
    if (!mp.enabled) return;
    const TcProblem tp = tps[blockIdx.x];
    const int nyh = (mp.n2 + TC_ROWS - 1) / TC_ROWS, nxt = (mp.n1 + TC_ROWS - 1) / TC_ROWS;
    const int slice = blockIdx.y, nslices = gridDim.y;
    // queries: top-2 trains
    for (int x = slice * blockDim.x + threadIdx.x; x < mp.n1; x += nslices * blockDim.x) {
        float b1 = -1e30f, b2 = -1e30f;
        int bh = 0;
        uint32_t bid = 0;
        for (int h = 0; h < nyh; ++h) {
            const uint2 e = tp.rowp[(size_t)h * mp.n1 + x];
            const float a1 = h_lo(e.x), a2 = h_hi(e.x);
            b2 = fmaxf(fminf(b1, a1), fmaxf(b2, a2));
            if (a1 > b1) { b1 = a1; bh = h; bid = e.y; }
        }
        const int d1 = (256 - (int)b1) >> 1;
        uint32_t k2 = KEY_NONE;
        int idx = 0;
        if (mp.n2 >= 2) {
            const int d2 = (256 - (int)b2) >> 1;
            k2 = ((uint32_t)d2 << 16) | 0xFFFEu;
            if ((float)d1 < __fmul_rn((float)d2, mp.nnr)) {
                const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(mp.d1 + (size_t)x * 32)),
                            a1 = __ldg(reinterpret_cast<const uint4*>(mp.d1 + (size_t)x * 32) + 1);
                idx = -1;
                if (d1 != d2) {
                    const int base = bh * TC_ROWS + (int)(bid & 3u) * 32 + (int)((bid >> 2) & 1u);
                    for (int i = 0; i < 16 && idx < 0; ++i) {
                        const int j = base + 2 * i;
                        if (j < mp.n2 && hamming256(a0, a1, mp.d2 + (size_t)j * 32) == d1) idx = j;
                    }
                }
                if (idx < 0) {   // tie accepted (nnr > 1): lowest index among all trains, as OpenCV orders them
                    for (int j = 0; j < mp.n2 && idx < 0; ++j)
                        if (hamming256(a0, a1, mp.d2 + (size_t)j * 32) == d1) idx = j;
                    if (idx < 0) idx = 0;
                }
            }
        }
        mp.rowpart[x] = make_uint2(((uint32_t)d1 << 16) | (uint32_t)idx, k2);
    }
    // trains: top-2 queries
    for (int y = slice * blockDim.x + threadIdx.x; y < mp.n2; y += nslices * blockDim.x) {
        const uint2 e = tp.colp[y];
        const int d1 = (256 - (int)h_lo(e.x)) >> 1;
        uint32_t k2 = KEY_NONE;
        int idx = 0;
        if (mp.n1 >= 2) {
            const int d2 = (256 - (int)h_hi(e.x)) >> 1;
            k2 = ((uint32_t)d2 << 16) | 0xFFFEu;
            if ((float)d1 < __fmul_rn((float)d2, mp.nnr)) {
                const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(mp.d2 + (size_t)y * 32)),
                            a1 = __ldg(reinterpret_cast<const uint4*>(mp.d2 + (size_t)y * 32) + 1);
                idx = -1;
                if (d1 != d2) {
                    for (int i = 0; i < nxt && idx < 0; ++i) {
                        const int j = (int)e.y + i * TC_ROWS;
                        if (j < mp.n1 && hamming256(a0, a1, mp.d1 + (size_t)j * 32) == d1) idx = j;
                    }
                }
                if (idx < 0) {
                    for (int j = 0; j < mp.n1 && idx < 0; ++j)
                        if (hamming256(a0, a1, mp.d1 + (size_t)j * 32) == d1) idx = j;
                    if (idx < 0) idx = 0;
                }
            }
        }
        mp.colpart[y] = make_uint2(((uint32_t)d1 << 16) | (uint32_t)idx, k2);
    }
}

cudaError_t launch_tc_resolve(const MatchProblem* mps, const TcProblem* tps, int n_problems, int slices, cudaStream_t stream) {
    if (n_problems <= 0) return cudaSuccess;
    tc_resolve_kernel<<<dim3(n_problems, slices > 0 ? slices : 1), 256, 0, stream>>>(mps, tps);
    return cudaGetLastError();
}

}  // namespace plstvo
