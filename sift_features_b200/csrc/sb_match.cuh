// sb_match.cuh -- brute-force L2 descriptor matching with cross-check on the 5th-generation tensor cores.
//
// What it replaces: the step AFTER the extraction path in the reference's examples
// (examples/sift-match.rs:30-35, examples/opencv-cross-match.rs:34-43: OpenCV BFMatcher(NORM_L2, crossCheck)
// over the (N,128) u8 descriptor matrices).
//
// ||a - b||^2 = |a|^2 + |b|^2 - 2 a.b with u8 operands is exact in s32, so the Gram matrix is one dense
// u8 x u8 contraction: tcgen05.mma kind::i8, A and B tiles staged in shared memory by TMA (128-byte rows =
// one SWIZZLE_128B atom row per descriptor), accumulators in TMEM (two 128 x 256 s32 stages = all 512 columns),
// and the distance matrix is never materialised: the epilogue warps read each accumulator stage with
// tcgen05.ld and keep a running (distance, index) minimum per query row while the next stage is being filled.
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocation + MMA issue (one elected lane), tcgen05.commit onto the mbarriers
//   warps 2..9  : epilogue, warp w owns TMEM lanes 32*(w%4) .. +31 (one query row per thread) and one half of the
//                 tile's 256 columns: two 64-column tcgen05.ld per tile and thread, then per column pair two IMADs
//                 (256 |b|^2 + column - 512 a.b, signed: the |a|^2 term is constant per row and added at the end)
//                 and one three-input minimum, with the tile's |b|^2 values staged once in shared memory
// The kernel is run in both directions (query->train, train->query); k_match_cross keeps the mutual pairs in
// ascending query order.  Ties resolve to the smallest index (integer distances: deterministic).
#pragma once
#include <cuda.h>

#include "sb_common.cuh"

namespace sb {

constexpr int MT_M = 128;        // query rows per CTA (UMMA M)
constexpr int MT_N = 256;        // train rows per tile (UMMA N)
constexpr int MT_K = 128;        // descriptor bytes == one swizzle-128B row
constexpr int MT_UK = 32;        // K per tcgen05.mma for 8-bit operands
constexpr int MT_STAGES = 3;     // train tiles in flight
constexpr int MT_EPI_WARPS = 8;
constexpr int MT_THREADS = 64 + 32 * MT_EPI_WARPS;
constexpr uint32_t MT_A_BYTES = MT_M * MT_K;
constexpr uint32_t MT_B_BYTES = MT_N * MT_K;
constexpr size_t MT_SMEM = 1024 + MT_A_BYTES + (size_t)MT_STAGES * MT_B_BYTES;   // 1 KB slack for the 1024-byte alignment
constexpr uint32_t MT_TMEM_COLS = 512;

// |b|^2 entry of a padding row: larger than any real (|b|^2 << 8) - 512 a.b, so a padding column never wins a minimum
constexpr uint32_t MT_NBP_PAD = 0x7fffff00u;
constexpr uint32_t MT_NBP_REAL_MAX = 0x7f800000u;   // real entries are at most (128 * 255^2 << 8) + 255 = 0x7f0080ff

struct MatchParams {
    const uint32_t* norm_a;        // [n_a] |a|^2
    const uint32_t* nbp;           // [ceil(n_b / MT_N) * MT_N]  (|b|^2 << 8) | (j & 255); MT_NBP_PAD beyond n_b
    uint32_t n_a, n_b;
    uint32_t tiles_per_cta;        // train tiles per CTA: blockIdx.y selects the range (the ranges are merged by atomicMin)
    unsigned long long* best;      // [n_a]  (distance^2 << 32) | argmin j; all-ones before the launch
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1),
                   "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// leading byte offset 1 (unused for swizzled K-major), stride byte offset = 8 rows x 128 B, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(const void* smem) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    return (uint64_t)((a >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = u8, both K-major, N >> 3, M >> 4
constexpr uint32_t MT_IDESC = (2u << 4) | ((uint32_t)(MT_N >> 3) << 17) | ((uint32_t)(MT_M >> 4) << 24);

__global__ void __launch_bounds__(MT_THREADS, 1) k_match_nn(const __grid_constant__ CUtensorMap tm_a,
                                                             const __grid_constant__ CUtensorMap tm_b, const MatchParams p) {
    extern __shared__ uint8_t mt_raw[];
    uint8_t* const sm = reinterpret_cast<uint8_t*>(((uintptr_t)mt_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* const sA = sm;
    uint8_t* const sB = sm + MT_A_BYTES;
    __shared__ __align__(8) uint64_t full_a, full_b[MT_STAGES], empty_b[MT_STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) uint32_t s_nbp[2][MT_N];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * MT_M;
    // this CTA's train tiles: [t_first, t_first + n_tiles)
    const int all_tiles = (int)((p.n_b + MT_N - 1) / MT_N);
    const int t_first = (int)(blockIdx.y * p.tiles_per_cta);
    const int n_tiles = min((int)p.tiles_per_cta, all_tiles - t_first);
    if (n_tiles <= 0) return;   // block-uniform, before any barrier / TMEM use
    if (tid == 0) {
        mbar_init(&full_a, 1);
        for (int s = 0; s < MT_STAGES; s++) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], MT_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // whole warp: allocate all of TMEM (one CTA per SM), publish the base address
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(MT_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {   // ---- TMA producer ----
            mbar_expect_tx(&full_a, MT_A_BYTES);
            tma_load_2d(sA, &tm_a, 0, m0, &full_a);
            for (int t = 0; t < n_tiles; t++) {
                const int s = t % MT_STAGES;
                if (t >= MT_STAGES) mbar_wait(&empty_b[s], (uint32_t)((t / MT_STAGES) - 1) & 1u);
                mbar_expect_tx(&full_b[s], MT_B_BYTES);
                tma_load_2d(sB + (size_t)s * MT_B_BYTES, &tm_b, 0, (t_first + t) * MT_N, &full_b[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---- MMA issuer ----
            mbar_wait(&full_a, 0);
            const uint64_t da = umma_desc_k_sw128(sA);
            for (int t = 0; t < n_tiles; t++) {
                const int s = t % MT_STAGES, acc = t & 1;
                if (t >= 2) mbar_wait(&tmem_empty[acc], (uint32_t)((t >> 1) - 1) & 1u);   // epilogue drained this stage
                mbar_wait(&full_b[s], (uint32_t)(t / MT_STAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t db = umma_desc_k_sw128(sB + (size_t)s * MT_B_BYTES);
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * MT_N;
#pragma unroll
                for (int k = 0; k < MT_K / MT_UK; k++) {
                    // advancing K inside the swizzle atom = advancing the start-address field by 32 B (>> 4)
                    const uint64_t ak = da + (uint64_t)(k * MT_UK >> 4), bk = db + (uint64_t)(k * MT_UK >> 4);
                    const uint32_t accumulate = k > 0;
                    asm volatile(
                        "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p; }"
                        ::"r"(d_tmem), "l"(ak), "l"(bk), "r"(MT_IDESC), "r"(accumulate)
                        : "memory");
                }
                // completion of the MMAs above frees the smem stage and publishes the accumulator stage
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             ::"r"((uint32_t)__cvta_generic_to_shared(&empty_b[s])) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_full[acc])) : "memory");
            }
        }
    } else {
        // ---- epilogue: thread <-> TMEM lane <-> query row, warp <-> one half of the tile's columns ----
        const int q = warp & 3;                      // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;            // columns [128 * half, 128 * half + 128) of every tile
        const int etid = tid - 64;                   // 0 .. 255 among the epilogue threads
        const int row = m0 + 32 * q + lane;
        const uint32_t na256 = (row < (int)p.n_a) ? (__ldg(p.norm_a + row) << 8) : 0u;
        uint32_t best_d2 = 0xffffffffu, best_j = 0;
        for (int t = 0; t < n_tiles; t++) {
            const int acc = t & 1;
            const int tg = t_first + t;   // global tile index
            // the tile's 256 (|b|^2 << 8 | column) words, one per epilogue thread, double-buffered: the named barrier
            // below also orders the reuse of a buffer (every warp passed the barrier of tile t-1 after reading tile t-2)
            s_nbp[acc][etid] = __ldg(p.nbp + (size_t)tg * MT_N + etid);
            asm volatile("bar.sync 1, %0;" ::"n"(32 * MT_EPI_WARPS) : "memory");
            mbar_wait(&tmem_full[acc], (uint32_t)(t >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)acc * MT_N + (uint32_t)half * 128u;
            uint32_t v[128];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t* u = v + 64 * c;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                    "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                    "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                      "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                      "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31]),
                      "=r"(u[32]), "=r"(u[33]), "=r"(u[34]), "=r"(u[35]), "=r"(u[36]), "=r"(u[37]), "=r"(u[38]), "=r"(u[39]),
                      "=r"(u[40]), "=r"(u[41]), "=r"(u[42]), "=r"(u[43]), "=r"(u[44]), "=r"(u[45]), "=r"(u[46]), "=r"(u[47]),
                      "=r"(u[48]), "=r"(u[49]), "=r"(u[50]), "=r"(u[51]), "=r"(u[52]), "=r"(u[53]), "=r"(u[54]), "=r"(u[55]),
                      "=r"(u[56]), "=r"(u[57]), "=r"(u[58]), "=r"(u[59]), "=r"(u[60]), "=r"(u[61]), "=r"(u[62]), "=r"(u[63])
                    : "r"(taddr + 64u * (uint32_t)c));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // the accumulator stage is in registers: hand it back before the arithmetic
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            // signed: 256 (|b|^2 - 2 a.b) + column lies in (-2^31, 2^31) because |a|^2, |b|^2 <= 128 * 255^2 < 2^23
            int tmin[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
            const uint4* nb4 = reinterpret_cast<const uint4*>(&s_nbp[acc][128 * half]);
#pragma unroll
            for (int k = 0; k < 128; k += 4) {
                const uint4 nb = nb4[k >> 2];   // the same address for every lane: one broadcast
                const int t0 = (int)nb.x - 512 * (int)v[k], t1 = (int)nb.y - 512 * (int)v[k + 1];
                const int t2 = (int)nb.z - 512 * (int)v[k + 2], t3 = (int)nb.w - 512 * (int)v[k + 3];
                tmin[(k >> 2) & 1] = __vimin3_s32(tmin[(k >> 2) & 1], t0, t1);
                tmin[2 + ((k >> 2) & 1)] = __vimin3_s32(tmin[2 + ((k >> 2) & 1)], t2, t3);
            }
            const int tile_min = min(min(tmin[0], tmin[1]), min(tmin[2], tmin[3]));
            if (tile_min < (int)MT_NBP_REAL_MAX) {   // a real column (padding columns sit at MT_NBP_PAD)
                const uint32_t full = (uint32_t)tile_min + na256;   // 256 d^2 + column, >= 0
                const uint32_t d2 = full >> 8;
                if (d2 < best_d2) { best_d2 = d2; best_j = (uint32_t)(tg * MT_N) + (full & 255u); }
            }
        }
        // merge with the other column half and the other tile ranges of this row: smaller distance first, then smaller index
        if (row < (int)p.n_a && best_d2 != 0xffffffffu) atomicMin(p.best + row, ((unsigned long long)best_d2 << 32) | best_j);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(MT_TMEM_COLS));
    }
}

// |x|^2 per descriptor: norm[row] and nbp[row] = (norm << 8) | (row & 255); one warp per row
__global__ void __launch_bounds__(256) k_match_prep(const uint8_t* __restrict__ d, uint32_t n, uint32_t* __restrict__ norm,
                                                     uint32_t* __restrict__ nbp, uint32_t n_pad) {
    const uint32_t row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    uint32_t s = 0;
    if (row < n) {
        const uchar4 v = reinterpret_cast<const uchar4*>(d + (size_t)row * DESC_SIZE)[lane];
        s = (uint32_t)v.x * v.x + (uint32_t)v.y * v.y + (uint32_t)v.z * v.z + (uint32_t)v.w * v.w;
#pragma unroll
        for (int k = 16; k >= 1; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    }
    if (lane == 0) {
        if (row < n) norm[row] = s;
        nbp[row] = row < n ? ((s << 8) | (row & 255u)) : MT_NBP_PAD;
    }
}

struct MatchOut { uint32_t query, train, dist2; };

// mutual nearest neighbours in ascending query order (one CTA; the lists are small)
__global__ void __launch_bounds__(1024) k_match_cross(const unsigned long long* __restrict__ best_q,
                                                       const unsigned long long* __restrict__ best_t, uint32_t n_q,
                                                       uint32_t n_t, MatchOut* __restrict__ out, uint32_t cap,
                                                       uint32_t* __restrict__ n_out) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t s_run;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_q; base += 1024) {
        const uint32_t i = base + tid;
        bool keep = false;
        uint32_t j = 0, d2 = 0;
        if (i < n_q && n_t > 0) {
            const unsigned long long b = best_q[i];
            j = (uint32_t)b; d2 = (uint32_t)(b >> 32);
            keep = j < n_t && (uint32_t)best_t[j] == i;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        if (warp == 0) {
            uint32_t v = wsum[lane], sc = v;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, sc, k);
                if (lane >= k) sc += u;
            }
            wsum[lane] = sc - v;   // exclusive warp offsets
        }
        __syncthreads();
        const uint32_t run = s_run;
        if (keep) {
            const uint32_t pos = run + wsum[warp] + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) out[pos] = MatchOut{i, j, d2};
        }
        __syncthreads();
        if (tid == 1023) s_run = run + wsum[31] + __popc(m);
        __syncthreads();
    }
    if (tid == 0) *n_out = s_run;
}

}  // namespace sb
