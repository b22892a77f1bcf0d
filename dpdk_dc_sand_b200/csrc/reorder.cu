// Stage 1 (stand-alone): pre-beamform reorder, bit-exact.
//
//   samples   [B][A][C][T][P=2][X=2] u8  ->  reordered [B][P][C][T/16][16][A][X] u8
//
// Replaces kernel `prebeamform_reorder`
// (reference: beamformer/beamforming/kernels/prebeamform_reorder_kernel.mako:37-92), which moves one
// 16-bit word per thread and scatters its stores at a stride of A words.  Here one CTA stages an
// [A antennas] x [<=64 samples] tile of 4-byte words {p0.re, p0.im, p1.re, p1.im} through shared memory:
//   - loads: each antenna row of the tile is 4*tt contiguous bytes -> 128-bit coalesced loads, one STS.128 each;
//   - stores: for a fixed pol the tile's output is ONE contiguous run of tt*A*2 bytes ([t][a][x]); a thread
//     gathers the 8 words of 8 consecutive (t, a) elements ONCE and emits the 16-byte chunk of BOTH pols
//     -> 128-bit, fully coalesced stores, 8 LDS + 8 PRMT per 32 output bytes;
//   - the rows of antenna group a/8 are rotated by 4*(a/8) words, so the 32 lanes of a gather (8 antenna groups
//     x 4 consecutive samples when A = 64) hit 32 different banks while the staging stores stay 16-byte
//     aligned.
// All global offsets are 64-bit (the reference's `int` indices overflow at B*A*C*T*P >= 2^31).
//
// Antenna counts that are multiples of 8 (16-byte output rows: every MeerKAT / SKA configuration) take the TMA form
// of the same transpose, `reorder_tma_kernel`: persistent CTAs (two per SM), one producer warp and eight transposing
// warps.  A tile is [64 antennas] x [32 samples]: ONE tensor-map TMA box load (128B-swizzled rows, one per antenna;
// antennas / samples past the end are zero-filled by the hardware) lands it in a 4-stage ring; warp g gathers antennas
// 8 g .. 8 g + 7 with lane = sample (8 conflict-free LDS.32, 8 PRMT) and writes both pols' 16-byte chunks into
// 128B-swizzled output boxes [32 samples][64 antennas x 2 B] (2 conflict-free STS.128); one TMA tensor store per pol
// writes the box (rows / antennas past the end are clipped).  No address arithmetic per element is left on the SM
// (the first version was bound by it: ncu `math_pipe_throttle`, profiles/r01_standalone_reorder_c3_ncu_summary.txt);
// per 8 KiB tile the SM issues ~180 warp instructions and 256 shared-memory wavefronts.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace dcbf {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(kThreads)
reorder_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int A, int C, int T, int tile_t,
               int n_tiles, uint32_t inv_a /* floor(2^32 / A) + 1 for A >= 2: e / A == umulhi(e, inv_a), e < 2^17 */) {
    extern __shared__ __align__(16) uint32_t raw[];  // [A][tile_t] words, row a rotated by 4 * (a / 8)

    const long long blk = blockIdx.x;
    const int tile = static_cast<int>(blk % n_tiles);
    const int c = static_cast<int>((blk / n_tiles) % C);
    const long long b = blk / (static_cast<long long>(n_tiles) * C);
    const int t0 = tile * tile_t;
    const int tt = min(tile_t, T - t0);  // multiple of 16
    const int mask = tile_t - 1;         // tile_t is a power of two

    // ---- coalesced 128-bit loads of the [A][tt] tile ----
    const int vec_per_row = tt >> 2;
    const int n_vec = A * vec_per_row;
    for (int i = threadIdx.x; i < n_vec; i += kThreads) {
        const int a = i / vec_per_row;
        const int v = i - a * vec_per_row;
        const size_t row = ((static_cast<size_t>(b) * A + a) * C + c) * static_cast<size_t>(T) + t0;
        const uint4 x = ldg_stream(reinterpret_cast<const uint4*>(in + row * 4) + v);
        *reinterpret_cast<uint4*>(raw + a * tile_t + ((4 * v + 4 * (a >> 3)) & mask)) = x;
    }
    __syncthreads();

    // ---- tt*A elements of 2 bytes per pol, contiguous in the output; 8 elements per 128-bit store ----
    const int n_chunk = (tt * A) >> 3;  // tt % 8 == 0
    const size_t base0 = ((static_cast<size_t>(b) * kPols * C + c) * static_cast<size_t>(T) + t0) * A * 2;
    uint4* dst0 = reinterpret_cast<uint4*>(out + base0);
    uint4* dst1 = reinterpret_cast<uint4*>(out + base0 + static_cast<size_t>(C) * T * A * 2);  // pol 1 plane
    const bool groups_of_8 = (A & 7) == 0;  // a chunk is then 8 antennas of ONE sample: one base address, constant stride
    for (int i = threadIdx.x; i < n_chunk; i += kThreads) {
        const int e = i << 3;
        // e / A (ncu: the gather was bound by its own integer arithmetic, not by memory)
        int t = inv_a ? static_cast<int>(__umulhi(static_cast<uint32_t>(e), inv_a)) : e;
        int a = e - t * A;
        uint32_t w[8];
        if (groups_of_8) {
            const uint32_t* src = raw + a * tile_t + ((t + 4 * (a >> 3)) & mask);
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = src[j * tile_t];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                w[j] = raw[a * tile_t + ((t + 4 * (a >> 3)) & mask)];
                if (++a == A) {
                    a = 0;
                    ++t;
                }
            }
        }
        uint4 v0, v1;  // low / high 16 bits of two words = pol 0 / pol 1 (re, im) of two elements
        v0.x = __byte_perm(w[0], w[1], 0x5410u);
        v0.y = __byte_perm(w[2], w[3], 0x5410u);
        v0.z = __byte_perm(w[4], w[5], 0x5410u);
        v0.w = __byte_perm(w[6], w[7], 0x5410u);
        v1.x = __byte_perm(w[0], w[1], 0x7632u);
        v1.y = __byte_perm(w[2], w[3], 0x7632u);
        v1.z = __byte_perm(w[4], w[5], 0x7632u);
        v1.w = __byte_perm(w[6], w[7], 0x7632u);
        stg_stream(dst0 + i, v0);
        stg_stream(dst1 + i, v1);
    }
}

// ------------------------------------------------------------------------------------------------------
// TMA form
// ------------------------------------------------------------------------------------------------------
constexpr int kTmaAnts = 64, kTmaT = 32;               // tile: 64 antennas x 32 samples x 4 B = 8 KiB
constexpr int kTmaTileBytes = kTmaAnts * kTmaT * 4;
constexpr int kTmaLoadStages = 4, kTmaOutStages = 4;    // 32 KiB of loads in flight per CTA, 2 CTAs per SM (6 stages: no faster)
constexpr int kTmaConsumerWarps = 8;
constexpr int kTmaThreads = (kTmaConsumerWarps + 1) * 32;
constexpr int kTmaSmem = 1024 + (kTmaLoadStages + kTmaOutStages) * kTmaTileBytes + 64;

struct ReorderTmaParams {
    int B, C, a_tiles, t_tiles;
    long long n_tiles;
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(kTmaThreads, 2)
reorder_tma_kernel(const __grid_constant__ ReorderTmaParams prm, const __grid_constant__ CUtensorMap tm_in,
                   const __grid_constant__ CUtensorMap tm_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t in_base = base, out_base = base + kTmaLoadStages * kTmaTileBytes;
    const uint32_t bar_base = out_base + kTmaOutStages * kTmaTileBytes;  // full[stages], empty[stages]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kTmaLoadStages; ++i) {
            mbar_init(bar_base + 8u * i, 1);
            mbar_init(bar_base + 8u * (kTmaLoadStages + i), kTmaConsumerWarps);
        }
        fence_mbar_init();
        prefetch_tensormap(&tm_in);
        prefetch_tensormap(&tm_out);
    }
    __syncthreads();
    // tile -> (b, c, sample chunk, antenna tile); consecutive tiles of a CTA's sequence are far apart on purpose
    // (tile = blockIdx + k * grid): neighbouring CTAs then work on neighbouring runs of DRAM
    auto decode = [&](long long tile, int* b, int* c, int* t0, int* a0) {
        const int at = static_cast<int>(tile % prm.a_tiles);
        long long r = tile / prm.a_tiles;
        const int tt = static_cast<int>(r % prm.t_tiles);
        r /= prm.t_tiles;
        *c = static_cast<int>(r % prm.C);
        *b = static_cast<int>(r / prm.C);
        *t0 = tt * kTmaT, *a0 = at * kTmaAnts;
    };
    if (warp == kTmaConsumerWarps) {
        // ---- producer: one box per tile ----
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++it) {
            const uint32_t st = it % kTmaLoadStages, ph = (it / kTmaLoadStages) & 1u;
            while (!mbar_try_wait(bar_base + 8u * (kTmaLoadStages + st), ph ^ 1u)) {
            }
            if (lane == 0) {
                int b, c, t0, a0;
                decode(tile, &b, &c, &t0, &a0);
                mbar_arrive_expect_tx(bar_base + 8u * st, kTmaTileBytes);
                tma_load_4d(in_base + st * kTmaTileBytes, &tm_in, bar_base + 8u * st, t0, c, a0, b);
            }
            __syncwarp();
        }
    } else {
        // ---- transpose: warp g <-> antennas 8 g .. 8 g + 7 of the tile, lane <-> sample ----
        const int g = warp;
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++it) {
            const uint32_t st = it % kTmaLoadStages, ph = (it / kTmaLoadStages) & 1u, os = it % kTmaOutStages;
            // the output boxes used kTmaOutStages tiles ago have been read by their stores (thread 0 committed them)
            if (threadIdx.x == 0) bulk_wait_group_read<kTmaOutStages - 1>();
            while (!mbar_try_wait(bar_base + 8u * st, ph)) {
            }
            named_bar_sync(1, kTmaConsumerWarps * 32);
            // raw row r = antenna, 128 B = 32 samples; 16-byte chunk j of row r sits at chunk j ^ (r & 7)
            const uint32_t src = in_base + st * kTmaTileBytes + (lane & 3) * 4;
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 8 * g + i;
                w[i] = ld_shared_u32(src + r * 128 + (((lane >> 2) ^ i) << 4));  // (r & 7) == i
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_base + 8u * (kTmaLoadStages + st));  // this warp is done with the stage
            // output box of pol p: row t = lane, 128 B = 64 antennas x (re, im); chunk g of row t sits at chunk g ^ (t & 7)
            const uint32_t dst = out_base + os * kTmaTileBytes + lane * 128 + ((g ^ (lane & 7)) << 4);
            st_shared_v4(dst, __byte_perm(w[0], w[1], 0x5410u), __byte_perm(w[2], w[3], 0x5410u),
                         __byte_perm(w[4], w[5], 0x5410u), __byte_perm(w[6], w[7], 0x5410u));
            st_shared_v4(dst + kTmaTileBytes / 2, __byte_perm(w[0], w[1], 0x7632u), __byte_perm(w[2], w[3], 0x7632u),
                         __byte_perm(w[4], w[5], 0x7632u), __byte_perm(w[6], w[7], 0x7632u));
            fence_proxy_async_smem();
            named_bar_sync(2, kTmaConsumerWarps * 32);
            if (threadIdx.x == 0) {
                int b, c, t0, a0;
                decode(tile, &b, &c, &t0, &a0);
                const int plane0 = (b * kPols) * prm.C + c;
                tma_store_3d(&tm_out, out_base + os * kTmaTileBytes, 2 * a0, t0, plane0);
                tma_store_3d(&tm_out, out_base + os * kTmaTileBytes + kTmaTileBytes / 2, 2 * a0, t0, plane0 + prm.C);
                bulk_commit_group();
            }
        }
        if (threadIdx.x == 0) bulk_wait_group_all();
    }
}

}  // namespace

static int launch_reorder_tma(const uint8_t* samples, uint8_t* reordered, int B, int A, int C, int T, cudaStream_t s) {
    EncodeTiledFn encode = nullptr;
    if (int e = get_encode_fn(&encode)) return e;
    alignas(64) CUtensorMap tm_in, tm_out;
    {
        // samples as 4-byte words {p0.re, p0.im, p1.re, p1.im}: [B][A][C][T], box [1][64][1][32], 128B swizzle
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(A),
                                    static_cast<cuuint64_t>(B)};
        const cuuint64_t strides[3] = {static_cast<cuuint64_t>(T) * 4, static_cast<cuuint64_t>(C) * T * 4,
                                       static_cast<cuuint64_t>(A) * C * T * 4};
        const cuuint32_t box[4] = {kTmaT, 1, kTmaAnts, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        if (encode(&tm_in, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<uint8_t*>(samples), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return record_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(reorder in)");
    }
    {
        // reordered as bytes [B*2*C][T][2A], box [1][32][128], 128B swizzle
        const cuuint64_t dims[3] = {static_cast<cuuint64_t>(2 * A), static_cast<cuuint64_t>(T),
                                    static_cast<cuuint64_t>(B) * kPols * C};
        const cuuint64_t strides[2] = {static_cast<cuuint64_t>(2 * A), static_cast<cuuint64_t>(T) * 2 * A};
        const cuuint32_t box[3] = {2 * kTmaAnts, kTmaT, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (encode(&tm_out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, reordered, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return record_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(reorder out)");
    }
    ReorderTmaParams p{};
    p.B = B, p.C = C;
    p.a_tiles = (A + kTmaAnts - 1) / kTmaAnts;
    p.t_tiles = (T + kTmaT - 1) / kTmaT;
    p.n_tiles = static_cast<long long>(B) * C * p.t_tiles * p.a_tiles;
    static int n_sms = 0;
    if (!n_sms) {
        int dev = 0;
        DCBF_CUDA_TRY(cudaGetDevice(&dev));
        DCBF_CUDA_TRY(cudaFuncSetAttribute(reorder_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmem));
        DCBF_CUDA_TRY(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long want = 2LL * n_sms;
    const unsigned grid = static_cast<unsigned>(p.n_tiles < want ? p.n_tiles : want);
    reorder_tma_kernel<<<grid, kTmaThreads, kTmaSmem, s>>>(p, tm_in, tm_out);
    DCBF_CHECK_LAUNCH("reorder_tma_kernel");
    return DCBF_OK;
}

int launch_reorder(const uint8_t* samples, uint8_t* reordered, int B, int A, int C, int T, cudaStream_t s) {
    // TMA form: 16-byte output rows (A % 8 == 0) and tensor-map extents within range
    if ((A & 7) == 0 && static_cast<long long>(B) * kPols * C <= 0x7fffffffLL && !std::getenv("DCBF_REORDER_NO_TMA"))
        return launch_reorder_tma(samples, reordered, B, A, C, T, s);
    // Power-of-two tile (in samples, 16..128): long contiguous DRAM runs per antenna, but small enough (<= 36 KiB of
    // staging) that six CTAs share an SM; huge arrays take whatever still fits.
    int tile_t = 128;
    while (tile_t > 16 && tile_t / 2 >= T) tile_t >>= 1;  // no point in a tile longer than the heap
    auto smem_for = [&](int tt) { return static_cast<size_t>(A) * tt * sizeof(uint32_t); };
    const size_t kGoodSmem = 36 * 1024, kMaxSmem = 200 * 1024;
    while (tile_t > 64 && smem_for(tile_t) > kGoodSmem) tile_t >>= 1;
    while (tile_t > 16 && smem_for(tile_t) > kMaxSmem) tile_t >>= 1;
    if (smem_for(tile_t) > kMaxSmem) return DCBF_ERR_UNSUPPORTED;
    const int n_tiles = (T + tile_t - 1) / tile_t;
    const long long n_blocks = static_cast<long long>(B) * C * n_tiles;
    if (n_blocks > 0x7fffffffLL) return DCBF_ERR_UNSUPPORTED;
    const size_t smem = smem_for(tile_t);
    if (smem > 48 * 1024)
        DCBF_CUDA_TRY(cudaFuncSetAttribute(reorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
    // e / A for the element indices of a tile (e < tile_t * A <= 51200) through one multiply; A = 1: e / 1 = e
    const uint32_t inv_a = A == 1 ? 0u : static_cast<uint32_t>((1ull << 32) / static_cast<unsigned>(A)) + 1u;
    reorder_kernel<<<static_cast<unsigned>(n_blocks), kThreads, smem, s>>>(samples, reordered, A, C, T, tile_t,
                                                                           n_tiles, inv_a);
    DCBF_CHECK_LAUNCH("reorder_kernel");
    return DCBF_OK;
}

}  // namespace dcbf
