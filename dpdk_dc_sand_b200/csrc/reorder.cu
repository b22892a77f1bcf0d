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
#include "common.cuh"

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

}  // namespace

int launch_reorder(const uint8_t* samples, uint8_t* reordered, int B, int A, int C, int T, cudaStream_t s) {
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
