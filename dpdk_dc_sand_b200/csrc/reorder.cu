// Stage 1 (stand-alone): pre-beamform reorder, bit-exact.
//
//   samples   [B][A][C][T][P=2][X=2] u8  ->  reordered [B][P][C][T/16][16][A][X] u8
//
// Replaces kernel `prebeamform_reorder`
// (reference: beamformer/beamforming/kernels/prebeamform_reorder_kernel.mako:37-92), which moves one
// 16-bit word per thread and scatters its stores at a stride of A words.  Here one CTA stages an
// [A antennas] x [<=64 samples] tile through shared memory:
//   - loads: each antenna row of the tile is 4*tt contiguous bytes -> 128-bit coalesced loads;
//   - stores: for a fixed pol the tile's output is ONE contiguous run of tt*A*2 bytes
//     ([t][a][x] with t the tile's samples) -> 128-bit, fully coalesced stores;
//   - the transpose itself is a 2-byte gather from shared memory (row pitch tt+1 words keeps the
//     gather at <=2-way bank conflicts for every antenna count, including odd ones).
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
               int n_tiles) {
    extern __shared__ uint32_t raw[];  // [A][tt + 1] words; word (a, t) = {p0.re, p0.im, p1.re, p1.im}

    const long long blk = blockIdx.x;
    const int tile = static_cast<int>(blk % n_tiles);
    const int c = static_cast<int>((blk / n_tiles) % C);
    const long long b = blk / (static_cast<long long>(n_tiles) * C);
    const int t0 = tile * tile_t;
    const int tt = min(tile_t, T - t0);  // multiple of 16
    const int pitch = tt + 1;

    // ---- coalesced 128-bit loads of the [A][tt] tile ----
    const int vec_per_row = tt >> 2;
    const int n_vec = A * vec_per_row;
    for (int i = threadIdx.x; i < n_vec; i += kThreads) {
        const int a = i / vec_per_row;
        const int v = i - a * vec_per_row;
        const size_t row = ((static_cast<size_t>(b) * A + a) * C + c) * static_cast<size_t>(T) + t0;
        const uint4 x = ldg_stream(reinterpret_cast<const uint4*>(in + row * 4) + v);
        uint32_t* dst = raw + a * pitch + 4 * v;
        dst[0] = x.x;
        dst[1] = x.y;
        dst[2] = x.z;
        dst[3] = x.w;
    }
    __syncthreads();

    // ---- per pol: tt*A elements of 2 bytes, contiguous in the output; 8 elements per 128-bit store ----
    const int n_chunk = (tt * A) >> 3;  // tt % 8 == 0
#pragma unroll
    for (int p = 0; p < kPols; ++p) {
        const size_t base = (((static_cast<size_t>(b) * kPols + p) * C + c) * static_cast<size_t>(T) + t0) * A * 2;
        uint4* dst = reinterpret_cast<uint4*>(out + base);
        const uint32_t sel = p ? 0x7632u : 0x5410u;  // pick the high/low 16 bits of two words
        for (int i = threadIdx.x; i < n_chunk; i += kThreads) {
            const int e = i << 3;
            int t = e / A;
            int a = e - t * A;
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                w[j] = raw[a * pitch + t];
                if (++a == A) {
                    a = 0;
                    ++t;
                }
            }
            uint4 v;
            v.x = __byte_perm(w[0], w[1], sel);
            v.y = __byte_perm(w[2], w[3], sel);
            v.z = __byte_perm(w[4], w[5], sel);
            v.w = __byte_perm(w[6], w[7], sel);
            stg_stream(dst + i, v);
        }
    }
}

}  // namespace

int launch_reorder(const uint8_t* samples, uint8_t* reordered, int B, int A, int C, int T, cudaStream_t s) {
    // Largest tile (in samples) whose staging buffer fits in shared memory.
    int tile_t = 64;
    auto smem_for = [&](int tt) { return static_cast<size_t>(A) * (tt + 1) * sizeof(uint32_t); };
    const size_t kMaxSmem = 200 * 1024;
    while (tile_t > 16 && smem_for(tile_t) > kMaxSmem) tile_t >>= 1;
    if (smem_for(tile_t) > kMaxSmem) return DCBF_ERR_UNSUPPORTED;
    if (tile_t > T) tile_t = T;
    const int n_tiles = (T + tile_t - 1) / tile_t;
    const long long n_blocks = static_cast<long long>(B) * C * n_tiles;
    if (n_blocks > 0x7fffffffLL) return DCBF_ERR_UNSUPPORTED;
    const size_t smem = smem_for(tile_t);
    if (smem > 48 * 1024)
        DCBF_CUDA_TRY(cudaFuncSetAttribute(reorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
    reorder_kernel<<<static_cast<unsigned>(n_blocks), kThreads, smem, s>>>(samples, reordered, A, C, T, tile_t,
                                                                           n_tiles);
    DCBF_CHECK_LAUNCH("reorder_kernel");
    return DCBF_OK;
}

}  // namespace dcbf
