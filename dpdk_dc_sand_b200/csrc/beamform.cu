// Stage 3 (stand-alone): coefficient x voltage contraction over antennas on materialised operands.
//
//   out[b,p,c,t,n] = sum_{j<2A} f32(reordered[b,p,c,t,j]) * coeffs[b,p,c,j,n],   n < 2M
//
// Replaces kernel `run_complex_mult` (reference: beamformer/beamforming/complex_mult_kernel.py:11-100), which
// launches 2A threads per output row that all compute the same 2M outputs with no operand reuse.  This is
// the FALLBACK of the stand-alone MatrixMultiply op, for shapes whose rows the TMA descriptors of the tcgen05
// kernel (beamform_tc.cu) cannot address (odd beam counts: 8M-byte rows), and its cross-check
// (DCBF_FLAG_DEBUG_CUDA_CORES): float32 operands, float32 accumulation in the reference's order (j ascending,
// one accumulator per output) on the CUDA cores with a shared-memory tiled 64x64x32 scheme -- compute-bound at
// ~30 TFLOP/s.  The production path is the fused tcgen05 kernel in fused.cu, which never materialises
// `reordered` or `coeffs`.
#include "common.cuh"

namespace dcbf {

namespace {

constexpr int kBM = 64;   // samples (t) per CTA tile
constexpr int kBN = 64;   // output columns (2*beam + re/im) per CTA tile
constexpr int kBK = 32;   // contraction chunk (2*ant + re/im)
constexpr int kThreads = 256;
constexpr int kPitch = kBM + 4;  // keeps float4 alignment, spreads the transpose stores over banks

template <bool kSigned>
__global__ void __launch_bounds__(kThreads)
beamform_kernel(const uint8_t* __restrict__ data, const float* __restrict__ coeffs, float* __restrict__ out, int T,
                int K2 /*2A*/, int N2 /*2M*/, int tiles_t, int tiles_n) {
    __shared__ __align__(16) float Ds[kBK][kPitch];  // [k][t]
    __shared__ __align__(16) float Cs[kBK][kBN];     // [k][n]

    const long long blk = blockIdx.x;
    const int tn = static_cast<int>(blk % tiles_n);
    const int tt = static_cast<int>((blk / tiles_n) % tiles_t);
    const size_t bpc = static_cast<size_t>(blk / (static_cast<long long>(tiles_n) * tiles_t));
    const int t0 = tt * kBM, n0 = tn * kBN;

    const uint8_t* d_base = data + bpc * static_cast<size_t>(T) * K2;
    const float* c_base = coeffs + bpc * static_cast<size_t>(K2) * N2;
    float* o_base = out + bpc * static_cast<size_t>(T) * N2;

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};

    for (int k0 = 0; k0 < K2; k0 += kBK) {
        // data tile: 64 samples x 16 (re,im) pairs, 2-byte loads (pairs are always 2-byte aligned)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = threadIdx.x + r * kThreads;
            const int t = i >> 4, jp = i & 15;
            const int j = k0 + 2 * jp;
            float re = 0.f, im = 0.f;
            if (t0 + t < T && j < K2) {
                const uint16_t w = *reinterpret_cast<const uint16_t*>(d_base + static_cast<size_t>(t0 + t) * K2 + j);
                if (kSigned) {
                    re = static_cast<float>(static_cast<int8_t>(w & 0xff));
                    im = static_cast<float>(static_cast<int8_t>(w >> 8));
                } else {
                    re = static_cast<float>(w & 0xff);
                    im = static_cast<float>(w >> 8);
                }
            }
            Ds[2 * jp][t] = re;
            Ds[2 * jp + 1][t] = im;
        }
        // coefficient tile: 32 rows x 64 columns, lanes along n
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = threadIdx.x + r * kThreads;
            const int kk = i >> 6, n = i & 63;
            float v = 0.f;
            if (k0 + kk < K2 && n0 + n < N2) v = __ldg(c_base + static_cast<size_t>(k0 + kk) * N2 + n0 + n);
            Cs[kk][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&Ds[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Cs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    const bool vec_ok = (N2 & 3) == 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty * 4 + i;
        if (t >= T) break;
        float* row = o_base + static_cast<size_t>(t) * N2 + n0 + tx * 4;
        const int n = n0 + tx * 4;
        if (vec_ok && n + 3 < N2) {
            *reinterpret_cast<float4*>(row) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < N2) row[j] = acc[i][j];
        }
    }
}

}  // namespace

int launch_beamform(const uint8_t* reordered, const float* coeffs, float* beams, int B, int C, int T, int A, int M,
                    unsigned flags, cudaStream_t s) {
    const int K2 = 2 * A, N2 = 2 * M;
    const int tiles_t = (T + kBM - 1) / kBM, tiles_n = (N2 + kBN - 1) / kBN;
    const long long n_blocks = static_cast<long long>(B) * kPols * C * tiles_t * tiles_n;
    if (n_blocks > 0x7fffffffLL) return DCBF_ERR_UNSUPPORTED;
    if (flags & DCBF_FLAG_SIGNED_INPUT)
        beamform_kernel<true><<<static_cast<unsigned>(n_blocks), kThreads, 0, s>>>(reordered, coeffs, beams, T, K2,
                                                                                  N2, tiles_t, tiles_n);
    else
        beamform_kernel<false><<<static_cast<unsigned>(n_blocks), kThreads, 0, s>>>(reordered, coeffs, beams, T, K2,
                                                                                   N2, tiles_t, tiles_n);
    DCBF_CHECK_LAUNCH("beamform_kernel");
    return DCBF_OK;
}

}  // namespace dcbf
