// Stage 2 (stand-alone): steering-coefficient generation.
//
//   delay_vals [C][M][A][4] f32  ->  coeffs [B][P][C][2A][2M] f32
//   block (ant a, beam m):  rows 2a,2a+1 / cols 2m,2m+1 = [[cos r, sin r], [-sin r, cos r]]
//   r = delay*ch*(-pi)/(N*Ts) + phase - delay*(N/2)*(-pi)/(N*Ts),   ch = c + C*xeng_id
//
// Replaces kernel `run_coeff_gen` (reference: beamformer/beamforming/coeff_generator.py:12-103) and the
// static (rate-free) case of the native precursor's calculate_beamweights_* kernels
// (beamformer_coefficient_generator/BeamformerKernels.cu:7-189).  Indexing follows the reference's CPU
// checker (beamformer/unit_test/coeff_generator_cpu.py:125-186): delay_vals is read at [c][beam][ant] and
// written at rows 2*ant, cols 2*beam (the reference GPU kernel transposes ant/beam between its read and
// its write; the tests never see it because they use uniform delays).
//
// Arithmetic: float64 with the reference's exact operation order (explicit _rn intrinsics, no FMA
// contraction), float64 sincos, rounded once to float32 -- i.e. the same value the reference computes under
// numba/numpy-1.x promotion rules, so the result is bit-identical to the float64 oracle except where
// libm's and CUDA's double cos differ in the last ulp AND that ulp straddles a float32 rounding boundary.
//
// Memory: one CTA owns a [<= 64 beams] x [32 ants] tile of one channel.  delay_vals is read with lanes along `ant`
// (contiguous 16-byte structs -> fully coalesced), the (cos, sin) pairs are transposed through shared memory into the
// OUTPUT layout of the tile (rows 2a, 2a+1 x columns 2m, 2m+1), and every (batch, pol) replica of the tile is then
// written by bulk copies (cp.async.bulk shared -> global: one copy per tile when the tile spans whole rows, one per
// row otherwise) -- the replicated writes cost no instructions and no second pass over the registers.  Row pitches
// that are not a multiple of 16 bytes (odd beam counts; for the half-precision output beam counts not divisible by
// 4) take the plain-store path.
//
// Output type: float32 (the reference operator's slot), or -- dcbf_coeffs_f16 -- fp16, the precursor's 16-bit
// output option (beamformer_coefficient_generator/BeamformerKernels.cu:113-115, 172-185: __floats2half2_rn of the
// float values), in the same [B][P][C][2A][2M] layout.
#include <cuda_fp16.h>

#include "common.cuh"

namespace dcbf {

namespace {

constexpr int kTile = 32;
constexpr int kThreads = 256;

// Per-batch time offsets of the time-varying form (dcbf_coeffs_tv): delay -> delay + delay_rate*dt,
// phase -> phase + phase_rate*dt, after beamformer_coefficient_generator/BeamformerKernels.cu:25-35.
struct BatchTimes {
    int n;  // 0: static steering (the reference Python path ignores the rate fields)
    double dt[DCBF_MAX_TV_BATCHES];
};

// x / d, correctly rounded, for the launch-invariant divisor d = N * Ts with r = RN(1 / d) computed once: q0 = RN(x r)
// is within an ulp of the quotient, the remainder x - q0 d is exact in one FMA, and RN(q0 + rem r) is then the
// correctly rounded quotient (Markstein's theorem for a correctly rounded reciprocal) -- the same value as the
// reference's float64 division, for 3 instructions instead of the ~20 of a general IEEE division.  Operands here are
// far from overflow / underflow (|x| < 1e12, d ~ 1e-6 .. 1e-3).
__device__ __forceinline__ double div_by_denom(double x, double d, double r) {
    const double q0 = __dmul_rn(x, r);
    const double rem = __fma_rn(-q0, d, x);
    return __fma_rn(rem, r, q0);
}

// float2 / __half2 of one (cos, sin)-like pair
template <typename OutT> struct Pair;
template <> struct Pair<float> {
    using type = float2;
    static __device__ __forceinline__ float2 make(float a, float b) { return make_float2(a, b); }
};
template <> struct Pair<__half> {
    using type = __half2;
    static __device__ __forceinline__ __half2 make(float a, float b) { return __floats2half2_rn(a, b); }
};

constexpr int kTileM = 64;  // beams per tile

template <typename OutT, bool kBulk>
__global__ void __launch_bounds__(kThreads)
coeffs_kernel(const float4* __restrict__ delay_vals, OutT* __restrict__ coeffs, int n_batches, int n_pols, int C, int A,
              int M, int chan_offset, double half_n, double denom, int tiles_a, int tiles_m,
              const __grid_constant__ BatchTimes times, const float* __restrict__ weights) {
    using P2 = typename Pair<OutT>::type;
    __shared__ float2 cs[kTile][kTile + 1];                       // [beam][ant] -> (cos, sin), 32 beams at a time
    __shared__ __align__(128) P2 tile[2 * kTile][kTileM];         // [row 2a + r][beam]: the tile in the output layout

    const long long blk = blockIdx.x;
    const int ta = static_cast<int>(blk % tiles_a);
    const int tm = static_cast<int>((blk / tiles_a) % tiles_m);
    const int c = static_cast<int>(blk / (static_cast<long long>(tiles_a) * tiles_m));
    const int a0 = ta * kTile, m0 = tm * kTileM;
    const int na = min(kTile, A - a0), nm = min(kTileM, M - m0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const double ch = static_cast<double>(c + chan_offset);
    const double neg_pi = -3.141592653589793;  // == -math.pi
    const double inv_denom = __drcp_rn(denom);

    const size_t row_len = 2 * static_cast<size_t>(M);
    const int n_groups = times.n > 0 ? n_batches : 1;          // distinct coefficient sets
    const int reps = times.n > 0 ? n_pols : n_batches * n_pols;  // (batch, pol) replicas per set
    for (int g = 0; g < n_groups; ++g) {
        for (int mh = 0; mh < nm; mh += kTile) {
            // phase 1: lanes along ant
            // all four loads of this thread first (ncu: the dependent load -> float64 chain -> next load sequence left
            // the warps waiting on long_scoreboard most of the time)
            float4 dvs[kTile / (kThreads / 32)];
#pragma unroll
            for (int r = 0; r < kTile / (kThreads / 32); ++r) {
                const int m = m0 + mh + warp + r * (kThreads / 32), a = a0 + lane;
                dvs[r] = (m < M && a < A) ? __ldg(delay_vals + (static_cast<size_t>(c) * M + m) * A + a) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < kTile / (kThreads / 32); ++r) {
                const int mi = warp + r * (kThreads / 32);
                const int m = m0 + mh + mi, a = a0 + lane;
                if (m < M && a < A) {
                    const float4 dv = dvs[r];
                    double delay = static_cast<double>(dv.x);
                    double phase = static_cast<double>(dv.z);
                    if (times.n > 0) {
                        delay = __dadd_rn(delay, __dmul_rn(static_cast<double>(dv.y), times.dt[g]));
                        phase = __dadd_rn(phase, __dmul_rn(static_cast<double>(dv.w), times.dt[g]));
                    }
                    // ((delay*ch)*(-pi))/(N*Ts) + phase      coeff_generator_cpu.py:143-150
                    const double initial = __dadd_rn(div_by_denom(__dmul_rn(__dmul_rn(delay, ch), neg_pi), denom, inv_denom), phase);
                    // ((delay*(N/2))*(-pi))/(N*Ts)           coeff_generator_cpu.py:155-160
                    const double centre = div_by_denom(__dmul_rn(__dmul_rn(delay, half_n), neg_pi), denom, inv_denom);
                    const double rot = __dsub_rn(initial, centre);
                    double sn, cn;
                    sincos(rot, &sn, &cn);
                    if (weights) {  // real per-(beam, antenna) weight (?beam-weights), applied in float64 before rounding
                        const double w = static_cast<double>(__ldg(weights + static_cast<size_t>(m) * A + a));
                        sn = __dmul_rn(sn, w);
                        cn = __dmul_rn(cn, w);
                    }
                    cs[mi][lane] = make_float2(static_cast<float>(cn), static_cast<float>(sn));
                }
            }
            __syncthreads();

            // phase 2: lanes along beam; rows 2a (cos, sin) and 2a+1 (-sin, cos)
            if (mh + lane < nm) {
                for (int ai = warp; ai < na; ai += kThreads / 32) {
                    const float2 v = cs[lane][ai];
                    const P2 r0 = Pair<OutT>::make(v.x, v.y), r1 = Pair<OutT>::make(-v.y, v.x);
                    if (kBulk) {
                        tile[2 * ai][mh + lane] = r0;
                        tile[2 * ai + 1][mh + lane] = r1;
                    } else {
                        for (int rep = 0; rep < reps; ++rep) {
                            OutT* base = coeffs + ((static_cast<size_t>(g * reps + rep) * C + c) * (2 * static_cast<size_t>(A)) + 2 * (a0 + ai)) * row_len;
                            reinterpret_cast<P2*>(base)[m0 + mh + lane] = r0;
                            reinterpret_cast<P2*>(base + row_len)[m0 + mh + lane] = r1;
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (kBulk) {
            // the tile -> every replica.  Rows of the tile are kTileM pairs apart in shared memory; in global memory they
            // are row_len elements apart, i.e. contiguous with the tile's rows exactly when the tile spans the whole row.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            const uint32_t row_bytes = static_cast<uint32_t>(nm) * sizeof(P2);
            const bool whole = nm == kTileM && M == kTileM;  // one copy covers all 2 * na rows
            if (whole) {
                if (threadIdx.x < reps) {
                    OutT* dst = coeffs + ((static_cast<size_t>(g * reps + threadIdx.x) * C + c) * (2 * static_cast<size_t>(A)) + 2 * a0) * row_len;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                                 "r"(static_cast<uint32_t>(__cvta_generic_to_shared(&tile[0][0]))), "r"(2u * na * row_bytes)
                                 : "memory");
                }
            } else {
                for (int i = threadIdx.x; i < 2 * na * reps; i += kThreads) {
                    const int rep = i / (2 * na), row = i - rep * (2 * na);
                    OutT* dst = coeffs + ((static_cast<size_t>(g * reps + rep) * C + c) * (2 * static_cast<size_t>(A)) + 2 * a0 + row) * row_len + 2 * m0;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                                 "r"(static_cast<uint32_t>(__cvta_generic_to_shared(&tile[row][0]))), "r"(row_bytes)
                                 : "memory");
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile is rewritten for the next set / released at exit
            __syncthreads();
        }
    }
}

}  // namespace

template <typename OutT>
static int launch_coeffs_t(const float* delay_vals, OutT* coeffs, int B, int P, int C, int N, int A, int M, int xeng_id,
                           double sample_period, const double* batch_dt_s, const float* beam_weights, cudaStream_t s) {
    BatchTimes times{};
    if (batch_dt_s) {
        if (B > DCBF_MAX_TV_BATCHES) return DCBF_ERR_UNSUPPORTED;
        times.n = B;
        for (int b = 0; b < B; ++b) times.dt[b] = batch_dt_s[b];
    }
    const int tiles_a = (A + kTile - 1) / kTile, tiles_m = (M + kTileM - 1) / kTileM;
    const long long n_blocks = static_cast<long long>(C) * tiles_a * tiles_m;
    if (n_blocks > 0x7fffffffLL) return DCBF_ERR_UNSUPPORTED;
    const double denom = static_cast<double>(N) * sample_period;  // python: (n_channels * sample_period)
    const double half_n = static_cast<double>(N) / 2.0;           // python: (n_channels / 2)
    // bulk copies need 16-byte aligned addresses and sizes: the row pitch 2 M elements and every tile's row segment
    const bool bulk = (2 * static_cast<size_t>(M) * sizeof(OutT)) % 16 == 0 && (reinterpret_cast<uintptr_t>(coeffs) & 15) == 0;
    auto kernel = bulk ? coeffs_kernel<OutT, true> : coeffs_kernel<OutT, false>;
    kernel<<<static_cast<unsigned>(n_blocks), kThreads, 0, s>>>(
        reinterpret_cast<const float4*>(delay_vals), coeffs, B, P, C, A, M, C * xeng_id, half_n, denom, tiles_a, tiles_m,
        times, beam_weights);
    DCBF_CHECK_LAUNCH("coeffs_kernel");
    return DCBF_OK;
}

int launch_coeffs(const float* delay_vals, float* coeffs, int B, int P, int C, int N, int A, int M, int xeng_id,
                  double sample_period, const double* batch_dt_s, const float* beam_weights, cudaStream_t s) {
    return launch_coeffs_t<float>(delay_vals, coeffs, B, P, C, N, A, M, xeng_id, sample_period, batch_dt_s, beam_weights, s);
}

int launch_coeffs_f16(const float* delay_vals, void* coeffs, int B, int P, int C, int N, int A, int M, int xeng_id,
                      double sample_period, const double* batch_dt_s, const float* beam_weights, cudaStream_t s) {
    return launch_coeffs_t<__half>(delay_vals, static_cast<__half*>(coeffs), B, P, C, N, A, M, xeng_id, sample_period, batch_dt_s,
                                   beam_weights, s);
}

}  // namespace dcbf
