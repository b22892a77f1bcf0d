// Stages 1+2+3 fused: reorder + steering coefficients + contraction over antennas, one pass over HBM.
//
//   samples [B][A][C][T][2][2] u8,  delay_vals [C][M][A][4] f32  ->  beams [B][2][C][T][2M] f32
//
// Replaces the three back-to-back launches of OpSequence.__call__
// (reference: beamformer/beamforming/beamform_op_sequence.py:141-154 ->
//  kernels/prebeamform_reorder_kernel.mako:37-92, coeff_generator.py:12-103, complex_mult_kernel.py:11-100)
// and the native precursor's fused single-pol kernel
// (beamformer_coefficient_generator/BeamformerKernels.cu:192-367).  Neither the reordered voltages nor the
// coefficients ever exist in HBM: the reordered layout [t][2a+x] IS the K-major A operand of the GEMM and
// the coefficient block [[cos,sin],[-sin,cos]] IS its B operand, both built in shared memory.
//
// Per (batch b, pol p, channel c):  D[T x 2M] = X[T x 2A] * W[2A x 2M]   (fp16 operands, fp32 accumulate)
//   X[t][2a+x]   = f16(sample[b][a][c][t][p][x] / 1024)   exact: |byte| <= 255
//   W[2a+x][2m+y] = 1024 * {{cos, sin}, {-sin, cos}}[x][y] of rot(c, m, a), carried as fp16 hi + fp16 lo
//                  (two accumulating MMAs; coefficient error ~2^-24) or fp16 hi only (DCBF_FLAG_FP16_COEFF);
//                  the two powers of two cancel in the product and keep small coefficients out of fp16 subnormals.
//
// One persistent CTA per SM takes channels from a dynamic queue.  Warp roles (608 threads):
//   warps 0-7     coeffs    : delay_vals (coalesced float4, one batch prefetched in registers) -> phase to float64
//                             accuracy in float pairs -> sincospi -> fp16 hi/lo -> 128B-swizzled B tiles, one
//                             channel ahead; lane 0 of warp 0 also draws the CTA's channels from the global queue
//   warps 8-11    convert   : raw bytes -> fp16, pol de-interleave, a<->t transpose into 64B-swizzled A tiles
//   warps 12-15   epilogue  : tcgen05.ld 32x32b -> 128B-swizzled staging tile -> TMA tensor store (32x32 boxes)
//                             (odd beam counts / ragged N tiles: 16x256b -> st.global.v2 from registers)
//   warp 16       producer  : one tensor-map TMA box [16 ant][128 samples][4 B] per slab -> raw ring
//   warps 17, 18  MMA       : one lane each issues the tcgen05.mma (M=128, N<=128, K=16) of one pol, accumulators in TMEM
// Pipelines (mbarrier full/empty pairs): raw ring (TMA->convert), A ring (convert->MMA), B double buffer
// (coeffs->MMA), TMEM accumulator double buffer (MMA->epilogue), per-warp staging pairs (bulk groups).
//
// Tiling: time tiles of 128 samples (UMMA M), slabs of 16 antennas (K = 32 = two MMA K-steps), B k-blocks of
// 32 antennas (one 128-byte swizzle row), N tiles of <=128 columns chosen so that one B tile set (all k-blocks,
// hi+lo) fits 64 KiB.  With more than one N tile the voltages of a channel are re-read (from L2).  Tiles of
// <= 64 columns run hi and lo as ONE N = 2 nt MMA (kMerged).
//
// Template specialisations: kProf (per-role blocked-time accounting, tools/prof_roles.py), kTv (per-heap delay /
// phase rates, dcbf_fused_tv), kQ8 (int8 requantised output, dcbf_fused_q8), kMerged, and kStream: when one B tile
// set would need several N tiles (many antennas x beams) or does not fit at all (> 512 antennas), B is streamed
// through a ring of 32-antenna k-blocks instead, with N tiles of up to 128 columns, two time tiles (all four
// (time tile, pol) accumulators) per unit, and (channel, N tile, time-tile pair) units drawn from the queue.
// Raw ring: 4 stages of their own plus, for narrow N tiles, up to 4 more in the unused tail of the B buffers.
#include <cuda.h>

#include <atomic>
#include <cmath>
#include <mutex>
#include <type_traits>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dcbf {

namespace {

// Warp roles, one or more whole warpgroups (4 warps) each so that every role can resize its register allocation
// with setmaxnreg: the kernel is launched with 72 registers per thread (28 warps), the producer / MMA group and
// convert drop to what they need, and the epilogue grows from what they released.
// The SM sub-partition arbiter favours the highest warp id, so ids are handed out in the order of urgency: the MMA
// issuers and the TMA producer (a few instructions each, everything else waits on them) on top, then the epilogue
// (drains TMEM into the dominant HBM stream), convert, and the coefficient warps (longest job, but needed a whole
// unit later) at the bottom.  Sixteen coefficient warps: every one of them is stalled ~85 % of its cycles (loads,
// dependent FMA chains, shared-memory stores -- ncu), so the role's rate scales with its warp count, and it is the
// role that bounds short launches (first B tile set) and many antennas x beams.
#ifndef DCBF_COEFF_WARPS
#define DCBF_COEFF_WARPS 16
#endif
constexpr int kCoeffWarp0 = 0, kCoeffWarps = DCBF_COEFF_WARPS;  // warps 0..15
constexpr int kConvertWarp0 = kCoeffWarps;                       // warps 16..19
constexpr int kEpilogueWarp0 = kCoeffWarps + 4;                  // warps 20..23 (warp % 4 = TMEM lane quarter)
constexpr int kProducerWarp = kCoeffWarps + 8;
constexpr int kMmaWarp = kCoeffWarps + 9;     // issues the pol-0 MMAs (and owns the TMEM allocation)
constexpr int kMmaWarp2 = kCoeffWarps + 10;   // issues the pol-1 MMAs: the issue path, not the tensor pipe, limits narrow tiles
constexpr int kThreads = (kCoeffWarps + 12) * 32;  // the last warp only keeps the last warpgroup whole
// int8 output: a second epilogue warpgroup, warps 28..31, one group per pol.  Quantising costs ~3 ALU instructions per
// value on top of the float32 epilogue, and the four epilogue warps share their schedulers with the coefficient role:
// they were busy 275 of 290 us at C3 while every other role waited for them.  The kernel then starts from 64 registers
// per thread (32 warps) and splits 72 (coefficients) / 72 (epilogue) / 40 / 40 -- or, in the build for whole tile sets of
// more than 64 columns (kQ8Wide: C3), 80 / 56 / 40 / 40: the coefficient role bounds that kernel, and its spills go
// to L2 (all of L1 is carved out for shared memory): an `ncu` capture had 20 % of its stall samples on local-memory
// reloads.  C3: 242.6 -> 220.4 us on the same box.  The merged-tile (<= 64 columns) and K-streamed builds keep 72 / 72:
// their epilogues need the registers (+9 ... +12 % with 56).
constexpr int kEpilogue2Warp0 = kCoeffWarps + 12;
constexpr int kThreadsQ8 = kThreads + 128;
static_assert(kThreadsQ8 * 64 <= 65536 && kCoeffWarps * 72 + 8 * 72 + 4 * 40 + 4 * 40 <= (kCoeffWarps + 16) * 64 &&
              kCoeffWarps * 80 + 8 * 56 + 4 * 40 + 4 * 40 <= (kCoeffWarps + 16) * 64, "register pool, int8 output");
// registers per thread at launch (the __launch_bounds__ cap) and per role after setmaxnreg; the roles' sum must fit the
// CTA's pool of kThreads * kRegsLaunch
#if DCBF_COEFF_WARPS == 16
#define DCBF_REGS_LAUNCH 72
#define DCBF_REGS_COEFF 72
#define DCBF_REGS_CONVERT 56
#define DCBF_REGS_EPILOGUE 104
#define DCBF_REGS_ISSUE 56
#else
#define DCBF_REGS_LAUNCH 96
#define DCBF_REGS_COEFF 96
#define DCBF_REGS_CONVERT 72
#define DCBF_REGS_EPILOGUE 144
#define DCBF_REGS_ISSUE 72
#endif
static_assert(kCoeffWarps % 4 == 0 && kThreads * DCBF_REGS_LAUNCH <= 65536, "launch register file");
static_assert(kCoeffWarps * DCBF_REGS_COEFF + 4 * (DCBF_REGS_CONVERT + DCBF_REGS_EPILOGUE + DCBF_REGS_ISSUE) <= (kCoeffWarps + 12) * DCBF_REGS_LAUNCH,
              "register pool of the CTA");
#define DCBF_STR2(x) #x
#define DCBF_STR(x) DCBF_STR2(x)
// sleep between failed probes of the waits that are a whole buffer ahead of their consumer (tc_common.cuh: mbar_wait_slow)
#ifndef DCBF_BACKOFF_NS
#define DCBF_BACKOFF_NS 200
#endif
constexpr int kTileT = 128;    // samples per MMA tile (UMMA M)
constexpr int kSlabAnts = 16;  // antennas per raw slab / A stage
constexpr int kKbAnts = 32;    // antennas per B k-block (128-byte swizzle row of fp16)
constexpr int kRawStages = 4;      // raw stages with their own shared memory ...
constexpr int kMaxRawStages = 8;   // ... plus up to 4 more in the unused tail of the two B buffers (narrow N tiles)
constexpr int kAopStages = 2;      // A stages with their own shared memory ...
constexpr int kMaxAopStages = 4;   // ... plus, with K-streamed B, two more behind a B ring of three (not four) slots
constexpr int kBopBufs = 2;
constexpr int kBopSlots = 4;    // kStream: B k-block ring (4 x 32 KiB) instead of 2 whole tile sets
constexpr int kBopSlotBytes = kBopBufs * 64 * 1024 / kBopSlots;
constexpr int kAccBufs = 2;
constexpr int kRawStageBytes = kSlabAnts * kTileT * 4;  // 8 KiB: [ant][t][pol][re,im]
constexpr int kAopTileBytes = kTileT * 64;              // 8 KiB: [t][32 fp16], 64B swizzle
constexpr int kAopStageBytes = 2 * kAopTileBytes;       // pol 0 + pol 1
constexpr int kBopBufBytes = 64 * 1024;
constexpr int kOutBoxBytes = 32 * 128;                  // 32 rows x 32 fp32 columns, 128B swizzle
constexpr int kOutStageBytes = 4 * 2 * kOutBoxBytes;    // 4 epilogue warps x 2 boxes
constexpr int kTmemCols = 512;

constexpr int kSmemData = kAopStages * kAopStageBytes + kBopBufs * kBopBufBytes + kRawStages * kRawStageBytes + kOutStageBytes;
constexpr int kCtlBytes = 640;  // 320 B of mbarriers + Control
constexpr int kSmemBytes = 1024 /*alignment slack*/ + kSmemData + kCtlBytes;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

enum Role { kRoleProducer = 1, kRoleMma = 2, kRoleEpilogue = 3, kRoleConvert = 4, kRoleCoeff = 5 };

struct FusedParams {
    const float4* dv;
    const uint8_t* samples;         // (the kernel reads them through tm_in; only the early L2 prefetch uses the pointer)
    float* out;
    int8_t* out_q8;                 // non-null: requantised int8 beams instead of float32 (dcbf_fused_q8)
    const float* gains;             // [M] per-beam quantisation gain (q8 only)
    const float* weights;           // optional [M][A] real per-(beam, antenna) weights folded into the coefficients
    unsigned long long* saturated;  // optional count of clipped values (q8 only)
    int* status;  // [0]=error code, [1]=role, [2]=barrier id, [3]=blockIdx
    int* sched;   // [0]=next channel counter (beyond the first gridDim.x), [1]=finished CTAs; self-resetting
    unsigned long long* prof;  // optional [grid][6 roles][4]: ns blocked per barrier class, [..][3] = role span
    int B, A, C, T, M;
    int slab_count;  // ceil(A / 16)
    int kb_count;    // ceil(A / 32)
    int nt;          // columns per N tile (multiple of 16, <= 128)
    int nt_count;    // number of N tiles
    int ht_count;    // ceil(T / 128)
    int parts;       // 2 = fp16 hi+lo coefficients, 1 = fp16 hi only
    int signed_in;
    // byte -> fp16 conversion of the voltages: byte b under the exponent byte `a_magic` is 2^(E-15) (1 + b / 1024), minus
    // `a_bias` = 2^(E-15) (times 1 + 128/1024 for int8 input after the ^0x80 re-bias) leaves b * 2^(E-25); E = 15 + the
    // caller's beam_weights_log2 (0: the plain 2^-10).  w_scale = 2^-beam_weights_log2 goes onto the weights.
    uint32_t a_magic, a_bias;
    float w_scale;
    int tma_store;   // 1: epilogue through shared memory + TMA tensor stores; 0: st.global from registers
    int merged;      // hi and lo coefficient rows form ONE N = 2 nt tile per MMA (nt <= 64); the epilogue adds the halves
    int q8_wide;     // q8 only: the N tile is the whole output row (32 / 64 / 128 bytes): one box per 32 rows
    int hg_count;    // kStream: groups of <= hg_size time tiles per (channel, N tile)
    int hg_size;     // 2, or 1 when every time tile has its own coefficient set (sub-heap time-varying steering)
    int sb_count;    // coefficient sets per (channel, N tile): 1; B with per-heap times; B * ht_count with per-tile times
    int set_tiles;   // accumulator tiles (batch, time tile) sharing one coefficient set: B * ht_count, ht_count or 1
    int raw_stages;  // depth of the raw TMA ring: kRawStages + extra stages placed behind the B tiles
    // whole-tile-set mode: the first n_whole channels are units of their own; each of the remaining C - n_whole
    // channels is cut into `split` units along its list of tile_count accumulator tiles (N tile, batch, time tile), so
    // that the last round of the persistent CTAs is a fraction of a channel instead of a whole one
    int n_whole, split, tile_count;
    // bsplit = 2: the pieces of a cut channel come in pairs that take one half of the N tile's beams each (first, because
    // halves of the beams also halve the coefficient work of a piece; pieces along the tile list repeat it)
    // (Tried on top: every CTA's FIRST channel as two such halves too, so that the first MMAs of a launch wait for half a
    // tile set of coefficients -- 41.6 -> 44.0 us at the 512-channel share of C3, 261.5 -> 263.1 us at C3: the second
    // conversion and the narrower MMAs of that channel cost more than the earlier start gains.)
    int bsplit;
    int dbg;         // developer experiments: 1 = no delay_vals loads, 2 = no phase / sin-cos arithmetic (and no B stores), 4 = no output stores,
                     // 8 = no L2 prefetches of delay_vals (C3: 268 -> 290 us without them; C5 share: no difference),
                     // 16 = half of the u8 -> fp16 conversion work, 32 = no epilogue staging stores / proxy fence,
                     // 64 = one MMA per tile (the last three on top of 4: which role carries the SM-side time)
    int pdl_wait;    // 1: wait for the preceding kernel of the stream (griddepcontrol.wait) after the prologue
    // Packed steering coefficients (whole-tile-set mode, static steering): the B tile set of a channel -- exactly the bytes
    // the coefficient role leaves in a B buffer -- kept in HBM between delay-model updates.  1: generate and write them
    // (no voltages are read, no beams written); 2: the hot path loads them with one bulk copy per channel instead of
    // evaluating 4096 phases and sin/cos pairs per channel and heap.
    int packed_mode;
    int packed_bytes;       // bytes of one channel's tile set (kb_count * parts * nt * 128)
    uint8_t* packed;        // [C][packed_bytes]
    int raw_extra_off;  // byte offset of the first extra stage inside each 64 KiB B buffer
    int aop_extra;      // whole tile sets: 1 = A stages 2 and 3 in the last 16 KiB of the two B buffers
    uint32_t inv_a;     // floor(2^32 / A) + 1: e / A == umulhi(e, inv_a) for every entry index of an N tile (A >= 2)
    double dt_s[DCBF_MAX_TV_BATCHES];  // time-varying steering: time offset (s) of every batch (heap)
    double sample_dt;                  // ... and seconds per sample inside a heap: != 0 gives every 128-sample time tile its
                                       // own coefficient set, evaluated at the tile's centre
    double chan_centre;      // absolute index of local channel 0, minus N/2
    double turns_per_delay;  // -1 / (N * Ts): half-turns of phase per (second of delay x channel offset)
};

// The two MMA operands carry a power-of-two scale that cancels in the product: the voltages enter as x * 2^-10 (exact
// in fp16: |x| <= 255), the steering coefficients as c * 2^10, split hi + lo.  A coefficient near zero (cos of a phase
// next to pi/2, a small beam weight) would otherwise put its hi part, and every lo part below 2^-3, into the fp16
// subnormal range, where the absolute resolution is stuck at 6e-8 instead of following the value: with the scale the
// pair resolves 2^-22 relative down to |c| ~ 1e-4 and 6e-11 absolute below that.  The products and the fp32 sums in
// TMEM are bit-for-bit what they would be without the scale wherever nothing was subnormal.
constexpr float kCoefScale = 1024.0f;

// u8 (or i8) pair -> half2 of x / 1024, exact.  `w` holds {p0.re, p0.im, p1.re, p1.im}; sel picks the pol.
__device__ __forceinline__ uint32_t bytes_to_half2(uint32_t w, uint32_t sel, uint32_t bias, uint32_t magic) {
    // bytes -> 0x3Cbb = 1 + b / 1024 (fp16), then subtract 1 (u8) or 1 + 128 / 1024 (i8 after the ^0x80 re-bias)
    const uint32_t h = __byte_perm(w, magic, sel);
    const __half2 r = __hsub2(*reinterpret_cast<const __half2*>(&h), *reinterpret_cast<const __half2*>(&bias));
    return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr double kInvPi = 0.318309886183790671538;
constexpr float kInvPiHi = static_cast<float>(kInvPi);
constexpr float kInvPiLo = static_cast<float>(kInvPi - static_cast<double>(kInvPiHi));
constexpr float kRint = 12582912.0f;  // 1.5 * 2^23: x + kRint - kRint = rint(x), and the low mantissa bits of x + kRint hold it

// Steering phase in half-turns:  x = delay * scale + phase / pi   with   scale = (ch - N/2) * (-1 / (N Ts))
// (reference: beamformer/unit_test/coeff_generator_cpu.py:143-165), delivered as a quadrant q = rint(2x) (its low two
// bits are what matter) and the remainder t = x - q/2 in [-1/4, 1/4].  The product delay * scale reaches tens to
// thousands of half-turns, so it is evaluated to float64 accuracy WITHOUT float64 instructions: scale = s_hi + s_lo
// and 1/pi are split into float pairs, products carry their exact FMA residuals, the big term is reduced mod 2
// exactly, and the rounding of the sum r0 + phase/pi is captured by a two-sum.  Angle error ~1e-7 rad.  (A leaner
// variant without the two-sum and the residual of phase/pi -- 6 instructions fewer, 2.8e-7 rad -- was measured: no
// faster at any configuration, the arithmetic of this role is off the critical path, and it cost four more cases of the
// reference's own op-sequence test, whose atol = 1e-4 sits ~1e-7 sum|x| above zero where the antennas cancel.)
// With kPair the delay and the phase are float pairs themselves (time-varying steering).
// Returns false when an operand is outside the range these tricks cover (|delay * scale| or |phase / pi| >= 2^20
// half-turns, i.e. milliseconds of delay); the caller then redoes the entry with steer_tq_f64.
template <bool kPair>
__device__ __forceinline__ bool steer_tq_fast(float d_hi, float d_lo, float ph_hi, float ph_lo, float s_hi, float s_lo, float* t,
                                              int* q) {
    const float p = d_hi * s_hi;
    const float u = ph_hi * kInvPiHi;
    float e = fmaf(d_hi, s_hi, -p);  // exact residual of p
    e = fmaf(d_hi, s_lo, e);
    if (kPair) e = fmaf(d_lo, s_hi, e);
    e += fmaf(ph_hi, kInvPiHi, -u);  // exact residual of u
    e = fmaf(ph_hi, kInvPiLo, e);
    if (kPair) e = fmaf(ph_lo, kInvPiHi, e);
    const float qf = fmaf(p, 0.5f, kRint) - kRint;  // rint(p / 2)
    const float r0 = fmaf(qf, -2.0f, p);            // p mod 2 in [-1, 1], exact
    const float s1 = r0 + u;
    const float bb = s1 - r0;
    const float err = (r0 - (s1 - bb)) + (u - bb);  // rounding error of s1 (two-sum)
    const float z = fmaf(s1, 2.0f, kRint);
    *q = __float_as_int(z);
    *t = fmaf(z - kRint, -0.5f, s1) + (e + err);
    return (fabsf(p) < 1048576.0f) & (fabsf(u) < 1048576.0f);
}
// The same in float64, for operands outside the range of steer_tq_fast (nothing physical; any finite input works).
// (Out of line and by value: a pointer result would force the callers' (t, q) arrays into local memory.)
struct Tq {
    float t;
    int q;
};
template <bool kPair>
__device__ __noinline__ Tq steer_tq_f64(float d_hi, float d_lo, float ph_hi, float ph_lo, double scale) {
    const double dd = kPair ? static_cast<double>(d_hi) + static_cast<double>(d_lo) : static_cast<double>(d_hi);
    const double pp = kPair ? static_cast<double>(ph_hi) + static_cast<double>(ph_lo) : static_cast<double>(ph_hi);
    const double x = fma(dd, scale, pp * kInvPi);
    const double xr = x - 2.0 * rint(0.5 * x);  // [-1, 1]
    const double k = rint(2.0 * xr);
    Tq r;
    r.t = static_cast<float>(fma(k, -0.5, xr));
    r.q = static_cast<int>(k);
    return r;
}

// kCoefScale * (sin, cos)(pi (t + q/2)) for t in [-1/4, 1/4]: odd / even Taylor polynomials in t of degree 9 / 10
// (truncation < 2e-9 and 2e-10; their constants carry the power-of-two scale, same roundings as unscaled), then the
// quadrant rotation.  Absolute error <= ~6e-8 of the unscaled value as evaluated in float32.  (Degree-7 / degree-6
// minimax fits -- three FFMA fewer, 9e-8 -- were measured: no faster, and two more cases of the reference's own
// op-sequence test fell out of its atol = 1e-4.)  The sign of the sine is returned flipped (`nsn` = -sin): that is
// the value the B operand holds next to the cosine.
__device__ __forceinline__ void sincos_quadrant(float t, int q, float* nsn, float* cs) {
    constexpr float K = kCoefScale;
    const float s = t * t;
    float ps = fmaf(s, K * 0.0821458866f, K * -0.599264529f);   // pi^9/9!, -pi^7/7!
    ps = fmaf(ps, s, K * 2.55016404f);                           // pi^5/5!
    ps = fmaf(ps, s, K * -5.16771278f);                          // -pi^3/3!
    ps = fmaf(ps * s, t, t * (K * 3.14159274f));                 // t*pi + t*s*(...)
    float pc = fmaf(s, K * -0.0258068914f, K * 0.235330630f);    // -pi^10/10!, pi^8/8!
    pc = fmaf(pc, s, K * -1.33526277f);                          // -pi^6/6!
    pc = fmaf(pc, s, K * 4.05871213f);                           // pi^4/4!
    pc = fmaf(pc, s, K * -4.93480220f);                          // -pi^2/2!
    pc = fmaf(pc, s, K);
    const bool swap = q & 1;
    const float a = swap ? pc : ps, b = swap ? ps : pc;
    // q mod 4: 0 -> (s, c); 1 -> (c, -s); 2 -> (-s, -c); 3 -> (-c, s)
    *nsn = __int_as_float(__float_as_int(a) ^ (~(q << 30) & 0x80000000));
    *cs = __int_as_float(__float_as_int(b) ^ (((q + 1) << 30) & 0x80000000));
}

// (delay_s, phase_rad) of a delay_vals entry whose four fields were loaded with ONE 128-bit instruction.  The two rate
// fields are folded in as `x | (rate & 0)`: a real instruction that keeps them -- and with them the vector load -- alive;
// with two fields dead ptxas narrows the load into two 32-bit loads, each of which walks the warp's 512-byte span again.
__device__ __forceinline__ float2 delay_and_phase(const float4& e) {
    uint32_t x = __float_as_uint(e.x), z = __float_as_uint(e.z);
    asm volatile("lop3.b32 %0, %0, %1, %2, 0xF8;" : "+r"(x) : "r"(__float_as_uint(e.y)), "r"(0u));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0xF8;" : "+r"(z) : "r"(__float_as_uint(e.w)), "r"(0u));
    return make_float2(__uint_as_float(x), __uint_as_float(z));
}

// One (beam, antenna) coefficient -> the four 32-bit words of the B operand: rows n = 2m (k = 2a: cos, 2a+1: -sin) and
// n = 2m+1 (sin, cos), each as fp16 hi and fp16 residual (lo).
__device__ __forceinline__ void coef_words(float nsn, float cs, uint32_t* hi0, uint32_t* hi1, uint32_t* lo0, uint32_t* lo1) {
    const uint32_t hi = pack_half2(cs, nsn);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float rc = cs - hf.x, rs = nsn - hf.y;
    *hi0 = hi;
    *hi1 = pack_half2(-nsn, cs);  // (same roundings as hi: the halves swapped, the sine's sign flipped)
    *lo0 = pack_half2(rc, rs);
    *lo1 = pack_half2(-rs, rc);
}

// Time offset of a coefficient set as a float pair: the heap's, or (sample_dt != 0) that of the centre of time tile h of
// the heap (the native precursor steps its coefficients per timestamp, dt = t * SAMPLING_PERIOD * FFT_SIZE,
// beamformer_coefficient_generator/BeamformerKernels.cu:153-156; here per 128-sample MMA tile).
__device__ __forceinline__ void set_time(const FusedParams& prm, int b, int h, float* dt_hi, float* dt_lo) {
    double dt = prm.dt_s[b];
    if (prm.sample_dt != 0.0) {
        const int n = min(kTileT, prm.T - h * kTileT);
        dt += (static_cast<double>(h * kTileT) + 0.5 * static_cast<double>(n - 1)) * prm.sample_dt;
    }
    *dt_hi = static_cast<float>(dt);
    *dt_lo = static_cast<float>(dt - static_cast<double>(*dt_hi));
}

// base + rate * dt as a float pair (dt = dt_hi + dt_lo): the model value at a heap's timestamp
// (beamformer_coefficient_generator/BeamformerKernels.cu:28-35: fDeltaDelay, fDeltaPhase).
__device__ __forceinline__ void advance_model(float base, float rate, float dt_hi, float dt_lo, float* hi, float* lo) {
    const float w = rate * dt_hi;
    float w_e = fmaf(rate, dt_hi, -w);
    w_e = fmaf(rate, dt_lo, w_e);
    const float sum = base + w;
    const float bb = sum - base;
    *hi = sum;
    *lo = ((base - (sum - bb)) + (w - bb)) + w_e;
}

// Four float32 beam values -> four int8 in one word.  The value is clipped BEFORE scaling (to +-127/gain, so that
// one FFMA then scales and rounds: adding 1.5*2^23 leaves the two's-complement int8 in the low mantissa byte,
// round-half-even), three byte permutes pack the word.  kCount also counts clipped values.
template <bool kCount>
__device__ __forceinline__ uint32_t quantise4(const uint32_t (&r)[32], int j, float gain, float limit, int* clipped) {
    uint32_t m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float v = __uint_as_float(r[4 * j + i]);
        const float cl = fminf(fmaxf(v, -limit), limit);
        if (kCount) *clipped += (cl != v);
        m[i] = __float_as_uint(fmaf(cl, gain, 12582912.0f));
    }
    return __byte_perm(__byte_perm(m[0], m[1], 0x0040u), __byte_perm(m[2], m[3], 0x0040u), 0x5410u);
}

// The same for values that are known to stay below 2^15 quantisation steps (the caller checks 510 * A * max|gain| -- the
// largest |beam * gain| 8-bit voltages can produce -- against it): scale, round and bias by 128 in one FFMA (an even
// bias keeps round-half-even), then clamp PAIRS of 16-bit results with one add-min-relu each, pack, and undo the bias on
// the packed word: 11 instructions per four values instead of 15.
__device__ __forceinline__ uint32_t quantise4_fast(const uint32_t (&r)[32], int j, float gain) {
    uint32_t m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = __float_as_uint(fmaf(__uint_as_float(r[4 * j + i]), gain, 12582912.0f + 128.0f));
    // low halves = k + 128;  relu(min(k + 127, 254)) = clip(k, -127, 127) + 127
    const uint32_t c01 = __viaddmin_s16x2_relu(__byte_perm(m[0], m[1], 0x5410u), 0xffffffffu, 0x00fe00feu);
    const uint32_t c23 = __viaddmin_s16x2_relu(__byte_perm(m[2], m[3], 0x5410u), 0xffffffffu, 0x00fe00feu);
    return (__byte_perm(c01, c23, 0x6420u) + 0x01010101u) ^ 0x80808080u;  // bytes t in [0, 254] -> t - 127 (two's complement)
}

// ------------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------------
template <bool kProf, bool kTv, bool kQ8, bool kMerged, bool kStream, bool kPair = false>
__global__ void __launch_bounds__(kQ8 ? kThreadsQ8 : kThreads, 1)
fused_beamform_kernel(const __grid_constant__ FusedParams prm, const __grid_constant__ CUtensorMap tm_in,
                      const __grid_constant__ CUtensorMap tm_out) {
    extern __shared__ uint8_t smem_raw[];
    const unsigned long long t_entry = kProf ? global_ns() : 0ull;
    // Programmatic dependent launch: with DCBF_FLAG_STREAMING the NEXT fused launch on the stream may start
    // filling SMs as soon as this launch's CTAs leave them (there is no griddepcontrol.wait anywhere: such
    // launches promise to be independent).  Without the launch attribute this is a no-op.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const uint32_t bop_base = smem_base;                                    // [buf][kb][part][nt x 128 B]
    const uint32_t ost_base = bop_base + kBopBufs * kBopBufBytes;           // [warp][2][32 x 128 B]
    const uint32_t aop_base = ost_base + kOutStageBytes;                    // [stage][pol][128 x 64 B]
    const uint32_t raw_base = aop_base + kAopStages * kAopStageBytes;       // [stage][ant][t][4 B]
    const uint32_t bar_base = raw_base + kRawStages * kRawStageBytes;       // 8-byte mbarriers
    Control* ctl = reinterpret_cast<Control*>(smem_gen + kSmemData + 320);

    // barrier ids (also reported by the watchdog)
    const int kRawFull = 0, kRawEmpty = kRawFull + kMaxRawStages, kAopFull = kRawEmpty + kMaxRawStages,
              kAopEmpty = kAopFull + kMaxAopStages, kBopFull = kAopEmpty + kMaxAopStages, kBopEmpty = kBopFull + kBopSlots,
              kAccFull = kBopEmpty + kBopSlots, kAccEmpty = kAccFull + kAccBufs, kNumBars = kAccEmpty + kAccBufs;
    static_assert(2 * (kMaxRawStages + kMaxAopStages + kBopSlots + kAccBufs) * 8 <= 320, "barrier area");
    // K-streamed B: every slab is a convert -> MMA -> convert round trip (fence, arrive, wake-up, commit), so the depth of
    // the A ring bounds the slab rate more than the look-ahead of the B ring does (ncu: every role of a two-stage
    // pipeline sat in its barrier waits).  Four A stages, the two extra ones in the fourth slot of the B ring.
    // Whole tile sets narrower than their 64 KiB buffers (C2, C4: 16-48 KiB) leave room for the same two extra stages, one
    // in the last 16 KiB of each buffer (prm.aop_extra): the A ring, not the work, set the slab rate there as well.
    constexpr uint32_t kAStages = kStream ? kMaxAopStages : kAopStages;  // (K-streamed builds: a compile-time constant)
    const uint32_t a_shift = kStream ? 2u : (prm.aop_extra ? 2u : 1u), a_mask = (1u << a_shift) - 1u;
    constexpr uint32_t kBSlots = kStream ? 3 : kBopSlots;
    auto aop_addr = [&](uint32_t as) {
        if (as < kAopStages) return aop_base + as * kAopStageBytes;
        return kStream ? bop_base + 3u * kBopSlotBytes + (as - kAopStages) * kAopStageBytes
                       : bop_base + (as - kAopStages) * kBopBufBytes + (kBopBufBytes - kAopStageBytes);
    };
    static_assert(320 + sizeof(Control) <= kCtlBytes, "control area");
    // raw stage s: its own 8 KiB for s < kRawStages, else alternately behind the tiles of B buffer 0 / 1
    const uint32_t raw_stages = static_cast<uint32_t>(prm.raw_stages);
    auto raw_addr = [&](uint32_t s) {
        return s < kRawStages ? raw_base + s * kRawStageBytes
                              : bop_base + ((s - kRawStages) & 1u) * kBopBufBytes + prm.raw_extra_off + ((s - kRawStages) >> 1) * kRawStageBytes;
    };
    auto bar = [&](int id) { return bar_base + 8u * static_cast<uint32_t>(id); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool kQ8Wide = kQ8 && !kMerged && !kStream;  // register split 80 / 56 instead of 72 / 72 (see kThreadsQ8)
    // Float32 output: the coefficient role gets 80 registers, convert 48, the epilogue 96, the issue warps 40 (the launch
    // bound of 72 per thread is what the 28 warps share).  At 72 the time-varying builds -- four fields per entry and the
    // time pair -- spilled in this role, to L2 (see kThreadsQ8): C3 with per-heap times 313 -> 287 us; the CTA-pair build
    // reloaded its per-unit counters from local memory (C5 on one GPU 2463 -> 2370 us); and the static builds can then
    // afford whole 128-bit delay_vals loads (kWhole128 below): C3 at the power cap 304 -> 290 us, C4-shaped 4096 channels
    // 198 -> 194 us, the 512-channel share of C3 40.8 -> 39.6 us (same box, interleaved).
    constexpr bool kTvSplit = !kQ8 && DCBF_COEFF_WARPS == 16;
    static_assert(kCoeffWarps * 80 + 4 * (48 + 96 + 40) <= (kCoeffWarps + 12) * DCBF_REGS_LAUNCH || DCBF_COEFF_WARPS != 16, "register pool, time-varying steering");

    // ---- one-time setup ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxRawStages; ++s) {
            mbar_init(bar(kRawFull + s), 1);
            mbar_init(bar(kRawEmpty + s), 4);
        }
        for (int s = 0; s < kMaxAopStages; ++s) {
            mbar_init(bar(kAopFull + s), kPair ? 8 : 4);  // (a pair's barriers of this kind live in CTA 0 and count both CTAs)
            mbar_init(bar(kAopEmpty + s), 2);  // one tcgen05.commit per MMA warp
        }
        for (int s = 0; s < kBopSlots; ++s) {
            mbar_init(bar(kBopFull + s), prm.packed_mode == 2 ? 1 : kPair ? 2 * kCoeffWarps : kCoeffWarps);  // (packed: one bulk copy)
            mbar_init(bar(kBopEmpty + s), 2);
        }
        for (int s = 0; s < kAccBufs; ++s) {
            mbar_init(bar(kAccFull + s), 2);
            mbar_init(bar(kAccEmpty + s), (kPair ? 8 : 4) * (kQ8 ? 2 : 1));  // one arrival per epilogue warp (and CTA)
        }
        ctl->abort = 0;
        ctl->chan_pub = 0;
        for (int i = 0; i < 24; ++i) ctl->wait_ns[i >> 2][i & 3] = 0;
        fence_mbar_init();
        (void)kNumBars;
    }
    // B tiles start as zeros: padding rows (k >= 2A, n >= 2M) are never written afterwards.
    // (only the part of each 64 KiB buffer the tiles occupy: the tail, if any, holds extra raw stages)
    {
        const int used16 = kStream ? kBopBufBytes / 16 : prm.raw_extra_off / 16;  // 16-byte units per buffer
        for (int buf = 0; buf < kBopBufs; ++buf) {
            uint4* z = reinterpret_cast<uint4*>(smem_gen + buf * kBopBufBytes);
            for (int i = threadIdx.x; i < used16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async_smem();
    }
    if (warp == kMmaWarp) {
        if (kPair) tmem_alloc_pair(smem_u32(&ctl->tmem_base), kTmemCols);
        else tmem_alloc(smem_u32(&ctl->tmem_base), kTmemCols);
    }
    if (warp == kProducerWarp && lane == 0) prefetch_tensormap(&tm_in);
    if (warp == kEpilogueWarp0 && lane == 0) prefetch_tensormap(&tm_out);
    if (!kPair && warp == kCoeffWarp0 + 1 && prm.pdl_wait && !(prm.dbg & 8) && prm.packed_mode != 1) {
        // The first unit's inputs -> L2 before anything else asks for memory (and, with programmatic dependent launch,
        // while the preceding kernel still drains: a prefetch returns no data, and L2 is the coherence point of global
        // memory, so this is safe ahead of the dependency wait).  The launch is otherwise serial until the first B tile
        // set is complete, and that set waits for these delay_vals at DRAM latency under the previous launch's write-back.
        const uint32_t w = blockIdx.x;
        uint32_t uc;
        int m0 = 0, b0 = 0, h0, hn = 1;
        if (kStream) {
            const uint32_t pc = static_cast<uint32_t>(prm.nt_count * prm.hg_count);
            uc = w / pc;
            const int r_ = static_cast<int>(w - uc * pc), it_ = r_ / prm.hg_count;
            m0 = it_ * (prm.nt >> 1), h0 = prm.hg_size * (r_ - it_ * prm.hg_count), hn = min(prm.hg_size, prm.ht_count - h0);
        } else {
            int j0 = 0;
            uc = w;
            if (w >= static_cast<uint32_t>(prm.n_whole)) {
                const uint32_t r_ = w - static_cast<uint32_t>(prm.n_whole), c_ = r_ / static_cast<uint32_t>(prm.split);
                uc = static_cast<uint32_t>(prm.n_whole) + c_;
                j0 = (static_cast<int>(r_ - c_ * static_cast<uint32_t>(prm.split)) / prm.bsplit) * prm.tile_count / (prm.split / prm.bsplit);
            }
            const int bh = prm.B * prm.ht_count;
            m0 = (j0 / bh) * (prm.nt >> 1), b0 = (j0 % bh) / prm.ht_count, h0 = j0 % prm.ht_count;
        }
        size_t dv_bytes = static_cast<size_t>(min(prm.nt >> 1, prm.M - m0)) * prm.A * 16;
        const char* dv_p = reinterpret_cast<const char*>(prm.dv + (static_cast<size_t>(uc) * prm.M + m0) * prm.A);
        if (prm.packed_mode == 2) {  // the channel's packed tile set instead of its delay_vals
            dv_bytes = static_cast<size_t>(prm.packed_bytes);
            dv_p = reinterpret_cast<const char*>(prm.packed) + static_cast<size_t>(uc) * dv_bytes;
        }
        for (size_t o = static_cast<size_t>(lane) * 8192; o < dv_bytes; o += 32 * 8192)
            bulk_prefetch_l2(dv_p + o, static_cast<uint32_t>(min(dv_bytes - o, static_cast<size_t>(8192))));
        const uint32_t row_bytes = static_cast<uint32_t>(min(hn * kTileT, prm.T - h0 * kTileT)) * 4u;
        for (int a = lane; a < prm.A; a += 32)
            bulk_prefetch_l2(prm.samples + ((static_cast<size_t>(b0) * prm.A + a) * prm.C + uc) * prm.T * 4 + static_cast<size_t>(h0) * kTileT * 4,
                             row_bytes);
    }
    // Programmatic dependent launch, default mode: everything above overlapped the tail of the preceding kernel of
    // the stream; from here on global memory is read and written, so wait until that kernel has completed and flushed.
    if (prm.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (kQ8 && warp == 0) {
        // q8: the per-beam gains ride on the coefficients as gain[m] / max|gain| (in [-1, 1], so the fp16 hi+lo split
        // keeps its precision); the epilogue then scales every column by the same max|gain| and clips at 127 / max|gain|
        float g = 0.f;
        for (int m = lane; m < prm.M; m += 32) g = fmaxf(g, fabsf(__ldg(prm.gains + m)));
#pragma unroll
        for (int o = 16; o; o >>= 1) g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, o));
        if (lane == 0) ctl->q8_gmax = g;
    }
    tc_fence_before();
    if (kPair) cluster_sync();  // the peer's barriers are initialised before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;
    // CTA pair (kPair): the two CTAs of a cluster share every unit -- CTA r converts time tile uh0 + r and generates the
    // coefficients of beams [64 r, 64 r + 64) of the N tile; CTA 0 issues the cta_group::2 MMAs (M = 256: 128 rows in each
    // CTA's tensor memory, each CTA supplying its own A rows and its half of the B rows) and owns the barriers that
    // collect both CTAs' arrivals; its commits are multicast to both CTAs.
    const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;

    const int A = prm.A, C = prm.C, T = prm.T, M = prm.M, B = prm.B;
    const int N2 = 2 * M;
    const int nt = prm.nt, parts = prm.parts;
    // Scheduling unit: a channel; with K-streamed B tiles a (channel, N tile, group of <= 2 time tiles) triple
    // (a channel of many beams takes long enough that whole channels balance badly over the CTAs; two time tiles are
    // what the TMEM accumulators of one unit hold)
    const uint32_t per_chan = kStream ? static_cast<uint32_t>(prm.nt_count * prm.hg_count) : 1u;
    const uint32_t n_units = kStream ? static_cast<uint32_t>(C) * per_chan
                                     : static_cast<uint32_t>(prm.n_whole + (C - prm.n_whole) * prm.split);
    // whole-tile-set mode: unit -> channel and the range [j0, j1) of its accumulator tiles; tile j is
    // (N tile j / (B ht), batch (j / ht) % B, time tile j % ht), the order every role walks them in
    auto unit_range = [&](uint32_t w, uint32_t* uc, int* j0, int* j1) {
        if (w < static_cast<uint32_t>(prm.n_whole)) {
            *uc = w, *j0 = 0, *j1 = prm.tile_count;
        } else {
            const uint32_t r_ = w - static_cast<uint32_t>(prm.n_whole), c_ = r_ / static_cast<uint32_t>(prm.split);
            const int s_ = static_cast<int>(r_ - c_ * static_cast<uint32_t>(prm.split));
            *uc = static_cast<uint32_t>(prm.n_whole) + c_;
            const int jp = s_ / prm.bsplit, jsplit = prm.split / prm.bsplit;
            *j0 = jp * prm.tile_count / jsplit, *j1 = (jp + 1) * prm.tile_count / jsplit;
        }
    };
    // beams of the N tile a unit covers: all nt / 2 of them, or (piece of a cut channel, bsplit = 2) one half
    auto unit_beams = [&](uint32_t w, int* um0, int* umt) {
        *um0 = 0, *umt = nt >> 1;
        if (!kStream && prm.bsplit > 1 && w >= static_cast<uint32_t>(prm.n_whole)) {
            const uint32_t r_ = w - static_cast<uint32_t>(prm.n_whole);
            *umt = (nt >> 1) / prm.bsplit;
            *um0 = static_cast<int>((r_ % static_cast<uint32_t>(prm.split)) % static_cast<uint32_t>(prm.bsplit)) * *umt;
        }
    };
    const int bh_count = prm.set_tiles;  // accumulator tiles per coefficient set
    // kStream: unit -> channel, N tile, first time tile and number of time tiles of the group
    auto unit_decode = [&](uint32_t w, uint32_t* uc, int* uit, int* uh0, int* uhn) {
        const uint32_t c_ = w / per_chan, r_ = w - c_ * per_chan;
        const int it_ = static_cast<int>(r_) / prm.hg_count, hg_ = static_cast<int>(r_) - it_ * prm.hg_count;
        *uc = c_, *uit = it_, *uh0 = prm.hg_size * hg_, *uhn = min(prm.hg_size, prm.ht_count - prm.hg_size * hg_);
    };
    const int nt_cta = kPair ? nt >> 1 : nt;  // B rows (output columns) this CTA's shared memory holds: a pair splits the N tile
    const uint32_t part_bytes = static_cast<uint32_t>(nt_cta * 128);        // one part of a k-block: [nt_cta rows][128 B]
    const uint32_t bop_kb_bytes = static_cast<uint32_t>(parts) * part_bytes;  // one k-block: [part][nt rows][128 B]
    // profiling: lane 0 of each role's first warp accounts blocked time per barrier class (slot) and role span
    const bool prof_lane = kProf && lane == 0 && (warp == kProducerWarp || warp == kMmaWarp || warp == kEpilogueWarp0 ||
                                         warp == kConvertWarp0 || warp == kCoeffWarp0);
    const int ps = prof_lane ? 0 : -100;
    const int my_role = warp == kProducerWarp ? kRoleProducer : (warp == kMmaWarp || warp == kMmaWarp2) ? kRoleMma
                        : warp >= kEpilogueWarp0 ? kRoleEpilogue : warp >= kConvertWarp0 ? kRoleConvert : kRoleCoeff;
    const unsigned long long role_t0 = prof_lane ? global_ns() : 0ull;
    const unsigned long long role_t0_cta = (kProf && threadIdx.x < 24) ? global_ns() : 0ull;

    if (warp >= kProducerWarp && warp < kProducerWarp + 4) {
        if constexpr (kQ8 || kTvSplit) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 " DCBF_STR(DCBF_REGS_ISSUE) ";");
    }
    if (warp == kProducerWarp) {
        // =================================== TMA producer ===================================
        uint32_t rs = 0, ph = 0;
        bool ok = prm.packed_mode != 1;  // (packing the coefficients: no voltages are read)
        // one raw slab: [16 antennas][128 samples] of (batch b, channel c) -> next ring stage
        auto load_slab = [&](int h, int c, int s, int b) {
            if (!mbar_wait<kProf, false, DCBF_BACKOFF_NS>(bar(kRawEmpty + rs), ph ^ 1u, ctl, prm.status, kRoleProducer, kRawEmpty + rs, ps + 0)) return false;
            if (elect_one()) {
                mbar_arrive_expect_tx(bar(kRawFull + rs), kRawStageBytes);
                tma_load_4d(raw_addr(rs), &tm_in, bar(kRawFull + rs), h * kTileT, c, s * kSlabAnts, b);
            }
            __syncwarp();
            if (++rs == raw_stages) rs = 0, ph ^= 1u;
            return true;
        };
        if constexpr (kStream) {
            // slab outer, time tile inner (every B k-block is then used for both time tiles before it is released)
            for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
                uint32_t uc;
                int uit, uh0, uhn;
                unit_decode(w, &uc, &uit, &uh0, &uhn);
                for (int b = 0; b < B && ok; ++b)
                    for (int s = 0; s < prm.slab_count && ok; ++s) {
                        if (kPair) {  // this CTA's time tile (past the end of the heap: an all-zero box, nothing is stored)
                            ok = load_slab(uh0 + static_cast<int>(cta_rank), static_cast<int>(uc), s, b);
                        } else {
                            for (int i = 0; i < uhn && ok; ++i) ok = load_slab(uh0 + i, static_cast<int>(uc), s, b);
                        }
                    }
            }
        } else {
            // per accumulator tile (N tile, batch, time tile): its antenna slabs
            for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
                uint32_t uc;
                int j0, j1;
                unit_range(w, &uc, &j0, &j1);
                int bh = j0 % (B * prm.ht_count), b = bh / prm.ht_count, h = bh - b * prm.ht_count;
                for (int j = j0; j < j1 && ok; ++j) {
                    for (int s = 0; s < prm.slab_count && ok; ++s) ok = load_slab(h, static_cast<int>(uc), s, b);
                    if (++h == prm.ht_count) {
                        h = 0;
                        if (++b == B) b = 0;
                    }
                }
            }
        }
    } else if (warp == kMmaWarp || warp == kMmaWarp2) {
        // =================================== MMA issuer ===================================
        // merged: the lo rows follow the hi rows in the B tile, so one N = 2 nt MMA replaces two N = nt MMAs and the
        // A tile is read from shared memory once; its two halves land in adjacent TMEM column ranges
        const int mma_parts = kMerged ? 1 : parts;
        const uint32_t acc_cols = static_cast<uint32_t>(kMerged ? 2 * nt : nt);  // TMEM columns per (buffer, pol)
        const uint32_t idesc = make_idesc_f16(static_cast<int>(acc_cols));
        const uint32_t b_lo0 = desc_lo(bop_base);
        const uint32_t part_lo = part_bytes >> 4, kb_lo = bop_kb_bytes >> 4;
        uint32_t slab = 0, unit = 0, step = 0;
        bool ok = true;
        if constexpr (kPair) {
            // CTA pair: CTA 0 issues for both.  One cta_group::2 MMA covers both time tiles of the unit (M = 256: rows
            // 0..127 = this CTA's tile, 128..255 = the peer's) and the whole N tile (each CTA holds half of its B rows);
            // accumulators: pol p at TMEM columns [p nt, (p + 1) nt) of both CTAs.  Commits go to both CTAs' barriers.
            if (cta_rank == 0) {
                const uint32_t pol = warp == kMmaWarp ? 0u : 1u;
                const uint32_t idesc2 = make_idesc_f16(nt, false, false, 256);
                const uint32_t d_tmem = tmem_base + pol * static_cast<uint32_t>(nt);
                uint32_t kstep = 0;
                for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
                    for (int b = 0; b < B && ok; ++b, ++unit) {
                        ok = mbar_wait<kProf>(bar(kAccEmpty), (unit & 1u) ^ 1u, ctl, prm.status, kRoleMma, kAccEmpty, ps + 1);
                        if (!ok) break;
                        tc_fence_after();
                        for (int s = 0; s < prm.slab_count && ok; ++s, ++slab) {
                            const uint32_t slot = kstep % kBSlots;
                            if ((s & 1) == 0)
                                ok = mbar_wait<kProf>(bar(kBopFull + slot), (kstep / kBSlots) & 1u, ctl, prm.status, kRoleMma, kBopFull + slot, ps + 0);
                            const uint32_t as = slab % kAStages;
                            if (ok) ok = mbar_wait<kProf>(bar(kAopFull + as), (slab / kAStages) & 1u, ctl, prm.status, kRoleMma, kAopFull + as, ps + 2);
                            if (!ok) break;
                            tc_fence_after();
                            const int n_ants = min(kSlabAnts, A - s * kSlabAnts);
                            const int k_steps = (n_ants + 7) >> 3;
                            const uint32_t b_lo = b_lo0 + slot * (kBopSlotBytes >> 4) + static_cast<uint32_t>(s & 1) * 4u;
                            const uint32_t a_lo = desc_lo(aop_addr(as)) + pol * (kAopTileBytes >> 4);
                            if (elect_one()) {
#pragma unroll
                                for (int part = 0; part < 2; ++part) {
                                    if (part < parts) {
#pragma unroll
                                        for (int kk = 0; kk < 2; ++kk) {
                                            if (kk < k_steps)
                                                umma_f16_pair(d_tmem, make_desc(a_lo + 2u * kk, kDescHiSw64),
                                                              make_desc(b_lo + part * part_lo + 2u * kk, kDescHiSw128), idesc2,
                                                              (s | part | kk) != 0);
                                        }
                                    }
                                }
                                umma_commit_pair(bar(kAopEmpty + as));
                                if ((s & 1) == 1 || s == prm.slab_count - 1) umma_commit_pair(bar(kBopEmpty + slot));  // last use of this k-block
                            }
                            __syncwarp();
                            if ((s & 1) == 1 || s == prm.slab_count - 1) ++kstep;
                        }
                        if (ok && elect_one()) umma_commit_pair(bar(kAccFull));
                        __syncwarp();
                    }
                }
            }
        } else if constexpr (kStream) {
            // K-streamed B: one ring slot per 32-antenna k-block; all (time tile, pol) accumulators of a
            // (channel, N tile, batch) unit are open at once (ht x 2 x nt TMEM columns), slabs outer, time tiles inner
            const uint32_t pol = warp == kMmaWarp ? 0u : 1u;
            uint32_t kstep = 0;
            for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
                uint32_t uc;
                int uit, uh0, uhn;
                unit_decode(w, &uc, &uit, &uh0, &uhn);
                for (int b = 0; b < B && ok; ++b, ++unit) {
                    ok = mbar_wait<kProf>(bar(kAccEmpty), (unit & 1u) ^ 1u, ctl, prm.status, kRoleMma, kAccEmpty, ps + 1);
                    if (!ok) break;
                    tc_fence_after();
                    for (int s = 0; s < prm.slab_count && ok; ++s) {
                        const uint32_t slot = kstep % kBSlots;
                        if ((s & 1) == 0)
                            ok = mbar_wait<kProf>(bar(kBopFull + slot), (kstep / kBSlots) & 1u, ctl, prm.status, kRoleMma, kBopFull + slot, ps + 0);
                        const int n_ants = min(kSlabAnts, A - s * kSlabAnts);
                        const int k_steps = (n_ants + 7) >> 3;
                        const uint32_t b_lo = b_lo0 + slot * (kBopSlotBytes >> 4) + static_cast<uint32_t>(s & 1) * 4u;
                        for (int h = 0; h < uhn && ok; ++h, ++slab) {
                            const uint32_t as = slab % kAStages;
                            ok = mbar_wait<kProf>(bar(kAopFull + as), (slab / kAStages) & 1u, ctl, prm.status, kRoleMma, kAopFull + as, ps + 2);
                            if (!ok) break;
                            tc_fence_after();
                            const uint32_t a_lo = desc_lo(aop_addr(as)) + pol * (kAopTileBytes >> 4);
                            const uint32_t d_tmem = tmem_base + (static_cast<uint32_t>(h) * kPols + pol) * static_cast<uint32_t>(nt);
                            if (elect_one()) {
#pragma unroll
                                for (int part = 0; part < 2; ++part) {
                                    if (part < parts) {
#pragma unroll
                                        for (int kk = 0; kk < 2; ++kk) {
                                            if (kk < k_steps)
                                                umma_f16(d_tmem, make_desc(a_lo + 2u * kk, kDescHiSw64),
                                                         make_desc(b_lo + part * part_lo + 2u * kk, kDescHiSw128), idesc,
                                                         (s | part | kk) != 0);
                                        }
                                    }
                                }
                                umma_commit(bar(kAopEmpty + as));
                            }
                            __syncwarp();
                        }
                        if (ok && ((s & 1) == 1 || s == prm.slab_count - 1)) {  // last use of this k-block
                            if (elect_one()) umma_commit(bar(kBopEmpty + slot));
                            __syncwarp();
                            ++kstep;
                        }
                    }
                    if (ok && elect_one()) umma_commit(bar(kAccFull));
                    __syncwarp();
                }
            }
        } else
        for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
            uint32_t uc;
            int j0, j1;
            unit_range(w, &uc, &j0, &j1);
            int um0, umt;
            unit_beams(w, &um0, &umt);
            const uint32_t idesc_u = 2 * umt == nt ? idesc : make_idesc_f16(2 * umt);  // (half of the beams: a narrower MMA)
            for (int j = j0; j < j1 && ok; ++step) {  // one (N tile, coefficient set) and the unit's tiles that use it
                const int jend = min(j1, (j / bh_count + 1) * bh_count);
                const uint32_t bb = step % kBopBufs;
                ok = mbar_wait<kProf>(bar(kBopFull + bb), (step / kBopBufs) & 1u, ctl, prm.status, kRoleMma, kBopFull + bb, ps + 0);
                if (prm.packed_mode == 1) {
                    // packing: the finished tile set goes to HBM as it lies in shared memory (the writers have fenced it for
                    // the async proxy); the buffer is handed back once the copy has read it.  One arrival per MMA warp,
                    // like the two commits of the normal path.
                    if (ok && warp == kMmaWarp && elect_one()) {
                        bulk_store(prm.packed + static_cast<size_t>(uc) * prm.packed_bytes, bop_base + bb * kBopBufBytes,
                                   static_cast<uint32_t>(prm.packed_bytes));
                        bulk_commit_group();
                        bulk_wait_group_read<0>();
                    }
                    __syncwarp();
                    if (ok && lane == 0) mbar_arrive(bar(kBopEmpty + bb));
                    j = jend;
                    continue;
                }
                for (; j < jend && ok; ++j, ++unit) {
                    const uint32_t ab = unit % kAccBufs;
                    ok = mbar_wait<kProf>(bar(kAccEmpty + ab), ((unit / kAccBufs) & 1u) ^ 1u, ctl, prm.status, kRoleMma, kAccEmpty + ab, ps + 1);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t d_tmem0 = tmem_base + ab * kPols * acc_cols;
                    for (int s = 0; s < prm.slab_count; ++s, ++slab) {
                        const uint32_t as = slab & a_mask;
                        ok = mbar_wait<kProf>(bar(kAopFull + as), (slab >> a_shift) & 1u, ctl, prm.status, kRoleMma, kAopFull + as, ps + 2);
                        if (!ok) break;
                        tc_fence_after();
                        const int n_ants = min(kSlabAnts, A - s * kSlabAnts);
                        const int k_steps = (n_ants + 7) >> 3;  // 8 antennas = 16 k per MMA
                        const uint32_t a_lo = desc_lo(aop_addr(as));
                        // B: k-block s/2, 64-byte half s%2 of its 128-byte rows
                        const uint32_t b_lo = b_lo0 + bb * (kBopBufBytes >> 4) + static_cast<uint32_t>(s >> 1) * kb_lo + static_cast<uint32_t>(s & 1) * 4u;
                        if (elect_one()) {
                            {
                                const uint32_t p = warp == kMmaWarp ? 0u : 1u;  // this warp's pol
                                const uint32_t d_tmem = d_tmem0 + p * acc_cols;
#pragma unroll
                                for (int part = 0; part < 2; ++part) {
                                    if (part < mma_parts) {
#pragma unroll
                                        for (int k = 0; k < 2; ++k) {
                                            if (k < k_steps && !((prm.dbg & 64) && (s | part | k) != 0))  // (ablation: one MMA per tile)
                                                umma_f16(d_tmem, make_desc(a_lo + p * (kAopTileBytes >> 4) + 2u * k, kDescHiSw64),
                                                         make_desc(b_lo + part * part_lo + 2u * k, kDescHiSw128), idesc_u,
                                                         (s | part | k) != 0);
                                        }
                                    }
                                }
                            }
                            umma_commit(bar(kAopEmpty + as));  // A stage free once these MMAs retire
                        }
                        __syncwarp();
                    }
                    if (ok && elect_one()) umma_commit(bar(kAccFull + ab));
                    __syncwarp();
                }
                if (ok && elect_one()) umma_commit(bar(kBopEmpty + bb));
                __syncwarp();
            }
        }
    } else if ((warp >= kEpilogueWarp0 && warp < kEpilogueWarp0 + 4) || (kQ8 && warp >= kEpilogue2Warp0)) {
        // =================================== epilogue ===================================
        if constexpr (kQ8Wide) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        else if constexpr (kQ8) asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
        else if constexpr (kTvSplit) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 " DCBF_STR(DCBF_REGS_EPILOGUE) ";");
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        // int8 output: two warpgroups, group g quantises pol g; every warp then has 4 KiB of staging (one wide box or
        // four 1 KiB blocks) instead of 8
        const int egroup = kQ8 && warp >= kEpilogue2Warp0 ? 1 : 0;
        const uint32_t ost = ost_base + (kQ8 ? static_cast<uint32_t>(egroup * 4 + q) * kOutBoxBytes
                                             : static_cast<uint32_t>(q) * (2 * kOutBoxBytes));
        const float q8_gain = kQ8 ? ctl->q8_gmax : 0.f, q8_limit = q8_gain > 0.f ? 127.0f / q8_gain : 0.f;
        // 8-bit voltages bound every beam by 510 A (|re| + |im| <= 510 per antenna, unit coefficients, gains <= max|gain|):
        // when that is below 2^15 quantisation steps the packed 16-bit clamp is exact (no counter, no weights: both need
        // the float comparison)
        const bool q8_fast = kQ8 && !prm.saturated && !prm.weights && 510.0f * static_cast<float>(A) * q8_gain < 32000.0f;
        constexpr bool merged = kMerged;  // a specialisation: the extra 32 registers must not weigh on the wide-tile build
        const uint32_t acc_cols = static_cast<uint32_t>(merged ? 2 * nt : nt);  // TMEM columns per (buffer, pol)
        // 32 accumulator columns of this thread's row; merged tiles keep the hi and lo coefficient parts in two
        // column ranges nt apart, which are summed here (the wait for both loads is then already done)
        auto ld32 = [&](uint32_t taddr, uint32_t (&r)[32]) {
            tmem_ld_32x32b_x32(taddr, r);
            if constexpr (merged) {
                uint32_t l[32];
                tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(nt), l);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(l[i]));
            }
        };
        uint32_t unit = 0, box = 0;
        int clipped = 0;
        bool ok = prm.packed_mode != 1;  // (packing the coefficients: no beams are formed)
        // One accumulator tile (128 rows x nt columns at TMEM column col0) -> beams[b][p][c][t0 ..][n0 ..], float32.
        auto store_tile_f32 = [&](uint32_t col0, int b, int p, uint32_t c, int t0, int n0, int ncols) {
            if (prm.tma_store) {
                const int row0 = t0 + 32 * q;  // this warp's 32 rows of the tile
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + col0;
                const int plane = (b * kPols + p) * C + static_cast<int>(c);
                for (int cb = 0; cb < ncols && row0 < T && n0 + cb < N2; cb += 32, ++box) {
                    uint32_t r[32];
                    unsigned long long tp0 = 0, tp1 = 0, tp2 = 0;
                    if (kProf && prof_lane) tp0 = global_ns();
                    ld32(taddr + cb, r);
                    const uint32_t sb = ost + (box & 1u) * kOutBoxBytes;
                    bulk_wait_group_read<1>();  // (issuing lane) the store that last read this box is done
                    __syncwarp();
                    if (kProf && prof_lane) tp1 = global_ns();
                    tmem_wait_ld();
                    if (kProf && prof_lane) tp2 = global_ns();
                    const uint32_t dst = sb + lane * 128;
                    if (!(prm.dbg & 32)) {  // (ablation: no staging stores, no proxy fence)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            st_shared_v4(dst + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                        fence_proxy_async_smem();
                    }
                    __syncwarp();
                    if (kProf && prof_lane) {  // slot 1: bulk-store read wait, slot 2: TMEM read + fence + stores
                        ctl->wait_ns[kRoleEpilogue][1] += tp1 - tp0;
                        ctl->wait_ns[kRoleEpilogue][2] += (tp2 - tp1) | ((global_ns() - tp2) << 32);
                    }
                    if (elect_one()) {
                        if (!(prm.dbg & 4)) tma_store_3d(&tm_out, sb, n0 + cb, row0, plane);
                        bulk_commit_group();
                    }
                }
            } else {
                float* tile_out = prm.out + (((static_cast<size_t>(b) * kPols + p) * C + c) * static_cast<size_t>(T) + t0) * N2 + n0;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int r_lo = 32 * q + 16 * half + (lane >> 2);
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q + 16 * half) << 16) + col0;
                    float* row_lo = tile_out + static_cast<size_t>(r_lo) * N2 + 2 * (lane & 3);
                    float* row_hi = row_lo + 8 * static_cast<size_t>(N2);
                    const bool v_lo = t0 + r_lo < T, v_hi = t0 + r_lo + 8 < T;
                    int cb = 0;
                    for (; cb + 64 <= ncols; cb += 64) {
                        uint32_t r[32];
                        tmem_ld_16x256b_x8(taddr + cb, r);
                        if constexpr (merged) {
                            uint32_t l[32];
                            tmem_ld_16x256b_x8(taddr + cb + static_cast<uint32_t>(nt), l);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(l[i]));
                        }
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int col = cb + 8 * i;
                            if (n0 + col + 2 * (lane & 3) < N2) {
                                if (v_lo) st_global_v2(row_lo + col, r[4 * i], r[4 * i + 1]);
                                if (v_hi) st_global_v2(row_hi + col, r[4 * i + 2], r[4 * i + 3]);
                            }
                        }
                    }
                    for (; cb < ncols; cb += 16) {
                        uint32_t r[8];
                        tmem_ld_16x256b_x2(taddr + cb, r);
                        if constexpr (merged) {
                            uint32_t l[8];
                            tmem_ld_16x256b_x2(taddr + cb + static_cast<uint32_t>(nt), l);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(l[i]));
                        }
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int col = cb + 8 * i;
                            if (n0 + col + 2 * (lane & 3) < N2) {
                                if (v_lo) st_global_v2(row_lo + col, r[4 * i], r[4 * i + 1]);
                                if (v_hi) st_global_v2(row_hi + col, r[4 * i + 2], r[4 * i + 3]);
                            }
                        }
                    }
                }
            }
        };
        if constexpr (kPair) {
            // this CTA's 128 rows (time tile uh0 + rank) of both pols; the accumulators go back to CTA 0's issuer
            for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
                uint32_t c;
                int it, uh0, uhn;
                unit_decode(w, &c, &it, &uh0, &uhn);
                for (int b = 0; b < B && ok; ++b, ++unit) {
                    ok = mbar_wait<kProf>(bar(kAccFull), unit & 1u, ctl, prm.status, kRoleEpilogue, kAccFull, ps + 0);
                    if (!ok) break;
                    tc_fence_after();
                    for (int p = 0; p < kPols; ++p)
                        store_tile_f32(static_cast<uint32_t>(p) * static_cast<uint32_t>(nt), b, p, c,
                                       (uh0 + static_cast<int>(cta_rank)) * kTileT, it * nt, nt);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(bar(kAccEmpty), 0);
                }
            }
        } else if constexpr (kStream) {
            // all (time tile, pol) accumulators of a (channel, N tile, batch) unit complete together
            for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
                uint32_t c;
                int it, uh0, uhn;
                unit_decode(w, &c, &it, &uh0, &uhn);
                    for (int b = 0; b < B && ok; ++b, ++unit) {
                        ok = mbar_wait<kProf>(bar(kAccFull), unit & 1u, ctl, prm.status, kRoleEpilogue, kAccFull, ps + 0);
                        if (!ok) break;
                        tc_fence_after();
                        for (int h = 0; h < uhn; ++h)
                            for (int p = kQ8 ? egroup : 0; p < (kQ8 ? egroup + 1 : kPols); ++p) {  // (int8 output: this warpgroup's pol)
                                const uint32_t col0 = (static_cast<uint32_t>(h) * kPols + p) * static_cast<uint32_t>(nt);
                                if constexpr (kQ8) {
                                    // requantised output, 32-column blocks through four rotating 1 KiB boxes (as below)
                                    const int row0 = (uh0 + h) * kTileT + 32 * q, n0 = it * nt;
                                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + col0;
                                    const int plane = (b * kPols + p) * C + static_cast<int>(c);
                                    for (int cb = 0; cb < nt && row0 < T && n0 + cb < N2; cb += 32, ++box) {
                                        uint32_t r[32], w[8];
                                        tmem_ld_32x32b_x32(taddr + cb, r);
                                        bulk_wait_group_read<3>();  // (issuing lane) the box used four blocks ago is free
                                        __syncwarp();
                                        tmem_wait_ld();
                                        const int valid = min(nt, N2 - n0) - cb;  // columns past the last beam: never stored, never "clipped"
                                        if (valid < 32) {
#pragma unroll
                                            for (int i = 0; i < 32; ++i)
                                                if (i >= valid) r[i] = 0u;
                                        }
#pragma unroll
                                        for (int j = 0; j < 8; ++j)
                                            w[j] = q8_fast ? quantise4_fast(r, j, q8_gain)
                                                   : prm.saturated ? quantise4<true>(r, j, q8_gain, q8_limit, &clipped)
                                                                   : quantise4<false>(r, j, q8_gain, q8_limit, &clipped);
                                        if (prm.tma_store) {
                                            const uint32_t sb = ost + (box & 3u) * 1024u;
                                            const uint32_t dst = sb + lane * 32;  // [32 rows][32 B], 32B swizzle
                                            const uint32_t x = static_cast<uint32_t>((lane >> 2) & 1) << 4;
                                            st_shared_v4(dst + x, w[0], w[1], w[2], w[3]);
                                            st_shared_v4(dst + (x ^ 16u), w[4], w[5], w[6], w[7]);
                                            fence_proxy_async_smem();
                                            __syncwarp();
                                            if (elect_one()) {
                                                tma_store_3d(&tm_out, sb, n0 + cb, row0, plane);
                                                bulk_commit_group();
                                            }
                                        } else if (row0 + lane < T) {  // ragged beam counts: 2-byte stores (row pitch 2M is even)
                                            uint8_t* rowp = reinterpret_cast<uint8_t*>(prm.out_q8) +
                                                            (static_cast<size_t>(plane) * T + row0 + lane) * static_cast<size_t>(N2) + n0 + cb;
#pragma unroll
                                            for (int j = 0; j < 16; ++j)
                                                if (n0 + cb + 2 * j < N2)
                                                    *reinterpret_cast<uint16_t*>(rowp + 2 * j) =
                                                        static_cast<uint16_t>((w[j >> 1] >> (16 * (j & 1))) & 0xffffu);
                                        }
                                    }
                                } else {
                                    store_tile_f32(col0, b, p, c, (uh0 + h) * kTileT, it * nt, nt);
                                }
                            }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(kAccEmpty));
                    }
            }
        } else
        for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
            uint32_t c;
            int j0, j1;
            unit_range(w, &c, &j0, &j1);
            int um0, umt;
            unit_beams(w, &um0, &umt);
            int jit = j0 / (B * prm.ht_count), jb = (j0 - jit * B * prm.ht_count) / prm.ht_count, jh = j0 % prm.ht_count;
            {
                {
                    for (int j = j0; j < j1 && ok; ++j, ++unit) {
                        const int n0 = jit * nt, b = jb, h = jh;  // this tile; then step the (N tile, batch, time tile) counters
                        if (++jh == prm.ht_count) {
                            jh = 0;
                            if (++jb == B) jb = 0, ++jit;
                        }
                        const uint32_t ab = unit % kAccBufs;
                        ok = mbar_wait<kProf>(bar(kAccFull + ab), (unit / kAccBufs) & 1u, ctl, prm.status, kRoleEpilogue, kAccFull + ab, ps + 0);
                        if (!ok) break;
                        tc_fence_after();
                        const int t0 = h * kTileT;
                        if constexpr (kQ8) {
                            // requantised output: thread = row, 32 columns -> 32 bytes per row; four 1 KiB staging boxes
                            // rotate so the bulk stores' shared-memory reads stay off the critical path.
                            const int row0 = t0 + 32 * q;
                            if (prm.q8_wide) {
                                // whole output rows (nt = 2M bytes: 32, 64 or 128) -> one box of 32 rows per pol:
                                // 4x fewer, 4x longer rows for the TMA store engine than 32-byte pieces
                                const uint32_t cmask = static_cast<uint32_t>(nt >> 4) - 1u;  // swizzle span = row length
                                for (int p = egroup; p <= egroup && row0 < T; ++p, ++box) {  // this warpgroup's pol
                                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + (ab * kPols + p) * acc_cols;
                                    const uint32_t sb = ost;
                                    unsigned long long tp0 = 0, tp1 = 0, tp2 = 0;
                                    if (kProf && prof_lane) tp0 = global_ns();
                                    bulk_wait_group_read<0>();  // (issuing lane) the store that last read this box is done
                                    __syncwarp();
                                    if (kProf && prof_lane) tp1 = global_ns();
                                    // 32 columns of row `lane` -> 16-byte chunks c0/16 and c0/16 + 1, XOR-swizzled by the
                                    // row's 128-byte group
                                    auto put32 = [&](const uint32_t (&r)[32], int c0) {
                                        uint32_t w[8];
#pragma unroll
                                        for (int j = 0; j < 8; ++j)
                                            w[j] = q8_fast ? quantise4_fast(r, j, q8_gain)
                                                   : prm.saturated ? quantise4<true>(r, j, q8_gain, q8_limit, &clipped)
                                                                   : quantise4<false>(r, j, q8_gain, q8_limit, &clipped);
                                        const uint32_t a0 = static_cast<uint32_t>(lane * nt + c0);
                                        const uint32_t a1 = a0 + 16u;
                                        st_shared_v4(sb + (a0 ^ (((a0 >> 7) & cmask) << 4)), w[0], w[1], w[2], w[3]);
                                        st_shared_v4(sb + (a1 ^ (((a1 >> 7) & cmask) << 4)), w[4], w[5], w[6], w[7]);
                                    };
                                    if constexpr (merged) {
                                        for (int cb = 0; cb < nt; cb += 32) {
                                            uint32_t r[32];
                                            ld32(taddr + cb, r);
                                            tmem_wait_ld();
                                            put32(r, cb);
                                        }
                                    } else {
                                        for (int cb = 0; cb < nt; cb += 32) {
                                            uint32_t r[32];
                                            tmem_ld_32x32b_x32(taddr + cb, r);
                                            tmem_wait_ld();
                                            put32(r, cb);
                                        }
                                    }
                                    if (kProf && prof_lane) tp2 = global_ns();
                                    fence_proxy_async_smem();
                                    __syncwarp();
                                    if (elect_one()) {
                                        tma_store_3d(&tm_out, sb, 0, row0, (b * kPols + p) * C + c);
                                        bulk_commit_group();
                                    }
                                    if (kProf && prof_lane) {  // slot 1: bulk-store read wait, slot 2: TMEM + quantise + stores | fence + issue
                                        ctl->wait_ns[kRoleEpilogue][1] += tp1 - tp0;
                                        ctl->wait_ns[kRoleEpilogue][2] += (tp2 - tp1) | ((global_ns() - tp2) << 32);
                                    }
                                }
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(bar(kAccEmpty + ab));
                                continue;
                            }
                            auto emit_block = [&](uint32_t (&r)[32], int cb, int plane, uint32_t sb) {
                                uint32_t w[8];
                                // columns past the tile / the last beam hold other accumulators' values: they are never
                                // stored, and must not count as clipped either
                                const int valid = min(nt, N2 - n0) - cb;
                                if (valid < 32) {
#pragma unroll
                                    for (int i = 0; i < 32; ++i)
                                        if (i >= valid) r[i] = 0u;
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    w[j] = q8_fast ? quantise4_fast(r, j, q8_gain)
                                           : prm.saturated ? quantise4<true>(r, j, q8_gain, q8_limit, &clipped)
                                                           : quantise4<false>(r, j, q8_gain, q8_limit, &clipped);
                                if (prm.tma_store) {
                                    const uint32_t dst = sb + lane * 32;  // [32 rows][32 B], 32B swizzle
                                    const uint32_t x = static_cast<uint32_t>((lane >> 2) & 1) << 4;
                                    st_shared_v4(dst + x, w[0], w[1], w[2], w[3]);
                                    st_shared_v4(dst + (x ^ 16u), w[4], w[5], w[6], w[7]);
                                    fence_proxy_async_smem();
                                    __syncwarp();
                                    if (elect_one()) {
                                        tma_store_3d(&tm_out, sb, n0 + cb, row0, plane);
                                        bulk_commit_group();
                                    }
                                } else if (row0 + lane < T) {  // ragged beam counts: 2-byte stores (row pitch 2M is even)
                                    uint8_t* rowp = reinterpret_cast<uint8_t*>(prm.out_q8) +
                                                    (static_cast<size_t>(plane) * T + row0 + lane) * static_cast<size_t>(N2) + n0 + cb;
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        if (n0 + cb + 2 * j < N2)
                                            *reinterpret_cast<uint16_t*>(rowp + 2 * j) =
                                                static_cast<uint16_t>((w[j >> 1] >> (16 * (j & 1))) & 0xffffu);
                                }
                            };
                            for (int p = egroup; p <= egroup; ++p) {  // this warpgroup's pol
                                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + (ab * kPols + p) * acc_cols;
                                const int plane = (b * kPols + p) * C + c;
                                for (int cb = 0; cb < nt && row0 < T && n0 + cb < N2; cb += 32, ++box) {
                                    uint32_t r[32];
                                    ld32(taddr + cb, r);
                                    bulk_wait_group_read<3>();  // (issuing lane) the box used four blocks ago is free
                                    __syncwarp();
                                    tmem_wait_ld();
                                    emit_block(r, cb, plane, ost + (box & 3u) * 1024u);
                                }
                            }
                        } else {
                            for (int p = 0; p < kPols; ++p) store_tile_f32((ab * kPols + p) * acc_cols, b, p, c, t0, n0 + 2 * um0, 2 * umt);
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(kAccEmpty + ab));
                    }
                }
            }
        }
        bulk_wait_group_all();  // (issuing lane) staging memory and the stores themselves are done before exit
        if (kQ8 && prm.saturated) {
            clipped = __reduce_add_sync(0xffffffffu, clipped);
            if (lane == 0 && clipped) atomicAdd(prm.saturated, static_cast<unsigned long long>(clipped));
        }
    } else if (warp >= kConvertWarp0 && warp < kConvertWarp0 + 4) {
        // =================================== convert ===================================
        if constexpr (kQ8) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        else if constexpr (kTvSplit) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 " DCBF_STR(DCBF_REGS_CONVERT) ";");
        // thread = one sample row t; per 4-antenna chunk: 4 conflict-free LDS.32, 8 PRMT + 8 HSUB2, 2 STS.128
        // (quarter-warps write 8 distinct 16-byte chunks of the 64B-swizzled rows: conflict-free)
        const int t = threadIdx.x - kConvertWarp0 * 32;
        const uint32_t bias = prm.a_bias, magic = prm.a_magic;  // (plain scale: 0x3C003C00 | 0x3C803C80 and 0x3C3C3C3C)
        const uint32_t flip = prm.signed_in ? 0x80808080u : 0u;
        const uint32_t sw = static_cast<uint32_t>((t >> 1) & 3);
        uint32_t slab = 0, rs = 0, rph = 0;
        bool ok = prm.packed_mode != 1;
        for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
            int tiles;  // (batch, time tile) pairs of the unit, each with slab_count slabs
            if (kStream) {
                uint32_t uc;
                int uit, uh0;
                unit_decode(w, &uc, &uit, &uh0, &tiles);
                tiles = kPair ? B : tiles * B;  // (a pair: one time tile per CTA)
            } else {
                uint32_t uc;
                int j0, j1;
                unit_range(w, &uc, &j0, &j1);
                tiles = j1 - j0;
            }
            {
                for (int bh = 0; bh < tiles && ok; ++bh)
                    for (int s = 0; s < prm.slab_count; ++s, ++slab) {
                        const uint32_t as = slab & a_mask;
                        ok = mbar_wait2<kProf>(bar(kRawFull + rs), rph, kRawFull + rs, bar(kAopEmpty + as),
                                               ((slab >> a_shift) & 1u) ^ 1u, kAopEmpty + as, ctl, prm.status, kRoleConvert, ps + 0);
                        if (!ok) break;
                        // antennas beyond A were zero-filled by the TMA box: byte 0 -> value 0 (u8), and
                        // 0 ^ 0x80 - 128 -> 0 (i8), so the K padding of the operand needs no special case
                        unsigned long long tc0 = 0, tc1 = 0;
                        if (kProf && prof_lane) tc0 = global_ns();
                        const uint32_t src = raw_addr(rs) + t * 4;
                        const uint32_t dst0 = aop_addr(as) + t * 64;
                        // all 16 loads first (the volatile shared-memory accesses keep their program order, so
                        // interleaving them with the stores would expose the load latency once per chunk)
                        uint32_t w[kSlabAnts];
#pragma unroll
                        for (int i = 0; i < kSlabAnts; ++i) w[i] = ((prm.dbg & 16) && i >= 8) ? 0u : ld_shared_u32(src + i * (kTileT * 4));
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if ((prm.dbg & 16) && j >= 2) continue;  // (ablation: half of the conversion work)
#pragma unroll
                            for (int i = 0; i < 4; ++i) w[4 * j + i] ^= flip;
                            const uint32_t off = (static_cast<uint32_t>(j) ^ sw) << 4;
                            st_shared_v4(dst0 + off, bytes_to_half2(w[4 * j], 0x4140u, bias, magic), bytes_to_half2(w[4 * j + 1], 0x4140u, bias, magic),
                                         bytes_to_half2(w[4 * j + 2], 0x4140u, bias, magic), bytes_to_half2(w[4 * j + 3], 0x4140u, bias, magic));
                            st_shared_v4(dst0 + kAopTileBytes + off, bytes_to_half2(w[4 * j], 0x4342u, bias, magic),
                                         bytes_to_half2(w[4 * j + 1], 0x4342u, bias, magic), bytes_to_half2(w[4 * j + 2], 0x4342u, bias, magic),
                                         bytes_to_half2(w[4 * j + 3], 0x4342u, bias, magic));
                        }
                        if (kProf && prof_lane) tc1 = global_ns();
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (kPair) mbar_arrive_cluster(bar(kAopFull + as), 0);  // the issuer (CTA 0) waits for both CTAs' tiles
                            else mbar_arrive(bar(kAopFull + as));
                            mbar_arrive(bar(kRawEmpty + rs));
                        }
                        if (++rs == raw_stages) rs = 0, rph ^= 1u;
                        if (kProf && prof_lane)  // slot 2: (LDS + convert + STS) | (fence + arrive) << 32
                            ctl->wait_ns[kRoleConvert][2] += (tc1 - tc0) | ((global_ns() - tc1) << 32);
                    }
            }
        }
    } else if (warp < kCoeffWarp0 + kCoeffWarps) {
        // =================================== steering coefficients ===================================
        if constexpr (kQ8Wide) asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");  // (from 64 at launch)
        else if constexpr (kQ8) asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
        else if constexpr (kTvSplit) asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 " DCBF_STR(DCBF_REGS_COEFF) ";");
        // delay_vals[c][m0 .. m0+mt) is one contiguous run of (beam, antenna) entries: the 256 threads walk it
        // with lane <-> consecutive entry, so every warp load is 512 contiguous bytes.  Each entry becomes four
        // 32-bit words (row 2m | row 2m+1) x (fp16 hi | fp16 lo residual); consecutive antennas are consecutive
        // words of one 128-byte B row, so each of the four STS.32 of a warp touches 32 different banks.
        // The loads of the NEXT batch (possibly the next channel's) are issued before the current batch is
        // evaluated, so HBM latency is covered by arithmetic rather than exposed once per batch.
        // kTv (time-varying steering): one coefficient set per heap, all four delay_vals fields are used.
        using Dv = typename std::conditional<kTv, float4, float2>::type;
#ifndef DCBF_COEFF_BATCH
#define DCBF_COEFF_BATCH 4
#endif
        constexpr int kBatch = kTv ? 4 : DCBF_COEFF_BATCH;
        constexpr int kIlp = kTv ? 2 : 4;  // entries evaluated together (bounded by the 72 registers of this role)
        constexpr int kStride = kCoeffWarps * 32;
        const int ctid = threadIdx.x - kCoeffWarp0 * 32;
        const int mt = nt >> 1;  // beams per N tile
        // CTA pair: this CTA generates (and warms) the coefficients of its half of the N tile's beams
        const int mt_cta = kPair ? mt >> 1 : mt, m_cta0 = kPair ? static_cast<int>(cta_rank) * mt_cta : 0;
        const int dm = kStride / A, da = kStride - dm * A;  // (beam, antenna) advance per kStride entries
        const int ml_first = ctid / A, a_first = ctid - ml_first * A;
        // common shape (A = 64, 32, ...): a thread keeps its antenna and moves 4k beams per step, so its B
        // address only advances by a constant
        const bool fast_addr = da == 0 && (dm & 3) == 0;
        const int sb_count = kTv ? prm.sb_count : 1;

        // channel scheduler (lane 0 of the first coefficient warp): keeps the sequence published one entry
        // beyond the load cursor; the atomic is issued a batch before its result is stored so its latency
        // is never waited for
        const bool is_sched = warp == kCoeffWarp0 && lane == 0;
        int sch_n = 0, sch_raw = 0;  // sch_raw: counter value still in flight (not touched until it is published)
        bool sch_end = false, sch_pending = false;
        auto warm_l2 = [&](int ch, int m0) {  // delay_vals of one (channel, N tile) step -> L2
            // (packed tile sets: the bulk copy itself is issued a whole buffer ahead of its use; warming L2 two units ahead
            // on top of that cost 10 % at C3 -- 283 vs 257 us -- the lines are evicted by the output stream and fetched twice)
            if (prm.packed_mode == 2) return;
            m0 += m_cta0;
            if (m0 >= M) return;
            const size_t bytes = static_cast<size_t>(min(mt_cta, M - m0)) * A * 16;
            const char* p = reinterpret_cast<const char*>(prm.dv + (static_cast<size_t>(ch) * M + m0) * A);
            if (prm.dbg & 8) return;
            for (size_t o = 0; o < bytes; o += 65536)
                bulk_prefetch_l2(p + o, static_cast<uint32_t>(min(bytes - o, static_cast<size_t>(65536))));
        };
        auto sch_publish = [&](int id) {
            if (static_cast<uint32_t>(id) >= n_units) {
                id = kChanSentinel;
                sch_end = true;
            } else if (sch_n > 0) {  // about one unit before the register loads get there
                if (kStream) {
                    warm_l2(id / static_cast<int>(per_chan), ((id % static_cast<int>(per_chan)) / prm.hg_count) * (nt >> 1));
                } else {
                    uint32_t uc;
                    int j0, j1;
                    unit_range(static_cast<uint32_t>(id), &uc, &j0, &j1);
                    warm_l2(static_cast<int>(uc), (j0 / bh_count / sb_count) * mt);
                }
            }
            ctl->chan_ring[sch_n & 7] = id;
            __threadfence_block();
            ctl->chan_pub = ++sch_n;
        };
        // (CTA pairs: both CTAs must walk the same sequence, so it is the static one: pair p takes units p, p + pairs, ...)
        auto sch_request = [&]() {
            if (!sch_end && !sch_pending) {
                sch_raw = kPair ? sch_n * static_cast<int>(gridDim.x >> 1) + static_cast<int>(blockIdx.x >> 1) - static_cast<int>(gridDim.x)
                                : atomicAdd(prm.sched, 1);
                sch_pending = true;
            }
        };
        auto sch_flush = [&]() {
            if (sch_pending) {
                sch_publish(static_cast<int>(gridDim.x) + sch_raw);
                sch_pending = false;
            }
        };
        if (is_sched) {
            sch_publish(static_cast<int>(kPair ? blockIdx.x >> 1 : blockIdx.x));
            sch_request();
        }
        __syncwarp();

        if constexpr (kStream) {
            // K-streamed B tiles: one step = one 32-antenna k-block of one (channel, N tile, batch) unit, written to
            // ring slot step % 4.  lane <-> antenna of the k-block (a 512-byte run of delay_vals per beam), warp w
            // takes beams w, w + 16, ...: 4 beams per warp and step.  Coefficients are regenerated per batch in
            // this mode (the ring does not keep them).
            constexpr int kPer = 64 / kCoeffWarps;
            static_assert(kPer == 4, "the K-streamed coefficient step is written for four entries per thread");
            const float ks_inv_gmax = kQ8 && ctl->q8_gmax > 0.f ? 1.0f / ctl->q8_gmax : 0.f;
            const int wl = warp - kCoeffWarp0;
            uint32_t nk = 0;
            int nw = sched_get(ctl, 0), nb = 0, nkb = 0;  // load cursor: unit, batch, k-block
            // delay_vals in flight: the loads of a step are issued kDepth steps ahead, into one of kDepth register sets
            // picked by the step's parity (one copy of the step body per set, so that the set is a compile-time choice).
            constexpr int kDepth = 1;  // (2 was tried: 301 -> 320 us at the C5 share even with 80 registers for this role)
            // All four fields of an entry are loaded with one 128-bit instruction and kept until the step consumes them
            // (delay_and_phase).  Invalid entries load a valid dummy address instead of being predicated (a predicated vector
            // load is split into scalar loads as well); only the stores are masked.
            float4 nxt[kDepth][kPer];
            // the cursor's unit, decoded once per unit (the divisions would otherwise sit in every k-block step): this
            // thread's entry (beam wl, antenna lane) of the unit's first k-block, and which of its four beams exist
            const float4* n_ptr = prm.dv;
            uint32_t n_mask = 0;  // bit u: beam wl + 16 u is inside the N tile; 0 past the last unit
            const size_t u_stride = static_cast<size_t>(kCoeffWarps) * A;  // entries between this thread's beams
            auto cursor_decode = [&]() {
                n_mask = 0;
                if (static_cast<uint32_t>(nw) < n_units) {
                    const int nc = nw / static_cast<int>(per_chan), nit = (nw - nc * static_cast<int>(per_chan)) / prm.hg_count;
                    const int n_mte = max(0, min(mt_cta, M - (nit * mt + m_cta0)));
#pragma unroll
                    for (int u = 0; u < kPer; ++u) n_mask |= (wl + kCoeffWarps * u < n_mte ? 1u : 0u) << u;
                    n_ptr = prm.dv + (static_cast<size_t>(nc) * M + nit * mt + m_cta0 + wl) * A + lane;
                }
                if (prm.dbg & 1) n_mask = 0;
            };
            cursor_decode();
            auto issue_loads = [&](auto set_c) {
                constexpr int set = decltype(set_c)::value;
                const float4* p = n_ptr + kKbAnts * nkb;
                const uint32_t mask = kKbAnts * nkb + lane < A ? n_mask : 0u;
#pragma unroll
                for (int u = 0; u < kPer; ++u) nxt[set][u] = ldg_nc_f4((mask & (1u << u)) ? p + u * u_stride : prm.dv);
            };
            auto advance_cursor = [&]() {
                if (is_sched) {
                    sch_flush();
                    if (sch_n <= static_cast<int>(nk) + 1) sch_request();
                }
                if (static_cast<uint32_t>(nw) >= n_units) return;
                if (++nkb < prm.kb_count) return;
                nkb = 0;
                if (++nb < B) return;
                nb = 0;
                ++nk;
                if (is_sched && sch_n <= static_cast<int>(nk)) {
                    sch_request();
                    sch_flush();
                }
                __syncwarp();
                nw = sched_get(ctl, nk);
                cursor_decode();
            };
            issue_loads(std::integral_constant<int, 0>{});
            if constexpr (kDepth > 1) {
                advance_cursor();
                issue_loads(std::integral_constant<int, kDepth - 1>{});
            }
            uint32_t kstep = 0;
            bool ok = true;
            for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
                uint32_t c;
                int it, uh0, uhn;
                unit_decode(w, &c, &it, &uh0, &uhn);
                // half-turns of phase per second of delay at this channel, as a float pair.  Pinned in registers: left
                // to itself the compiler re-derives the pair from the channel index in every k-block step, and the
                // float64 pipe those five instructions run on is narrow (ncu: 27 % of this role's stall samples).
                float s_hi, s_lo;
                {
                    const double scale = (static_cast<double>(c) + prm.chan_centre) * prm.turns_per_delay;
                    s_hi = static_cast<float>(scale), s_lo = static_cast<float>(scale - static_cast<double>(s_hi));
                    asm volatile("" : "+f"(s_hi), "+f"(s_lo));
                }
                const int m0 = it * mt + m_cta0, mte = max(0, min(mt_cta, M - m0));
                const float* w_tile = prm.weights ? prm.weights + static_cast<size_t>(m0) * A : nullptr;
                const bool scaled = kQ8 || w_tile != nullptr;
                uint32_t u_mask = 0;  // bit u: beam wl + 16 u exists
#pragma unroll
                for (int u = 0; u < kPer; ++u) u_mask |= (wl + kCoeffWarps * u < mte ? 1u : 0u) << u;
                for (int b = 0; b < B && ok; ++b) {
                    float dt_hi = 0.f, dt_lo = 0.f;
                    if constexpr (kTv) set_time(prm, b, uh0, &dt_hi, &dt_lo);  // this batch's (and time tile's) coefficients
                    for (int kb = 0; kb < prm.kb_count && ok; ++kb, ++kstep) {
                      auto do_step = [&](auto set_c) {
                        constexpr int set = decltype(set_c)::value;
                        const uint32_t slot = kstep % kBSlots;
                        Dv v[kPer];
#pragma unroll
                        for (int u = 0; u < kPer; ++u) {
                            if constexpr (kTv) v[u] = nxt[set][u];
                            else v[u] = delay_and_phase(nxt[set][u]);
                        }
                        advance_cursor();
                        issue_loads(set_c);
                        const int a = kKbAnts * kb + lane;
                        const uint32_t st_mask = a < A ? u_mask : 0u;  // entries of this thread that are stored
                        ok = mbar_wait<kProf, false, DCBF_BACKOFF_NS>(bar(kBopEmpty + slot), ((kstep / kBSlots) & 1u) ^ 1u, ctl, prm.status, kRoleCoeff, kBopEmpty + slot, ps + 0);
                        if (!ok) return;
                        // beam m = wl + 16 u: row 2 m has the same swizzle phase for every u, so the four words of
                        // entry u sit at a constant 4096-byte stride from those of entry 0 (immediate offsets)
                        const uint32_t row0 = 2u * static_cast<uint32_t>(wl);
                        const uint32_t d0 = bop_base + slot * kBopSlotBytes + row0 * 128u +
                                            (((static_cast<uint32_t>(lane) >> 2) ^ (row0 & 7u)) << 4) +
                                            ((static_cast<uint32_t>(lane) & 3u) << 2);
                        const uint32_t d1 = (d0 + 128u) ^ 16u;  // row + 1: swizzle phase (row & 7) | 1
                        const uint32_t d0l = d0 + part_bytes, d1l = d1 + part_bytes;
                        // kGroup entries at a time as one straight-line region, so that their dependent chains interleave:
                        // reduced phases (float pairs; a rare out-of-range group is redone in float64), sin / cos, fp16
                        // split.  Two copies of the region: with and without the optional real factor per entry (?beam-weights
                        // and / or the requantisation gain of the beam relative to the largest one; tiny tables that stay in
                        // L1 / L2) -- as one region the plain path's arithmetic would share a scoreboard with those loads and
                        // through it wait for the delay_vals just requested for the NEXT step (ncu: 10 % of this role's samples)
                        constexpr int kGroup = kTv ? 2 : 4;
                        auto generate = [&](auto scaled_c) {
                            constexpr bool kScaled = decltype(scaled_c)::value;
                            float f[kPer];
                            if constexpr (kScaled) {
#pragma unroll
                                for (int u = 0; u < kPer; ++u) {
                                    const int m = wl + kCoeffWarps * u;
                                    f[u] = 1.0f;
                                    if (st_mask & (1u << u)) {
                                        if (w_tile) f[u] = __ldg(w_tile + static_cast<size_t>(m) * A + a) * prm.w_scale;
                                        if constexpr (kQ8) f[u] *= __ldg(prm.gains + m0 + m) * ks_inv_gmax;
                                    }
                                }
                            }
#pragma unroll
                            for (int g = 0; g < kPer; g += kGroup) {
                                float t[kGroup];
                                int q[kGroup];
                                bool in_range = true;
                                if constexpr (kTv) {
                                    float d_hi[kGroup], d_lo[kGroup], ph_hi[kGroup], ph_lo[kGroup];
#pragma unroll
                                    for (int u = 0; u < kGroup; ++u) {
                                        advance_model(v[g + u].x, v[g + u].y, dt_hi, dt_lo, &d_hi[u], &d_lo[u]);
                                        advance_model(v[g + u].z, v[g + u].w, dt_hi, dt_lo, &ph_hi[u], &ph_lo[u]);
                                        in_range &= steer_tq_fast<true>(d_hi[u], d_lo[u], ph_hi[u], ph_lo[u], s_hi, s_lo, &t[u], &q[u]);
                                    }
                                    if (!in_range) {
                                        const double scale = (static_cast<double>(c) + prm.chan_centre) * prm.turns_per_delay;
#pragma unroll
                                        for (int u = 0; u < kGroup; ++u) {
                                            const Tq r = steer_tq_f64<true>(d_hi[u], d_lo[u], ph_hi[u], ph_lo[u], scale);
                                            t[u] = r.t, q[u] = r.q;
                                        }
                                    }
                                } else {
#pragma unroll
                                    for (int u = 0; u < kGroup; ++u)
                                        in_range &= steer_tq_fast<false>(v[g + u].x, 0.f, v[g + u].y, 0.f, s_hi, s_lo, &t[u], &q[u]);
                                    if (!in_range) {
                                        const double scale = (static_cast<double>(c) + prm.chan_centre) * prm.turns_per_delay;
#pragma unroll
                                        for (int u = 0; u < kGroup; ++u) {
                                            const Tq r = steer_tq_f64<false>(v[g + u].x, 0.f, v[g + u].y, 0.f, scale);
                                            t[u] = r.t, q[u] = r.q;
                                        }
                                    }
                                }
                                auto emit = [&](auto gu_c, int u) {
                                    constexpr int gu = decltype(gu_c)::value;
                                    float nsn, cs;
                                    sincos_quadrant(t[u], q[u], &nsn, &cs);
                                    if constexpr (kScaled) {
                                        nsn *= f[gu];
                                        cs *= f[gu];
                                    }
                                    uint32_t hi0, hi1, lo0, lo1;
                                    coef_words(nsn, cs, &hi0, &hi1, &lo0, &lo1);
                                    if (st_mask & (1u << gu)) {
                                        constexpr int kOff = gu * (2 * kCoeffWarps * 128);
                                        st_shared_u32_off<kOff>(d0, hi0);
                                        st_shared_u32_off<kOff>(d1, hi1);
                                        if (parts > 1) {
                                            st_shared_u32_off<kOff>(d0l, lo0);
                                            st_shared_u32_off<kOff>(d1l, lo1);
                                        }
                                    }
                                };
                                if constexpr (kGroup == 4) {
                                    emit(std::integral_constant<int, 0>{}, 0);
                                    emit(std::integral_constant<int, 1>{}, 1);
                                    emit(std::integral_constant<int, 2>{}, 2);
                                    emit(std::integral_constant<int, 3>{}, 3);
                                } else {
                                    if (g == 0) {
                                        emit(std::integral_constant<int, 0>{}, 0);
                                        emit(std::integral_constant<int, 1>{}, 1);
                                    } else {
                                        emit(std::integral_constant<int, 2>{}, 0);
                                        emit(std::integral_constant<int, 3>{}, 1);
                                    }
                                }
                            }
                        };
                        if (!(prm.dbg & 2)) {
                            if (scaled) generate(std::true_type{});
                            else generate(std::false_type{});
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (kPair) mbar_arrive_cluster(bar(kBopFull + slot), 0);  // the issuer (CTA 0) waits for both halves
                            else mbar_arrive(bar(kBopFull + slot));
                        }
                      };
                      if (kDepth > 1 && (kstep & 1u)) do_step(std::integral_constant<int, kDepth - 1>{});
                      else do_step(std::integral_constant<int, 0>{});
                    }
                }
            }
        } else if (prm.packed_mode == 2) {
            // packed tile sets: one bulk copy per unit into the free B buffer (asynchronous proxy to asynchronous proxy: no
            // fence), issued by the scheduler's warp a whole buffer ahead of the MMAs; the other coefficient warps have
            // nothing to do in this mode
            if (warp == kCoeffWarp0) {
                uint32_t step = 0;
                bool ok = true;
                for (uint32_t k = 0;; ++k) {
                    if (is_sched) {  // keep the sequence published one unit beyond this one
                        while (!sch_end && sch_n <= static_cast<int>(k) + 1) {
                            sch_request();
                            sch_flush();
                        }
                    }
                    __syncwarp();
                    const uint32_t w = static_cast<uint32_t>(sched_get(ctl, k));
                    if (!ok || w >= n_units) break;
                    uint32_t c;
                    int j0, j1;
                    unit_range(w, &c, &j0, &j1);
                    for (int isb = j0 / bh_count, isb_last = (j1 - 1) / bh_count; isb <= isb_last && ok; ++isb, ++step) {
                        const uint32_t bb = step % kBopBufs;
                        ok = mbar_wait<kProf, false, DCBF_BACKOFF_NS>(bar(kBopEmpty + bb), ((step / kBopBufs) & 1u) ^ 1u, ctl, prm.status, kRoleCoeff, kBopEmpty + bb, ps + 0);
                        if (ok && lane == 0) {
                            mbar_arrive_expect_tx(bar(kBopFull + bb), static_cast<uint32_t>(prm.packed_bytes));
                            bulk_load(bop_base + bb * kBopBufBytes, prm.packed + static_cast<size_t>(c) * prm.packed_bytes,
                                      static_cast<uint32_t>(prm.packed_bytes), bar(kBopFull + bb));
                        }
                        __syncwarp();
                    }
                }
            }
        } else {
        // cursor of the batch whose loads are in flight: (unit sequence index, coefficient set of the unit, entry)
        uint32_t nk = 0;
        int nw = sched_get(ctl, 0), n_ch = 0, nisb = 0, nisb_last = 0, ne0 = ctid;
        int n_entries = 0, n_um0 = 0, n_umt = mt;
        const float4* n_src = prm.dv;
        // (static: (delay_s, phase_rad), the two rate fields are ignored like the reference does.  The compiler would narrow
        // the 128-bit load to two 32-bit loads then, each walking the warp's 512-byte span again; it is kept whole as in
        // the K-streamed step (delay_and_phase).  With 72 registers for this role the eight more registers in flight cost
        // more than the LSU passes saved -- C3 262 -> 264.5 us; with 80 it is a gain everywhere, see kTvSplit)
        constexpr bool kWhole128 = !kTv;
        using Nx = typename std::conditional<kTv || kWhole128, float4, float2>::type;
        Nx nxt[kBatch];
        auto cursor_unit = [&]() {  // the cursor's unit -> channel and its range of (N tile, coefficient set) steps
            if (static_cast<uint32_t>(nw) < n_units) {
                uint32_t uc;
                int j0, j1;
                unit_range(static_cast<uint32_t>(nw), &uc, &j0, &j1);
                n_ch = static_cast<int>(uc), nisb = j0 / bh_count, nisb_last = (j1 - 1) / bh_count;
                unit_beams(static_cast<uint32_t>(nw), &n_um0, &n_umt);
            }
        };
        auto cursor_set = [&]() {  // delay_vals of the cursor's (channel, N tile)
            const int m0 = (nisb / sb_count) * mt + n_um0;
            n_entries = min(n_umt, M - m0) * A;
            n_src = prm.dv + (static_cast<size_t>(n_ch) * M + m0) * A;
        };
        cursor_unit();
        if (static_cast<uint32_t>(nw) < n_units) cursor_set();
        auto issue_loads = [&]() {
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int e = ne0 + u * kStride;
                if constexpr (kWhole128) {
                    nxt[u] = ldg_nc_f4((e < n_entries && !(prm.dbg & 1)) ? n_src + e : prm.dv);  // (a dummy address, not a predicate)
                } else {
                    float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e < n_entries && !(prm.dbg & 1)) t4 = ldg_nc_f4(n_src + e);
                    if constexpr (kTv) nxt[u] = t4;
                    else nxt[u] = make_float2(t4.x, t4.z);
                }
            }
        };
        auto advance_cursor = [&]() {
            if (is_sched) {  // publish what has come back, ask for the entry after the cursor's next unit
                sch_flush();
                if (sch_n <= static_cast<int>(nk) + 1) sch_request();
            }
            if (static_cast<uint32_t>(nw) >= n_units) return;
            ne0 += kStride * kBatch;
            if (ne0 - ctid < n_entries) return;
            ne0 = ctid;
            if (nisb < nisb_last) {
                // the unit's next step: the same delay_vals again for the next heap's coefficient set (L2 hits), or
                // the next N tile (warmed when its predecessor was started; warm the one after it)
                ++nisb;
                cursor_set();
                if (is_sched && nisb % sb_count == 0 && nisb + sb_count <= nisb_last) warm_l2(n_ch, (nisb / sb_count + 1) * mt);
                return;
            }
            ++nk;
            if (is_sched && sch_n <= static_cast<int>(nk)) {  // not published yet: do it now (rare)
                sch_request();
                sch_flush();
            }
            __syncwarp();
            nw = sched_get(ctl, nk);
            n_entries = 0;
            cursor_unit();
            if (static_cast<uint32_t>(nw) < n_units) {
                cursor_set();
                if (is_sched && nisb + sb_count <= nisb_last) warm_l2(n_ch, (nisb / sb_count + 1) * mt);
            }
        };
        issue_loads();

        uint32_t step = 0;
        bool ok = true;
        for (uint32_t k = 0, w; ok && (w = sched_get(ctl, k)) < n_units; ++k) {
            uint32_t c;
            int j0, j1;
            unit_range(w, &c, &j0, &j1);
            int um0, umt;
            unit_beams(w, &um0, &umt);
            const double scale = (static_cast<double>(c) + prm.chan_centre) * prm.turns_per_delay;  // half-turns per second of delay
            const float s_hi = static_cast<float>(scale), s_lo = static_cast<float>(scale - static_cast<double>(s_hi));
            float dt_hi = 0.f, dt_lo = 0.f;
            const float* w_tile = nullptr;
            const float* g_tile = nullptr;
            const float q8_inv_gmax = kQ8 && ctl->q8_gmax > 0.f ? 1.0f / ctl->q8_gmax : 0.f;
            const bool two_parts = parts > 1;
            // quadrant + remainder of one entry's steering phase, float pairs only; false: out of range, redo in float64
            auto phase_fast = [&](const Dv& dv, float* t, int* q) {
                if constexpr (kTv) {
                    float d_hi, d_lo, ph_hi, ph_lo;
                    advance_model(dv.x, dv.y, dt_hi, dt_lo, &d_hi, &d_lo);
                    advance_model(dv.z, dv.w, dt_hi, dt_lo, &ph_hi, &ph_lo);
                    return steer_tq_fast<true>(d_hi, d_lo, ph_hi, ph_lo, s_hi, s_lo, t, q);
                } else {
                    return steer_tq_fast<false>(dv.x, 0.f, dv.y, 0.f, s_hi, s_lo, t, q);
                }
            };
            auto phase_f64 = [&](const Dv& dv) {
                if constexpr (kTv) {
                    float d_hi, d_lo, ph_hi, ph_lo;
                    advance_model(dv.x, dv.y, dt_hi, dt_lo, &d_hi, &d_lo);
                    advance_model(dv.z, dv.w, dt_hi, dt_lo, &ph_hi, &ph_lo);
                    return steer_tq_f64<true>(d_hi, d_lo, ph_hi, ph_lo, scale);
                } else {
                    return steer_tq_f64<false>(dv.x, 0.f, dv.y, 0.f, scale);
                }
            };
            // (quadrant, remainder) of one (beam, antenna) entry -> four 32-bit words of the B tile at d0 (row 2m) and d1
            // (row 2m+1); `valid` only guards the stores and the optional table reads (the arithmetic is branch-free)
            auto finish = [&](float t, int q, uint32_t d0, int e, bool valid) {
                float nsn, cs;
                sincos_quadrant(t, q, &nsn, &cs);
                if (w_tile && valid) {  // ?beam-weights: real weight of this (beam, antenna); the table is tiny and stays in L1/L2
                    const float w = __ldg(w_tile + e) * prm.w_scale;
                    cs *= w;
                    nsn *= w;
                }
                if constexpr (kQ8) {  // requantisation gain of this entry's beam, relative to the largest one
                    // (loading the batch's gains ahead of the wait and the phase arithmetic was measured: four more live
                    // registers, more spills, 220 -> 238 us)
                    if (valid) {
                        const float g = __ldg(g_tile + (A == 1 ? static_cast<uint32_t>(e) : __umulhi(static_cast<uint32_t>(e), prm.inv_a))) * q8_inv_gmax;
                        cs *= g;
                        nsn *= g;
                    }
                }
                uint32_t hi0, hi1, lo0, lo1;
                coef_words(nsn, cs, &hi0, &hi1, &lo0, &lo1);
                const uint32_t d1 = (d0 + 128u) ^ 16u;  // row + 1: swizzle phase (row & 7) | 1
                if (valid) {
                    st_shared_u32(d0, hi0);
                    st_shared_u32(d1, hi1);
                    if (two_parts) {
                        st_shared_u32(d0 + part_bytes, lo0);
                        st_shared_u32(d1 + part_bytes, lo1);
                    }
                }
            };
            auto b_addr = [&](uint32_t buf, int ml, int a) {
                const int row = 2 * ml, al = a & (kKbAnts - 1);
                return buf + static_cast<uint32_t>(a >> 5) * bop_kb_bytes + static_cast<uint32_t>(row) * 128u +
                       (static_cast<uint32_t>(((al >> 2) ^ (row & 7)) << 4) | static_cast<uint32_t>((al & 3) << 2));
            };
            for (int isb = j0 / bh_count, isb_last = (j1 - 1) / bh_count; isb <= isb_last && ok; ++isb, ++step) {
                const int it = isb / sb_count, sb = isb - it * sb_count;
                if constexpr (kTv) {  // set sb of the N tile: batch sb, or (per-tile times) batch sb / ht_count, time tile sb % ht_count
                    const int sets_per_batch = sb_count / B;
                    set_time(prm, sb / sets_per_batch, sb % sets_per_batch, &dt_hi, &dt_lo);
                }
                const uint32_t bb = step % kBopBufs;
                const int m0 = it * mt + um0;
                const int entries = min(umt, M - m0) * A;
                w_tile = prm.weights ? prm.weights + static_cast<size_t>(m0) * A : nullptr;  // entry e <-> [m0 + e / A][e % A]
                if (kQ8) g_tile = prm.gains + m0;
                bool waited = false;
                const uint32_t buf = bop_base + bb * kBopBufBytes;
                int ml = ml_first, a = a_first;
                uint32_t d_fast = b_addr(buf, ml_first, a_first);
                const uint32_t d_step = static_cast<uint32_t>(dm) * 256u;  // dm beams = 2 dm rows of 128 B
                for (int e0 = ctid; e0 - ctid < entries; e0 += kStride * kBatch) {  // e0 - ctid is warp-uniform
                    const unsigned long long tl0 = kProf && prof_lane ? global_ns() : 0ull;
                    Dv v[kBatch];
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        if constexpr (kWhole128) v[u] = delay_and_phase(nxt[u]);
                        else v[u] = nxt[u];
                    }
                    if (kProf && prof_lane) {  // developer probe: time spent waiting for this batch's delay_vals to land
                        float sink = 0.f;
#pragma unroll
                        for (int u = 0; u < kBatch; ++u) sink += v[u].x;
                        asm volatile("" ::"f"(sink));
                        ctl->wait_ns[kRoleCoeff][1] += global_ns() - tl0;
                    }
                    advance_cursor();
                    issue_loads();
                    if (!waited) {
                        ok = mbar_wait<kProf, false, DCBF_BACKOFF_NS>(bar(kBopEmpty + bb), ((step / kBopBufs) & 1u) ^ 1u, ctl, prm.status, kRoleCoeff, kBopEmpty + bb, ps + 0);
                        waited = true;
                        if (!ok) break;
                    }
                    // kIlp entries at a time in straight-line passes, so that their dependent chains interleave: reduced
                    // phases (float pairs; the rare out-of-range entry is redone in float64), then sin/cos, fp16 split, stores
                    uint32_t d0[kBatch];
                    if (fast_addr) {
#pragma unroll
                        for (int u = 0; u < kBatch; ++u) d0[u] = d_fast + static_cast<uint32_t>(u) * d_step;
                        d_fast += kBatch * d_step;
                    } else {
#pragma unroll
                        for (int u = 0; u < kBatch; ++u) {
                            d0[u] = b_addr(buf, ml, a);
                            ml += dm;
                            a += da;
                            if (a >= A) {
                                a -= A;
                                ++ml;
                            }
                        }
                    }
                    const bool whole = (e0 - ctid) + kStride * kBatch <= entries;  // batch inside the tile (uniform over the role)
                    // entries of the tile from this batch on (uniform over the role).  A batch that reaches past the end of
                    // a narrow tile (C2: 1024 entries for the 2048 of a batch; C4: 2560, so its second batch holds 512)
                    // evaluates only the pair that exists (the arithmetic is branch-free per entry): C2 40.1 -> 39.2 us, the
                    // C4 share 193.5 -> 192.7 us, C3 unchanged (same box, interleaved)
                    const int left = entries - (e0 - ctid);
                    auto evaluate = [&](auto n_c, int g) {
                        constexpr int n = decltype(n_c)::value;
                        float ph_t[n];
                        int ph_q[n];
                        bool in_range = true;
#pragma unroll
                        for (int u = 0; u < n; ++u) in_range &= phase_fast(v[g + u], &ph_t[u], &ph_q[u]);
                        if (!in_range) {
#pragma unroll
                            for (int u = 0; u < n; ++u) {
                                const Tq r = phase_f64(v[g + u]);
                                ph_t[u] = r.t, ph_q[u] = r.q;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < n; ++u)
                            finish(ph_t[u], ph_q[u], d0[g + u], e0 + (g + u) * kStride, whole || e0 + (g + u) * kStride < entries);
                    };
                    if (prm.dbg & 2) {  // (ablation: no phase / sin-cos arithmetic, no B stores)
                    } else if constexpr (kIlp == 4 && kBatch == 4) {
                        if (left <= 2 * kStride) evaluate(std::integral_constant<int, 2>{}, 0);
                        else evaluate(std::integral_constant<int, 4>{}, 0);
                    } else {
#pragma unroll
                        for (int g = 0; g < kBatch; g += kIlp)
                            if (g == 0 || g * kStride < left) evaluate(std::integral_constant<int, kIlp>{}, g);
                    }
                }
                if (!ok) break;
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(kBopFull + bb));
                if (kProf && prof_lane && step == 0) ctl->wait_ns[kRoleCoeff][2] = global_ns() - role_t0;  // first set done
            }
        }
        }
    }

    // ---- teardown ----
    if (prof_lane) ctl->wait_ns[my_role][3] = global_ns() - role_t0;
    tc_fence_before();
    if (kPair) cluster_sync();  // nothing of the peer (arrivals, multicast commits, MMA reads of this CTA's tiles) is still under way
    else __syncthreads();
    if (threadIdx.x == 0) {  // last CTA out re-arms the channel counter for the next launch that uses this slot
        __threadfence();
        if (atomicAdd(prm.sched + 1, 1) == static_cast<int>(gridDim.x) - 1) {
            atomicExch(prm.sched, 0);
            atomicExch(prm.sched + 1, 0);
        }
    }
    if (kProf && prm.prof && threadIdx.x < 24) {
        unsigned long long v = ctl->wait_ns[threadIdx.x >> 2][threadIdx.x & 3];
        // role index 0 is unused: absolute times of CTA entry, first role start (thread 0's view) and exit
        if (threadIdx.x == 0) v = t_entry;
        if (threadIdx.x == 1) v = role_t0_cta;
        if (threadIdx.x == 2) v = global_ns();
        prm.prof[blockIdx.x * 24 + threadIdx.x] = v;
    }
    if (warp == kMmaWarp) {
        tc_fence_after();
        if (kPair) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

constexpr int kLiveSlots = 64;   // launches in flight at once on one device that can share the pool without interfering
constexpr int kSchedSlots = 2 * kLiveSlots;  // the second half belongs to launches captured into CUDA graphs, which keep
                                             // their slot for every replay and must not meet a live launch on it
using KernelFn = void (*)(const FusedParams, const CUtensorMap, const CUtensorMap);
// index = 6 * q8 + 2 * variant (0 plain, 1 profiling, 2 time-varying) + merged; [12] .. [16] = K-streamed B (plain,
// profiling, time-varying, int8 output, int8 output + time-varying)
constexpr int kNumKernels = 19;  // [17], [18] = K-streamed B on CTA pairs (cta_group::2): plain, profiling
KernelFn const kKernels[kNumKernels] = {
    fused_beamform_kernel<false, false, false, false, false>, fused_beamform_kernel<false, false, false, true, false>,
    fused_beamform_kernel<true, false, false, false, false>,  fused_beamform_kernel<true, false, false, true, false>,
    fused_beamform_kernel<false, true, false, false, false>,  fused_beamform_kernel<false, true, false, true, false>,
    fused_beamform_kernel<false, false, true, false, false>,  fused_beamform_kernel<false, false, true, true, false>,
    fused_beamform_kernel<true, false, true, false, false>,   fused_beamform_kernel<true, false, true, true, false>,
    fused_beamform_kernel<false, true, true, false, false>,   fused_beamform_kernel<false, true, true, true, false>,
    fused_beamform_kernel<false, false, false, false, true>,  fused_beamform_kernel<true, false, false, false, true>,
    fused_beamform_kernel<false, true, false, false, true>,
    fused_beamform_kernel<false, false, true, false, true>,   fused_beamform_kernel<false, true, true, false, true>,
    fused_beamform_kernel<false, false, false, false, true, true>,
    fused_beamform_kernel<true, false, false, false, true, true>,
};

int* g_status_dev[64] = {};  // per-device: 4-int status block + kSchedSlots x {next, done} channel counters
unsigned long long* g_prof_dev = nullptr;  // set by fused_set_profile_buffer (developer aid)

}  // namespace

// Picks the N tiling: nt columns per tile (multiple of 16, <= 128) such that kb_count * parts * nt * 128 B <= 64 KiB.
// With more than one N tile nt is a multiple of 32 when the budget allows, so that the 32-column TMA store boxes
// never straddle two tiles.
static void pick_n_tiling(int A, int M, int parts, int* kb_count, int* nt, int* nt_count) {
    const int kbc = (A + kKbAnts - 1) / kKbAnts;
    const int n_pad = ((2 * M + 15) / 16) * 16;
    int nt_max = (kBopBufBytes / (kbc * parts * 128)) & ~15;
    if (nt_max > 128) nt_max = 128;
    *kb_count = kbc;
    if (nt_max <= 0) {
        *nt = 0;
        *nt_count = 0;
    } else if (n_pad <= nt_max) {
        *nt = n_pad;
        *nt_count = 1;
    } else {
        const int gran = (nt_max & ~31) >= 32 ? 32 : 16;
        const int cap = gran == 32 ? (nt_max & ~31) : nt_max;
        const int count = (n_pad + cap - 1) / cap;
        *nt_count = count;
        *nt = ((((n_pad + count - 1) / count) + gran - 1) / gran) * gran;
    }
}

// Status block layout (ints): [0..3] error code / role / barrier / CTA, [4..5] device-visible address of the host
// flag, [6..7] unused, [kStatusInts ...] the channel-queue counters.
constexpr int kStatusInts = 8;
static std::mutex g_init_mu;            // one-time per-device initialisation (status block, kernel attributes)
static int* g_status_host[64] = {};     // per-device: page-locked, device-mapped flag a failing kernel raises

int get_status_block(int** out) {
    int dev = 0;
    DCBF_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DCBF_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lock(g_init_mu);
    if (!g_status_dev[dev]) {
        int* p = nullptr;
        int* h = nullptr;
        int* h_dev = nullptr;
        DCBF_CUDA_TRY(cudaMalloc(&p, (kStatusInts + 2 * kSchedSlots) * sizeof(int)));
        DCBF_CUDA_TRY(cudaMemset(p, 0, (kStatusInts + 2 * kSchedSlots) * sizeof(int)));
        DCBF_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&h), 64, cudaHostAllocMapped | cudaHostAllocPortable));
        h[0] = 0;
        DCBF_CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h_dev), h, 0));
        DCBF_CUDA_TRY(cudaMemcpy(p + 4, &h_dev, sizeof(h_dev), cudaMemcpyHostToDevice));
        // the memset / copy above run on the legacy stream, which non-blocking streams do not wait for
        DCBF_CUDA_TRY(cudaDeviceSynchronize());
        g_status_host[dev] = h;
        g_status_dev[dev] = p;
    }
    *out = g_status_dev[dev];
    return DCBF_OK;
}

int get_encode_fn(EncodeTiledFn* out) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        DCBF_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) return record_cuda_error(cudaErrorNotSupported, "cuTensorMapEncodeTiled lookup");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    *out = fn;
    return DCBF_OK;
}

int launch_fused(const uint8_t* samples, const float* delay_vals, float* beams, int B, int A, int C, int N, int T,
                 int M, long long first_chan, double sample_period, const double* batch_dt_s, unsigned flags,
                 cudaStream_t s, const QuantisedOut* q8, const float* beam_weights, double sample_dt_s, int weights_log2,
                 int packed_mode, uint8_t* packed) {
    FusedParams p{};
    p.packed_mode = packed_mode;
    p.packed = packed;
    p.dv = reinterpret_cast<const float4*>(delay_vals);
    p.samples = samples;
    p.out = beams;
    p.weights = beam_weights;
    if (q8) {
        p.out_q8 = q8->beams;
        p.gains = q8->gains;
        p.saturated = q8->saturated;
    }
    p.B = B, p.A = A, p.C = C, p.T = T, p.M = M;
    p.parts = (flags & DCBF_FLAG_FP16_COEFF) ? 1 : 2;
    p.signed_in = (flags & DCBF_FLAG_SIGNED_INPUT) ? 1 : 0;
    {
        const uint32_t e_field = static_cast<uint32_t>(15 + (beam_weights ? weights_log2 : 0));  // fp16 exponent field, 1 .. 30
        const uint32_t half_bias = (e_field << 10) | (p.signed_in ? 0x80u : 0u);
        p.a_magic = (e_field << 2) * 0x01010101u;
        p.a_bias = half_bias * 0x00010001u;
        p.w_scale = std::ldexp(1.0f, -(beam_weights ? weights_log2 : 0));
    }
    pick_n_tiling(A, M, p.parts, &p.kb_count, &p.nt, &p.nt_count);
    const bool no_whole_tiles = p.nt < 16;  // > 512 antennas (hi+lo): not even a 16-column tile set fits a 64 KiB buffer
    p.slab_count = (A + kSlabAnts - 1) / kSlabAnts;
    p.inv_a = static_cast<uint32_t>((1ull << 32) / static_cast<unsigned>(A)) + 1u;
    p.ht_count = (T + kTileT - 1) / kTileT;
    p.chan_centre = static_cast<double>(first_chan) - static_cast<double>(N) / 2.0;
    p.turns_per_delay = -1.0 / (static_cast<double>(N) * sample_period);
    // TMA stores need a 16-byte row pitch (even beam count) and 32-column boxes that stay inside their N tile
    p.sb_count = 1;
    p.set_tiles = B * p.ht_count;
    p.hg_size = 2;
    if (batch_dt_s) {  // time-varying steering: one coefficient set per heap
        if (B > DCBF_MAX_TV_BATCHES) return DCBF_ERR_UNSUPPORTED;
        const bool per_tile = sample_dt_s != 0.0 && p.ht_count > 1;
        p.sb_count = per_tile ? B * p.ht_count : B;
        p.set_tiles = per_tile ? 1 : p.ht_count;
        p.hg_size = per_tile ? 1 : 2;
        p.sample_dt = sample_dt_s;
        for (int b = 0; b < B; ++b) p.dt_s[b] = batch_dt_s[b];
    }
    // Many antennas x beams: a whole B tile set no longer fits 64 KiB with a useful width (the voltages would be
    // re-converted for every narrow N tile).  Stream B by 32-antenna k-blocks instead: N tiles of up to 128 columns.
    const bool kstream = (p.nt_count > 1 || no_whole_tiles) && !(flags & DCBF_FLAG_DEBUG_NO_KSTREAM);
    p.hg_count = (p.ht_count + p.hg_size - 1) / p.hg_size;
    if (no_whole_tiles && !kstream) return DCBF_ERR_UNSUPPORTED;  // (the k-block ring itself has no antenna limit)
    // CTA pairs for the plain K-streamed case with at least two time tiles and more than 128 output columns: the two CTAs
    // of a cluster take one time tile each and half of the coefficients each, N tiles of up to 256 columns (one
    // cta_group::2 MMA of M = 256): per output byte half the coefficient and conversion work and ~40 % less shared-memory
    // traffic (every B row is read by the tensor cores once for 256 output rows instead of once for 128).
    const int n_pad = ((2 * M + 15) / 16) * 16;
    const bool pair = kstream && !batch_dt_s && !q8 && T > kTileT && n_pad > 128 && !(flags & DCBF_FLAG_DEBUG_NO_PAIR);
    if (kstream) {
        const int cap = pair ? 256 : 128;
        p.nt_count = (n_pad + cap - 1) / cap;
        p.nt = ((((n_pad + p.nt_count - 1) / p.nt_count) + 31) / 32) * 32;
    }
    p.merged = !kstream && p.parts == 2 && p.nt <= 64;
    // Raw ring depth: narrow N tiles leave the tail of both 64 KiB B buffers unused; every spare 8 KiB becomes one
    // more TMA stage in flight (HBM latency x bandwidth per SM is ~40 KiB, the dedicated 4 stages hold 32 KiB)
    p.raw_stages = kRawStages;
    if (!kstream) {
        const int used = p.kb_count * p.parts * p.nt * 128;
        // (the raw ring comes first: at C4, where 16 KiB per buffer are spare, two more A stages instead of four more raw
        // stages cost 3 %; with both -- C2, 32 beams -- the extra A stages are worth 3-4 %)
        p.aop_extra = kBopBufBytes - used >= kAopStageBytes + ((kMaxRawStages - kRawStages) / 2) * kRawStageBytes &&
                      !(flags & DCBF_FLAG_DEBUG_TWO_A_STAGES);
        const int spare = (kBopBufBytes - used - (p.aop_extra ? kAopStageBytes : 0)) / kRawStageBytes;
        p.raw_extra_off = used;
        p.raw_stages = kRawStages + 2 * std::min(spare, (kMaxRawStages - kRawStages) / 2);
    }
    if (packed_mode) {  // whole tile sets, static steering (the shapes packed_bytes() answers for)
        if (kstream || batch_dt_s || !packed) return DCBF_ERR_UNSUPPORTED;
        p.packed_bytes = p.kb_count * p.parts * p.nt * 128;
        if (packed_mode == 1) flags |= DCBF_FLAG_DEBUG_WHOLE_CHANNELS;  // every channel is generated (and written) once
        else flags |= DCBF_FLAG_DEBUG_NO_BEAM_PIECES;                   // a piece of a cut channel loads the whole tile set
    }
    p.tma_store = !(flags & DCBF_FLAG_DEBUG_DIRECT_EPILOGUE) && (M % (q8 ? 8 : 2) == 0) && (p.nt_count == 1 || p.nt % 32 == 0);
    if (static_cast<long long>(B) * kPols * C > 0x7fffffffLL) return DCBF_ERR_UNSUPPORTED;
    if (int e = get_status_block(&p.status)) return e;
    static std::atomic<unsigned> ticket[64], captured_ticket[64];  // per device, like the block the slots live in
    int slot_dev = 0;
    DCBF_CUDA_TRY(cudaGetDevice(&slot_dev));
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    DCBF_CUDA_TRY(cudaStreamIsCapturing(s, &capturing));
    const unsigned slot = capturing == cudaStreamCaptureStatusActive
                              ? kLiveSlots + captured_ticket[slot_dev & 63].fetch_add(1, std::memory_order_relaxed) % kLiveSlots
                              : ticket[slot_dev & 63].fetch_add(1, std::memory_order_relaxed) % kLiveSlots;
    p.sched = p.status + kStatusInts + 2 * slot;
    p.prof = g_prof_dev;

    EncodeTiledFn encode = nullptr;
    if (int e = get_encode_fn(&encode)) return e;
    alignas(64) CUtensorMap tm_in, tm_out;
    {
        // samples as 4-byte words {p0.re, p0.im, p1.re, p1.im}: [B][A][C][T], box [1][16][1][128]
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(A),
                                    static_cast<cuuint64_t>(B)};
        const cuuint64_t strides[3] = {static_cast<cuuint64_t>(T) * 4, static_cast<cuuint64_t>(C) * T * 4,
                                       static_cast<cuuint64_t>(A) * C * T * 4};
        const cuuint32_t box[4] = {kTileT, 1, kSlabAnts, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = encode(&tm_in, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<uint8_t*>(samples), dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return record_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(samples)");
    }
    p.q8_wide = q8 && !kstream && p.tma_store && p.nt_count == 1 && (p.nt == 32 || p.nt == 64 || p.nt == 128);
    if (p.tma_store && q8) {
        // requantised beams as [B*2*C][T][2M] int8; box [1][32][32] with 32B swizzle, or -- when one N tile is the
        // whole row -- box [1][32][nt] with the swizzle span equal to the row length
        const cuuint64_t dims[3] = {static_cast<cuuint64_t>(2 * M), static_cast<cuuint64_t>(T),
                                    static_cast<cuuint64_t>(B) * kPols * C};
        const cuuint64_t strides[2] = {static_cast<cuuint64_t>(2 * M), static_cast<cuuint64_t>(T) * 2 * M};
        const cuuint32_t box[3] = {static_cast<cuuint32_t>(p.q8_wide ? p.nt : 32), 32, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUtensorMapSwizzle sw = !p.q8_wide || p.nt == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                      : p.nt == 64             ? CU_TENSOR_MAP_SWIZZLE_64B
                                                               : CU_TENSOR_MAP_SWIZZLE_128B;
        const CUresult r = encode(&tm_out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, q8->beams, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return record_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(beams_q8)");
    } else if (p.tma_store) {
        // beams as [B*2*C][T][2M] fp32, box [1][32][32], 128B swizzle
        const cuuint64_t dims[3] = {static_cast<cuuint64_t>(2 * M), static_cast<cuuint64_t>(T),
                                    static_cast<cuuint64_t>(B) * kPols * C};
        const cuuint64_t strides[2] = {static_cast<cuuint64_t>(2 * M) * 4, static_cast<cuuint64_t>(T) * 2 * M * 4};
        const cuuint32_t box[3] = {32, 32, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, beams, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return record_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(beams)");
    } else {
        tm_out = tm_in;  // never dereferenced
    }

    static int n_sms[64] = {};
    int dev = 0;
    DCBF_CUDA_TRY(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(g_init_mu);
        if (!n_sms[dev]) {
            for (int i = 0; i < kNumKernels; ++i)
                DCBF_CUDA_TRY(cudaFuncSetAttribute(kKernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
            DCBF_CUDA_TRY(cudaDeviceGetAttribute(&n_sms[dev], cudaDevAttrMultiProcessorCount, dev));
        }
    }
    // Whole-tile-set mode: C channels over G persistent CTAs leave a last round of R = C mod G channels in which
    // G - R CTAs idle for a whole channel time (C3 cut over 8 GPUs: 512 channels = 3.46 per CTA, i.e. 4 rounds).  The
    // last R channels are therefore cut into `split` units each along their accumulator tiles (the coefficients of a
    // cut channel are generated once per piece, which is why only the tail is cut).
    p.tile_count = p.nt_count * B * p.ht_count;
    p.n_whole = C;
    p.split = 1;
    p.bsplit = 1;
    if (!kstream && !(flags & (DCBF_FLAG_DEBUG_WHOLE_CHANNELS | DCBF_FLAG_STREAMING))) {  // (overlapped launches balance themselves)
        const int rem = C % n_sms[dev];
        // float32 output through TMA stores, one N tile of 128 or 256 columns with beams in both halves: a piece can be one
        // half of the beams (64-column MMAs; the voltages are converted once per piece either way)
        const bool by_beams = !q8 && p.tma_store && !p.merged && p.nt_count == 1 && p.nt % 64 == 0 && 4 * M > p.nt &&
                              !(flags & DCBF_FLAG_DEBUG_NO_BEAM_PIECES);
        const int room = rem ? n_sms[dev] / rem : 1;  // pieces per cut channel that still keep every CTA at one piece
        if (by_beams && room >= 2) {
            p.n_whole = C - rem;
            p.bsplit = 2;
            p.split = 2 * std::min(p.tile_count, room / 2);
        } else if (std::min(p.tile_count, room) > 1) {
            p.n_whole = C - rem;
            p.split = std::min(p.tile_count, room);
        }
    }
    const long long units = kstream ? static_cast<long long>(C) * p.nt_count * p.hg_count
                                    : static_cast<long long>(p.n_whole) + static_cast<long long>(C - p.n_whole) * p.split;
    if (units > 0x7ffffff0LL) return DCBF_ERR_UNSUPPORTED;  // what the CTAs draw from the queue
    int grid = units < n_sms[dev] ? static_cast<int>(units) : n_sms[dev];
    if (pair) {  // one cluster of two CTAs per TPC the driver will co-schedule
        static int n_pairs[64] = {};
        if (!n_pairs[dev]) {
            cudaLaunchConfig_t occ{};
            occ.gridDim = dim3(2 * n_sms[dev]);
            occ.blockDim = dim3(kThreads);
            occ.dynamicSmemBytes = kSmemBytes;
            cudaLaunchAttribute ca[1];
            ca[0].id = cudaLaunchAttributeClusterDimension;
            ca[0].val.clusterDim.x = 2, ca[0].val.clusterDim.y = 1, ca[0].val.clusterDim.z = 1;
            occ.attrs = ca;
            occ.numAttrs = 1;
            int n = 0;
            DCBF_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, kKernels[17], &occ));
            n_pairs[dev] = n > 0 ? n : n_sms[dev] / 2;
        }
        grid = 2 * static_cast<int>(units < n_pairs[dev] ? units : n_pairs[dev]);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(q8 ? kThreadsQ8 : kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2, attr[1].val.clusterDim.y = 1, attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    // Every launch may start while the preceding kernel of the stream drains (programmatic dependent launch).  By
    // default the kernel then waits for that kernel's completion (griddepcontrol.wait) right after its prologue, before
    // it touches global memory: launch latency, barrier / TMEM set-up and the zeroing of the B tiles are off the
    // critical path, nothing else changes.  DCBF_FLAG_STREAMING drops the wait (the caller promises independence).
    cfg.numAttrs = (flags & DCBF_FLAG_DEBUG_NO_PDL) ? 0 : 1;
    if (pair) {
        if (!cfg.numAttrs) attr[0] = attr[1];
        ++cfg.numAttrs;
    }
    p.pdl_wait = (flags & DCBF_FLAG_STREAMING) ? 0 : 1;
    p.dbg = (flags >> 16) & 127;  // developer experiments
    // (int8 output, variant, merged) specialisation; variant: 0 plain, 1 profiling, 2 time-varying steering (the
    // profiler has no time-varying build)
    const int variant = batch_dt_s ? 2 : p.prof ? 1 : 0;
    auto kernel = pair ? kKernels[17 + (p.prof ? 1 : 0)] : kstream ? kKernels[q8 ? (batch_dt_s ? 16 : 15) : batch_dt_s ? 14 : 12 + (p.prof ? 1 : 0)] : kKernels[(q8 ? 6 : 0) + 2 * variant + (p.merged ? 1 : 0)];
    DCBF_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, p, tm_in, tm_out));
    DCBF_CHECK_LAUNCH("fused_beamform_kernel");
    return DCBF_OK;
}

// Non-blocking: has any tcgen05 kernel of this device raised its watchdog since the status was last cleared?  (Reads a
// page-locked flag the failing kernel writes; meaningful for work the caller has already synchronised with.)
int fused_status_poll() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return DCBF_OK;
    const int* h = g_status_host[dev];
    return h ? *const_cast<const volatile int*>(h) : DCBF_OK;
}

int fused_status(int* role, int* barrier, int* block) {
    int* blk = nullptr;
    if (int e = get_status_block(&blk)) return e;
    // every stream of the device, non-blocking ones included (a copy on the legacy stream would not wait for those)
    DCBF_CUDA_TRY(cudaDeviceSynchronize());
    int h[4] = {};
    DCBF_CUDA_TRY(cudaMemcpy(h, blk, sizeof(h), cudaMemcpyDeviceToHost));
    if (h[0] != 0) {  // clear the status and re-arm the channel counters (an aborted launch leaves them mid-count)
        DCBF_CUDA_TRY(cudaMemset(blk, 0, 4 * sizeof(int)));
        DCBF_CUDA_TRY(cudaMemset(blk + kStatusInts, 0, 2 * kSchedSlots * sizeof(int)));
        DCBF_CUDA_TRY(cudaDeviceSynchronize());
        int dev = 0;
        DCBF_CUDA_TRY(cudaGetDevice(&dev));
        if (g_status_host[dev]) *const_cast<volatile int*>(g_status_host[dev]) = 0;
    }
    if (role) *role = h[1];
    if (barrier) *barrier = h[2];
    if (block) *block = h[3];
    return h[0];
}

void fused_set_profile_buffer(unsigned long long* dev_ptr) { g_prof_dev = dev_ptr; }

void fused_tiling(int A, int M, unsigned flags, int* kb_count, int* nt, int* nt_count) {
    pick_n_tiling(A, M, (flags & DCBF_FLAG_FP16_COEFF) ? 1 : 2, kb_count, nt, nt_count);
}

}  // namespace dcbf
