// Stage 3 (stand-alone) on the tensor cores: coefficient x voltage contraction on MATERIALISED operands.
//
//   out[u, t, n] = sum_{j < 2A} f32(reordered[u, t, j]) * coeffs[u, j, n],      u = (b, p, c),  n < 2M
//
// Replaces kernel `run_complex_mult` (reference: beamformer/beamforming/complex_mult_kernel.py:11-100) behind
// `MatrixMultiply._run` (matrix_multiply.py:155-163).  The coefficient slot is an INPUT of this operator, so the
// values are arbitrary float32, not the unit-modulus pairs the fused kernel generates itself.  They are split on
// the fly into THREE bfloat16 terms (hi + mid + lo = 24 significand bits, full float32 exponent range), the 8-bit
// voltages are exact in bfloat16 too, and each product term is a kind::f16 tcgen05.mma (bf16 x bf16) accumulated
// in float32 in TMEM: the result differs from a float32 evaluation only by accumulation order.
//
// HBM-bound: per unit it reads T x 2A bytes + 2A x 2M x 4 bytes and writes T x 2M x 4 bytes, each once.
//
// One persistent CTA per SM, 20 warps:
//   warps 0-7   split     W k-block (TMA-staged float32, 128B-swizzled 32x32 boxes) -> 3 bf16 K-major B tiles
//   warps 8-11  convert   X stage ([128 t][32 B] bytes, 32B-swizzled) -> bf16 K-major A stage (64B swizzle)
//   warps 12-15 epilogue  TMEM -> registers -> 128B-swizzled staging -> TMA tensor store
//   warp 16     X producer (TMA), warp 17 W producer (TMA), warps 18-19 MMA issuers (one per time tile of a group)
// A work item is (unit, N tile of <= 128 columns, group of <= 2 time tiles); its k-blocks of 32 contraction
// elements (16 antennas) stream through a 2-slot B ring while both time tiles' accumulators stay open, and the
// accumulators are double-buffered across items so the epilogue of one overlaps the MMAs of the next.
//
// Antenna counts whose sample rows (2A bytes) are not a multiple of 16 bytes are fetched through an [8 samples x 2A]
// view, eight small boxes per stage.  Odd beam counts (8M bytes per coefficient / output row not a multiple of 16:
// no tensor map can describe those rows) run here too: the split warps then read their coefficients straight from
// global memory (8 consecutive columns per quarter-warp: whole 32-byte sectors) instead of TMA-staged boxes, and the
// epilogue stores from registers (tcgen05.ld 16x256b -> st.global.v2, as the fused kernel does for odd beam counts).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dcbf {

namespace {

constexpr int kSplitWarps = 8;
constexpr int kConvertWarp0 = 8;
constexpr int kEpilogueWarp0 = 12;  // warp % 4 = TMEM lane quarter
constexpr int kXProducerWarp = 16;
constexpr int kWProducerWarp = 17;
constexpr int kMmaWarp = 18;   // MMAs of a group's first time tile (and owns the TMEM allocation)
constexpr int kMmaWarp2 = 19;  // second time tile: the issue path, not the tensor pipe, limits narrow tiles
constexpr int kThreads = 20 * 32;
constexpr int kTileT = 128;  // samples per MMA tile (UMMA M)
constexpr int kKb = 32;      // contraction elements per k-block: 16 antennas x (re, im) = two K = 16 MMA steps
constexpr int kParts = 3;    // bf16 hi + mid + lo
constexpr int kNtMax = 128;
constexpr int kBSlots = 2;
constexpr int kBSlotBytes = kParts * kNtMax * 64;  // [part][n][32 bf16], 64B swizzle
constexpr int kWStages = 4;
constexpr int kWBoxBytes = 32 * 128;               // 32 k rows x 32 float32 columns, 128B swizzle
constexpr int kWStageBytes = (kNtMax / 32) * kWBoxBytes;
constexpr int kXStages = 12;
constexpr int kXStageBytes = kTileT * kKb;         // [t][32 bytes], 32B swizzle
constexpr int kXSplitRowBytes = kKb + 16;          // split mode: a stage is 8 boxes [16 rows][48 B] (32 wanted bytes at a
constexpr int kXSplitBoxBytes = 16 * kXSplitRowBytes;  // 2-byte-granular offset inside a 16-byte-aligned window)
constexpr int kXSplitStageBytes = 8 * kXSplitBoxBytes;
constexpr int kXSplitStages = kXStages * kXStageBytes / kXSplitStageBytes;
constexpr int kAStages = 4;
constexpr int kAStageBytes = kTileT * 64;          // [t][32 bf16], 64B swizzle
constexpr int kOutBoxBytes = 32 * 128;             // 32 rows x 32 float32 columns, 128B swizzle
constexpr int kOutStageBytes = 4 * 2 * kOutBoxBytes;
constexpr int kAccBufs = 2;
constexpr int kGroupTiles = 2;  // time tiles whose accumulators are open together
constexpr int kTmemCols = 512;  // kAccBufs x kGroupTiles x kNtMax

constexpr int kSmemData = kBSlots * kBSlotBytes + kWStages * kWStageBytes + kOutStageBytes + kAStages * kAStageBytes +
                          kXStages * kXStageBytes;
constexpr int kBarBytes = 384;
constexpr int kCtlBytes = 640;
constexpr int kSmemBytes = 1024 /*alignment slack*/ + kSmemData + kCtlBytes;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

enum Role { kRoleXProducer = 1, kRoleMma = 2, kRoleEpilogue = 3, kRoleConvert = 4, kRoleSplit = 5, kRoleWProducer = 6 };

struct TcParams {
    int* status;      // shared with the fused kernel: [0] = error code, [1] = role, [2] = barrier id, [3] = CTA
    long long items;  // units x N tiles x time-tile groups
    int T, K2, N2;    // samples, 2 x antennas, 2 x beams
    int kb_count;     // ceil(K2 / 32)
    int nt, nt_count; // columns per N tile (multiple of 32, <= 128), number of N tiles
    int ht_count;     // ceil(T / 128)
    int hg_count;     // ceil(ht_count / 2)
    int signed_in;
    int x_split;      // 2A is not a multiple of 16 bytes: X rows are fetched as 8 interleaved boxes (see launch_beamform_tc)
    int mma_warps;    // 2: one issuing warp per time tile of a group (narrow tiles are issue-bound); 1: wide tiles
    int direct;       // odd beam count: coefficients by plain loads, beams by plain stores (rows are only 8-byte aligned)
    const float* w;   // coefficients [units][K2][N2] (direct mode)
    float* out;       // beams [units][T][N2] (direct mode)
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// Bytes kLo, kLo + 1 of w (u8, or i8 re-biased by ^0x80) -> bfloat16 pair, exact (|value| <= 255 needs 8 significand
// bits).  bf16 has no room for the fp16-style 0x64bb trick, so: byte -> 0x4B0000bb = 2^23 + b (float32), minus the
// bias (2^23, or 2^23 + 128 for i8), then one packed conversion.
template <int kLo>
__device__ __forceinline__ uint32_t byte_pair_to_bf16x2(uint32_t w, float bias) {
    const float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u + kLo)) - bias;
    const float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7441u + kLo)) - bias;
    return pack_bf16x2(f0, f1);
}

__global__ void __launch_bounds__(kThreads, 1)
beamform_tc_kernel(const __grid_constant__ TcParams prm, const __grid_constant__ CUtensorMap tm_x,
                   const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const uint32_t b_base = smem_base;                                // [slot][part][n][64 B]
    const uint32_t w_base = b_base + kBSlots * kBSlotBytes;           // [stage][n / 32][k][128 B]
    const uint32_t ost_base = w_base + kWStages * kWStageBytes;       // [warp][2][32 x 128 B]
    const uint32_t a_base = ost_base + kOutStageBytes;                // [stage][t][64 B]
    const uint32_t x_base = a_base + kAStages * kAStageBytes;         // [stage][t][32 B]
    const uint32_t bar_base = x_base + kXStages * kXStageBytes;
    Control* ctl = reinterpret_cast<Control*>(smem_gen + kSmemData + kBarBytes);
    static_assert(kBarBytes + sizeof(Control) <= kCtlBytes, "control area");

    // barrier ids (also reported by the watchdog)
    constexpr int kXFull = 0, kXEmpty = kXFull + kXStages, kWFull = kXEmpty + kXStages, kWEmpty = kWFull + kWStages,
                  kAFull = kWEmpty + kWStages, kAEmpty = kAFull + kAStages, kBFull = kAEmpty + kAStages,
                  kBEmpty = kBFull + kBSlots, kAccFull = kBEmpty + kBSlots, kAccEmpty = kAccFull + kAccBufs,
                  kNumBars = kAccEmpty + kAccBufs;
    static_assert(kNumBars * 8 <= kBarBytes, "barrier area");
    auto bar = [&](int id) { return bar_base + 8u * static_cast<uint32_t>(id); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kXStages; ++s) {
            mbar_init(bar(kXFull + s), 1);
            mbar_init(bar(kXEmpty + s), 4);
        }
        for (int s = 0; s < kWStages; ++s) {
            mbar_init(bar(kWFull + s), 1);
            mbar_init(bar(kWEmpty + s), kSplitWarps);
        }
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(bar(kAFull + s), 4);
            mbar_init(bar(kAEmpty + s), 1);
        }
        for (int s = 0; s < kBSlots; ++s) {
            mbar_init(bar(kBFull + s), kSplitWarps);
            mbar_init(bar(kBEmpty + s), prm.mma_warps);  // one tcgen05.commit per MMA warp
        }
        for (int s = 0; s < kAccBufs; ++s) {
            mbar_init(bar(kAccFull + s), prm.mma_warps);
            mbar_init(bar(kAccEmpty + s), 4);
        }
        ctl->abort = 0;
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(&ctl->tmem_base), kTmemCols);
    if (warp == kXProducerWarp && lane == 0) prefetch_tensormap(&tm_x);
    if (warp == kWProducerWarp && lane == 0) prefetch_tensormap(&tm_w);
    if (warp == kEpilogueWarp0 && lane == 0) prefetch_tensormap(&tm_out);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;

    const int nt = prm.nt, kb_count = prm.kb_count;
    const uint32_t x_stages = prm.x_split ? kXSplitStages : kXStages, x_pitch = prm.x_split ? kXSplitStageBytes : kXStageBytes;
    const uint32_t part_bytes = static_cast<uint32_t>(nt) * 64u;
    const long long items = prm.items, stride = gridDim.x;
    const int per_unit = prm.nt_count * prm.hg_count;
    // item -> (unit, N tile, time-tile group); N tiles and groups of one unit are adjacent so X / W re-reads hit L2
    auto decode = [&](long long item, int* u, int* it, int* hg, int* hn) {
        const int r = static_cast<int>(item % per_unit);
        *u = static_cast<int>(item / per_unit);
        *it = r / prm.hg_count;
        *hg = r % prm.hg_count;
        *hn = min(kGroupTiles, prm.ht_count - *hg * kGroupTiles);
    };
    bool ok = true;

    if (warp == kXProducerWarp) {
        // =================================== X producer ===================================
        uint32_t xs = 0, ph = 0;
        for (long long item = blockIdx.x; item < items && ok; item += stride) {
            int u, it, hg, hn;
            decode(item, &u, &it, &hg, &hn);
            for (int s = 0; s < kb_count && ok; ++s)
                for (int h = 0; h < hn; ++h) {
                    ok = mbar_wait<false>(bar(kXEmpty + xs), ph ^ 1u, ctl, prm.status, kRoleXProducer, kXEmpty + xs);
                    if (!ok) break;
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar(kXFull + xs), x_pitch);
                        if (!prm.x_split) {
                            tma_load_3d(x_base + xs * x_pitch, &tm_x, bar(kXFull + xs), s * kKb, (hg * kGroupTiles + h) * kTileT, u);
                        } else {  // rows of 8 samples: sample 8 r + i of the tile is bytes [i * 2A, (i + 1) * 2A) of row r
                            // (the innermost TMA coordinate must be a multiple of 16 bytes: fetch the aligned 48-byte window)
                            for (int i = 0; i < 8; ++i)
                                tma_load_3d(x_base + xs * x_pitch + i * kXSplitBoxBytes, &tm_x, bar(kXFull + xs),
                                            (i * prm.K2 + s * kKb) & ~15, (hg * kGroupTiles + h) * (kTileT / 8), u);
                        }
                    }
                    __syncwarp();
                    if (++xs == x_stages) xs = 0, ph ^= 1u;
                }
        }
    } else if (warp == kWProducerWarp) {
        // =================================== W producer ===================================
        uint32_t ws = 0, ph = 0;
        const int boxes = nt >> 5;
        for (long long item = blockIdx.x; item < items && ok && !prm.direct; item += stride) {
            int u, it, hg, hn;
            decode(item, &u, &it, &hg, &hn);
            for (int s = 0; s < kb_count; ++s) {
                ok = mbar_wait<false>(bar(kWEmpty + ws), ph ^ 1u, ctl, prm.status, kRoleWProducer, kWEmpty + ws);
                if (!ok) break;
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar(kWFull + ws), static_cast<uint32_t>(boxes) * kWBoxBytes);
                    for (int g = 0; g < boxes; ++g)  // columns / rows outside the tensor are zero-filled
                        tma_load_3d(w_base + ws * kWStageBytes + g * kWBoxBytes, &tm_w, bar(kWFull + ws), it * nt + g * 32,
                                    s * kKb, u);
                }
                __syncwarp();
                if (++ws == kWStages) ws = 0, ph ^= 1u;
            }
        }
    } else if (warp == kMmaWarp || (warp == kMmaWarp2 && prm.mma_warps == 2)) {
        // =================================== MMA issuers ===================================
        const int my_h = prm.mma_warps == 1 ? -1 : warp == kMmaWarp ? 0 : 1;  // -1: every time tile
        const uint32_t idesc = make_idesc_f16(nt, true, true);  // bf16 x bf16 (kind::f16 rejects mixed fp16 / bf16 operands)
        const uint32_t a_lo0 = desc_lo(a_base), b_lo0 = desc_lo(b_base), part_lo = part_bytes >> 4;
        uint32_t as = 0, aph = 0, bs = 0, bph = 0, n_item = 0;
        for (long long item = blockIdx.x; item < items && ok; item += stride, ++n_item) {
            int u, it, hg, hn;
            decode(item, &u, &it, &hg, &hn);
            const uint32_t ab = n_item & 1u;
            ok = mbar_wait<false>(bar(kAccEmpty + ab), ((n_item >> 1) & 1u) ^ 1u, ctl, prm.status, kRoleMma, kAccEmpty + ab);
            if (!ok) break;
            tc_fence_after();
            for (int s = 0; s < kb_count && ok; ++s) {
                ok = mbar_wait<false>(bar(kBFull + bs), bph, ctl, prm.status, kRoleMma, kBFull + bs);
                const int k_steps = min(2, (prm.K2 - s * kKb + 15) >> 4);
                const uint32_t b_lo = b_lo0 + bs * (kBSlotBytes >> 4);
                for (int h = 0; h < hn && ok; ++h) {
                    if (my_h >= 0 && h != my_h) {  // the other warp's stage
                        if (++as == kAStages) as = 0, aph ^= 1u;
                        continue;
                    }
                    ok = mbar_wait<false>(bar(kAFull + as), aph, ctl, prm.status, kRoleMma, kAFull + as);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + as * (kAStageBytes >> 4);
                    const uint32_t d_tmem = tmem_base + (ab * kGroupTiles + static_cast<uint32_t>(h)) * kNtMax;
                    if (elect_one()) {
#pragma unroll
                        for (int part = 0; part < kParts; ++part)
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
                                if (kk < k_steps)
                                    umma_f16(d_tmem, make_desc(a_lo + 2u * kk, kDescHiSw64),
                                             make_desc(b_lo + part * part_lo + 2u * kk, kDescHiSw64), idesc, (s | part | kk) != 0);
                        umma_commit(bar(kAEmpty + as));  // A stage free once these MMAs retire
                    }
                    __syncwarp();
                    if (++as == kAStages) as = 0, aph ^= 1u;
                }
                if (ok && elect_one()) umma_commit(bar(kBEmpty + bs));
                __syncwarp();
                if (++bs == kBSlots) bs = 0, bph ^= 1u;
            }
            if (ok && elect_one()) umma_commit(bar(kAccFull + ab));
            __syncwarp();
        }
    } else if (warp >= kEpilogueWarp0 && warp < kEpilogueWarp0 + 4) {
        // =================================== epilogue ===================================
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        const uint32_t ost = ost_base + static_cast<uint32_t>(q) * (2 * kOutBoxBytes);
        uint32_t n_item = 0, box = 0;
        for (long long item = blockIdx.x; item < items && ok; item += stride, ++n_item) {
            int u, it, hg, hn;
            decode(item, &u, &it, &hg, &hn);
            const uint32_t ab = n_item & 1u;
            ok = mbar_wait<false>(bar(kAccFull + ab), (n_item >> 1) & 1u, ctl, prm.status, kRoleEpilogue, kAccFull + ab);
            if (!ok) break;
            tc_fence_after();
            for (int h = 0; h < hn; ++h) {
                const int row0 = (hg * kGroupTiles + h) * kTileT + 32 * q;  // this warp's 32 rows of the tile
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + (ab * kGroupTiles + static_cast<uint32_t>(h)) * kNtMax;
                if (prm.direct) {
                    // 16 rows x 16 columns per load: lane holds row lane/4 (+8), columns 2 (lane%4) + {0,1} (+8): a quad
                    // writes 32 contiguous bytes of a row
                    const int n0 = it * nt;
                    for (int half = 0; half < 2; ++half) {
                        const int r_lo = row0 + 16 * half + (lane >> 2);
                        const uint32_t ta = taddr + (static_cast<uint32_t>(16 * half) << 16);
                        float* row_lo = prm.out + (static_cast<size_t>(u) * prm.T + r_lo) * prm.N2 + n0 + 2 * (lane & 3);
                        float* row_hi = row_lo + 8 * static_cast<size_t>(prm.N2);
                        const bool v_lo = r_lo < prm.T, v_hi = r_lo + 8 < prm.T;
                        for (int cb = 0; cb < nt; cb += 16) {
                            uint32_t r[8];
                            tmem_ld_16x256b_x2(ta + cb, r);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const int col = cb + 8 * i;
                                if (n0 + col + 2 * (lane & 3) < prm.N2) {  // (N2 is even: a pair is inside or outside as a whole)
                                    if (v_lo) st_global_v2(row_lo + col, r[4 * i], r[4 * i + 1]);
                                    if (v_hi) st_global_v2(row_hi + col, r[4 * i + 2], r[4 * i + 3]);
                                }
                            }
                        }
                    }
                    continue;
                }
                for (int cb = 0; cb < nt && row0 < prm.T && it * nt + cb < prm.N2; cb += 32, ++box) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr + cb, r);
                    const uint32_t sb = ost + (box & 1u) * kOutBoxBytes;
                    bulk_wait_group_read<1>();  // (issuing lane) the store that last read this box is done
                    __syncwarp();
                    tmem_wait_ld();
                    const uint32_t dst = sb + lane * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_v4(dst + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (elect_one()) {
                        tma_store_3d(&tm_out, sb, it * nt + cb, row0, u);
                        bulk_commit_group();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kAccEmpty + ab));
        }
        bulk_wait_group_all();  // (issuing lane) staging memory and the stores themselves are done before exit
    } else if (warp >= kConvertWarp0 && warp < kConvertWarp0 + 4) {
        // =================================== convert ===================================
        // thread = one sample row t: 2 LDS.128 (32B swizzle: conflict-free), 32 PRMT + 32 FADD + 16 F2F, 4 STS.128
        const int t = threadIdx.x - kConvertWarp0 * 32;
        const float bias = prm.signed_in ? 8388736.0f : 8388608.0f;  // 2^23 (+ 128 after the ^0x80 re-bias of i8)
        const uint32_t flip = prm.signed_in ? 0x80808080u : 0u;
        const uint32_t sw_in = static_cast<uint32_t>((t >> 2) & 1), sw_out = static_cast<uint32_t>((t >> 1) & 3);
        // the two 16-byte halves of this thread's row inside a stage.  Split mode: sample t = 8 r + i is row r of box i,
        // its 32 bytes start (i * 2A) % 16 bytes into the row's 48-byte window: nine 4-byte loads and a funnel shift.
        // TMA destinations must be 128-byte aligned, so the boxes cannot be skewed against the banks and the eight
        // lanes of a quarter-warp (eight boxes, same row) collide -- the price of an antenna count TMA cannot address.
        const uint32_t src0 = static_cast<uint32_t>(t * 32) + ((0u ^ sw_in) << 4), src1 = static_cast<uint32_t>(t * 32) + ((1u ^ sw_in) << 4);
        const uint32_t split_off = static_cast<uint32_t>((t & 7) * prm.K2) & 15u;
        const uint32_t split_src = static_cast<uint32_t>((t & 7) * kXSplitBoxBytes + (t >> 3) * kXSplitRowBytes) + (split_off & ~3u);
        const uint32_t split_shift = (split_off & 2u) * 8u;
        uint32_t xs = 0, xph = 0, as = 0, aph = 0;
        for (long long item = blockIdx.x; item < items && ok; item += stride) {
            int u, it, hg, hn;
            decode(item, &u, &it, &hg, &hn);
            for (int sh = 0; sh < kb_count * hn; ++sh) {
                ok = mbar_wait2<false>(bar(kXFull + xs), xph, kXFull + xs, bar(kAEmpty + as), aph ^ 1u, kAEmpty + as, ctl,
                                       prm.status, kRoleConvert);
                if (!ok) break;
                // antennas beyond A and samples beyond T were zero-filled by the TMA box: byte 0 -> value 0 (u8),
                // and 0 ^ 0x80 - 128 -> 0 (i8), so the operand padding needs no special case (split mode: the bytes
                // past 2A belong to the next sample, but they only ever meet the zero-filled rows k >= 2A of W)
                const uint32_t src = x_base + xs * x_pitch;
                uint32_t w[8];
                {
                    if (!prm.x_split) {
                        uint32_t c0[4], c1[4];
                        ld_shared_v4(src + src0, c0);
                        ld_shared_v4(src + src1, c1);
#pragma unroll
                        for (int i = 0; i < 4; ++i) w[i] = c0[i] ^ flip, w[4 + i] = c1[i] ^ flip;
                    } else {
                        uint32_t c[9];
#pragma unroll
                        for (int i = 0; i < 9; ++i) c[i] = ld_shared_u32(src + split_src + 4u * i);
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = __funnelshift_r(c[i], c[i + 1], split_shift) ^ flip;
                    }
                }
                const uint32_t dst = a_base + as * kAStageBytes + t * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    st_shared_v4(dst + ((static_cast<uint32_t>(j) ^ sw_out) << 4), byte_pair_to_bf16x2<0>(w[2 * j], bias),
                                 byte_pair_to_bf16x2<2>(w[2 * j], bias), byte_pair_to_bf16x2<0>(w[2 * j + 1], bias),
                                 byte_pair_to_bf16x2<2>(w[2 * j + 1], bias));
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar(kAFull + as));
                    mbar_arrive(bar(kXEmpty + xs));
                }
                if (++xs == x_stages) xs = 0, xph ^= 1u;
                if (++as == kAStages) as = 0, aph ^= 1u;
            }
        }
    } else if (warp < kSplitWarps) {
        // =================================== coefficient split ===================================
        // A warp step covers 8 k x 8 n of the k-block: lane = (k pair kp, column nn).  Both the 128B-swizzled float32
        // reads (rows 2kp, 2kp+1 of an 8-row group) and the 64B-swizzled bf16-pair writes touch 32 distinct banks.
        const int kp = lane & 3, nn = lane >> 2;
        const int steps = nt >> 1;  // (k group of 8) x (column group of 8) = 4 x nt / 8
        // Step j = warp + 8 i covers k group j & 3 = warp & 3 and columns 8 (j >> 2) + nn: two steps 16 apart are 32
        // columns = one W box and 32 B rows apart, so every swizzle term below is fixed per thread.
        uint32_t src0[2], src1[2], dst[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int j = warp + i * kSplitWarps;
            const int kg = j & 3, n = (j >> 2) * 8 + nn;  // n < 32
            const int k = kg * 8 + 2 * kp;
            const uint32_t col = static_cast<uint32_t>(n & 3) * 4u, chunk = static_cast<uint32_t>(n >> 2);
            src0[i] = col + k * 128 + ((chunk ^ static_cast<uint32_t>(k & 7)) << 4);
            src1[i] = col + (k + 1) * 128 + ((chunk ^ static_cast<uint32_t>((k + 1) & 7)) << 4);
            dst[i] = static_cast<uint32_t>(n) * 64u + ((static_cast<uint32_t>(kg) ^ static_cast<uint32_t>((n >> 1) & 3)) << 4) +
                     static_cast<uint32_t>(kp) * 4u;
        }
        const int rounds = steps >> 4;  // nt / 32
        uint32_t ws = 0, wph = 0, bs = 0, bph = 0;
        for (long long item = blockIdx.x; item < items && ok; item += stride) {
            int u, it, hg, hn;
            decode(item, &u, &it, &hg, &hn);
            for (int s = 0; s < kb_count; ++s) {
                if (prm.direct)
                    ok = mbar_wait<false>(bar(kBEmpty + bs), bph ^ 1u, ctl, prm.status, kRoleSplit, kBEmpty + bs);
                else
                    ok = mbar_wait2<false>(bar(kWFull + ws), wph, kWFull + ws, bar(kBEmpty + bs), bph ^ 1u, kBEmpty + bs, ctl,
                                           prm.status, kRoleSplit);
                if (!ok) break;
                uint32_t wst = w_base + ws * kWStageBytes, bsl = b_base + bs * kBSlotBytes;
                for (int r = 0; r < rounds; ++r, wst += kWBoxBytes, bsl += 32 * 64) {
                    float w0[2], w1[2];
                    if (prm.direct) {
                        // element (k, n) of the unit's [K2][N2] coefficients; rows / columns past the end are zero like the
                        // hardware's fill.  A quarter-warp reads 8 consecutive columns of one row: one 32-byte sector.
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int j = warp + i * kSplitWarps;
                            const int k = s * kKb + (j & 3) * 8 + 2 * kp, n = it * nt + r * 32 + (j >> 2) * 8 + nn;
                            const float* src = prm.w + (static_cast<size_t>(u) * prm.K2 + k) * prm.N2 + n;
                            w0[i] = (k < prm.K2 && n < prm.N2) ? __ldg(src) : 0.f;
                            w1[i] = (k + 1 < prm.K2 && n < prm.N2) ? __ldg(src + prm.N2) : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            w0[i] = __uint_as_float(ld_shared_u32(wst + src0[i]));
                            w1[i] = __uint_as_float(ld_shared_u32(wst + src1[i]));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        // w = hi + mid + lo, each term a bfloat16; the residuals are exact in float32
                        const uint32_t p_hi = pack_bf16x2(w0[i], w1[i]);
                        const float r0 = w0[i] - __uint_as_float(p_hi << 16), r1 = w1[i] - __uint_as_float(p_hi & 0xffff0000u);
                        const uint32_t p_mid = pack_bf16x2(r0, r1);
                        const float s0 = r0 - __uint_as_float(p_mid << 16), s1 = r1 - __uint_as_float(p_mid & 0xffff0000u);
                        const uint32_t p_lo = pack_bf16x2(s0, s1);
                        st_shared_u32(bsl + dst[i], p_hi);
                        st_shared_u32(bsl + dst[i] + part_bytes, p_mid);
                        st_shared_u32(bsl + dst[i] + 2u * part_bytes, p_lo);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar(kBFull + bs));
                    if (!prm.direct) mbar_arrive(bar(kWEmpty + ws));
                }
                if (++ws == kWStages) ws = 0, wph ^= 1u;
                if (++bs == kBSlots) bs = 0, bph ^= 1u;
            }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

bool beamform_tc_supported(const void* reordered, const void* coeffs, const void* beams, int A, int M) {
    (void)A;  // any antenna count: rows that TMA cannot address one by one are fetched eight at a time
    (void)M;  // any beam count: odd ones read their coefficients and write their beams without tensor maps
    return aligned16(reordered) && aligned16(coeffs) && aligned16(beams);
}

int launch_beamform_tc(const uint8_t* reordered, const float* coeffs, float* beams, int B, int C, int T, int A, int M,
                       unsigned flags, cudaStream_t s) {
    const long long units = static_cast<long long>(B) * kPols * C;
    if (units > 0x7fffffffLL) return DCBF_ERR_UNSUPPORTED;
    TcParams p{};
    p.T = T, p.K2 = 2 * A, p.N2 = 2 * M;
    p.kb_count = (p.K2 + kKb - 1) / kKb;
    const int n_pad = ((p.N2 + 31) / 32) * 32;
    p.nt_count = (n_pad + kNtMax - 1) / kNtMax;
    p.nt = ((((n_pad + p.nt_count - 1) / p.nt_count) + 31) / 32) * 32;
    p.ht_count = (T + kTileT - 1) / kTileT;
    p.hg_count = (p.ht_count + kGroupTiles - 1) / kGroupTiles;
    p.items = units * p.nt_count * p.hg_count;
    p.signed_in = (flags & DCBF_FLAG_SIGNED_INPUT) ? 1 : 0;
    p.mma_warps = p.nt <= 64 ? 2 : 1;
    p.x_split = (p.K2 % 16) != 0;
    p.direct = (M % 2) != 0;
    p.w = coeffs;
    p.out = beams;
    if (int e = get_status_block(&p.status)) return e;

    EncodeTiledFn encode = nullptr;
    if (int e = get_encode_fn(&encode)) return e;
    alignas(64) CUtensorMap tm_x, tm_w, tm_out;
    const cuuint32_t estr[3] = {1, 1, 1};
    if (p.x_split) {
        // 2A bytes per sample is not a multiple of 16, which a tensor map needs of every stride: view the voltages as
        // [units][T / 8][8 x 2A] (stride 16 A) and fetch a [128 samples][32 B] stage as 8 boxes [16 rows][32 B], box i
        // = samples 8 r + i at byte offset i * 2A + k rounded down to 16 (48-byte rows, no swizzle).
        const cuuint64_t dims[3] = {static_cast<cuuint64_t>(8 * p.K2), static_cast<cuuint64_t>(T / 8), static_cast<cuuint64_t>(units)};
        const cuuint64_t strides[2] = {static_cast<cuuint64_t>(8 * p.K2), static_cast<cuuint64_t>(T) * p.K2};
        const cuuint32_t box[3] = {kXSplitRowBytes, kTileT / 8, 1};
        const CUresult r = encode(&tm_x, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(reordered), dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return record_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(reordered, split)");
    } else {
        // voltages as bytes [units][T][2A]; box [1][128][32], 32B swizzle
        const cuuint64_t dims[3] = {static_cast<cuuint64_t>(p.K2), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(units)};
        const cuuint64_t strides[2] = {static_cast<cuuint64_t>(p.K2), static_cast<cuuint64_t>(T) * p.K2};
        const cuuint32_t box[3] = {kKb, kTileT, 1};
        const CUresult r = encode(&tm_x, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(reordered), dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return record_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(reordered)");
    }
    if (!p.direct) {
    {
        // coefficients as float32 [units][2A][2M]; box [1][32][32], 128B swizzle
        const cuuint64_t dims[3] = {static_cast<cuuint64_t>(p.N2), static_cast<cuuint64_t>(p.K2), static_cast<cuuint64_t>(units)};
        const cuuint64_t strides[2] = {static_cast<cuuint64_t>(p.N2) * 4, static_cast<cuuint64_t>(p.K2) * p.N2 * 4};
        const cuuint32_t box[3] = {32, kKb, 1};
        const CUresult r = encode(&tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(coeffs), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return record_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(coeffs)");
    }
    {
        // beams as float32 [units][T][2M]; box [1][32][32], 128B swizzle
        const cuuint64_t dims[3] = {static_cast<cuuint64_t>(p.N2), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(units)};
        const cuuint64_t strides[2] = {static_cast<cuuint64_t>(p.N2) * 4, static_cast<cuuint64_t>(T) * p.N2 * 4};
        const cuuint32_t box[3] = {32, 32, 1};
        const CUresult r = encode(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, beams, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return record_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(beams)");
    }

    } else {
        tm_w = tm_x;  // never dereferenced in direct mode (odd beam counts: rows are only 8-byte aligned)
        tm_out = tm_x;
    }

    static int n_sms[64] = {};
    int dev = 0;
    DCBF_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DCBF_ERR_UNSUPPORTED;
    if (!n_sms[dev]) {
        DCBF_CUDA_TRY(cudaFuncSetAttribute(beamform_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        DCBF_CUDA_TRY(cudaDeviceGetAttribute(&n_sms[dev], cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = p.items < n_sms[dev] ? static_cast<int>(p.items) : n_sms[dev];
    beamform_tc_kernel<<<grid, kThreads, kSmemBytes, s>>>(p, tm_x, tm_w, tm_out);
    DCBF_CHECK_LAUNCH("beamform_tc_kernel");
    return DCBF_OK;
}

}  // namespace dcbf
