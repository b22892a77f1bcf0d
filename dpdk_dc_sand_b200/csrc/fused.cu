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
//   X[t][2a+x]   = f16(sample[b][a][c][t][p][x])          exact: |byte| <= 255
//   W[2a+x][2m+y] = {{cos, sin}, {-sin, cos}}[x][y] of rot(c, m, a), carried as fp16 hi + fp16 lo
//                  (two accumulating MMAs; coefficient error ~2^-24) or fp16 hi only (DCBF_FLAG_FP16_COEFF).
//
// One persistent CTA per SM walks channels c = blockIdx.x, +gridDim.x, ...  Warp roles (576 threads):
//   warps 0-7     coeffs    : delay_vals (coalesced float4) -> f64 phase -> sincospi -> swizzled B tiles
//   warps 8-11    convert   : raw bytes -> fp16, pol de-interleave, a<->t transpose into the 128B-swizzled A tiles
//   warps 12-15   epilogue  : tcgen05.ld 16x256b -> full-sector st.global.v2 straight from registers
//   warp 16       producer  : 1-D TMA bulk copies  in[b][a][c][t0:t0+128] (512 B runs) -> raw ring, mbarrier tx
//   warp 17       MMA       : one lane issues tcgen05.mma (M=128, N<=128, K=16), accumulators in TMEM
// Pipelines (mbarrier full/empty pairs): raw ring (TMA->convert), A ring (convert->MMA), B double buffer
// (coeffs->MMA, one channel ahead), TMEM accumulator double buffer (MMA->epilogue).
//
// Tiling: time tiles of 128 samples (UMMA M), k-blocks of 32 antennas (64 fp16 = one 128-byte swizzle row),
// N tiles of <=128 columns (64 beams) chosen so that one B tile set (all k-blocks, hi+lo) fits 64 KiB.  With
// more than one N tile the voltages of a channel are re-read (they then come from L2).
#include <cuda_fp16.h>

#include "common.cuh"

namespace dcbf {

namespace {

// Warp roles.  The SM sub-partition arbiter favours the highest warp id, so ids are handed out in the order
// of urgency: the MMA issuer and the TMA producer (a few instructions each, everything else waits on them)
// on top, then the epilogue (drains TMEM into the dominant HBM stream), convert, and the coefficient warps
// (longest job, but needed a whole channel later) at the bottom.
constexpr int kCoeffWarp0 = 0, kCoeffWarps = 8;  // warps 0..7
constexpr int kConvertWarp0 = 8;                 // warps 8..11
constexpr int kEpilogueWarp0 = 12;               // warps 12..15 (warp % 4 = TMEM lane quarter)
constexpr int kProducerWarp = 16;
constexpr int kMmaWarp = 17;
constexpr int kThreads = 18 * 32;
constexpr int kTileT = 128;    // samples per MMA tile (UMMA M)
constexpr int kKbAnts = 32;    // antennas per k-block
constexpr int kRawStages = 2;
constexpr int kAopStages = 2;
constexpr int kBopBufs = 2;
constexpr int kAccBufs = 2;
constexpr int kRawStageBytes = kKbAnts * kTileT * 4;  // 16 KiB: [ant][t][pol][re,im]
constexpr int kAopTileBytes = kTileT * 128;           // 16 KiB: [t][64 fp16], 128B swizzle
constexpr int kAopStageBytes = 2 * kAopTileBytes;     // pol 0 + pol 1
constexpr int kBopBufBytes = 64 * 1024;
constexpr int kTmemCols = 512;
constexpr unsigned long long kWatchdogNs = 2000000000ull;  // 2 s without progress on one barrier = dead-lock

constexpr int kSmemData = kAopStages * kAopStageBytes + kBopBufs * kBopBufBytes + kRawStages * kRawStageBytes;
constexpr int kSmemBytes = 1024 /*alignment slack*/ + kSmemData + 512 /*barriers + control*/;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

enum Role { kRoleProducer = 1, kRoleMma = 2, kRoleEpilogue = 3, kRoleConvert = 4, kRoleCoeff = 5 };

struct FusedParams {
    const uint8_t* in;
    const float4* dv;
    float* out;
    int* status;  // [0]=error code, [1]=role, [2]=barrier id, [3]=blockIdx
    unsigned long long* prof;  // optional [grid][6 roles][4]: ns blocked per barrier class, [..][3] = role span
    int B, A, C, T, M;
    int kb_count;   // ceil(A / 32)
    int nt;         // columns per N tile (multiple of 16, <= 128)
    int nt_count;   // number of N tiles
    int ht_count;   // ceil(T / 128)
    int parts;      // 2 = fp16 hi+lo coefficients, 1 = fp16 hi only
    int signed_in;
    int rowwise_epilogue;  // debug: 32x32b row-per-thread epilogue
    double chan_centre;    // absolute index of local channel 0, minus N/2
    double turns_per_delay;  // -1 / (N * Ts): half-turns of phase per (second of delay x channel offset)
};

// ------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 1-D TMA: global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 x fp16 -> fp32, single CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 lanes x 16 columns: r[0..1] = row lane/4, cols 2*(lane%4)+{0,1}; r[2..3] = row lane/4 + 8; r[4..7] = +8 columns.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// 16 lanes x 64 columns (8 repeats of the 8-column pattern above).
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 columns: thread = row, r[j] = column j.
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void st_global_v2(float* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// Shared-memory matrix descriptor: K-major, 128-byte swizzle, rows 128 B apart, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);
    d |= 1ull << 16;             // leading byte offset (unused for swizzled K-major), canonical value 1
    d |= (1024ull >> 4) << 32;   // stride byte offset between 8-row atoms
    d |= 1ull << 46;             // descriptor version (Blackwell)
    d |= 2ull << 61;             // SWIZZLE_128B
    return d;
}
// Instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, M = 128.
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(kTileT >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------
// Barrier wait with a dead-lock guard: on timeout the CTA aborts cooperatively (no hang, no trap) and the
// host sees DCBF_ERR_TIMEOUT through dcbf_fused_status().
// ------------------------------------------------------------------------------------------------------
struct Control {
    uint32_t tmem_base;
    volatile int abort;
    unsigned long long wait_ns[6][4];  // [role][slot]: time lane 0 of a role's first warp spent blocked
};

__device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, Control* ctl, int* status, int role, int id) {
    const unsigned long long t0 = global_ns();
    for (;;) {
        // up to 64 hardware-suspended probes in a 7-instruction loop, then one look at the abort flag / clock
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .u32 n;\n\t"
            "mov.u32 n, 0;\n"
            "DCBF_WAIT_AGAIN:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "@p bra DCBF_WAIT_DONE;\n\t"
            "add.u32 n, n, 1;\n\t"
            "setp.lt.u32 q, n, 64;\n\t"
            "@q bra DCBF_WAIT_AGAIN;\n"
            "DCBF_WAIT_DONE:\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(100000u)
            : "memory");
        if (ok) return true;
        if (ctl->abort) return false;
        if (global_ns() - t0 > kWatchdogNs) {
            ctl->abort = 1;
            if (atomicCAS(status, 0, DCBF_ERR_TIMEOUT) == 0) {
                status[1] = role;
                status[2] = id;
                status[3] = static_cast<int>(blockIdx.x);
            }
            return false;
        }
    }
}
// Warp-collective: every lane waits; the result is made warp-uniform.
// `slot` >= 0 on exactly one lane of a role makes that lane account its blocked time (profiling aid; the
// first try_wait may itself suspend the thread, so the whole call is timed).
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, Control* ctl, int* status, int role, int id,
                                          int slot = -1) {
    unsigned long long t0 = 0;
    if (slot >= 0) t0 = global_ns();
    bool ok = mbar_try_wait(bar, parity) != 0;
    if (!ok) ok = mbar_wait_slow(bar, parity, ctl, status, role, id);
    if (slot >= 0) ctl->wait_ns[role][slot] += global_ns() - t0;
    return __all_sync(0xffffffffu, ok);
}

// u8 (or i8) pair -> half2, exact.  `w` holds {p0.re, p0.im, p1.re, p1.im}; sel picks the pol.
__device__ __forceinline__ uint32_t bytes_to_half2(uint32_t w, uint32_t sel, uint32_t bias) {
    // bytes -> 0x64bb = 1024 + b (fp16), then subtract 1024 (u8) or 1152 (i8 after the ^0x80 re-bias)
    const uint32_t h = __byte_perm(w, 0x64646464u, sel);
    const __half2 r = __hsub2(*reinterpret_cast<const __half2*>(&h), *reinterpret_cast<const __half2*>(&bias));
    return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// sin(pi r), cos(pi r) for r in [-1, 1] (half-turns, already range-reduced in float64).  Quadrant split
// q = rint(2r), t = r - q/2 in [-1/4, 1/4], odd/even Taylor polynomials in t (truncation < 2e-9 and 2e-10),
// then the quadrant rotation.  Absolute error <= ~1.2e-7; no special cases (r is always finite here).
__device__ __forceinline__ void sincospi_reduced(float r, float* sn, float* cs) {
    const float z = fmaf(r, 2.0f, 12582912.0f);  // 1.5 * 2^23: the low mantissa bits now hold rint(2r)
    const int q = __float_as_int(z);
    const float t = fmaf(z - 12582912.0f, -0.5f, r);
    const float s = t * t;
    float ps = fmaf(s, 0.0821458866f, -0.599264529f);   // pi^9/9!, -pi^7/7!
    ps = fmaf(ps, s, 2.55016404f);                       // pi^5/5!
    ps = fmaf(ps, s, -5.16771278f);                      // -pi^3/3!
    ps = fmaf(ps * s, t, t * 3.14159274f) ;              // t*pi + t*s*(...)
    float pc = fmaf(s, -0.0258068914f, 0.235330630f);    // -pi^10/10!, pi^8/8!
    pc = fmaf(pc, s, -1.33526277f);                      // -pi^6/6!
    pc = fmaf(pc, s, 4.05871213f);                       // pi^4/4!
    pc = fmaf(pc, s, -4.93480220f);                      // -pi^2/2!
    pc = fmaf(pc, s, 1.0f);
    const bool swap = q & 1;
    const float a = swap ? pc : ps, b = swap ? ps : pc;
    // q mod 4: 0 -> (s, c); 1 -> (c, -s); 2 -> (-s, -c); 3 -> (-c, s)
    *sn = __int_as_float(__float_as_int(a) ^ ((q << 30) & 0x80000000));
    *cs = __int_as_float(__float_as_int(b) ^ (((q + 1) << 30) & 0x80000000));
}

// ------------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) fused_beamform_kernel(const FusedParams prm) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const uint32_t aop_base = smem_base;                                    // [stage][pol][128 x 128 B]
    const uint32_t bop_base = aop_base + kAopStages * kAopStageBytes;       // [buf][kb][part][nt x 128 B]
    const uint32_t raw_base = bop_base + kBopBufs * kBopBufBytes;           // [stage][ant][t][4 B]
    const uint32_t bar_base = raw_base + kRawStages * kRawStageBytes;       // 8-byte mbarriers
    Control* ctl = reinterpret_cast<Control*>(smem_gen + kSmemData + 192);

    // barrier ids (also reported by the watchdog)
    const int kRawFull = 0, kRawEmpty = kRawFull + kRawStages, kAopFull = kRawEmpty + kRawStages,
              kAopEmpty = kAopFull + kAopStages, kBopFull = kAopEmpty + kAopStages, kBopEmpty = kBopFull + kBopBufs,
              kAccFull = kBopEmpty + kBopBufs, kAccEmpty = kAccFull + kAccBufs, kNumBars = kAccEmpty + kAccBufs;
    static_assert(2 * (kRawStages + kAopStages + kBopBufs + kAccBufs) * 8 <= 192, "barrier area");
    auto bar = [&](int id) { return bar_base + 8u * static_cast<uint32_t>(id); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < kRawStages; ++s) {
            mbar_init(bar(kRawFull + s), 1);
            mbar_init(bar(kRawEmpty + s), 4);
        }
        for (int s = 0; s < kAopStages; ++s) {
            mbar_init(bar(kAopFull + s), 4);
            mbar_init(bar(kAopEmpty + s), 1);
        }
        for (int s = 0; s < kBopBufs; ++s) {
            mbar_init(bar(kBopFull + s), kCoeffWarps);
            mbar_init(bar(kBopEmpty + s), 1);
        }
        for (int s = 0; s < kAccBufs; ++s) {
            mbar_init(bar(kAccFull + s), 1);
            mbar_init(bar(kAccEmpty + s), 4);
        }
        ctl->abort = 0;
        for (int i = 0; i < 24; ++i) ctl->wait_ns[i >> 2][i & 3] = 0;
        fence_mbar_init();
        (void)kNumBars;
    }
    // B tiles start as zeros: padding rows (k >= 2A, n >= 2M) are never written afterwards.
    {
        uint4* z = reinterpret_cast<uint4*>(smem_gen + kAopStages * kAopStageBytes);
        for (int i = threadIdx.x; i < kBopBufs * kBopBufBytes / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(&ctl->tmem_base), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;

    const int A = prm.A, C = prm.C, T = prm.T, M = prm.M, B = prm.B;
    const int N2 = 2 * M;
    const int nt = prm.nt, parts = prm.parts;
    const uint32_t bop_kb_bytes = static_cast<uint32_t>(parts * nt * 128);  // one k-block: [part][nt rows][128 B]
    // profiling: lane 0 of each role's first warp accounts blocked time per barrier class (slot) and role span
    const bool prof_lane = lane == 0 && (warp == kProducerWarp || warp == kMmaWarp || warp == kEpilogueWarp0 || warp == kConvertWarp0 || warp == kCoeffWarp0);
    const int ps = prof_lane ? 0 : -100;
    const int my_role = warp == kProducerWarp ? kRoleProducer : warp == kMmaWarp ? kRoleMma : warp >= kEpilogueWarp0 ? kRoleEpilogue : warp >= kConvertWarp0 ? kRoleConvert : kRoleCoeff;
    const unsigned long long role_t0 = prof_lane ? global_ns() : 0ull;

    if (warp == kProducerWarp) {
        // =================================== TMA producer ===================================
        uint32_t slab = 0;
        bool ok = true;
        for (int c = blockIdx.x; c < C && ok; c += gridDim.x)
            for (int it = 0; it < prm.nt_count && ok; ++it)
                for (int b = 0; b < B && ok; ++b)
                    for (int h = 0; h < prm.ht_count && ok; ++h) {
                        const int t0 = h * kTileT;
                        const int rows = min(kTileT, T - t0);
                        for (int kb = 0; kb < prm.kb_count; ++kb, ++slab) {
                            const uint32_t rs = slab % kRawStages, ph = (slab / kRawStages) & 1u;
                            ok = mbar_wait(bar(kRawEmpty + rs), ph ^ 1u, ctl, prm.status, kRoleProducer, kRawEmpty + rs, ps + 0);
                            if (!ok) break;
                            const int a0 = kb * kKbAnts;
                            const int n_ants = min(kKbAnts, A - a0);
                            if (lane == 0)
                                mbar_arrive_expect_tx(bar(kRawFull + rs), static_cast<uint32_t>(n_ants * rows * 4));
                            __syncwarp();
                            if (lane < n_ants) {
                                const size_t row = ((static_cast<size_t>(b) * A + (a0 + lane)) * C + c) * static_cast<size_t>(T) + t0;
                                bulk_g2s(raw_base + rs * kRawStageBytes + lane * (kTileT * 4), prm.in + row * 4,
                                         static_cast<uint32_t>(rows * 4), bar(kRawFull + rs));
                            }
                        }
                    }
    } else if (warp == kMmaWarp) {
        // =================================== MMA issuer ===================================
        const uint32_t idesc = make_idesc_f16(nt);
        uint32_t slab = 0, unit = 0, step = 0;
        bool ok = true;
        for (int c = blockIdx.x; c < C && ok; c += gridDim.x)
            for (int it = 0; it < prm.nt_count && ok; ++it, ++step) {
                const uint32_t bb = step % kBopBufs;
                ok = mbar_wait(bar(kBopFull + bb), (step / kBopBufs) & 1u, ctl, prm.status, kRoleMma, kBopFull + bb, ps + 0);
                for (int bh = 0; bh < B * prm.ht_count && ok; ++bh, ++unit) {
                    const uint32_t ab = unit % kAccBufs;
                    ok = mbar_wait(bar(kAccEmpty + ab), ((unit / kAccBufs) & 1u) ^ 1u, ctl, prm.status, kRoleMma, kAccEmpty + ab, ps + 1);
                    if (!ok) break;
                    tc_fence_after();
                    for (int kb = 0; kb < prm.kb_count; ++kb, ++slab) {
                        const uint32_t as = slab % kAopStages;
                        ok = mbar_wait(bar(kAopFull + as), (slab / kAopStages) & 1u, ctl, prm.status, kRoleMma, kAopFull + as, ps + 2);
                        if (!ok) break;
                        tc_fence_after();
                        if (lane == 0) {
                            const int n_ants = min(kKbAnts, A - kb * kKbAnts);
                            const int k16_steps = (n_ants + 7) >> 3;  // 8 antennas = 16 k per MMA
                            for (int p = 0; p < kPols; ++p) {
                                const uint32_t d_tmem = tmem_base + (ab * kPols + p) * static_cast<uint32_t>(nt);
                                const uint64_t a_desc = make_kmajor_sw128_desc(aop_base + as * kAopStageBytes + p * kAopTileBytes);
                                for (int part = 0; part < parts; ++part) {
                                    const uint64_t b_desc = make_kmajor_sw128_desc(
                                        bop_base + bb * kBopBufBytes + kb * bop_kb_bytes + part * (nt * 128));
                                    for (int k = 0; k < k16_steps; ++k) {
                                        // +32 B per K=16 step inside the 128-byte swizzle row (encoded >> 4)
                                        umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | part | k) != 0);
                                    }
                                }
                            }
                            umma_commit(bar(kAopEmpty + as));  // A stage free once these MMAs retire
                        }
                        __syncwarp();
                    }
                    if (ok && lane == 0) umma_commit(bar(kAccFull + ab));
                    __syncwarp();
                }
                if (ok && lane == 0) umma_commit(bar(kBopEmpty + bb));
                __syncwarp();
            }
    } else if (warp >= kEpilogueWarp0 && warp < kEpilogueWarp0 + 4) {
        // =================================== epilogue ===================================
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        uint32_t unit = 0;
        bool ok = true;
        for (int c = blockIdx.x; c < C && ok; c += gridDim.x)
            for (int it = 0; it < prm.nt_count && ok; ++it) {
                const int n0 = it * nt;
                for (int b = 0; b < B && ok; ++b)
                    for (int h = 0; h < prm.ht_count && ok; ++h, ++unit) {
                        const uint32_t ab = unit % kAccBufs;
                        ok = mbar_wait(bar(kAccFull + ab), (unit / kAccBufs) & 1u, ctl, prm.status, kRoleEpilogue, kAccFull + ab, ps + 0);
                        if (!ok) break;
                        tc_fence_after();
                        const int t0 = h * kTileT;
                        for (int p = 0; p < kPols; ++p) {
                            const uint32_t col0 = (ab * kPols + p) * static_cast<uint32_t>(nt);
                            float* tile_out = prm.out + (((static_cast<size_t>(b) * kPols + p) * C + c) * static_cast<size_t>(T) + t0) * N2 + n0;
                            if (!prm.rowwise_epilogue) {
#pragma unroll 1
                                for (int half = 0; half < 2; ++half) {
                                    const int r_lo = 32 * q + 16 * half + (lane >> 2);
                                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q + 16 * half) << 16) + col0;
                                    float* row_lo = tile_out + static_cast<size_t>(r_lo) * N2 + 2 * (lane & 3);
                                    float* row_hi = row_lo + 8 * static_cast<size_t>(N2);
                                    const bool v_lo = t0 + r_lo < T, v_hi = t0 + r_lo + 8 < T;
                                    int cb = 0;
                                    for (; cb + 64 <= nt; cb += 64) {
                                        uint32_t r[32];
                                        tmem_ld_16x256b_x8(taddr + cb, r);
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
                                    for (; cb < nt; cb += 16) {
                                        uint32_t r[8];
                                        tmem_ld_16x256b_x2(taddr + cb, r);
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
                            } else {
                                const int row = 32 * q + lane;
                                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + col0;
                                float* rowp = tile_out + static_cast<size_t>(row) * N2;
                                for (int cb = 0; cb < nt; cb += 16) {
                                    uint32_t r[16];
                                    tmem_ld_32x32b_x16(taddr + cb, r);
                                    tmem_wait_ld();
                                    if (t0 + row < T) {
#pragma unroll
                                        for (int j = 0; j < 16; j += 2)
                                            if (n0 + cb + j < N2) st_global_v2(rowp + cb + j, r[j], r[j + 1]);
                                    }
                                }
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(kAccEmpty + ab));
                    }
            }
    } else if (warp >= kConvertWarp0 && warp < kConvertWarp0 + 4) {
        // =================================== convert ===================================
        // thread = one sample row t; per 4-antenna chunk: 4 conflict-free LDS.32, 8 PRMT+HSUB2, 2 STS.128
        const int t = threadIdx.x - kConvertWarp0 * 32;
        const uint32_t bias = prm.signed_in ? 0x64806480u : 0x64006400u;  // 1152 | 1024 as fp16 pairs
        const uint32_t flip = prm.signed_in ? 0x80808080u : 0u;
        uint32_t slab = 0;
        bool ok = true;
        for (int c = blockIdx.x; c < C && ok; c += gridDim.x)
            for (int it = 0; it < prm.nt_count && ok; ++it)
                for (int bh = 0; bh < B * prm.ht_count && ok; ++bh)
                    for (int kb = 0; kb < prm.kb_count; ++kb, ++slab) {
                        const uint32_t rs = slab % kRawStages, as = slab % kAopStages;
                        ok = mbar_wait(bar(kRawFull + rs), (slab / kRawStages) & 1u, ctl, prm.status, kRoleConvert, kRawFull + rs, ps + 0);
                        if (ok)
                            ok = mbar_wait(bar(kAopEmpty + as), ((slab / kAopStages) & 1u) ^ 1u, ctl, prm.status, kRoleConvert, kAopEmpty + as, ps + 1);
                        if (!ok) break;
                        const int n_ants = min(kKbAnts, A - kb * kKbAnts);
                        const int n_chunks = 2 * ((n_ants + 7) >> 3);  // 4-antenna chunks inside the padded K extent
                        const uint32_t src = raw_base + rs * kRawStageBytes + t * 4;
                        const uint32_t dst0 = aop_base + as * kAopStageBytes + t * 128;
                        for (int j = 0; j < n_chunks; ++j) {
                            uint32_t w[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int a = 4 * j + i;
                                w[i] = (a < n_ants) ? (ld_shared_u32(src + a * (kTileT * 4)) ^ flip) : flip;
                            }
                            const uint32_t off = static_cast<uint32_t>((j ^ (t & 7)) << 4);
                            // padding antennas: byte 0 (u8) / 0x80^0x80 -> value 0 after the bias subtraction
                            st_shared_v4(dst0 + off, bytes_to_half2(w[0], 0x4140u, bias), bytes_to_half2(w[1], 0x4140u, bias),
                                         bytes_to_half2(w[2], 0x4140u, bias), bytes_to_half2(w[3], 0x4140u, bias));
                            st_shared_v4(dst0 + kAopTileBytes + off, bytes_to_half2(w[0], 0x4342u, bias),
                                         bytes_to_half2(w[1], 0x4342u, bias), bytes_to_half2(w[2], 0x4342u, bias),
                                         bytes_to_half2(w[3], 0x4342u, bias));
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(bar(kAopFull + as));
                            mbar_arrive(bar(kRawEmpty + rs));
                        }
                    }
    } else if (warp < kCoeffWarp0 + kCoeffWarps) {
        // =================================== steering coefficients ===================================
        // delay_vals[c][m0 .. m0+mt) is one contiguous run of (beam, antenna) entries: the 256 threads walk it
        // with lane <-> consecutive entry, so every warp load is 512 contiguous bytes.  Each entry becomes four
        // 32-bit words (row 2m | row 2m+1) x (fp16 hi | fp16 lo residual); consecutive antennas are consecutive
        // words of one 128-byte B row, so each of the four STS.32 of a warp touches 32 different banks.
        const int ctid = threadIdx.x - kCoeffWarp0 * 32;
        const int mt = nt >> 1;  // beams per N tile
        constexpr int kBatch = 8;
        constexpr int kStride = kCoeffWarps * 32;
        const int dm = kStride / A, da = kStride - dm * A;  // (beam, antenna) advance per kStride entries
        const int ml_first = ctid / A, a_first = ctid - ml_first * A;
        const uint32_t part_bytes = static_cast<uint32_t>(nt * 128);
        const double kInvPi = 0.318309886183790671538;
        uint32_t step = 0;
        bool ok = true;
        for (int c = blockIdx.x; c < C && ok; c += gridDim.x) {
            const double chan = static_cast<double>(c) + prm.chan_centre;
            const double scale = chan * prm.turns_per_delay;  // half-turns per second of delay
            for (int it = 0; it < prm.nt_count && ok; ++it, ++step) {
                const uint32_t bb = step % kBopBufs;
                const int m0 = it * mt;
                const int entries = min(mt, M - m0) * A;
                const float4* src = prm.dv + (static_cast<size_t>(c) * M + m0) * A;
                // warm L2 for the step after this one (same channel next N tile, or next channel's first)
                if (warp == kCoeffWarp0 && lane == 0) {
                    int nc = c, nit = it + 1;
                    if (nit == prm.nt_count) {
                        nit = 0;
                        nc = c + gridDim.x;
                    }
                    if (nc < C) {
                        const int nm0 = nit * mt;
                        const int nm = min(mt, M - nm0);
                        const size_t bytes = static_cast<size_t>(nm) * A * 16;
                        const char* p = reinterpret_cast<const char*>(prm.dv + (static_cast<size_t>(nc) * M + nm0) * A);
                        for (size_t o = 0; o < bytes; o += 65536)
                            bulk_prefetch_l2(p + o, static_cast<uint32_t>(min(bytes - o, static_cast<size_t>(65536))));
                    }
                }
                bool waited = false;
                const uint32_t buf = bop_base + bb * kBopBufBytes;
                int ml = ml_first, a = a_first;
                for (int e0 = ctid; e0 < entries + ctid; e0 += kStride * kBatch) {  // e0 - ctid is warp-uniform
                    float2 v[kBatch];  // (delay_s, phase_rad); the two rate fields are ignored like the reference does
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int e = e0 + u * kStride;
                        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (e < entries) t4 = ldg_nc_f4(src + e);
                        v[u] = make_float2(t4.x, t4.z);
                    }
                    if (!waited) {  // the loads above are already in flight while we wait for the buffer
                        ok = mbar_wait(bar(kBopEmpty + bb), ((step / kBopBufs) & 1u) ^ 1u, ctl, prm.status, kRoleCoeff, kBopEmpty + bb, ps + 0);
                        waited = true;
                        if (!ok) break;
                    }
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int e = e0 + u * kStride;
                        if (e < entries) {
                            // rot/pi = delay * (ch - N/2) * (-1/(N Ts)) + phase/pi   (coeff_generator_cpu.py:143-165)
                            const double x = fma(static_cast<double>(v[u].x), scale, static_cast<double>(v[u].y) * kInvPi);
                            const float r = static_cast<float>(x - 2.0 * rint(0.5 * x));  // [-1, 1] half-turns
                            float sn, cs;
                            sincospi_reduced(r, &sn, &cs);
                            // fp16 hi + fp16 residual of (cos, sin)
                            const uint32_t hi = pack_half2(cs, sn);
                            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
                            const uint32_t lo = pack_half2(cs - hf.x, sn - hf.y);
                            // B^T rows: n = 2m -> (k=2a: cos, k=2a+1: -sin);  n = 2m+1 -> (sin, cos)
                            const int row = 2 * ml;
                            const int al = a & (kKbAnts - 1);
                            const uint32_t d0 = buf + static_cast<uint32_t>(a >> 5) * bop_kb_bytes + static_cast<uint32_t>(row) * 128u +
                                                (static_cast<uint32_t>(((al >> 2) ^ (row & 7)) << 4) | static_cast<uint32_t>((al & 3) << 2));
                            const uint32_t d1 = (d0 + 128u) ^ 16u;  // row + 1: swizzle phase (row & 7) | 1
                            st_shared_u32(d0, hi ^ 0x80000000u);
                            st_shared_u32(d1, __byte_perm(hi, 0u, 0x1032u));
                            if (parts > 1) {
                                st_shared_u32(d0 + part_bytes, lo ^ 0x80000000u);
                                st_shared_u32(d1 + part_bytes, __byte_perm(lo, 0u, 0x1032u));
                            }
                        }
                        ml += dm;
                        a += da;
                        if (a >= A) {
                            a -= A;
                            ++ml;
                        }
                    }
                }
                if (!ok) break;
                if (!waited)  // this warp had no entries in this step; it still takes part in the hand-shake
                    ok = mbar_wait(bar(kBopEmpty + bb), ((step / kBopBufs) & 1u) ^ 1u, ctl, prm.status, kRoleCoeff, kBopEmpty + bb, ps + 0);
                if (!ok) break;
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(kBopFull + bb));
            }
        }
    }

    // ---- teardown ----
    if (prof_lane) ctl->wait_ns[my_role][3] = global_ns() - role_t0;
    tc_fence_before();
    __syncthreads();
    if (prm.prof && threadIdx.x < 24) prm.prof[blockIdx.x * 24 + threadIdx.x] = ctl->wait_ns[threadIdx.x >> 2][threadIdx.x & 3];
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

int* g_status_dev[64] = {};  // per-device 4-int status block, allocated on first use
unsigned long long* g_prof_dev = nullptr;  // set by fused_set_profile_buffer (developer aid)

}  // namespace

// Picks the N tiling: nt columns per tile (multiple of 16, <= 128) such that kb_count * parts * nt * 128 B <= 64 KiB.
static void pick_n_tiling(int A, int M, int parts, int* kb_count, int* nt, int* nt_count) {
    const int kbc = (A + kKbAnts - 1) / kKbAnts;
    const int n_pad = ((2 * M + 15) / 16) * 16;
    int nt_max = (kBopBufBytes / (kbc * parts * 128)) & ~15;
    if (nt_max > 128) nt_max = 128;
    const int count = nt_max > 0 ? (n_pad + nt_max - 1) / nt_max : 0;
    *kb_count = kbc;
    *nt_count = count;
    *nt = count > 0 ? ((((n_pad + count - 1) / count) + 15) / 16) * 16 : 0;
}

static int get_status_block(int** out) {
    int dev = 0;
    DCBF_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DCBF_ERR_UNSUPPORTED;
    if (!g_status_dev[dev]) {
        int* p = nullptr;
        DCBF_CUDA_TRY(cudaMalloc(&p, 4 * sizeof(int)));
        DCBF_CUDA_TRY(cudaMemset(p, 0, 4 * sizeof(int)));
        g_status_dev[dev] = p;
    }
    *out = g_status_dev[dev];
    return DCBF_OK;
}

int launch_fused(const uint8_t* samples, const float* delay_vals, float* beams, int B, int A, int C, int N, int T,
                 int M, long long first_chan, double sample_period, unsigned flags, cudaStream_t s) {
    FusedParams p{};
    p.in = samples;
    p.dv = reinterpret_cast<const float4*>(delay_vals);
    p.out = beams;
    p.B = B, p.A = A, p.C = C, p.T = T, p.M = M;
    p.parts = (flags & DCBF_FLAG_FP16_COEFF) ? 1 : 2;
    p.signed_in = (flags & DCBF_FLAG_SIGNED_INPUT) ? 1 : 0;
    p.rowwise_epilogue = (flags & DCBF_FLAG_DEBUG_ROWWISE_EPILOGUE) ? 1 : 0;
    pick_n_tiling(A, M, p.parts, &p.kb_count, &p.nt, &p.nt_count);
    if (p.nt < 16) return DCBF_ERR_UNSUPPORTED;  // more than 128 k-blocks (4096 antennas)
    p.ht_count = (T + kTileT - 1) / kTileT;
    p.chan_centre = static_cast<double>(first_chan) - static_cast<double>(N) / 2.0;
    p.turns_per_delay = -1.0 / (static_cast<double>(N) * sample_period);
    if (int e = get_status_block(&p.status)) return e;
    p.prof = g_prof_dev;

    static int n_sms[64] = {};
    int dev = 0;
    DCBF_CUDA_TRY(cudaGetDevice(&dev));
    if (!n_sms[dev]) {
        DCBF_CUDA_TRY(cudaDeviceGetAttribute(&n_sms[dev], cudaDevAttrMultiProcessorCount, dev));
        DCBF_CUDA_TRY(cudaFuncSetAttribute(fused_beamform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    }
    const int grid = C < n_sms[dev] ? C : n_sms[dev];
    fused_beamform_kernel<<<grid, kThreads, kSmemBytes, s>>>(p);
    DCBF_CHECK_LAUNCH("fused_beamform_kernel");
    return DCBF_OK;
}

int fused_status(int* role, int* barrier, int* block) {
    int* blk = nullptr;
    if (int e = get_status_block(&blk)) return e;
    int h[4] = {};
    DCBF_CUDA_TRY(cudaMemcpy(h, blk, sizeof(h), cudaMemcpyDeviceToHost));  // synchronises with prior work
    if (h[0] != 0) DCBF_CUDA_TRY(cudaMemset(blk, 0, sizeof(h)));
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
