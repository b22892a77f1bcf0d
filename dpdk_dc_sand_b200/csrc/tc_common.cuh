// PTX wrappers, UMMA descriptors and the watchdog-guarded mbarrier waits shared by the tcgen05 kernels of libdcbf
// (fused.cu, beamform_tc.cu).  sm_100a only.  Everything here is internal linkage.
#pragma once

#include <cuda.h>
#include <stdint.h>

#include "common.cuh"

namespace dcbf {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int get_encode_fn(EncodeTiledFn* out);  // fused.cu
// Per-device status block shared by the tcgen05 kernels: [0] = error code, [1] = role, [2] = barrier id, [3] = CTA
// (read and cleared by dcbf_fused_status), [4..5] = device-visible address of a page-locked host flag raised with the
// error code; the fused kernel's channel counters follow at [8].
int get_status_block(int** out);        // fused.cu

namespace {

constexpr unsigned long long kWatchdogNs = 2000000000ull;  // 2 s without progress on one barrier = dead-lock

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
// kCluster: acquire at cluster scope (the barrier also collects arrivals from the other CTA of a pair).
template <bool kCluster = false>
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    if (kCluster)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    return ok;
}
// True on exactly one lane of a converged warp.  ptxas knows the guarded region is single-threaded, so
// warp-uniform operands of tcgen05 / TMA instructions go straight to uniform registers.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
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

// Tensor-map TMA: one [16 ant] x [128 sample] box of 4-byte words, global -> shared, bytes counted on an mbarrier.
// Out-of-range antennas / samples are zero-filled by the hardware and still count towards the box's bytes.
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// Tensor-map TMA store: one [32 row] x [32 column] fp32 box, shared -> global; rows/columns outside the tensor are clipped.
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// Plain (non-tensor) bulk copies: contiguous bytes, 16-byte aligned, size a multiple of 16.
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_group_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
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
// ---- CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of a TPC share one MMA; CTA rank 0 issues ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.  Default semantics (release at
// CTA scope), as in every 2-SM pipeline of CUTLASS: what the arrival publishes are this CTA's own shared-memory writes,
// already made visible to the async proxy by fence.proxy.async, which the pair's tensor cores read in place.  (The
// .release.cluster form costs a GPU-scope MEMBAR + ERRBAR per arrival: 10 % of all stall samples in the first build.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 rem;\n\t"
        "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t}" ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {  // same warp id in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem, 128 rows in each CTA] (+)= A[smem, each CTA its 128 rows] * B[smem, each CTA half of the N rows]
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// Arrive on the mbarrier at this offset in BOTH CTAs of the pair once every MMA issued so far has completed.
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
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
// 32 lanes x 32 columns: thread = row (TMEM lane), r[j] = column j.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// 32 lanes x 64 columns: thread = row (TMEM lane), r[j] = column j.  One load (one TMEM round trip, several hundred
// cycles) for twice the columns of the x32 form.
__device__ __forceinline__ void tmem_ld_32x32b_x64(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
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
// (the offset becomes the instruction's immediate)
template <int kOff>
__device__ __forceinline__ void st_shared_u32_off(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(kOff) : "memory");
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t (&v)[4]) {
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr) : "memory");
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

// Shared-memory matrix descriptors, K-major.  Low word: start address >> 4 (14 bits) | leading-byte-offset field
// (unused for swizzled K-major, canonical value 1).  High word: stride between 8-row atoms >> 4, descriptor
// version 1 (Blackwell), swizzle mode.  A K=16 step advances the start address by 32 B (+2 in the low word).
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3fffu) | (1u << 16); }
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // B tiles: 128 B rows, 128B swizzle
constexpr uint32_t kDescHiSw64 = (512u >> 4) | (1u << 14) | (4u << 29);    // A tiles: 64 B rows, 64B swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) {
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
// Instruction descriptor for kind::f16: fp32 accumulate, both operands K-major, M = 128, N = n.
// a_bf16 / b_bf16 select bfloat16 instead of fp16 for that operand (the two formats are independent fields).
// m = 128, or 256 for a cta_group::2 instruction (128 rows in each CTA of the pair).
__device__ __forceinline__ uint32_t make_idesc_f16(int n, bool a_bf16 = false, bool b_bf16 = false, int m = 128) {
    return (1u << 4) | (a_bf16 ? 1u << 7 : 0u) | (b_bf16 ? 1u << 10 : 0u) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------
// Barrier wait with a dead-lock guard: on timeout the CTA aborts cooperatively (no hang, no trap) and the
// host sees DCBF_ERR_TIMEOUT through dcbf_fused_status().
// ------------------------------------------------------------------------------------------------------
struct Control {
    uint32_t tmem_base;
    volatile int abort;
    unsigned long long wait_ns[6][4];  // [role][slot]: time lane 0 of a role's first warp spent blocked
    volatile int chan_pub;             // number of entries of this CTA's channel sequence published so far
    volatile int chan_ring[8];         // channel sequence, entry k at [k % 8]; kChanSentinel ends it
    float q8_gmax;                     // fused q8 output: max |gain| over the beams
};
constexpr int kChanSentinel = 0x7fffffff;

// k-th channel this CTA works on.  Channels are handed out dynamically (first one = blockIdx.x, the rest from
// a global counter) by one lane of the coefficient role, which is the first to need them; everybody else
// reads the sequence from shared memory.  Balances the finish times of the 148 persistent CTAs to within one
// channel instead of the fixed 27-or-28 split, and absorbs per-SM speed differences.
__device__ __forceinline__ int sched_get(Control* ctl, uint32_t k) {
    while (ctl->chan_pub <= static_cast<int>(k)) {
        if (ctl->abort) return kChanSentinel;
        __nanosleep(32);
    }
    return ctl->chan_ring[k & 7];
}

// kSleepNs > 0: sleep between failed probes.  A suspended try_wait is woken by every arrival on ANY barrier of the CTA, so a
// role that waits long (the coefficient warps for a free B buffer, the producer for a free raw stage) otherwise spins
// through this loop at full rate: at C3 23 % of all executed warp instructions were the sixteen coefficient warps doing
// exactly that -- issue slots and power taken from the roles that have work.  Only for waits whose wake-up latency is
// off the critical path (the waiter is a whole buffer ahead).
template <int kSleepNs = 0>
__device__ __forceinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, Control* ctl, int* status, int role, int id) {
    const unsigned long long t0 = global_ns();
    for (;;) {
        // up to 64 hardware-suspended probes in a 7-instruction loop, then one look at the abort flag / clock
        uint32_t ok;
        if constexpr (kSleepNs > 0) {
            asm volatile(
                "{\n\t.reg .pred p, q;\n\t.reg .u32 n;\n\t"
                "mov.u32 n, 0;\n"
                "DCBF_WAITS_AGAIN:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "@p bra DCBF_WAITS_DONE;\n\t"
                "nanosleep.u32 %4;\n\t"
                "add.u32 n, n, 1;\n\t"
                "setp.lt.u32 q, n, 64;\n\t"
                "@q bra DCBF_WAITS_AGAIN;\n"
                "DCBF_WAITS_DONE:\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(bar), "r"(parity), "r"(100000u), "r"(static_cast<uint32_t>(kSleepNs))
                : "memory");
        } else {
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
        }
        if (ok) return true;
        if (ctl->abort) return false;
        if (global_ns() - t0 > kWatchdogNs) {
            ctl->abort = 1;
            if (atomicCAS(status, 0, DCBF_ERR_TIMEOUT) == 0) {
                status[1] = role;
                status[2] = id;
                status[3] = static_cast<int>(blockIdx.x);
                // page-locked host flag (address kept in the status block): lets the host notice without a copy
                volatile int* host_flag = *reinterpret_cast<volatile int* volatile*>(status + 4);
                if (host_flag) *host_flag = DCBF_ERR_TIMEOUT;
                __threadfence_system();
            }
            return false;
        }
    }
}
// Warp-collective: every lane waits; the result is made warp-uniform.
// kProf builds only: `slot` >= 0 on exactly one lane of a role makes that lane account its blocked time (the
// first try_wait may itself suspend the thread, so the whole call is timed).
template <bool kProf, bool kCluster = false, int kSleepNs = 0>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, Control* ctl, int* status, int role, int id,
                                          int slot = -1) {
    unsigned long long t0 = 0;
    if (kProf && slot >= 0) t0 = global_ns();
    bool ok = mbar_try_wait<kCluster>(bar, parity) != 0;
    if (!ok) {
        ok = mbar_wait_slow<kSleepNs>(bar, parity, ctl, status, role, id);
        if (kCluster && ok) mbar_try_wait<true>(bar, parity);  // (completed: this probe only adds the cluster-scope acquire)
    }
    if (kProf && slot >= 0) ctl->wait_ns[role][slot] += global_ns() - t0;
    return __all_sync(0xffffffffu, ok);
}
// Two barriers at once: both probes are in flight together (a probe of an already-completed phase still
// costs ~90 cycles), the slow path is only entered for the one that is really pending.
template <bool kProf>
__device__ __forceinline__ bool mbar_wait2(uint32_t bar_a, uint32_t parity_a, int id_a, uint32_t bar_b, uint32_t parity_b,
                                           int id_b, Control* ctl, int* status, int role, int slot = -1) {
    unsigned long long t0 = 0;
    if (kProf && slot >= 0) t0 = global_ns();
    const bool ok_a = mbar_try_wait(bar_a, parity_a) != 0;
    const bool ok_b = mbar_try_wait(bar_b, parity_b) != 0;
    bool ok = true;
    if (!ok_a) ok = mbar_wait_slow(bar_a, parity_a, ctl, status, role, id_a);
    if (kProf && slot >= 0) {
        const unsigned long long t1 = global_ns();
        ctl->wait_ns[role][slot] += t1 - t0;
        t0 = t1;
    }
    if (ok && !ok_b) ok = mbar_wait_slow(bar_b, parity_b, ctl, status, role, id_b);
    if (kProf && slot >= 0) ctl->wait_ns[role][slot + 1] += global_ns() - t0;
    return __all_sync(0xffffffffu, ok);
}

}  // namespace

}  // namespace dcbf
