/*
 * dcbf.h -- C ABI of libdcbf.so, the B200 (sm_100a) tied-array beamforming hot path.
 *
 * This is the drop-in boundary: plain C, device pointers + sizes, no torch / C++ types.
 * Each entry point replaces one launch site of the reference (paths relative to the
 * reference root, magnate3/dpdk_dc_sand):
 *
 *   dcbf_reorder   <- PreBeamformReorder._run          beamformer/beamforming/prebeamform_reorder.py:171-186
 *                     (kernel prebeamform_reorder,     beamformer/beamforming/kernels/prebeamform_reorder_kernel.mako:37-92)
 *   dcbf_coeffs    <- CoeffGenerator._run              beamformer/beamforming/coeff_generator.py:209-250
 *                     (kernel run_coeff_gen,           beamformer/beamforming/coeff_generator.py:12-103;
 *                      indexing follows the CPU oracle beamformer/unit_test/coeff_generator_cpu.py:120-186)
 *   dcbf_beamform  <- ComplexMultKernel.complex_mult   beamformer/beamforming/complex_mult_kernel.py:106-162
 *                     (kernel run_complex_mult,        beamformer/beamforming/complex_mult_kernel.py:11-100)
 *   dcbf_fused     <- OpSequence.__call__              beamformer/beamforming/beamform_op_sequence.py:117-157
 *                     (the three launches above fused; also supersedes the native precursor
 *                      calculate_beamweights_and_beamform_single_channel,
 *                      beamformer_coefficient_generator/BeamformerKernels.cu:192-367)
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller (16-byte aligned, as any
 *     cudaMalloc / torch allocation is).  The library allocates nothing persistent.
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 *     and is NOT synchronised (the reference calls cuda.synchronize() after every op;
 *     callers that need that call cudaStreamSynchronize themselves).
 *   - The device is the caller's current device (the reference hard-codes device 0).
 *   - Return value: DCBF_OK (0) or a negative dcbf_status; dcbf_strerror() names it.
 *   - Arrays are C-order with exactly the reference's shapes (no padding):
 *       samples     uint8  [B][A][C][T][2 pols][2 re,im]
 *       reordered   uint8  [B][2][C][T/16][16][A][2]
 *       delay_vals  float  [C][M][A][4]  {delay_s, delay_rate, phase_rad, phase_rate}
 *       coeffs      float  [B][P][C][2A][2M]   block (a,m) = [[cos, sin], [-sin, cos]]
 *       beams       float  [B][2][C][T/16][16][2M]  (col 2m = Re, 2m+1 = Im of beam m)
 *   - T must be a positive multiple of 16 (reference: prebeamform_reorder.py:59-65).
 */
#ifndef DCBF_H_
#define DCBF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCBF_VERSION 1

typedef void* dcbf_stream_t; /* cudaStream_t */

typedef enum dcbf_status {
    DCBF_OK = 0,
    DCBF_ERR_INVALID_ARG = -1, /* null pointer, non-positive dimension, T % 16 != 0, misaligned pointer */
    DCBF_ERR_UNSUPPORTED = -2, /* shape outside what the kernels were built for: more than DCBF_MAX_TV_BATCHES heaps
                                  with per-heap times; index overflow */
    DCBF_ERR_CUDA = -3,        /* a CUDA runtime call failed; see dcbf_last_cuda_error() */
    DCBF_ERR_NO_DEVICE = -4,   /* no sm_100 device is current */
    DCBF_ERR_TIMEOUT = -5      /* in-kernel watchdog fired (pipeline dead-lock guard) */
} dcbf_status;

/* dcbf_fused / dcbf_beamform flags */
#define DCBF_FLAG_SIGNED_INPUT 0x1u /* bytes are int8 (F-engine native); default: uint8 like the reference API */
#define DCBF_FLAG_FP16_COEFF 0x2u   /* dcbf_fused: round the steering coefficients once to fp16 (error <= 2^-12 per
                                       component) instead of the default fp16 hi+lo pair (~2^-24); halves tensor work.
                                       Measured error: under a tenth of the 2^-10 * sum|x| output budget; outside the 1e-4
                                       of the reference's own unit tests, hence optional.  On a power-capped board this
                                       is the mode that keeps the kernel HBM-bound (0.91-0.94 of the copy bandwidth
                                       sustained against 0.81-0.85: the tensor cores' power share lowers the SM clock) */
#define DCBF_FLAG_STREAMING 0x4u    /* dcbf_fused: this call is independent of the kernel queued just before it on the
                                       stream (it reads nothing that kernel writes and writes nothing that kernel
                                       reads or writes, e.g. consecutive heaps into alternating output buffers), so its
                                       CTAs may start on SMs that kernel has already left (programmatic dependent
                                       launch).  Copies and events keep their normal stream ordering. */
#define DCBF_FLAG_DEBUG_NO_KSTREAM 0x200u /* dcbf_fused: keep whole B tile sets even when several N tiles are needed (cross-check) */
#define DCBF_FLAG_DEBUG_CUDA_CORES 0x400u /* dcbf_beamform: float32 CUDA-core kernel even where the tcgen05 one applies (cross-check) */
#define DCBF_FLAG_DEBUG_DIRECT_EPILOGUE 0x100u /* dcbf_fused: st.global from registers instead of TMA stores (cross-check) */
#define DCBF_FLAG_DEBUG_WHOLE_CHANNELS 0x800u /* dcbf_fused: never cut the channels of the last scheduling round into pieces (cross-check) */
#define DCBF_FLAG_DEBUG_NO_PDL 0x1000u /* dcbf_fused: plain stream-ordered launch (no programmatic dependent launch attribute) */
#define DCBF_FLAG_DEBUG_NO_PAIR 0x2000u /* dcbf_fused: many antennas x beams on single CTAs instead of cta_group::2 CTA pairs (cross-check) */
#define DCBF_FLAG_DEBUG_NO_BEAM_PIECES 0x4000u /* dcbf_fused: the channels of the last scheduling round are only cut along their tile list, not into halves of the beams (cross-check) */
#define DCBF_FLAG_DEBUG_TWO_A_STAGES 0x8000u /* dcbf_fused: narrow whole tile sets keep the two-stage converted-voltage ring (cross-check / A-B) */

int dcbf_version(void);
const char* dcbf_strerror(int status);
/* Text of the last CUDA error seen by this library on the calling thread ("" if none). */
const char* dcbf_last_cuda_error(void);

/* Stage 1.  samples [B][A][C][T][2][2] u8 -> reordered [B][2][C][T/16][16][A][2] u8.  Bit-exact. */
int dcbf_reorder(const uint8_t* samples, uint8_t* reordered, int n_batches, int n_ants, int n_chans,
                 int n_samples, dcbf_stream_t stream);

/* Stage 2.  delay_vals [C][M][A][4] f32 -> coeffs [B][P][C][2A][2M] f32.
 * rot = delay*ch*(-pi)/(N*Ts) + phase - delay*(N/2)*(-pi)/(N*Ts), ch = c + C*xeng_id, evaluated in
 * float64 with the reference's operation order, cos/sin in float64, stored as float32. */
int dcbf_coeffs(const float* delay_vals, float* coeffs, int n_batches, int n_pols, int n_chans,
                int n_chans_total, int n_ants, int n_beams, int xeng_id, double sample_period,
                dcbf_stream_t stream);

/* Stage 2, time-varying form (next-row feature; the reference Python path ignores the two rate fields, its native
 * precursor beamformer_coefficient_generator/BeamformerKernels.cu:25-35 does not): the coefficients of batch b are
 * evaluated with delay + delay_rate*dt and phase + phase_rate*dt, dt = batch_dt_s[b] seconds since the delay
 * model's reference time (HOST array of n_batches doubles, read during the call; n_batches <= DCBF_MAX_TV_BATCHES). */
#define DCBF_MAX_TV_BATCHES 64
int dcbf_coeffs_tv(const float* delay_vals, float* coeffs, int n_batches, int n_pols, int n_chans,
                   int n_chans_total, int n_ants, int n_beams, int xeng_id, double sample_period,
                   const double* batch_dt_s, dcbf_stream_t stream);

/* Stage 3.  out[b,p,c,t,n] = sum_j f32(reordered[b,p,c,t,j]) * coeffs[b,p,c,j,n], fp32 accumulate.
 * The coefficients are arbitrary float32 (this slot is an input of the reference operator).  Runs on tcgen05: every
 * coefficient is split into three bfloat16 terms (24 significand bits, full float32 exponent range), the 8-bit
 * voltages are exact in bfloat16, accumulation is float32 in TMEM -- the result differs from a float32 evaluation by
 * accumulation order only.  (Odd beam counts, whose 8M-byte rows no tensor map can describe, read their coefficients
 * and write their beams with plain accesses in the same kernel.)  DCBF_FLAG_DEBUG_CUDA_CORES selects the float32
 * CUDA-core kernel that accumulates in the reference's order (cross-checks).  flags: DCBF_FLAG_SIGNED_INPUT. */
int dcbf_beamform(const uint8_t* reordered, const float* coeffs, float* beams, int n_batches, int n_chans,
                  int n_samples, int n_ants, int n_beams, unsigned flags, dcbf_stream_t stream);

/* Stages 1+2+3 in one pass: every voltage byte and every delay value is read from HBM once,
 * neither `reordered` nor `coeffs` is materialised.  tcgen05 (fp16 operands, fp32 accumulate in TMEM).
 * n_chans_total / xeng_id / sample_period as in dcbf_coeffs.  Does not synchronise.
 *
 * Launch ordering: every call is a programmatic dependent launch.  Its prologue (barrier / tensor-memory set-up, L2
 * prefetch of its first inputs) may overlap the tail of the kernel queued before it on the stream; it then waits for
 * that kernel to complete and flush before it reads or writes global memory, so the usual stream semantics hold.
 * DCBF_FLAG_STREAMING removes that wait (see the flag).
 *
 * CUDA graphs: launches can be captured and replayed.  A captured launch keeps one of 64 per-device channel-queue
 * slots for every replay; launches that can run CONCURRENTLY must not share a slot, so: at most 64 captured launches per
 * device may be in flight at once, and one graph must not be replayed concurrently with itself (replays on one stream,
 * or serialised by events, are fine).  Live launches draw from a separate pool of 64 slots per device. */
int dcbf_fused(const uint8_t* samples, const float* delay_vals, float* beams, int n_batches, int n_ants,
               int n_chans, int n_chans_total, int n_samples, int n_beams, int xeng_id, double sample_period,
               unsigned flags, dcbf_stream_t stream);

/* dcbf_fused with time-varying steering: one coefficient set per batch (heap), see dcbf_coeffs_tv. */
int dcbf_fused_tv(const uint8_t* samples, const float* delay_vals, float* beams, int n_batches, int n_ants,
                  int n_chans, int n_chans_total, int n_samples, int n_beams, int xeng_id, double sample_period,
                  const double* batch_dt_s, unsigned flags, dcbf_stream_t stream);

/* Fused path with the beam post-stage folded into the epilogue (next-row feature, SURVEY 8f-2: the reference's
 * output "reorder further in the pipeline is to be expected", beamformer_coefficient_generator/BeamformerKernels.cuh:142-143;
 * tied-array-channelised-voltage streams are 8-bit):
 *     beams_q8[b][p][c][t][2m + x] = int8( clip( rint( beam * beam_gains[m] ), -127, 127 ) )      (round half to even)
 * beam_gains: device float[n_beams]; saturated: optional device counter incremented by the number of clipped values
 * (NULL to skip); batch_dt_s: NULL or per-batch time offsets as in dcbf_fused_tv.  Output traffic drops 4x. */
int dcbf_fused_q8(const uint8_t* samples, const float* delay_vals, const float* beam_gains, int8_t* beams_q8,
                  unsigned long long* saturated, int n_batches, int n_ants, int n_chans, int n_chans_total,
                  int n_samples, int n_beams, int xeng_id, double sample_period, const double* batch_dt_s,
                  unsigned flags, dcbf_stream_t stream);
/* Algorithmic HBM bytes of one dcbf_fused_q8 call (in + delay_vals + gains + int8 out). */
unsigned long long dcbf_fused_q8_bytes(int n_batches, int n_ants, int n_chans, int n_samples, int n_beams);

/* General form of the fused call: every optional feature through one options block (zero-initialise, set
 * struct_size = sizeof(dcbf_fused_options), fill what is needed; NULL opts == dcbf_fused).
 *   batch_dt_s    HOST double[n_batches]  per-heap time offsets, as dcbf_fused_tv
 *   sample_dt_s   seconds between consecutive samples of a heap (the precursor's SAMPLING_PERIOD * FFT_SIZE,
 *                 beamformer_coefficient_generator/BeamformerKernels.cu:153-156, which re-evaluates its coefficients per
 *                 timestamp).  Non-zero (needs batch_dt_s = the time of each heap's FIRST sample): every 128-sample
 *                 time tile of a heap is steered with its own coefficient set, evaluated at the tile's centre
 *                 batch_dt_s[b] + (t0 + (n - 1) / 2) * sample_dt_s, instead of one set per heap -- the residual phase
 *                 drift inside a tile is rate * 64 * sample_dt_s.  0 = one set per heap at batch_dt_s[b].
 *   beam_weights  DEVICE float[n_beams][n_ants]  real weight of every input on every beam, multiplied into the
 *                 steering coefficient -- what the control plane's `?beam-weights <stream> w_0 .. w_{A-1}` request carries
 *                 (reference: ngkcs/ngkcs/corr3_servlet.py:140-153, which only forwards it); update it between calls
 *                 like delay_vals, nothing is cached
 *                 Range: the weights enter the tensor cores as fp16 pairs scaled by 2^10 / 2^beam_weights_log2, so
 *                 max|w| <= 2^beam_weights_log2 must hold (default 0: |w| <= 1; up to 60 is still safe).
 *   beam_weights_log2   e in [-14, 15]: the caller's bound max|w| <= 2^e.  The kernel multiplies the weights by 2^-e and
 *                 the voltages by 2^e (both exact: the voltages' scale is the exponent of the byte -> fp16 conversion),
 *                 so weights from 6e-5 to 3e4 keep the full precision of the coefficient pair.  The Python operators set
 *                 it from the weights themselves.
 *   beams_q8 / beam_gains / saturated   int8 output as dcbf_fused_q8 (then `beams` may be NULL) */
typedef struct dcbf_fused_options {
    size_t struct_size;
    const double* batch_dt_s;
    const float* beam_weights;
    const float* beam_gains;
    int8_t* beams_q8;
    unsigned long long* saturated;
    double sample_dt_s;
    int beam_weights_log2;
} dcbf_fused_options;
int dcbf_fused_ex(const uint8_t* samples, const float* delay_vals, float* beams, int n_batches, int n_ants,
                  int n_chans, int n_chans_total, int n_samples, int n_beams, int xeng_id, double sample_period,
                  const dcbf_fused_options* opts, unsigned flags, dcbf_stream_t stream);
/* dcbf_coeffs with the same options (either pointer may be NULL). */
int dcbf_coeffs_ex(const float* delay_vals, float* coeffs, int n_batches, int n_pols, int n_chans,
                   int n_chans_total, int n_ants, int n_beams, int xeng_id, double sample_period,
                   const double* batch_dt_s, const float* beam_weights, dcbf_stream_t stream);

/* dcbf_coeffs_ex with half-precision output: coeffs_f16 [B][P][C][2A][2M] fp16 (IEEE binary16), every value the
 * round-to-nearest fp16 of the float32 coefficient -- the 16-bit output option of the native precursor
 * (beamformer_coefficient_generator/BeamformerKernels.cu:113-115, 172-185: __floats2half2_rn of (real, imag)), in this
 * library's real-expanded layout.  Halves the coefficient traffic of a consumer that does not need 24 bits. */
int dcbf_coeffs_f16(const float* delay_vals, void* coeffs_f16, int n_batches, int n_pols, int n_chans,
                    int n_chans_total, int n_ants, int n_beams, int xeng_id, double sample_period,
                    const double* batch_dt_s, const float* beam_weights, dcbf_stream_t stream);

/* Synchronises the whole current device (cudaDeviceSynchronize: every stream, non-blocking ones included), then
 * returns the status the dcbf_fused / dcbf_beamform kernels left behind: DCBF_OK, or DCBF_ERR_TIMEOUT if an in-kernel
 * pipeline wait exceeded its 2 s guard (the kernel then exits early instead of hanging, leaving partly written beams;
 * *role / *barrier / *block say who waited on what).  Clears the status. */
int dcbf_fused_status(int* role, int* barrier, int* block);
/* Non-blocking form: DCBF_OK, or the error code a kernel has raised since the status was last cleared (a page-locked
 * flag the failing kernel writes).  It covers work the caller has already synchronised with (after
 * cudaStreamSynchronize / an event wait); dcbf_host_plan_run* check it themselves and return the error. */
int dcbf_fused_status_poll(void);

/* Packed steering coefficients: the delay model changes at control-plane cadence (the reference rebuilds its
 * coefficient array on every call of the sequence, beamformer/beamforming/beamform_op_sequence.py:117-157; its native
 * precursor has stand-alone generators that leave a coefficient array in device memory for a later beamformer,
 * beamformer_coefficient_generator/BeamformerKernels.cu:7-54, 56-119), the heaps arrive every fraction of a
 * millisecond.  dcbf_fused_pack_coeffs evaluates the steering coefficients of one
 * delay model ONCE and stores them in the layout the tensor cores read (fp16 hi + lo rows, 128-byte swizzle: the very
 * bytes dcbf_fused builds per channel in shared memory); dcbf_fused_packed is dcbf_fused with those tile sets loaded by
 * one bulk copy per channel instead of 4096 phase / sin-cos evaluations per channel and heap.  Same HBM traffic
 * (a tile set is as large as the channel's delay_vals for the hi+lo pair), bit-identical beams, a fraction of the
 * SM-side work -- which is what a power-capped board runs out of first.
 * Static steering, shapes whose tile set fits one shared-memory buffer (dcbf_fused_tiling nt_count == 1:
 * up to 64 antennas x 64 beams, 128 x 32, 256 x 16 ...); otherwise dcbf_fused_packed_bytes returns 0 and the two calls
 * DCBF_ERR_UNSUPPORTED.  flags: DCBF_FLAG_FP16_COEFF must be the same in all three calls; DCBF_FLAG_SIGNED_INPUT and
 * DCBF_FLAG_STREAMING as for dcbf_fused (a streaming launch must not directly follow the pack it reads on the stream:
 * it would not wait for it).  beam_weights: optional [n_beams][n_ants] real weights folded into the
 * coefficients (as dcbf_fused_ex with beam_weights_log2 = 0), or NULL. */
unsigned long long dcbf_fused_packed_bytes(int n_ants, int n_chans, int n_beams, unsigned flags);
int dcbf_fused_pack_coeffs(const float* delay_vals, void* packed, int n_ants, int n_chans, int n_chans_total, int n_beams,
                           int xeng_id, double sample_period, const float* beam_weights, unsigned flags,
                           dcbf_stream_t stream);
int dcbf_fused_packed(const uint8_t* samples, const void* packed, float* beams, int n_batches, int n_ants, int n_chans,
                      int n_chans_total, int n_samples, int n_beams, int xeng_id, double sample_period, unsigned flags,
                      dcbf_stream_t stream);

/* The same for the int8 requantised output (dcbf_fused_q8): the per-beam gains ride on the coefficients (relative to
 * max|gain|), so they are part of the packed tile sets -- pack again when the gains change -- and are passed to the hot
 * call as well (its epilogue scales by max|gain|).  Bit-identical to dcbf_fused_q8 with the same arguments. */
int dcbf_fused_pack_coeffs_q8(const float* delay_vals, const float* beam_gains, void* packed, int n_ants, int n_chans,
                              int n_chans_total, int n_beams, int xeng_id, double sample_period, unsigned flags,
                              dcbf_stream_t stream);
int dcbf_fused_packed_q8(const uint8_t* samples, const void* packed, const float* beam_gains, int8_t* beams_q8,
                         unsigned long long* saturated, int n_batches, int n_ants, int n_chans, int n_chans_total,
                         int n_samples, int n_beams, int xeng_id, double sample_period, unsigned flags,
                         dcbf_stream_t stream);

/* Developer aid: when non-NULL, every dcbf_fused CTA writes 24 uint64 to dev_ptr[blockIdx*24 + role*4 + slot]:
 * nanoseconds each warp role spent blocked per barrier class (slot 0..2) and the role's span (slot 3).
 * The buffer must hold 24 * (number of SMs) entries. */
void dcbf_debug_set_profile_buffer(unsigned long long* dev_ptr);

/* The whole-tile-set tiling of dcbf_fused for (n_ants, n_beams, flags): k-blocks of 32 antennas, N tiles of *nt
 * columns such that one tile set (all k-blocks, hi+lo) fits a 64 KiB buffer.  When that needs more than one N tile
 * (*nt_count > 1) the kernel streams B by k-blocks instead and uses N tiles
 * of up to 128 columns (ceil(2 n_beams / 128) of them), two 128-sample time tiles at a time. */
void dcbf_fused_tiling(int n_ants, int n_beams, unsigned flags, int* kb_count, int* nt, int* nt_count);

/* ---- host-buffer entry point (what a caller holding HOST arrays binds to) --------------------------------
 * Replaces the reference's set() / op() / get() round trip (beamformer/unit_test/beamform_op_sequence_test.py:156-163).
 * A plan owns n_slots device workspaces + streams; dcbf_host_plan_run cuts the channel axis into chunks of
 * chunk_chans channels and pipelines H2D -> dcbf_fused -> D2H across the slots, then blocks until `beams` is
 * complete in host memory.  Host arrays have the full reference shapes; pin them (cudaHostRegister /
 * torch pin_memory) for the copies to overlap.  chunk_chans <= 0: one chunk; n_slots <= 0: 3. */
typedef void* dcbf_host_plan_t;
int dcbf_host_plan_create(dcbf_host_plan_t* plan, int n_batches, int n_ants, int n_chans, int n_chans_total,
                          int n_samples, int n_beams, int xeng_id, double sample_period, unsigned flags,
                          int chunk_chans, int n_slots);
int dcbf_host_plan_run(dcbf_host_plan_t plan, const uint8_t* samples, const float* delay_vals, float* beams);
/* Resident delay model: upload the whole HOST [C][M][A][4] table once (and again whenever the control plane changes
 * it: the new table goes into the copy that runs already issued do not read); runs that pass delay_vals = NULL then
 * use it, so the per-step host-to-device traffic is the voltages alone.  Blocks until the table is on the device. */
int dcbf_host_plan_set_delay_vals(dcbf_host_plan_t plan, const float* delay_vals);
/* Requantised-output runs of a plan (dcbf_fused_q8 per chunk): set the per-beam gains (HOST float[n_beams]) once,
 * then run; beams_q8 is a HOST int8 array [B][2][C][T/16][16][2M]; *saturated (may be NULL) gets the clip count. */
int dcbf_host_plan_set_gains(dcbf_host_plan_t plan, const float* beam_gains);
int dcbf_host_plan_run_q8(dcbf_host_plan_t plan, const uint8_t* samples, const float* delay_vals, int8_t* beams_q8,
                          unsigned long long* saturated);
int dcbf_host_plan_destroy(dcbf_host_plan_t plan);

/* ---- ingest: F-engine heaps -> [B][A][C][T][2][2] chunks in page-locked memory ---------------------------------
 * The stage before the path (reference: fgpu_send_prototype/fgpu_send_prototype.py:20-22,55-60 defines the heap an
 * F-engine sends: timestamp 0x1600, feng_id 0x4101, feng_raw 0x4300 = int8 [n_chans][n_samples][2][2];
 * beamformer/README.md:5-6 the batch layout).  A ring of n_chunks chunks of n_batches consecutive heaps x n_ants;
 * heaps are placed by (timestamp, feng_id) in any arrival order; n_chunks/2 chunks can be receiving at once, the
 * rest hold finished chunks until released.  A chunk is finished when all its heaps have arrived, or when newer
 * heaps push the window past it (missing heaps are zero-filled and reported).  No network code: the receive loop
 * calls dcbf_ingest_packet (raw SPEAD packets), dcbf_ingest_heap (copy of a complete heap) or
 * dcbf_ingest_heap_ptr + dcbf_ingest_heap_done (write in place).
 * timestamp_step = ADC samples between consecutive heaps of one antenna; pinned = 0 uses ordinary memory (no GPU
 * needed, for tests).  Thread-safe.  dcbf_ingest_heap returns DCBF_ERR_UNSUPPORTED for a heap that was dropped
 * (too old, duplicate of a closed chunk, or no free chunk because the consumer is behind). */
typedef void* dcbf_ingest_t;
int dcbf_ingest_create(dcbf_ingest_t* ingest, int n_chunks, int n_batches, int n_ants, int n_chans, int n_samples,
                       long long timestamp_step, int pinned);
int dcbf_ingest_heap(dcbf_ingest_t ingest, long long timestamp, int feng_id, const void* payload);
int dcbf_ingest_heap_ptr(dcbf_ingest_t ingest, long long timestamp, int feng_id, void** dst);
int dcbf_ingest_heap_done(dcbf_ingest_t ingest, long long timestamp, int feng_id);
/* One SPEAD-64-48 packet as received from the network (the wire format of fgpu_send_prototype.py:18,55-60 and of the
 * MeerKAT F-engines): its payload is written straight to its place in the chunk (heap offset item 0x0003), the heap
 * counts as arrived once all its bytes have.  timestamp 0x1600 and feng_id 0x4101 are taken from the packet, or, for
 * packets that do not repeat the item pointers, from the heap's first packet (a packet that overtakes it is dropped);
 * default_feng_id is used when the sender has no 0x4101 item (the prototype: one sender = one antenna), -1 = none.
 * Descriptor heaps and heaps of another sub-band (dcbf_ingest_set_frequency, item 0x4103) return
 * DCBF_ERR_UNSUPPORTED; malformed packets DCBF_ERR_INVALID_ARG and count in n_bad. */
int dcbf_ingest_packet(dcbf_ingest_t ingest, const void* packet, size_t length, int default_feng_id);
int dcbf_ingest_set_frequency(dcbf_ingest_t ingest, long long first_channel);
/* Next finished chunk in time order: returns 1 and sets *samples (the `samples` argument of dcbf_host_plan_run),
 * *first_timestamp, *n_missing and present[n_batches * n_ants] (each may be NULL); 0 if none is ready; flush != 0
 * also hands out chunks that are still incomplete (end of stream).  Give the chunk back with dcbf_ingest_release. */
int dcbf_ingest_pop(dcbf_ingest_t ingest, int flush, const uint8_t** samples, long long* first_timestamp,
                    int* n_missing, uint8_t* present);
int dcbf_ingest_release(dcbf_ingest_t ingest, const uint8_t* samples);
int dcbf_ingest_stats(dcbf_ingest_t ingest, unsigned long long* n_late, unsigned long long* n_duplicate,
                      unsigned long long* n_bad);
int dcbf_ingest_destroy(dcbf_ingest_t ingest);

/* Number of kernel launches this library has issued since load (all entry points); used by bench.py
 * to report gpu_launches. */
unsigned long long dcbf_launch_count(void);

/* Algorithmic HBM bytes of one dcbf_fused call (in + delay_vals + out). */
unsigned long long dcbf_fused_bytes(int n_batches, int n_ants, int n_chans, int n_samples, int n_beams);

#ifdef __cplusplus
}
#endif
#endif /* DCBF_H_ */
