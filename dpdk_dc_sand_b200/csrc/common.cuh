// Shared host/device helpers for libdcbf (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "dcbf.h"

namespace dcbf {

constexpr int kPols = 2;            // reference: prebeamform_reorder.py:54 ("Hardcoded to 2")
constexpr int kSamplesPerBlock = 16;  // reference: prebeamform_reorder.py:59

// Records a CUDA error for dcbf_last_cuda_error() and maps it to DCBF_ERR_CUDA.
int record_cuda_error(cudaError_t err, const char* what);
void count_launch(unsigned n = 1);

#define DCBF_CUDA_TRY(expr)                                              \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) return ::dcbf::record_cuda_error(_e, #expr); \
    } while (0)

// Launch errors right after a kernel launch (cudaGetLastError returns AND clears the error; it is recorded for
// dcbf_last_cuda_error and the call fails).
#define DCBF_CHECK_LAUNCH(name)                                         \
    do {                                                                \
        cudaError_t _e = cudaGetLastError();                            \
        if (_e != cudaSuccess) return ::dcbf::record_cuda_error(_e, name); \
        ::dcbf::count_launch();                                         \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Kernel launchers (one per translation unit).
int launch_reorder(const uint8_t* samples, uint8_t* reordered, int B, int A, int C, int T, cudaStream_t s);
// batch_dt_s: NULL (static steering) or B host doubles (seconds since the delay model's reference time)
int launch_coeffs(const float* delay_vals, float* coeffs, int B, int P, int C, int N, int A, int M, int xeng_id,
                  double sample_period, const double* batch_dt_s, const float* beam_weights, cudaStream_t s);
// fp16 output in the same layout (the precursor's 16-bit coefficient option)
int launch_coeffs_f16(const float* delay_vals, void* coeffs, int B, int P, int C, int N, int A, int M, int xeng_id,
                      double sample_period, const double* batch_dt_s, const float* beam_weights, cudaStream_t s);
int launch_beamform(const uint8_t* reordered, const float* coeffs, float* beams, int B, int C, int T, int A, int M,
                    unsigned flags, cudaStream_t s);
// tcgen05 version (beamform_tc.cu) for shapes its TMA descriptors can express: even beam count
bool beamform_tc_supported(const void* reordered, const void* coeffs, const void* beams, int A, int M);
int launch_beamform_tc(const uint8_t* reordered, const float* coeffs, float* beams, int B, int C, int T, int A, int M,
                       unsigned flags, cudaStream_t s);
// Requantised-output variant of the fused kernel (dcbf_fused_q8): int8 beams, per-beam gains, saturation counter.
struct QuantisedOut {
    int8_t* beams;
    const float* gains;
    unsigned long long* saturated;
};
// first_chan = absolute F-engine channel of local channel 0 (n_chans * xeng_id for a whole stream).
int launch_fused(const uint8_t* samples, const float* delay_vals, float* beams, int B, int A, int C, int N, int T,
                 int M, long long first_chan, double sample_period, const double* batch_dt_s, unsigned flags,
                 cudaStream_t s, const QuantisedOut* q8 = nullptr, const float* beam_weights = nullptr,
                 double sample_dt_s = 0.0, int weights_log2 = 0, int packed_mode = 0, uint8_t* packed = nullptr);
int fused_status(int* role, int* barrier, int* block);
int fused_status_poll();
void fused_set_profile_buffer(unsigned long long* dev_ptr);
void fused_tiling(int A, int M, unsigned flags, int* kb_count, int* nt, int* nt_count);

}  // namespace dcbf
