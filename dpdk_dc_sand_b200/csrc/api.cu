// extern "C" entry points of libdcbf.so (declared in include/dcbf.h): argument validation,
// error bookkeeping and dispatch to the kernel launchers.  No kernels live here.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace dcbf {

static thread_local char g_last_cuda_error[256] = "";
static std::atomic<unsigned long long> g_launches{0};

int record_cuda_error(cudaError_t err, const char* what) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s (%s)", what, cudaGetErrorName(err),
             cudaGetErrorString(err));
    return DCBF_ERR_CUDA;
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int check_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return DCBF_ERR_NO_DEVICE;
    }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return DCBF_ERR_NO_DEVICE;
    }
    // The library carries sm_100a SASS only: anything else cannot run it.
    return major == 10 ? DCBF_OK : DCBF_ERR_NO_DEVICE;
}

static bool bad_t(int T) { return T <= 0 || (T % kSamplesPerBlock) != 0; }

}  // namespace dcbf

using namespace dcbf;

#pragma GCC visibility push(default)
extern "C" {

int dcbf_version(void) { return DCBF_VERSION; }

const char* dcbf_strerror(int status) {
    switch (status) {
        case DCBF_OK: return "ok";
        case DCBF_ERR_INVALID_ARG: return "invalid argument (null/misaligned pointer, non-positive dimension or n_samples % 16 != 0)";
        case DCBF_ERR_UNSUPPORTED: return "unsupported shape for this build";
        case DCBF_ERR_CUDA: return "CUDA runtime error (see dcbf_last_cuda_error)";
        case DCBF_ERR_NO_DEVICE: return "current device is not an sm_100 (B200) GPU";
        case DCBF_ERR_TIMEOUT: return "in-kernel watchdog fired";
        default: return "unknown dcbf status";
    }
}

const char* dcbf_last_cuda_error(void) { return g_last_cuda_error; }

unsigned long long dcbf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

unsigned long long dcbf_fused_bytes(int B, int A, int C, int T, int M) {
    const unsigned long long in = 1ull * B * A * C * T * kPols * 2;
    const unsigned long long dv = 1ull * C * M * A * 16;
    const unsigned long long out = 1ull * B * kPols * C * T * M * 8;
    return in + dv + out;
}

int dcbf_reorder(const uint8_t* samples, uint8_t* reordered, int B, int A, int C, int T, dcbf_stream_t stream) {
    if (!samples || !reordered || B <= 0 || A <= 0 || C <= 0 || bad_t(T)) return DCBF_ERR_INVALID_ARG;
    if (!aligned16(samples) || !aligned16(reordered)) return DCBF_ERR_INVALID_ARG;
    if (int e = check_device()) return e;
    return launch_reorder(samples, reordered, B, A, C, T, static_cast<cudaStream_t>(stream));
}

int dcbf_coeffs(const float* delay_vals, float* coeffs, int B, int P, int C, int N, int A, int M, int xeng_id,
                double sample_period, dcbf_stream_t stream) {
    if (!delay_vals || !coeffs || B <= 0 || P <= 0 || C <= 0 || N <= 0 || A <= 0 || M <= 0 || xeng_id < 0 ||
        !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(delay_vals) || !aligned16(coeffs)) return DCBF_ERR_INVALID_ARG;
    if (int e = check_device()) return e;
    return launch_coeffs(delay_vals, coeffs, B, P, C, N, A, M, xeng_id, sample_period, nullptr, nullptr,
                         static_cast<cudaStream_t>(stream));
}

int dcbf_coeffs_tv(const float* delay_vals, float* coeffs, int B, int P, int C, int N, int A, int M, int xeng_id,
                   double sample_period, const double* batch_dt_s, dcbf_stream_t stream) {
    if (!delay_vals || !coeffs || !batch_dt_s || B <= 0 || P <= 0 || C <= 0 || N <= 0 || A <= 0 || M <= 0 ||
        xeng_id < 0 || !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(delay_vals) || !aligned16(coeffs)) return DCBF_ERR_INVALID_ARG;
    if (B > DCBF_MAX_TV_BATCHES) return DCBF_ERR_UNSUPPORTED;
    if (int e = check_device()) return e;
    return launch_coeffs(delay_vals, coeffs, B, P, C, N, A, M, xeng_id, sample_period, batch_dt_s, nullptr,
                         static_cast<cudaStream_t>(stream));
}

int dcbf_coeffs_ex(const float* delay_vals, float* coeffs, int B, int P, int C, int N, int A, int M, int xeng_id,
                   double sample_period, const double* batch_dt_s, const float* beam_weights, dcbf_stream_t stream) {
    if (!delay_vals || !coeffs || B <= 0 || P <= 0 || C <= 0 || N <= 0 || A <= 0 || M <= 0 || xeng_id < 0 ||
        !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(delay_vals) || !aligned16(coeffs)) return DCBF_ERR_INVALID_ARG;
    if (batch_dt_s && B > DCBF_MAX_TV_BATCHES) return DCBF_ERR_UNSUPPORTED;
    if (int e = check_device()) return e;
    return launch_coeffs(delay_vals, coeffs, B, P, C, N, A, M, xeng_id, sample_period, batch_dt_s, beam_weights,
                         static_cast<cudaStream_t>(stream));
}

int dcbf_coeffs_f16(const float* delay_vals, void* coeffs_f16, int B, int P, int C, int N, int A, int M, int xeng_id,
                    double sample_period, const double* batch_dt_s, const float* beam_weights, dcbf_stream_t stream) {
    if (!delay_vals || !coeffs_f16 || B <= 0 || P <= 0 || C <= 0 || N <= 0 || A <= 0 || M <= 0 || xeng_id < 0 ||
        !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(delay_vals) || (reinterpret_cast<uintptr_t>(coeffs_f16) & 3u)) return DCBF_ERR_INVALID_ARG;
    if (batch_dt_s && B > DCBF_MAX_TV_BATCHES) return DCBF_ERR_UNSUPPORTED;
    if (int e = check_device()) return e;
    return launch_coeffs_f16(delay_vals, coeffs_f16, B, P, C, N, A, M, xeng_id, sample_period, batch_dt_s, beam_weights,
                             static_cast<cudaStream_t>(stream));
}

int dcbf_beamform(const uint8_t* reordered, const float* coeffs, float* beams, int B, int C, int T, int A, int M,
                  unsigned flags, dcbf_stream_t stream) {
    if (!reordered || !coeffs || !beams || B <= 0 || C <= 0 || A <= 0 || M <= 0 || bad_t(T))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(reordered) || !aligned16(coeffs) || !aligned16(beams)) return DCBF_ERR_INVALID_ARG;
    if (int e = check_device()) return e;
    if (!(flags & DCBF_FLAG_DEBUG_CUDA_CORES) && beamform_tc_supported(reordered, coeffs, beams, A, M))
        return launch_beamform_tc(reordered, coeffs, beams, B, C, T, A, M, flags, static_cast<cudaStream_t>(stream));
    return launch_beamform(reordered, coeffs, beams, B, C, T, A, M, flags, static_cast<cudaStream_t>(stream));
}

int dcbf_fused(const uint8_t* samples, const float* delay_vals, float* beams, int B, int A, int C, int N, int T, int M,
               int xeng_id, double sample_period, unsigned flags, dcbf_stream_t stream) {
    if (!samples || !delay_vals || !beams || B <= 0 || A <= 0 || C <= 0 || N <= 0 || M <= 0 || xeng_id < 0 ||
        bad_t(T) || !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(samples) || !aligned16(delay_vals) || !aligned16(beams)) return DCBF_ERR_INVALID_ARG;
    if (int e = check_device()) return e;
    return launch_fused(samples, delay_vals, beams, B, A, C, N, T, M, static_cast<long long>(C) * xeng_id, sample_period,
                        nullptr, flags, static_cast<cudaStream_t>(stream));
}

int dcbf_fused_tv(const uint8_t* samples, const float* delay_vals, float* beams, int B, int A, int C, int N, int T,
                  int M, int xeng_id, double sample_period, const double* batch_dt_s, unsigned flags,
                  dcbf_stream_t stream) {
    if (!samples || !delay_vals || !beams || !batch_dt_s || B <= 0 || A <= 0 || C <= 0 || N <= 0 || M <= 0 ||
        xeng_id < 0 || bad_t(T) || !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(samples) || !aligned16(delay_vals) || !aligned16(beams)) return DCBF_ERR_INVALID_ARG;
    if (B > DCBF_MAX_TV_BATCHES) return DCBF_ERR_UNSUPPORTED;
    if (int e = check_device()) return e;
    return launch_fused(samples, delay_vals, beams, B, A, C, N, T, M, static_cast<long long>(C) * xeng_id, sample_period,
                        batch_dt_s, flags, static_cast<cudaStream_t>(stream));
}

int dcbf_fused_q8(const uint8_t* samples, const float* delay_vals, const float* beam_gains, int8_t* beams_q8,
                  unsigned long long* saturated, int B, int A, int C, int N, int T, int M, int xeng_id,
                  double sample_period, const double* batch_dt_s, unsigned flags, dcbf_stream_t stream) {
    if (!samples || !delay_vals || !beam_gains || !beams_q8 || B <= 0 || A <= 0 || C <= 0 || N <= 0 || M <= 0 ||
        xeng_id < 0 || bad_t(T) || !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(samples) || !aligned16(delay_vals) || !aligned16(beams_q8)) return DCBF_ERR_INVALID_ARG;
    if (batch_dt_s && B > DCBF_MAX_TV_BATCHES) return DCBF_ERR_UNSUPPORTED;
    if (int e = check_device()) return e;
    const QuantisedOut q8{beams_q8, beam_gains, saturated};
    return launch_fused(samples, delay_vals, nullptr, B, A, C, N, T, M, static_cast<long long>(C) * xeng_id, sample_period,
                        batch_dt_s, flags, static_cast<cudaStream_t>(stream), &q8);
}

unsigned long long dcbf_fused_q8_bytes(int B, int A, int C, int T, int M) {
    const unsigned long long in = 1ull * B * A * C * T * kPols * 2;
    const unsigned long long dv = 1ull * C * M * A * 16;
    const unsigned long long out = 1ull * B * kPols * C * T * M * 2;
    return in + dv + out + 4ull * M;
}

int dcbf_fused_ex(const uint8_t* samples, const float* delay_vals, float* beams, int B, int A, int C, int N, int T,
                  int M, int xeng_id, double sample_period, const dcbf_fused_options* opts, unsigned flags,
                  dcbf_stream_t stream) {
    dcbf_fused_options o{};
    if (opts) {
        if (opts->struct_size < sizeof(size_t) || opts->struct_size > sizeof(o)) return DCBF_ERR_INVALID_ARG;
        memcpy(&o, opts, opts->struct_size);  // older, shorter structs leave the newer fields zero
    }
    const bool q8 = o.beams_q8 != nullptr;
    if (!samples || !delay_vals || (!beams && !q8) || (q8 && !o.beam_gains) || B <= 0 || A <= 0 || C <= 0 || N <= 0 ||
        M <= 0 || xeng_id < 0 || bad_t(T) || !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(samples) || !aligned16(delay_vals) || !aligned16(q8 ? static_cast<void*>(o.beams_q8) : beams))
        return DCBF_ERR_INVALID_ARG;
    if (o.batch_dt_s && B > DCBF_MAX_TV_BATCHES) return DCBF_ERR_UNSUPPORTED;
    if (!(o.sample_dt_s == o.sample_dt_s) || (o.sample_dt_s != 0.0 && !o.batch_dt_s)) return DCBF_ERR_INVALID_ARG;
    if (o.beam_weights_log2 < -14 || o.beam_weights_log2 > 15) return DCBF_ERR_INVALID_ARG;
    if (int e = check_device()) return e;
    const QuantisedOut qo{o.beams_q8, o.beam_gains, o.saturated};
    return launch_fused(samples, delay_vals, beams, B, A, C, N, T, M, static_cast<long long>(C) * xeng_id, sample_period,
                        o.batch_dt_s, flags, static_cast<cudaStream_t>(stream), q8 ? &qo : nullptr, o.beam_weights,
                        o.batch_dt_s ? o.sample_dt_s : 0.0, o.beam_weights ? o.beam_weights_log2 : 0);
}

unsigned long long dcbf_fused_packed_bytes(int A, int C, int M, unsigned flags) {
    if (A <= 0 || C <= 0 || M <= 0) return 0;
    int kb = 0, nt = 0, ntc = 0;
    fused_tiling(A, M, flags, &kb, &nt, &ntc);
    if (ntc != 1 || nt < 16) return 0;  // K-streamed shapes keep no whole tile set
    const int parts = (flags & DCBF_FLAG_FP16_COEFF) ? 1 : 2;
    return 1ull * C * kb * parts * nt * 128;
}

int dcbf_fused_pack_coeffs(const float* delay_vals, void* packed, int A, int C, int N, int M, int xeng_id,
                           double sample_period, const float* beam_weights, unsigned flags, dcbf_stream_t stream) {
    if (!delay_vals || !packed || A <= 0 || C <= 0 || N <= 0 || M <= 0 || xeng_id < 0 || !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(delay_vals) || !aligned16(packed)) return DCBF_ERR_INVALID_ARG;
    if (!dcbf_fused_packed_bytes(A, C, M, flags)) return DCBF_ERR_UNSUPPORTED;
    if (int e = check_device()) return e;
    // (the kernel of the hot path with its voltage and beam roles idle: the tile sets are bit for bit what dcbf_fused
    // builds in shared memory; the voltage / beam pointers only have to be valid addresses)
    uint8_t* img = static_cast<uint8_t*>(packed);
    return launch_fused(img, delay_vals, reinterpret_cast<float*>(img), 1, A, C, N, 128, M, static_cast<long long>(C) * xeng_id,
                        sample_period, nullptr, flags & (DCBF_FLAG_FP16_COEFF | DCBF_FLAG_SIGNED_INPUT), static_cast<cudaStream_t>(stream),
                        nullptr, beam_weights, 0.0, 0, 1, img);
}

int dcbf_fused_packed(const uint8_t* samples, const void* packed, float* beams, int B, int A, int C, int N, int T, int M,
                      int xeng_id, double sample_period, unsigned flags, dcbf_stream_t stream) {
    if (!samples || !packed || !beams || B <= 0 || A <= 0 || C <= 0 || N <= 0 || M <= 0 || xeng_id < 0 || bad_t(T) ||
        !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(samples) || !aligned16(packed) || !aligned16(beams)) return DCBF_ERR_INVALID_ARG;
    if (!dcbf_fused_packed_bytes(A, C, M, flags)) return DCBF_ERR_UNSUPPORTED;
    if (int e = check_device()) return e;
    return launch_fused(samples, nullptr, beams, B, A, C, N, T, M, static_cast<long long>(C) * xeng_id, sample_period, nullptr,
                        flags, static_cast<cudaStream_t>(stream), nullptr, nullptr, 0.0, 0, 2,
                        const_cast<uint8_t*>(static_cast<const uint8_t*>(packed)));
}

int dcbf_fused_pack_coeffs_q8(const float* delay_vals, const float* beam_gains, void* packed, int A, int C, int N, int M,
                              int xeng_id, double sample_period, unsigned flags, dcbf_stream_t stream) {
    if (!delay_vals || !beam_gains || !packed || A <= 0 || C <= 0 || N <= 0 || M <= 0 || xeng_id < 0 || !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(delay_vals) || !aligned16(packed)) return DCBF_ERR_INVALID_ARG;
    if (!dcbf_fused_packed_bytes(A, C, M, flags)) return DCBF_ERR_UNSUPPORTED;
    if (int e = check_device()) return e;
    uint8_t* img = static_cast<uint8_t*>(packed);
    const QuantisedOut q8{reinterpret_cast<int8_t*>(img), beam_gains, nullptr};  // (the int8 kernel: its coefficients carry gain / max|gain|)
    return launch_fused(img, delay_vals, nullptr, 1, A, C, N, 128, M, static_cast<long long>(C) * xeng_id, sample_period, nullptr,
                        flags & (DCBF_FLAG_FP16_COEFF | DCBF_FLAG_SIGNED_INPUT), static_cast<cudaStream_t>(stream), &q8, nullptr,
                        0.0, 0, 1, img);
}

int dcbf_fused_packed_q8(const uint8_t* samples, const void* packed, const float* beam_gains, int8_t* beams_q8,
                         unsigned long long* saturated, int B, int A, int C, int N, int T, int M, int xeng_id,
                         double sample_period, unsigned flags, dcbf_stream_t stream) {
    if (!samples || !packed || !beam_gains || !beams_q8 || B <= 0 || A <= 0 || C <= 0 || N <= 0 || M <= 0 || xeng_id < 0 ||
        bad_t(T) || !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (!aligned16(samples) || !aligned16(packed) || !aligned16(beams_q8)) return DCBF_ERR_INVALID_ARG;
    if (!dcbf_fused_packed_bytes(A, C, M, flags)) return DCBF_ERR_UNSUPPORTED;
    if (int e = check_device()) return e;
    const QuantisedOut q8{beams_q8, beam_gains, saturated};
    return launch_fused(samples, nullptr, nullptr, B, A, C, N, T, M, static_cast<long long>(C) * xeng_id, sample_period, nullptr,
                        flags, static_cast<cudaStream_t>(stream), &q8, nullptr, 0.0, 0, 2,
                        const_cast<uint8_t*>(static_cast<const uint8_t*>(packed)));
}

int dcbf_fused_status(int* role, int* barrier, int* block) {
    if (int e = check_device()) return e;
    return fused_status(role, barrier, block);
}

int dcbf_fused_status_poll(void) { return fused_status_poll(); }

void dcbf_debug_set_profile_buffer(unsigned long long* dev_ptr) { fused_set_profile_buffer(dev_ptr); }

void dcbf_fused_tiling(int A, int M, unsigned flags, int* kb_count, int* nt, int* nt_count) {
    int a = 0, b = 0, c = 0;
    if (A > 0 && M > 0) fused_tiling(A, M, flags, &a, &b, &c);
    if (kb_count) *kb_count = a;
    if (nt) *nt = b;
    if (nt_count) *nt_count = c;
}

}  // extern "C"
#pragma GCC visibility pop
