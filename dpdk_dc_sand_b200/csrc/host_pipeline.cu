// Host-buffer entry point of the fused path: what a caller that owns HOST arrays (the reference's tests do:
// `buf.set(queue, host)` ... `op()` ... `buf.get(queue, host)`, beamform_op_sequence_test.py:156-163) binds to.
//
// A plan owns `n_slots` device workspaces and streams.  The channel axis is cut into chunks; chunk i runs on
// slot i % n_slots as  H2D(samples chunk, strided 2-D copy) -> H2D(delay_vals chunk) -> fused kernel ->
// D2H(beams chunk, strided 2-D copy),  so the copy engines (one per direction) and the SMs overlap across
// chunks.  Channels are independent, so chunking changes nothing in the result.
#include <vector>

#include "common.cuh"

namespace dcbf {

struct HostPlan {
    int B, A, C, N, T, M, xeng_id;
    double sample_period;
    unsigned flags;
    int chunk;  // channels per chunk
    int device;
    struct Slot {
        cudaStream_t stream = nullptr;
        uint8_t* samples = nullptr;
        float* delay_vals = nullptr;
        float* beams = nullptr;
    };
    std::vector<Slot> slots;
    // Resident delay model (dcbf_host_plan_set_delay_vals): the whole [C][M][A][4] table on the device, double-buffered so
    // that an update lands in the copy the kernels of a run in flight do not read.  Runs that pass delay_vals = NULL use
    // it: the per-step H2D traffic is then the voltages alone (the delay model changes at control-plane cadence).
    float* dv_resident[2] = {nullptr, nullptr};
    int dv_active = -1;
    float* gains = nullptr;                    // device [M], set by dcbf_host_plan_set_gains (q8 runs)
    unsigned long long* saturated = nullptr;   // device counter (q8 runs)
};

static int destroy_plan(HostPlan* p) {
    if (!p) return DCBF_OK;
    for (auto& s : p->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.samples);
        cudaFree(s.delay_vals);
        cudaFree(s.beams);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    cudaFree(p->dv_resident[0]);
    cudaFree(p->dv_resident[1]);
    cudaFree(p->gains);
    cudaFree(p->saturated);
    delete p;
    cudaGetLastError();
    return DCBF_OK;
}

static int create_plan(HostPlan* p) {
    const size_t in_b = static_cast<size_t>(p->B) * p->A * p->chunk * p->T * 4;
    const size_t dv_b = static_cast<size_t>(p->chunk) * p->M * p->A * 16;
    const size_t out_b = static_cast<size_t>(p->B) * kPols * p->chunk * p->T * p->M * 8;
    for (auto& s : p->slots) {
        DCBF_CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        DCBF_CUDA_TRY(cudaMalloc(&s.samples, in_b));
        DCBF_CUDA_TRY(cudaMalloc(&s.delay_vals, dv_b));
        DCBF_CUDA_TRY(cudaMalloc(&s.beams, out_b));
    }
    return DCBF_OK;
}

// h_beams: float32 beams, or int8 requantised beams when q8 (then *h_saturated receives the clip count).
static int run_plan(HostPlan* p, const uint8_t* h_samples, const float* h_dv, void* h_beams, bool q8,
                    unsigned long long* h_saturated) {
    int dev = 0;
    DCBF_CUDA_TRY(cudaGetDevice(&dev));
    if (dev != p->device) return DCBF_ERR_INVALID_ARG;
    if (!h_dv && p->dv_active < 0) return DCBF_ERR_INVALID_ARG;  // dcbf_host_plan_set_delay_vals first
    if (q8) {
        if (!p->gains) return DCBF_ERR_INVALID_ARG;  // dcbf_host_plan_set_gains first
        if (!p->saturated) DCBF_CUDA_TRY(cudaMalloc(&p->saturated, sizeof(unsigned long long)));
        DCBF_CUDA_TRY(cudaMemset(p->saturated, 0, sizeof(unsigned long long)));
    }
    const size_t samp_chan = static_cast<size_t>(p->T) * 4;                   // bytes per (b, a, c)
    const size_t beam_chan = static_cast<size_t>(p->T) * p->M * (q8 ? 2 : 8);  // bytes per (b, p, c)
    int i = 0;
    for (int c0 = 0; c0 < p->C; c0 += p->chunk, ++i) {
        auto& s = p->slots[i % p->slots.size()];
        const int cc = (p->C - c0 < p->chunk) ? p->C - c0 : p->chunk;
        // samples[b][a][c0:c0+cc] : B*A rows of cc*T*4 bytes, pitch C*T*4 -> dense [B][A][cc][T][4]
        DCBF_CUDA_TRY(cudaMemcpy2DAsync(s.samples, cc * samp_chan, h_samples + c0 * samp_chan, p->C * samp_chan,
                                        cc * samp_chan, static_cast<size_t>(p->B) * p->A, cudaMemcpyHostToDevice,
                                        s.stream));
        const float* dv_dev = s.delay_vals;
        if (h_dv)
            DCBF_CUDA_TRY(cudaMemcpyAsync(s.delay_vals, h_dv + static_cast<size_t>(c0) * p->M * p->A * 4,
                                          static_cast<size_t>(cc) * p->M * p->A * 16, cudaMemcpyHostToDevice, s.stream));
        else
            dv_dev = p->dv_resident[p->dv_active] + static_cast<size_t>(c0) * p->M * p->A * 4;
        const long long first_chan = static_cast<long long>(p->C) * p->xeng_id + c0;
        const QuantisedOut qo{reinterpret_cast<int8_t*>(s.beams), p->gains, p->saturated};
        if (int e = launch_fused(s.samples, dv_dev, s.beams, p->B, p->A, cc, p->N, p->T, p->M, first_chan,
                                 p->sample_period, nullptr, p->flags, s.stream, q8 ? &qo : nullptr, nullptr))
            return e;
        DCBF_CUDA_TRY(cudaMemcpy2DAsync(reinterpret_cast<uint8_t*>(h_beams) + c0 * beam_chan, p->C * beam_chan, s.beams,
                                        cc * beam_chan, cc * beam_chan, static_cast<size_t>(p->B) * kPols,
                                        cudaMemcpyDeviceToHost, s.stream));
    }
    for (auto& s : p->slots) DCBF_CUDA_TRY(cudaStreamSynchronize(s.stream));
    if (fused_status_poll() != DCBF_OK) return fused_status(nullptr, nullptr, nullptr);  // a watchdog fired: the beams are not valid
    if (q8 && h_saturated)
        DCBF_CUDA_TRY(cudaMemcpy(h_saturated, p->saturated, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return DCBF_OK;
}

}  // namespace dcbf

using namespace dcbf;

extern "C" {

__attribute__((visibility("default"))) int dcbf_host_plan_create(dcbf_host_plan_t* plan, int B, int A, int C, int N,
                                                                 int T, int M, int xeng_id, double sample_period,
                                                                 unsigned flags, int chunk_chans, int n_slots) {
    if (!plan || B <= 0 || A <= 0 || C <= 0 || N <= 0 || M <= 0 || T <= 0 || (T % kSamplesPerBlock) || xeng_id < 0 ||
        !(sample_period > 0.0))
        return DCBF_ERR_INVALID_ARG;
    if (chunk_chans <= 0 || chunk_chans > C) chunk_chans = C;
    if (n_slots <= 0) n_slots = 3;
    const int n_chunks = (C + chunk_chans - 1) / chunk_chans;
    if (n_slots > n_chunks) n_slots = n_chunks;
    auto* p = new HostPlan{B, A, C, N, T, M, xeng_id, sample_period, flags, chunk_chans, 0, {}};
    if (cudaGetDevice(&p->device) != cudaSuccess) {
        delete p;
        cudaGetLastError();
        return DCBF_ERR_NO_DEVICE;
    }
    p->slots.resize(n_slots);
    if (int e = create_plan(p)) {
        destroy_plan(p);
        return e;
    }
    *plan = p;
    return DCBF_OK;
}

__attribute__((visibility("default"))) int dcbf_host_plan_run(dcbf_host_plan_t plan, const uint8_t* samples,
                                                              const float* delay_vals, float* beams) {
    if (!plan || !samples || !beams) return DCBF_ERR_INVALID_ARG;  // delay_vals may be NULL: resident delay model
    return run_plan(static_cast<HostPlan*>(plan), samples, delay_vals, beams, false, nullptr);
}

__attribute__((visibility("default"))) int dcbf_host_plan_set_delay_vals(dcbf_host_plan_t plan, const float* delay_vals) {
    auto* p = static_cast<HostPlan*>(plan);
    if (!p || !delay_vals) return DCBF_ERR_INVALID_ARG;
    int dev = 0;
    DCBF_CUDA_TRY(cudaGetDevice(&dev));
    if (dev != p->device) return DCBF_ERR_INVALID_ARG;
    const size_t bytes = static_cast<size_t>(p->C) * p->M * p->A * 16;
    const int next = p->dv_active == 0 ? 1 : 0;  // never the copy the kernels of the last run read
    if (!p->dv_resident[next]) DCBF_CUDA_TRY(cudaMalloc(&p->dv_resident[next], bytes));
    cudaStream_t s = p->slots[0].stream;
    DCBF_CUDA_TRY(cudaMemcpyAsync(p->dv_resident[next], delay_vals, bytes, cudaMemcpyHostToDevice, s));
    DCBF_CUDA_TRY(cudaStreamSynchronize(s));
    p->dv_active = next;
    return DCBF_OK;
}

__attribute__((visibility("default"))) int dcbf_host_plan_set_gains(dcbf_host_plan_t plan, const float* beam_gains) {
    auto* p = static_cast<HostPlan*>(plan);
    if (!p || !beam_gains) return DCBF_ERR_INVALID_ARG;
    if (!p->gains) DCBF_CUDA_TRY(cudaMalloc(&p->gains, sizeof(float) * p->M));
    DCBF_CUDA_TRY(cudaMemcpy(p->gains, beam_gains, sizeof(float) * p->M, cudaMemcpyHostToDevice));
    return DCBF_OK;
}

__attribute__((visibility("default"))) int dcbf_host_plan_run_q8(dcbf_host_plan_t plan, const uint8_t* samples,
                                                                 const float* delay_vals, int8_t* beams_q8,
                                                                 unsigned long long* saturated) {
    if (!plan || !samples || !beams_q8) return DCBF_ERR_INVALID_ARG;
    return run_plan(static_cast<HostPlan*>(plan), samples, delay_vals, beams_q8, true, saturated);
}

__attribute__((visibility("default"))) int dcbf_host_plan_destroy(dcbf_host_plan_t plan) {
    return destroy_plan(static_cast<HostPlan*>(plan));
}

}  // extern "C"
