// Ingest side of the path: F-engine heaps -> the device layout [B][A][C][T][2][2] the beamformer consumes.
//
// An F-engine sends one SPEAD heap per (antenna, timestamp): item 0x1600 = timestamp (ADC samples), 0x4101 =
// feng_id, 0x4103 = first channel, 0x4300 = feng_raw int8 [n_chans][n_spectra_per_heap][n_pols][2]
// (reference: fgpu_send_prototype/fgpu_send_prototype.py:20-22,55-60; batches of heaps land as
// [batch][ant][chan][time][pol][complex], beamformer/README.md:5-6).  The engine that runs the beamformer has to
// gather the heaps of all antennas for `n_batches` consecutive timestamps into one contiguous chunk before it can
// launch.  This file is that assembly stage: a ring of page-locked chunks, heaps placed by (timestamp, feng_id) in
// any arrival order, chunks handed out in time order once complete -- or forced out, zero-filled and flagged, when
// newer data pushes the window forward (a lost packet must not stall the stream).
//
// No network code lives here (the reference's DPDK / ibverbs receivers are out of scope): the receive loop calls
// dcbf_ingest_heap (copy) or dcbf_ingest_heap_ptr (write in place) for every heap it completes.  The chunk returned
// by dcbf_ingest_pop is exactly the `samples` argument of dcbf_host_plan_run / the source of the H2D copy.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace dcbf {

struct Ingest {
    int n_chunks, window, B, A, C, T;  // window: chunks that may be open (receiving heaps) at the same time
    long long step;        // ADC samples between consecutive heaps of one antenna
    size_t heap_bytes;     // C * T * 4
    size_t chunk_bytes;    // B * A * heap_bytes
    bool pinned;
    struct Chunk {
        uint8_t* data = nullptr;
        long long index = -1;          // chunk number = timestamp / (B * step); -1: free
        std::vector<uint8_t> present;  // [B * A]
        // [B * A] byte ranges [begin, end) of each heap placed so far by dcbf_ingest_packet, sorted and disjoint: a
        // repeated or overlapping packet must not count twice (the heap would be reported complete while another
        // packet's bytes are still those of the slot's previous chunk)
        std::vector<std::vector<std::pair<uint32_t, uint32_t>>> received;
        int n_present = 0;
        bool handed_out = false;
    };
    std::vector<Chunk> chunks;
    std::deque<int> ready;             // chunk slots waiting for dcbf_ingest_pop, in time order
    long long newest = -1;             // highest chunk index seen
    unsigned long long n_late = 0, n_duplicate = 0, n_bad = 0;
    long long expect_frequency = -1;   // dcbf_ingest_packet: drop heaps of another sub-band (-1: accept any)
    // dcbf_ingest_packet: (timestamp, feng_id) of heaps whose later packets do not repeat the item pointers
    struct HeapKey {
        long long timestamp;
        int feng_id;
    };
    std::unordered_map<unsigned long long, HeapKey> heap_keys;
    std::deque<unsigned long long> heap_key_order;
    std::mutex mu;
};

static void free_ingest(Ingest* g) {
    if (!g) return;
    for (auto& c : g->chunks) {
        if (!c.data) continue;
        if (g->pinned)
            cudaFreeHost(c.data);
        else
            free(c.data);
    }
    if (g->pinned) cudaGetLastError();
    delete g;
}

// Marks a chunk complete-or-forced: zero-fills what never arrived and queues it for the consumer.
static void finish_chunk(Ingest* g, int slot) {
    auto& c = g->chunks[slot];
    if (c.n_present < g->B * g->A)
        for (int i = 0; i < g->B * g->A; ++i)
            if (!c.present[i]) memset(c.data + static_cast<size_t>(i) * g->heap_bytes, 0, g->heap_bytes);
    g->ready.push_back(slot);
}

// Slot holding chunk `index`, opening (and, if the ring is full, forcing out) chunks as needed.  -1: too old / no room.
static int slot_for(Ingest* g, long long index) {
    for (int s = 0; s < g->n_chunks; ++s)
        if (g->chunks[s].index == index) return g->chunks[s].handed_out || g->chunks[s].n_present < 0 ? -1 : s;
    if (index <= g->newest - g->window) return -1;  // older than the window
    // force out every open chunk that falls behind the new window, oldest first
    if (index > g->newest) {
        for (;;) {
            int oldest = -1;
            for (int s = 0; s < g->n_chunks; ++s) {
                auto& c = g->chunks[s];
                if (c.index >= 0 && c.n_present >= 0 && c.index <= index - g->window &&
                    (oldest < 0 || c.index < g->chunks[oldest].index))
                    oldest = s;
            }
            if (oldest < 0) break;
            finish_chunk(g, oldest);
            g->chunks[oldest].n_present = -1 - g->chunks[oldest].n_present;  // closed: late heaps are refused
        }
        g->newest = index;
    }
    for (int s = 0; s < g->n_chunks; ++s) {
        auto& c = g->chunks[s];
        if (c.index < 0) {
            c.index = index;
            c.n_present = 0;
            c.handed_out = false;
            std::fill(c.present.begin(), c.present.end(), 0);
            for (auto& r : c.received) r.clear();
            return s;
        }
    }
    return -1;  // every slot is queued or with the consumer: back-pressure
}

static int locate(Ingest* g, long long timestamp, int feng_id, int* slot, int* cell) {
    if (timestamp < 0 || timestamp % g->step || feng_id < 0 || feng_id >= g->A) {
        ++g->n_bad;
        return DCBF_ERR_INVALID_ARG;
    }
    const long long heap = timestamp / g->step;
    const long long index = heap / g->B;
    const int s = slot_for(g, index);
    if (s < 0) {
        ++g->n_late;
        return DCBF_ERR_UNSUPPORTED;  // dropped (too old, or no free chunk)
    }
    *slot = s;
    *cell = static_cast<int>(heap % g->B) * g->A + feng_id;
    return DCBF_OK;
}

static void mark_present(Ingest* g, int slot, int cell) {
    auto& c = g->chunks[slot];
    if (c.present[cell]) {
        ++g->n_duplicate;
        return;
    }
    c.present[cell] = 1;
    if (++c.n_present == g->B * g->A) {
        finish_chunk(g, slot);
        c.n_present = -1 - c.n_present;  // closed
    }
}

}  // namespace dcbf

using namespace dcbf;

#pragma GCC visibility push(default)
extern "C" {

int dcbf_ingest_create(dcbf_ingest_t* ingest, int n_chunks, int n_batches, int n_ants, int n_chans, int n_samples,
                       long long timestamp_step, int pinned) {
    if (!ingest || n_chunks < 2 || n_batches <= 0 || n_ants <= 0 || n_chans <= 0 || n_samples <= 0 ||
        n_samples % kSamplesPerBlock || timestamp_step <= 0)
        return DCBF_ERR_INVALID_ARG;
    auto* g = new Ingest{};
    g->n_chunks = n_chunks, g->B = n_batches, g->A = n_ants, g->C = n_chans, g->T = n_samples;
    g->window = n_chunks / 2;  // the other slots hold finished chunks until the consumer has released them
    g->step = timestamp_step;
    g->heap_bytes = static_cast<size_t>(n_chans) * n_samples * 4;
    g->chunk_bytes = g->heap_bytes * n_batches * n_ants;
    g->pinned = pinned != 0;
    g->chunks.resize(n_chunks);
    for (auto& c : g->chunks) {
        if (g->pinned) {
            if (cudaHostAlloc(reinterpret_cast<void**>(&c.data), g->chunk_bytes, cudaHostAllocDefault) != cudaSuccess) {
                const int e = record_cuda_error(cudaGetLastError(), "cudaHostAlloc(ingest chunk)");
                c.data = nullptr;
                free_ingest(g);
                return e;
            }
        } else if (posix_memalign(reinterpret_cast<void**>(&c.data), 4096, g->chunk_bytes) != 0) {
            c.data = nullptr;
            free_ingest(g);
            return DCBF_ERR_UNSUPPORTED;
        }
        c.present.assign(static_cast<size_t>(n_batches) * n_ants, 0);
        c.received.assign(static_cast<size_t>(n_batches) * n_ants, {});
    }
    *ingest = g;
    return DCBF_OK;
}

int dcbf_ingest_heap(dcbf_ingest_t ingest, long long timestamp, int feng_id, const void* payload) {
    auto* g = static_cast<Ingest*>(ingest);
    if (!g || !payload) return DCBF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->mu);
    int slot = 0, cell = 0;
    if (int e = locate(g, timestamp, feng_id, &slot, &cell)) return e;
    if (!g->chunks[slot].present[cell])
        memcpy(g->chunks[slot].data + static_cast<size_t>(cell) * g->heap_bytes, payload, g->heap_bytes);
    mark_present(g, slot, cell);
    return DCBF_OK;
}

int dcbf_ingest_heap_ptr(dcbf_ingest_t ingest, long long timestamp, int feng_id, void** dst) {
    auto* g = static_cast<Ingest*>(ingest);
    if (!g || !dst) return DCBF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->mu);
    int slot = 0, cell = 0;
    if (int e = locate(g, timestamp, feng_id, &slot, &cell)) return e;
    *dst = g->chunks[slot].data + static_cast<size_t>(cell) * g->heap_bytes;
    return DCBF_OK;
}

int dcbf_ingest_heap_done(dcbf_ingest_t ingest, long long timestamp, int feng_id) {
    auto* g = static_cast<Ingest*>(ingest);
    if (!g) return DCBF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->mu);
    int slot = 0, cell = 0;
    if (int e = locate(g, timestamp, feng_id, &slot, &cell)) return e;
    mark_present(g, slot, cell);
    return DCBF_OK;
}

// SPEAD-64-48 (flavour 4, 64, 48: what MeerKAT F-engines and fgpu_send_prototype.py:18 emit).  Packet = 8-byte header
// {0x53, 0x04, item-pointer id bytes = 2, heap-address bytes = 6, reserved u16, n_items u16}, n_items big-endian
// 64-bit item pointers {immediate flag : 1, id : 15, value or payload address : 48}, then the payload bytes.
int dcbf_ingest_packet(dcbf_ingest_t ingest, const void* packet, size_t length, int default_feng_id) {
    auto* g = static_cast<Ingest*>(ingest);
    if (!g || !packet) return DCBF_ERR_INVALID_ARG;
    const auto* p = static_cast<const uint8_t*>(packet);
    std::lock_guard<std::mutex> lock(g->mu);
    auto bad = [&]() {
        ++g->n_bad;
        return DCBF_ERR_INVALID_ARG;
    };
    if (length < 8 || p[0] != 0x53 || p[1] != 0x04 || p[2] != 2 || p[3] != 6) return bad();
    const size_t n_items = (static_cast<size_t>(p[6]) << 8) | p[7];
    if (length < 8 + 8 * n_items) return bad();
    long long heap_cnt = -1, heap_size = -1, heap_offset = -1, payload_len = -1, timestamp = -1, feng_id = -1,
              frequency = -1, raw_addr = -1;
    bool descriptor = false;
    for (size_t i = 0; i < n_items; ++i) {
        unsigned long long w = 0;
        for (int b = 0; b < 8; ++b) w = (w << 8) | p[8 + 8 * i + b];
        const bool immediate = (w >> 63) != 0;
        const unsigned id = static_cast<unsigned>((w >> 48) & 0x7fffu);
        const long long v = static_cast<long long>(w & 0xffffffffffffull);
        switch (id) {
            case 0x0001: heap_cnt = v; break;
            case 0x0002: heap_size = v; break;
            case 0x0003: heap_offset = v; break;
            case 0x0004: payload_len = v; break;
            case 0x0005: descriptor = true; break;
            case 0x1600: if (immediate) timestamp = v; break;
            case 0x4101: if (immediate) feng_id = v; break;
            case 0x4103: if (immediate) frequency = v; break;
            case 0x4300: if (!immediate) raw_addr = v; break;
            default: break;  // other items (stream control, engine-specific immediates) do not concern the layout
        }
    }
    const size_t header = 8 + 8 * n_items;
    if (heap_cnt < 0 || heap_offset < 0 || payload_len < 0 || static_cast<size_t>(payload_len) != length - header) return bad();
    if (descriptor) return DCBF_ERR_UNSUPPORTED;  // descriptor heaps carry no data
    if (timestamp >= 0) {  // first packet of a heap (or a sender that repeats the pointers): remember the heap's place
        if (feng_id < 0) feng_id = default_feng_id;
        if (feng_id < 0 || feng_id >= g->A) return bad();  // 48-bit field: range-check before it is narrowed to int
        if (g->heap_keys.find(static_cast<unsigned long long>(heap_cnt)) == g->heap_keys.end()) {
            g->heap_key_order.push_back(static_cast<unsigned long long>(heap_cnt));
            if (g->heap_key_order.size() > 4096) {
                g->heap_keys.erase(g->heap_key_order.front());
                g->heap_key_order.pop_front();
            }
        }
        g->heap_keys[static_cast<unsigned long long>(heap_cnt)] = {timestamp, static_cast<int>(feng_id)};
        if (g->expect_frequency >= 0 && frequency >= 0 && frequency != g->expect_frequency) {
            g->heap_keys[static_cast<unsigned long long>(heap_cnt)].feng_id = -2;  // another engine's sub-band
            return DCBF_ERR_UNSUPPORTED;
        }
    } else {
        const auto it = g->heap_keys.find(static_cast<unsigned long long>(heap_cnt));
        if (it == g->heap_keys.end()) {  // overtook its heap's first packet: cannot be placed
            ++g->n_late;
            return DCBF_ERR_UNSUPPORTED;
        }
        timestamp = it->second.timestamp;
        feng_id = it->second.feng_id;
        if (feng_id == -2) return DCBF_ERR_UNSUPPORTED;
    }
    if (raw_addr < 0) raw_addr = 0;  // the payload of an F-engine heap is the feng_raw item alone
    if ((heap_size >= 0 && static_cast<size_t>(heap_size) != g->heap_bytes) || raw_addr != 0 ||
        static_cast<size_t>(heap_offset) + static_cast<size_t>(payload_len) > g->heap_bytes)
        return bad();
    int slot = 0, cell = 0;
    if (int e = locate(g, timestamp, static_cast<int>(feng_id), &slot, &cell)) return e;
    auto& c = g->chunks[slot];
    if (c.present[cell]) {
        ++g->n_duplicate;
        return DCBF_OK;
    }
    // merge [heap_offset, heap_offset + payload_len) into the heap's received ranges
    auto& ranges = c.received[cell];
    uint32_t lo = static_cast<uint32_t>(heap_offset), hi = lo + static_cast<uint32_t>(payload_len);
    bool covered = false;
    std::vector<std::pair<uint32_t, uint32_t>> merged;
    merged.reserve(ranges.size() + 1);
    for (const auto& r : ranges) {
        if (r.first <= lo && hi <= r.second) covered = true;
        if (r.second < lo || hi < r.first) {
            merged.push_back(r);  // disjoint and not adjacent
        } else {
            lo = std::min(lo, r.first);
            hi = std::max(hi, r.second);
        }
    }
    if (covered && payload_len > 0) {  // a retransmission of bytes that are already in place
        ++g->n_duplicate;
        return DCBF_OK;
    }
    memcpy(c.data + static_cast<size_t>(cell) * g->heap_bytes + heap_offset, p + header, static_cast<size_t>(payload_len));
    merged.emplace_back(lo, hi);
    std::sort(merged.begin(), merged.end());
    ranges.swap(merged);
    if (ranges.size() == 1 && ranges[0].first == 0 && ranges[0].second >= g->heap_bytes) mark_present(g, slot, cell);
    return DCBF_OK;
}

int dcbf_ingest_set_frequency(dcbf_ingest_t ingest, long long first_channel) {
    auto* g = static_cast<Ingest*>(ingest);
    if (!g) return DCBF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->mu);
    g->expect_frequency = first_channel;
    return DCBF_OK;
}

int dcbf_ingest_pop(dcbf_ingest_t ingest, int flush, const uint8_t** samples, long long* first_timestamp,
                    int* n_missing, uint8_t* present) {
    auto* g = static_cast<Ingest*>(ingest);
    if (!g || !samples) return DCBF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->mu);
    if (g->ready.empty() && flush) {  // end of stream: push out the oldest open chunk as it is
        int oldest = -1;
        for (int s = 0; s < g->n_chunks; ++s) {
            auto& c = g->chunks[s];
            if (c.index >= 0 && c.n_present >= 0 && !c.handed_out && (oldest < 0 || c.index < g->chunks[oldest].index))
                oldest = s;
        }
        if (oldest >= 0) {
            finish_chunk(g, oldest);
            g->chunks[oldest].n_present = -1 - g->chunks[oldest].n_present;
        }
    }
    if (g->ready.empty()) return 0;
    // strictly in time order: the oldest finished chunk, and only if no older chunk is still receiving heaps
    size_t pick = 0;
    for (size_t i = 1; i < g->ready.size(); ++i)
        if (g->chunks[g->ready[i]].index < g->chunks[g->ready[pick]].index) pick = i;
    const int slot = g->ready[pick];
    if (!flush)
        for (const auto& o : g->chunks)
            if (o.index >= 0 && o.n_present >= 0 && o.index < g->chunks[slot].index) return 0;
    g->ready.erase(g->ready.begin() + static_cast<long>(pick));
    auto& c = g->chunks[slot];
    c.handed_out = true;
    *samples = c.data;
    if (first_timestamp) *first_timestamp = c.index * g->B * g->step;
    const int have = -1 - c.n_present;
    if (n_missing) *n_missing = g->B * g->A - have;
    if (present) memcpy(present, c.present.data(), c.present.size());
    return 1;
}

int dcbf_ingest_release(dcbf_ingest_t ingest, const uint8_t* samples) {
    auto* g = static_cast<Ingest*>(ingest);
    if (!g || !samples) return DCBF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->mu);
    for (auto& c : g->chunks)
        if (c.data == samples && c.handed_out) {
            c.handed_out = false;
            c.index = -1;
            c.n_present = 0;
            return DCBF_OK;
        }
    return DCBF_ERR_INVALID_ARG;
}

int dcbf_ingest_stats(dcbf_ingest_t ingest, unsigned long long* n_late, unsigned long long* n_duplicate,
                      unsigned long long* n_bad) {
    auto* g = static_cast<Ingest*>(ingest);
    if (!g) return DCBF_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g->mu);
    if (n_late) *n_late = g->n_late;
    if (n_duplicate) *n_duplicate = g->n_duplicate;
    if (n_bad) *n_bad = g->n_bad;
    return DCBF_OK;
}

int dcbf_ingest_destroy(dcbf_ingest_t ingest) {
    free_ingest(static_cast<Ingest*>(ingest));
    return DCBF_OK;
}

}  // extern "C"
#pragma GCC visibility pop
