// capi.cu — the extern "C" boundary of librtbvh_rs.so: scenes and the batch traversal entry points of
// include/rtbvh_gpu.h.  (The legacy rtbvh_ffi entry points of include/rtbvh.h live in legacy.cu.)
//
// Ownership model mirrors rtbvh_ffi's StructureManager (rtbvh_ffi/src/lib.rs:12-127): a process-global
// table guarded by a reader/writer lock, ids never reused.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "build.cuh"
#include "traverse.cuh"

using namespace rtb;

namespace rtb {
thread_local std::string g_last_error;

ResultCode fail(const char* what, cudaError_t e) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return Error;
}
ResultCode fail(const char* what) {
    g_last_error = what;
    return Error;
}
}  // namespace rtb

#define RTB_CUDA(call)                                  \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return fail(#call, e__); \
    } while (0)

namespace {

constexpr int kTickets = 64;                     // outstanding asynchronous submissions per scene
constexpr int kPipeStreams = 4;                  // H2D / kernel / D2H of consecutive chunks overlap
constexpr size_t kChunkRays = size_t(1) << 21;   // initial capacity of one staging slot: 2 Mi rays (64 MiB of RTRay)
constexpr size_t kMaxSlotRays = size_t(1) << 23; // the gated flavour grows the slots up to 8 Mi rays: one launch per 8 M-ray batch
constexpr uint32_t kCounterSlots = 1024;
constexpr size_t kSubChunkRays = size_t(1) << 19;  // gated pipeline: watermark granularity (measured: 256 Ki 1.48, 512 Ki 1.52, 1 Mi 1.38 Grays/s)
constexpr size_t kMinChunkRays = size_t(1) << 17;  // gated pipeline: smallest launch (the tail of a batch ramps down to this)
constexpr int kReadySlots = 64;
constexpr int kMarkSlots = 4096;

// RTBVH_TRACE_MODE selects the single-ray kernel variant (A/B measurements).  Measured on B200, config 2 (1 Mi-triangle
// soup, 8 M primary rays per launch, scripts/trace_ab.py, profiles/r5_trace_ab.md): persistent (one node visit + its leaf
// slots per iteration) 1 887-1 889 Mrays/s; phased (warp-wide node / triangle phases) 1 780 with the majority rule, 1 862-1 877
// with a triangle phase as soon as a third as many lanes wait for it; staged tree top (RTB_TOPK = 21 / 85 / 341 nodes in shared
// memory) 1 844-1 860; software-pipelined node fetch 1 460.  Configs 4 / 5 (incoherent): 390 / 94 persistent, 393 / 90 phased.
#ifndef RTB_DEFAULT_TRACE_MODE
#define RTB_DEFAULT_TRACE_MODE kTracePersistent
#endif
constexpr int kDefaultTraceMode = RTB_DEFAULT_TRACE_MODE;
int persistent_mode() {
    static const int v = [] {
        const char* e = std::getenv("RTBVH_TRACE_MODE");
        if (!e) return kDefaultTraceMode;
        const std::string m(e);
        if (m == "static") return (int)kTraceStatic;
        if (m == "persistent") return (int)kTracePersistent;
        if (m == "phased") return (int)kTracePhased;
        if (m == "coop") return (int)kTraceCoop;
        return kDefaultTraceMode;
    }();
    return v;
}

// the refill-kernel flavour for calls that need one (input gate, split input, fused gather)
int refill_mode() {
    const int m = persistent_mode();
    return (m == kTracePersistent || m == kTracePhased) ? m : kDefaultTraceMode;
}

struct Scene {
    int device = 0;
    DeviceTree bvh{nullptr, 0, nullptr, 0, nullptr, 0};
    DeviceTree mbvh{nullptr, 0, nullptr, 0, nullptr, 0};
    void* d_bvh_nodes = nullptr;
    void* d_mbvh_nodes = nullptr;
    TriRec* d_tris_bvh = nullptr;   // leaf order of the Bvh's indices
    TriRec* d_tris_mbvh = nullptr;  // leaf order of the Mbvh's indices (may alias d_tris_bvh)
    uint32_t* d_idx_bvh = nullptr;   // prim_indices of the trees (kept for refit: re-gathering the triangle records)
    uint32_t* d_idx_mbvh = nullptr;  // may alias d_idx_bvh
    uint32_t tri_count = 0;
    float4* d_top = nullptr;          // staged-top table of the Mbvh (DeviceTree::top), rebuilt after every refit
    uint32_t* d_top_count = nullptr;
    ResidentRefit refit_cache;       // topology analysis + scratch of rtbvh_gpu_scene_refit*
    float* d_refit_verts = nullptr;  // staging of the host-buffer refit call
    size_t refit_verts_bytes = 0;
    std::mutex refit_mutex;
    uint32_t* d_overflow = nullptr;
    float bounds[6] = {0, 0, 0, 1, 1, 1};  // root box (keys of the optional ray sort)
    std::atomic<int> sort_rays{0};
    std::atomic<uint32_t> tile_w{0};  // rtbvh_gpu_scene_set_ray_tiling: row length of image-ordered batches (0 = off)
    // work-order hint for the device-resident single-ray calls: 8x8 pixel tiles over the whole 8-row bands of the batch
    void apply_tiling(PeerDests& pd, size_t n) const {
        const uint32_t w = tile_w.load();
        if (w == 0 || sort_rays.load()) return;
        const unsigned long long band = 8ull * w;
        pd.tile_w = w;
        pd.tile_n = (unsigned long long)n / band * band;
        if (pd.tile_n == 0) pd.tile_w = 0;
    }
    const float* sort_bounds() const { return sort_rays.load() ? bounds : nullptr; }
    // work counters of the persistent kernels: every launch takes the next slot and zeroes it on its
    // own stream, so launches on different streams never share a counter
    unsigned long long* d_counters = nullptr;
    std::atomic<uint32_t> next_counter{0};
    unsigned long long* counter_slot() { return d_counters + (next_counter.fetch_add(1) % kCounterSlots); }
    // host-buffer pipeline (lazily created)
    cudaStream_t streams[kPipeStreams] = {};
    void* d_in[kPipeStreams] = {};
    void* d_out[kPipeStreams] = {};
    std::mutex pipe_mutex;
    size_t slot_rays = 0;    // capacity of one staging slot
    uint64_t chunk_seq = 0;  // staging slot rotation across batches
    uint64_t gated_seq = 0;  // watermark slot rotation (gated flavour)
    int marks_used = 0;      // pinned watermark source values in flight (memcpy fallback of the gated flavour)
    struct Ticket {
        uint64_t id = 0;
        cudaEvent_t ev[kPipeStreams] = {};
    };
    Ticket tickets[kTickets];
    uint64_t next_ticket = 1;
    // gated pipeline (single rays): one copy stream feeds launches that start before their input has arrived
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_start[kPipeStreams] = {}, ev_done[kPipeStreams] = {};
    cudaEvent_t ev_refit = nullptr;          // scratch event: orders a scene refit against the pipeline streams
    cudaEvent_t ev_refit_done = nullptr;     // end of the latest refit (refits share one scratch: they are chained)
    unsigned long long* d_ready = nullptr;   // kReadySlots watermarks (rays delivered per launch)
    unsigned long long* h_marks = nullptr;   // pinned source values of the watermark copies

    ~Scene() {
        int prev_device = -1;
        cudaGetDevice(&prev_device);
        cudaSetDevice(device);
        for (int i = 0; i < kPipeStreams; i++) {
            if (streams[i]) cudaStreamDestroy(streams[i]);
            cudaFree(d_in[i]);
            cudaFree(d_out[i]);
            if (ev_start[i]) cudaEventDestroy(ev_start[i]);
            if (ev_done[i]) cudaEventDestroy(ev_done[i]);
        }
        for (auto& t : tickets)
            for (auto e : t.ev)
                if (e) cudaEventDestroy(e);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (ev_refit) cudaEventDestroy(ev_refit);
        if (ev_refit_done) cudaEventDestroy(ev_refit_done);
        cudaFree(d_ready);
        if (h_marks) cudaFreeHost(h_marks);
        dev_block_free(d_bvh_nodes);  // the large blocks go back to the scene block cache (build.cuh)
        dev_block_free(d_mbvh_nodes);
        if (d_tris_mbvh != d_tris_bvh) dev_block_free(d_tris_mbvh);
        dev_block_free(d_tris_bvh);
        if (d_idx_mbvh != d_idx_bvh) dev_block_free(d_idx_mbvh);
        dev_block_free(d_idx_bvh);
        dev_block_free(d_refit_verts);
        cudaFree(d_top);
        cudaFree(d_top_count);
        cudaFree(d_overflow);
        cudaFree(d_counters);
        if (prev_device >= 0 && prev_device != device) cudaSetDevice(prev_device);
    }
};

// Makes `device` current for the scope and restores the caller's device afterwards (the C ABI must not leave the
// calling thread on another device).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;  // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    bool ok() const { return err == cudaSuccess; }
};

struct SceneTable {
    std::shared_mutex mu;
    std::vector<std::shared_ptr<Scene>> scenes;  // index = handle - 1; freed entries become null
} g_scenes;

std::shared_ptr<Scene> get_scene(RTGpuScene h) {
    std::shared_lock<std::shared_mutex> lk(g_scenes.mu);
    if (h == 0 || h > g_scenes.scenes.size()) return nullptr;
    return g_scenes.scenes[h - 1];
}

// Device-pointer entry points launch on the caller's stream, i.e. on the caller's current device: the scene must live there.
bool on_scene_device(const Scene& s) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return cur == s.device;
}

const DeviceTree* pick_tree(const Scene& s, RTTreeKind kind) {
    const DeviceTree* t = kind == RT_TREE_MBVH ? &s.mbvh : (kind == RT_TREE_BVH ? &s.bvh : nullptr);
    if (!t || !t->nodes) return nullptr;
    return t;
}

// Staging slots hold `slot_rays` rays each (RTRay records, or origins at offset 0 and directions at slot_rays * 12; a packet
// chunk is slot_rays / 4 * 112 B < this).  want_rays > slot_rays: drain the pipeline and grow the slots.
ResultCode ensure_pipeline(Scene& s, size_t want_rays = 0) {
    if (s.streams[0] && want_rays > s.slot_rays) {
        // grow: the new slots are allocated first and committed only when every allocation has succeeded, so a failed
        // cudaMalloc leaves the old (smaller, still valid) pipeline in place
        void* nin[kPipeStreams] = {};
        void* nout[kPipeStreams] = {};
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < kPipeStreams && e == cudaSuccess; i++) {
            e = cudaMalloc(&nin[i], want_rays * sizeof(RTRay));
            if (e == cudaSuccess) e = cudaMalloc(&nout[i], want_rays * sizeof(RTHit));
        }
        if (e == cudaSuccess) {
            for (int i = 0; i < kPipeStreams && e == cudaSuccess; i++) e = cudaStreamSynchronize(s.streams[i]);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s.copy_stream);
        }
        if (e != cudaSuccess) {
            for (int i = 0; i < kPipeStreams; i++) {
                cudaFree(nin[i]);
                cudaFree(nout[i]);
            }
            return fail("host-buffer pipeline: growing the staging slots failed", e);
        }
        for (int i = 0; i < kPipeStreams; i++) {
            cudaFree(s.d_in[i]);
            cudaFree(s.d_out[i]);
            s.d_in[i] = nin[i];
            s.d_out[i] = nout[i];
        }
        s.slot_rays = want_rays;
        return Ok;
    }
    if (s.streams[0]) return Ok;
    // first use: if anything fails the half-built pipeline is torn down again (streams[0] stays null: the next call retries)
    const size_t rays = std::max(kChunkRays, want_rays);
    auto teardown = [&]() {
        for (int i = 0; i < kPipeStreams; i++) {
            if (s.streams[i]) cudaStreamDestroy(s.streams[i]);
            s.streams[i] = nullptr;
            cudaFree(s.d_in[i]);
            cudaFree(s.d_out[i]);
            s.d_in[i] = s.d_out[i] = nullptr;
            if (s.ev_start[i]) cudaEventDestroy(s.ev_start[i]);
            if (s.ev_done[i]) cudaEventDestroy(s.ev_done[i]);
            s.ev_start[i] = s.ev_done[i] = nullptr;
        }
        if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
        s.copy_stream = nullptr;
        cudaFree(s.d_ready);
        s.d_ready = nullptr;
        if (s.h_marks) cudaFreeHost(s.h_marks);
        s.h_marks = nullptr;
        s.slot_rays = 0;
    };
    cudaError_t e = cudaSuccess;
    cudaStream_t st[kPipeStreams] = {};
    for (int i = 0; i < kPipeStreams && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_in[i], rays * sizeof(RTRay));
        if (e == cudaSuccess) e = cudaMalloc(&s.d_out[i], rays * sizeof(RTHit));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.ev_start[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.ev_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&s.d_ready, kReadySlots * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaHostAlloc(&s.h_marks, kMarkSlots * sizeof(unsigned long long), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        for (int i = 0; i < kPipeStreams; i++) {
            s.streams[i] = st[i];  // so that teardown destroys them
        }
        teardown();
        return fail("host-buffer pipeline: set-up failed", e);
    }
    for (int i = 0; i < kPipeStreams; i++) s.streams[i] = st[i];  // committed last: streams[0] != null <=> pipeline complete
    s.slot_rays = rays;
    return Ok;
}

// slot capacity the gated flavour wants for a batch of `units` rays: every launch ends with a drain phase, so the fewer
// launches per batch the better — the slots grow with the batch, up to 8 Mi rays
size_t gated_slot_rays(size_t units) {
    size_t want = kChunkRays;
    while (want < units && want < kMaxSlotRays) want <<= 1;
    return want;
}

// Builds (or, after a refit, refreshes) the staged-top table of the scene's Mbvh.  `read_count`: also fetch the slot count
// (scene creation; a refit keeps the topology and with it the count).
ResultCode scene_build_top(Scene& s, cudaStream_t st, bool read_count) {
    const int cap = top_table_capacity();
    if (cap == 0 || !s.d_mbvh_nodes || s.mbvh.node_count == 0) return Ok;
    if (!s.d_top) {
        RTB_CUDA(cudaMalloc((void**)&s.d_top, (size_t)cap * 128));
        RTB_CUDA(cudaMalloc((void**)&s.d_top_count, sizeof(uint32_t)));
    }
    RTB_CUDA(launch_build_top_table((const float4*)s.d_mbvh_nodes, s.mbvh.node_count, s.d_top, s.d_top_count, st));
    if (read_count) {
        uint32_t cnt = 0;
        RTB_CUDA(cudaMemcpyAsync(&cnt, s.d_top_count, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        RTB_CUDA(cudaStreamSynchronize(st));
        s.mbvh.top = s.d_top;
        s.mbvh.top_count = cnt;
    }
    return Ok;
}

enum HostMode { kHostAuto = 0, kHostStaged = 1, kHostGated = 2 };
// Measured on B200, 8 M rays per step, two steps in flight (profiles/r2_host_pipeline.md), staged (16 chunks over 4 streams)
// vs gated (one launch per batch that starts before its input has arrived): RTRay records 1.37 vs 1.51 Grays/s, split
// origin / direction input 1.39 vs 1.74-1.83; blocking calls 1.35 vs 1.36.  Auto = gated wherever the call allows it (single
// rays, caller's order, default kernel); RTBVH_HOST_MODE=staged|gated forces one.
// Rejected: kernels reading pinned host rays directly over PCIe 1.11; reading rays and writing records directly 0.74.
int host_mode() {
    static const int v = [] {
        const char* e = std::getenv("RTBVH_HOST_MODE");
        if (!e) return (int)kHostAuto;
        const std::string m(e);
        if (m == "staged") return (int)kHostStaged;
        if (m == "gated") return (int)kHostGated;
        return (int)kHostAuto;
    }();
    return v;
}

// Reads (and clears) the scene's stack-overflow flag once the pipeline streams have drained.
ResultCode check_overflow(Scene& s) {
    uint32_t ovf = 0;
    RTB_CUDA(cudaMemcpy(&ovf, s.d_overflow, sizeof(ovf), cudaMemcpyDeviceToHost));
    if (ovf) {
        RTB_CUDA(cudaMemset(s.d_overflow, 0, sizeof(uint32_t)));
        return fail("traversal stack overflow (> 128 entries)");
    }
    return Ok;
}

// Enqueues one host-buffer batch: chunks flow through kPipeStreams streams, each doing H2D -> kernel -> D2H into its
// own staging slot, so the copy engines and the SMs overlap across chunks — and across consecutive batches when the
// caller does not wait in between (the *_async entry points).  The caller holds s.pipe_mutex.
// unit_in / unit_out are bytes per ray (single) or per packet; rays_per_unit is 1 or 4.
// `in2` (optional): a second input array of unit_in2 bytes per unit, staged slot_rays * 12 bytes into the slot (the split
// origin / direction format).
template <class Launch>
ResultCode enqueue_host_batch(Scene& s, const void* in, size_t units, size_t unit_in, size_t unit_out, size_t rays_per_unit,
                              void* out, Launch&& launch, const void* in2 = nullptr, size_t unit_in2 = 0) {
    // Chunk size: a synchronous call drains its pipeline before returning, so the first H2D and the last kernel + D2H
    // are not overlapped with anything; many small chunks keep that fill/drain cost low (a batch is cut into >= 16
    // chunks), a floor of 256 Ki rays keeps every launch big enough to fill the machine.  RTBVH_CHUNK_RAYS overrides.
    static const size_t forced = [] {
        const char* e = std::getenv("RTBVH_CHUNK_RAYS");
        return e ? (size_t)std::strtoull(e, nullptr, 10) : size_t(0);
    }();
    size_t chunk_rays = forced ? forced : std::max<size_t>(size_t(1) << 18, (units * rays_per_unit + 15) / 16);
    chunk_rays = std::min(chunk_rays, kChunkRays);
    const size_t chunk_units = std::max<size_t>(1, chunk_rays / rays_per_unit);
    size_t done = 0;
    static const bool trace = std::getenv("RTBVH_PIPE_TRACE") != nullptr;  // debug: per-chunk timeline on stderr
    std::vector<cudaEvent_t> ev;
    if (trace) {
        ev.resize(1);
        cudaEventCreate(&ev[0]);
        cudaEventRecord(ev[0], s.streams[0]);
    }
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
    };
    while (done < units) {
        const int k = (int)(s.chunk_seq++ % kPipeStreams);  // slots rotate across batches
        const size_t m = units - done < chunk_units ? units - done : chunk_units;
        cudaStream_t st = s.streams[k];
        mark(st);
        RTB_CUDA(cudaMemcpyAsync(s.d_in[k], (const char*)in + done * unit_in, m * unit_in, cudaMemcpyHostToDevice, st));
        if (in2)
            RTB_CUDA(cudaMemcpyAsync((char*)s.d_in[k] + s.slot_rays * 12, (const char*)in2 + done * unit_in2, m * unit_in2,
                                     cudaMemcpyHostToDevice, st));
        mark(st);
        RTB_CUDA(launch(s.d_in[k], m, s.d_out[k], (const unsigned long long*)nullptr, st));
        mark(st);
        RTB_CUDA(cudaMemcpyAsync((char*)out + done * unit_out, s.d_out[k], m * unit_out, cudaMemcpyDeviceToHost, st));
        mark(st);
        done += m;
    }
    if (trace) {
        for (int i = 0; i < kPipeStreams; i++) cudaStreamSynchronize(s.streams[i]);
        for (size_t c = 0; 1 + 4 * c + 3 < ev.size(); c++) {
            float t[4];
            for (int j = 0; j < 4; j++) cudaEventElapsedTime(&t[j], ev[0], ev[1 + 4 * c + j]);
            std::fprintf(stderr, "chunk %2zu: h2d %.3f-%.3f  kernel -%.3f  d2h -%.3f ms\n", c, t[0], t[1], t[2], t[3]);
        }
        for (auto e : ev) cudaEventDestroy(e);
    }
    return Ok;
}

// Gated flavour of enqueue_host_batch for single rays (RTBVH_HOST_MODE=gated): the batch is cut into a few large launches;
// ONE copy stream uploads the rays back to back in 8 MiB pieces, each followed by a watermark write, and every launch starts
// as soon as its staging slot is free — its warps wait on the watermark for ranges that have not arrived
// (PeerDests::ready).  The copy engine never waits for a kernel boundary and a launch is never smaller than its slot, so the
// per-launch drain (warps that can no longer refill) is paid once per 2 Mi rays instead of once per 0.5 Mi.
// `ramp`: a blocking call shrinks the last launches (down to 128 Ki rays) so that little is left to trace after the last
// byte has crossed PCIe; a stream of asynchronous batches keeps them large.  The caller holds s.pipe_mutex.
template <class Launch>
ResultCode enqueue_host_batch_gated(Scene& s, const void* in, size_t n, size_t unit_in, size_t unit_out, void* out, bool ramp,
                                    Launch&& launch, const void* in2) {
    cudaStream_t cp = s.copy_stream;
    // watermark writes: cuStreamWriteValue64 when the driver exports it (fetched at run time: no link-time libcuda
    // dependency), else an 8-byte H2D copy from a pinned table
    typedef int (*WriteValue64)(cudaStream_t, unsigned long long, unsigned long long, unsigned int);
    static const WriteValue64 write_value = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (std::getenv("RTBVH_GATE_MEMCPY") != nullptr) return (WriteValue64) nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return (WriteValue64) nullptr;
        }
        return (WriteValue64)fn;
    }();
    static const size_t sub_rays = [] {
        const char* e = std::getenv("RTBVH_SUBCHUNK_RAYS");
        const size_t v = e ? (size_t)std::strtoull(e, nullptr, 10) : 0;
        return v ? v : kSubChunkRays;
    }();
    size_t done = 0;
    while (done < n) {
        const uint64_t seq = s.gated_seq++;
        const int k = (int)(s.chunk_seq++ % kPipeStreams);
        const size_t left = n - done;
        size_t m = std::min(s.slot_rays, left);
        if (ramp) {
            m = std::min(s.slot_rays, std::max(kMinChunkRays, ((left / 2 + kMinChunkRays - 1) / kMinChunkRays) * kMinChunkRays));
            if (m > left || left - m < kMinChunkRays / 2) m = std::min(left, s.slot_rays);
        }
        unsigned long long* ready = s.d_ready + (seq % kReadySlots);
        if (!s.d_in[k] || !s.d_out[k]) return fail("host-buffer pipeline: staging slot missing");
        // the staging slot (and the watermark slot, at most kPipeStreams launches are in flight) is free once everything
        // enqueued on streams[k] so far — the launch that used the slot and its D2H — has finished
        RTB_CUDA(cudaEventRecord(s.ev_done[k], s.streams[k]));
        RTB_CUDA(cudaStreamWaitEvent(cp, s.ev_done[k], 0));
        RTB_CUDA(cudaMemsetAsync(ready, 0, sizeof(unsigned long long), cp));
        RTB_CUDA(cudaEventRecord(s.ev_start[k], cp));
        // ORDER MATTERS: every upload and watermark write of this launch is queued on the copy stream BEFORE the kernel is
        // launched.  The kernel still starts (on streams[k], behind ev_start only) before its input has arrived, but nothing
        // it waits for is host-ordered behind the launch call — so the call also returns when launches are synchronous
        // (ncu, compute-sanitizer, CUDA_LAUNCH_BLOCKING=1), where a launch-then-feed order never gets to feed.
        // If feeding fails (e.g. a bad host pointer) no launch is waiting yet: report the error, nothing spins.
        for (size_t off = 0; off < m; off += sub_rays) {
            const size_t sub = std::min(sub_rays, m - off);
            RTB_CUDA(cudaMemcpyAsync((char*)s.d_in[k] + off * unit_in, (const char*)in + (done + off) * unit_in, sub * unit_in,
                                     cudaMemcpyHostToDevice, cp));
            if (in2)
                RTB_CUDA(cudaMemcpyAsync((char*)s.d_in[k] + s.slot_rays * 12 + off * unit_in, (const char*)in2 + (done + off) * unit_in,
                                         sub * unit_in, cudaMemcpyHostToDevice, cp));
            if (write_value) {  // stream memory operation: no DMA set-up, ordered behind the copies like any stream work
                if (write_value(cp, (unsigned long long)(uintptr_t)ready, off + sub, 0) != 0)
                    return fail("host-buffer pipeline: cuStreamWriteValue64 failed");
            } else {
                if (s.marks_used == kMarkSlots) {  // the pinned source values of in-flight watermark copies must stay intact
                    RTB_CUDA(cudaStreamSynchronize(cp));
                    s.marks_used = 0;
                }
                s.h_marks[s.marks_used] = off + sub;
                RTB_CUDA(cudaMemcpyAsync(ready, &s.h_marks[s.marks_used], sizeof(unsigned long long), cudaMemcpyHostToDevice, cp));
                s.marks_used++;
            }
        }
        RTB_CUDA(cudaStreamWaitEvent(s.streams[k], s.ev_start[k], 0));
        RTB_CUDA(launch(s.d_in[k], m, s.d_out[k], (const unsigned long long*)ready, s.streams[k]));
        // with a pageable `out` this call blocks until the launch has finished
        RTB_CUDA(cudaMemcpyAsync((char*)out + done * unit_out, s.d_out[k], m * unit_out, cudaMemcpyDeviceToHost, s.streams[k]));
        done += m;
    }
    return Ok;
}

// gate_ok: the call is a single-ray call in the caller's order (no ray sorting, default kernel), i.e. it may use the
// gated flavour when RTBVH_HOST_MODE selects it.
template <class Launch>
ResultCode enqueue_any(Scene& s, const void* in, size_t units, size_t unit_in, size_t unit_out, size_t rays_per_unit, void* out,
                       Launch&& launch, const void* in2, size_t unit_in2, bool gate_ok, bool blocking) {
    const int mode = host_mode();
    const bool gated = mode != kHostStaged;
    // The gated flavour queues all uploads of a launch before the launch itself (see there); cudaMemcpyAsync from PAGEABLE
    // memory returns only once the source has been staged, which would serialise upload and traversal — pageable inputs
    // therefore take the staged flavour (chunk i traces while chunk i+1 is staged).  RTBVH_HOST_MODE=gated forces gated.
    auto pinned = [](const void* p) {
        if (!p) return true;
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
    };
    const bool overlap_ok = mode == kHostGated || (pinned(in) && pinned(in2));
    if (gate_ok && gated && overlap_ok && (in2 == nullptr || unit_in2 == unit_in)) {
        // every launch ends with a drain phase (its warps can no longer refill; the longest rays of a launch alone take
        // ~0.25 ms): the fewer launches per batch the better, so the slots grow with the batch (up to 8 Mi rays)
        if (ensure_pipeline(s, gated_slot_rays(units)) != Ok) return Error;
        return enqueue_host_batch_gated(s, in, units, unit_in, unit_out, out, blocking, launch, in2);
    }
    return enqueue_host_batch(s, in, units, unit_in, unit_out, rays_per_unit, out, launch, in2, unit_in2);
}

// Synchronous host-buffer call: enqueue, drain, check.
template <class Launch>
ResultCode run_host_batch(Scene& s, const void* in, size_t units, size_t unit_in, size_t unit_out, size_t rays_per_unit,
                          void* out, Launch&& launch, const void* in2 = nullptr, size_t unit_in2 = 0, bool gate_ok = false) {
    if (units == 0) return Ok;
    if (!in || !out) return fail("null host buffer");
    std::lock_guard<std::mutex> lk(s.pipe_mutex);
    DeviceGuard dg(s.device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    if (ensure_pipeline(s, gate_ok ? gated_slot_rays(units) : 0) != Ok) return Error;
    if (enqueue_any(s, in, units, unit_in, unit_out, rays_per_unit, out, launch, in2, unit_in2, gate_ok, true) != Ok) return Error;
    RTB_CUDA(cudaStreamSynchronize(s.copy_stream));
    for (int i = 0; i < kPipeStreams; i++) RTB_CUDA(cudaStreamSynchronize(s.streams[i]));
    return check_overflow(s);
}

// Asynchronous host-buffer call: enqueue and hand out a ticket; rtbvh_gpu_wait(ticket) blocks until this batch's records
// are in `out`.  Consecutive submissions share the staging slots and streams, so batch k+1 uploads while batch k still
// traces / downloads: the per-call pipeline fill and drain disappear from a steady stream of batches.
template <class Launch>
ResultCode submit_host_batch(Scene& s, const void* in, size_t units, size_t unit_in, size_t unit_out, size_t rays_per_unit,
                             void* out, uint64_t* ticket, Launch&& launch, const void* in2 = nullptr, size_t unit_in2 = 0,
                             bool gate_ok = false) {
    if (!ticket) return fail("null ticket");
    if (units != 0 && (!in || !out)) return fail("null host buffer");
    std::unique_lock<std::mutex> lk(s.pipe_mutex);
    DeviceGuard dg(s.device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    if (ensure_pipeline(s, gate_ok ? gated_slot_rays(units) : 0) != Ok) return Error;
    const uint64_t id = s.next_ticket++;
    Scene::Ticket& t = s.tickets[id % kTickets];
    if (t.id != 0) {  // the ring wrapped onto a ticket nobody waited for: it must have completed before its events are reused
        for (int i = 0; i < kPipeStreams; i++) RTB_CUDA(cudaEventSynchronize(t.ev[i]));
    }
    if (units != 0 && enqueue_any(s, in, units, unit_in, unit_out, rays_per_unit, out, launch, in2, unit_in2, gate_ok, false) != Ok)
        return Error;
    for (int i = 0; i < kPipeStreams; i++) {
        if (!t.ev[i]) RTB_CUDA(cudaEventCreateWithFlags(&t.ev[i], cudaEventDisableTiming));
        RTB_CUDA(cudaEventRecord(t.ev[i], s.streams[i]));
    }
    t.id = id;
    *ticket = id;
    return Ok;
}

ResultCode wait_ticket(Scene& s, uint64_t ticket) {
    DeviceGuard dg(s.device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    cudaEvent_t ev[kPipeStreams] = {};
    {
        std::lock_guard<std::mutex> lk(s.pipe_mutex);
        if (!s.streams[0]) return Ok;  // nothing was ever submitted
        if (ticket == 0) {             // everything submitted so far
            for (int i = 0; i < kPipeStreams; i++) RTB_CUDA(cudaStreamSynchronize(s.streams[i]));
            return check_overflow(s);
        }
        if (ticket >= s.next_ticket) return fail("unknown ticket");
        const Scene::Ticket& t = s.tickets[ticket % kTickets];
        // a recycled slot means submit_host_batch already waited for this ticket before reusing its events
        if (t.id != ticket) return check_overflow(s);
        for (int i = 0; i < kPipeStreams; i++) ev[i] = t.ev[i];
    }
    for (int i = 0; i < kPipeStreams; i++) RTB_CUDA(cudaEventSynchronize(ev[i]));
    return check_overflow(s);
}

ResultCode upload(void** dst, const void* src, size_t bytes) {
    RTB_CUDA(cudaMalloc(dst, bytes ? bytes : 16));
    if (bytes) RTB_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return Ok;
}

}  // namespace

extern "C" {

int rtbvh_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

ResultCode rtbvh_gpu_set_device(int device) {
    RTB_CUDA(cudaSetDevice(device));
    return Ok;
}

const char* rtbvh_gpu_last_error(void) { return g_last_error.c_str(); }

ResultCode rtbvh_gpu_scene_create(const RTBvh* bvh, const RTMbvh* mbvh, const float* vertices, size_t vertex_stride,
                                  size_t triangle_count, RTGpuScene* scene) {
    if (!scene || !vertices || (!bvh && !mbvh)) return fail("rtbvh_gpu_scene_create: null argument");
    if (vertex_stride != 12 && vertex_stride != 16) return fail("vertex_stride must be 12 or 16 bytes");
    if (bvh && (!bvh->nodes || !bvh->indices)) return fail("RTBvh with null pointers");
    if (mbvh && (!mbvh->nodes || !mbvh->indices)) return fail("RTMbvh with null pointers");
    if (rtbvh_gpu_device_count() == 0) return fail("no CUDA device: the traversal path has no CPU fallback");
    auto s = std::make_shared<Scene>();
    RTB_CUDA(cudaGetDevice(&s->device));
    RTB_CUDA(cudaMalloc(&s->d_overflow, sizeof(uint32_t)));
    RTB_CUDA(cudaMemset(s->d_overflow, 0, sizeof(uint32_t)));
    RTB_CUDA(cudaMalloc(&s->d_counters, kCounterSlots * sizeof(unsigned long long)));
    float* d_verts = nullptr;
    const size_t vbytes = triangle_count * 3 * vertex_stride;
    if (upload((void**)&d_verts, vertices, vbytes) != Ok) return Error;
    s->tri_count = (uint32_t)triangle_count;
    auto gather = [&](const uint32_t* indices, uint32_t index_count, TriRec** out, uint32_t** d_idx_out) -> ResultCode {
        uint32_t* d_idx = nullptr;
        if (upload((void**)&d_idx, indices, (size_t)index_count * 4) != Ok) return Error;
        *d_idx_out = d_idx;
        RTB_CUDA(cudaMalloc((void**)out, (size_t)(index_count ? index_count : 1) * sizeof(TriRec)));
        RTB_CUDA(launch_gather_tris(d_verts, (uint32_t)(vertex_stride / 4), d_idx, index_count, (uint32_t)triangle_count,
                                    *out, 0));
        RTB_CUDA(cudaDeviceSynchronize());
        return Ok;
    };
    ResultCode rc = Ok;
    if (bvh) {
        rc = upload(&s->d_bvh_nodes, bvh->nodes, (size_t)bvh->node_count * sizeof(RTBvhNode));
        if (rc == Ok) rc = gather(bvh->indices, bvh->index_count, &s->d_tris_bvh, &s->d_idx_bvh);
        s->bvh = DeviceTree{(const float4*)s->d_bvh_nodes, bvh->node_count, s->d_tris_bvh, bvh->index_count, nullptr, 0};
    }
    if (rc == Ok && mbvh) {
        rc = upload(&s->d_mbvh_nodes, mbvh->nodes, (size_t)mbvh->node_count * sizeof(RTMbvhNode));
        const bool same = bvh && bvh->index_count == mbvh->index_count &&
                          (bvh->indices == mbvh->indices ||
                           std::memcmp(bvh->indices, mbvh->indices, (size_t)bvh->index_count * 4) == 0);
        if (rc == Ok) {
            if (same) {
                s->d_tris_mbvh = s->d_tris_bvh;  // Mbvh keeps a clone of the Bvh's prim_indices (src/bvh.rs:399-403)
                s->d_idx_mbvh = s->d_idx_bvh;
            } else {
                rc = gather(mbvh->indices, mbvh->index_count, &s->d_tris_mbvh, &s->d_idx_mbvh);
            }
        }
        s->mbvh = DeviceTree{(const float4*)s->d_mbvh_nodes, mbvh->node_count, s->d_tris_mbvh, mbvh->index_count, nullptr, 0};
    }
    cudaFree(d_verts);
    if (rc != Ok) return rc;
    if (mbvh && scene_build_top(*s, 0, true) != Ok) return Error;
    if (bvh && bvh->node_count) {
        const RTAabb& r = bvh->nodes[0].aabb;
        for (int k = 0; k < 3; k++) {
            s->bounds[k] = r.min[k];
            s->bounds[3 + k] = r.max[k];
        }
    } else if (mbvh && mbvh->node_count) {
        const RTMbvhNode& r = mbvh->nodes[0];
        const float* mn[3] = {r.min_x, r.min_y, r.min_z};
        const float* mx[3] = {r.max_x, r.max_y, r.max_z};
        for (int k = 0; k < 3; k++) {
            s->bounds[k] = std::min(std::min(mn[k][0], mn[k][1]), std::min(mn[k][2], mn[k][3]));
            s->bounds[3 + k] = std::max(std::max(mx[k][0], mx[k][1]), std::max(mx[k][2], mx[k][3]));
        }
    }
    std::unique_lock<std::shared_mutex> lk(g_scenes.mu);
    g_scenes.scenes.push_back(s);
    *scene = (RTGpuScene)g_scenes.scenes.size();
    return Ok;
}

// ---- build into a scene without host mirrors ----------------------------------------------------------------------
static ResultCode scene_build_common(const float* vertices, bool on_device, size_t vertex_stride, size_t triangle_count,
                                     size_t prims_per_leaf, BvhType type, int want_mbvh, RTGpuScene* scene) {
    if (!scene || !vertices) return fail("rtbvh_gpu_scene_build: null argument");
    if (vertex_stride != 12 && vertex_stride != 16) return fail("vertex_stride must be 12 or 16 bytes");
    if (triangle_count == 0) return NoPrimitives;
    if (rtbvh_gpu_device_count() == 0) return fail("no CUDA device: the builders run on the GPU only (no CPU fallback)");
    auto s = std::make_shared<Scene>();
    RTB_CUDA(cudaGetDevice(&s->device));
    RTB_CUDA(cudaMalloc(&s->d_overflow, sizeof(uint32_t)));
    RTB_CUDA(cudaMemset(s->d_overflow, 0, sizeof(uint32_t)));
    RTB_CUDA(cudaMalloc(&s->d_counters, kCounterSlots * sizeof(unsigned long long)));
    ResidentTrees rt;
    const ResultCode rc = gpu_build_resident(vertices, on_device, vertex_stride, triangle_count, prims_per_leaf, (uint32_t)type,
                                             want_mbvh != 0, &rt);
    s->d_bvh_nodes = rt.d_nodes;  // owned by the scene from here on (freed by ~Scene also on the error paths below)
    s->d_idx_bvh = rt.d_indices;
    s->d_mbvh_nodes = rt.d_mnodes;
    s->d_idx_mbvh = rt.d_mnodes ? rt.d_indices : nullptr;
    s->d_refit_verts = rt.d_vertices;
    s->refit_verts_bytes = rt.d_vertices ? triangle_count * 3 * vertex_stride : 0;
    if (rc != Ok) return rc;
    s->tri_count = (uint32_t)triangle_count;
    const float* d_verts = on_device ? vertices : rt.d_vertices;
    RTB_CUDA(dev_block_alloc((void**)&s->d_tris_bvh, (size_t)rt.index_count * sizeof(TriRec)));
    RTB_CUDA(launch_gather_tris(d_verts, (uint32_t)(vertex_stride / 4), rt.d_indices, rt.index_count, (uint32_t)triangle_count,
                                s->d_tris_bvh, 0));
    s->bvh = DeviceTree{(const float4*)s->d_bvh_nodes, rt.node_count, s->d_tris_bvh, rt.index_count, nullptr, 0};
    if (rt.d_mnodes) {
        s->d_tris_mbvh = s->d_tris_bvh;
        s->mbvh = DeviceTree{(const float4*)s->d_mbvh_nodes, rt.m_count, s->d_tris_mbvh, rt.index_count, nullptr, 0};
        if (scene_build_top(*s, 0, true) != Ok) return Error;
    }
    float4 root[2];
    RTB_CUDA(cudaMemcpy(root, s->d_bvh_nodes, 32, cudaMemcpyDeviceToHost));  // also drains the gather
    s->bounds[0] = root[0].x; s->bounds[1] = root[0].y; s->bounds[2] = root[0].z;
    s->bounds[3] = root[1].x; s->bounds[4] = root[1].y; s->bounds[5] = root[1].z;
    std::unique_lock<std::shared_mutex> lk(g_scenes.mu);
    g_scenes.scenes.push_back(s);
    *scene = (RTGpuScene)g_scenes.scenes.size();
    return Ok;
}
ResultCode rtbvh_gpu_scene_build(const float* vertices, size_t vertex_stride, size_t triangle_count, size_t prims_per_leaf,
                                 BvhType type, int want_mbvh, RTGpuScene* scene) {
    return scene_build_common(vertices, false, vertex_stride, triangle_count, prims_per_leaf, type, want_mbvh, scene);
}
ResultCode rtbvh_gpu_scene_build_device(const float* d_vertices, size_t vertex_stride, size_t triangle_count,
                                        size_t prims_per_leaf, BvhType type, int want_mbvh, RTGpuScene* scene) {
    return scene_build_common(d_vertices, true, vertex_stride, triangle_count, prims_per_leaf, type, want_mbvh, scene);
}
ResultCode rtbvh_gpu_scene_tree_size(RTGpuScene h, RTTreeKind tree, uint32_t* node_count, uint32_t* index_count) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    if (node_count) *node_count = t->node_count;
    if (index_count) *index_count = t->index_count;
    return Ok;
}
ResultCode rtbvh_gpu_scene_read_indices(RTGpuScene h, RTTreeKind tree, uint32_t* out, size_t count) {
    auto s = get_scene(h);
    if (!s || !out) return fail("unknown scene / null buffer");
    const DeviceTree* t = pick_tree(*s, tree);
    const uint32_t* d = tree == RT_TREE_MBVH ? s->d_idx_mbvh : s->d_idx_bvh;
    if (!t || !d) return fail("scene has no such tree");
    if (count != t->index_count) return fail("rtbvh_gpu_scene_read_indices: count must equal the tree's index_count");
    DeviceGuard dg(s->device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    RTB_CUDA(cudaMemcpy(out, d, count * 4, cudaMemcpyDeviceToHost));
    return Ok;
}

ResultCode rtbvh_gpu_trim_workspace(void) { return gpu_trim_workspace(); }

// ---- dynamic scenes ------------------------------------------------------------------------------
static ResultCode scene_refit_on(Scene& s, const float* d_vertices, size_t vertex_stride, size_t triangle_count, cudaStream_t st) {
    if (vertex_stride != 12 && vertex_stride != 16) return fail("vertex_stride must be 12 or 16 bytes");
    if (triangle_count != s.tri_count) return fail("scene refit: the triangle count must not change (rebuild instead)");
    if (!s.d_bvh_nodes) return fail("scene refit needs the binary Bvh in the scene (an Mbvh alone has no refit in the reference either)");
    if (s.d_mbvh_nodes && s.d_idx_mbvh != s.d_idx_bvh) return fail("scene refit: the scene's Mbvh is not the 4-wide collapse of its Bvh");
    const uint32_t vs = (uint32_t)(vertex_stride / 4);
    if (gpu_refit_resident(&s.refit_cache, (float4*)s.d_bvh_nodes, (uint32_t)s.bvh.node_count, s.d_idx_bvh, (uint32_t)s.bvh.index_count,
                           d_vertices, vs, s.tri_count, (float4*)s.d_mbvh_nodes, (uint32_t)s.mbvh.node_count, st) != Ok)
        return Error;
    RTB_CUDA(launch_gather_tris(d_vertices, vs, s.d_idx_bvh, (uint32_t)s.bvh.index_count, s.tri_count, s.d_tris_bvh, st));
    return scene_build_top(s, st, false);  // the slot boxes of the staged top follow the refitted Mbvh
}
// Ordering of a refit against the scene's own work (the caller holds refit_mutex and pipe_mutex):
//  * before: the refit stream waits for everything queued so far on the host-buffer pipeline streams (those batches still
//    read the old nodes and records) and for the previous refit (all refits of a scene share one scratch area);
//  * after: the pipeline streams wait for the refit, so host-buffer batches submitted later see the refitted trees.
// Device-resident traversal calls on the CALLER's streams are ordered by the caller (same stream, or events), as for any
// stream-ordered API; rtbvh_gpu.h says so.
static ResultCode refit_order_begin(Scene& s, cudaStream_t st) {
    if (!s.ev_refit) RTB_CUDA(cudaEventCreateWithFlags(&s.ev_refit, cudaEventDisableTiming));
    if (s.streams[0]) {
        for (int i = 0; i < kPipeStreams; i++) {
            RTB_CUDA(cudaEventRecord(s.ev_refit, s.streams[i]));
            RTB_CUDA(cudaStreamWaitEvent(st, s.ev_refit, 0));
        }
    }
    if (s.ev_refit_done) RTB_CUDA(cudaStreamWaitEvent(st, s.ev_refit_done, 0));
    return Ok;
}
static ResultCode refit_order_end(Scene& s, cudaStream_t st) {
    if (!s.ev_refit_done) RTB_CUDA(cudaEventCreateWithFlags(&s.ev_refit_done, cudaEventDisableTiming));
    RTB_CUDA(cudaEventRecord(s.ev_refit_done, st));
    if (s.streams[0])
        for (int i = 0; i < kPipeStreams; i++) RTB_CUDA(cudaStreamWaitEvent(s.streams[i], s.ev_refit_done, 0));
    return Ok;
}
ResultCode rtbvh_gpu_scene_refit_device(RTGpuScene h, const float* d_vertices, size_t vertex_stride, size_t triangle_count,
                                        void* stream) {
    auto s = get_scene(h);
    if (!s || !d_vertices) return fail("unknown scene / null vertices");
    std::lock_guard<std::mutex> lk(s->refit_mutex);
    std::lock_guard<std::mutex> lp(s->pipe_mutex);
    DeviceGuard dg(s->device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    if (refit_order_begin(*s, (cudaStream_t)stream) != Ok) return Error;
    const ResultCode rc = scene_refit_on(*s, d_vertices, vertex_stride, triangle_count, (cudaStream_t)stream);
    if (refit_order_end(*s, (cudaStream_t)stream) != Ok) return Error;  // also after a failed enqueue: whatever was queued is chained
    return rc;
}
ResultCode rtbvh_gpu_scene_refit(RTGpuScene h, const float* vertices, size_t vertex_stride, size_t triangle_count) {
    auto s = get_scene(h);
    if (!s || !vertices) return fail("unknown scene / null vertices");
    std::lock_guard<std::mutex> lk(s->refit_mutex);
    std::lock_guard<std::mutex> lp(s->pipe_mutex);
    DeviceGuard dg(s->device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    // blocking flavour: drain the scene's pipeline (outstanding *_async batches still traverse the old trees), refit on
    // the legacy stream, wait for it
    if (s->streams[0]) {
        for (int i = 0; i < kPipeStreams; i++) RTB_CUDA(cudaStreamSynchronize(s->streams[i]));
        RTB_CUDA(cudaStreamSynchronize(s->copy_stream));
    }
    const size_t bytes = triangle_count * 3 * vertex_stride;
    if (bytes > s->refit_verts_bytes) {
        RTB_CUDA(cudaDeviceSynchronize());  // an earlier refit_device may still read the old staging buffer
        dev_block_free(s->d_refit_verts);
        s->d_refit_verts = nullptr;
        s->refit_verts_bytes = 0;
        RTB_CUDA(dev_block_alloc((void**)&s->d_refit_verts, bytes ? bytes : 16));
        s->refit_verts_bytes = bytes;
    }
    if (refit_order_begin(*s, 0) != Ok) return Error;
    RTB_CUDA(cudaMemcpyAsync(s->d_refit_verts, vertices, bytes, cudaMemcpyHostToDevice, 0));
    const ResultCode rc = scene_refit_on(*s, s->d_refit_verts, vertex_stride, triangle_count, 0);
    if (refit_order_end(*s, 0) != Ok) return Error;
    if (rc != Ok) return rc;
    RTB_CUDA(cudaStreamSynchronize(0));
    float4 root[2];  // the keys of the optional ray sort follow the new root box
    RTB_CUDA(cudaMemcpy(root, s->d_bvh_nodes, 32, cudaMemcpyDeviceToHost));
    s->bounds[0] = root[0].x; s->bounds[1] = root[0].y; s->bounds[2] = root[0].z;
    s->bounds[3] = root[1].x; s->bounds[4] = root[1].y; s->bounds[5] = root[1].z;
    return Ok;
}
ResultCode rtbvh_gpu_scene_read_nodes(RTGpuScene h, RTTreeKind tree, void* out, size_t bytes) {
    auto s = get_scene(h);
    if (!s || !out) return fail("unknown scene / null buffer");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    const size_t have = (size_t)t->node_count * (tree == RT_TREE_MBVH ? sizeof(RTMbvhNode) : sizeof(RTBvhNode));
    if (bytes != have) return fail("rtbvh_gpu_scene_read_nodes: buffer size must be node_count * sizeof(node)");
    DeviceGuard dg(s->device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    RTB_CUDA(cudaDeviceSynchronize());
    RTB_CUDA(cudaMemcpy(out, t->nodes, bytes, cudaMemcpyDeviceToHost));
    return Ok;
}

ResultCode rtbvh_gpu_scene_set_ray_sorting(RTGpuScene h, int enable) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    s->sort_rays.store(enable ? 1 : 0);
    return Ok;
}

ResultCode rtbvh_gpu_scene_set_ray_tiling(RTGpuScene h, uint32_t row_length) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    if (row_length % 8 != 0) return fail("rtbvh_gpu_scene_set_ray_tiling: the row length must be a multiple of 8 (0 = off)");
    s->tile_w.store(row_length);
    return Ok;
}

ResultCode rtbvh_gpu_scene_free(RTGpuScene h) {
    std::unique_lock<std::shared_mutex> lk(g_scenes.mu);
    if (h == 0 || h > g_scenes.scenes.size() || !g_scenes.scenes[h - 1]) return fail("unknown scene");
    g_scenes.scenes[h - 1].reset();
    return Ok;
}

// ---- scene replication: build once, copy device to device ------------------------------------
// The six device arrays of a scene, in this order: Bvh nodes, Bvh prim_indices, Bvh-order triangle records, Mbvh nodes,
// Mbvh prim_indices, Mbvh-order triangle records (the last two usually alias the Bvh's: kSharedLeafOrder).
namespace {
constexpr uint32_t kExportMagic = 0x52544258u;  // "RTBX"
constexpr uint32_t kHasBvh = 1u, kHasMbvh = 2u, kSharedLeafOrder = 4u;
struct SceneBlob {
    uint32_t magic, version;
    int32_t device;
    uint32_t flags, tri_count;
    uint32_t bvh_nodes, bvh_indices, mbvh_nodes, mbvh_indices;
    float bounds[6];
    cudaIpcMemHandle_t handles[6];
};
static_assert(sizeof(SceneBlob) <= sizeof(RTGpuSceneExport), "export blob must fit the public POD");

void scene_describe(const Scene& s, SceneBlob& b, const void* src[6], size_t bytes[6]) {
    std::memset(&b, 0, sizeof(b));
    b.magic = kExportMagic;
    b.version = 1;
    b.device = s.device;
    b.tri_count = s.tri_count;
    if (s.d_bvh_nodes) b.flags |= kHasBvh;
    if (s.d_mbvh_nodes) b.flags |= kHasMbvh;
    if (s.d_bvh_nodes && s.d_mbvh_nodes && s.d_tris_mbvh == s.d_tris_bvh) b.flags |= kSharedLeafOrder;
    b.bvh_nodes = s.bvh.node_count;
    b.bvh_indices = s.bvh.index_count;
    b.mbvh_nodes = s.mbvh.node_count;
    b.mbvh_indices = s.mbvh.index_count;
    std::memcpy(b.bounds, s.bounds, sizeof(b.bounds));
    const bool own_m = (b.flags & kHasMbvh) && !(b.flags & kSharedLeafOrder);
    src[0] = s.d_bvh_nodes;   bytes[0] = (b.flags & kHasBvh) ? (size_t)b.bvh_nodes * 32 : 0;
    src[1] = s.d_idx_bvh;     bytes[1] = (b.flags & kHasBvh) ? (size_t)b.bvh_indices * 4 : 0;
    src[2] = s.d_tris_bvh;    bytes[2] = (b.flags & kHasBvh) ? (size_t)b.bvh_indices * sizeof(TriRec) : 0;
    src[3] = s.d_mbvh_nodes;  bytes[3] = (b.flags & kHasMbvh) ? (size_t)b.mbvh_nodes * 128 : 0;
    src[4] = s.d_idx_mbvh;    bytes[4] = own_m ? (size_t)b.mbvh_indices * 4 : 0;
    src[5] = s.d_tris_mbvh;   bytes[5] = own_m ? (size_t)b.mbvh_indices * sizeof(TriRec) : 0;
}

// New scene on the CURRENT device from six source arrays that are readable from it: `src_device` >= 0 selects
// cudaMemcpyPeer (clone inside one process), -1 a plain device-to-device copy (cudaIpc-mapped peer memory).
ResultCode scene_replicate(const SceneBlob& b, const void* const src[6], const size_t bytes[6], int src_device, RTGpuScene* out) {
    auto s = std::make_shared<Scene>();
    RTB_CUDA(cudaGetDevice(&s->device));
    RTB_CUDA(cudaMalloc(&s->d_overflow, sizeof(uint32_t)));
    RTB_CUDA(cudaMemset(s->d_overflow, 0, sizeof(uint32_t)));
    RTB_CUDA(cudaMalloc(&s->d_counters, kCounterSlots * sizeof(unsigned long long)));
    void* dst[6] = {};
    for (int k = 0; k < 6; k++) {
        if (bytes[k] == 0) continue;
        cudaError_t e = dev_block_alloc(&dst[k], bytes[k]);
        if (e == cudaSuccess)
            e = src_device >= 0 ? cudaMemcpyPeer(dst[k], s->device, src[k], src_device, bytes[k])
                                : cudaMemcpy(dst[k], src[k], bytes[k], cudaMemcpyDeviceToDevice);
        // the scene owns whatever has been allocated so far (freed by ~Scene on the error path)
        if (k == 0) s->d_bvh_nodes = dst[k];
        if (k == 1) s->d_idx_bvh = (uint32_t*)dst[k];
        if (k == 2) s->d_tris_bvh = (TriRec*)dst[k];
        if (k == 3) s->d_mbvh_nodes = dst[k];
        if (k == 4) s->d_idx_mbvh = (uint32_t*)dst[k];
        if (k == 5) s->d_tris_mbvh = (TriRec*)dst[k];
        if (e != cudaSuccess) return fail("scene replication: device-to-device copy failed", e);
    }
    s->tri_count = b.tri_count;
    std::memcpy(s->bounds, b.bounds, sizeof(b.bounds));
    if (b.flags & kHasBvh)
        s->bvh = DeviceTree{(const float4*)s->d_bvh_nodes, b.bvh_nodes, s->d_tris_bvh, b.bvh_indices, nullptr, 0};
    if (b.flags & kHasMbvh) {
        if (b.flags & kSharedLeafOrder) {
            s->d_idx_mbvh = s->d_idx_bvh;
            s->d_tris_mbvh = s->d_tris_bvh;
        }
        s->mbvh = DeviceTree{(const float4*)s->d_mbvh_nodes, b.mbvh_nodes, s->d_tris_mbvh, b.mbvh_indices, nullptr, 0};
        if (scene_build_top(*s, 0, true) != Ok) return Error;
    }
    RTB_CUDA(cudaDeviceSynchronize());
    std::unique_lock<std::shared_mutex> lk(g_scenes.mu);
    g_scenes.scenes.push_back(s);
    *out = (RTGpuScene)g_scenes.scenes.size();
    return Ok;
}
}  // namespace

ResultCode rtbvh_gpu_scene_export(RTGpuScene h, RTGpuSceneExport* out) {
    auto s = get_scene(h);
    if (!s || !out) return fail("rtbvh_gpu_scene_export: unknown scene / null argument");
    DeviceGuard dg(s->device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    SceneBlob b;
    const void* src[6];
    size_t bytes[6];
    scene_describe(*s, b, src, bytes);
    RTB_CUDA(cudaDeviceSynchronize());  // builds / refits enqueued on the scene have finished before a peer reads it
    for (int k = 0; k < 6; k++)
        if (bytes[k]) RTB_CUDA(cudaIpcGetMemHandle(&b.handles[k], const_cast<void*>(src[k])));
    std::memset(out, 0, sizeof(*out));
    std::memcpy(out->bytes, &b, sizeof(b));
    return Ok;
}

ResultCode rtbvh_gpu_scene_import(const RTGpuSceneExport* exported, RTGpuScene* scene) {
    if (!exported || !scene) return fail("rtbvh_gpu_scene_import: null argument");
    if (rtbvh_gpu_device_count() == 0) return fail("no CUDA device: the traversal path has no CPU fallback");
    SceneBlob b;
    std::memcpy(&b, exported->bytes, sizeof(b));
    if (b.magic != kExportMagic || b.version != 1 || !(b.flags & (kHasBvh | kHasMbvh)))
        return fail("rtbvh_gpu_scene_import: not a scene export of this library version");
    const size_t want[6] = {(b.flags & kHasBvh) ? (size_t)b.bvh_nodes * 32 : 0,
                            (b.flags & kHasBvh) ? (size_t)b.bvh_indices * 4 : 0,
                            (b.flags & kHasBvh) ? (size_t)b.bvh_indices * sizeof(TriRec) : 0,
                            (b.flags & kHasMbvh) ? (size_t)b.mbvh_nodes * 128 : 0,
                            ((b.flags & kHasMbvh) && !(b.flags & kSharedLeafOrder)) ? (size_t)b.mbvh_indices * 4 : 0,
                            ((b.flags & kHasMbvh) && !(b.flags & kSharedLeafOrder)) ? (size_t)b.mbvh_indices * sizeof(TriRec) : 0};
    void* mapped[6] = {};
    ResultCode rc = Ok;
    for (int k = 0; k < 6 && rc == Ok; k++) {
        if (!want[k]) continue;
        const cudaError_t e = cudaIpcOpenMemHandle(&mapped[k], b.handles[k], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            mapped[k] = nullptr;
            rc = fail("rtbvh_gpu_scene_import: cudaIpcOpenMemHandle (exports are opened by ANOTHER process; peer access needed)", e);
        }
    }
    if (rc == Ok) rc = scene_replicate(b, mapped, want, -1, scene);
    for (int k = 0; k < 6; k++)
        if (mapped[k]) cudaIpcCloseMemHandle(mapped[k]);
    return rc;
}

ResultCode rtbvh_gpu_scene_clone(RTGpuScene h, int device, RTGpuScene* clone) {
    auto s = get_scene(h);
    if (!s || !clone) return fail("rtbvh_gpu_scene_clone: unknown scene / null argument");
    if (device < 0 || device >= rtbvh_gpu_device_count()) return fail("rtbvh_gpu_scene_clone: no such device");
    SceneBlob b;
    const void* src[6];
    size_t bytes[6];
    scene_describe(*s, b, src, bytes);
    {
        DeviceGuard dg(s->device);
        if (!dg.ok()) return fail("cudaSetDevice", dg.err);
        RTB_CUDA(cudaDeviceSynchronize());
    }
    DeviceGuard dg(device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    return scene_replicate(b, src, bytes, s->device, clone);
}

// Packet kernels (RTBVH_PACKET_MODE = static | persistent | lane).  Measured on config 2's frames as packets of four
// x-adjacent pixels (scripts/trace_ab.py --packets, profiles/r5_trace_ab.md), closest / any hit Mrays/s:
//   Mbvh: quad static 1 135 / 1 221, quad persistent 1 288 / 1 520, one lane per packet 2 008 / 2 183
//   Bvh:  quad static   404 / 1 103, quad persistent   360,         one lane per packet   743 / 2 102
// Default: one lane per packet for both trees.
int packet_mode(RTTreeKind tree) {
    static const int forced = [] {
        const char* e = std::getenv("RTBVH_PACKET_MODE");
        if (!e) return -1;
        const std::string m(e);
        if (m == "persistent") return (int)kTracePersistent;
        if (m == "static") return (int)kTraceStatic;
        if (m == "lane") return (int)kTraceLane;
        return -1;
    }();
    if (forced >= 0) return forced;
    (void)tree;
    return (int)kTraceLane;
}

// ---- device-resident, asynchronous ---------------------------------------------------------------
ResultCode rtbvh_gpu_intersect_device(RTGpuScene h, RTTreeKind tree, const RTRay* d_rays, size_t n, RTHit* d_hits,
                                      void* stream) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    if (!on_scene_device(*s)) return fail("the scene lives on another device than the current one (cudaSetDevice / rtbvh_gpu_set_device first)");
    PeerDests pd{};
    s->apply_tiling(pd, n);
    RTB_CUDA(launch_trace_single(*t, tree, false, d_rays, n, d_hits, nullptr, s->counter_slot(), s->d_overflow,
                                 persistent_mode(), s->sort_bounds(), pd.tile_w ? &pd : nullptr, (cudaStream_t)stream));
    return Ok;
}
ResultCode rtbvh_gpu_occluded_device(RTGpuScene h, RTTreeKind tree, const RTRay* d_rays, size_t n, uint8_t* d_occ,
                                     void* stream) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    if (!on_scene_device(*s)) return fail("the scene lives on another device than the current one (cudaSetDevice / rtbvh_gpu_set_device first)");
    PeerDests pd{};
    s->apply_tiling(pd, n);
    RTB_CUDA(launch_trace_single(*t, tree, true, d_rays, n, nullptr, d_occ, s->counter_slot(), s->d_overflow,
                                 persistent_mode(), s->sort_bounds(), pd.tile_w ? &pd : nullptr, (cudaStream_t)stream));
    return Ok;
}
ResultCode rtbvh_gpu_intersect_packets_device(RTGpuScene h, RTTreeKind tree, const RTRayPacket4* d_packets, size_t n,
                                              float t_min, RTHitPacket4* d_hits, void* stream) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    if (!on_scene_device(*s)) return fail("the scene lives on another device than the current one (cudaSetDevice / rtbvh_gpu_set_device first)");
    RTB_CUDA(launch_trace_packets(*t, tree, false, d_packets, n, t_min, d_hits, nullptr, s->counter_slot(), s->d_overflow,
                                  packet_mode(tree), (cudaStream_t)stream));
    return Ok;
}
ResultCode rtbvh_gpu_occluded_packets_device(RTGpuScene h, RTTreeKind tree, const RTRayPacket4* d_packets, size_t n,
                                             float t_min, uint8_t* d_occ, void* stream) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    if (!on_scene_device(*s)) return fail("the scene lives on another device than the current one (cudaSetDevice / rtbvh_gpu_set_device first)");
    RTB_CUDA(launch_trace_packets(*t, tree, true, d_packets, n, t_min, nullptr, d_occ, s->counter_slot(), s->d_overflow,
                                  packet_mode(tree), (cudaStream_t)stream));
    return Ok;
}
// ---- multi-GPU: gather fused into the traversal kernel (P2P stores into cudaIpc-mapped peer buffers) -----------
static ResultCode scatter_call(RTGpuScene h, RTTreeKind tree, bool any, const RTRay* d_rays, size_t n, void* d_local,
                               void* const* dests, int dest_count, size_t dest_offset, void* stream) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    if (!on_scene_device(*s)) return fail("the scene lives on another device than the current one (cudaSetDevice / rtbvh_gpu_set_device first)");
    if (dest_count < 0 || dest_count > 8 || (dest_count > 0 && !dests)) return fail("0..8 destinations");
    PeerDests pd{};
    for (int k = 0; k < dest_count; k++) pd.p[k] = dests[k];
    pd.count = dest_count;
    pd.offset = dest_offset;
    s->apply_tiling(pd, n);
    // Store shape of the fused gather (RTBVH_GATHER_PUSH=1 / 0 forces one).  Measured, 8 M rays per step and rank
    // (profiles/r5k_gather_ab_n2.txt, r5n_gather_ab.txt): per-ray stores cost +0.2 % for the own buffer and +2.4 % for 7 peers;
    // the chunk-wise push halves the peer part (+1.1 %) but its bookkeeping costs +1.7 % whatever the peer count:
    // 2 GPUs 4.26 vs 4.33 ms per step, 8 GPUs 4.42 vs 4.41.  Default: push only beyond 4 destinations.
    static const int push_mode = [] {
        const char* e = std::getenv("RTBVH_GATHER_PUSH");
        return e ? (e[0] == '0' ? 0 : 1) : -1;
    }();
    const bool push = push_mode < 0 ? dest_count > 4 : push_mode == 1;
    pd.push = (push && dest_count > 0 && d_local != nullptr) ? 1 : 0;
    RTB_CUDA(launch_trace_single(*t, tree, any, d_rays, n, any ? nullptr : (RTHit*)d_local, any ? (uint8_t*)d_local : nullptr,
                                 s->counter_slot(), s->d_overflow, refill_mode(), s->sort_bounds(), &pd, (cudaStream_t)stream));
    return Ok;
}
ResultCode rtbvh_gpu_intersect_device_scatter(RTGpuScene h, RTTreeKind tree, const RTRay* d_rays, size_t n, RTHit* d_hits,
                                              void* const* dests, int dest_count, size_t dest_offset, void* stream) {
    return scatter_call(h, tree, false, d_rays, n, d_hits, dests, dest_count, dest_offset, stream);
}
ResultCode rtbvh_gpu_occluded_device_scatter(RTGpuScene h, RTTreeKind tree, const RTRay* d_rays, size_t n, uint8_t* d_occ,
                                             void* const* dests, int dest_count, size_t dest_offset, void* stream) {
    return scatter_call(h, tree, true, d_rays, n, d_occ, dests, dest_count, dest_offset, stream);
}
ResultCode rtbvh_gpu_peer_buffer_create(size_t bytes, void** d_ptr, unsigned char* handle64) {
    if (!d_ptr || !handle64) return fail("null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    RTB_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 256));
    RTB_CUDA(cudaMemset(*d_ptr, 0, bytes ? bytes : 256));
    cudaIpcMemHandle_t hnd;
    RTB_CUDA(cudaIpcGetMemHandle(&hnd, *d_ptr));
    std::memcpy(handle64, &hnd, 64);
    return Ok;
}
ResultCode rtbvh_gpu_peer_buffer_open(const unsigned char* handle64, void** d_ptr) {
    if (!d_ptr || !handle64) return fail("null argument");
    cudaIpcMemHandle_t hnd;
    std::memcpy(&hnd, handle64, 64);
    RTB_CUDA(cudaIpcOpenMemHandle(d_ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
    return Ok;
}
// Cross-GPU step barrier on a stream (one tiny kernel): thread k publishes `value` into slot `rank` of peer k's
// flag array (system-scope release after a system fence, so every P2P record written by earlier kernels of this
// stream is visible first), then spins until slot k of the own flag array has reached `value`.
__global__ void peer_barrier_kernel(PeerDests flags, int rank, unsigned long long value) {
    const int k = threadIdx.x;
    if (k >= flags.count) return;
    __threadfence_system();
    void* pk = nullptr;
    void* pr = nullptr;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j == k) pk = flags.p[j];
        if (j == rank) pr = flags.p[j];
    }
    unsigned long long* remote = static_cast<unsigned long long*>(pk) + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(value) : "memory");
    const unsigned long long* mine = static_cast<const unsigned long long*>(pr) + k;
    unsigned long long seen;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
    } while (seen < value);
}
ResultCode rtbvh_gpu_peer_barrier(void* const* flag_arrays, int count, int rank, uint64_t value, void* stream) {
    if (!flag_arrays || count < 1 || count > 8 || rank < 0 || rank >= count) return fail("1..8 flag arrays, rank < count");
    PeerDests pd{};
    for (int k = 0; k < count; k++) pd.p[k] = flag_arrays[k];
    pd.count = count;
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pd, rank, (unsigned long long)value);
    RTB_CUDA(cudaGetLastError());
    return Ok;
}
ResultCode rtbvh_gpu_peer_buffer_close(void* d_ptr) {
    RTB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return Ok;
}
ResultCode rtbvh_gpu_peer_buffer_free(void* d_ptr) {
    RTB_CUDA(cudaFree(d_ptr));
    return Ok;
}

ResultCode rtbvh_gpu_scene_stack_overflowed(RTGpuScene h, uint32_t* overflowed) {
    auto s = get_scene(h);
    if (!s || !overflowed) return fail("unknown scene");
    DeviceGuard dg(s->device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    RTB_CUDA(cudaDeviceSynchronize());
    RTB_CUDA(cudaMemcpy(overflowed, s->d_overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    RTB_CUDA(cudaMemset(s->d_overflow, 0, sizeof(uint32_t)));
    return Ok;
}

// ---- host buffers ------------------------------------------------------------------------------
// One launch of the single-ray kernel for a staged chunk: RTRay records in `din`; `ready` is the gated flavour's watermark.
static ResultCode rtray_host_call(RTGpuScene h, RTTreeKind tree, bool any, const RTRay* rays, size_t n, void* out, uint64_t* ticket,
                                  bool async) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    const bool gate_ok = !s->sort_bounds() && persistent_mode() == refill_mode();
    auto launch = [&](void* din, size_t m, void* dout, const unsigned long long* ready, cudaStream_t st) {
        PeerDests pd{};
        pd.ready = ready;
        return launch_trace_single(*t, tree, any, (const RTRay*)din, m, any ? nullptr : (RTHit*)dout, any ? (uint8_t*)dout : nullptr,
                                   s->counter_slot(), s->d_overflow, ready ? refill_mode() : persistent_mode(),
                                   ready ? nullptr : s->sort_bounds(), ready ? &pd : nullptr, st);
    };
    const size_t unit_out = any ? 1 : sizeof(RTHit);
    if (async) return submit_host_batch(*s, rays, n, sizeof(RTRay), unit_out, 1, out, ticket, launch, nullptr, 0, gate_ok);
    return run_host_batch(*s, rays, n, sizeof(RTRay), unit_out, 1, out, launch, nullptr, 0, gate_ok);
}
ResultCode rtbvh_gpu_intersect(RTGpuScene h, RTTreeKind tree, const RTRay* rays, size_t n, RTHit* hits) {
    return rtray_host_call(h, tree, false, rays, n, hits, nullptr, false);
}
ResultCode rtbvh_gpu_occluded(RTGpuScene h, RTTreeKind tree, const RTRay* rays, size_t n, uint8_t* occluded) {
    return rtray_host_call(h, tree, true, rays, n, occluded, nullptr, false);
}
// ---- host buffers, asynchronous: submit / wait ------------------------------------------------------------------
ResultCode rtbvh_gpu_intersect_async(RTGpuScene h, RTTreeKind tree, const RTRay* rays, size_t n, RTHit* hits, uint64_t* ticket) {
    return rtray_host_call(h, tree, false, rays, n, hits, ticket, true);
}
ResultCode rtbvh_gpu_occluded_async(RTGpuScene h, RTTreeKind tree, const RTRay* rays, size_t n, uint8_t* occluded, uint64_t* ticket) {
    return rtray_host_call(h, tree, true, rays, n, occluded, ticket, true);
}
// ---- split ray input: origins[3n], directions[3n], common t_min / t_max (24 B per ray across PCIe) -----------------
static cudaError_t launch_od(Scene& s, const DeviceTree& t, RTTreeKind tree, bool any, const float* d_origins,
                             const float* d_directions, size_t m, float t_min, float t_max, void* d_out,
                             const unsigned long long* ready, cudaStream_t st) {
    PeerDests pd{};
    pd.ready = ready;
    pd.directions = d_directions;
    pd.t_min = t_min;
    pd.t_max = t_max;
    return launch_trace_single(t, tree, any, reinterpret_cast<const RTRay*>(d_origins), m, any ? nullptr : (RTHit*)d_out,
                               any ? (uint8_t*)d_out : nullptr, s.counter_slot(), s.d_overflow, refill_mode(), nullptr, &pd, st);
}
static ResultCode od_host_call(RTGpuScene h, RTTreeKind tree, bool any, const float* origins, const float* directions, size_t n,
                               float t_min, float t_max, void* out, uint64_t* ticket) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    if (n != 0 && !directions) return fail("null host buffer");
    auto launch = [&](void* din, size_t m, void* dout, const unsigned long long* ready, cudaStream_t st) {
        return launch_od(*s, *t, tree, any, (const float*)din, (const float*)((const char*)din + s->slot_rays * 12), m, t_min, t_max,
                         dout, ready, st);
    };
    const size_t unit_out = any ? 1 : sizeof(RTHit);
    if (ticket) return submit_host_batch(*s, origins, n, 12, unit_out, 1, out, ticket, launch, directions, 12, true);
    return run_host_batch(*s, origins, n, 12, unit_out, 1, out, launch, directions, 12, true);
}
ResultCode rtbvh_gpu_intersect_od(RTGpuScene h, RTTreeKind tree, const float* origins, const float* directions, size_t n,
                                  float t_min, float t_max, RTHit* hits) {
    return od_host_call(h, tree, false, origins, directions, n, t_min, t_max, hits, nullptr);
}
ResultCode rtbvh_gpu_occluded_od(RTGpuScene h, RTTreeKind tree, const float* origins, const float* directions, size_t n,
                                 float t_min, float t_max, uint8_t* occluded) {
    return od_host_call(h, tree, true, origins, directions, n, t_min, t_max, occluded, nullptr);
}
ResultCode rtbvh_gpu_intersect_od_async(RTGpuScene h, RTTreeKind tree, const float* origins, const float* directions, size_t n,
                                        float t_min, float t_max, RTHit* hits, uint64_t* ticket) {
    if (!ticket) return fail("null ticket");
    return od_host_call(h, tree, false, origins, directions, n, t_min, t_max, hits, ticket);
}
ResultCode rtbvh_gpu_occluded_od_async(RTGpuScene h, RTTreeKind tree, const float* origins, const float* directions, size_t n,
                                       float t_min, float t_max, uint8_t* occluded, uint64_t* ticket) {
    if (!ticket) return fail("null ticket");
    return od_host_call(h, tree, true, origins, directions, n, t_min, t_max, occluded, ticket);
}
ResultCode rtbvh_gpu_intersect_od_device(RTGpuScene h, RTTreeKind tree, const float* d_origins, const float* d_directions,
                                         size_t n, float t_min, float t_max, RTHit* d_hits, void* stream) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    if (!on_scene_device(*s)) return fail("the scene lives on another device than the current one (cudaSetDevice / rtbvh_gpu_set_device first)");
    if (n != 0 && (!d_origins || !d_directions || !d_hits)) return fail("null device buffer");
    RTB_CUDA(launch_od(*s, *t, tree, false, d_origins, d_directions, n, t_min, t_max, d_hits, nullptr, (cudaStream_t)stream));
    return Ok;
}
// ---- primary rays generated on the device: no ray upload at all ---------------------------------------------------------
// The batch is `frames` frames of width x height camera rays (CameraView3D::generate_ray, shared/src/lib.rs:157-165; the
// arithmetic of rtbvh_gpu_generate_camera_rays_device).  Per chunk of whole frames (or of rows, when one frame exceeds a
// staging slot): ray generation -> traversal in 8x8 tiles -> D2H of the records, on the pipeline streams, so consecutive
// chunks and consecutive submissions overlap like the host-ray calls do.  Only the records cross PCIe.
static ResultCode camera_call(RTGpuScene h, RTTreeKind tree, bool any, const float pos[3], const float p1[3], const float right[3],
                              const float up[3], uint32_t width, uint32_t height, uint64_t seed, uint64_t first_frame,
                              uint32_t frames, void* out, uint64_t* ticket) {
    auto sp = get_scene(h);
    if (!sp) return fail("unknown scene");
    Scene& s = *sp;
    const DeviceTree* t = pick_tree(s, tree);
    if (!t) return fail("scene has no such tree");
    if (!pos || !p1 || !right || !up || !ticket) return fail("null argument");
    const size_t per_frame = (size_t)width * height, total = per_frame * frames;
    if (total != 0 && !out) return fail("null host buffer");
    std::unique_lock<std::mutex> lk(s.pipe_mutex);
    DeviceGuard dg(s.device);
    if (!dg.ok()) return fail("cudaSetDevice", dg.err);
    if (ensure_pipeline(s, gated_slot_rays(total)) != Ok) return Error;
    const uint64_t id = s.next_ticket++;
    Scene::Ticket& tk = s.tickets[id % kTickets];
    if (tk.id != 0)
        for (int i = 0; i < kPipeStreams; i++) RTB_CUDA(cudaEventSynchronize(tk.ev[i]));
    const size_t unit_out = any ? 1 : sizeof(RTHit);
    if (total != 0) {
        // rows per chunk: whole frames while they fit a slot, else a multiple of 8 rows (tile bands)
        size_t rows_per_chunk = s.slot_rays / width;
        if (rows_per_chunk == 0) return fail("camera batch: a single row exceeds the staging slot");
        if (rows_per_chunk >= height) rows_per_chunk = rows_per_chunk / height * height;
        else if (rows_per_chunk >= 8) rows_per_chunk &= ~size_t(7);
        const size_t total_rows = (size_t)height * frames;
        for (size_t row = 0; row < total_rows; row += rows_per_chunk) {
            const size_t rows = std::min(rows_per_chunk, total_rows - row);
            const int k = (int)(s.chunk_seq++ % kPipeStreams);
            cudaStream_t st = s.streams[k];
            RTRay* d_rays = (RTRay*)s.d_in[k];
            for (size_t r0 = row; r0 < row + rows;) {  // one generator launch per frame segment
                const size_t f = r0 / height, y0 = r0 % height;
                const size_t seg = std::min((size_t)height - y0, row + rows - r0);
                RTB_CUDA(launch_camera_rays(pos, p1, right, up, width, height, (uint32_t)y0, (uint32_t)seg, seed, first_frame + f,
                                            d_rays + (r0 - row) * width, st));
                r0 += seg;
            }
            const size_t m = rows * width;
            PeerDests pd{};
            if (width % 8 == 0 && !s.sort_rays.load()) {
                pd.tile_w = width;
                pd.tile_n = (unsigned long long)(m / (8ull * width) * (8ull * width));
                if (pd.tile_n == 0) pd.tile_w = 0;
            }
            RTB_CUDA(launch_trace_single(*t, tree, any, d_rays, m, any ? nullptr : (RTHit*)s.d_out[k], any ? (uint8_t*)s.d_out[k] : nullptr,
                                         s.counter_slot(), s.d_overflow, persistent_mode(), s.sort_bounds(), pd.tile_w ? &pd : nullptr, st));
            RTB_CUDA(cudaMemcpyAsync((char*)out + row * width * unit_out, s.d_out[k], m * unit_out, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int i = 0; i < kPipeStreams; i++) {
        if (!tk.ev[i]) RTB_CUDA(cudaEventCreateWithFlags(&tk.ev[i], cudaEventDisableTiming));
        RTB_CUDA(cudaEventRecord(tk.ev[i], s.streams[i]));
    }
    tk.id = id;
    *ticket = id;
    return Ok;
}
ResultCode rtbvh_gpu_intersect_camera_async(RTGpuScene h, RTTreeKind tree, const float pos[3], const float p1[3], const float right[3],
                                            const float up[3], uint32_t width, uint32_t height, uint64_t jitter_seed,
                                            uint64_t first_frame, uint32_t frames, RTHit* hits, uint64_t* ticket) {
    return camera_call(h, tree, false, pos, p1, right, up, width, height, jitter_seed, first_frame, frames, hits, ticket);
}
ResultCode rtbvh_gpu_occluded_camera_async(RTGpuScene h, RTTreeKind tree, const float pos[3], const float p1[3], const float right[3],
                                           const float up[3], uint32_t width, uint32_t height, uint64_t jitter_seed,
                                           uint64_t first_frame, uint32_t frames, uint8_t* occluded, uint64_t* ticket) {
    return camera_call(h, tree, true, pos, p1, right, up, width, height, jitter_seed, first_frame, frames, occluded, ticket);
}
ResultCode rtbvh_gpu_wait(RTGpuScene h, uint64_t ticket) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    return wait_ticket(*s, ticket);
}
ResultCode rtbvh_gpu_host_alloc(size_t bytes, void** ptr) {
    if (!ptr) return fail("null argument");
    RTB_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return Ok;
}
ResultCode rtbvh_gpu_host_free(void* ptr) {
    RTB_CUDA(cudaFreeHost(ptr));
    return Ok;
}

ResultCode rtbvh_gpu_intersect_packets(RTGpuScene h, RTTreeKind tree, const RTRayPacket4* packets, size_t n,
                                       float t_min, RTHitPacket4* hits) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    return run_host_batch(*s, packets, n, sizeof(RTRayPacket4), sizeof(RTHitPacket4), 4, hits,
                          [&](void* din, size_t m, void* dout, const unsigned long long*, cudaStream_t st) {
                              return launch_trace_packets(*t, tree, false, (const RTRayPacket4*)din, m, t_min,
                                                          (RTHitPacket4*)dout, nullptr, s->counter_slot(), s->d_overflow,
                                                          packet_mode(tree), st);
                          });
}
ResultCode rtbvh_gpu_occluded_packets(RTGpuScene h, RTTreeKind tree, const RTRayPacket4* packets, size_t n,
                                      float t_min, uint8_t* occluded) {
    auto s = get_scene(h);
    if (!s) return fail("unknown scene");
    const DeviceTree* t = pick_tree(*s, tree);
    if (!t) return fail("scene has no such tree");
    return run_host_batch(*s, packets, n, sizeof(RTRayPacket4), 4, 4, occluded,
                          [&](void* din, size_t m, void* dout, const unsigned long long*, cudaStream_t st) {
                              return launch_trace_packets(*t, tree, true, (const RTRayPacket4*)din, m, t_min, nullptr,
                                                          (uint8_t*)dout, s->counter_slot(), s->d_overflow, packet_mode(tree), st);
                          });
}

ResultCode rtbvh_gpu_generate_camera_rays_device(const float pos[3], const float p1[3], const float right[3],
                                                 const float up[3], uint32_t width, uint32_t height, uint32_t row0,
                                                 uint32_t rows, uint64_t jitter_seed, uint64_t frame, RTRay* d_rays,
                                                 void* stream) {
    if (!pos || !p1 || !right || !up || !d_rays) return fail("null argument");
    RTB_CUDA(launch_camera_rays(pos, p1, right, up, width, height, row0, rows, jitter_seed, frame, d_rays,
                                (cudaStream_t)stream));
    return Ok;
}

}  // extern "C"
