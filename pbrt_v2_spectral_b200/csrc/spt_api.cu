// spt_api.cu — the C ABI of include/spt.h: scene upload, wavefront scheduling, film.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false --shared -Xcompiler -fPIC
// There is no CPU path: without a CUDA device every computing entry point returns SPT_ERR_CUDA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ctime>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "launch.h"
#include "wide.h"

struct SptScene;
static void collect_class_times(SptScene *s);
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return SPT_ERR_CUDA; } } while (0)
#define CUP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return nullptr; } } while (0)

namespace {

// Device allocations go through a per-process cache of freed blocks keyed by (device, size): a host
// that renders frame after frame (scene_create -> render -> download -> destroy) asks for the same
// sizes every frame, and cudaMalloc/cudaFree of tens of buffers (the film alone is 65 MB, the wave
// state gigabytes) costs milliseconds and synchronises the device. spt_trim() returns the cache to
// the driver.
struct BlockCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> free_blocks;
    // dev: the device the block lives on (the caller's current device at allocation time; DevMem records it)
    void *get(size_t bytes, int dev) {
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = free_blocks.find({dev, bytes});
            if (it != free_blocks.end()) { void *p = it->second; free_blocks.erase(it); return p; }
        }
        void *p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            trim();                                   // out of memory: give the cache back and retry once
            if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        }
        return p;
    }
    void put(void *p, size_t bytes, int dev) {
        std::lock_guard<std::mutex> lk(mu);
        free_blocks.insert({{dev, bytes}, p});
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu);
        int cur = 0; cudaGetDevice(&cur);
        for (auto &kv : free_blocks) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
        cudaSetDevice(cur);
        free_blocks.clear();
    }
};
BlockCache g_blocks;

// RAII over the current device: handles remember the device they were created on (spt_set_device supports several
// devices per process) and every entry point that touches a handle's memory switches to it first.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct DevMem {
    std::vector<std::pair<void *, size_t>> ptrs;
    int dev = -1;                // the device every block of this holder lives on (set by the first allocation)
    template <typename T> T *alloc(size_t n) {
        if (n == 0) n = 1;
        size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
        if (dev < 0) { dev = 0; cudaGetDevice(&dev); }
        void *p = g_blocks.get(bytes, dev);
        if (!p) return nullptr;
        ptrs.push_back({p, bytes});
        return (T *)p;
    }
    template <typename T> T *upload(const T *host, size_t n, bool settle = true) {
        T *d = alloc<T>(n);
        // cudaMemcpy from pageable memory returns when the data is in the driver's staging buffer - the DMA into `d` may still be
        // in flight, ordered on the legacy stream only. The render lanes are non-blocking streams, which that stream does not
        // order: wait here, or a kernel launched next could read the tail of the array before it lands.
        // (settle = false: the caller uploads a batch and synchronises the legacy stream once)
        if (d && n && host && (cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess ||
                               (settle && cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess))) return nullptr;
        return d;
    }
    // callers synchronise the device (or the stream that last touched the blocks) before releasing
    void release() { for (auto &p : ptrs) g_blocks.put(p.first, p.second, dev); ptrs.clear(); }
};

int num_sms() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

}  // namespace

#define SPT_MAX_LANES 4
#define SPT_MAX_FRAMES 4            // frames in flight on a scene (spt_render_begin)
// Streams and events outlive the scene that created them: a host that builds a scene per frame (the reference's renderer does)
// would otherwise pay 4 stream + ~80 event creations per frame (0.3 ms). Keyed by device; emptied by spt_trim().
#define SPT_POOL_DEVICES 64
struct HandlePool {
    std::mutex mu;
    std::vector<cudaStream_t> streams[SPT_POOL_DEVICES];
    std::vector<cudaEvent_t> timing_events[SPT_POOL_DEVICES], plain_events[SPT_POOL_DEVICES];
    static int slot(int dev) { return dev >= 0 && dev < SPT_POOL_DEVICES ? dev : -1; }
    cudaStream_t stream(int dev) {
        { std::lock_guard<std::mutex> g(mu); int k = slot(dev); if (k >= 0 && !streams[k].empty()) { cudaStream_t st = streams[k].back(); streams[k].pop_back(); return st; } }
        cudaStream_t st = nullptr;
        return cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess ? st : nullptr;
    }
    cudaEvent_t event(int dev, bool timing) {
        { std::lock_guard<std::mutex> g(mu); int k = slot(dev); auto *v = timing ? timing_events : plain_events;
          if (k >= 0 && !v[k].empty()) { cudaEvent_t e = v[k].back(); v[k].pop_back(); return e; } }
        cudaEvent_t e = nullptr;
        return cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming) == cudaSuccess ? e : nullptr;
    }
    // the caller has synchronised the device: nothing is pending on what comes back
    void put(int dev, cudaStream_t st) { if (!st) return; std::lock_guard<std::mutex> g(mu); int k = slot(dev); if (k >= 0 && streams[k].size() < 64) streams[k].push_back(st); else cudaStreamDestroy(st); }
    void put(int dev, cudaEvent_t e, bool timing) {
        if (!e) return;
        std::lock_guard<std::mutex> g(mu); int k = slot(dev); auto *v = timing ? timing_events : plain_events;
        if (k >= 0 && v[k].size() < 4096) v[k].push_back(e); else cudaEventDestroy(e);
    }
    void trim(int dev) {
        std::lock_guard<std::mutex> g(mu); int k = slot(dev); if (k < 0) return;
        for (cudaStream_t st : streams[k]) cudaStreamDestroy(st);
        for (cudaEvent_t e : timing_events[k]) cudaEventDestroy(e);
        for (cudaEvent_t e : plain_events[k]) cudaEventDestroy(e);
        streams[k].clear(); timing_events[k].clear(); plain_events[k].clear();
    }
};
static HandlePool g_handles;

struct SptScene {
    int device = 0;                  // the device the scene lives on: every entry point switches to it (DeviceGuard)
    DevMem mem;
    DevMem trace_scratch;            // spt_trace_*_dev: split rays + counters, grown on demand, released with the scene
    uint64_t trace_scratch_n = 0;
    float4 *ts_ro = nullptr, *ts_rd = nullptr;
    uint32_t *ts_cnt = nullptr, *ts_slot = nullptr;
    DevScene dev;
    std::vector<uint32_t> prim_id_host;
    uint32_t *prim_id_dev = nullptr;
    DevMem wide_mem;                 // fast traversal layout (wide.h), built on the first spt_scene_set_traversal(FAST)
    float4 *wnodes = nullptr;
    uint32_t wroot = 0xffffffffu;
    bool counters_on = false;
    int trace_variant = 1;           // trace_kernels.cuh: 0 reference nodes, 1 pair nodes (default)
    uint32_t fetch_threshold = 14;
    bool merge_trace = true;         // one persistent trace launch per bounce over all its ray queues (SPT_MERGE_TRACE=0: one per queue)
    int max_lanes = 4;               // spt_scene_set_lanes: the most streams a frame's waves are dealt to (1 = one stream: per-kernel
                                     // timing is then exact). A frame that fits two waves uses two lanes - fewer, larger waves win
                                     // (profiles/r01_rank_emulation.log) - a frame of many memory-capped waves uses all of them.
    cudaEvent_t evjoin[SPT_MAX_LANES] = {};
    bool has_env = false;            // an infinite light is present (escaped camera rays pick up Le)
    int direct_slots = 1;            // sum of the lights' n_samples: slots per camera sample under directlighting
    bool direct_pow2 = true;         // every light's n_samples is a power of two (the generated sampler needs it)
    unsigned long long *counters = nullptr;
    // Wave state, allocated on first use (the blocks come back from the block cache frame after frame).
    // Two LANES = two streams, each with its own wave buffers: spt_render deals the waves of a frame to
    // them alternately, so the drain of one wave's persistent trace kernel (a few long rays keep a
    // handful of warps busy for ~50-90 us while the machine empties) is covered by the other wave's
    // kernels instead of being paid 18 times per frame per GPU.
    struct Lane {
        cudaStream_t stream = nullptr;
        DevMem mem;
        WaveBuffers wb{};
        cudaEvent_t last = nullptr;  // previous mark of this lane
        bool have_last = false;
    } lane[SPT_MAX_LANES];
    DevMem counts_mem;
    uint32_t *counts = nullptr;      // device queue lengths: per wave, (max_depth+2) rows of SPT_ROW words
    size_t counts_len = 0;
    cudaStream_t stream = nullptr;   // = lane[0].stream
    cudaStream_t bk = nullptr;       // book-keeping: end-of-frame join of the lanes, frame-end event, counter read-back
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    SptStats stats{};
    bool class_times_pending = false;    // the per-launch event deltas of the last render are folded into stats.class_ms on demand
    uint64_t launches = 0;
    // per-launch timing: an event after every launch of a render, attributed to the launch's class;
    // deltas are taken between consecutive events of the same lane
    struct Mark { cudaEvent_t e; int cls; int lane; };
    std::vector<Mark> marks;
    size_t ev_used = 0;
    bool marks_on = true;            // off for a frame enqueued behind another one (the two would record the same events)
    void mark(int cls, int ln = 0) {
        if (marks_on) {
            if (ev_used == marks.size()) marks.push_back(Mark{g_handles.event(device, true), 0, 0});
            marks[ev_used].cls = cls; marks[ev_used].lane = ln;
            cudaEventRecord(marks[ev_used++].e, lane[ln].stream);
        }
        if (cls >= 0) { ++launches; if (marks_on) ++stats.class_launches[cls]; }
    }
    // Frames in flight (spt_render_begin ... spt_render_end): what spt_render_end needs to finish the frame's statistics. Two
    // records: the next frame is enqueued while the current one runs, so the device goes from one to the other without
    // waiting for the host.
    struct FrameRec {
        bool timed = false, tree = false, empty = false;
        size_t n_waves = 0, per_wave = 0;
        int depth = 0, n_lanes = 0, spp = 0, nranks = 1;
        uint64_t local_tiles = 0;
        RenderCfg cfg;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr, evc = nullptr;
        uint32_t *hc = nullptr;      // page-locked copy of the frame's counter rows
        size_t hc_cap = 0;
        std::vector<uint32_t> hc_tree;
    } frame[SPT_MAX_FRAMES];
    int f_head = 0, f_count = 0;     // oldest record, records in flight
};

static void destroy_lanes(SptScene *s) {
    g_handles.put(s->device, s->bk); s->bk = nullptr;
    for (int k = 0; k < SPT_MAX_LANES; ++k) {
        g_handles.put(s->device, s->lane[k].stream); s->lane[k].stream = nullptr;
        g_handles.put(s->device, s->evjoin[k], false); s->evjoin[k] = nullptr;
    }
}
static bool make_lanes(SptScene *s) {
    if (!(s->bk = g_handles.stream(s->device))) return false;
    for (int k = 0; k < SPT_MAX_LANES; ++k)
        if (!(s->lane[k].stream = g_handles.stream(s->device)) || !(s->evjoin[k] = g_handles.event(s->device, false))) { destroy_lanes(s); return false; }
    return true;
}

struct SptFilm {
    int device = 0;
    SptFilmDesc desc;
    float *pix = nullptr;           // [y][x][NB+1]
    bool owned = true;
    float *table = nullptr;
    void *ipc_base = nullptr;       // pixels opened from another process's film (spt_film_open_ipc): closed on destroy
    float *split = nullptr;         // download staging: [y][x][NB] followed by [y][x] weights
    DevMem mem;                     // owned device blocks (pixels unless external, filter table, staging)
    size_t npix() const { return (size_t)desc.x_pixel_count * desc.y_pixel_count; }
};

extern "C" {

int spt_nbands(void) { return NB; }
const char *spt_last_error(void) { return g_err.c_str(); }
int spt_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
int spt_set_device(int ordinal) { CU(cudaSetDevice(ordinal)); return SPT_OK; }
void *spt_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { g_err = "cudaMallocHost failed"; cudaGetLastError(); return nullptr; }
    return p;
}
void spt_host_free(void *p) { if (p) cudaFreeHost(p); }
void spt_trim(void) {
    cudaDeviceSynchronize();
    g_blocks.trim();
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) g_handles.trim(dev);
}

static double now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

SptScene *spt_scene_create(const SptSceneDesc *d) {
    if (!d) { g_err = "null scene desc"; return nullptr; }
    const bool timing = getenv("SPT_TIMING") != nullptr;
    double t_mark = now_ms();
    auto lap = [&](const char *what) { if (timing) { double t = now_ms(); fprintf(stderr, "[spt_scene_create] %-28s %.3f ms\n", what, t - t_mark); t_mark = t; } };
    if (d->nbands != NB) { g_err = "scene band count does not match the library's SPT_NBANDS"; return nullptr; }
    if (spt_device_count() <= 0) { g_err = "no CUDA device: this library has no CPU path"; return nullptr; }
    if (d->n_materials > 0xffffu || d->n_lights > 0xfffeu) { g_err = "more than 65535 materials or 65534 lights"; return nullptr; }
    // indices the host and the kernels dereference: checked here once, so that a malformed table is an error, not a fault
    {
        auto bad = [&](const char *what) { g_err = std::string("scene tables: ") + what; return (SptScene *)nullptr; };
        if (d->n_prims && (!d->prim_kind || !d->prim_flags || !d->prim_id || !d->prim_data || !d->prim_material || !d->prim_light || !d->prim_xform))
            return bad("a per-slot array is NULL");
        for (uint32_t p = 0; p < d->n_prims; ++p) {
            if (d->prim_material[p] < 0 || (uint32_t)d->prim_material[p] >= d->n_materials) return bad("prim_material out of range");
            if (d->prim_light[p] < -1 || d->prim_light[p] >= (int32_t)d->n_lights) return bad("prim_light out of range");
            if (d->prim_kind[p] == SPT_PRIM_TRIANGLE) { if (d->prim_data[p] >= d->n_tris) return bad("prim_data: triangle number out of range"); }
            else if (d->prim_kind[p] == SPT_PRIM_SPHERE || d->prim_kind[p] == SPT_PRIM_DISK) { if (d->prim_data[p] >= d->n_quadrics) return bad("prim_data: quadric row out of range"); }
            else return bad("unknown primitive kind");
            if (d->prim_xform[p] < 0 || (d->n_xforms && (uint32_t)d->prim_xform[p] >= d->n_xforms)) return bad("prim_xform out of range");
        }
        for (size_t k = 0; k < (size_t)d->n_tris * 3; ++k)
            if (d->tri_vidx[k] < 0 || (uint32_t)d->tri_vidx[k] >= d->n_verts) return bad("tri_vidx out of range");
        for (uint32_t q = 0; q < d->n_quadrics; ++q)
            if (d->quadrics[q].xform < 0 || (uint32_t)d->quadrics[q].xform >= d->n_xforms) return bad("quadric xform out of range");
        for (uint32_t li = 0; li < d->n_lights; ++li) {
            const SptLight &l = d->lights[li];
            if (l.type == SPT_LIGHT_AREA && (l.shape_first < 0 || l.shape_count < 0 || (uint64_t)l.shape_first + (uint64_t)l.shape_count > d->n_light_shapes))
                return bad("light shape range out of bounds");
            if (l.type == SPT_LIGHT_INFINITE && (l.xform < 0 || (uint32_t)l.xform >= d->n_xforms)) return bad("infinite light xform out of range");
        }
        for (uint32_t k = 0; k < d->n_light_shapes; ++k) {
            const SptLightShape &ls = d->light_shapes[k];
            if (ls.kind == SPT_PRIM_TRIANGLE ? (ls.data < 0 || (uint32_t)ls.data >= d->n_tris) : (ls.data < 0 || (uint32_t)ls.data >= d->n_quadrics))
                return bad("light shape data out of range");
        }
    }
    if (NB < NBP) {
        // rows are SPT_BAND_PITCH wide; what lies beyond the valid bands must be zero (the lane-per-band kernels read whole rows)
        auto padded = [](const float *row) { for (int c = NB; c < NBP; ++c) if (row[c] != 0.f) return false; return true; };
        bool ok = padded(d->tables.cie_y);
        for (int k = 0; k < 7; ++k) ok = ok && padded(d->tables.rgb_illum[k]) && padded(d->tables.rgb_refl[k]);
        for (uint32_t k = 0; k < d->n_materials; ++k) ok = ok && padded(d->materials[k].spec0) && padded(d->materials[k].spec1);
        for (uint32_t k = 0; k < d->n_lights; ++k) ok = ok && padded(d->lights[k].spectrum);
        for (uint32_t k = 0; k < d->n_brdf_nodes; ++k) ok = ok && padded(d->brdf_spectra + (size_t)k * NBP);
        if (!ok) { g_err = "a spectrum row has non-zero padding beyond SPT_NBANDS"; return nullptr; }
    }
    SptScene *s = new SptScene();
    cudaGetDevice(&s->device);
    if (const char *e = getenv("SPT_TRACE_VARIANT")) s->trace_variant = atoi(e);
    if (const char *e = getenv("SPT_FETCH_THRESHOLD")) s->fetch_threshold = (uint32_t)atoi(e);
    if (const char *e = getenv("SPT_MERGE_TRACE")) s->merge_trace = atoi(e) != 0;
    if (const char *e = getenv("SPT_LANES")) s->max_lanes = std::min(std::max(atoi(e), 1), SPT_MAX_LANES);
    DevScene &v = s->dev;
    memset(&v, 0, sizeof(v));
    // What the traversal kernels read - leaf flags in the reference nodes, pair nodes, per-slot triangle vertices - is built
    // on the DEVICE from the uploaded arrays (spt_build.cu); SPT_HOST_RELAYOUT=1 keeps the first, host-side builder (1.7 ms per
    // scene on killeroo) for A/B checks. Both give the same bytes.
    const bool host_relayout = getenv("SPT_HOST_RELAYOUT") != nullptr;
    std::vector<uint8_t> nodes;
    std::vector<float4> pn, tv;
    uint32_t root_code = 0xffffffffu;
    bool pairs_ok = d->n_prims < (1u << 27) - 1u;
    if (host_relayout) {
    nodes.assign((const uint8_t *)d->bvh_nodes, (const uint8_t *)d->bvh_nodes + (size_t)d->n_nodes * 32);
    for (uint32_t n = 0; n < d->n_nodes; ++n) {
        uint8_t *nd = &nodes[(size_t)n * 32];
        uint32_t off; memcpy(&off, nd + 24, 4);
        uint8_t np = nd[28];
        nd[30] = 0; nd[31] = 0;
        for (uint32_t i = 0; i < np; ++i)
            if (off + i < d->n_prims && d->prim_kind[off + i] != SPT_PRIM_TRIANGLE) nd[30] = 1;
    }
    lap("node copy + leaf flags");
    // pair nodes: compact array over the interior nodes of the reference's depth-first layout. Child
    // codes pack a leaf's {hasQuadric, nPrims-1, first slot} into one word; a tree that does not fit
    // (leaves of more than 8 primitives, 2^27 primitives) is walked on the reference layout (variant 0).
    {
        struct RefNode { float b[6]; uint32_t off; uint8_t np, axis, hasq, pad; };
        const RefNode *rn = (const RefNode *)nodes.data();
        std::vector<uint32_t> pidx(d->n_nodes, 0xffffffffu);
        uint32_t nint = 0;
        for (uint32_t n = 0; n < d->n_nodes; ++n) { if (rn[n].np == 0) pidx[n] = nint++; else if (rn[n].np > 8) pairs_ok = false; }
        if (nint >= 0x80000000u) pairs_ok = false;
        auto code = [&](uint32_t c) {
            return rn[c].np ? (0x80000000u | (rn[c].hasq ? 0x40000000u : 0u) | ((uint32_t)(rn[c].np - 1) << 27) | rn[c].off) : pidx[c];
        };
        if (pairs_ok) {
            pn.assign((size_t)nint * 4, make_float4(0, 0, 0, 0));
            for (uint32_t n = 0; n < d->n_nodes; ++n) {
                if (rn[n].np) continue;
                uint32_t c0 = n + 1, c1 = rn[n].off;
                if (c0 >= d->n_nodes || c1 >= d->n_nodes) { g_err = "malformed BVH: child index out of range"; s->mem.release(); delete s; return nullptr; }
                float4 *q = &pn[(size_t)pidx[n] * 4];
                const float *a = rn[c0].b, *b = rn[c1].b;
                q[0] = make_float4(a[0], a[1], a[2], a[3]);
                q[1] = make_float4(a[4], a[5], b[0], b[1]);
                q[2] = make_float4(b[2], b[3], b[4], b[5]);
                uint32_t w[4] = { code(c0), code(c1), (uint32_t)(rn[n].axis & 3), 0u };
                memcpy(&q[3], w, 16);
            }
            if (d->n_nodes) root_code = code(0);
        } else s->trace_variant = 0;
    }
    lap("pair nodes");
    // pre-gathered triangle vertices per BVH slot
    tv.assign((size_t)d->n_prims * 3, make_float4(0, 0, 0, 0));
    for (uint32_t p = 0; p < d->n_prims; ++p) {
        if (d->prim_kind[p] != SPT_PRIM_TRIANGLE) continue;
        const int32_t *vi = d->tri_vidx + 3 * (size_t)d->prim_data[p];
        for (int k = 0; k < 3; ++k) {
            const float *P = d->P + 3 * (size_t)vi[k];
            tv[(size_t)p * 3 + k] = make_float4(P[0], P[1], P[2], 0.f);
        }
    }
    }
    // Distribution1D of each area light's ShapeSet (montecarlo.h:48-68), same fp32 recurrence
    std::vector<float> cdf(d->n_light_shapes + d->n_lights + 1, 0.f);
    for (uint32_t li = 0; li < d->n_lights; ++li) {
        const SptLight &l = d->lights[li];
        if (l.type != SPT_LIGHT_AREA) continue;
        float *c = &cdf[l.shape_first + li];
        int n = l.shape_count;
        c[0] = 0.f;
        for (int i = 1; i < n + 1; ++i) c[i] = c[i - 1] + d->light_shapes[l.shape_first + i - 1].area / n;
        float funcInt = c[n];
        if (funcInt == 0.f) for (int i = 1; i < n + 1; ++i) c[i] = float(i) / float(n);
        else for (int i = 1; i < n + 1; ++i) c[i] /= funcInt;
    }
    lap("triangle gather + light cdf");
    DevMem &m = s->mem;
    bool ok = true;
#define UP(dst, src, n) do { dst = m.upload(src, (size_t)(n), false); if (!dst) ok = false; } while (0)
    float4 *nodes4 = nullptr, *tv_dev = nullptr, *pn_dev = nullptr;
    if (host_relayout) {
        UP(nodes4, (const float4 *)nodes.data(), (size_t)d->n_nodes * 2);
        UP(tv_dev, tv.data(), tv.size());
        UP(pn_dev, pn.data(), pn.size());
    } else {
        UP(nodes4, (const float4 *)d->bvh_nodes, (size_t)d->n_nodes * 2);
        tv_dev = m.alloc<float4>((size_t)d->n_prims * 3);
        pn_dev = m.alloc<float4>(((size_t)d->n_nodes / 2 + 1) * 4);       // a binary tree of n nodes has (n - 1) / 2 interior nodes
        if (!tv_dev || !pn_dev) ok = false;
    }
    v.nodes = nodes4; v.tri_verts = tv_dev; v.pnodes = pn_dev;
    v.n_nodes = d->n_nodes; v.n_prims = d->n_prims;
    UP(v.prim_kind, d->prim_kind, d->n_prims); UP(v.prim_flags, d->prim_flags, d->n_prims);
    UP(v.prim_id, d->prim_id, d->n_prims); UP(v.prim_data, d->prim_data, d->n_prims);
    UP(v.prim_material, d->prim_material, d->n_prims); UP(v.prim_light, d->prim_light, d->n_prims);
    UP(v.prim_xform, d->prim_xform, d->n_prims);
    UP(v.tri_vidx, d->tri_vidx, (size_t)d->n_tris * 3);
    UP(v.P, d->P, (size_t)d->n_verts * 3); UP(v.N, d->N, (size_t)d->n_verts * 3); UP(v.UV, d->UV, (size_t)d->n_verts * 2);
    if (!host_relayout && ok) {
        DevMem scratch;
        void *sc_dev = scratch.alloc<uint8_t>(spt_relayout_scratch_bytes(d->n_nodes));
        uint32_t status[2] = { 0, 0xffffffffu };
        cudaError_t e = sc_dev ? spt_launch_relayout(0, nodes4, d->n_nodes, v.prim_kind, v.prim_data, d->n_prims, v.tri_vidx, v.P,
                                                     pn_dev, tv_dev, sc_dev, status) : cudaErrorMemoryAllocation;
        scratch.release();
        if (e != cudaSuccess) { g_err = std::string("scene re-layout failed: ") + cudaGetErrorString(e); m.release(); delete s; return nullptr; }
        if (status[0] & 1u) { g_err = "malformed BVH: child index out of range"; m.release(); delete s; return nullptr; }     // whatever the variant
        if (status[0] & 2u) pairs_ok = false;
        if (pairs_ok) root_code = d->n_nodes ? status[1] : 0xffffffffu; else s->trace_variant = 0;
    }
    v.root_code = root_code;
    UP(v.quadrics, d->quadrics, d->n_quadrics); UP(v.xforms, d->xforms, d->n_xforms);
    // materials: p1 of a mirror / glass row records which of its spectra are not black (the BxDFs GetBSDF adds)
    std::vector<SptMaterial> mats(d->materials, d->materials + d->n_materials);
    for (SptMaterial &mt : mats)
        if (mt.type == SPT_MAT_MIRROR || mt.type == SPT_MAT_GLASS) {
            int mask = 0;
            for (int c = 0; c < NB; ++c) { if (mt.spec0[c] != 0.f) mask |= 1; if (mt.type == SPT_MAT_GLASS && mt.spec1[c] != 0.f) mask |= 2; }
            mt.p1 = (float)mask;
            v.has_specular = 1;
        }
    for (const SptMaterial &mt : mats) {
        if (mt.type == SPT_MAT_SUBSTRATE || mt.type == SPT_MAT_MEASURED || mt.tex_kd >= 0 || mt.tex_bump >= 0) v.has_ext = 1;
        if (mt.type == SPT_MAT_MEASURED) {
            v.has_measured = 1;
            if (mt.brdf < 0 || mt.brdf >= (int32_t)d->n_brdfs) { g_err = "measured material references a BRDF table that is not in the scene"; s->mem.release(); delete s; return nullptr; }
        }
        if (mt.tex_kd >= (int32_t)d->n_textures || mt.tex_bump >= (int32_t)d->n_textures) { g_err = "material references a texture that is not in the scene"; s->mem.release(); delete s; return nullptr; }
        if (mt.tex_kd >= 0 && d->textures[mt.tex_kd].channels != 3) { g_err = "Kd texture is not an RGB image map"; s->mem.release(); delete s; return nullptr; }
        if (mt.tex_bump >= 0 && d->textures[mt.tex_bump].channels != 1) { g_err = "bump texture is not a float image map"; s->mem.release(); delete s; return nullptr; }
    }
    for (uint32_t t = 0; t < d->n_textures; ++t) {
        const SptTexture &tx = d->textures[t];
        if (tx.width <= 0 || tx.height <= 0 || (tx.width & (tx.width - 1)) || (tx.height & (tx.height - 1)) || tx.n_levels < 1) {
            g_err = "image texture: level 0 must have power-of-two sides (MIPMap resamples to them)"; s->mem.release(); delete s; return nullptr;
        }
    }
    if (d->env_w > 0 && ((d->env_w & (d->env_w - 1)) || (d->env_h & (d->env_h - 1)))) {
        g_err = "environment map resolution must be a power of two"; s->mem.release(); delete s; return nullptr;
    }
    if (d->n_textures && !d->ewa_weight_lut) { g_err = "image textures need ewa_weight_lut"; s->mem.release(); delete s; return nullptr; }
    for (uint32_t p = 0; p < d->n_prims; ++p) {
        const SptMaterial &mt = d->materials[d->prim_material[p]];
        if ((mt.tex_kd >= 0 || mt.tex_bump >= 0) && d->prim_kind[p] != SPT_PRIM_TRIANGLE) {
            g_err = "textured materials are supported on triangles only"; s->mem.release(); delete s; return nullptr;
        }
    }
    UP(v.materials, mats.data(), d->n_materials); UP(v.lights, d->lights, d->n_lights);
    UP(v.textures, d->textures, d->n_textures); UP(v.tex_texels, d->tex_texels, d->n_texels);
    UP(v.ewa_lut, d->ewa_weight_lut, d->ewa_weight_lut ? 128 : 0);
    for (uint32_t b = 0; b < d->n_brdfs; ++b)
        if (d->brdfs[b].n_nodes == 0) {      // half-angle table
            const SptBrdfTable &t = d->brdfs[b];
            const uint64_t n = (uint64_t)t.n_theta_h * t.n_theta_d * t.n_phi_d;
            if (n == 0 || t.n_theta_h > 4096 || t.n_theta_d > 4096 || t.n_phi_d > 4096 || t.rgb_offset + 3 * n > d->n_merl_floats || !d->merl_rgb) {
                g_err = "malformed half-angle BRDF table"; s->mem.release(); delete s; return nullptr;
            }
        } else if ((uint64_t)d->brdfs[b].node_first + d->brdfs[b].n_nodes > d->n_brdf_nodes || d->brdfs[b].n_nodes > (1u << 16)) {
            g_err = "malformed BRDF table"; s->mem.release(); delete s; return nullptr;       // 2^16 nodes: the look-up's stack of 32 covers depth 16
        }
    UP(v.brdfs, d->brdfs, d->n_brdfs); UP(v.brdf_nodes, d->brdf_nodes, d->n_brdf_nodes);
    UP(v.brdf_spectra, d->brdf_spectra, (size_t)d->n_brdf_nodes * NBP);
    UP(v.merl_rgb, d->merl_rgb, (size_t)d->n_merl_floats);
    UP(v.light_shapes, d->light_shapes, d->n_light_shapes);
    UP(v.light_cdf, cdf.data(), cdf.size());
    v.n_lights = d->n_lights;
    for (uint32_t li = 0; li < d->n_lights; ++li) if (d->lights[li].type == SPT_LIGHT_INFINITE) s->has_env = true;
    {
        int N = 0;
        for (uint32_t li = 0; li < d->n_lights; ++li) {
            int ns = d->lights[li].n_samples;
            if (ns < 1) { g_err = "light n_samples must be >= 1"; s->mem.release(); delete s; return nullptr; }
            if (ns & (ns - 1)) s->direct_pow2 = false;
            N += ns;
        }
        s->direct_slots = N > 0 ? N : 1;
    }
    UP(v.tables, &d->tables, 1);
    v.env_w = d->env_w; v.env_h = d->env_h;
    size_t ew = (size_t)d->env_w, eh = (size_t)d->env_h;
    UP(v.env_rgb, d->env_rgb, 3 * ew * eh); UP(v.env_func, d->env_func, ew * eh);
    UP(v.env_cdf, d->env_cdf, (ew + 1) * eh); UP(v.env_func_int, d->env_func_int, eh);
    UP(v.env_marg_func, d->env_marg_func, eh); UP(v.env_marg_cdf, d->env_marg_cdf, eh ? eh + 1 : 0);
    v.env_marg_int = d->env_marg_int;
#undef UP
    if (cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) ok = false;      // every table has landed (see DevMem::upload)
    lap("uploads");
    s->counters = m.alloc<unsigned long long>(4);
    if (!ok || !s->counters || cudaMemset(s->counters, 0, 32) != cudaSuccess ||
        !make_lanes(s) ||
        !(s->ev0 = g_handles.event(s->device, true)) || !(s->ev1 = g_handles.event(s->device, true))) {
        g_err = std::string("scene upload failed: ") + cudaGetErrorString(cudaGetLastError());
        destroy_lanes(s);
        if (s->ev0) cudaEventDestroy(s->ev0);
        if (s->ev1) cudaEventDestroy(s->ev1);
        m.release();
        delete s;
        return nullptr;
    }
    v.counters = nullptr;
    s->stream = s->lane[0].stream;
    lap("streams + events");
    return s;
}

void spt_scene_destroy(SptScene *s) {
    if (!s) return;
    DeviceGuard dg(s->device);
    cudaDeviceSynchronize();
    s->mem.release();
    s->wide_mem.release();
    s->trace_scratch.release();
    s->counts_mem.release();
    for (auto &ln : s->lane) ln.mem.release();
    for (auto &fr : s->frame) {
        g_handles.put(s->device, fr.ev0, true); g_handles.put(s->device, fr.ev1, true); g_handles.put(s->device, fr.evc, false);
        if (fr.hc) cudaFreeHost(fr.hc);
    }
    destroy_lanes(s);
    g_handles.put(s->device, s->ev0, true);
    g_handles.put(s->device, s->ev1, true);
    for (auto &m : s->marks) g_handles.put(s->device, m.e, true);
    delete s;
}

int spt_scene_enable_counters(SptScene *s, int on) {
    if (!s) return fail(SPT_ERR_ARG, "null scene");
    DeviceGuard dg(s->device);
    s->counters_on = on != 0;
    s->dev.counters = on ? s->counters : nullptr;
    CU(cudaMemset(s->counters, 0, 32));
    s->stats.node_visits_closest = s->stats.prim_tests_closest = 0;
    s->stats.node_visits_any = s->stats.prim_tests_any = 0;
    return SPT_OK;
}

int spt_scene_set_traversal(SptScene *s, int mode) {
    if (!s || (mode != SPT_TRAVERSAL_EXACT && mode != SPT_TRAVERSAL_FAST)) return fail(SPT_ERR_ARG, "unknown traversal mode");
    DeviceGuard dg(s->device);
    CU(cudaDeviceSynchronize());
    if (mode == SPT_TRAVERSAL_EXACT) { s->dev.wnodes = nullptr; return SPT_OK; }
    if (s->trace_variant == 0) return fail(SPT_ERR_UNSUPP, "this tree does not pack into child codes (leaves of more than 8 primitives): exact traversal only");
    if (!s->wnodes) {
        // collapse the reference's binary tree (already in HBM, leaf flags set by the re-layout) into 4-wide nodes on the host
        std::vector<uint8_t> ref((size_t)s->dev.n_nodes * 32);
        CU(cudaMemcpy(ref.data(), s->dev.nodes, ref.size(), cudaMemcpyDeviceToHost));
        std::vector<W4Node> wide;
        if (!spt_build_w4(ref.data(), s->dev.n_nodes, &wide, &s->wroot)) return fail(SPT_ERR_UNSUPP, "the BVH does not collapse into the wide layout");
        static_assert(sizeof(W4Node) == 128, "W4Node is 8 x float4");
        s->wnodes = s->wide_mem.upload((const float4 *)wide.data(), wide.size() * 8);
        if (!s->wnodes) return fail(SPT_ERR_CUDA, "out of device memory for the wide BVH");
    }
    s->dev.wnodes = s->wnodes; s->dev.wroot = s->wroot;
    return SPT_OK;
}

int spt_scene_set_lanes(SptScene *s, int lanes) {
    if (!s || lanes < 1 || lanes > SPT_MAX_LANES) return fail(SPT_ERR_ARG, "lanes must be 1..4");
    s->max_lanes = lanes;
    return SPT_OK;
}

double spt_last_render_ms(SptScene *s) { return s ? s->stats.render_ms : 0.0; }

int spt_get_stats(SptScene *s, SptStats *out) {
    if (!s || !out) return fail(SPT_ERR_ARG, "null argument");
    DeviceGuard dg(s->device);
    if (s->counters_on) {
        unsigned long long c[4];
        CU(cudaMemcpy(c, s->counters, 32, cudaMemcpyDeviceToHost));
        s->stats.node_visits_closest = c[0]; s->stats.prim_tests_closest = c[1];
        s->stats.node_visits_any = c[2]; s->stats.prim_tests_any = c[3];
    }
    if (s->class_times_pending) collect_class_times(s);
    s->stats.kernel_launches = s->launches;
    *out = s->stats;
    return SPT_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
#define SPT_ROW 16

// directlighting on a scene with specular materials: slots per camera sample, the sample itself + the nodes of its
// SpecularReflect / SpecularTransmit tree (most samples spawn none; a glass object a few per sample that meets it)
#define SPT_TREE_SLOTS 4
static int tree_slots() {                 // SPT_TREE_SLOTS in the environment: tests force the split-and-retry path of spt_render with 2
    if (const char *e = getenv("SPT_TREE_SLOTS")) return std::min(std::max(atoi(e), 2), 64);
    return SPT_TREE_SLOTS;
}
static int ensure_wave(SptScene *s, int n_lanes, uint32_t cap, int max_depth, size_t n_waves, uint32_t sub = 1, int lane_base = 0) {
    cap = (cap + 31u) & ~31u;
    const size_t jcap = (size_t)cap * sub;
    if (jcap > 0xffffffffull) return fail(SPT_ERR_ARG, "wave too large: paths x jobs per path exceeds 2^32 (lower wave_pixels)");
    size_t need_counts = SPT_MAX_FRAMES * n_waves * (size_t)(max_depth + 3) * SPT_ROW;      // frames in flight: one part each
    for (int li = lane_base; li < lane_base + n_lanes; ++li) {
        SptScene::Lane &ln = s->lane[li];
        if (ln.wb.cap >= cap && ln.wb.jcap >= jcap) continue;
        ln.mem.release();
        DevMem &m = ln.mem;
        WaveBuffers &w = ln.wb;
        w.cap = cap; w.jcap = (uint32_t)jcap;
        bool ok = true;
#define AL(field, T, n) do { field = m.alloc<T>((size_t)(n)); if (!field) ok = false; } while (0)
        AL(w.ray_o, float4, cap); AL(w.ray_d, float4, cap); AL(w.hit_slot, uint32_t, cap); AL(w.hit_t, float, cap);
        AL(w.g0, float4, jcap); AL(w.g1, float4, jcap); AL(w.g2, float4, jcap); AL(w.g3, float4, cap);
        AL(w.mis_slot, uint32_t, jcap); AL(w.mis_t, float, jcap); AL(w.sh_slot, uint32_t, jcap);
        AL(w.rec0, float4, jcap); AL(w.rec1, float4, jcap); AL(w.rec2, float4, jcap);
        AL(w.laux, float4, jcap); AL(w.pflags, uint32_t, cap); AL(w.root, uint32_t, s->dev.has_specular ? cap : 1);
        AL(w.rec3, float4, s->dev.has_ext ? jcap : 1); AL(w.rec4, float4, s->dev.has_ext ? jcap : 1);
        AL(w.frow, float, s->dev.has_measured ? jcap * 3 * NBP : 1);
        AL(w.img_xy, float2, cap);
        AL(w.T[0], float, (size_t)cap * NBP); AL(w.T[1], float, (size_t)cap * NBP); AL(w.L, float, (size_t)cap * NBP);
        AL(w.pathQ[0], uint32_t, cap); AL(w.pathQ[1], uint32_t, cap); AL(w.shadowQ, uint32_t, jcap); AL(w.misQ, uint32_t, jcap);
        AL(w.hitQ, uint32_t, cap); AL(w.missQ, uint32_t, cap); AL(w.misAnyQ, uint32_t, jcap);
#undef AL
        if (!ok) { w.cap = 0; w.jcap = 0; m.release(); return fail(SPT_ERR_CUDA, "out of device memory for wave buffers"); }
    }
    if (s->counts_len < need_counts) {
        s->counts_mem.release();
        s->counts = s->counts_mem.alloc<uint32_t>(need_counts);
        if (!s->counts) { s->counts_len = 0; return fail(SPT_ERR_CUDA, "out of device memory for queue counters"); }
        s->counts_len = need_counts;
    }
    return SPT_OK;
}

// Row of device counters of bounce b (SPT_ROW words, zeroed per frame):
//   0 path rays into the bounce, 1 shadow rays, 2 MIS rays traced as closest-hit queries, 3 hits, 4..7 work cursors of the
//   trace launch that carries the bounce's path rays, 8 MIS rays elided (could not reach the light), 9 MIS rays traced as
//   any-hit queries (towards an infinite light), 10 escaped rays
// One trace launch over up to four ray queues (variant chosen per scene; SPT_TRACE_VARIANT / SPT_MERGE_TRACE override).
struct SegList {
    TraceMultiArgs a;
    SegList() { memset(&a, 0, sizeof(a)); }
    void add(bool any, const uint32_t *queue, const uint32_t *count, const float4 *ro, const float4 *rd, uint32_t *out_slot, float *out_t) {
        TraceSeg &g = a.seg[a.nseg++];
        g.queue = queue; g.count = count; g.ro = ro; g.rd = rd; g.out_slot = out_slot; g.out_t = out_t; g.any = any ? 1u : 0u;
    }
};
static void launch_trace(SptScene *s, cudaStream_t st, int grid, SegList &sl, uint32_t *work) {
    sl.a.work = work;
    sl.a.fetch_threshold = s->fetch_threshold;
    spt_launch_trace_multi(s->trace_variant, s->merge_trace, s->counters_on, grid, st, s->dev, sl.a);
}

// Runs one wave: K1, then per bounce b
//   trace   { path rays of b | shadow, MIS, MIS-any rays of b-1 }     one persistent launch
//   K6b     k_addlight(b-1): L += T (Le + Ld)                          (needs the shadow / MIS verdicts just traced)
//   compact hits of b, [escaped rays -> environment light]
//   K5      k_shade(b): records + shadow / MIS rays of b
//   K6a     k_advance(b): T' = T f|cos|/pdf, Russian roulette, path rays of b+1      (not at the last bounce)
// and a last trace { shadow, MIS, MIS-any of max_depth } + k_addlight(max_depth).
static void run_wave(SptScene *s, const RenderCfg &cfg, const SampleSource &src, uint32_t *counts, int li = 0) {
    cudaStream_t st = s->lane[li].stream;
    const WaveBuffers &wb = s->lane[li].wb;
    const DevScene &sc = s->dev;
    int sms = num_sms();
    uint32_t n = cfg.n_samples;
    int gridN = (int)std::min<uint64_t>(((uint64_t)n + 255) / 256, (uint64_t)sms * 16);
    if (gridN < 1) gridN = 1;
    s->mark(-1, li);
    spt_launch_gen_camera(gridN, st, cfg, src, wb, counts + 0);
    s->mark(SPT_K_GEN, li);
    const uint64_t nj = (uint64_t)n * (uint64_t)std::max(cfg.sub, 1);
    int gridT = (int)std::min<uint64_t>(((uint64_t)n + 127) / 128, (uint64_t)sms * 16);
    if (gridT < 1) gridT = 1;
    int gridP = (int)std::min<uint64_t>((nj + 127) / 128, (uint64_t)sms * 8);      // persistent trace kernels: resident blocks only
    if (gridP < 1) gridP = 1;
    int gridC = (int)std::min<uint64_t>(((uint64_t)n + 2047) / 2048, (uint64_t)sms * 8);
    if (gridC < 1) gridC = 1;
    const bool lights = sc.n_lights > 0;
    for (int b = 0; b <= cfg.max_depth + 1; ++b) {
        uint32_t *row = counts + SPT_ROW * b, *prev = b > 0 ? counts + SPT_ROW * (b - 1) : nullptr, *next = counts + SPT_ROW * (b + 1);
        const bool last = b == cfg.max_depth + 1;                                            // only the light rays of max_depth are left
        uint32_t *q = b == 0 ? nullptr : wb.pathQ[b & 1], *qn = wb.pathQ[(b + 1) & 1];      // camera rays: sample order, no queue
        SegList sl;
        if (!last) sl.add(false, q, row + 0, wb.ray_o, wb.ray_d, wb.hit_slot, wb.hit_t);
        if (prev && lights) {
            sl.add(true, wb.shadowQ, prev + 1, wb.g0, wb.g1, wb.sh_slot, nullptr);
            sl.add(false, wb.misQ, prev + 2, wb.g0, wb.g2, wb.mis_slot, wb.mis_t);
            // EstimateDirect's BSDF-sampled ray towards an INFINITE light contributes Le iff it escapes
            // (integrator.cpp:151-158: a hit primitive never is that light): an any-hit query
            if (s->has_env) sl.add(true, wb.misAnyQ, prev + 9, wb.g0, wb.g2, wb.mis_slot, nullptr);
        }
        if (sl.a.nseg) {
            launch_trace(s, st, gridP, sl, row + 4);
            s->mark(SPT_K_TRACE_PATH, li);
            if (!s->merge_trace || s->trace_variant == 0) s->launches += sl.a.nseg - 1;
        }
        if (prev) {
            spt_launch_addlight(gridT, st, sc, cfg, wb, b - 1, wb.hitQ, prev + 3);
            s->mark(SPT_K_ACCUMULATE, li);
        }
        if (last) break;
        // camera rays that escape pick up the environment light (samplerrenderer.cpp:239-243)
        uint32_t *mq = (s->has_env && (b == 0 || sc.has_specular || cfg.tree)) ? wb.missQ : nullptr;
        spt_launch_compact_hits(gridC, st, q, row + 0, wb.hit_slot, wb.hitQ, row + 3, mq, row + 10, (b == 0 && !s->has_env) ? wb.L : nullptr);
        s->mark(SPT_K_SHADE, li);
        if (mq) { spt_launch_miss_env(gridT, st, sc, wb, b, cfg.tree, mq, row + 10); s->mark(SPT_K_SHADE, li); }
        spt_launch_shade(gridT, st, sc, cfg, src, wb, b, wb.hitQ, row + 3, row + 1, row + 2, row + 8, row + 9, qn, next + 0, counts + 11);
        s->mark(SPT_K_SHADE, li);
        if (cfg.tree) {
            // the nodes K5 spawned (SpecularReflect / SpecularTransmit) are the next level's rays; their throughput rows
            if (b < cfg.max_depth) { spt_launch_spawn_T(gridT, st, sc, wb, b, qn, next + 0); s->mark(SPT_K_ADVANCE, li); }
        } else if (b < cfg.max_depth) {
            spt_launch_advance(gridT, st, sc, cfg, wb, b, wb.hitQ, row + 3, qn, next + 0);
            s->mark(SPT_K_ADVANCE, li);
        }
    }
}

// after the stream has drained: fold the per-launch event deltas into stats.class_ms
static void collect_class_times(SptScene *s) {
    s->class_times_pending = false;
    cudaEvent_t last[SPT_MAX_LANES] = {};
    bool seen[SPT_K_CLASSES] = {};
    for (size_t k = 0; k < s->ev_used; ++k) {
        const SptScene::Mark &m = s->marks[k];
        if (m.cls >= 0 && last[m.lane]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, last[m.lane], m.e) == cudaSuccess) {
                s->stats.class_ms[m.cls] += ms;
                // SPT_K_SHADE's first mark is the hit compaction; the kernel itself is the class's longest launch of bounce 0
                if (!seen[m.cls] || (m.cls == SPT_K_SHADE && m.lane == 0 && k < 8 && ms > s->stats.first_launch_ms[m.cls])) s->stats.first_launch_ms[m.cls] = ms;
                seen[m.cls] = true;
            }
        }
        last[m.lane] = m.e;
    }
    s->ev_used = 0;
}
static void reset_class_stats(SptScene *s) {
    s->class_times_pending = false;
    for (int k = 0; k < SPT_K_CLASSES; ++k) {
        s->stats.class_ms[k] = 0.; s->stats.class_launches[k] = 0; s->stats.class_rays[k] = 0;
        s->stats.first_launch_ms[k] = 0.; s->stats.first_launch_units[k] = 0;
    }
    s->stats.mis_rays_elided = 0; s->stats.first_vertices = 0;
    s->ev_used = 0;
}

static void add_ray_stats(SptScene *s, const std::vector<uint32_t> &counts, int max_depth, size_t n_waves) {
    for (size_t w = 0; w < n_waves; ++w)
        for (int b = 0; b <= max_depth; ++b) {
            const uint32_t *row = &counts[(w * (size_t)(max_depth + 3) + b) * SPT_ROW];
            s->stats.closest_rays += row[0] + row[2];
            s->stats.any_rays += row[1];
            s->stats.class_rays[SPT_K_TRACE_PATH] += row[0];
            s->stats.class_rays[SPT_K_SHADE] += row[3];
            s->stats.class_rays[SPT_K_ACCUMULATE] += row[3];
            if (b < max_depth) s->stats.class_rays[SPT_K_ADVANCE] += row[3];
            s->stats.class_rays[SPT_K_TRACE_SHADOW] += row[1];
            s->stats.class_rays[SPT_K_TRACE_MIS] += row[2] + row[9];
            s->stats.any_rays += row[9];
            s->stats.mis_rays_elided += row[8];
            if (b == 0) s->stats.first_vertices += row[3];
        }
}

extern "C" {

int spt_camera_rays(const SptCameraDesc *cam, const float *samples, uint64_t n, float *out_rays) {
    if (!cam || !samples || !out_rays) return fail(SPT_ERR_ARG, "null argument");
    if (spt_device_count() <= 0) return fail(SPT_ERR_CUDA, "no CUDA device: this library has no CPU path");
    if (n == 0) return SPT_OK;
    DevMem m;
    float *ds = m.upload(samples, n * 5), *dr = m.alloc<float>(n * 8);
    if (!ds || !dr) { m.release(); return fail(SPT_ERR_CUDA, "device allocation failed"); }
    spt_launch_camera_rays(0, *cam, ds, (uint32_t)n, dr);
    cudaError_t e = cudaMemcpy(out_rays, dr, n * 8 * sizeof(float), cudaMemcpyDeviceToHost);
    m.release();
    if (e != cudaSuccess) return fail(SPT_ERR_CUDA, cudaGetErrorString(e));
    return SPT_OK;
}

static int trace_dev(SptScene *s, bool any, const float4 *ro, const float4 *rd, uint64_t n, uint32_t *slot, float *t,
                     uint32_t *count_dev) {
    int sms = num_sms();
    int grid = (int)std::min<uint64_t>((n + 127) / 128, (uint64_t)sms * 8);
    if (grid < 1) grid = 1;
    cudaStream_t st = s->stream;
    cudaEventRecord(s->ev0, st);
    SegList sl;
    sl.add(any, nullptr, count_dev, ro, rd, slot, t);
    launch_trace(s, st, grid, sl, count_dev + 1);
    cudaEventRecord(s->ev1, st);
    s->launches += 1;
    CU(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev0, s->ev1);
    s->stats.trace_ms = ms;
    if (any) s->stats.any_rays += n; else s->stats.closest_rays += n;
    return SPT_OK;
}

#define SPT_NOT_IN_FLIGHT(s) do { if ((s)->f_count) return fail(SPT_ERR_ARG, "a frame is in flight on this scene: spt_render_end first"); } while (0)
static int trace_host(SptScene *s, bool any, const float *rays, uint64_t n, uint32_t *out_slot, uint32_t *out_id,
                      float *out_t, uint8_t *out_hit) {
    if (!s || !rays) return fail(SPT_ERR_ARG, "null argument");
    SPT_NOT_IN_FLIGHT(s);
    if (n == 0) return SPT_OK;
    if (n > 0x7fffffffull) return fail(SPT_ERR_ARG, "too many rays for one call");
    DeviceGuard dg(s->device);
    DevMem m;
    float *dr = m.upload(rays, n * 8);
    float4 *ro = m.alloc<float4>(n), *rd = m.alloc<float4>(n);
    uint32_t *slot = m.alloc<uint32_t>(n), *ids = m.alloc<uint32_t>(n), *cnt = m.alloc<uint32_t>(2);
    float *t = m.alloc<float>(n);
    uint8_t *flag = m.alloc<uint8_t>(n);
    if (!dr || !ro || !rd || !slot || !ids || !cnt || !t || !flag) { m.release(); return fail(SPT_ERR_CUDA, "device allocation failed"); }
    uint32_t n32 = (uint32_t)n;
    uint32_t cw[2] = { n32, 0 };
    cudaMemcpyAsync(cnt, cw, 8, cudaMemcpyHostToDevice, s->stream);
    spt_launch_split_rays(s->stream, dr, n32, ro, rd);
    int rc = trace_dev(s, any, ro, rd, n, slot, t, cnt);
    if (rc == SPT_OK) {
        cudaError_t e = cudaSuccess;
        if (any) {
            spt_launch_slot_to_flag(s->stream, slot, n32, flag);
            if (out_hit) e = cudaMemcpyAsync(out_hit, flag, n, cudaMemcpyDeviceToHost, s->stream);
        } else {
            spt_launch_slot_to_id(s->stream, slot, s->dev.prim_id, n32, ids);
            if (out_slot) e = cudaMemcpyAsync(out_slot, slot, n * 4, cudaMemcpyDeviceToHost, s->stream);
            if (e == cudaSuccess && out_id) e = cudaMemcpyAsync(out_id, ids, n * 4, cudaMemcpyDeviceToHost, s->stream);
            if (e == cudaSuccess && out_t) e = cudaMemcpyAsync(out_t, t, n * 4, cudaMemcpyDeviceToHost, s->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) rc = fail(SPT_ERR_CUDA, cudaGetErrorString(e));
    }
    m.release();
    return rc;
}

int spt_trace_closest(SptScene *s, const float *rays, uint64_t n, uint32_t *out_slot, uint32_t *out_prim_id, float *out_t) {
    return trace_host(s, false, rays, n, out_slot, out_prim_id, out_t, nullptr);
}
int spt_trace_any(SptScene *s, const float *rays, uint64_t n, uint8_t *out_hit) {
    return trace_host(s, true, rays, n, nullptr, nullptr, nullptr, out_hit);
}

// Device-resident variants: rays_dev is n x 8 floats in HBM; outputs are device arrays.
static int trace_resident(SptScene *s, bool any, const float *rays_dev, uint64_t n, uint32_t *slot_dev, float *t_dev,
                          uint8_t *hit_dev) {
    if (!s || !rays_dev) return fail(SPT_ERR_ARG, "null argument");
    SPT_NOT_IN_FLIGHT(s);
    if (n == 0) return SPT_OK;
    if (n > 0x7fffffffull) return fail(SPT_ERR_ARG, "too many rays for one call");
    DeviceGuard dg(s->device);
    // scratch (split rays + count) belongs to the scene: same device, released by spt_scene_destroy
    DevMem &scratch = s->trace_scratch;
    uint64_t &scratch_n = s->trace_scratch_n;
    float4 *&ro = s->ts_ro, *&rd = s->ts_rd;
    uint32_t *&cnt = s->ts_cnt, *&slot_tmp = s->ts_slot;
    if (scratch_n < n) {
        cudaStreamSynchronize(s->stream);
        scratch.release();
        ro = scratch.alloc<float4>(n); rd = scratch.alloc<float4>(n); cnt = scratch.alloc<uint32_t>(2);
        slot_tmp = scratch.alloc<uint32_t>(n);
        if (!ro || !rd || !cnt || !slot_tmp) { scratch.release(); scratch_n = 0; return fail(SPT_ERR_CUDA, "device allocation failed"); }
        scratch_n = n;
    }
    uint32_t n32 = (uint32_t)n;
    uint32_t cw[2] = { n32, 0 };
    CU(cudaMemcpyAsync(cnt, cw, 8, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));        // cw is on this frame's stack
    spt_launch_split_rays(s->stream, rays_dev, n32, ro, rd);
    uint32_t *slot = any ? slot_tmp : (slot_dev ? slot_dev : slot_tmp);
    int rc = trace_dev(s, any, ro, rd, n, slot, t_dev, cnt);
    if (rc != SPT_OK) return rc;
    if (any && hit_dev) { spt_launch_slot_to_flag(s->stream, slot, n32, hit_dev); CU(cudaStreamSynchronize(s->stream)); }
    return SPT_OK;
}
int spt_trace_closest_dev(SptScene *s, const float *rays_dev, uint64_t n, uint32_t *out_slot_dev, float *out_t_dev) {
    if (!out_t_dev) return fail(SPT_ERR_ARG, "out_t_dev is required");
    return trace_resident(s, false, rays_dev, n, out_slot_dev, out_t_dev, nullptr);
}
int spt_trace_any_dev(SptScene *s, const float *rays_dev, uint64_t n, uint8_t *out_hit_dev) {
    return trace_resident(s, true, rays_dev, n, nullptr, nullptr, out_hit_dev);
}

int spt_shade_samples(SptScene *s, const SptCameraDesc *cam, int32_t integrator, int32_t max_depth, int32_t spp, const float *samples,
                      const float *rng, int32_t n_rng, uint64_t n, float *out_L) {
    if (!s || !cam || !samples || !out_L) return fail(SPT_ERR_ARG, "null argument");
    SPT_NOT_IN_FLIGHT(s);
    if (spp <= 0) return fail(SPT_ERR_ARG, "spp must be positive");
    if (integrator != SPT_INTEGRATOR_PATH && integrator != SPT_INTEGRATOR_DIRECT_ALL && integrator != SPT_INTEGRATOR_DIRECT_ONE)
        return fail(SPT_ERR_ARG, "unknown integrator");
    const bool direct = integrator == SPT_INTEGRATOR_DIRECT_ALL, directOne = integrator == SPT_INTEGRATOR_DIRECT_ONE;
    if (directOne && s->dev.has_specular) return fail(SPT_ERR_UNSUPP, "directlighting strategy \"one\" with specular materials is not supported");
    if (direct && s->dev.n_lights == 0) return fail(SPT_ERR_UNSUPP, "directlighting needs at least one light");
    if (n == 0) return SPT_OK;
    const int sub = direct ? s->direct_slots : 1;
    const int stride = direct ? 7 + 6 * s->direct_slots : (directOne ? 14 : 37);
    // directlighting: max_depth is DirectLightingIntegrator::maxDepth, the depth of the SpecularReflect / SpecularTransmit recursion
    // (directlighting.cpp:97-107); without specular materials there is one level. Strategy "one" is the path integrator's first
    // vertex: emitted light + UniformSampleOneLight.
    const bool tree = direct && s->dev.has_specular && max_depth > 1;
    max_depth = tree ? max_depth - 1 : ((direct || directOne) ? 0 : max_depth);
    // this entry point (tests, diagnostics) sizes the node pool for the full binary tree of up to five levels: no retries here
    const uint64_t slots = n * (uint64_t)(tree ? std::min(31, (2 << std::min(max_depth, 4)) - 1) : 1);
    if (slots * (uint64_t)sub > (1u << 26)) return fail(SPT_ERR_ARG, "too many samples for one call");
    DeviceGuard dg(s->device);
    int rc = ensure_wave(s, 1, (uint32_t)slots, max_depth, 1, (uint32_t)sub);
    if (rc != SPT_OK) return rc;
    DevMem m;
    float *dsmp = m.upload(samples, n * (size_t)stride);
    float *drng = (rng && n_rng > 0) ? m.upload(rng, n * (size_t)n_rng) : nullptr;
    float *dout = m.alloc<float>(n * NB);
    if (!dsmp || !dout || (rng && n_rng > 0 && !drng)) { m.release(); return fail(SPT_ERR_CUDA, "device allocation failed"); }
    RenderCfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.cam = *cam; cfg.spp = 1; cfg.spp_shift = 0; cfg.max_depth = max_depth; cfg.n_samples = (uint32_t)n;
    cfg.integrator = integrator; cfg.sub = sub; cfg.tree = tree ? 1 : 0;
    cfg.tile = 1; cfg.tile_shift = 0; cfg.tilesX = 1; cfg.tilesY = 1; cfg.nranks = 1;
    cfg.diff_scale = 1.f / sqrtf((float)spp);
    SampleSource src; src.smp = dsmp; src.stride = stride; src.rng = drng; src.n_rng = drng ? n_rng : 0; src.seed = 0; src.spp = 1;
    size_t nc = (size_t)(max_depth + 3) * SPT_ROW;
    cudaMemsetAsync(s->counts, 0, nc * 4, s->stream);
    reset_class_stats(s);
    run_wave(s, cfg, src, s->counts);
    spt_launch_gather_L(s->stream, s->lane[0].wb.L, s->lane[0].wb.cap, (uint32_t)n, dout);
    std::vector<uint32_t> hc(nc);
    cudaMemcpyAsync(hc.data(), s->counts, nc * 4, cudaMemcpyDeviceToHost, s->stream);
    cudaError_t e = cudaMemcpyAsync(out_L, dout, n * NB * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    m.release();
    if (e != cudaSuccess) return fail(SPT_ERR_CUDA, cudaGetErrorString(e));
    if (tree && hc[12]) return fail(SPT_ERR_UNSUPP, "specular tree: more nodes than the full binary tree of five levels per sample");
    s->class_times_pending = true;
    s->stats.camera_samples += n;
    add_ray_stats(s, hc, max_depth, 1);
    return SPT_OK;
}

// ---- film ---------------------------------------------------------------------------------------
static SptFilm *film_new(const SptFilmDesc *d, float *ext) {
    if (!d) { g_err = "null film desc"; return nullptr; }
    if (d->x_pixel_count <= 0 || d->y_pixel_count <= 0) { g_err = "empty film"; return nullptr; }
    if (spt_device_count() <= 0) { g_err = "no CUDA device: this library has no CPU path"; return nullptr; }
    SptFilm *f = new SptFilm();
    cudaGetDevice(&f->device);
    f->desc = *d;
    f->owned = ext == nullptr;
    f->pix = ext;
    size_t bytes = f->npix() * (NB + 1) * sizeof(float);
    if (!ext) {
        f->pix = f->mem.alloc<float>(bytes / sizeof(float));
        // the lanes are non-blocking streams: the zeroing must have FINISHED before a render may add to the film
        if (!f->pix || cudaMemset(f->pix, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            g_err = std::string("film allocation failed: ") + cudaGetErrorString(cudaGetLastError());
            f->mem.release(); delete f; return nullptr;
        }
    }
    f->table = f->mem.upload(d->filter_table, 256);
    if (!f->table) {
        g_err = "film filter table upload failed";
        f->mem.release(); delete f; return nullptr;
    }
    return f;
}
SptFilm *spt_film_create(const SptFilmDesc *d) { return film_new(d, nullptr); }
SptFilm *spt_film_create_external(const SptFilmDesc *d, float *pixels_dev) {
    if (!pixels_dev) { g_err = "null device buffer"; return nullptr; }
    return film_new(d, pixels_dev);
}
void spt_film_destroy(SptFilm *f) {
    if (!f) return;
    DeviceGuard dg(f->device);
    cudaDeviceSynchronize();
    f->mem.release();
    if (f->ipc_base) cudaIpcCloseMemHandle(f->ipc_base);
    delete f;
}
int spt_film_clear(SptFilm *f) {
    if (!f) return fail(SPT_ERR_ARG, "null film");
    DeviceGuard dg(f->device);
    // ordered against every stream of the device on both sides (the render lanes do not synchronise with the legacy stream)
    CU(cudaDeviceSynchronize());
    CU(cudaMemset(f->pix, 0, f->npix() * (NB + 1) * sizeof(float)));
    CU(cudaDeviceSynchronize());
    return SPT_OK;
}
int spt_film_clear_idle(SptFilm *f) {
    if (!f) return fail(SPT_ERR_ARG, "null film");
    DeviceGuard dg(f->device);
    cudaStream_t st = g_handles.stream(f->device);
    if (!st) return fail(SPT_ERR_CUDA, "stream creation failed");
    cudaError_t e = cudaMemsetAsync(f->pix, 0, f->npix() * (NB + 1) * sizeof(float), st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    g_handles.put(f->device, st);
    if (e != cudaSuccess) return fail(SPT_ERR_CUDA, cudaGetErrorString(e));
    return SPT_OK;
}
float *spt_film_device_ptr(SptFilm *f) { return f ? f->pix : nullptr; }

int spt_film_download(SptFilm *f, float *c, float *weight) {
    if (!f) return fail(SPT_ERR_ARG, "null film");
    DeviceGuard dg(f->device);
    size_t np = f->npix();
    if (!f->split) f->split = f->mem.alloc<float>(np * (NB + 1));
    if (!f->split) return fail(SPT_ERR_CUDA, "out of device memory for the film staging buffer");
    // [y][x][NB+1] -> {[y][x][NB], [y][x]} on the device, then one copy per host array (DMA at full
    // rate when the caller's buffers are page-locked, spt_host_alloc)
    CU(cudaDeviceSynchronize());
    unsigned g = (unsigned)std::min<size_t>((np * (NB + 1) + 255) / 256, (size_t)num_sms() * 16);
    spt_launch_film_split((int)g, 0, f->pix, np, f->split, f->split + np * NB);
    if (c) CU(cudaMemcpy(c, f->split, np * NB * sizeof(float), cudaMemcpyDeviceToHost));
    if (weight) CU(cudaMemcpy(weight, f->split + np * NB, np * sizeof(float), cudaMemcpyDeviceToHost));
    CU(cudaGetLastError());
    return SPT_OK;
}

int spt_film_write_dat(SptFilm *f, const char *path) {
    if (!f || !path) return fail(SPT_ERR_ARG, "null argument");
    size_t np = f->npix();
    std::vector<float> c(np * NB);
    int rc = spt_film_download(f, c.data(), nullptr);
    if (rc != SPT_OK) return rc;
    FILE *fp = fopen(path, "wb");
    if (!fp) return fail(SPT_ERR_IO, std::string("cannot open ") + path);
    int W = f->desc.x_pixel_count, H = f->desc.y_pixel_count;
    // spectralImage.cpp:344-352: dimensions line, then the lens line (focalLength fStop fieldOfView;
    // uninitialised in the reference for non-lens cameras, SURVEY.md F5 — zeros here)
    fprintf(fp, "%d %d %d\n", W, H, NB);
    fprintf(fp, "0 0 0\n");
    // spectralImage.cpp:283-296,320-369: clamp at zero, no division by weightSum, [band][x][y] doubles
    std::vector<double> plane(np);
    for (int b = 0; b < NB; ++b) {
        for (int x = 0; x < W; ++x)
            for (int y = 0; y < H; ++y) {
                float v = c[((size_t)y * W + x) * NB + b];
                plane[(size_t)x * H + y] = (double)(v > 0.f ? v : 0.f);
            }
        fwrite(plane.data(), sizeof(double), np, fp);
    }
    fclose(fp);
    return SPT_OK;
}

int spt_film_add_samples(SptFilm *f, const SptSpectralTables *tables, const float *image_xy, const float *L, uint64_t n) {
    if (!f || !tables || !image_xy || !L) return fail(SPT_ERR_ARG, "null argument");
    if (n == 0) return SPT_OK;
    if (n > 0x7fffffffull) return fail(SPT_ERR_ARG, "too many samples for one call");
    DeviceGuard dg(f->device);
    DevMem m;
    SptSpectralTables *dt = m.upload(tables, 1);
    float2 *dxy = (float2 *)m.upload(image_xy, n * 2);
    float *dl = m.upload(L, n * NB), *soa = m.alloc<float>(((n + 31) / 32 * 32) * NBP);
    if (!dt || !dxy || !dl || !soa) { m.release(); return fail(SPT_ERR_CUDA, "device allocation failed"); }
    spt_launch_scatter_L(0, dl, (uint32_t)n, (uint32_t)n, soa);
    FilmView fv; fv.d = f->desc; fv.pix = f->pix; fv.table = f->table;
    unsigned gw = (unsigned)std::min<uint64_t>((n * 32 + 255) / 256, (uint64_t)num_sms() * 16);
    spt_launch_film_add((int)gw, 0, fv, dt, dxy, soa, (uint32_t)n, (uint32_t)n, 1);
    cudaError_t e = cudaDeviceSynchronize();
    m.release();
    if (e != cudaSuccess) return fail(SPT_ERR_CUDA, cudaGetErrorString(e));
    return SPT_OK;
}

// ---- the whole job -------------------------------------------------------------------------------
// pipelined: the caller keeps frames in flight (spt_render_begin); false: one frame, the host waits for it (spt_render)
static int render_begin(SptScene *s, const SptCameraDesc *cam, SptFilm *film, const SptRenderParams *rp, bool pipelined) {
    if (!s || !cam || !film || !rp) return fail(SPT_ERR_ARG, "null argument");
    if (film->device != s->device) return fail(SPT_ERR_ARG, "scene and film live on different devices");
    if (s->f_count >= SPT_MAX_FRAMES) return fail(SPT_ERR_ARG, "four frames are already in flight: spt_render_end first");
    DeviceGuard dg(s->device);
    if (rp->spp <= 0 || (rp->spp & (rp->spp - 1))) return fail(SPT_ERR_ARG, "spp must be a power of two (LDSampler rounds up)");
    if (rp->max_depth < 0 || rp->max_depth > 64) return fail(SPT_ERR_ARG, "max_depth out of range");
    int nranks = rp->tile_nranks > 0 ? rp->tile_nranks : 1;
    if (rp->tile_rank < 0 || rp->tile_rank >= nranks) return fail(SPT_ERR_ARG, "tile_rank out of range");
    if (rp->integrator != SPT_INTEGRATOR_PATH && rp->integrator != SPT_INTEGRATOR_DIRECT_ALL && rp->integrator != SPT_INTEGRATOR_DIRECT_ONE)
        return fail(SPT_ERR_ARG, "unknown integrator");
    const bool direct = rp->integrator == SPT_INTEGRATOR_DIRECT_ALL, directOne = rp->integrator == SPT_INTEGRATOR_DIRECT_ONE;
    if (directOne && s->dev.has_specular) return fail(SPT_ERR_UNSUPP, "directlighting strategy \"one\" with specular materials is not supported");
    if (direct && s->dev.has_specular && s->dev.has_ext) return fail(SPT_ERR_UNSUPP, "directlighting with specular AND textured / substrate / measured materials is not supported");
    if (direct && s->dev.n_lights == 0) return fail(SPT_ERR_UNSUPP, "directlighting needs at least one light");
    if (direct && !s->direct_pow2) return fail(SPT_ERR_ARG, "directlighting: every light's n_samples must be a power of two (Sampler::RoundSize)");
    const int sub = direct ? s->direct_slots : 1;                 // jobs per camera hit
    RenderCfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    // directlighting: one level, or - on scenes with specular materials - the levels of its SpecularReflect / SpecularTransmit tree
    const bool tree = direct && s->dev.has_specular && rp->max_depth > 1;
    cfg.cam = *cam; cfg.spp = rp->spp; cfg.max_depth = tree ? rp->max_depth - 1 : ((direct || directOne) ? 0 : rp->max_depth);
    cfg.integrator = rp->integrator; cfg.sub = sub; cfg.tree = tree ? 1 : 0;
    cfg.diff_scale = 1.f / sqrtf((float)rp->spp);
    for (cfg.spp_shift = 0; (1 << cfg.spp_shift) < cfg.spp; ++cfg.spp_shift) {}
    cfg.x0 = rp->x_start; cfg.y0 = rp->y_start; cfg.x1 = rp->x_end; cfg.y1 = rp->y_end;
    const SptFilmDesc &fd = film->desc;
    if (rp->skip_border) {
        cfg.x1 = std::min(cfg.x1, fd.x_pixel_start + fd.x_pixel_count);
        cfg.y1 = std::min(cfg.y1, fd.y_pixel_start + fd.y_pixel_count);
    }
    if (cfg.x1 <= cfg.x0 || cfg.y1 <= cfg.y0) return fail(SPT_ERR_ARG, "empty sample extent");
    cfg.tile = rp->tile_size > 0 ? rp->tile_size : 32;
    if ((cfg.tile & (cfg.tile - 1)) || cfg.tile > 4096) return fail(SPT_ERR_ARG, "tile_size must be a power of two, at most 4096");
    for (cfg.tile_shift = 0; (1 << cfg.tile_shift) < cfg.tile; ++cfg.tile_shift) {}
    cfg.tilesX = (cfg.x1 - cfg.x0 + cfg.tile - 1) / cfg.tile;
    cfg.tilesY = (cfg.y1 - cfg.y0 + cfg.tile - 1) / cfg.tile;
    cfg.rank = rp->tile_rank; cfg.nranks = nranks;
    cfg.seed = (uint32_t)(rp->seed ^ (rp->seed >> 32));
    uint64_t ntiles = (uint64_t)cfg.tilesX * cfg.tilesY;
    uint64_t local_tiles = ntiles > (uint64_t)cfg.rank ? (ntiles - cfg.rank + nranks - 1) / nranks : 0;
    uint64_t local_pixels = local_tiles * (uint64_t)cfg.tile * cfg.tile;
    const int slot = (s->f_head + s->f_count) % SPT_MAX_FRAMES;            // frame records: a ring
    SptScene::FrameRec &fr = s->frame[slot];
    fr.timed = s->f_count == 0; fr.tree = false; fr.empty = false; fr.hc_tree.clear();
    if (local_pixels == 0) {            // a rank that owns no tile (more ranks than tiles): nothing to render
        fr.empty = true;
        ++s->f_count;
        return SPT_OK;
    }
    // Waves: by default the rank's pixels are cut into a multiple of max_lanes waves, each at most
    // 2^25 / max_lanes paths (17 GB of state over all lanes), dealt to the lanes in turn; small jobs
    // (< 2^19 paths per wave) use fewer lanes, down to one wave on one lane.
    const uint64_t slots_pp = (uint64_t)rp->spp * (tree ? tree_slots() : 1);     // path slots per pixel (tree: the samples + their nodes)
    const uint64_t mem_pp = slots_pp * (uint64_t)std::max(1, (sub + 3) / 4);   // directlighting: ~140 B per job on top of ~450 B per path
    const int depth = cfg.max_depth;
    const uint64_t local_samples = local_pixels * slots_pp;
    int want_lanes = s->max_lanes;
    if (want_lanes > 2 && local_pixels * mem_pp <= (1ull << 25)) want_lanes = 2;
    while (want_lanes > 1 && local_samples < ((uint64_t)want_lanes << 19)) --want_lanes;
    if (const char *e = getenv("SPT_FORCE_LANES")) want_lanes = std::min(std::max(atoi(e), 1), s->max_lanes);      // A/B runs (profiles/tools)
    // Frames kept in flight on a small share of the image (several GPUs): every launch pays a fixed drain tail - the longest
    // ray of a persistent trace launch is ~50-100 us of dependent L2 round trips whatever the launch's size - so a frame is
    // ONE wave (half as many launches as two half-size waves) and consecutive frames go to different lanes, which keeps two
    // or three independent kernel chains on the device as the two waves of one frame did. Rank 0's tile set of 8 GPUs, one
    // GPU: 5.50 ms per frame as two waves, 5.17 as one wave with two frames in flight, 5.07 with three
    // (profiles/r02_pipe_modes.log, r02_pipe_depth.log; SPT_PIPE_MODE=0 restores the two-wave frames).
    int lane_base = 0, rot = 0;             // rot: lanes the one-wave frames rotate over (their wave buffers are sized together)
    static const int pipe_mode = [] { const char *e = getenv("SPT_PIPE_MODE"); return e ? atoi(e) : 1; }();
    static const int pipe_lanes = [] { const char *e = getenv("SPT_PIPE_LANES"); return e ? std::min(std::max(atoi(e), 1), SPT_MAX_LANES) : 4; }();
    if (pipelined && !tree && s->max_lanes >= 2 && rp->wave_pixels <= 0) {
        // up to 2^23 paths per frame (config 1 on four GPUs or more): above that the tails are < 1 % of a frame (2 GPUs: 19.70 ms
        // as two waves, 19.55 as one), and the wave state stays at most 2^25 paths over the lanes
        if (pipe_mode == 1 && local_pixels * mem_pp <= (1ull << 23)) {
            want_lanes = 1;
            rot = std::min(s->max_lanes, pipe_lanes);
            lane_base = slot % rot;
        }
    }
    const uint64_t lane_cap_pixels = std::max<uint64_t>(1, ((1u << 25) / (uint64_t)want_lanes) / mem_pp);
    uint64_t wave_pixels;
    if (rp->wave_pixels > 0) wave_pixels = (uint64_t)rp->wave_pixels;
    else {
        uint64_t nw = std::max<uint64_t>((uint64_t)want_lanes, (local_pixels + lane_cap_pixels - 1) / lane_cap_pixels);
        nw = (nw + want_lanes - 1) / want_lanes * want_lanes;
        wave_pixels = (local_pixels + nw - 1) / nw;
    }
    wave_pixels = std::max<uint64_t>(1, std::min<uint64_t>(wave_pixels, local_pixels));
    if (wave_pixels * slots_pp > (1ull << 27)) wave_pixels = std::max<uint64_t>(1, (1ull << 27) / slots_pp);
    if (wave_pixels * slots_pp * (uint64_t)sub > 0xffffffffull) wave_pixels = std::max<uint64_t>(1, 0xffffffffull / (slots_pp * (uint64_t)sub));
    size_t n_waves;
    int n_lanes, rc;
    // a frame is still running in the wave buffers: they may only be re-allocated (a bigger frame) once it has drained; the
    // specular tree checks every pixel range on the host, so it starts on an idle device as well
    if (s->f_count > 0) {
        const size_t nw0 = (size_t)((local_pixels + wave_pixels - 1) / wave_pixels);
        const int nl0 = tree ? 1 : (int)std::min<size_t>((size_t)(s->max_lanes - lane_base), std::max<size_t>(nw0, 1));
        const uint32_t cap0 = ((uint32_t)(wave_pixels * slots_pp) + 31u) & ~31u;
        bool grow = tree || s->counts_len < SPT_MAX_FRAMES * nw0 * (size_t)(depth + 3) * SPT_ROW;
        for (int li = rot ? 0 : lane_base; li < (rot ? rot : lane_base + nl0); ++li) if (s->lane[li].wb.cap < cap0 || s->lane[li].wb.jcap < (size_t)cap0 * sub) grow = true;
        if (grow) CU(cudaDeviceSynchronize());
    }
    for (;;) {
        n_waves = (size_t)((local_pixels + wave_pixels - 1) / wave_pixels);
        n_lanes = tree ? 1 : (int)std::min<size_t>((size_t)(s->max_lanes - lane_base), std::max<size_t>(n_waves, 1));     // tree: every range is checked before it reaches the film
        if (rot) n_lanes = 1;              // a frame's waves (more than one only on a part short of memory) stay on its lane
        rc = ensure_wave(s, rot ? rot : n_lanes, (uint32_t)(wave_pixels * slots_pp), depth, std::max<size_t>(n_waves, 1), (uint32_t)sub, rot ? 0 : lane_base);
        // a part with less free memory than the default sizing assumes: smaller waves instead of an error
        if (rc != SPT_ERR_CUDA || wave_pixels * slots_pp <= (1u << 16)) break;
        cudaDeviceSynchronize();
        for (auto &ln : s->lane) { ln.mem.release(); ln.wb.cap = 0; ln.wb.jcap = 0; }
        g_blocks.trim();
        wave_pixels = (wave_pixels + 1) / 2;
    }
    if (rc != SPT_OK) return rc;
    size_t per_wave = (size_t)(depth + 3) * SPT_ROW;
    cudaStream_t st = s->lane[lane_base].stream;
    // the frame's counter rows: this record's half of the array, zeroed wave by wave on the lane that runs the wave
    uint32_t *counts = s->counts + (size_t)slot * (s->counts_len / SPT_MAX_FRAMES);
    SampleSource src; src.smp = nullptr; src.stride = 0; src.rng = nullptr; src.n_rng = 0; src.seed = cfg.seed; src.spp = (uint32_t)rp->spp;
    FilmView fv; fv.d = film->desc; fv.pix = film->pix; fv.table = film->table;
    if (!fr.ev0 && (!(fr.ev0 = g_handles.event(s->device, true)) || !(fr.ev1 = g_handles.event(s->device, true)) ||
                    !(fr.evc = g_handles.event(s->device, false)))) return fail(SPT_ERR_CUDA, "event creation failed");
    const size_t hc_words = std::max<size_t>(n_waves, 1) * per_wave;
    if (fr.hc_cap < hc_words) {
        if (fr.hc) cudaFreeHost(fr.hc);
        fr.hc = nullptr; fr.hc_cap = 0;
        if (cudaMallocHost((void **)&fr.hc, hc_words * 4) != cudaSuccess) return fail(SPT_ERR_CUDA, "cudaMallocHost failed");
        fr.hc_cap = hc_words;
    }
    s->marks_on = fr.timed;
    if (fr.timed) reset_class_stats(s);
    CU(cudaEventRecord(fr.ev0, st));
    // fork - only for a frame that starts on an idle scene. A frame enqueued behind another one needs no common start: every lane
    // goes on with its next wave in stream order (its wave buffers are its own, the counter rows this record's), so one lane's
    // last kernels of frame k run beside the other lane's first kernels of frame k + 1 instead of beside an idle half machine.
    if (fr.timed) for (int k = 1; k < n_lanes; ++k) CU(cudaStreamWaitEvent(s->lane[lane_base + k].stream, fr.ev0, 0));
    std::vector<uint32_t> &hc_tree = fr.hc_tree;          // tree mode: the counter blocks of the pixel ranges that fitted, in the order they ran
    if (tree) {
        // The node pool of a wave is what its buffers hold beyond the camera samples: (SPT_TREE_SLOTS - 1) per sample on average.
        // A range whose trees need more (a window full of glass) is NOT added to the film; it is cut in two and each half runs in
        // the same buffers with twice the pool per sample - down to the full binary tree, which always fits.
        const WaveBuffers &wb = s->lane[0].wb;
        std::vector<std::pair<uint64_t, uint64_t>> todo;                     // {first pixel, pixels}, a stack
        for (size_t w = n_waves; w-- > 0;) todo.push_back({(uint64_t)w * wave_pixels, std::min<uint64_t>(wave_pixels, local_pixels - (uint64_t)w * wave_pixels)});
        std::vector<uint32_t> blk(per_wave);
        while (!todo.empty()) {
            const uint64_t base = todo.back().first, np = todo.back().second;
            todo.pop_back();
            CU(cudaMemsetAsync(counts, 0, per_wave * 4, st));
            cfg.pixel_base = base;
            cfg.n_samples = (uint32_t)(np * (uint64_t)rp->spp);
            run_wave(s, cfg, src, counts, 0);
            CU(cudaMemcpyAsync(blk.data(), counts, per_wave * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (blk[12]) {
                if (np == 1) return fail(SPT_ERR_UNSUPP, "specular tree: one pixel's samples spawn more nodes than a wave holds");
                todo.push_back({base + np / 2, np - np / 2});
                todo.push_back({base, np / 2});
                continue;
            }
            unsigned gw = (unsigned)std::min<uint64_t>((np * 32 + 255) / 256, (uint64_t)num_sms() * 16);
            spt_launch_film_add((int)gw, st, fv, s->dev.tables, wb.img_xy, wb.L, wb.cap, (uint32_t)(np * rp->spp), rp->spp);
            s->mark(SPT_K_FILM, 0);
            hc_tree.insert(hc_tree.end(), blk.begin(), blk.end());
        }
    } else
    for (size_t w = 0; w < n_waves; ++w) {
        const int li = lane_base + (int)(w % (size_t)n_lanes);
        cfg.pixel_base = (uint64_t)w * wave_pixels;
        uint64_t np = std::min<uint64_t>(wave_pixels, local_pixels - cfg.pixel_base);
        cfg.n_samples = (uint32_t)(np * (uint64_t)rp->spp);
        CU(cudaMemsetAsync(counts + w * per_wave, 0, per_wave * 4, s->lane[li].stream));
        run_wave(s, cfg, src, counts + w * per_wave, li);
        unsigned gw = (unsigned)std::min<uint64_t>((np * 32 + 255) / 256, (uint64_t)num_sms() * 16);
        const WaveBuffers &wb = s->lane[li].wb;
        spt_launch_film_add((int)gw, s->lane[li].stream, fv, s->dev.tables, wb.img_xy, wb.L, wb.cap, (uint32_t)(np * rp->spp), rp->spp);
        s->mark(SPT_K_FILM, li);
    }
    // join - on the book-keeping stream, not on a lane: no lane waits for another one, so a frame enqueued behind this one
    // starts on each lane as soon as that lane is done here
    for (int k = lane_base; k < lane_base + n_lanes; ++k) {
        CU(cudaEventRecord(s->evjoin[k], s->lane[k].stream));
        CU(cudaStreamWaitEvent(s->bk, s->evjoin[k], 0));
    }
    CU(cudaEventRecord(fr.ev1, s->bk));
    CU(cudaMemcpyAsync(fr.hc, counts, hc_words * 4, cudaMemcpyDeviceToHost, s->bk));
    CU(cudaEventRecord(fr.evc, s->bk));
    s->marks_on = true;
    fr.tree = tree; fr.n_waves = n_waves; fr.per_wave = per_wave; fr.depth = depth; fr.n_lanes = n_lanes; fr.spp = rp->spp;
    fr.nranks = nranks; fr.local_tiles = local_tiles; fr.cfg = cfg;
    ++s->f_count;
    return SPT_OK;
}

int spt_render_end(SptScene *s) {
    if (!s) return fail(SPT_ERR_ARG, "null scene");
    if (s->f_count == 0) return fail(SPT_ERR_ARG, "no frame in flight");
    DeviceGuard dg(s->device);
    SptScene::FrameRec &fr = s->frame[s->f_head];
    s->f_head = (s->f_head + 1) % SPT_MAX_FRAMES; --s->f_count;
    if (fr.empty) {
        if (fr.timed) reset_class_stats(s);
        s->stats.render_ms = 0.; s->stats.lanes_used = 0;
        return SPT_OK;
    }
    CU(cudaEventSynchronize(fr.evc));
    CU(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, fr.ev0, fr.ev1);
    s->stats.render_ms = ms;
    s->stats.lanes_used = fr.n_lanes;
    s->class_times_pending = fr.timed;      // ~80 cudaEventElapsedTime calls: only when spt_get_stats asks (not between frames)
    const RenderCfg &cfg = fr.cfg;
    size_t n_waves = fr.n_waves;
    const size_t per_wave = fr.per_wave;
    std::vector<uint32_t> hc;
    if (fr.tree) { hc.swap(fr.hc_tree); n_waves = hc.size() / per_wave; }
    else hc.assign(fr.hc, fr.hc + std::max<size_t>(n_waves, 1) * per_wave);
    // sample slots of tiles that overhang the sample extent carry rays that cannot hit anything: not samples
    uint64_t slots = 0, valid_pixels = 0;
    for (size_t w = 0; w < n_waves; ++w) slots += hc[w * per_wave];
    for (uint64_t k = 0; k < fr.local_tiles; ++k) {
        uint64_t t = k * (uint64_t)fr.nranks + (uint64_t)cfg.rank;
        int tx = (int)(t % cfg.tilesX), ty = (int)(t / cfg.tilesX);
        int w = std::min(cfg.tile, cfg.x1 - (cfg.x0 + tx * cfg.tile)), h = std::min(cfg.tile, cfg.y1 - (cfg.y0 + ty * cfg.tile));
        if (w > 0 && h > 0) valid_pixels += (uint64_t)w * h;
    }
    const uint64_t samples = valid_pixels * (uint64_t)fr.spp;
    s->stats.class_rays[SPT_K_GEN] = samples; s->stats.class_rays[SPT_K_FILM] = samples;
    s->stats.camera_samples += samples;
    if (n_waves) add_ray_stats(s, hc, fr.depth, n_waves);
    if (n_waves) {   // the first wave's bounce 0: what the first launch of each class worked on (sample slots incl. the few that overhang the extent)
        uint64_t *u = s->stats.first_launch_units;
        u[SPT_K_GEN] = hc[0]; u[SPT_K_FILM] = hc[0]; u[SPT_K_TRACE_PATH] = hc[0];
        u[SPT_K_SHADE] = hc[3]; u[SPT_K_ACCUMULATE] = hc[3]; u[SPT_K_ADVANCE] = hc[3];
    }
    const uint64_t overhang = slots > samples ? slots - samples : 0;
    s->stats.closest_rays -= overhang;
    s->stats.class_rays[SPT_K_TRACE_PATH] -= overhang;
    return SPT_OK;
}

int spt_render(SptScene *s, const SptCameraDesc *cam, SptFilm *film, const SptRenderParams *rp) {
    if (s && s->f_count) return fail(SPT_ERR_ARG, "a frame begun with spt_render_begin is in flight: spt_render_end first");
    int rc = render_begin(s, cam, film, rp, false);
    return rc == SPT_OK ? spt_render_end(s) : rc;
}
int spt_render_begin(SptScene *s, const SptCameraDesc *cam, SptFilm *film, const SptRenderParams *rp) { return render_begin(s, cam, film, rp, true); }


// ---- several GPUs ------------------------------------------------------------------------------------
int spt_film_ipc_export(SptFilm *f, uint8_t handle[SPT_IPC_HANDLE_BYTES]) {
    if (!f || !handle) return fail(SPT_ERR_ARG, "null argument");
    if (!f->owned) return fail(SPT_ERR_ARG, "only films allocated by the library can be exported");
    static_assert(sizeof(cudaIpcMemHandle_t) == SPT_IPC_HANDLE_BYTES, "IPC handle size");
    DeviceGuard dg(f->device);
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, f->pix));
    memcpy(handle, &h, sizeof(h));
    return SPT_OK;
}
SptFilm *spt_film_open_ipc(const SptFilmDesc *d, const uint8_t handle[SPT_IPC_HANDLE_BYTES]) {
    if (!d || !handle) { g_err = "null argument"; return nullptr; }
    if (spt_device_count() <= 0) { g_err = "no CUDA device: this library has no CPU path"; return nullptr; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    CUP(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    SptFilm *f = film_new(d, (float *)p);
    if (!f) { cudaIpcCloseMemHandle(p); return nullptr; }
    f->ipc_base = p;
    return f;
}

}  // extern "C"

struct SptMulti {
    std::vector<int> devices;
    std::vector<SptScene *> scenes;
    std::vector<SptFilm *> films;        // films[0] owns the pixels (devices[0]); the others alias them through peer access
    std::vector<int> rc;
    std::vector<std::string> err;
};

// runs fn(k) on one host thread per device, with that device current; gathers status + message per device
template <typename F> static int multi_run(SptMulti *m, F fn) {
    const size_t n = m->devices.size();
    std::vector<std::thread> th;
    for (size_t k = 0; k < n; ++k)
        th.emplace_back([m, k, &fn]() {
            m->rc[k] = SPT_OK; m->err[k].clear();
            if (cudaSetDevice(m->devices[k]) != cudaSuccess) { m->rc[k] = SPT_ERR_CUDA; m->err[k] = "cudaSetDevice failed"; cudaGetLastError(); return; }
            m->rc[k] = fn(k);
            if (m->rc[k] != SPT_OK) m->err[k] = g_err;
        });
    for (auto &t : th) t.join();
    for (size_t k = 0; k < n; ++k)
        if (m->rc[k] != SPT_OK) { g_err = "device " + std::to_string(m->devices[k]) + ": " + m->err[k]; return m->rc[k]; }
    return SPT_OK;
}

extern "C" {

void spt_multi_destroy(SptMulti *m) {
    if (!m) return;
    multi_run(m, [m](size_t k) {
        if (k > 0 && k < m->films.size() && m->films[k]) spt_film_destroy(m->films[k]);
        if (k < m->scenes.size() && m->scenes[k]) spt_scene_destroy(m->scenes[k]);
        return SPT_OK;
    });
    if (!m->films.empty() && m->films[0]) spt_film_destroy(m->films[0]);
    delete m;
}

SptMulti *spt_multi_create(const SptSceneDesc *scene, const SptFilmDesc *film, int n_devices, const int *devices) {
    if (!scene || !film) { g_err = "null argument"; return nullptr; }
    const int visible = spt_device_count();
    if (visible <= 0) { g_err = "no CUDA device: this library has no CPU path"; return nullptr; }
    if (n_devices <= 0) { n_devices = visible; devices = nullptr; }
    SptMulti *m = new SptMulti();
    for (int k = 0; k < n_devices; ++k) {
        const int d = devices ? devices[k] : k;
        if (d < 0 || d >= visible) { g_err = "device ordinal out of range"; delete m; return nullptr; }
        for (int dd : m->devices) if (dd == d) { g_err = "a device is listed twice"; delete m; return nullptr; }
        m->devices.push_back(d);
    }
    const size_t n = m->devices.size();
    m->scenes.assign(n, nullptr); m->films.assign(n, nullptr); m->rc.assign(n, SPT_OK); m->err.assign(n, "");
    int prev = 0; cudaGetDevice(&prev);
    // the film: on the first device; every other device reaches it through peer access (NVLink / NVSwitch)
    for (size_t k = 1; k < n; ++k) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, m->devices[k], m->devices[0]) != cudaSuccess || !can) {
            g_err = "device " + std::to_string(m->devices[k]) + " cannot access the film's device " + std::to_string(m->devices[0]) + " (no peer access)";
            cudaGetLastError(); delete m; return nullptr;
        }
    }
    cudaSetDevice(m->devices[0]);
    m->films[0] = spt_film_create(film);
    cudaSetDevice(prev);
    if (!m->films[0]) { delete m; return nullptr; }
    float *pix0 = m->films[0]->pix;
    int rc = multi_run(m, [m, scene, film, pix0](size_t k) {
        if (k > 0) {
            cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[0], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { g_err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); cudaGetLastError(); return SPT_ERR_CUDA; }
            cudaGetLastError();
            m->films[k] = spt_film_create_external(film, pix0);
            if (!m->films[k]) return SPT_ERR_CUDA;
        }
        m->scenes[k] = spt_scene_create(scene);
        return m->scenes[k] ? SPT_OK : SPT_ERR_CUDA;
    });
    cudaSetDevice(prev);
    if (rc != SPT_OK) { std::string keep = g_err; spt_multi_destroy(m); g_err = keep; return nullptr; }
    return m;
}

int spt_multi_device_count(SptMulti *m) { return m ? (int)m->devices.size() : 0; }
SptFilm *spt_multi_film(SptMulti *m) { return m ? m->films[0] : nullptr; }

int spt_multi_render(SptMulti *m, const SptCameraDesc *cam, const SptRenderParams *rp) {
    if (!m || !cam || !rp) return fail(SPT_ERR_ARG, "null argument");
    const int n = (int)m->devices.size();
    // every device renders its tile set into the one film and returns when its streams have drained: the join below is
    // the end-of-frame barrier
    return multi_run(m, [m, cam, rp, n](size_t k) {
        SptRenderParams p = *rp;
        p.tile_rank = (int)k; p.tile_nranks = n;
        if (p.tile_size <= 0) p.tile_size = 32;
        return spt_render(m->scenes[k], cam, m->films[k], &p);
    });
}

int spt_multi_get_stats(SptMulti *m, int index, SptStats *out) {
    if (!m || index < 0 || index >= (int)m->devices.size()) return fail(SPT_ERR_ARG, "device index out of range");
    return spt_get_stats(m->scenes[index], out);
}
double spt_multi_last_render_ms(SptMulti *m) {
    double t = 0.;
    if (m) for (SptScene *s : m->scenes) if (s) t = std::max(t, s->stats.render_ms);
    return t;
}

}  // extern "C"
