// agofrt_cabi.cu -- host side of libagofrt.so: the C ABI declared in include/agofrt.h.
//
// No CPU compute path lives here: the only arithmetic done on the host is O(nbin) / O(frames)
// planning (the exact bin-threshold table, the per-job single-pass proof from coordinate bounds,
// the atom permutation).  Every pair evaluation happens in agofrt_kernels.cu.
#include "agofrt.h"
#include "agofrt_kernels.cuh"

#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <limits>
#include <memory>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

using namespace agofrt;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// Nothing throws across the C ABI: every entry point is a function-try-block ending here (std::vector growth in
// the planning code is the realistic source: a block of billions of jobs, a trajectory of billions of atoms).
static int on_exception() noexcept {
    try {
        throw;
    } catch (const std::bad_alloc &) {
        return fail(AGOFRT_ERR_TOO_LARGE, "out of host memory while planning the call");
    } catch (const std::exception &e) {
        return fail(AGOFRT_ERR_INTERNAL, "unexpected exception: %s", e.what());
    } catch (...) {
        return fail(AGOFRT_ERR_INTERNAL, "unexpected exception");
    }
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(AGOFRT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                       \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen) so that whichever libnccl.so.2 the process already carries --
// the system one, or the one a python launcher has loaded -- is the one used.
// ---------------------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, void *) = nullptr;   // NCCL >= 2.18
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi &nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    // NCCL writes its version banner / debug lines to stdout unless told otherwise; stdout belongs to the
    // caller's data (the CLI prints g(r,t) there), so default the debug stream to stderr.
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return api;
#define LOAD(field, sym)                                                             \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym));       \
    if (!api.field) return api;
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(GetErrorString, "ncclGetErrorString")
    LOAD(Broadcast, "ncclBroadcast")
    LOAD(AllGather, "ncclAllGather")
#undef LOAD
    api.CommSplit = reinterpret_cast<decltype(api.CommSplit)>(dlsym(api.handle, "ncclCommSplit"));   // optional
    api.ok = true;
    return api;
}

#define NC(call)                                                                                            \
    do {                                                                                                    \
        ncclResult_t r__ = (call);                                                                          \
        if (r__ != ncclSuccess)                                                                             \
            return fail(AGOFRT_ERR_NCCL, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r__),     \
                        __FILE__, __LINE__);                                                                \
    } while (0)

// ---------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------
struct Dev {
    int id = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    size_t smem_per_sm = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    ncclComm_t comm = nullptr;
    ncclComm_t comm_up = nullptr;   // window exchange (agofrt_traj_upload_ex, AGOFRT_UP_SHARED): a communicator of its
                                    // own, because a window may be uploaded by a second host thread while the first
                                    // is inside agofrt_block (whose all-reduce runs on `comm`)
    double *peak_sink = nullptr;
    // page-locked staging for uploads from pageable host memory (two slots, filled by several host threads while the
    // previous slot is on its way to the device)
    double *hstage[2] = {nullptr, nullptr};
    size_t hstage_bytes = 0;
    cudaEvent_t hstage_free[2] = {nullptr, nullptr};
    std::mutex *hstage_mutex = nullptr;
    // The two large buffers of a window (SoA positions, AoS staging) are kept when a window is destroyed and handed to
    // the next one that fits: a caller that builds a trajectory object per analysis (python: Trajectory(...) per call)
    // would otherwise pay cudaMalloc + cudaFree of gigabytes every time (measured: 0.2-0.7 s per 2.3 GB window).
    struct Spare {
        void *ptr = nullptr;
        size_t bytes = 0;
    };
    Spare *spare = nullptr;   // [kSpares], guarded by hstage_mutex
    // ... and the small ones of a plan (thresholds, histogram, counters): a cudaFree right after a long kernel was
    // measured at up to 1.2 s (r2x), so objects that are created and destroyed per analysis recycle their buffers
    std::vector<Spare> *pool = nullptr;   // guarded by hstage_mutex
};
constexpr int kSpares = 4;
constexpr size_t kPoolEntries = 64, kPoolMaxBytes = 32u << 20;

// device memory for the short-lived small buffers: recycled by size class (multiples of 256 bytes)
static cudaError_t pool_malloc(Dev &dv, void **ptr, size_t bytes) {
    const size_t want = (std::max<size_t>(bytes, 1) + 255) / 256 * 256;
    {
        std::lock_guard<std::mutex> lock(*dv.hstage_mutex);
        for (size_t k = 0; k < dv.pool->size(); ++k)
            if ((*dv.pool)[k].bytes == want) {
                *ptr = (*dv.pool)[k].ptr;
                dv.pool->erase(dv.pool->begin() + k);
                return cudaSuccess;
            }
    }
    return cudaMalloc(ptr, want);
}
template <class T>
static cudaError_t pool_malloc(Dev &dv, T **ptr, size_t bytes) {
    return pool_malloc(dv, reinterpret_cast<void **>(ptr), bytes);
}
static void pool_free(Dev &dv, void *ptr, size_t bytes) {
    if (!ptr) return;
    const size_t have = (std::max<size_t>(bytes, 1) + 255) / 256 * 256;
    if (have <= kPoolMaxBytes) {
        std::lock_guard<std::mutex> lock(*dv.hstage_mutex);
        if (dv.pool->size() < kPoolEntries) {
            dv.pool->push_back(Dev::Spare{ptr, have});
            return;
        }
    }
    cudaFree(ptr);
}

static void *spare_take(Dev &dv, size_t bytes) {
    std::lock_guard<std::mutex> lock(*dv.hstage_mutex);
    int best = -1;
    for (int k = 0; k < kSpares; ++k)
        if (dv.spare[k].ptr && dv.spare[k].bytes >= bytes && dv.spare[k].bytes <= bytes + bytes / 4 + (1u << 20) &&
            (best < 0 || dv.spare[k].bytes < dv.spare[best].bytes))
            best = k;
    if (best < 0) return nullptr;
    void *p = dv.spare[best].ptr;
    dv.spare[best] = Dev::Spare();
    return p;
}

// keeps the buffer if it is large enough to matter and a slot is free (or holds a smaller one); else frees it
static void spare_give(Dev &dv, void *ptr, size_t bytes) {
    if (!ptr) return;
    if (bytes >= (4u << 20)) {
        std::lock_guard<std::mutex> lock(*dv.hstage_mutex);
        int slot = -1;
        for (int k = 0; k < kSpares && slot < 0; ++k)
            if (!dv.spare[k].ptr) slot = k;
        if (slot < 0) {
            int smallest = 0;
            for (int k = 1; k < kSpares; ++k)
                if (dv.spare[k].bytes < dv.spare[smallest].bytes) smallest = k;
            if (dv.spare[smallest].bytes < bytes) {
                cudaFree(dv.spare[smallest].ptr);
                slot = smallest;
            }
        }
        if (slot >= 0) {
            dv.spare[slot].ptr = ptr;
            dv.spare[slot].bytes = bytes;
            return;
        }
    }
    cudaFree(ptr);
}

struct agofrt_ctx {
    std::vector<Dev> devs;
    int first_rank = 0;
    int world = 0;       // 0: not sharded beyond the local devices
    bool comm_ready = false;
    bool shard_only = false;
    // page-locked read-back buffers of destroyed plans, recycled by size class like Dev::pool
    std::vector<Dev::Spare> host_pool;
    std::mutex host_pool_mutex;
};

static cudaError_t host_pool_alloc(agofrt_ctx *ctx, void **ptr, size_t bytes) {
    const size_t want = (std::max<size_t>(bytes, 1) + 255) / 256 * 256;
    {
        std::lock_guard<std::mutex> lock(ctx->host_pool_mutex);
        for (size_t k = 0; k < ctx->host_pool.size(); ++k)
            if (ctx->host_pool[k].bytes == want) {
                *ptr = ctx->host_pool[k].ptr;
                ctx->host_pool.erase(ctx->host_pool.begin() + k);
                return cudaSuccess;
            }
    }
    return cudaHostAlloc(ptr, want, cudaHostAllocPortable);
}
static void host_pool_free(agofrt_ctx *ctx, void *ptr, size_t bytes) {
    if (!ptr) return;
    const size_t have = (std::max<size_t>(bytes, 1) + 255) / 256 * 256;
    if (have <= kPoolMaxBytes) {
        std::lock_guard<std::mutex> lock(ctx->host_pool_mutex);
        if (ctx->host_pool.size() < kPoolEntries) {
            ctx->host_pool.push_back(Dev::Spare{ptr, have});
            return;
        }
    }
    cudaFreeHost(ptr);
}

struct TrajDev {
    double *pos = nullptr;       // [max_frames][3][npad]
    double *box6 = nullptr;      // [max_frames][6]
    double *bounds = nullptr;    // [max_frames][6]
    double *stage = nullptr;     // AoS staging
    double *box_stage = nullptr; // [max_frames][stride]
    int *perm = nullptr;         // [npad]
    int *type_pad = nullptr;     // [npad]
    int *type_start = nullptr;   // [ntypes+1]
    unsigned int *flags = nullptr;  // [8]: 0 inf seen, 1 wrap cap hit (block), 2 NaN in a real atom, 3 wrap cap hit (upload),
                                    //      4 records with an unknown atom id, 5 records whose type changed (agofrt_traj_upload_records)
    double *probe = nullptr;        // [4]: result of agofrt_traj_d2_pair
    double *rot = nullptr;          // [max_frames][9] rotation matrix Q of every frame (agofrt_traj_set_rotation), or NULL
    unsigned long long *nb_hist = nullptr;  // neighbour-count histogram [ntypes][natoms+1] (agofrt_neighbour_hist)
    unsigned int *nb_counts = nullptr;      // per-atom neighbour counts [frames of the call][ntypes][npad]
    size_t nb_counts_len = 0;
    int *nb_frames = nullptr;
    size_t nb_frames_cap = 0;
    int *id_table = nullptr;        // LAMMPS atom id -> slot of the caller's atom order (agofrt_traj_set_ids), -1 = unknown id
    int *slot_type = nullptr;       // [natoms] raw LAMMPS type every slot is expected to carry
    size_t pos_bytes = 0, stage_bytes = 0;   // sizes of pos / stage as allocated (they may come from the context's spares)
    double *raw = nullptr;          // staging of raw dump records, 8 doubles per atom (agofrt_traj_upload_records)
    size_t raw_frames = 0;
    cudaStream_t up = nullptr;      // uploads of THIS window run here: they can overlap the pair kernels of another
                                    // window of the same context (which run on the device's main stream)
};

struct agofrt_traj {
    agofrt_ctx *ctx = nullptr;
    size_t natoms = 0, max_frames = 0;
    int npad = 0, ntypes = 0, stride = 6;
    std::vector<int> type_id, type_start, type_pad, perm;
    std::vector<TrajDev> dev;
    size_t stage_frames = 0;
    // current window
    size_t first_frame = 0, nframes = 0;
    std::vector<double> box6;    // host copy [nframes][6]
    std::vector<double> bounds;  // host copy [nframes][6]
    bool has_inf = false;
    bool has_nan = false;   // NaN coordinates in the input (they are never in range, as in the reference)
    bool bad_box = false;
    size_t rot_first = 0, rot_frames = 0;   // frames whose rotation matrices are on the devices
    std::vector<double> cm;      // per-type centres of mass of the window frames [nframes][ntypes][3] (agofrt_traj_set_cm)
    size_t cm_first = 0, cm_frames = 0;
    bool perm_valid = false;     // the permutation is kept over uploads and refreshed every kPermRefresh frames
    size_t perm_frame = 0;       // first frame of the window it was built from
    size_t id_table_len = 0;     // entries of the id -> slot table on the devices (0: agofrt_traj_set_ids not called)
    std::vector<int> id_table_host;
    std::vector<int> slot_of;    // atom -> device slot (built on demand by agofrt_traj_d2_pair)
    size_t slot_of_first = 0;
    bool slot_of_stale = true;
};

struct PlanDev {
    double *thr = nullptr, *thr_full = nullptr;
    unsigned long long *ghist = nullptr;
    size_t ghist_len = 0;
    Job *jobs = nullptr;
    size_t jobs_cap = 0;
    unsigned int *counter = nullptr;      // [1]
    unsigned long long *edges = nullptr;  // [1]
    unsigned long long *batch = nullptr;  // agofrt_blocks: [nblocks][len] counts of whole blocks
    size_t batch_len = 0;
};

struct agofrt_plan {
    agofrt_ctx *ctx = nullptr;   // outlives windows: the plan may be retargeted or destroyed after its first window is gone
    agofrt_traj *traj = nullptr;
    double rmin = 0, rmax = 0, dr = 0, rmin2 = 0, rmax2 = 0;
    unsigned nbin = 0;
    std::vector<double> thr_full, thr;
    bool empty_range = false;
    unsigned hlo = 0, hspan = 0;
    float inv_dr = 0, c0 = 0;
    float c0h = 0, lim = 0;      // safe-zone binning: c0 - 0.5, 0.5 - eps
    int ntypes = 0;
    float qmax = 0;              // ... clamp of the bin coordinate: nbin + 0.25
    int glo = 0;                 // ... guard bins below bin 0 in every shared-memory histogram row
    bool safe_ok = false;        // validated on the device when the plan was made
    double q_reach = 0;          // largest bin coordinate the float path may meet before it must give up
    bool safe2_ok = false;       // the two-floor form (MODE_SAFE2) is valid for this plan: integer c0, validated on the device
    float inv_lo = 0, inv_hi = 0, bias0 = 0, smax = 0;
    std::vector<PlanDev> dev;
    unsigned long long *host_counts = nullptr;  // pinned
    size_t host_counts_len = 0;
    unsigned long long *host_edges = nullptr;   // pinned [ndev]
    unsigned int *host_flags = nullptr;         // pinned [ndev]
    size_t last_len = 0;                        // words of the counts the last agofrt_block left in dev[0].ghist
    bool last_valid = false;                    // ... and whether they are the complete (all-reduced) counts
    unsigned batch_blocks = 0;                  // agofrt_blocks: blocks held in dev[*].batch, each batch_block_len words
    size_t batch_block_len = 0;
};

// ---------------------------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------------------------
extern "C" const char *agofrt_version(void) { return "agofrt 0.1 (sm_100a)"; }
extern "C" const char *agofrt_last_error(void) { return g_last_error.c_str(); }

extern "C" int agofrt_device_count(int *count) try {
    if (!count) return fail(AGOFRT_ERR_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(AGOFRT_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// Page-locked when a CUDA driver is present.  Without one (authoring container, CPU-only tests of the
// file reader) the buffer is plain aligned memory: allocation is not computation, and every entry point
// that computes still fails with AGOFRT_ERR_CUDA.
static std::mutex g_pageable_mutex;
static std::unordered_set<void *> g_pageable;
// Page-locking costs milliseconds per call (cudaHostAlloc of the two 29 MB window buffers of the C1 run: 6-31 ms,
// r2ab): buffers of 1 MiB .. 256 MiB that are given back are kept (at most 8, 512 MiB in all) and handed to the next
// request of the same size -- a process that builds a trajectory object per analysis pays for them once.
static std::unordered_map<void *, size_t> g_pinned;                  // live page-locked buffers and their sizes
static std::vector<std::pair<void *, size_t>> g_pinned_pool;
static size_t g_pinned_pool_bytes = 0;
constexpr size_t kPinnedPoolMin = 1u << 20, kPinnedPoolMax = 256u << 20, kPinnedPoolTotal = 512u << 20, kPinnedPoolEntries = 8;

extern "C" int agofrt_host_alloc(void **ptr, size_t bytes) try {
    if (!ptr) return fail(AGOFRT_ERR_ARG, "ptr is NULL");
    *ptr = nullptr;
    const size_t want = (std::max<size_t>(bytes, 1) + 4095) / 4096 * 4096;
    {
        std::lock_guard<std::mutex> lock(g_pageable_mutex);
        for (size_t k = 0; k < g_pinned_pool.size(); ++k)
            if (g_pinned_pool[k].second == want) {
                *ptr = g_pinned_pool[k].first;
                g_pinned_pool_bytes -= want;
                g_pinned_pool.erase(g_pinned_pool.begin() + k);
                g_pinned[*ptr] = want;
                return AGOFRT_OK;
            }
    }
    const cudaError_t e = cudaHostAlloc(ptr, want, cudaHostAllocPortable);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lock(g_pageable_mutex);
        g_pinned[*ptr] = want;
        return AGOFRT_OK;
    }
    cudaGetLastError();
    if (e == cudaErrorInsufficientDriver || e == cudaErrorNoDevice) {
        void *p = nullptr;
        if (posix_memalign(&p, 4096, want) != 0) return fail(AGOFRT_ERR_INTERNAL, "out of host memory (%zu bytes)", bytes);
        std::lock_guard<std::mutex> lock(g_pageable_mutex);
        g_pageable.insert(p);
        *ptr = p;
        return AGOFRT_OK;
    }
    return fail(AGOFRT_ERR_CUDA, "cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
} catch (...) {
    return on_exception();
}
extern "C" int agofrt_host_free(void *ptr) try {
    if (!ptr) return AGOFRT_OK;
    {
        std::lock_guard<std::mutex> lock(g_pageable_mutex);
        auto it = g_pageable.find(ptr);
        if (it != g_pageable.end()) {
            g_pageable.erase(it);
            free(ptr);
            return AGOFRT_OK;
        }
        auto pin = g_pinned.find(ptr);
        if (pin != g_pinned.end()) {
            const size_t have = pin->second;
            g_pinned.erase(pin);
            if (have >= kPinnedPoolMin && have <= kPinnedPoolMax && g_pinned_pool.size() < kPinnedPoolEntries &&
                g_pinned_pool_bytes + have <= kPinnedPoolTotal) {
                g_pinned_pool.emplace_back(ptr, have);
                g_pinned_pool_bytes += have;
                return AGOFRT_OK;
            }
        }
    }
    CU(cudaFreeHost(ptr));
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" int agofrt_ctx_create(agofrt_ctx **out, const int *devices, int ndev) try {
    if (!out) return fail(AGOFRT_ERR_ARG, "ctx is NULL");
    *out = nullptr;
    int avail = 0;
    cudaError_t e = cudaGetDeviceCount(&avail);
    if (e != cudaSuccess || avail <= 0)
        return fail(AGOFRT_ERR_CUDA, "no usable CUDA device (%s); libagofrt has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
    std::vector<int> ids;
    if (ndev == -1) {
        for (int i = 0; i < avail; ++i) ids.push_back(i);
    } else if (ndev == 0 || devices == nullptr) {
        ids.push_back(0);
    } else {
        for (int i = 0; i < ndev; ++i) {
            if (devices[i] < 0 || devices[i] >= avail)
                return fail(AGOFRT_ERR_ARG, "device %d not in [0,%d)", devices[i], avail);
            ids.push_back(devices[i]);
        }
    }
    auto ctx = std::make_unique<agofrt_ctx>();
    for (int id : ids) {
        Dev d;
        d.id = id;
        CU(cudaSetDevice(id));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, id));
        if (prop.major < 10)
            return fail(AGOFRT_ERR_CUDA, "device %d is sm_%d%d; libagofrt is built for sm_100a only", id, prop.major,
                        prop.minor);
        d.sm_count = prop.multiProcessorCount;
        d.smem_optin = prop.sharedMemPerBlockOptin;
        d.smem_per_sm = prop.sharedMemPerMultiprocessor;
        CU(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&d.ev_begin));
        CU(cudaEventCreate(&d.ev_end));
        CU(cudaEventCreate(&d.ev_k0));
        CU(cudaEventCreate(&d.ev_k1));
        CU(prepare_pair_kernels(d.smem_optin));
        CU(cudaEventCreateWithFlags(&d.hstage_free[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&d.hstage_free[1], cudaEventDisableTiming));
        d.hstage_mutex = new std::mutex();
        d.spare = new Dev::Spare[kSpares];
        d.pool = new std::vector<Dev::Spare>();
        ctx->devs.push_back(d);
    }
    *out = ctx.release();
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_ctx_destroy(agofrt_ctx *ctx) try {
    if (!ctx) return AGOFRT_OK;
    for (Dev &d : ctx->devs) {
        cudaSetDevice(d.id);
        if (d.comm_up && nccl_api().ok) nccl_api().CommDestroy(d.comm_up);
        if (d.comm && nccl_api().ok) nccl_api().CommDestroy(d.comm);
        if (d.peak_sink) cudaFree(d.peak_sink);
        for (int k = 0; k < 2; ++k) {
            if (d.hstage[k]) cudaFreeHost(d.hstage[k]);
            if (d.hstage_free[k]) cudaEventDestroy(d.hstage_free[k]);
        }
        if (d.spare)
            for (int k = 0; k < kSpares; ++k) cudaFree(d.spare[k].ptr);
        delete[] d.spare;
        if (d.pool)
            for (const Dev::Spare &sp : *d.pool) cudaFree(sp.ptr);
        delete d.pool;
        delete d.hstage_mutex;
        if (d.stream) cudaStreamDestroy(d.stream);
        if (d.ev_begin) cudaEventDestroy(d.ev_begin);
        if (d.ev_end) cudaEventDestroy(d.ev_end);
        if (d.ev_k0) cudaEventDestroy(d.ev_k0);
        if (d.ev_k1) cudaEventDestroy(d.ev_k1);
    }
    for (const Dev::Spare &sp : ctx->host_pool) cudaFreeHost(sp.ptr);
    delete ctx;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_ctx_ndev(const agofrt_ctx *ctx) { return ctx ? static_cast<int>(ctx->devs.size()) : 0; }

extern "C" int agofrt_comm_unique_id(char id[AGOFRT_COMM_ID_BYTES]) try {
    static_assert(AGOFRT_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    if (!id) return fail(AGOFRT_ERR_ARG, "id is NULL");
    NcclApi &api = nccl_api();
    if (!api.ok) return fail(AGOFRT_ERR_NCCL, "libnccl.so.2 could not be loaded");
    ncclUniqueId u;
    NC(api.GetUniqueId(&u));
    memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

static int comm_init(agofrt_ctx *ctx, const ncclUniqueId &u, int first_rank, int world) {
    NcclApi &api = nccl_api();
    if (!api.ok) return fail(AGOFRT_ERR_NCCL, "libnccl.so.2 could not be loaded");
    const int nloc = static_cast<int>(ctx->devs.size());
    if (first_rank < 0 || first_rank + nloc > world) return fail(AGOFRT_ERR_ARG, "rank range outside world");
    // NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION/WARN (NCCL_DEBUG_FILE is only honoured
    // above that level).  stdout is the caller's data channel (the CLI prints g(r,t) there): point fd 1 at
    // stderr while the communicator is created.
    fflush(stdout);
    const int saved_stdout = dup(1);
    if (saved_stdout >= 0) dup2(2, 1);
    auto restore = [&]() {
        if (saved_stdout >= 0) {
            fflush(stdout);
            dup2(saved_stdout, 1);
            close(saved_stdout);
        }
    };
    ncclResult_t nr = api.GroupStart();
    for (int i = 0; i < nloc && nr == ncclSuccess; ++i) {
        if (cudaSetDevice(ctx->devs[i].id) != cudaSuccess) {
            nr = ncclUnhandledCudaError;
            break;
        }
        nr = api.CommInitRank(&ctx->devs[i].comm, world, u, first_rank + i);
    }
    const ncclResult_t ne = api.GroupEnd();
    restore();
    if (nr == ncclSuccess) nr = ne;
    if (nr != ncclSuccess) return fail(AGOFRT_ERR_NCCL, "NCCL communicator creation failed: %s", api.GetErrorString(nr));
    // ... and a second one over the same ranks for the window exchange (optional: without ncclCommSplit the shared
    // upload falls back to one full host-to-device copy per device)
    if (api.CommSplit) {
        ncclResult_t sr = api.GroupStart();
        for (int i = 0; i < nloc && sr == ncclSuccess; ++i) {
            if (cudaSetDevice(ctx->devs[i].id) != cudaSuccess) {
                sr = ncclUnhandledCudaError;
                break;
            }
            sr = api.CommSplit(ctx->devs[i].comm, 0, first_rank + i, &ctx->devs[i].comm_up, nullptr);
        }
        const ncclResult_t se = api.GroupEnd();
        if (sr == ncclSuccess) sr = se;
        if (sr != ncclSuccess)
            for (int i = 0; i < nloc; ++i) ctx->devs[i].comm_up = nullptr;
    }
    ctx->first_rank = first_rank;
    ctx->world = world;
    ctx->comm_ready = true;
    ctx->shard_only = false;
    return AGOFRT_OK;
}

extern "C" int agofrt_comm_join(agofrt_ctx *ctx, const char id[AGOFRT_COMM_ID_BYTES], int first_rank, int world) try {
    if (!ctx || !id) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (ctx->comm_ready) return fail(AGOFRT_ERR_ARG, "context already has a communicator");
    ncclUniqueId u;
    memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
    return comm_init(ctx, u, first_rank, world);
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_ctx_set_shard(agofrt_ctx *ctx, int first_rank, int world) try {
    if (!ctx) return fail(AGOFRT_ERR_ARG, "NULL argument");
    const int nloc = static_cast<int>(ctx->devs.size());
    if (world < nloc || first_rank < 0 || first_rank + nloc > world)
        return fail(AGOFRT_ERR_ARG, "rank range outside world");
    if (ctx->comm_ready) return fail(AGOFRT_ERR_ARG, "context already has a communicator");
    ctx->first_rank = first_rank;
    ctx->world = world;
    ctx->shard_only = true;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_shard_range(uint64_t units, int rank, int world, uint64_t *begin, uint64_t *end) try {
    if (!begin || !end || world <= 0 || rank < 0 || rank >= world) return fail(AGOFRT_ERR_ARG, "bad shard arguments");
    // 128-bit products: units * world never overflows
    *begin = static_cast<uint64_t>(static_cast<unsigned __int128>(units) * rank / world);
    *end = static_cast<uint64_t>(static_cast<unsigned __int128>(units) * (rank + 1) / world);
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_block_share(unsigned nblocks, int rank, int world, unsigned block, unsigned *part_a, unsigned *part_b) try {
    if (!part_a || !part_b) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (world <= 0 || rank < 0 || rank >= world || block >= nblocks) return fail(AGOFRT_ERR_ARG, "bad rank / world / block");
    // in units of 1/world block: the rank owns [rank * nblocks, (rank + 1) * nblocks), the block is [block * world, (block + 1) * world)
    const uint64_t W = static_cast<uint64_t>(world), own_a = static_cast<uint64_t>(rank) * nblocks,
                   own_b = (static_cast<uint64_t>(rank) + 1) * nblocks, blk_a = static_cast<uint64_t>(block) * W, blk_b = blk_a + W;
    const uint64_t a = std::max(own_a, blk_a), b = std::min(own_b, blk_b);
    *part_a = b > a ? static_cast<unsigned>(a - blk_a) : 0u;
    *part_b = b > a ? static_cast<unsigned>(b - blk_a) : 0u;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// a multi-device context without a joined communicator is its own communicator
static int ensure_local_comm(agofrt_ctx *ctx) {
    if (ctx->comm_ready || ctx->shard_only || ctx->devs.size() <= 1) return AGOFRT_OK;
    NcclApi &api = nccl_api();
    if (!api.ok) return fail(AGOFRT_ERR_NCCL, "libnccl.so.2 could not be loaded (needed for >1 device)");
    ncclUniqueId u;
    NC(api.GetUniqueId(&u));
    return comm_init(ctx, u, 0, static_cast<int>(ctx->devs.size()));
}

// ---------------------------------------------------------------------------------------------
// trajectory window
// ---------------------------------------------------------------------------------------------
static void free_traj_dev(agofrt_traj *t) {
    for (size_t i = 0; i < t->dev.size(); ++i) {
        cudaSetDevice(t->ctx->devs[i].id);
        TrajDev &d = t->dev[i];
        Dev &dv = t->ctx->devs[i];
        const size_t npad1 = std::max(t->npad, 1);
        spare_give(dv, d.pos, d.pos_bytes);
        pool_free(dv, d.box6, t->max_frames * 6 * sizeof(double));
        pool_free(dv, d.bounds, t->max_frames * 6 * sizeof(double));
        spare_give(dv, d.stage, d.stage_bytes);
        pool_free(dv, d.box_stage, t->max_frames * t->stride * sizeof(double));
        pool_free(dv, d.perm, npad1 * sizeof(int));
        pool_free(dv, d.type_pad, npad1 * sizeof(int));
        pool_free(dv, d.type_start, (t->ntypes + 1) * sizeof(int));
        pool_free(dv, d.flags, 8 * sizeof(unsigned int));
        pool_free(dv, d.probe, 4 * sizeof(double));
        cudaFree(d.rot);
        cudaFree(d.nb_hist);
        cudaFree(d.nb_frames);
        cudaFree(d.nb_counts);
        cudaFree(d.id_table);
        cudaFree(d.slot_type);
        cudaFree(d.raw);
        if (d.up) cudaStreamDestroy(d.up);
    }
}

extern "C" int agofrt_traj_create(agofrt_traj **out, agofrt_ctx *ctx, size_t natoms, int box_stride,
                                  const int *type_id, int ntypes, size_t max_frames) try {
    if (!out || !ctx) return fail(AGOFRT_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (box_stride != 6 && box_stride != 9) return fail(AGOFRT_ERR_ARG, "box_stride must be 6 or 9");
    if (ntypes <= 0 || ntypes > 64) return fail(AGOFRT_ERR_ARG, "ntypes must be in [1,64]");
    if (natoms > 0 && !type_id) return fail(AGOFRT_ERR_ARG, "type_id is NULL");
    if (natoms > (1u << 30)) return fail(AGOFRT_ERR_ARG, "natoms too large");
    if (max_frames == 0) return fail(AGOFRT_ERR_ARG, "max_frames must be > 0");
    auto t = std::make_unique<agofrt_traj>();
    t->ctx = ctx;
    t->natoms = natoms;
    t->ntypes = ntypes;
    t->stride = box_stride;
    t->max_frames = max_frames;
    t->type_id.assign(type_id, type_id + natoms);
    std::vector<size_t> cnt(ntypes, 0);
    for (size_t i = 0; i < natoms; ++i) {
        if (type_id[i] < 0 || type_id[i] >= ntypes)
            return fail(AGOFRT_ERR_ARG, "type_id[%zu]=%d not in [0,%d)", i, type_id[i], ntypes);
        cnt[type_id[i]]++;
    }
    t->type_start.resize(ntypes + 1);
    size_t off = 0;
    for (int k = 0; k < ntypes; ++k) {
        t->type_start[k] = static_cast<int>(off);
        off += (cnt[k] + kPadGroup - 1) / kPadGroup * kPadGroup;
    }
    t->type_start[ntypes] = static_cast<int>(off);
    t->npad = static_cast<int>(off);

    t->type_pad.assign(t->npad, 0);
    for (int k = 0; k < ntypes; ++k)
        for (int s = t->type_start[k]; s < t->type_start[k + 1]; ++s) t->type_pad[s] = k;
    t->perm.assign(t->npad, -1);

    // staging: at most ~256 MiB of AoS frames at a time
    const size_t frame_bytes = std::max<size_t>(natoms * 3 * sizeof(double), 1);
    t->stage_frames = std::max<size_t>(1, std::min<size_t>(max_frames, (256u << 20) / frame_bytes));

    const size_t npad1 = std::max(t->npad, 1);
    t->dev.resize(ctx->devs.size());
    for (size_t i = 0; i < ctx->devs.size(); ++i) {
        CU(cudaSetDevice(ctx->devs[i].id));
        TrajDev &d = t->dev[i];
        const size_t want_pos = max_frames * 3 * npad1 * sizeof(double), want_stage = t->stage_frames * frame_bytes;
        d.pos = static_cast<double *>(spare_take(ctx->devs[i], want_pos));
        cudaError_t e = d.pos ? cudaSuccess : cudaMalloc(&d.pos, want_pos);
        if (e != cudaSuccess) {
            // the spares of this device may be what is in the way: give them back to the driver and try once more
            cudaGetLastError();
            {
                std::lock_guard<std::mutex> lock(*ctx->devs[i].hstage_mutex);
                for (int k = 0; k < kSpares; ++k) {
                    cudaFree(ctx->devs[i].spare[k].ptr);
                    ctx->devs[i].spare[k] = Dev::Spare();
                }
            }
            e = cudaMalloc(&d.pos, want_pos);
        }
        if (e != cudaSuccess) {
            free_traj_dev(t.get());
            return fail(AGOFRT_ERR_CUDA, "cudaMalloc of the %zu-frame window (%zu bytes) failed: %s", max_frames,
                        max_frames * 3 * npad1 * sizeof(double), cudaGetErrorString(e));
        }
        d.pos_bytes = want_pos;
        CU(pool_malloc(ctx->devs[i], &d.box6, max_frames * 6 * sizeof(double)));
        CU(pool_malloc(ctx->devs[i], &d.bounds, max_frames * 6 * sizeof(double)));
        d.stage = static_cast<double *>(spare_take(ctx->devs[i], want_stage));
        if (!d.stage) CU(cudaMalloc(&d.stage, want_stage));
        d.stage_bytes = want_stage;
        CU(pool_malloc(ctx->devs[i], &d.box_stage, max_frames * box_stride * sizeof(double)));
        CU(pool_malloc(ctx->devs[i], &d.perm, npad1 * sizeof(int)));
        CU(pool_malloc(ctx->devs[i], &d.type_pad, npad1 * sizeof(int)));
        CU(pool_malloc(ctx->devs[i], &d.type_start, (ntypes + 1) * sizeof(int)));
        CU(pool_malloc(ctx->devs[i], &d.flags, 8 * sizeof(unsigned int)));
        CU(pool_malloc(ctx->devs[i], &d.probe, 4 * sizeof(double)));
        CU(cudaStreamCreateWithFlags(&d.up, cudaStreamNonBlocking));
        CU(cudaMemset(d.flags, 0, 8 * sizeof(unsigned int)));
        if (t->npad > 0) CU(cudaMemcpy(d.type_pad, t->type_pad.data(), t->npad * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d.type_start, t->type_start.data(), (ntypes + 1) * sizeof(int), cudaMemcpyHostToDevice));
    }
    *out = t.release();
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_traj_destroy(agofrt_traj *t) try {
    if (!t) return AGOFRT_OK;
    const bool debug = getenv("AGOFRT_DEBUG") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    free_traj_dev(t);
    delete t;
    if (debug)
        fprintf(stderr, "[agofrt] window released in %.1f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

static inline uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// Type-major, Morton-ordered permutation from the first frame of the window.  Only the ORDER of the
// device slots depends on it -- never a count -- so a poor key costs speed, not correctness.
static void build_perm(agofrt_traj *t, const double *pos0, const double *box_row) {
    const size_t n = t->natoms;
    std::vector<uint64_t> key(n);
    const double Lx = 2 * box_row[3], Ly = 2 * box_row[4], Lz = 2 * box_row[5];
    const bool tri = t->stride == 9;
    const double xy = tri ? box_row[6] : 0.0, xz = tri ? box_row[7] : 0.0, yz = tri ? box_row[8] : 0.0;
    const bool okbox = Lx > 0 && Ly > 0 && Lz > 0 && std::isfinite(Lx) && std::isfinite(Ly) && std::isfinite(Lz);
    int bits = 1;
    while (bits < 10 && (static_cast<size_t>(1) << (3 * bits)) * 4 < n) ++bits;
    const double scale = static_cast<double>(1u << bits);
    for (size_t i = 0; i < n; ++i) {
        uint32_t m = 0;
        if (okbox) {
            const double x = pos0[3 * i], y = pos0[3 * i + 1], z = pos0[3 * i + 2];
            double sz = z / Lz;
            double sy = (y - yz * sz) / Ly;
            double sx = (x - xy * sy - xz * sz) / Lx;
            sx -= std::floor(sx);
            sy -= std::floor(sy);
            sz -= std::floor(sz);
            if (std::isfinite(sx) && std::isfinite(sy) && std::isfinite(sz)) {
                const uint32_t cx = std::min<uint32_t>(static_cast<uint32_t>(sx * scale), (1u << bits) - 1);
                const uint32_t cy = std::min<uint32_t>(static_cast<uint32_t>(sy * scale), (1u << bits) - 1);
                const uint32_t cz = std::min<uint32_t>(static_cast<uint32_t>(sz * scale), (1u << bits) - 1);
                m = spread3(cx) | (spread3(cy) << 1) | (spread3(cz) << 2);
            }
        }
        key[i] = (static_cast<uint64_t>(t->type_id[i]) << 32) | m;
    }
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
    std::fill(t->perm.begin(), t->perm.end(), -1);
    std::vector<int> fill(t->type_start.begin(), t->type_start.end() - 1);
    for (size_t k = 0; k < n; ++k) {
        const int a = order[k];
        t->perm[fill[t->type_id[a]]++] = a;
    }
}

// Copy `bytes` from pageable host memory into a page-locked slot with several host threads (one thread moves
// about 10 GB/s; a PCIe 5 x16 link takes 50).
static void parallel_copy(void *dst, const void *src, size_t bytes) {
    const size_t min_per_thread = 4u << 20;
    size_t nth = std::min<size_t>({8, std::max(1u, std::thread::hardware_concurrency()), bytes / min_per_thread});
    if (nth <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> pool;
    const size_t each = (bytes / nth + 4095) & ~static_cast<size_t>(4095);
    for (size_t k = 0; k < nth; ++k) {
        const size_t o = k * each;
        if (o >= bytes) break;
        const size_t n = std::min(each, bytes - o);
        pool.emplace_back([=]() { memcpy(static_cast<char *>(dst) + o, static_cast<const char *>(src) + o, n); });
    }
    for (std::thread &th : pool) th.join();
}

// the same for a list of pieces (the chunks of many small frames): the pieces are dealt to the threads by bytes
struct CopyPiece {
    void *dst;
    const void *src;
    size_t bytes;
};
static void parallel_copy_pieces(const std::vector<CopyPiece> &pieces) {
    size_t total = 0;
    for (const CopyPiece &p : pieces) total += p.bytes;
    const size_t nth = std::min<size_t>({8, std::max(1u, std::thread::hardware_concurrency()), std::max<size_t>(1, total / (1u << 20))});
    if (nth <= 1 || pieces.size() < 2 * nth) {
        for (const CopyPiece &p : pieces) parallel_copy(p.dst, p.src, p.bytes);
        return;
    }
    std::vector<std::thread> pool;
    const size_t each = (pieces.size() + nth - 1) / nth;
    for (size_t k = 0; k < nth; ++k) {
        const size_t a = k * each, b = std::min(pieces.size(), a + each);
        if (a >= b) break;
        pool.emplace_back([&pieces, a, b]() {
            for (size_t i = a; i < b; ++i) memcpy(pieces[i].dst, pieces[i].src, pieces[i].bytes);
        });
    }
    for (std::thread &th : pool) th.join();
}

static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// The window upload.  flags: AGOFRT_UP_WRAP applies BaseTrajectory::pbc_wrap on the device before the layout change,
// AGOFRT_UP_WRITEBACK hands the wrapped frames back in pos_back, AGOFRT_UP_SHARED deals the frames to the devices of
// the communicator -- every device copies, wraps and lays out only its share, then the shares travel device to device
// (grouped ncclBroadcast = an all-gather with unequal counts, over NVLink) -- so the window crosses PCIe ONCE per box
// instead of once per GPU.  Pageable source memory goes through two page-locked slots filled by several host threads.
// raw dump records as the source of a window (agofrt_traj_upload_records): frame f = the chunks
// [frame_chunk[f], frame_chunk[f+1]) of chunk_ptr / chunk_atoms, natoms records of 8 doubles in all
struct RecordSource {
    const void *const *chunk_ptr;
    const int *chunk_atoms;
    const size_t *frame_chunk;
};

static int upload_impl(agofrt_traj *t, size_t first_frame, size_t nframes, const double *pos_in, const double *box_internal,
                       unsigned flags, double *pos_back, const RecordSource *rec = nullptr) {
    if (!t) return fail(AGOFRT_ERR_ARG, "traj is NULL");
    if (nframes > t->max_frames) return fail(AGOFRT_ERR_ARG, "window of %zu frames > max_frames %zu", nframes, t->max_frames);
    if (nframes > 0 && (!box_internal || (t->natoms > 0 && !pos_in && !rec))) return fail(AGOFRT_ERR_ARG, "NULL buffer");
    const bool wrap = (flags & AGOFRT_UP_WRAP) != 0;
    if (!(flags & AGOFRT_UP_WRITEBACK) || (!wrap && !rec)) pos_back = nullptr;
    agofrt_ctx *ctx = t->ctx;
    const bool debug = getenv("AGOFRT_DEBUG") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point a) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
    };
    double ms_perm = 0, ms_copy = 0, ms_setup = 0, ms_loop = 0;
    t->first_frame = first_frame;
    t->nframes = nframes;
    t->box6.assign(nframes * 6, 0.0);
    t->bounds.assign(nframes * 6, 0.0);
    t->has_inf = false;
    t->has_nan = false;
    t->bad_box = false;
    if (nframes == 0) return AGOFRT_OK;

    for (size_t f = 0; f < nframes; ++f) {
        const double *r = box_internal + f * t->stride;
        double *o = &t->box6[f * 6];
        o[0] = r[3];
        o[1] = r[4];
        o[2] = r[5];
        o[3] = t->stride == 9 ? r[6] : 0.0;
        o[4] = t->stride == 9 ? r[7] : 0.0;
        o[5] = t->stride == 9 ? r[8] : 0.0;
        for (int k = 0; k < 3; ++k)
            if (!(o[k] > 0.0) || !std::isfinite(o[k])) t->bad_box = true;
        for (int k = 3; k < 6; ++k)
            if (!std::isfinite(o[k])) t->bad_box = true;
    }
    if (wrap && t->bad_box) {
        t->nframes = 0;
        return fail(AGOFRT_ERR_NONFINITE, "a box edge of the window is not positive and finite (the reference's wrap would not terminate)");
    }
    // The spatial sort only serves locality (the group filter of the sparse kernels); atoms of a condensed
    // phase move little over a few hundred frames, so the permutation of an earlier window is kept --
    // at 1M atoms the host sort costs as much as a whole block on 8 GPUs.
    constexpr size_t kPermRefresh = 256;
    const size_t moved = first_frame > t->perm_frame ? first_frame - t->perm_frame : t->perm_frame - first_frame;
    const bool new_perm = t->natoms > 0 && (!t->perm_valid || moved >= kPermRefresh);
    std::vector<double> frame0;
    if (new_perm && rec) {
        // the spatial sort wants the positions of the first frame in the caller's atom order: one frame parsed on the
        // host (ids resolved through the host copy of the table)
        frame0.assign(t->natoms * 3, 0.0);
        for (size_t c = rec->frame_chunk[0]; c < rec->frame_chunk[1]; ++c) {
            const double *r = static_cast<const double *>(rec->chunk_ptr[c]);
            for (int a = 0; a < rec->chunk_atoms[c]; ++a, r += 8) {
                const long id = std::lround(r[0]);
                if (id < 0 || static_cast<size_t>(id) >= t->id_table_host.size() || t->id_table_host[id] < 0) continue;
                double *o = &frame0[static_cast<size_t>(t->id_table_host[id]) * 3];
                o[0] = r[2];
                o[1] = r[3];
                o[2] = r[4];
            }
        }
    }
    if (new_perm) {
        const auto t0 = std::chrono::steady_clock::now();
        build_perm(t, rec ? frame0.data() : pos_in, box_internal);   // (shared uploads: frame 0 of the window must be valid on every rank)
        ms_perm = since(t0);
        t->perm_valid = true;
        t->perm_frame = first_frame;
        t->slot_of_stale = true;
    }

    // ---- who takes which frames ----
    const int nloc = static_cast<int>(t->dev.size());
    if (flags & AGOFRT_UP_SHARED) {
        const int rc = ensure_local_comm(ctx);
        if (rc != AGOFRT_OK) return rc;
    }
    const int world = ctx->world > 0 ? ctx->world : nloc;
    const int first_rank = ctx->world > 0 ? ctx->first_rank : 0;
    bool shared = (flags & AGOFRT_UP_SHARED) && ctx->comm_ready && world > 1 && t->npad > 0;
    for (int i = 0; i < nloc && shared; ++i) shared = ctx->devs[i].comm_up != nullptr;
    std::vector<size_t> fb(nloc, 0), fe(nloc, nframes);
    if (shared)
        for (int i = 0; i < nloc; ++i) {
            uint64_t b0 = 0, e0 = 0;
            agofrt_shard_range(nframes, first_rank + i, world, &b0, &e0);
            fb[i] = b0;
            fe[i] = e0;
        }

    const size_t frame_elems = t->natoms * 3, frame_bytes = frame_elems * sizeof(double);
    for (int i = 0; i < nloc; ++i) {
        Dev &dv = ctx->devs[i];
        TrajDev &d = t->dev[i];
        CU(cudaSetDevice(dv.id));
        {
            // [1] belongs to the kernels of agofrt_block; the rest is this upload's
            CU(cudaMemsetAsync(d.flags, 0, sizeof(unsigned int), d.up));
            CU(cudaMemsetAsync(d.flags + 2, 0, 6 * sizeof(unsigned int), d.up));
        }
        if (t->npad > 0 && new_perm)
            CU(cudaMemcpyAsync(d.perm, t->perm.data(), t->npad * sizeof(int), cudaMemcpyHostToDevice, d.up));
        CU(cudaMemcpyAsync(d.box_stage, box_internal, nframes * t->stride * sizeof(double), cudaMemcpyHostToDevice, d.up));
        CU(launch_pack_box(d.box_stage, t->stride, static_cast<int>(nframes), d.box6, d.up));
    }
    ms_setup = since(t_begin);
    if (t->npad > 0) {
        const bool pinned_src = !rec && is_pinned(pos_in);
        // bytes per frame on the host side of the copy: AoS positions, or the raw records (8 doubles per atom)
        const size_t src_frame_bytes = rec ? t->natoms * 8 * sizeof(double) : frame_bytes;
        // frames per step: what the device staging holds; from pageable memory also what a 64 MiB host slot holds
        size_t step = t->stage_frames;
        if (!pinned_src) step = std::max<size_t>(1, std::min<size_t>(step, (64u << 20) / std::max<size_t>(src_frame_bytes, 1)));
        size_t longest = 0;
        for (int i = 0; i < nloc; ++i) longest = std::max(longest, fe[i] - fb[i]);
        std::vector<std::unique_lock<std::mutex>> locks;
        if (!pinned_src)
            for (int i = 0; i < nloc; ++i) {
                Dev &dv = ctx->devs[i];
                locks.emplace_back(*dv.hstage_mutex);
                if (dv.hstage_bytes < step * src_frame_bytes) {
                    CU(cudaSetDevice(dv.id));
                    for (int k = 0; k < 2; ++k) {
                        if (dv.hstage[k]) cudaFreeHost(dv.hstage[k]);
                        dv.hstage[k] = nullptr;
                    }
                    dv.hstage_bytes = 0;
                    for (int k = 0; k < 2; ++k)
                        CU(cudaHostAlloc(reinterpret_cast<void **>(&dv.hstage[k]), step * src_frame_bytes, cudaHostAllocPortable));
                    dv.hstage_bytes = step * src_frame_bytes;
                }
            }
        if (rec)
            for (int i = 0; i < nloc; ++i) {
                TrajDev &d = t->dev[i];
                if (d.raw_frames < step) {
                    CU(cudaSetDevice(ctx->devs[i].id));
                    cudaFree(d.raw);
                    d.raw = nullptr;
                    d.raw_frames = 0;
                    CU(cudaMalloc(&d.raw, step * src_frame_bytes));
                    d.raw_frames = step;
                }
            }
        size_t slot = 0;
        for (size_t o = 0; o < longest; o += step, ++slot) {
            for (int i = 0; i < nloc; ++i) {
                if (fb[i] + o >= fe[i]) continue;
                const size_t f0 = fb[i] + o, nf = std::min(step, fe[i] - f0);
                Dev &dv = ctx->devs[i];
                TrajDev &d = t->dev[i];
                CU(cudaSetDevice(dv.id));
                const double *src = rec ? nullptr : pos_in + f0 * frame_elems;
                if (!pinned_src) {
                    double *hs = dv.hstage[slot & 1];
                    CU(cudaEventSynchronize(dv.hstage_free[slot & 1]));   // the copy that last read this slot is done
                    const auto t0 = std::chrono::steady_clock::now();
                    if (rec) {
                        // the chunks of every frame, one after the other: natoms records per frame
                        char *w = reinterpret_cast<char *>(hs);
                        std::vector<CopyPiece> pieces;
                        for (size_t f = f0; f < f0 + nf; ++f) {
                            size_t atoms = 0;
                            for (size_t c = rec->frame_chunk[f]; c < rec->frame_chunk[f + 1]; ++c) {
                                const size_t nb = static_cast<size_t>(rec->chunk_atoms[c]) * 8 * sizeof(double);
                                atoms += static_cast<size_t>(rec->chunk_atoms[c]);
                                if (atoms > t->natoms) return fail(AGOFRT_ERR_ARG, "frame %zu holds more records than atoms", first_frame + f);
                                pieces.push_back({w, rec->chunk_ptr[c], nb});
                                w += nb;
                            }
                            if (atoms != t->natoms) return fail(AGOFRT_ERR_ARG, "frame %zu holds %zu records for %zu atoms", first_frame + f, atoms, t->natoms);
                        }
                        parallel_copy_pieces(pieces);
                    } else {
                        parallel_copy(hs, src, nf * frame_bytes);
                    }
                    ms_copy += since(t0);
                    src = hs;
                }
                if (rec) {
                    CU(cudaMemcpyAsync(d.raw, src, nf * src_frame_bytes, cudaMemcpyHostToDevice, d.up));
                    CU(cudaEventRecord(dv.hstage_free[slot & 1], d.up));
                    // header part of Trajectory::set_access_at's frame loop on the device: id -> slot, scatter x y z
                    CU(launch_parse_records(d.raw, static_cast<int>(t->natoms), static_cast<int>(nf), d.id_table,
                                            static_cast<int>(t->id_table_len), d.slot_type, d.stage, d.flags, d.up));
                } else {
                    CU(cudaMemcpyAsync(d.stage, src, nf * frame_bytes, cudaMemcpyHostToDevice, d.up));
                    if (!pinned_src) CU(cudaEventRecord(dv.hstage_free[slot & 1], d.up));
                }
                if (wrap) {
                    CU(launch_pbc_wrap(d.stage, static_cast<int>(t->natoms), static_cast<int>(nf), d.box_stage + f0 * t->stride,
                                       t->stride, d.flags + 3, d.up));
                }
                if (pos_back && (shared || i == 0))
                    CU(cudaMemcpyAsync(pos_back + f0 * frame_elems, d.stage, nf * frame_bytes, cudaMemcpyDeviceToHost, d.up));
                CU(launch_gather_soa(d.stage, d.perm, static_cast<int>(t->natoms), t->npad, static_cast<int>(nf),
                                     d.pos + f0 * 3 * static_cast<size_t>(t->npad), d.up));
            }
        }
        ms_loop = since(t_begin) - ms_setup;
        // coordinate bounds per frame: every device for its share, or device 0 for the replicated window
        for (int i = 0; i < nloc; ++i) {
            if (!shared && i > 0) break;
            if (fe[i] <= fb[i]) continue;
            Dev &dv = ctx->devs[i];
            TrajDev &d = t->dev[i];
            CU(cudaSetDevice(dv.id));
            CU(launch_frame_bounds(d.pos + fb[i] * 3 * static_cast<size_t>(t->npad), d.perm, t->npad, static_cast<int>(fe[i] - fb[i]),
                                   d.bounds + fb[i] * 6, d.flags, d.up));
        }
        if (shared) {
            // the shares change hands: rank r's frames and their bounds go to everybody, flags are combined
            NcclApi &api = nccl_api();
            NC(api.GroupStart());
            if (nframes % static_cast<size_t>(world) == 0) {
                // equal shares: one in-place all-gather of the positions and one of the bounds
                const size_t per = nframes / static_cast<size_t>(world);
                for (int i = 0; i < nloc; ++i) {
                    TrajDev &d = t->dev[i];
                    const size_t r = static_cast<size_t>(first_rank + i);
                    NC(api.AllGather(d.pos + r * per * 3 * static_cast<size_t>(t->npad), d.pos, per * 3 * static_cast<size_t>(t->npad),
                                     ncclDouble, ctx->devs[i].comm_up, d.up));
                    NC(api.AllGather(d.bounds + r * per * 6, d.bounds, per * 6, ncclDouble, ctx->devs[i].comm_up, d.up));
                }
            } else {
                for (int r = 0; r < world; ++r) {
                    uint64_t b0 = 0, e0 = 0;
                    agofrt_shard_range(nframes, r, world, &b0, &e0);
                    if (e0 <= b0) continue;
                    for (int i = 0; i < nloc; ++i) {
                        Dev &dv = ctx->devs[i];
                        TrajDev &d = t->dev[i];
                        double *ppos = d.pos + b0 * 3 * static_cast<size_t>(t->npad);
                        NC(api.Broadcast(ppos, ppos, (e0 - b0) * 3 * static_cast<size_t>(t->npad), ncclDouble, r, dv.comm_up, d.up));
                        NC(api.Broadcast(d.bounds + b0 * 6, d.bounds + b0 * 6, (e0 - b0) * 6, ncclDouble, r, dv.comm_up, d.up));
                    }
                }
            }
            for (int i = 0; i < nloc; ++i)
                NC(api.AllReduce(t->dev[i].flags, t->dev[i].flags, 8, ncclUint32, ncclMax, ctx->devs[i].comm_up, t->dev[i].up));
            NC(api.GroupEnd());
        }
        Dev &dv = ctx->devs[0];
        TrajDev &d = t->dev[0];
        CU(cudaSetDevice(dv.id));
        unsigned int hflags[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        CU(cudaMemcpyAsync(t->bounds.data(), d.bounds, nframes * 6 * sizeof(double), cudaMemcpyDeviceToHost, d.up));
        CU(cudaMemcpyAsync(hflags, d.flags, sizeof(hflags), cudaMemcpyDeviceToHost, d.up));
        CU(cudaStreamSynchronize(d.up));
        t->has_inf = hflags[0] != 0;
        t->has_nan = hflags[2] != 0;
        if (hflags[3] || hflags[4] || hflags[5]) {
            for (int i = 0; i < nloc; ++i) {
                cudaSetDevice(ctx->devs[i].id);
                cudaStreamSynchronize(t->dev[i].up);
            }
            t->nframes = 0;
            if (hflags[4])
                return fail(AGOFRT_ERR_ARG, "a record of the window carries an atom id that was not in the table (agofrt_traj_set_ids)");
            if (hflags[5]) return fail(AGOFRT_ERR_RETYPED, "the type of an atom changes inside the window");
            return fail(AGOFRT_ERR_NONFINITE, "minimum image did not converge while wrapping (non-finite or absurdly far coordinate)");
        }
    }
    for (int i = 0; i < nloc; ++i) {
        CU(cudaSetDevice(ctx->devs[i].id));
        CU(cudaStreamSynchronize(t->dev[i].up));
    }
    if (debug)
        fprintf(stderr, "[agofrt] upload of %zu frames x %zu atoms (%s%s%s): %.1f ms (setup %.1f of which permutation %.1f, copy loop %.1f of which host staging copies %.1f)\n", nframes,
                t->natoms, wrap ? "wrap " : "", shared ? "shared " : "replicated ", rec ? "records" : (is_pinned(pos_in) ? "pinned" : "pageable"),
                since(t_begin), ms_setup, ms_perm, ms_loop, ms_copy);
    return AGOFRT_OK;
}

extern "C" int agofrt_traj_upload(agofrt_traj *t, size_t first_frame, size_t nframes, const double *pos_aos,
                                  const double *box_internal) try {
    return upload_impl(t, first_frame, nframes, pos_aos, box_internal, 0, nullptr);
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_traj_upload_wrap(agofrt_traj *t, size_t first_frame, size_t nframes, double *pos_aos_inout,
                                       const double *box_internal) try {
    return upload_impl(t, first_frame, nframes, pos_aos_inout, box_internal, AGOFRT_UP_WRAP | AGOFRT_UP_WRITEBACK, pos_aos_inout);
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_traj_upload_ex(agofrt_traj *t, size_t first_frame, size_t nframes, const double *pos_aos,
                                     const double *box_internal, unsigned flags, double *pos_wrapped_out) try {
    if ((flags & AGOFRT_UP_WRITEBACK) && !(flags & AGOFRT_UP_WRAP))
        return fail(AGOFRT_ERR_ARG, "AGOFRT_UP_WRITEBACK needs AGOFRT_UP_WRAP");
    if ((flags & AGOFRT_UP_WRITEBACK) && !pos_wrapped_out && nframes > 0 && t && t->natoms > 0)
        return fail(AGOFRT_ERR_ARG, "AGOFRT_UP_WRITEBACK needs pos_wrapped_out");
    return upload_impl(t, first_frame, nframes, pos_aos, box_internal, flags, pos_wrapped_out);
} catch (...) {
    return on_exception();
}

// The id -> slot table of a LAMMPS dump (Trajectory's id_map, reference lib/src/trajectory.cpp:133-189) and the raw
// type every slot carries, on the devices: what agofrt_traj_upload_records resolves records with.
extern "C" int agofrt_traj_set_ids(agofrt_traj *t, const int *slot_to_id, const int *slot_raw_type) try {
    if (!t || ((!slot_to_id || !slot_raw_type) && t->natoms > 0)) return fail(AGOFRT_ERR_ARG, "NULL argument");
    int max_id = -1;
    for (size_t k = 0; k < t->natoms; ++k) {
        if (slot_to_id[k] < 0) return fail(AGOFRT_ERR_ARG, "negative atom id");
        max_id = std::max(max_id, slot_to_id[k]);
    }
    if (static_cast<size_t>(max_id) > t->natoms * 8 + 1024) return fail(AGOFRT_ERR_TOO_LARGE, "atom ids are too sparse for a flat table (largest id %d for %zu atoms)", max_id, t->natoms);
    t->id_table_host.assign(static_cast<size_t>(max_id + 1), -1);
    for (size_t k = 0; k < t->natoms; ++k) {
        if (t->id_table_host[slot_to_id[k]] >= 0) return fail(AGOFRT_ERR_ARG, "atom id %d appears twice", slot_to_id[k]);
        t->id_table_host[slot_to_id[k]] = static_cast<int>(k);
    }
    t->id_table_len = t->id_table_host.size();
    for (size_t i = 0; i < t->dev.size(); ++i) {
        TrajDev &d = t->dev[i];
        CU(cudaSetDevice(t->ctx->devs[i].id));
        cudaFree(d.id_table);
        cudaFree(d.slot_type);
        d.id_table = d.slot_type = nullptr;
        CU(cudaMalloc(&d.id_table, std::max<size_t>(t->id_table_len, 1) * sizeof(int)));
        CU(cudaMalloc(&d.slot_type, std::max<size_t>(t->natoms, 1) * sizeof(int)));
        if (t->id_table_len) CU(cudaMemcpy(d.id_table, t->id_table_host.data(), t->id_table_len * sizeof(int), cudaMemcpyHostToDevice));
        if (t->natoms) CU(cudaMemcpy(d.slot_type, slot_raw_type, t->natoms * sizeof(int), cudaMemcpyHostToDevice));
    }
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_traj_upload_records(agofrt_traj *t, size_t first_frame, size_t nframes, const void *const *chunk_ptr,
                                          const int *chunk_atoms, const size_t *frame_chunk, const double *box_internal,
                                          unsigned flags, double *pos_out) try {
    if (!t) return fail(AGOFRT_ERR_ARG, "traj is NULL");
    if (nframes > 0 && (!chunk_ptr || !chunk_atoms || !frame_chunk)) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (t->natoms > 0 && t->id_table_len == 0) return fail(AGOFRT_ERR_ARG, "agofrt_traj_set_ids must be called before agofrt_traj_upload_records");
    if ((flags & AGOFRT_UP_WRITEBACK) && !pos_out && nframes > 0 && t->natoms > 0) return fail(AGOFRT_ERR_ARG, "AGOFRT_UP_WRITEBACK needs pos_out");
    RecordSource rec{chunk_ptr, chunk_atoms, frame_chunk};
    return upload_impl(t, first_frame, nframes, nullptr, box_internal, flags, pos_out, &rec);
} catch (...) {
    return on_exception();
}

// Per-frame rotation matrices next to the window (Trajectory_numpy keeps Q of the QR rotation that brings a general cell
// into the LAMMPS frame, reference lib/src/trajectory_numpy.cpp:120,131, lib/include/triclinic.h:71-73): device-resident
// like positions and cells, for kernels that must rotate vectors back to the laboratory frame.  g(r,t) itself is
// rotation-invariant and never reads them.
extern "C" int agofrt_traj_set_rotation(agofrt_traj *t, size_t first_frame, size_t nframes, const double *q) try {
    if (!t || (!q && nframes > 0)) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (nframes > t->max_frames) return fail(AGOFRT_ERR_ARG, "%zu rotation matrices for a window of at most %zu frames", nframes, t->max_frames);
    for (size_t i = 0; i < t->dev.size(); ++i) {
        TrajDev &d = t->dev[i];
        CU(cudaSetDevice(t->ctx->devs[i].id));
        if (!d.rot) CU(cudaMalloc(&d.rot, t->max_frames * 9 * sizeof(double)));
        if (nframes) CU(cudaMemcpyAsync(d.rot, q, nframes * 9 * sizeof(double), cudaMemcpyHostToDevice, d.up));
        CU(cudaStreamSynchronize(d.up));
    }
    t->rot_first = first_frame;
    t->rot_frames = nframes;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_traj_get_rotation(agofrt_traj *t, size_t frame, double *q9) try {
    if (!t || !q9) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (frame < t->rot_first || frame >= t->rot_first + t->rot_frames)
        return fail(AGOFRT_ERR_WINDOW, "no rotation matrix on the device for frame %zu", frame);
    TrajDev &d = t->dev[t->dev.size() - 1];   // (the last device: every device holds a copy)
    CU(cudaSetDevice(t->ctx->devs[t->dev.size() - 1].id));
    CU(cudaMemcpy(q9, d.rot + (frame - t->rot_first) * 9, 9 * sizeof(double), cudaMemcpyDeviceToHost));
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// Frames [first_frame, first_frame + nframes) of the device window back in the caller's atom order: what the host
// would hold after Trajectory::set_access_at / the Trajectory_numpy constructor (wrapped when the window was uploaded
// with AGOFRT_UP_WRAP).  The host classes call it the first time somebody asks for host positions.
extern "C" int agofrt_traj_download(agofrt_traj *t, size_t first_frame, size_t nframes, double *pos_aos) try {
    if (!t || (!pos_aos && nframes > 0)) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (nframes == 0 || t->natoms == 0) return AGOFRT_OK;
    if (first_frame < t->first_frame || first_frame + nframes > t->first_frame + t->nframes)
        return fail(AGOFRT_ERR_WINDOW, "frames [%zu,%zu) are not in the uploaded window", first_frame, first_frame + nframes);
    Dev &dv = t->ctx->devs[0];
    TrajDev &d = t->dev[0];
    CU(cudaSetDevice(dv.id));
    const size_t frame_elems = t->natoms * 3;
    for (size_t f0 = 0; f0 < nframes; f0 += t->stage_frames) {
        const size_t nf = std::min(t->stage_frames, nframes - f0);
        const size_t rel = first_frame - t->first_frame + f0;
        CU(launch_scatter_aos(d.pos + rel * 3 * static_cast<size_t>(t->npad), d.perm, static_cast<int>(t->natoms), t->npad,
                              static_cast<int>(nf), d.stage, d.up));
        CU(cudaMemcpyAsync(pos_aos + f0 * frame_elems, d.stage, nf * frame_elems * sizeof(double), cudaMemcpyDeviceToHost, d.up));
    }
    CU(cudaStreamSynchronize(d.up));
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_traj_download_frame(agofrt_traj *t, size_t frame, double *pos_aos) try {
    if (!t || !pos_aos) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (frame < t->first_frame || frame >= t->first_frame + t->nframes)
        return fail(AGOFRT_ERR_WINDOW, "frame %zu is not in the uploaded window", frame);
    if (t->natoms == 0) return AGOFRT_OK;
    Dev &dv = t->ctx->devs[0];
    TrajDev &d = t->dev[0];
    CU(cudaSetDevice(dv.id));
    const size_t rel = frame - t->first_frame;
    CU(launch_scatter_aos(d.pos + rel * 3 * static_cast<size_t>(t->npad), d.perm, static_cast<int>(t->natoms), t->npad, 1,
                          d.stage, dv.stream));
    CU(cudaMemcpyAsync(pos_aos, d.stage, t->natoms * 3 * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
    CU(cudaStreamSynchronize(dv.stream));
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_pbc_wrap(agofrt_ctx *ctx, double *pos_aos, size_t nframes, size_t natoms,
                               const double *box_internal, int box_stride) try {
    if (!ctx) return fail(AGOFRT_ERR_ARG, "ctx is NULL");
    if (box_stride != 6 && box_stride != 9) return fail(AGOFRT_ERR_ARG, "box_stride must be 6 or 9");
    if (nframes == 0 || natoms == 0) return AGOFRT_OK;
    if (!pos_aos || !box_internal) return fail(AGOFRT_ERR_ARG, "NULL buffer");
    for (size_t f = 0; f < nframes; ++f)
        for (int k = 3; k < box_stride; ++k) {
            const double v = box_internal[f * box_stride + k];
            if (!std::isfinite(v) || (k < 6 && !(v > 0.0)))
                return fail(AGOFRT_ERR_NONFINITE, "frame %zu: box entry %d = %g (the reference would not terminate)", f, k, v);
        }
    Dev &dv = ctx->devs[0];
    CU(cudaSetDevice(dv.id));
    const size_t frame_bytes = natoms * 3 * sizeof(double);
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(nframes, (256u << 20) / frame_bytes));
    double *dpos = nullptr, *dbox = nullptr;
    unsigned int *dflag = nullptr;
    CU(cudaMalloc(&dpos, chunk * frame_bytes));
    CU(cudaMalloc(&dbox, nframes * box_stride * sizeof(double)));
    CU(cudaMalloc(&dflag, sizeof(unsigned int)));
    int rc = AGOFRT_OK;
    auto body = [&]() -> int {
        CU(cudaMemsetAsync(dflag, 0, sizeof(unsigned int), dv.stream));
        CU(cudaMemcpyAsync(dbox, box_internal, nframes * box_stride * sizeof(double), cudaMemcpyHostToDevice, dv.stream));
        for (size_t f0 = 0; f0 < nframes; f0 += chunk) {
            const size_t nf = std::min(chunk, nframes - f0);
            CU(cudaMemcpyAsync(dpos, pos_aos + f0 * natoms * 3, nf * frame_bytes, cudaMemcpyHostToDevice, dv.stream));
            CU(launch_pbc_wrap(dpos, static_cast<int>(natoms), static_cast<int>(nf), dbox + f0 * box_stride, box_stride,
                               dflag, dv.stream));
            CU(cudaMemcpyAsync(pos_aos + f0 * natoms * 3, dpos, nf * frame_bytes, cudaMemcpyDeviceToHost, dv.stream));
        }
        unsigned int flag = 0;
        CU(cudaMemcpyAsync(&flag, dflag, sizeof(flag), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaStreamSynchronize(dv.stream));
        if (flag) return fail(AGOFRT_ERR_NONFINITE, "minimum image did not converge (non-finite or absurdly far coordinate)");
        return AGOFRT_OK;
    };
    rc = body();
    cudaFree(dpos);
    cudaFree(dbox);
    cudaFree(dflag);
    return rc;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_traj_d2_all(agofrt_traj *t, size_t frame_i, size_t frame_j, double *out) try {
    if (!t || !out) return fail(AGOFRT_ERR_ARG, "NULL argument");
    for (size_t f : {frame_i, frame_j})
        if (f < t->first_frame || f >= t->first_frame + t->nframes)
            return fail(AGOFRT_ERR_WINDOW, "frame %zu is not in the uploaded window", f);
    if (t->natoms == 0) return AGOFRT_OK;
    if (t->natoms > 4096) return fail(AGOFRT_ERR_ARG, "d2_all is a small-N probe (natoms <= 4096)");
    if (t->bad_box || t->has_inf) return fail(AGOFRT_ERR_NONFINITE, "non-finite coordinates or invalid box");
    Dev &dv = t->ctx->devs[0];
    TrajDev &d = t->dev[0];
    CU(cudaSetDevice(dv.id));
    const size_t ri = frame_i - t->first_frame, rj = frame_j - t->first_frame;
    double *dout = nullptr;
    const size_t bytes = t->natoms * t->natoms * 4 * sizeof(double);
    CU(cudaMalloc(&dout, bytes));
    auto body = [&]() -> int {
        CU(cudaMemsetAsync(d.flags + 1, 0, sizeof(unsigned int), dv.stream));
        CU(launch_d2_all(d.pos + ri * 3 * static_cast<size_t>(t->npad), d.pos + rj * 3 * static_cast<size_t>(t->npad),
                         d.box6 + ri * 6, t->stride == 9, d.perm, static_cast<int>(t->natoms), t->npad, dout,
                         d.flags + 1, dv.stream));
        CU(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, dv.stream));
        unsigned int flag = 0;
        CU(cudaMemcpyAsync(&flag, d.flags + 1, sizeof(flag), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaStreamSynchronize(dv.stream));
        if (flag) return fail(AGOFRT_ERR_NONFINITE, "minimum image did not converge");
        return AGOFRT_OK;
    };
    const int rc = body();
    cudaFree(dout);
    return rc;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_traj_d2_pair(agofrt_traj *t, size_t atom_i, size_t atom_j, size_t frame_i, size_t frame_j,
                                   double *out4) try {
    if (!t || !out4) return fail(AGOFRT_ERR_ARG, "NULL argument");
    for (size_t f : {frame_i, frame_j})
        if (f < t->first_frame || f >= t->first_frame + t->nframes)
            return fail(AGOFRT_ERR_WINDOW, "frame %zu is not in the uploaded window", f);
    if (atom_i >= t->natoms || atom_j >= t->natoms) return fail(AGOFRT_ERR_ARG, "Atom index out of range");
    if (t->bad_box || t->has_inf) return fail(AGOFRT_ERR_NONFINITE, "non-finite coordinates or invalid box");
    // slot of an atom = inverse of the upload's permutation
    if (t->slot_of.size() != t->natoms || t->slot_of_first != t->first_frame || t->slot_of_stale) {
        t->slot_of.assign(t->natoms, -1);
        for (int s = 0; s < t->npad; ++s)
            if (t->perm[s] >= 0) t->slot_of[t->perm[s]] = s;
        t->slot_of_first = t->first_frame;
        t->slot_of_stale = false;
    }
    Dev &dv = t->ctx->devs[0];
    TrajDev &d = t->dev[0];
    CU(cudaSetDevice(dv.id));
    const size_t ri = frame_i - t->first_frame, rj = frame_j - t->first_frame;
    // the AoS staging buffer is free between uploads: 4 doubles of it receive the result
    CU(cudaMemsetAsync(d.flags + 1, 0, sizeof(unsigned int), dv.stream));
    CU(launch_d2_pair(d.pos + ri * 3 * static_cast<size_t>(t->npad), d.pos + rj * 3 * static_cast<size_t>(t->npad),
                      d.box6 + ri * 6, t->stride == 9, t->slot_of[atom_i], t->slot_of[atom_j], t->npad, d.probe,
                      d.flags + 1, dv.stream));
    unsigned int flag = 0;
    CU(cudaMemcpyAsync(out4, d.probe, 4 * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
    CU(cudaMemcpyAsync(&flag, d.flags + 1, sizeof(flag), cudaMemcpyDeviceToHost, dv.stream));
    CU(cudaStreamSynchronize(dv.stream));
    if (flag) return fail(AGOFRT_ERR_NONFINITE, "minimum image did not converge");
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// ---------------------------------------------------------------------------------------------
// plan: the exact threshold table
// ---------------------------------------------------------------------------------------------
// The reference's bin index for an accepted pair (lib/src/gofrt.cpp:114-117), clipped to
// [-1, nbin]: sqrt in double, subtract and divide in double, ROUND TO FLOAT, floorf, (int).
// This translation unit is compiled without FMA contraction and without fast-math.
static int ref_bin_clipped(double d2, double rmin, double dr, unsigned nbin) {
    volatile double d = std::sqrt(d2);
    volatile double q = (d - rmin) / dr;
    volatile float qf = static_cast<float>(q);
    const float f = floorf(qf);
    if (!(f >= 0.0f)) return -1;
    if (f >= static_cast<float>(nbin)) return static_cast<int>(nbin);
    return static_cast<int>(f);
}

static inline double bits_to_double(uint64_t b) {
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static inline uint64_t double_to_bits(double d) {
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
}

// thresholds[k] = smallest non-negative double d2 with ref_bin_clipped(d2) >= k, k = 0..nbin.
// ref_bin_clipped is monotone non-decreasing in d2 (a composition of monotone roundings), so a
// bisection over the bit patterns of the non-negative doubles finds it exactly.
static void build_thresholds(double rmin, double dr, unsigned nbin, std::vector<double> &thr) {
    thr.resize(nbin + 1);
    const uint64_t inf_bits = 0x7ff0000000000000ull;
    uint64_t lo_start = 0;
    for (unsigned k = 0; k <= nbin; ++k) {
        uint64_t lo = lo_start, hi = inf_bits;  // answer in [lo, hi]
        if (ref_bin_clipped(bits_to_double(lo), rmin, dr, nbin) >= static_cast<int>(k)) {
            thr[k] = bits_to_double(lo);
            continue;
        }
        // invariant: f(lo) < k <= f(hi)   (f(+inf) = nbin >= k)
        while (hi - lo > 1) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (ref_bin_clipped(bits_to_double(mid), rmin, dr, nbin) >= static_cast<int>(k))
                hi = mid;
            else
                lo = mid;
        }
        thr[k] = bits_to_double(hi);
        lo_start = lo;
    }
}

// Exact bin of d2 from the folded table (the kernel's own definition): k with thr[k] <= d2 < thr[k+1], else -1.
static int exact_bin(const std::vector<double> &thr, unsigned nbin, double d2) {
    if (!(d2 >= thr[0]) || !(d2 < thr[nbin])) return -1;
    unsigned lo = 0, hi = nbin;  // thr[lo] <= d2 < thr[hi]
    while (hi - lo > 1) {
        const unsigned mid = (lo + hi) / 2;
        if (d2 >= thr[mid]) lo = mid; else hi = mid;
    }
    return static_cast<int>(lo);
}

// Run the device's float guess on probes at, next to and between all bin thresholds; a single
// unflagged wrong guess disables the safe-zone mode for this plan (the threshold mode is always exact).
static int validate_safe_zone(agofrt_plan *p) {
    const unsigned nbin = p->nbin;
    std::vector<double> probes;
    auto add_around = [&](double t) {
        if (!std::isfinite(t)) return;
        uint64_t b = double_to_bits(t);
        for (int d = -3; d <= 3; ++d) {
            const uint64_t bb = b + static_cast<uint64_t>(static_cast<int64_t>(d));
            if (d < 0 && b < static_cast<uint64_t>(-d)) continue;
            probes.push_back(bits_to_double(bb));
        }
    };
    probes.push_back(0.0);
    for (unsigned k = 0; k <= nbin; ++k) {
        add_around(p->thr[k]);
        add_around(p->thr_full[k]);
    }
    for (unsigned k = 0; k < nbin; ++k) {
        const double a = p->thr[k], b = p->thr[k + 1];
        if (!(a < b) || !std::isfinite(b)) continue;
        for (int i = 1; i < 32; ++i) probes.push_back(a + (b - a) * (i / 32.0));
        // also points a relative 1e-7 .. 1e-3 inside either edge (where the float guess is weakest)
        for (double rel : {1e-7, 1e-6, 1e-5, 1e-4, 1e-3}) {
            probes.push_back(a + (b - a) * rel);
            probes.push_back(b - (b - a) * rel);
        }
    }
    const double top = p->thr[nbin];
    if (std::isfinite(top))
        for (double f : {1.0000001, 1.001, 1.5, 4.0, 100.0, 1e6, 1e30, 1e300}) probes.push_back(top * f + 1e-300);
    probes.push_back(std::numeric_limits<double>::infinity());
    probes.push_back(std::numeric_limits<double>::quiet_NaN());   // ghost slots
    for (double tiny : {4.9e-324, 1e-310, 1e-300, 1e-45, 1e-38, 1e-30}) probes.push_back(tiny);
    std::vector<int> expected(probes.size());
    for (size_t i = 0; i < probes.size(); ++i) expected[i] = exact_bin(p->thr, nbin, probes[i]);

    Dev &dv = p->traj->ctx->devs[0];
    CU(cudaSetDevice(dv.id));
    double *dprobe = nullptr;
    int *dexp = nullptr;
    unsigned int *dbad = nullptr;
    CU(pool_malloc(dv, &dprobe, probes.size() * sizeof(double)));
    CU(pool_malloc(dv, &dexp, probes.size() * sizeof(int)));
    CU(pool_malloc(dv, &dbad, sizeof(unsigned int)));
    unsigned int bad = 0;
    auto body = [&]() -> int {
        CU(cudaMemcpyAsync(dprobe, probes.data(), probes.size() * sizeof(double), cudaMemcpyHostToDevice, dv.stream));
        CU(cudaMemcpyAsync(dexp, expected.data(), probes.size() * sizeof(int), cudaMemcpyHostToDevice, dv.stream));
        CU(cudaMemsetAsync(dbad, 0, sizeof(unsigned int), dv.stream));
        CU(launch_validate_safe(dprobe, dexp, static_cast<int>(probes.size()), p->inv_dr, p->c0h, p->lim, p->qmax,
                                static_cast<int>(nbin), p->glo, dbad, dv.stream));
        CU(cudaMemcpyAsync(&bad, dbad, sizeof(bad), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaStreamSynchronize(dv.stream));
        return AGOFRT_OK;
    };
    int rc = body();
    if (rc == AGOFRT_OK && bad != 0) {
        p->safe_ok = false;
        p->safe2_ok = false;
        p->glo = 0;
    }
    // the two-floor form of the guess (MODE_SAFE2), on the same probes
    if (rc == AGOFRT_OK && p->safe_ok && p->safe2_ok) {
        unsigned int bad2 = 0;
        auto body2 = [&]() -> int {
            CU(cudaMemsetAsync(dbad, 0, sizeof(unsigned int), dv.stream));
            CU(launch_validate_safe2(dprobe, dexp, static_cast<int>(probes.size()), p->inv_lo, p->inv_hi, p->bias0, p->smax,
                                     static_cast<int>(nbin), p->glo, dbad, dv.stream));
            CU(cudaMemcpyAsync(&bad2, dbad, sizeof(bad2), cudaMemcpyDeviceToHost, dv.stream));
            CU(cudaStreamSynchronize(dv.stream));
            return AGOFRT_OK;
        };
        rc = body2();
        if (bad2 != 0) p->safe2_ok = false;
        if (getenv("AGOFRT_DEBUG")) fprintf(stderr, "[agofrt] plan rmin %g dr %g nbin %u: two-floor validation, %u bad probes of %zu\n", p->rmin, p->dr, nbin, bad2, probes.size());
    }
    pool_free(dv, dprobe, probes.size() * sizeof(double));
    pool_free(dv, dexp, probes.size() * sizeof(int));
    pool_free(dv, dbad, sizeof(unsigned int));
    return rc;
}

extern "C" int agofrt_plan_create(agofrt_plan **out, agofrt_traj *traj, double rmin, double rmax, unsigned nbin) try {
    if (!out || !traj) return fail(AGOFRT_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (nbin == 0 || nbin > (1u << 24)) return fail(AGOFRT_ERR_ARG, "nbin must be in [1, 2^24]");
    auto p = std::make_unique<agofrt_plan>();
    p->traj = traj;
    p->ctx = traj->ctx;
    p->ntypes = traj->ntypes;
    p->rmin = rmin;
    p->rmax = rmax;
    p->nbin = nbin;
    // reference lib/src/gofrt.cpp:27-29
    p->dr = (rmax - rmin) / nbin;
    p->rmax2 = rmax * rmax;
    p->rmin2 = rmin * rmin;

    const int nt = traj->ntypes;
    if (p->dr > 0 && std::isfinite(p->dr) && std::isfinite(rmin)) {
        build_thresholds(rmin, p->dr, nbin, p->thr_full);
    } else {
        // dr <= 0 or NaN: the reference's idx is negative, NaN or huge for every pair -> nothing counted
        p->thr_full.assign(nbin + 1, std::numeric_limits<double>::infinity());
    }
    // fold the inclusive range test rmin2 <= d2 <= rmax2 (lib/src/gofrt.cpp:104) into the table
    const double lo = std::max(p->rmin2, p->thr_full[0]);
    const double hi_excl = std::min(std::nextafter(p->rmax2, std::numeric_limits<double>::infinity()), p->thr_full[nbin]);
    p->thr.resize(nbin + 1);
    if (!(lo < hi_excl)) {
        p->empty_range = true;
        std::fill(p->thr.begin(), p->thr.end(), std::numeric_limits<double>::infinity());
        p->hlo = 1;
        p->hspan = 0;
    } else {
        for (unsigned k = 0; k <= nbin; ++k) p->thr[k] = std::min(std::max(p->thr_full[k], lo), hi_excl);
        const uint64_t lob = double_to_bits(lo), hib = double_to_bits(hi_excl) - 1;  // 0 <= lo < hi_excl
        p->hlo = static_cast<unsigned>(lob >> 32);
        p->hspan = static_cast<unsigned>(hib >> 32) - p->hlo;
    }
    p->inv_dr = static_cast<float>(1.0 / p->dr);
    p->c0 = static_cast<float>(-rmin / p->dr);
    if (!std::isfinite(p->inv_dr)) p->inv_dr = 0.0f;
    if (!std::isfinite(p->c0)) p->c0 = 0.0f;

    // Safe-zone binning (MODE_SAFE): the float guess q = sqrtf(d2)*inv_dr + c0 differs from the
    // reference's float quotient by at most q_top * (2^-22 + 2^-23 + 2^-24): truncation of d2 to
    // float (2^-24 after the root), sqrt.approx (2^-23), inv_dr and FFMA roundings (2^-24 each) and
    // the reference's own rounding to float (2^-24).  A guess farther than eps = 2*bound from
    // both edges of its bin is therefore the reference's bin; everything closer goes through the
    // exact bracket search.  The claim is then CHECKED on the device against the exact table.
    p->safe_ok = false;
    if (!p->empty_range && p->dr > 0 && std::isfinite(p->dr)) {
        const double q_top = static_cast<double>(nbin) + std::fabs(rmin) / p->dr;
        const double bound = q_top * (std::ldexp(1.0, -22) + std::ldexp(1.0, -23) + std::ldexp(1.0, -24));
        const double eps = 2.0 * bound + std::ldexp(1.0, -20);
        if (eps <= std::ldexp(1.0, -6) && q_top < std::ldexp(1.0, 21)) {
            p->c0h = static_cast<float>(-rmin / p->dr - 0.5);
            p->lim = static_cast<float>(0.5 - eps);
            p->qmax = static_cast<float>(nbin) + 0.25f;
            p->q_reach = std::ldexp(1.0, 21);
            // the smallest guess is the one of d2 = 0: rint(c0h); two spare words for its roundings
            const double lowest = std::floor(static_cast<double>(p->c0h));
            const double guards = lowest < 0 ? -lowest + 2 : 2;
            if (guards <= 4096) {
                p->glo = static_cast<int>(guards);
                p->safe_ok = true;
                // MODE_SAFE2 wants c0 = -rmin/dr to be an integer as a float (rmin = 0, or a multiple of dr): then
                // 1.5*2^23 + c0 + row is an integer and one round-down FFMA yields the histogram word
                if (p->c0 == std::floor(p->c0) && std::fabs(p->c0) < 4000.0f) {
                    const double inv = 1.0 / p->dr, delta = std::ldexp(1.0, -20);
                    p->inv_lo = static_cast<float>(inv * (1.0 - delta));
                    p->inv_hi = static_cast<float>(inv * (1.0 + delta));
                    p->bias0 = 12582912.0f + p->c0;
                    // clamp of sqrt(d2): the middle of the guard bin that ends every row (bin coordinate nbin + 0.5)
                    p->smax = static_cast<float>((static_cast<double>(nbin) + 0.5) * p->dr + rmin);
                    p->safe2_ok = p->inv_lo < p->inv_hi && std::isfinite(p->inv_hi) && std::isfinite(p->smax);
                }
            }
        }
    }
    // shared-memory budget; the guard bins are given up before the plan is refused
    for (const Dev &d : traj->ctx->devs) {
        if (pair_kernel_smem_bytes(nt, static_cast<int>(nbin), static_cast<int>(nbin), p->glo, true) > d.smem_optin && p->glo > 0) {
            p->glo = 0;
            p->safe_ok = false;
            p->safe2_ok = false;
        }
        const size_t smem = pair_kernel_smem_bytes(nt, static_cast<int>(nbin), static_cast<int>(nbin), p->glo, true);
        if (smem > d.smem_optin)
            return fail(AGOFRT_ERR_TOO_LARGE,
                        "histogram of %d type-pair rows x %u bins needs %zu bytes of shared memory (> %zu)",
                        nt * (nt + 1), nbin, smem, d.smem_optin);
    }

    p->dev.resize(traj->ctx->devs.size());
    for (size_t i = 0; i < p->dev.size(); ++i) {
        CU(cudaSetDevice(traj->ctx->devs[i].id));
        PlanDev &d = p->dev[i];
        Dev &dv = traj->ctx->devs[i];
        CU(pool_malloc(dv, &d.thr, (nbin + 1) * sizeof(double)));
        CU(pool_malloc(dv, &d.thr_full, (nbin + 1) * sizeof(double)));
        CU(pool_malloc(dv, &d.counter, sizeof(unsigned int)));
        CU(pool_malloc(dv, &d.edges, sizeof(unsigned long long)));
        CU(cudaMemcpy(d.thr, p->thr.data(), (nbin + 1) * sizeof(double), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d.thr_full, p->thr_full.data(), (nbin + 1) * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (p->safe_ok) {
        const int rcv = validate_safe_zone(p.get());
        if (rcv != AGOFRT_OK) return rcv;
    }
    CU(host_pool_alloc(traj->ctx, reinterpret_cast<void **>(&p->host_edges), sizeof(unsigned long long) * p->dev.size()));
    CU(host_pool_alloc(traj->ctx, reinterpret_cast<void **>(&p->host_flags), sizeof(unsigned int) * p->dev.size()));
    *out = p.release();
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_plan_retarget(agofrt_plan *p, agofrt_traj *traj) try {
    if (!p || !traj) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (traj == p->traj) return AGOFRT_OK;
    if (traj->ctx != p->ctx || traj->ntypes != p->ntypes)
        return fail(AGOFRT_ERR_ARG, "the new window belongs to another context or has another number of types");
    p->traj = traj;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_plan_destroy(agofrt_plan *p) try {
    if (!p) return AGOFRT_OK;
    const bool debug = getenv("AGOFRT_DEBUG") != nullptr;
    if (debug) {   // is anything still running on the devices?  (how long a device-wide wait takes right now)
        const auto s0 = std::chrono::steady_clock::now();
        for (size_t i = 0; i < p->dev.size(); ++i) {
            cudaSetDevice(p->ctx->devs[i].id);
            cudaDeviceSynchronize();
        }
        fprintf(stderr, "[agofrt] device-wide wait before the plan is released: %.1f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s0).count());
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < p->dev.size(); ++i) {
        cudaSetDevice(p->ctx->devs[i].id);
        PlanDev &d = p->dev[i];
        Dev &dv = p->ctx->devs[i];
        pool_free(dv, d.thr, (p->nbin + 1) * sizeof(double));
        pool_free(dv, d.thr_full, (p->nbin + 1) * sizeof(double));
        pool_free(dv, d.ghist, d.ghist_len * sizeof(unsigned long long));
        pool_free(dv, d.jobs, d.jobs_cap * sizeof(Job));
        pool_free(dv, d.counter, sizeof(unsigned int));
        pool_free(dv, d.edges, sizeof(unsigned long long));
        spare_give(dv, d.batch, d.batch_len * sizeof(unsigned long long));
    }
    const auto t1 = std::chrono::steady_clock::now();
    host_pool_free(p->ctx, p->host_counts, p->host_counts_len * sizeof(unsigned long long));
    host_pool_free(p->ctx, p->host_edges, sizeof(unsigned long long) * p->dev.size());
    host_pool_free(p->ctx, p->host_flags, sizeof(unsigned int) * p->dev.size());
    delete p;
    if (debug)
        fprintf(stderr, "[agofrt] plan released in %.1f ms (device buffers %.1f)\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(),
                std::chrono::duration<double, std::milli>(t1 - t0).count());
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_plan_info(const agofrt_plan *p, int *safe_zone_ok, int *two_floor_ok, int *guard_bins) try {
    if (!p) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (safe_zone_ok) *safe_zone_ok = p->safe_ok ? 1 : 0;
    if (two_floor_ok) *two_floor_ok = p->safe2_ok ? 1 : 0;
    if (guard_bins) *guard_bins = p->glo;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_plan_thresholds(const agofrt_plan *p, double *thresholds) try {
    if (!p || !thresholds) return fail(AGOFRT_ERR_ARG, "NULL argument");
    memcpy(thresholds, p->thr_full.data(), (p->nbin + 1) * sizeof(double));
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// ---------------------------------------------------------------------------------------------
// one block
// ---------------------------------------------------------------------------------------------
// Is one image per dimension PROVABLY enough for every pair of this (frame fi, frame fj) job?
// |xi-xj| is bounded by the coordinate ranges of the two frames (the rounded difference is
// monotone in its operands), each tilt correction adds at most |tilt|, and a value with
// |d| <= 3*l_half lands in [-l_half, l_half] after one +-2*l_half.  2.99 leaves room for the
// roundings of the bound itself.
static bool job_is_single_pass(const agofrt_traj *t, size_t fi, size_t fj) {
    const double *b = &t->box6[fi * 6];
    const double *bi = &t->bounds[fi * 6], *bj = &t->bounds[fj * 6];
    double D[3];
    for (int c = 0; c < 3; ++c) {
        const double a = std::fabs(bi[3 + c] - bj[c]);  // max_i - min_j
        const double d = std::fabs(bi[c] - bj[3 + c]);  // min_i - max_j
        D[c] = std::max(a, d);
        if (!std::isfinite(D[c])) return false;
    }
    const double xy = std::fabs(b[3]), xz = std::fabs(b[4]), yz = std::fabs(b[5]);
    const double m = 2.99;
    if (!(D[2] <= m * b[2])) return false;
    if (!(D[1] + yz <= m * b[1])) return false;
    if (!(D[0] + xz + xy <= m * b[0])) return false;
    return true;
}

// The same proof for EVERY pair of frames in [f0, f1] at once, from the bounds of the whole range: the rounded
// difference is monotone in its operands, so the per-job D of job_is_single_pass never exceeds max - min over
// the range, and a range that passes here would pass job by job (the converse need not hold: the caller then
// falls back to the per-job test).  Turns the classification of a block of 10^5 small jobs from one test per
// job into one test per frame.
static bool range_is_single_pass(const agofrt_traj *t, size_t f0, size_t f1) {
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t f = f0; f <= f1; ++f)
        for (int c = 0; c < 3; ++c) {
            lo[c] = std::min(lo[c], t->bounds[f * 6 + c]);
            hi[c] = std::max(hi[c], t->bounds[f * 6 + 3 + c]);
        }
    double D[3];
    for (int c = 0; c < 3; ++c) {
        D[c] = std::fabs(hi[c] - lo[c]);
        if (!std::isfinite(D[c])) return false;
    }
    const double m = 2.99;
    for (size_t f = f0; f <= f1; ++f) {
        const double *b = &t->box6[f * 6];
        const double xy = std::fabs(b[3]), xz = std::fabs(b[4]), yz = std::fabs(b[5]);
        if (!(D[2] <= m * b[2])) return false;
        if (!(D[1] + yz <= m * b[1])) return false;
        if (!(D[0] + xz + xy <= m * b[0])) return false;
    }
    return true;
}

// A whole block on ONE device, enqueued and not waited for (agofrt_blocks): where it runs and where its counts go
struct BlockTarget {
    int dev;                    // local device
    unsigned long long *ghist;  // [len] on that device
    bool first;                 // first block of the batch on this device: reset the error flag
    // the part [part_a, part_b) / part_den of the block's work units that this device takes (agofrt_blocks shares out
    // blocks AND parts of blocks when the blocks do not divide among the devices); part_den == 0: all of it
    uint32_t part_a = 0, part_b = 0, part_den = 0;
};
constexpr int kNotBatchable = 1;   // block_impl with a target: the block has no regular job list (caller falls back)

static int block_impl(agofrt_plan *p, size_t primo, unsigned ntimesteps, unsigned leff, unsigned skip, unsigned every,
                      unsigned options, uint64_t *counts_out, uint64_t *edge_pairs_out, agofrt_stats *stats,
                      const BlockTarget *tg) {
    if (!p) return fail(AGOFRT_ERR_ARG, "plan is NULL");
    agofrt_traj *t = p->traj;
    agofrt_ctx *ctx = t->ctx;
    if (skip == 0) skip = 1;    // reference lib/include/calculatemultithread.h:44-45
    if (every == 0) every = 1;
    const int nt = t->ntypes;
    const size_t rowlen = static_cast<size_t>(nt) * (nt + 1) * p->nbin;
    const size_t len = static_cast<size_t>(leff) * rowlen;
    const bool on_device = (options & AGOFRT_OPT_ON_DEVICE) != 0;
    if (len > 0 && !counts_out && !on_device && !tg) return fail(AGOFRT_ERR_ARG, "counts_out is NULL");
    p->last_valid = false;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (edge_pairs_out) *edge_pairs_out = 0;
    const bool want_edges = (options & AGOFRT_OPT_EDGES) != 0 || edge_pairs_out != nullptr;

    int rc = ensure_local_comm(ctx);
    if (rc != AGOFRT_OK) return rc;
    const int nloc = static_cast<int>(ctx->devs.size());
    // (a target: the whole block on one device, no sharding)
    const int world = tg ? 1 : (ctx->world > 0 ? ctx->world : nloc);
    const int first_rank = tg ? 0 : (ctx->world > 0 ? ctx->first_rank : 0);

    // ---- the (lag, origin) jobs, in the reference's loop order (calculatemultithread.h:114-115) ----
    std::vector<Job> jobs_fast, jobs_gen;
    uint64_t njobs = 0, n_fast = 0, n_gen = 0;
    bool all_fast = false, implicit = false;
    if (ntimesteps > 0 && leff > 0) {
        // the last frame the loops really touch: last origin + last lag
        const size_t last = primo + static_cast<size_t>((ntimesteps - 1) / skip) * skip +
                            static_cast<size_t>((leff - 1) / every) * every;
        if (primo < t->first_frame || last >= t->first_frame + t->nframes)
            return fail(AGOFRT_ERR_WINDOW,
                        "block needs frames [%zu,%zu] but the device window holds [%zu,%zu)", primo, last,
                        t->first_frame, t->first_frame + t->nframes);
        if (t->natoms > 0 && (t->bad_box || t->has_inf))
            return fail(AGOFRT_ERR_NONFINITE,
                        "the window holds an infinite coordinate or a non-positive / non-finite box edge "
                        "(the reference's minimum image would not terminate)");
        const bool may_fast = !(options & AGOFRT_OPT_FORCE_GENERAL);
        all_fast = may_fast && range_is_single_pass(t, primo - t->first_frame, last - t->first_frame);
        const size_t expect = static_cast<size_t>((leff + every - 1) / every) * ((ntimesteps + skip - 1) / skip);
        if (expect >= 0xF0000000ull)   // (the same limit as on the work units below, before any memory is asked for)
            return fail(AGOFRT_ERR_ARG, "too many work units in one block (%zu (lag, origin) jobs)", expect);
        // When every job takes the same kernel (the whole frame range proven single-pass, or the general kernel
        // forced) the list is regular -- job k = (lag k / origins, origin k % origins) -- and the kernels derive it
        // themselves: nothing is built or uploaded (C1: 142 884 jobs per block).  Otherwise: an explicit list per kernel.
        implicit = (all_fast || !may_fast) && !(options & AGOFRT_OPT_EXPLICIT_JOBS);
        if (tg && !implicit) return kNotBatchable;
        if (implicit) {
            njobs = expect;
            (all_fast ? n_fast : n_gen) = expect;
        } else {
            (may_fast ? jobs_fast : jobs_gen).reserve(expect);
            for (unsigned tl = 0; tl < leff; tl += every)
                for (unsigned im = 0; im < ntimesteps; im += skip) {
                    const size_t fi = primo + im - t->first_frame;
                    Job j{static_cast<int>(fi), static_cast<int>(fi + tl), static_cast<int>(tl)};
                    const bool fast = all_fast || (may_fast && job_is_single_pass(t, fi, fi + tl));
                    (fast ? jobs_fast : jobs_gen).push_back(j);
                    ++njobs;
                }
            n_fast = jobs_fast.size();
            n_gen = jobs_gen.size();
        }
    }
    const unsigned norig = skip ? (ntimesteps + skip - 1) / skip : 0;
    const uint64_t n2 = static_cast<uint64_t>(t->natoms) * t->natoms;
    const bool nothing = njobs == 0 || t->natoms == 0 || p->empty_range || len == 0;

    // ---- work units ----
    const int n_itiles = std::max(1, (t->npad + kTileI - 1) / kTileI);
    int total_ctas = 0;
    for (const Dev &d : ctx->devs) total_ctas += kMinBlocks * d.sm_count;
    total_ctas = total_ctas / nloc * world;
    int n_jchunks = 1;
    {
        const int max_chunks = std::max(1, (t->npad + 4 * kTileJ - 1) / (4 * kTileJ));
        // enough units that the last one a CTA draws is a small part of its share: 64 per CTA (with 16, the 8-GPU run of
        // the default bench step had 17 units of 17 ms per CTA: a tail of up to 6 %)
        const uint64_t want = 64ull * total_ctas;
        const uint64_t base = std::max<uint64_t>(1, njobs * n_itiles);
        n_jchunks = static_cast<int>(std::min<uint64_t>(max_chunks, (want + base - 1) / base));
        n_jchunks = std::max(1, n_jchunks);
    }
    int jchunk = (t->npad + n_jchunks - 1) / n_jchunks;
    jchunk = std::max(kTileJ, (jchunk + kTileJ - 1) / kTileJ * kTileJ);
    n_jchunks = std::max(1, (t->npad + jchunk - 1) / jchunk);
    const uint64_t per_job = static_cast<uint64_t>(n_itiles) * n_jchunks;
    if (njobs * per_job >= 0xF0000000ull) return fail(AGOFRT_ERR_ARG, "too many work units in one block (%llu)",
                                                      static_cast<unsigned long long>(njobs * per_job));

    // ---- small systems: a contiguous range of jobs per CTA, batches of jobs per warp (pair_small_kernel) ----
    // Default up to kSmallDefault slots; above (up to kSmallMax) opt-in.
    const bool small = t->npad > 0 && !(options & AGOFRT_OPT_NO_SMALL) &&
                       (t->npad <= kSmallDefault || ((options & AGOFRT_OPT_SMALL) && t->npad <= kSmallMax));
    // A warp of pair_small_kernel works on `small_jb` jobs at a time, their i slots packed over its lanes (64 per
    // round): the batch size that wastes the fewest lanes of the last round, among those whose j frames fit the
    // warp's slice of the stage area with two CTAs per SM, and that leave a batch or more to every warp.
    int small_jb = 1, small_js = std::max(t->npad, 2), small_nb = 1;
    bool aggregate = p->nbin * static_cast<unsigned>(nt * (nt + 1) / 2) < 64;  // few counters: collisions are the rule
    if (options & AGOFRT_OPT_AGGREGATE) aggregate = true;
    if (options & AGOFRT_OPT_NO_AGGREGATE) aggregate = false;
    if (small) {
        // slots between the j frames of a batch: the threads of a warp read the frames of up to three consecutive jobs
        // with 16-byte loads -- keep those on different banks (4-byte words: 2 * js mod 32 away from 0)
        const int w = (2 * small_js) % 32;
        if (w < 4 || w > 28) small_js += 2;
        const size_t per_cta = (ctx->devs[0].smem_per_sm - 2048) / kMinBlocks;   // 1 KiB per CTA belongs to the system
        const size_t per_job = static_cast<size_t>(kThreads / 32) * 3 * small_js * sizeof(double);
        uint64_t total_warps = 0;
        for (const Dev &d : ctx->devs) total_warps += static_cast<uint64_t>(kMinBlocks) * d.sm_count * (kThreads / 32);
        int by_jobs = static_cast<int>(std::min<uint64_t>(8, std::max<uint64_t>(1, njobs / std::max<uint64_t>(total_warps, 1))));
        if (const char *e = getenv("AGOFRT_SMALL_BATCH")) by_jobs = std::max(1, std::min(8, atoi(e)));   // (tests: large batches on few jobs)
        // the best batch (fewest empty lanes in its last round; ties: the larger) that fits beside `nb` histograms
        auto best_batch = [&](int nb, int *jb_out) {
            const size_t fixed = pair_small_kernel_smem_bytes(nt, static_cast<int>(p->nbin), static_cast<int>(p->nbin), p->glo,
                                                              want_edges, 0, 0, nb);
            const int fit = per_cta > fixed ? static_cast<int>((per_cta - fixed) / per_job) : 0;
            const int jb_max = std::max(1, std::min(std::min(8, fit), by_jobs));
            double best = -1;
            for (int jb = 1; jb <= jb_max; ++jb) {
                const int slots = jb * t->npad, rounds = (slots + 32 * kIPT - 1) / (32 * kIPT);
                const double eff = static_cast<double>(slots) / (rounds * 32 * kIPT);
                if (eff >= best - 1e-9) {
                    best = std::max(best, eff);
                    *jb_out = jb;
                }
            }
            return fit >= 1 ? best : 0.0;
        };
        const double eff1 = best_batch(1, &small_jb);
        // histograms a CTA keeps at a time: the lags its share of the jobs touches (one merge at the end instead of one
        // per lag), as long as they fit without costing the batch its shape.  The warp-aggregated binning keys on the
        // bin alone: one histogram.
        const uint64_t jobs_per_cta = std::max<uint64_t>(1, njobs / std::max(1, total_ctas));
        int want_nb = static_cast<int>(std::min<uint64_t>(8, jobs_per_cta / std::max(1u, norig) + 2));
        if (const char *e = getenv("AGOFRT_SMALL_HISTS")) want_nb = std::max(1, std::min(8, atoi(e)));
        if (aggregate) want_nb = 1;
        for (int nb = want_nb; nb >= 2; --nb) {
            int jb = 1;
            if (best_batch(nb, &jb) >= eff1 - 1e-9 && jb >= small_jb) {
                small_nb = nb;
                small_jb = jb;
                break;
            }
        }
    }

    const bool tri = t->stride == 9;
    // safe-zone binning needs the bin coordinate of every reachable distance to stay below 2^21
    bool use_safe = p->safe_ok && !(options & AGOFRT_OPT_NO_SAFE) && !t->has_nan;
    if (use_safe) {
        double reach = 0;
        for (size_t f = 0; f < t->nframes; ++f) {
            const double *b = &t->box6[f * 6];
            const double ex = 3 * b[0] + std::fabs(b[3]) + std::fabs(b[4]), ey = 3 * b[1] + std::fabs(b[5]), ez = 3 * b[2];
            reach = std::max(reach, std::sqrt(ex * ex + ey * ey + ez * ez));
        }
        if (!(reach / p->dr < p->q_reach)) use_safe = false;
    }
    // one box for the whole window (NVT / NVE)?  Then it travels as a kernel parameter.
    bool same_box = t->nframes > 0;
    for (size_t f = 1; f < t->nframes && same_box; ++f)
        same_box = memcmp(&t->box6[f * 6], &t->box6[0], 6 * sizeof(double)) == 0;
    // dense or sparse?  The share of pairs within rmax, from the sphere / cell volume ratio of the
    // smallest cell of the window; above ~15% nearly every group of 8 pairs holds an in-range pair.
    bool dense = false;
    {
        double vmin = std::numeric_limits<double>::infinity();
        for (size_t f = 0; f < t->nframes; ++f) {
            const double *b = &t->box6[f * 6];
            vmin = std::min(vmin, 8.0 * b[0] * b[1] * b[2]);
        }
        const double rm = std::max(p->rmax, 0.0), r0 = std::max(p->rmin, 0.0);
        const double share = 4.18879020478639 * (rm * rm * rm - r0 * r0 * r0) / vmin;
        dense = share > 0.15;   // the dense kernel only drops the group filter
        if (options & AGOFRT_OPT_DENSE) dense = true;
        if (options & AGOFRT_OPT_SPARSE) dense = false;
    }
    const size_t smem = pair_kernel_smem_bytes(nt, static_cast<int>(p->nbin), static_cast<int>(p->nbin), p->glo, want_edges);
    const size_t smem_small = small ? pair_small_kernel_smem_bytes(nt, static_cast<int>(p->nbin), static_cast<int>(p->nbin), p->glo,
                                                                   want_edges, small_jb, small_js, small_nb)
                                    : 0;
    if (smem_small > ctx->devs[0].smem_optin)
        return fail(AGOFRT_ERR_ARG, "the small-system kernel needs %zu bytes of shared memory (limit %zu)", smem_small,
                    ctx->devs[0].smem_optin);
    // dense windows: the two-floor form of the safe-zone binning (MODE_SAFE2) where the plan allows it
    const int nhi = static_cast<int>(p->nbin);
    // (opt-in: measured on C2 it is 6 % slower than the clamped form although it issues two instructions per pair
    // fewer -- DESIGN.md section 4, profiles/r2f_variants.txt)
    const bool use_safe2 = use_safe && dense && p->safe2_ok && !aggregate && !want_edges && !small && (options & AGOFRT_OPT_SAFE2);

    // ---- pinned read-back buffer ----
    if (!tg && len > p->host_counts_len) {
        host_pool_free(ctx, p->host_counts, p->host_counts_len * sizeof(unsigned long long));
        p->host_counts = nullptr;
        p->host_counts_len = 0;
        CU(host_pool_alloc(ctx, reinterpret_cast<void **>(&p->host_counts), len * sizeof(unsigned long long)));
        p->host_counts_len = len;
    }

    unsigned launches = 0, modes_used = 0;
    uint64_t my_pairs = 0;
    // ---- enqueue on every local device ----
    for (int i = 0; i < nloc; ++i) {
        if (tg && i != tg->dev) continue;
        Dev &dv = ctx->devs[i];
        TrajDev &td = t->dev[i];
        PlanDev &pd = p->dev[i];
        CU(cudaSetDevice(dv.id));
        unsigned long long *const ghist = tg ? tg->ghist : nullptr;
        if (!tg && len > pd.ghist_len) {
            pool_free(dv, pd.ghist, pd.ghist_len * sizeof(unsigned long long));
            pd.ghist = nullptr;
            pd.ghist_len = 0;
            CU(pool_malloc(dv, &pd.ghist, len * sizeof(unsigned long long)));
            pd.ghist_len = len;
        }
        const size_t njall = jobs_fast.size() + jobs_gen.size();   // (0 with implicit jobs)
        if (njall > pd.jobs_cap) {
            pool_free(dv, pd.jobs, pd.jobs_cap * sizeof(Job));
            pd.jobs = nullptr;
            pd.jobs_cap = 0;
            CU(pool_malloc(dv, &pd.jobs, njall * sizeof(Job)));
            pd.jobs_cap = njall;
        }
        if (!tg) CU(cudaEventRecord(dv.ev_begin, dv.stream));
        if (len > 0) CU(cudaMemsetAsync(tg ? ghist : pd.ghist, 0, len * sizeof(unsigned long long), dv.stream));
        CU(cudaMemsetAsync(pd.edges, 0, sizeof(unsigned long long), dv.stream));
        if (!tg || tg->first) CU(cudaMemsetAsync(td.flags + 1, 0, sizeof(unsigned int), dv.stream));
        if (!jobs_fast.empty())
            CU(cudaMemcpyAsync(pd.jobs, jobs_fast.data(), jobs_fast.size() * sizeof(Job), cudaMemcpyHostToDevice, dv.stream));
        if (!jobs_gen.empty())
            CU(cudaMemcpyAsync(pd.jobs + jobs_fast.size(), jobs_gen.data(), jobs_gen.size() * sizeof(Job),
                               cudaMemcpyHostToDevice, dv.stream));
        if (!tg) CU(cudaEventRecord(dv.ev_k0, dv.stream));
        if (!nothing) {
            const int g = tg ? 0 : first_rank + i;
            for (int pass = 0; pass < 2; ++pass) {
                const uint64_t nlist = pass == 0 ? n_fast : n_gen;
                if (nlist == 0) continue;
                const uint64_t units = small ? nlist : nlist * per_job;   // the small-system kernel shares out jobs
                uint64_t ub = 0, ue = 0;
                agofrt_shard_range(units, g, world, &ub, &ue);
                if (tg && tg->part_den) {
                    ub = static_cast<uint64_t>(static_cast<unsigned __int128>(units) * tg->part_a / tg->part_den);
                    ue = static_cast<uint64_t>(static_cast<unsigned __int128>(units) * tg->part_b / tg->part_den);
                }
                if (ue <= ub) continue;
                PairParams pp;
                pp.pos = td.pos;
                pp.box = td.box6;
                pp.type_pad = td.type_pad;
                pp.type_start = td.type_start;
                pp.jobs = pd.jobs + (pass == 0 ? 0 : jobs_fast.size());
                pp.imp = implicit ? 1 : 0;
                pp.imp_f0 = static_cast<int>(primo - t->first_frame);
                pp.imp_norig = static_cast<int>(norig);
                pp.imp_skip = static_cast<int>(skip);
                pp.imp_every = static_cast<int>(every);
                pp.ghist = tg ? ghist : pd.ghist;
                pp.edges = pd.edges;
                pp.counter = pd.counter;
                pp.error_flag = td.flags + 1;
                pp.thr = pd.thr;
                pp.thr_full = pd.thr_full;
                pp.rmin2 = p->rmin2;
                pp.rmax2 = p->rmax2;
                pp.unit_begin = static_cast<unsigned>(ub);
                pp.unit_end = static_cast<unsigned>(ue);
                pp.npad = t->npad;
                pp.ntypes = nt;
                pp.nbin = static_cast<int>(p->nbin);
                pp.n_itiles = small ? small_jb : n_itiles;
                pp.n_jchunks = small ? small_js : n_jchunks;
                pp.jchunk = jchunk;
                pp.small_nb = small_nb;
                pp.small_w16 = 16;
                if (small) {
                    // the jobs of lag 0 lead the list (lag-major order)
                    uint64_t lag0 = 0;
                    if (implicit) {
                        lag0 = norig;
                    } else {
                        const std::vector<Job> &list = pass == 0 ? jobs_fast : jobs_gen;
                        while (lag0 < list.size() && list[lag0].tout == 0) ++lag0;
                    }
                    pp.jchunk = static_cast<int>(lag0);
                }
                pp.inv_dr = p->inv_dr;
                pp.c0 = p->c0;
                pp.hlo = p->hlo;
                pp.hspan = p->hspan;
                pp.c0h = p->c0h;
                pp.lim = p->lim;
                pp.qmax = p->qmax;
                pp.glo = p->glo;
                pp.hhi = p->hlo + p->hspan;
                pp.nhi = nhi;
                pp.inv_lo = p->inv_lo;
                pp.inv_hi = p->inv_hi;
                pp.bias0 = p->bias0;
                pp.smax = p->smax;
                pp.skew = (options & AGOFRT_OPT_SKEW) ? 1 : 0;
                int mode = kModeThr;
                if (want_edges)
                    mode = kModeEdges;
                else if (aggregate)
                    mode = kModeAgg;
                else if (pass == 0 && use_safe)
                    mode = use_safe2 ? kModeSafe2 : (dense ? kModeSafeDense : kModeSafe);
                // lag 0 through the safe-zone binning: the self pairs are left out of the main pass, three instructions per
                // pair (pair_small_kernel)
                if (mode == kModeSafe || mode == kModeSafeDense) pp.small_w16 = 18;
                const bool ubox = same_box && pass == 0 && !(options & AGOFRT_OPT_NO_UBOX) &&
                                  (mode == kModeThr || mode == kModeSafe || mode == kModeSafeDense || mode == kModeSafe2);
                if (ubox) {
                    const double *b = &t->box6[0];
                    for (int k = 0; k < 6; ++k) pp.ubox[k] = b[k];
                    for (int k = 0; k < 3; ++k) pp.ubox[6 + k] = -2.0 * b[k];
                    for (int k = 0; k < 3; ++k) pp.ubox[9 + k] = -b[3 + k];
                } else {
                    for (int k = 0; k < 12; ++k) pp.ubox[k] = 0.0;
                }
                const int variant = (tri ? 1 : 0) | (pass == 0 ? 2 : 0) | (mode << 2) | (ubox ? 32 : 0) | (small ? 64 : 0);
                // small systems: a CTA per batch round of its eight warps, at most
                const uint64_t want_ctas = small ? (ue - ub + (kThreads / 32) * small_jb - 1) / ((kThreads / 32) * small_jb) : ue - ub;
                const int grid = static_cast<int>(std::min<uint64_t>(want_ctas, static_cast<uint64_t>(kMinBlocks) * dv.sm_count));
                CU(cudaMemsetAsync(pd.counter, 0, sizeof(unsigned int), dv.stream));
                CU(launch_pair_kernel(variant, grid, small ? smem_small : smem, dv.stream, pp));
                ++launches;
                modes_used |= 1u << mode;
                if (small) {
                    modes_used |= 1u << 8;
                    my_pairs += (ue - ub) * n2;
                } else {
                    // pair evaluations of this shard, counted on real atoms: units are equal-sized
                    my_pairs += static_cast<uint64_t>(static_cast<double>(ue - ub) / static_cast<double>(per_job) * n2 + 0.5);
                }
            }
        }
        if (!tg) CU(cudaEventRecord(dv.ev_k1, dv.stream));
    }
    if (tg) {
        // enqueued; the caller waits, exchanges and checks the flags for the whole batch
        if (stats) {
            stats->pair_evals = nothing ? 0 : my_pairs;
            stats->pair_evals_total = njobs * n2;
            stats->jobs = njobs;
            stats->jobs_fast = n_fast;
            stats->launches = launches;
            stats->kernel_modes = modes_used;
        }
        return AGOFRT_OK;
    }
    // ---- combine: one all-reduce of the integer histograms (replaces mp.h:35-41) ----
    if (ctx->comm_ready && world > 1 && len > 0) {
        NcclApi &api = nccl_api();
        NC(api.GroupStart());
        for (int i = 0; i < nloc; ++i) {
            Dev &dv = ctx->devs[i];
            PlanDev &pd = p->dev[i];
            NC(api.AllReduce(pd.ghist, pd.ghist, len, ncclUint64, ncclSum, dv.comm, dv.stream));
            if (want_edges) NC(api.AllReduce(pd.edges, pd.edges, 1, ncclUint64, ncclSum, dv.comm, dv.stream));
        }
        NC(api.GroupEnd());
    }
    // ---- read back ----
    {
        Dev &dv = ctx->devs[0];
        PlanDev &pd = p->dev[0];
        CU(cudaSetDevice(dv.id));
        if (len > 0 && counts_out)
            CU(cudaMemcpyAsync(p->host_counts, pd.ghist, len * sizeof(unsigned long long), cudaMemcpyDeviceToHost, dv.stream));
    }
    for (int i = 0; i < nloc; ++i) {
        Dev &dv = ctx->devs[i];
        CU(cudaSetDevice(dv.id));
        CU(cudaMemcpyAsync(&p->host_edges[i], p->dev[i].edges, sizeof(unsigned long long), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaMemcpyAsync(&p->host_flags[i], t->dev[i].flags + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaEventRecord(dv.ev_end, dv.stream));
    }
    double kernel_ms = 0, total_ms = 0;
    for (int i = 0; i < nloc; ++i) {
        Dev &dv = ctx->devs[i];
        CU(cudaSetDevice(dv.id));
        CU(cudaStreamSynchronize(dv.stream));
        float a = 0, b = 0;
        CU(cudaEventElapsedTime(&a, dv.ev_k0, dv.ev_k1));
        CU(cudaEventElapsedTime(&b, dv.ev_begin, dv.ev_end));
        kernel_ms = std::max<double>(kernel_ms, a);
        total_ms = std::max<double>(total_ms, b);
        if (p->host_flags[i])
            return fail(AGOFRT_ERR_NONFINITE, "minimum image did not converge within %d images", kWrapCap);
    }
    // a sharded context without communicator and several local devices: partial counts are summed
    // here only because they are all in this process (integer sum, order-free)
    // the counts in dev[0].ghist are the whole job's when one device did it all or the all-reduce ran
    p->last_len = len;
    p->last_valid = nloc == 1 ? (world == 1 || (ctx->comm_ready && world > 1)) : ctx->comm_ready;
    if (on_device && !p->last_valid)
        return fail(AGOFRT_ERR_ARG, "AGOFRT_OPT_ON_DEVICE needs the complete counts on the device: a sharded context "
                                    "without communicator only holds partial counts");
    if (len > 0 && counts_out) {
        memcpy(counts_out, p->host_counts, len * sizeof(uint64_t));
        if (!ctx->comm_ready && nloc > 1) {
            for (int i = 1; i < nloc; ++i) {
                Dev &dv = ctx->devs[i];
                CU(cudaSetDevice(dv.id));
                CU(cudaMemcpy(p->host_counts, p->dev[i].ghist, len * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                for (size_t k = 0; k < len; ++k) counts_out[k] += p->host_counts[k];
            }
        }
    }
    if (edge_pairs_out) {
        if (ctx->comm_ready && world > 1) {
            *edge_pairs_out = p->host_edges[0];
        } else {
            uint64_t s = 0;
            for (int i = 0; i < nloc; ++i) s += p->host_edges[i];
            *edge_pairs_out = s;
        }
    }
    if (stats) {
        stats->kernel_ms = kernel_ms;
        stats->total_ms = total_ms;
        stats->pair_evals = nothing ? 0 : my_pairs;
        stats->pair_evals_total = njobs * n2;
        stats->jobs = njobs;
        stats->jobs_fast = n_fast;
        stats->launches = launches;
        stats->ndev_local = static_cast<uint32_t>(nloc);
        stats->world = static_cast<uint32_t>(world);
        stats->kernel_modes = modes_used;
    }
    return AGOFRT_OK;
}

extern "C" int agofrt_block(agofrt_plan *p, size_t primo, unsigned ntimesteps, unsigned leff, unsigned skip,
                            unsigned every, unsigned options, uint64_t *counts_out, uint64_t *edge_pairs_out,
                            agofrt_stats *stats) try {
    return block_impl(p, primo, ntimesteps, leff, skip, every, options, counts_out, edge_pairs_out, stats, nullptr);
} catch (...) {
    return on_exception();
}

// Many small blocks: WHOLE blocks are dealt to the devices (block b to device b mod world), each block a launch of
// its own on its device's stream, nothing waited for in between; then every device receives every block (grouped
// ncclBroadcast over NVLink), so that the Welford update can run, in block order, wherever it likes.  Replaces the
// round-robin of blocks over MPI ranks of the reference (lib/include/blockaverage.h:146-186).
extern "C" int agofrt_blocks(agofrt_plan *p, size_t primo0, size_t stride, unsigned nblocks, unsigned ntimesteps, unsigned leff,
                             unsigned skip, unsigned every, unsigned options, agofrt_stats *stats) try {
    if (!p) return fail(AGOFRT_ERR_ARG, "plan is NULL");
    if (options & (AGOFRT_OPT_EDGES)) return fail(AGOFRT_ERR_ARG, "agofrt_blocks does not count edge pairs (use agofrt_block)");
    agofrt_traj *t = p->traj;
    agofrt_ctx *ctx = t->ctx;
    if (stats) memset(stats, 0, sizeof(*stats));
    p->batch_blocks = 0;
    p->last_valid = false;
    int rc = ensure_local_comm(ctx);
    if (rc != AGOFRT_OK) return rc;
    const int nloc = static_cast<int>(ctx->devs.size());
    const int world = ctx->world > 0 ? ctx->world : nloc;
    const int first_rank = ctx->world > 0 ? ctx->first_rank : 0;
    if (world > 1 && !ctx->comm_ready) return fail(AGOFRT_ERR_ARG, "agofrt_blocks on a sharded context needs a communicator");
    const int nt = t->ntypes;
    const size_t len = static_cast<size_t>(leff) * nt * (nt + 1) * p->nbin;
    if (nblocks == 0 || len == 0) return AGOFRT_OK;
    const bool debug = getenv("AGOFRT_DEBUG") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since0 = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    double ms_alloc = 0, ms_enqueue = 0;
    if (static_cast<double>(nblocks) * static_cast<double>(len) * 8.0 > 2.0e9)
        return fail(AGOFRT_ERR_TOO_LARGE, "%u blocks of %zu counters do not fit the batch buffer (use agofrt_block)", nblocks, len);
    for (int i = 0; i < nloc; ++i) {
        PlanDev &pd = p->dev[i];
        CU(cudaSetDevice(ctx->devs[i].id));
        if (static_cast<size_t>(nblocks) * len > pd.batch_len) {
            spare_give(ctx->devs[i], pd.batch, pd.batch_len * sizeof(unsigned long long));
            pd.batch = nullptr;
            pd.batch_len = 0;
            const size_t want = static_cast<size_t>(nblocks) * len * sizeof(unsigned long long);
            pd.batch = static_cast<unsigned long long *>(spare_take(ctx->devs[i], want));
            if (!pd.batch) CU(cudaMalloc(&pd.batch, want));
            pd.batch_len = static_cast<size_t>(nblocks) * len;
        }
        CU(cudaEventRecord(ctx->devs[i].ev_begin, ctx->devs[i].stream));
    }
    ms_alloc = since0();
    agofrt_stats sum;
    memset(&sum, 0, sizeof(sum));
    // Contiguous runs of blocks per device.  When the blocks divide among the devices, whole blocks (every device then
    // sends its run to the others in one piece); otherwise device r takes the work units [r, r+1) * nblocks / world of
    // the batch -- whole blocks and PARTS of blocks (20 blocks on 8 GPUs: 2.5 each) -- into zeroed slots, and the
    // integer counts of all devices are summed (one all-reduce over the batch).
    const bool whole = world <= 1 || nblocks % static_cast<unsigned>(world) == 0;
    // One host thread per local device: 20 blocks are some 100 driver calls, and issued by one thread for 8 devices the
    // last device started 1.2 ms after the first (C1 on 8 GPUs: 3.2 ms of which 1.6 ms kernels, r2ab).
    std::vector<int> rcs(nloc, AGOFRT_OK);
    std::vector<std::string> errs(nloc);
    std::vector<agofrt_stats> sums(nloc);
    auto enqueue_device = [&](int i) -> int {
        agofrt_stats &sm = sums[i];
        memset(&sm, 0, sizeof(sm));
        if (!whole) {
            CU(cudaSetDevice(ctx->devs[i].id));
            CU(cudaMemsetAsync(p->dev[i].batch, 0, static_cast<size_t>(nblocks) * len * sizeof(unsigned long long), ctx->devs[i].stream));
        }
        bool first = true;
        for (unsigned b = 0; b < nblocks; ++b) {
            unsigned part_a = 0, part_b = 0;
            const int rcs_share = agofrt_block_share(nblocks, first_rank + i, world, b, &part_a, &part_b);
            if (rcs_share != AGOFRT_OK) return rcs_share;
            if (part_b <= part_a) continue;
            BlockTarget tg{i, p->dev[i].batch + static_cast<size_t>(b) * len, first};
            first = false;
            if (!whole) {
                tg.part_a = part_a;
                tg.part_b = part_b;
                tg.part_den = static_cast<uint32_t>(world);
            }
            agofrt_stats st;
            memset(&st, 0, sizeof(st));
            const int rcb = block_impl(p, primo0 + static_cast<size_t>(b) * stride, ntimesteps, leff, skip, every,
                                       options | AGOFRT_OPT_ON_DEVICE, nullptr, nullptr, &st, &tg);
            if (rcb == kNotBatchable)
                return fail(AGOFRT_ERR_ARG, "block %llu has no regular job list (single-pass minimum image not proven for its whole "
                                            "frame range): run the blocks one by one with agofrt_block", static_cast<unsigned long long>(b));
            if (rcb != AGOFRT_OK) return rcb;
            sm.pair_evals += st.pair_evals;
            sm.jobs += (whole || tg.part_a == 0) ? st.jobs : 0;   // a block shared by two devices counts once
            sm.jobs_fast += (whole || tg.part_a == 0) ? st.jobs_fast : 0;
            sm.launches += st.launches;
            sm.kernel_modes |= st.kernel_modes;
        }
        return AGOFRT_OK;
    };
    if (nloc == 1) {
        rcs[0] = enqueue_device(0);
        if (rcs[0] != AGOFRT_OK) return rcs[0];
    } else {
        std::vector<std::thread> workers;
        for (int i = 0; i < nloc; ++i)
            workers.emplace_back([&, i]() {
                try {
                    rcs[i] = enqueue_device(i);
                } catch (...) {
                    rcs[i] = on_exception();
                }
                if (rcs[i] != AGOFRT_OK) errs[i] = g_last_error;   // (the message lives in the worker's thread)
            });
        for (std::thread &w : workers) w.join();
        for (int i = 0; i < nloc; ++i)
            if (rcs[i] != AGOFRT_OK) return fail(rcs[i], "%s", errs[i].c_str());
    }
    for (int i = 0; i < nloc; ++i) {
        sum.pair_evals += sums[i].pair_evals;
        sum.jobs += sums[i].jobs;
        sum.jobs_fast += sums[i].jobs_fast;
        sum.launches += sums[i].launches;
        sum.kernel_modes |= sums[i].kernel_modes;
    }
    for (int i = 0; i < nloc; ++i) {
        CU(cudaSetDevice(ctx->devs[i].id));
        CU(cudaEventRecord(ctx->devs[i].ev_k1, ctx->devs[i].stream));
    }
    ms_enqueue = since0() - ms_alloc;
    // every device receives every block: the run of device r travels in one piece, or the partial counts are summed
    if (world > 1) {
        NcclApi &api = nccl_api();
        NC(api.GroupStart());
        if (whole) {
            const size_t per = static_cast<size_t>(nblocks / static_cast<unsigned>(world)) * len;
            for (int i = 0; i < nloc; ++i)
                NC(api.AllGather(p->dev[i].batch + static_cast<size_t>(first_rank + i) * per, p->dev[i].batch, per, ncclUint64,
                                 ctx->devs[i].comm, ctx->devs[i].stream));
        } else {
            for (int i = 0; i < nloc; ++i)
                NC(api.AllReduce(p->dev[i].batch, p->dev[i].batch, static_cast<size_t>(nblocks) * len, ncclUint64, ncclSum,
                                 ctx->devs[i].comm, ctx->devs[i].stream));
        }
        NC(api.GroupEnd());
    }
    double kernel_ms = 0, total_ms = 0;
    for (int i = 0; i < nloc; ++i) {
        Dev &dv = ctx->devs[i];
        CU(cudaSetDevice(dv.id));
        // what agofrt_block leaves behind: the last block in the plan's own histogram (agofrt_plan_last_counts, agofrt_blockavg_push)
        PlanDev &pd = p->dev[i];
        if (len > pd.ghist_len) {
            pool_free(dv, pd.ghist, pd.ghist_len * sizeof(unsigned long long));
            pd.ghist = nullptr;
            pd.ghist_len = 0;
            CU(pool_malloc(dv, &pd.ghist, len * sizeof(unsigned long long)));
            pd.ghist_len = len;
        }
        CU(cudaMemcpyAsync(pd.ghist, pd.batch + static_cast<size_t>(nblocks - 1) * len, len * sizeof(unsigned long long),
                           cudaMemcpyDeviceToDevice, dv.stream));
        CU(cudaMemcpyAsync(&p->host_flags[i], t->dev[i].flags + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaEventRecord(dv.ev_end, dv.stream));
    }
    for (int i = 0; i < nloc; ++i) {
        Dev &dv = ctx->devs[i];
        CU(cudaSetDevice(dv.id));
        CU(cudaStreamSynchronize(dv.stream));
        float a = 0, b = 0;
        CU(cudaEventElapsedTime(&a, dv.ev_begin, dv.ev_k1));
        CU(cudaEventElapsedTime(&b, dv.ev_begin, dv.ev_end));
        kernel_ms = std::max<double>(kernel_ms, a);
        total_ms = std::max<double>(total_ms, b);
        if (p->host_flags[i]) return fail(AGOFRT_ERR_NONFINITE, "minimum image did not converge within %d images", kWrapCap);
    }
    if (debug)
        fprintf(stderr, "[agofrt] batch of %u blocks x %zu counters on %d device(s): %.1f ms (buffers %.1f, enqueue %.1f, device %.1f)\n", nblocks, len,
                world, since0(), ms_alloc, ms_enqueue, total_ms);
    p->batch_blocks = nblocks;
    p->batch_block_len = len;
    p->last_len = len;
    p->last_valid = true;
    if (stats) {
        *stats = sum;
        stats->kernel_ms = kernel_ms;
        stats->total_ms = total_ms;
        stats->pair_evals_total = static_cast<uint64_t>(nblocks) * static_cast<uint64_t>((leff + (every ? every : 1) - 1) / (every ? every : 1)) *
                                  static_cast<uint64_t>((ntimesteps + (skip ? skip : 1) - 1) / (skip ? skip : 1)) *
                                  static_cast<uint64_t>(t->natoms) * t->natoms;
        stats->ndev_local = static_cast<uint32_t>(nloc);
        stats->world = static_cast<uint32_t>(world);
    }
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_plan_block_counts(agofrt_plan *p, unsigned block, uint64_t *counts_out, size_t len) try {
    if (!p || (!counts_out && len > 0)) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (block >= p->batch_blocks) return fail(AGOFRT_ERR_ARG, "block %u is not one of the %u blocks of the last agofrt_blocks", block, p->batch_blocks);
    if (len != p->batch_block_len) return fail(AGOFRT_ERR_ARG, "a block holds %zu counters, not %zu", p->batch_block_len, len);
    if (len == 0) return AGOFRT_OK;
    Dev &dv = p->ctx->devs[0];
    CU(cudaSetDevice(dv.id));
    CU(cudaMemcpyAsync(counts_out, p->dev[0].batch + static_cast<size_t>(block) * len, len * sizeof(uint64_t), cudaMemcpyDeviceToHost, dv.stream));
    CU(cudaStreamSynchronize(dv.stream));
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// ---------------------------------------------------------------------------------------------
// block averages on the device (MediaVar<Gofrt>, reference lib/include/calcoliblocchi.h:21-65)
// ---------------------------------------------------------------------------------------------
struct agofrt_blockavg {
    agofrt_ctx *ctx = nullptr;
    double *mean = nullptr, *var = nullptr;   // device 0 of the context
    size_t cap = 0, len = 0;
    unsigned blocks = 0;                      // MediaVar::iblock
    bool begun = false;
};

extern "C" int agofrt_blockavg_create(agofrt_blockavg **acc, agofrt_ctx *ctx) try {
    if (!acc || !ctx) return fail(AGOFRT_ERR_ARG, "NULL argument");
    *acc = new agofrt_blockavg;
    (*acc)->ctx = ctx;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_blockavg_destroy(agofrt_blockavg *a) try {
    if (!a) return AGOFRT_OK;
    cudaSetDevice(a->ctx->devs[0].id);
    cudaFree(a->mean);
    cudaFree(a->var);
    delete a;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_blockavg_begin(agofrt_blockavg *a, size_t len) try {
    if (!a) return fail(AGOFRT_ERR_ARG, "acc is NULL");
    Dev &dv = a->ctx->devs[0];
    CU(cudaSetDevice(dv.id));
    if (len > a->cap) {
        cudaFree(a->mean);
        cudaFree(a->var);
        a->mean = a->var = nullptr;
        a->cap = 0;
        CU(cudaMalloc(&a->mean, len * sizeof(double)));
        CU(cudaMalloc(&a->var, len * sizeof(double)));
        a->cap = len;
    }
    a->len = len;
    a->blocks = 0;
    a->begun = true;
    if (len > 0) {
        CU(cudaMemsetAsync(a->mean, 0, len * sizeof(double), dv.stream));   // +0.0, VectorOp::azzera
        CU(cudaMemsetAsync(a->var, 0, len * sizeof(double), dv.stream));
    }
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_blockavg_push(agofrt_blockavg *a, agofrt_plan *p, double incr) try {
    if (!a || !p) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (!a->begun) return fail(AGOFRT_ERR_ARG, "agofrt_blockavg_push before agofrt_blockavg_begin");
    if (p->ctx != a->ctx) return fail(AGOFRT_ERR_ARG, "plan and accumulator belong to different contexts");
    if (!p->last_valid) return fail(AGOFRT_ERR_ARG, "the plan holds no complete block on the device (run agofrt_block first)");
    // VectorOp's "Trying to operate on VectorOp of different sizes!" (reference lib/include/operazionisulista.h:48)
    if (p->last_len != a->len)
        return fail(AGOFRT_ERR_ARG, "block of %zu elements pushed into an average of %zu", p->last_len, a->len);
    Dev &dv = a->ctx->devs[0];
    CU(cudaSetDevice(dv.id));
    // same stream as the pair kernels and the all-reduce of the block: ordered after them, and before the next
    // agofrt_block zeroes the counts
    CU(launch_blockavg_push(p->dev[0].ghist, incr, a->blocks, a->mean, a->var, a->len, dv.sm_count, dv.stream));
    ++a->blocks;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// all the blocks of the plan's last agofrt_blocks, in block order, in one launch
extern "C" int agofrt_blockavg_push_blocks(agofrt_blockavg *a, agofrt_plan *p, double incr) try {
    if (!a || !p) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (!a->begun) return fail(AGOFRT_ERR_ARG, "agofrt_blockavg_push_blocks before agofrt_blockavg_begin");
    if (p->ctx != a->ctx) return fail(AGOFRT_ERR_ARG, "plan and accumulator belong to different contexts");
    if (p->batch_blocks == 0) return fail(AGOFRT_ERR_ARG, "the plan holds no batch of blocks (run agofrt_blocks first)");
    if (p->batch_block_len != a->len)
        return fail(AGOFRT_ERR_ARG, "blocks of %zu elements pushed into an average of %zu", p->batch_block_len, a->len);
    Dev &dv = a->ctx->devs[0];
    CU(cudaSetDevice(dv.id));
    CU(launch_blockavg_push_blocks(p->dev[0].batch, p->batch_blocks, incr, a->blocks, a->mean, a->var, a->len, dv.sm_count, dv.stream));
    a->blocks += p->batch_blocks;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_blockavg_end(agofrt_blockavg *a, unsigned n_b, double *mean_out, double *var_out) try {
    if (!a) return fail(AGOFRT_ERR_ARG, "acc is NULL");
    if (!a->begun) return fail(AGOFRT_ERR_ARG, "agofrt_blockavg_end before agofrt_blockavg_begin");
    Dev &dv = a->ctx->devs[0];
    CU(cudaSetDevice(dv.id));
    if (a->len > 0 && mean_out)
        CU(cudaMemcpyAsync(mean_out, a->mean, a->len * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
    if (a->len > 0 && var_out)
        CU(cudaMemcpyAsync(var_out, a->var, a->len * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
    CU(cudaStreamSynchronize(dv.stream));
    // MediaVar::calcola_end: *Tvar /= ((n_b-1)*n_b), unsigned arithmetic, one pass over the result -- on the host, so
    // that even the degenerate n_b = 1 (0/0) gives the host's own NaN
    if (var_out) {
        const double denom = static_cast<double>((n_b - 1u) * n_b);
        for (size_t k = 0; k < a->len; ++k) var_out[k] /= denom;
    }
    a->begun = false;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_plan_last_counts(agofrt_plan *p, uint64_t *counts_out, size_t len) try {
    if (!p || (len > 0 && !counts_out)) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (!p->last_valid) return fail(AGOFRT_ERR_ARG, "the plan holds no complete block on the device (run agofrt_block first)");
    if (len != p->last_len) return fail(AGOFRT_ERR_ARG, "the last block has %zu words, not %zu", p->last_len, len);
    if (len == 0) return AGOFRT_OK;
    Dev &dv = p->ctx->devs[0];
    CU(cudaSetDevice(dv.id));
    // (straight into the caller's buffer: the pinned staging of agofrt_block need not exist -- agofrt_blocks has none)
    CU(cudaMemcpyAsync(counts_out, p->dev[0].ghist, len * sizeof(unsigned long long), cudaMemcpyDeviceToHost, dv.stream));
    CU(cudaStreamSynchronize(dv.stream));
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// ---------------------------------------------------------------------------------------------
// neighbour-count histogram (IstogrammaAtomiRaggio::calculate)
// ---------------------------------------------------------------------------------------------
extern "C" int agofrt_neighbour_hist(agofrt_traj *t, double r, size_t tstart, unsigned ntimesteps, unsigned skip,
                                     uint64_t *hist_inout, agofrt_stats *stats) try {
    if (!t) return fail(AGOFRT_ERR_ARG, "traj is NULL");
    agofrt_ctx *ctx = t->ctx;
    if (skip < 1) skip = 1;   // reference lib/src/istogrammaatomiraggio.cpp:19
    if (stats) memset(stats, 0, sizeof(*stats));
    const size_t hstride = t->natoms + 1;
    const size_t hlen = static_cast<size_t>(t->ntypes) * hstride;
    if (ntimesteps == 0 || t->natoms == 0) return AGOFRT_OK;
    if (!hist_inout) return fail(AGOFRT_ERR_ARG, "hist is NULL");
    const size_t last = tstart + static_cast<size_t>((ntimesteps - 1) / skip) * skip;
    if (tstart < t->first_frame || last >= t->first_frame + t->nframes)
        return fail(AGOFRT_ERR_WINDOW, "frames [%zu,%zu] are not all in the device window [%zu,%zu)", tstart, last,
                    t->first_frame, t->first_frame + t->nframes);
    if (t->bad_box || t->has_inf)
        return fail(AGOFRT_ERR_NONFINITE, "the window holds an infinite coordinate or a non-positive / non-finite box edge");
    int rc = ensure_local_comm(ctx);
    if (rc != AGOFRT_OK) return rc;
    const int nloc = static_cast<int>(ctx->devs.size());
    const int world = ctx->world > 0 ? ctx->world : nloc;
    const int first_rank = ctx->world > 0 ? ctx->first_rank : 0;

    std::vector<int> fr_fast, fr_gen;
    for (size_t f = tstart; f <= last; f += skip) {
        const size_t rel = f - t->first_frame;
        (job_is_single_pass(t, rel, rel) && !t->has_nan ? fr_fast : fr_gen).push_back(static_cast<int>(rel));
    }
    const size_t nfr = fr_fast.size() + fr_gen.size();
    const int tile = neighbour_tile_atoms();
    const int n_itiles = std::max(1, (t->npad + tile - 1) / tile);
    // j chunks: enough work units for every SM of every device even when the call lists few frames
    int total_ctas = 0;
    for (const Dev &d : ctx->devs) total_ctas += 8 * d.sm_count;
    total_ctas = total_ctas / nloc * world;
    int n_jchunks = static_cast<int>(std::min<uint64_t>((t->npad + 1023) / 1024,
                                                        (4ull * total_ctas + nfr * n_itiles - 1) / std::max<uint64_t>(1, nfr * n_itiles)));
    n_jchunks = std::max(1, n_jchunks);
    int jchunk = ((t->npad + n_jchunks - 1) / n_jchunks + 7) / 8 * 8;
    n_jchunks = std::max(1, (t->npad + jchunk - 1) / jchunk);
    const bool tri = t->stride == 9;
    double kernel_ms = 0;
    unsigned launches = 0;
    for (int i = 0; i < nloc; ++i) {
        Dev &dv = ctx->devs[i];
        TrajDev &td = t->dev[i];
        CU(cudaSetDevice(dv.id));
        if (!td.nb_hist) CU(cudaMalloc(&td.nb_hist, hlen * sizeof(unsigned long long)));
        if (nfr > td.nb_frames_cap) {
            cudaFree(td.nb_frames);
            td.nb_frames = nullptr;
            td.nb_frames_cap = 0;
            CU(cudaMalloc(&td.nb_frames, nfr * sizeof(int)));
            td.nb_frames_cap = nfr;
        }
        const size_t ncounts = nfr * static_cast<size_t>(t->ntypes) * t->npad;
        if (ncounts > td.nb_counts_len) {
            cudaFree(td.nb_counts);
            td.nb_counts = nullptr;
            td.nb_counts_len = 0;
            CU(cudaMalloc(&td.nb_counts, ncounts * sizeof(unsigned int)));
            td.nb_counts_len = ncounts;
        }
        CU(cudaMemsetAsync(td.nb_hist, 0, hlen * sizeof(unsigned long long), dv.stream));
        CU(cudaMemsetAsync(td.nb_counts, 0, ncounts * sizeof(unsigned int), dv.stream));
        CU(cudaMemsetAsync(td.flags + 1, 0, sizeof(unsigned int), dv.stream));
        if (!fr_fast.empty())
            CU(cudaMemcpyAsync(td.nb_frames, fr_fast.data(), fr_fast.size() * sizeof(int), cudaMemcpyHostToDevice, dv.stream));
        if (!fr_gen.empty())
            CU(cudaMemcpyAsync(td.nb_frames + fr_fast.size(), fr_gen.data(), fr_gen.size() * sizeof(int),
                               cudaMemcpyHostToDevice, dv.stream));
        CU(cudaEventRecord(dv.ev_k0, dv.stream));
        for (int pass = 0; pass < 2; ++pass) {
            const std::vector<int> &list = pass == 0 ? fr_fast : fr_gen;
            if (list.empty()) continue;
            // the (frame, i tile) pairs are dealt to the devices; all the j chunks of a pair run on its device, so the
            // per-atom counts are complete there and only the histogram is all-reduced
            uint64_t vb = 0, ve = 0;
            agofrt_shard_range(static_cast<uint64_t>(list.size()) * n_itiles, first_rank + i, world, &vb, &ve);
            if (ve <= vb) continue;
            NeighbourParams np;
            np.pos = td.pos;
            np.box = td.box6;
            np.perm = td.perm;
            np.type_start = td.type_start;
            np.frames = td.nb_frames + (pass == 0 ? 0 : fr_fast.size());
            np.hist = td.nb_hist;
            np.counts = td.nb_counts + (pass == 0 ? 0 : fr_fast.size()) * static_cast<size_t>(t->ntypes) * t->npad;
            np.error_flag = td.flags + 1;
            np.r2 = r * r;   // reference lib/src/istogrammaatomiraggio.cpp:17
            np.unit_begin = static_cast<unsigned>(vb * n_jchunks);
            np.unit_end = static_cast<unsigned>(ve * n_jchunks);
            np.npad = t->npad;
            np.ntypes = t->ntypes;
            np.n_itiles = n_itiles;
            np.n_jchunks = n_jchunks;
            np.jchunk = jchunk;
            np.hist_stride = hstride;
            const uint64_t units = (ve - vb) * n_jchunks;
            const int grid = static_cast<int>(std::min<uint64_t>(units, static_cast<uint64_t>(dv.sm_count) * 8));
            CU(launch_neighbour_kernel(tri, pass == 0, grid, dv.stream, np));
            const uint64_t words = (ve - vb) * static_cast<uint64_t>(t->ntypes) * tile;
            CU(launch_neighbour_finish(static_cast<int>(std::min<uint64_t>((words + 255) / 256, static_cast<uint64_t>(dv.sm_count) * 8)),
                                       dv.stream, np, static_cast<unsigned>(vb), static_cast<unsigned>(ve)));
            launches += 2;
        }
        CU(cudaEventRecord(dv.ev_k1, dv.stream));
    }
    if (ctx->comm_ready && world > 1) {
        NcclApi &api = nccl_api();
        NC(api.GroupStart());
        for (int i = 0; i < nloc; ++i)
            NC(api.AllReduce(t->dev[i].nb_hist, t->dev[i].nb_hist, hlen, ncclUint64, ncclSum, ctx->devs[i].comm, ctx->devs[i].stream));
        NC(api.GroupEnd());
    }
    std::vector<unsigned long long> host(hlen);
    const int nread = (ctx->comm_ready && world > 1) ? 1 : nloc;   // without a communicator the local partials are summed here
    for (int i = 0; i < nloc; ++i) {
        Dev &dv = ctx->devs[i];
        CU(cudaSetDevice(dv.id));
        unsigned int flag = 0;
        CU(cudaMemcpyAsync(&flag, t->dev[i].flags + 1, sizeof(flag), cudaMemcpyDeviceToHost, dv.stream));
        if (i < nread) CU(cudaMemcpyAsync(host.data(), t->dev[i].nb_hist, hlen * sizeof(unsigned long long), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaStreamSynchronize(dv.stream));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, dv.ev_k0, dv.ev_k1));
        kernel_ms = std::max<double>(kernel_ms, ms);
        if (flag) return fail(AGOFRT_ERR_NONFINITE, "minimum image did not converge within %d images", kWrapCap);
        if (i < nread)
            for (size_t k = 0; k < hlen; ++k) hist_inout[k] += host[k];
    }
    if (stats) {
        stats->kernel_ms = kernel_ms;
        stats->total_ms = kernel_ms;
        stats->pair_evals_total = static_cast<uint64_t>(nfr) * t->natoms * t->natoms;
        stats->pair_evals = stats->pair_evals_total / static_cast<uint64_t>(world) * nloc;
        stats->jobs = nfr;
        stats->jobs_fast = fr_fast.size();
        stats->launches = launches;
        stats->ndev_local = static_cast<uint32_t>(nloc);
        stats->world = static_cast<uint32_t>(world);
    }
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// ---------------------------------------------------------------------------------------------
// mean square displacement (MSD<T>::calculate)
// ---------------------------------------------------------------------------------------------
extern "C" int agofrt_traj_set_cm(agofrt_traj *t, size_t first_frame, size_t nframes, const double *cm) try {
    if (!t) return fail(AGOFRT_ERR_ARG, "traj is NULL");
    if (nframes > 0 && !cm) return fail(AGOFRT_ERR_ARG, "cm is NULL");
    t->cm.assign(cm, cm + nframes * static_cast<size_t>(t->ntypes) * 3);
    t->cm_first = first_frame;
    t->cm_frames = nframes;
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

extern "C" int agofrt_msd(agofrt_traj *t, size_t primo, unsigned ntimesteps, unsigned leff, unsigned skip, int cm_msd,
                          int cm_self, double *out, agofrt_stats *stats) try {
    if (!t) return fail(AGOFRT_ERR_ARG, "traj is NULL");
    if (skip < 1) skip = 1;
    if (stats) memset(stats, 0, sizeof(*stats));
    const int nt = t->ntypes, f_cm = cm_msd ? 2 : 1;
    const size_t olen = static_cast<size_t>(leff) * f_cm * nt;
    if (olen == 0) return AGOFRT_OK;
    if (!out) return fail(AGOFRT_ERR_ARG, "out is NULL");
    if (ntimesteps == 0) {
        std::fill(out, out + olen, 0.0);
        return AGOFRT_OK;
    }
    const size_t last = primo + static_cast<size_t>((ntimesteps - 1) / skip) * skip + (leff - 1);
    if (primo < t->first_frame || last >= t->first_frame + t->nframes)
        return fail(AGOFRT_ERR_WINDOW, "the calculation needs frames [%zu,%zu] but the device window holds [%zu,%zu)", primo, last,
                    t->first_frame, t->first_frame + t->nframes);
    const bool need_cm = cm_msd || cm_self;
    if (need_cm && (t->cm_first != t->first_frame || t->cm_frames < t->nframes))
        return fail(AGOFRT_ERR_ARG, "centres of mass of the uploaded window were not set (agofrt_traj_set_cm)");
    // tiles of 256 real atoms, never across a type boundary (real atoms come first in every type group)
    std::vector<int> type_count(nt, 0);
    for (size_t a = 0; a < t->natoms; ++a) type_count[t->type_id[a]]++;
    const int tile = msd_tile_atoms();
    std::vector<int> tile_type, tile_start, tile_count;
    for (int ty = 0; ty < nt; ++ty)
        for (int o = 0; o < type_count[ty]; o += tile) {
            tile_type.push_back(ty);
            tile_start.push_back(t->type_start[ty] + o);
            tile_count.push_back(std::min(tile, type_count[ty] - o));
        }
    const int ntiles = static_cast<int>(tile_type.size());
    Dev &dv = t->ctx->devs[0];   // a bandwidth-bound O(N) pass per (lag, origin): one device is plenty
    TrajDev &td = t->dev[0];
    CU(cudaSetDevice(dv.id));
    int *d_tiles = nullptr;
    double *d_cm = nullptr, *d_partial = nullptr, *d_out = nullptr;
    auto body = [&]() -> int {
        CU(cudaMalloc(&d_tiles, (3 * static_cast<size_t>(std::max(ntiles, 1)) + nt) * sizeof(int)));
        CU(cudaMalloc(&d_partial, std::max<size_t>(1, static_cast<size_t>(leff) * ntiles) * sizeof(double)));
        CU(cudaMalloc(&d_out, olen * sizeof(double)));
        if (need_cm) {
            CU(cudaMalloc(&d_cm, std::max<size_t>(1, t->cm.size()) * sizeof(double)));
            CU(cudaMemcpyAsync(d_cm, t->cm.data(), t->cm.size() * sizeof(double), cudaMemcpyHostToDevice, dv.stream));
        }
        if (ntiles > 0) {
            CU(cudaMemcpyAsync(d_tiles, tile_type.data(), ntiles * sizeof(int), cudaMemcpyHostToDevice, dv.stream));
            CU(cudaMemcpyAsync(d_tiles + ntiles, tile_start.data(), ntiles * sizeof(int), cudaMemcpyHostToDevice, dv.stream));
            CU(cudaMemcpyAsync(d_tiles + 2 * ntiles, tile_count.data(), ntiles * sizeof(int), cudaMemcpyHostToDevice, dv.stream));
        }
        CU(cudaMemcpyAsync(d_tiles + 3 * static_cast<size_t>(std::max(ntiles, 1)), type_count.data(), nt * sizeof(int),
                           cudaMemcpyHostToDevice, dv.stream));
        MsdParams mp;
        mp.pos = td.pos;
        mp.cm = d_cm;
        mp.tile_type = d_tiles;
        mp.tile_start = d_tiles + ntiles;
        mp.tile_count = d_tiles + 2 * ntiles;
        mp.type_count = d_tiles + 3 * static_cast<size_t>(std::max(ntiles, 1));
        mp.partial = d_partial;
        mp.out = d_out;
        mp.npad = t->npad;
        mp.ntypes = nt;
        mp.ntiles = ntiles;
        mp.leff = static_cast<int>(leff);
        mp.f0 = static_cast<int>(primo - t->first_frame);
        mp.ntimesteps = static_cast<int>(ntimesteps);
        mp.skip = static_cast<int>(skip);
        mp.cm_msd = cm_msd ? 1 : 0;
        mp.cm_self = cm_self ? 1 : 0;
        CU(cudaEventRecord(dv.ev_k0, dv.stream));
        CU(launch_msd(mp, dv.stream));
        CU(cudaEventRecord(dv.ev_k1, dv.stream));
        CU(cudaMemcpyAsync(out, d_out, olen * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaStreamSynchronize(dv.stream));
        return AGOFRT_OK;
    };
    const int rc = body();
    float ms = 0;
    if (rc == AGOFRT_OK) cudaEventElapsedTime(&ms, dv.ev_k0, dv.ev_k1);
    cudaFree(d_tiles);
    cudaFree(d_cm);
    cudaFree(d_partial);
    cudaFree(d_out);
    if (rc != AGOFRT_OK) return rc;
    if (stats) {
        const uint64_t norig = (ntimesteps + skip - 1) / skip;
        stats->kernel_ms = ms;
        stats->total_ms = ms;
        stats->jobs = static_cast<uint64_t>(leff) * norig;
        stats->pair_evals = stats->pair_evals_total = stats->jobs * t->natoms;   // displacement evaluations
        stats->launches = 2;
        stats->ndev_local = 1;
        stats->world = 1;
    }
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// ---------------------------------------------------------------------------------------------
// neighbour lists and spherical-harmonic densities of one frame (device 0)
// ---------------------------------------------------------------------------------------------
static int atom_tables(agofrt_traj *t, int **d_slot, int **d_type) {
    // atom (caller's numbering) -> device slot, and its dense type
    std::vector<int> slot(t->natoms, 0);
    for (int s = 0; s < t->npad; ++s)
        if (t->perm[s] >= 0) slot[t->perm[s]] = s;
    CU(cudaMalloc(d_slot, std::max<size_t>(t->natoms, 1) * sizeof(int)));
    CU(cudaMalloc(d_type, std::max<size_t>(t->natoms, 1) * sizeof(int)));
    if (t->natoms) {
        CU(cudaMemcpy(*d_slot, slot.data(), t->natoms * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(*d_type, t->type_id.data(), t->natoms * sizeof(int), cudaMemcpyHostToDevice));
    }
    return AGOFRT_OK;
}

extern "C" int agofrt_neighbours(agofrt_traj *t, size_t frame, const uint64_t *nneigh, const double *cutoff2, int sort,
                                 uint64_t *list_out, double *rpos_out) try {
    if (!t || !nneigh || !cutoff2 || !list_out || !rpos_out) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (t->ntypes > kLsMaxTypes) return fail(AGOFRT_ERR_TOO_LARGE, "neighbour lists handle at most %d atom types", kLsMaxTypes);
    if (frame < t->first_frame || frame >= t->first_frame + t->nframes)
        return fail(AGOFRT_ERR_WINDOW, "frame %zu is not in the uploaded window", frame);
    if (t->bad_box || t->has_inf) return fail(AGOFRT_ERR_NONFINITE, "the window holds an infinite coordinate or a non-positive / non-finite box edge");
    NeighListParams p;
    size_t words = 0, doubles = 0;
    for (int k = 0; k < t->ntypes; ++k) {
        p.nneigh[k] = nneigh[k];
        p.cutoff2[k] = cutoff2[k];
        p.list_offset[k] = words;
        p.rpos_offset[k] = doubles;
        words += (nneigh[k] + 1) * t->natoms;
        doubles += nneigh[k] * t->natoms * 4;
    }
    if (t->natoms == 0) return AGOFRT_OK;
    Dev &dv = t->ctx->devs[0];
    TrajDev &td = t->dev[0];
    CU(cudaSetDevice(dv.id));
    int *d_slot = nullptr, *d_type = nullptr;
    unsigned long long *d_list = nullptr;
    double *d_rpos = nullptr;
    unsigned int *d_flags = nullptr;
    unsigned int flags[2] = {0, 0};
    auto body = [&]() -> int {
        int rc = atom_tables(t, &d_slot, &d_type);
        if (rc != AGOFRT_OK) return rc;
        CU(cudaMalloc(&d_list, words * sizeof(unsigned long long)));
        CU(cudaMalloc(&d_rpos, std::max<size_t>(doubles, 1) * sizeof(double)));
        CU(cudaMalloc(&d_flags, 2 * sizeof(unsigned int)));
        CU(cudaMemsetAsync(d_list, 0, words * sizeof(unsigned long long), dv.stream));   // update_neigh zeroes the lists (neighbour.cpp:10-12)
        CU(cudaMemsetAsync(d_rpos, 0, std::max<size_t>(doubles, 1) * sizeof(double), dv.stream));   // (the reference leaves these uninitialised)
        CU(cudaMemsetAsync(d_flags, 0, 2 * sizeof(unsigned int), dv.stream));
        p.pos = td.pos;
        p.box = td.box6;
        p.atom_slot = d_slot;
        p.atom_type = d_type;
        p.list = d_list;
        p.rpos = d_rpos;
        p.flags = d_flags;
        p.natoms = static_cast<int>(t->natoms);
        p.npad = t->npad;
        p.ntypes = t->ntypes;
        p.frame = static_cast<int>(frame - t->first_frame);
        p.triclinic = t->stride == 9 ? 1 : 0;
        p.sort = sort ? 1 : 0;
        CU(launch_neigh_list(p, dv.stream));
        CU(cudaMemcpyAsync(list_out, d_list, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost, dv.stream));
        if (doubles) CU(cudaMemcpyAsync(rpos_out, d_rpos, doubles * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaStreamSynchronize(dv.stream));
        return AGOFRT_OK;
    };
    const int rc = body();
    cudaFree(d_slot);
    cudaFree(d_type);
    cudaFree(d_list);
    cudaFree(d_rpos);
    cudaFree(d_flags);
    if (rc != AGOFRT_OK) return rc;
    if (flags[1]) return fail(AGOFRT_ERR_NONFINITE, "minimum image did not converge within %d images", kWrapCap);
    if (flags[0]) return fail(AGOFRT_ERR_TOO_LARGE, "Too many neighbours in shell!");
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// SpecialFunctions::realSpericalHarmonics_coeff<double>(l, m, l-m+1) (reference lib/include/specialfunctions.h:34-50),
// the same operations in the same order; compiled without FMA contraction and fast-math (build.py)
static double real_sh_coeff(int l, int m) {
    const double pi = 3.14159265358979323846264338327950288419716939937510;
    if (m == 0) return std::sqrt(static_cast<double>(2 * l + 1) / (4 * pi));
    if (m < 0) m = -m;
    double value = 1.0;
    for (long v = l - m + 1; v <= l + m; ++v) value = value * static_cast<double>(v);
    return std::sqrt(static_cast<double>(2 * l + 1) / (4 * pi) * 2 / value) * std::pow(-1, m);
}

extern "C" int agofrt_sh_density(agofrt_traj *t, size_t frame, int lmax, unsigned nbin, const double *rminmax, double *result_out,
                                 int *counter_out) try {
    if (!t || !rminmax || !result_out) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (lmax < 0 || lmax > kShMaxL) return fail(AGOFRT_ERR_ARG, "lmax must be in [0, %d]", kShMaxL);
    if (frame < t->first_frame || frame >= t->first_frame + t->nframes)
        return fail(AGOFRT_ERR_WINDOW, "frame %zu is not in the uploaded window", frame);
    if (t->bad_box || t->has_inf) return fail(AGOFRT_ERR_NONFINITE, "the window holds an infinite coordinate or a non-positive / non-finite box edge");
    const int nt = t->ntypes, nl = (lmax + 1) * (lmax + 1);
    const size_t cells = t->natoms * static_cast<size_t>(nt) * nbin;
    if (cells == 0) return AGOFRT_OK;
    std::vector<double> rmin(nt * nt), dr(nt * nt), coeff(nl, 0.0);
    for (int k = 0; k < nt * nt; ++k) {
        rmin[k] = rminmax[2 * k];
        dr[k] = (rminmax[2 * k + 1] - rminmax[2 * k]) / static_cast<double>(nbin);   // sphericalbase.h:39-41
    }
    for (int l = 0; l <= lmax; ++l)
        for (int m = 0; m <= l; ++m) coeff[l * (lmax + 1) + m] = real_sh_coeff(l, m);
    Dev &dv = t->ctx->devs[0];
    TrajDev &td = t->dev[0];
    CU(cudaSetDevice(dv.id));
    int *d_slot = nullptr, *d_type = nullptr, *d_counter = nullptr;
    double *d_res = nullptr, *d_par = nullptr;
    unsigned int *d_flags = nullptr;
    unsigned int flags[2] = {0, 0};
    auto body = [&]() -> int {
        int rc = atom_tables(t, &d_slot, &d_type);
        if (rc != AGOFRT_OK) return rc;
        CU(cudaMalloc(&d_res, cells * nl * sizeof(double)));
        CU(cudaMalloc(&d_counter, cells * sizeof(int)));
        CU(cudaMalloc(&d_par, (2 * nt * nt + nl) * sizeof(double)));
        CU(cudaMalloc(&d_flags, 2 * sizeof(unsigned int)));
        CU(cudaMemsetAsync(d_res, 0, cells * nl * sizeof(double), dv.stream));
        CU(cudaMemsetAsync(d_counter, 0, cells * sizeof(int), dv.stream));
        CU(cudaMemsetAsync(d_flags, 0, 2 * sizeof(unsigned int), dv.stream));
        CU(cudaMemcpyAsync(d_par, rmin.data(), nt * nt * sizeof(double), cudaMemcpyHostToDevice, dv.stream));
        CU(cudaMemcpyAsync(d_par + nt * nt, dr.data(), nt * nt * sizeof(double), cudaMemcpyHostToDevice, dv.stream));
        CU(cudaMemcpyAsync(d_par + 2 * nt * nt, coeff.data(), nl * sizeof(double), cudaMemcpyHostToDevice, dv.stream));
        ShDensityParams p;
        p.pos = td.pos;
        p.box = td.box6;
        p.atom_slot = d_slot;
        p.atom_type = d_type;
        p.rmin = d_par;
        p.dr = d_par + nt * nt;
        p.coeff = d_par + 2 * nt * nt;
        p.result = d_res;
        p.counter = d_counter;
        p.flags = d_flags;
        p.natoms = static_cast<int>(t->natoms);
        p.npad = t->npad;
        p.ntypes = nt;
        p.frame = static_cast<int>(frame - t->first_frame);
        p.triclinic = t->stride == 9 ? 1 : 0;
        p.lmax = lmax;
        p.nbin = static_cast<int>(nbin);
        CU(launch_sh_density(p, dv.stream));
        CU(cudaMemcpyAsync(result_out, d_res, cells * nl * sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
        if (counter_out) CU(cudaMemcpyAsync(counter_out, d_counter, cells * sizeof(int), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, dv.stream));
        CU(cudaStreamSynchronize(dv.stream));
        return AGOFRT_OK;
    };
    const int rc = body();
    cudaFree(d_slot);
    cudaFree(d_type);
    cudaFree(d_res);
    cudaFree(d_counter);
    cudaFree(d_par);
    cudaFree(d_flags);
    if (rc != AGOFRT_OK) return rc;
    if (flags[1]) return fail(AGOFRT_ERR_NONFINITE, "minimum image did not converge within %d images", kWrapCap);
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}

// ---------------------------------------------------------------------------------------------
// FP64 issue-rate microbenchmark
// ---------------------------------------------------------------------------------------------
extern "C" int agofrt_fp64_peak(agofrt_ctx *ctx, int local_device, double seconds, double *dfma_per_second) try {
    if (!ctx || !dfma_per_second) return fail(AGOFRT_ERR_ARG, "NULL argument");
    if (local_device < 0 || local_device >= static_cast<int>(ctx->devs.size()))
        return fail(AGOFRT_ERR_ARG, "local_device out of range");
    Dev &dv = ctx->devs[local_device];
    CU(cudaSetDevice(dv.id));
    if (!dv.peak_sink) CU(cudaMalloc(&dv.peak_sink, sizeof(double)));
    const int blocks = dv.sm_count * 8;
    unsigned long long count = 0;
    // calibrate
    int iters = 2000;
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        CU(cudaEventRecord(dv.ev_k0, dv.stream));
        CU(launch_dfma_peak(dv.peak_sink, blocks, iters, dv.stream, &count));
        CU(cudaEventRecord(dv.ev_k1, dv.stream));
        CU(cudaStreamSynchronize(dv.stream));
        CU(cudaEventElapsedTime(&ms, dv.ev_k0, dv.ev_k1));
    }
    if (!(seconds > 0)) seconds = 0.5;
    const double per_launch_ms = std::max(ms, 1e-3f);
    int nlaunch = static_cast<int>(std::min(200.0, std::max(3.0, seconds * 1e3 / per_launch_ms)));
    // Sustained rate: the median launch of the run's last three quarters.  The first quarter is warm-up -- a GPU
    // that idled while the host generated a trajectory needs a few hundred milliseconds to reach its clocks,
    // and an average over that ramp under-reports the peak (and flatters every roofline fraction).
    std::vector<double> rates;
    rates.reserve(nlaunch);
    for (int k = 0; k < nlaunch; ++k) {
        CU(cudaEventRecord(dv.ev_k0, dv.stream));
        CU(launch_dfma_peak(dv.peak_sink, blocks, iters, dv.stream, &count));
        CU(cudaEventRecord(dv.ev_k1, dv.stream));
        CU(cudaStreamSynchronize(dv.stream));
        CU(cudaEventElapsedTime(&ms, dv.ev_k0, dv.ev_k1));
        if (k >= nlaunch / 4) rates.push_back(count / (std::max(ms, 1e-6f) * 1e-3));
    }
    std::sort(rates.begin(), rates.end());
    *dfma_per_second = rates[rates.size() / 2];
    return AGOFRT_OK;
} catch (...) {
    return on_exception();
}
