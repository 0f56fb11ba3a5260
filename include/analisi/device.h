// device.h -- thin C++ layer of the host classes over the C ABI of libagofrt.so (include/agofrt.h).
//
// The reference has no such layer: its Gofrt reaches the trajectory buffers directly
// (lib/src/gofrt.cpp:101-103).  Here every pair evaluation happens on the GPU, so the host classes
// (BaseTrajectory, Gofrt) talk to the device only through these few calls.  Errors of the C ABI
// become std::runtime_error, the reference's own error convention (lib/src/gofrt.cpp:81-83), so the
// CLI still exits with code 1 and python still sees a RuntimeError.
#ifndef ANALISI_B200_DEVICE_H
#define ANALISI_B200_DEVICE_H

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "agofrt.h"

namespace analisi_device {

inline void check(int rc, const char *what) {
    if (rc != AGOFRT_OK)
        throw std::runtime_error(std::string(what) + ": " + agofrt_last_error() + "\n");
}

// The GPUs of this process.  Default: every visible device (one node, work units of a block sharded
// over them, one NCCL all-reduce of the integer histograms per block).  ANALISI_DEVICES="0,2" selects
// a subset.  There is no CPU path: without a usable sm_100 device the first use throws.
class Context {
public:
    static Context &instance();
    agofrt_ctx *handle() { return ctx_; }
    int ndev() const { return agofrt_ctx_ndev(ctx_); }
    // One process per GPU (torchrun, mpirun): rank 0 makes the id, the launcher broadcasts the bytes, every process
    // joins with its rank.  Afterwards the work units of every block are sharded over all processes, the integer
    // histograms are all-reduced, and windows are uploaded once per box (every process copies its share of the frames,
    // the shares travel GPU to GPU).  Replaces the reference's Mp (lib/include/mp.h:26-55).
    static std::string comm_unique_id();
    void comm_join(const std::string &id, int rank, int world);
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;

private:
    Context();
    ~Context();
    agofrt_ctx *ctx_ = nullptr;
};

// Page-locked host buffer for trajectory windows (the reference uses fftw_malloc as an aligned
// allocator, lib/src/trajectory.cpp:362-371); pinned memory makes the H2D upload a single DMA.
class PinnedBuffer {
public:
    PinnedBuffer() = default;
    ~PinnedBuffer() { release(); }
    PinnedBuffer(const PinnedBuffer &) = delete;
    PinnedBuffer &operator=(const PinnedBuffer &) = delete;
    void resize(size_t ndoubles);   // contents are not preserved
    void release();
    double *data() { return ptr_; }
    size_t size() const { return n_; }
    void swap(PinnedBuffer &o) {
        double *p = ptr_;
        ptr_ = o.ptr_;
        o.ptr_ = p;
        size_t n = n_;
        n_ = o.n_;
        o.n_ = n;
    }

private:
    double *ptr_ = nullptr;
    size_t n_ = 0;
};

// Device-resident window of one trajectory (agofrt_traj) + what it currently holds.
class Window {
public:
    Window() = default;
    ~Window() { release(); }
    Window(const Window &) = delete;
    Window &operator=(const Window &) = delete;

    // (re)create for natoms / box stride / dense type ids / capacity in frames
    void create(size_t natoms, int box_stride, const int *type_id, int ntypes, size_t max_frames);
    void release();
    bool valid() const { return traj_ != nullptr; }
    size_t capacity() const { return cap_; }
    // replace the device window by frames [first, first+n) from host buffers
    void upload(size_t first, size_t n, const double *pos_aos, const double *box_internal);
    // the same with BaseTrajectory::pbc_wrap applied on the device; pos_aos comes back wrapped
    void upload_wrap(size_t first, size_t n, double *pos_aos_inout, const double *box_internal);
    // general form (agofrt_traj_upload_ex): pos_aos is only read and may be pageable; `wrap` wraps on the device;
    // wrapped_out (may be NULL, may be pos_aos) receives the wrapped frames; the frames are dealt to the GPUs of the
    // process / communicator and exchanged device to device
    void upload_shared(size_t first, size_t n, const double *pos_aos, const double *box_internal, bool wrap, double *wrapped_out);
    // frames of the device window back to the host, caller's atom order
    void download(size_t first, size_t n, double *pos_aos_out);
    // LAMMPS dump records as the source (agofrt_traj_set_ids / agofrt_traj_upload_records): the id -> slot table once per
    // window object, then frames as lists of chunks of raw records.  upload_records returns false when an atom's type
    // changes inside the window (the caller then reads that window on the host, with the reference's warning).
    // per-frame rotation matrices next to the window (agofrt_traj_set_rotation)
    void set_rotation(size_t first, size_t n, const double *q9);
    void get_rotation(size_t frame, double *q9);
    bool ids_set() const { return ids_set_; }
    void set_ids(const int *slot_to_id, const int *slot_raw_type);
    bool upload_records(size_t first, size_t n, const void *const *chunk_ptr, const int *chunk_atoms, const size_t *frame_chunk,
                        const double *box_internal, bool wrap);
    void swap(Window &o) {
        agofrt_traj *t = traj_;
        traj_ = o.traj_;
        o.traj_ = t;
        size_t c = cap_;
        cap_ = o.cap_;
        o.cap_ = c;
        bool i = ids_set_;
        ids_set_ = o.ids_set_;
        o.ids_set_ = i;
    }
    agofrt_traj *handle() { return traj_; }
    // bumped on every create(): plans made on an older handle must be rebuilt
    uint64_t generation() const { return generation_; }

private:
    agofrt_traj *traj_ = nullptr;
    size_t cap_ = 0;
    uint64_t generation_ = 0;
    bool ids_set_ = false;
};

// BaseTrajectory::pbc_wrap (reference lib/include/basetrajectory.h:145-161) of whole frames, on the GPU
void pbc_wrap(double *pos_aos, size_t nframes, size_t natoms, const double *box_internal, int box_stride);

}  // namespace analisi_device

#endif
