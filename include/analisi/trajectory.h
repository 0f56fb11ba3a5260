// trajectory.h -- Trajectory: a LAMMPS binary dump read through mmap with a sliding window.
//
// Mirrors the reference's lib/include/trajectory.h:37-139 / lib/src/trajectory.cpp:44-707 for what
// the g(r,t) path and its callers use: open + atom-id map + types from frame 0, window size
// (set_data_access_block_size), window position (set_access_at: lazy frame index, overlap reuse,
// per-atom scatter by id, box conversion, optional wrap), window-relative accessors.
// Differences that are this repository's design, not the reference's:
//   * window buffers are page-locked (agofrt_host_alloc) so the upload to the GPUs is one DMA;
//   * the wrap of freshly read frames runs on the GPU (BaseTrajectory::pbc_wrap_frames);
//   * velocities and per-type centres of mass are only read when asked for
//     (set_load_velocities; g(r,t) never touches them -- at 1M atoms they would double the I/O);
//   * one stderr summary line per type instead of one line per atom;
//   * callers that walk the file in equal steps find the next window read ahead by a background thread
//     (ANALISI_PREFETCH=0 turns it off);
//   * the frames of a window are parsed and scattered by several host threads, ids resolved through a
//     flat table when they are compact (the reference does one std::map::at per atom and frame,
//     lib/src/trajectory.cpp:633, single-threaded).
#ifndef ANALISI_B200_TRAJECTORY_H
#define ANALISI_B200_TRAJECTORY_H

#include <cstdint>
#include <exception>
#include <sstream>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "analisi/basetrajectory.h"

class Trajectory : public BaseTrajectory<Trajectory> {
public:
    explicit Trajectory(std::string filename);
    ~Trajectory();

    template <bool SAFE = true>
    double *positions(const size_t &timestep, const size_t &atomo) {
        double *p = window_ptr<SAFE, true>(buffer_positions, timestep, atomo, 3 * natoms, 3);
        sync_host();   // a window parsed on the GPUs reaches the host the first time somebody asks for it
        return p;
    }
    template <bool SAFE = true>
    double *velocity(const size_t &timestep, const size_t &atomo) {
        if (!buffer_velocity) return nullptr;
        return window_ptr<SAFE, true>(buffer_velocity, timestep, atomo, 3 * natoms, 3);
    }
    template <bool SAFE = true>
    double *box(const size_t &timestep) {
        return window_ptr<SAFE, false>(buffer_boxes, timestep, 0, buffer_boxes_stride, 0);
    }
    template <bool SAFE = true>
    double *positions_cm(const size_t &timestep, const size_t &tipo) {
        if (!cm_pos.data()) return nullptr;
        return window_ptr<SAFE, false>(cm_pos.data(), timestep, tipo, 3 * ntypes, 3);
    }
    template <bool SAFE = true>
    double *velocity_cm(const size_t &timestep, const size_t &tipo) {
        if (!cm_vel.data()) return nullptr;
        return window_ptr<SAFE, false>(cm_vel.data(), timestep, tipo, 3 * ntypes, 3);
    }
    double *box_last() {
        if (loaded_timesteps > 0 && buffer_boxes) return buffer_boxes + buffer_boxes_stride * (loaded_timesteps - 1);
        throw std::runtime_error("No data is loaded!\n");
    }

    using BaseTrajectory<Trajectory>::Errori;
    Errori set_data_access_block_size(const size_t &timesteps);
    Errori set_access_at(const size_t &timestep);
    int64_t get_timestep_lammps(size_t timestep);
    void index_all();
    int *get_lammps_id();     // new int[natoms], caller owns (as in the reference)
    int *get_lammps_type();   // new int[natoms], caller owns

    // default true (what the reference always does); the CLI g(r,t) branch turns it off
    void set_load_velocities(bool v) { load_velocities = v; }
    // the caller will ask for windows `stride` frames apart (BlockAverageG does): read ahead from the first one
    void set_access_stride_hint(size_t stride) { stride_hint = stride; }
    // the window lives on the GPUs (parsed there from the raw records); fetch it for the host accessors
    void materialise_host_positions() { download_window(pos_buf.data()); }

private:
    template <bool SAFE, bool ATOM>
    double *window_ptr(double *base, const size_t &timestep, const size_t &atomo, const size_t &stride1,
                       const size_t &stride2) {
        if constexpr (SAFE) {
            if constexpr (ATOM) {
                if (atomo >= static_cast<size_t>(natoms)) {
                    std::stringstream ss;
                    ss << "Requested atom index (" << atomo << ") is not in the range [0, " << natoms - 1 << "]\n";
                    throw std::runtime_error(ss.str());
                }
            }
            const bool inside = window_loaded && timestep >= static_cast<size_t>(current_timestep) &&
                                timestep - current_timestep < static_cast<size_t>(loaded_timesteps);
            if (!inside) {
                // like the reference: an access outside the window moves the window there
                if (!set_access_at(timestep)) throw std::runtime_error("Error loading the file\n");
                if (timestep < static_cast<size_t>(current_timestep) ||
                    timestep - current_timestep >= static_cast<size_t>(loaded_timesteps))
                    throw std::runtime_error("requested timestep is out of range");
            }
        }
        return base + (timestep - current_timestep) * stride1 + atomo * stride2;
    }

    size_t frame_bytes(size_t offset, LammpsFrameHeader &head, std::vector<LammpsChunk> *chunks);
    void ensure_indexed(size_t upto);
    void read_frame_into_slot(size_t frame, size_t slot);
    void read_frame_to(size_t frame, double *P, double *box_row, double *V, double *cm_p, double *cm_v);
    void read_frames(size_t first, size_t last, size_t origin, double *P0, double *B0, bool own_window);
    void start_prefetch(size_t target);
    void cancel_prefetch();
    // Device-side ingest: the records of a window go to the GPUs as they are in the file and are parsed there (id -> slot,
    // scatter, box rows from the headers, wrap).  Possible when positions are all that is wanted (no velocities, no
    // centres of mass) and the ids are compact enough for a flat table.  ANALISI_DEVICE_PARSE=0 keeps the host parser.
    bool device_parse_possible() const;
    struct RecordTable {
        std::vector<const void *> ptr;
        std::vector<int> atoms;
        std::vector<size_t> frame_chunk;
    };
    void gather_records(size_t first, size_t n, RecordTable &tab, double *B0);

    int fd = -1;
    char *file = nullptr;
    size_t fsize = 0;
    std::vector<size_t> offsets;          // byte offset of frame k, valid for k <= indexed_upto
    std::vector<int64_t> lammps_steps;    // LAMMPS timestep of frame k, valid once the frame was parsed
    size_t indexed_upto = 0;
    bool window_loaded = false;
    bool load_velocities = true;
    std::mutex type_change_mutex;   // the type-change warning of read_frame_to (several reader threads)
    std::unordered_map<int, int> id_to_slot;
    std::vector<int> slot_to_id;
    std::vector<int> dense_slot;          // id -> slot when the ids are compact (else id_to_slot is used)
    std::vector<int> raw_type, type_id;
    analisi_device::PinnedBuffer pos_buf, vel_buf;
    std::vector<double> boxes, cm_pos, cm_vel;
    size_t window_capacity = 0;
    // read-ahead of the next window (see trajectory.cpp: start_prefetch)
    analisi_device::PinnedBuffer pos_alt;
    std::vector<double> boxes_alt;
    std::thread prefetch_thread;
    std::exception_ptr prefetch_error;
    size_t prefetch_target = 0, stride_hint = 0;
    bool prefetch_valid = false, prefetch_enabled = true, prefetch_uploaded = false, prefetch_records = false;
};

#endif
