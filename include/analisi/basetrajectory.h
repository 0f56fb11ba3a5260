// basetrajectory.h -- BaseTrajectory<T>: what every trajectory container offers to the calculations.
//
// Mirrors the public surface of the reference's lib/include/basetrajectory.h:45-331 that the g(r,t)
// path and its callers use (same names, argument meaning and error behaviour), on top of a
// DEVICE-RESIDENT copy of the loaded window:
//
//   * the host window (positions [frame][atom][3], one internal box row per frame
//     [xlo,ylo,zlo,lx/2,ly/2,lz/2(,xy,xz,yz)], dense type ids) is what the accessors
//     positions<>() / box<>() / get_type() serve, exactly as in the reference;
//   * device_window() hands Gofrt the agofrt_traj that holds the same window on every GPU of the
//     process (coalesced SoA float64), uploading it when the host window
//     changed since the last upload;
//   * the arithmetic of BaseTrajectory -- pbc_wrap (reference :145-161) and d2_minImage
//     (reference :168-268) -- runs on the GPU through the C ABI (agofrt_pbc_wrap,
//     agofrt_traj_d2_pair); there is no host implementation of the minimum image here.
#ifndef ANALISI_B200_BASETRAJECTORY_H
#define ANALISI_B200_BASETRAJECTORY_H

#include <sys/types.h>

#include <algorithm>
#include <cstddef>
#include <fstream>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "analisi/device.h"
#include "analisi/lammps_struct.h"

template <class T>
class BaseTrajectory {
public:
    enum class BoxFormat { Lammps_ortho, Lammps_triclinic, Cell_vectors, Invalid };
    enum Errori { non_inizializzato = 0, oltre_fine_file = 2, Ok = 1 };

    BaseTrajectory() = default;
    BaseTrajectory(const BaseTrajectory &) = delete;
    BaseTrajectory &operator=(const BaseTrajectory &) = delete;

    // ---- accessors served by the derived container (CRTP, as in the reference :63-68) ----
    template <bool SAFE = true>
    double *positions(const int &timestep, const int &atomo) {
        return static_cast<T *>(this)->template positions<SAFE>(timestep, atomo);
    }
    template <bool SAFE = true>
    double *velocity(const int &timestep, const int &atomo) {
        return static_cast<T *>(this)->template velocity<SAFE>(timestep, atomo);
    }
    template <bool SAFE = true>
    double *box(const int &timestep) {
        return static_cast<T *>(this)->template box<SAFE>(timestep);
    }
    double *box_last() { return static_cast<T *>(this)->box_last(); }

    std::vector<unsigned int> get_types() {
        get_ntypes();
        return types;
    }
    unsigned int get_type(const unsigned int &atomo) {
        if (atomo < static_cast<size_t>(natoms)) return buffer_type_id[atomo];
        throw std::runtime_error("Atom index out of range\n");
    }
    Errori set_data_access_block_size(const size_t &) {
        std::cerr << "Warning: doing nothing (not reading in blocks)" << std::endl;
        return Errori::Ok;
    }
    Errori set_access_at(const size_t &) {
        std::cerr << "Warning: doing nothing (not reading in blocks)" << std::endl;
        return Errori::Ok;
    }

    // must be set before the window is loaded
    void set_pbc_wrap(bool p) { wrap_pbc = p; }
    bool get_pbc_wrap() const { return wrap_pbc; }

    // [xlo,xhi,ylo,yhi,zlo,zhi] -> [xlo,ylo,zlo,(xhi-xlo)/2,(yhi-ylo)/2,(zhi-zlo)/2]  (reference :94-105)
    static void lammps_to_internal(double *c) {
        const double xlo = c[0], xhi = c[1], ylo = c[2], yhi = c[3], zlo = c[4], zhi = c[5];
        c[0] = xlo;
        c[1] = ylo;
        c[2] = zlo;
        c[3] = (xhi - xlo) / 2;
        c[4] = (yhi - ylo) / 2;
        c[5] = (zhi - zlo) / 2;
    }
    // and back (reference :109-120): hi = half*2 + lo
    static void internal_to_lammps(double *c) {
        const double xlo = c[0], ylo = c[1], zlo = c[2], hx = c[3], hy = c[4], hz = c[5];
        c[0] = xlo;
        c[1] = xlo + hx * 2;
        c[2] = ylo;
        c[3] = hy * 2 + ylo;
        c[4] = zlo;
        c[5] = hz * 2 + zlo;
    }

    double *positions_data() {
        sync_host();
        return buffer_positions;
    }
    double *velocity_data() { return buffer_velocity; }
    int get_type_min() { return min_type; }
    int get_type_max() { return max_type; }
    size_t get_natoms() const { return natoms; }
    size_t get_ntimesteps() const { return n_timesteps; }
    ssize_t get_current_timestep() const { return current_timestep; }
    size_t get_nloaded_timesteps() const { return loaded_timesteps; }
    bool is_triclinic() const { return triclinic; }
    size_t get_box_stride() const { return buffer_boxes_stride; }

    std::vector<ssize_t> get_shape() {
        return {static_cast<ssize_t>(loaded_timesteps), static_cast<ssize_t>(natoms), 3};
    }
    std::vector<ssize_t> get_stride() {
        return {static_cast<ssize_t>(natoms * 3 * sizeof(double)), static_cast<ssize_t>(3 * sizeof(double)),
                static_cast<ssize_t>(sizeof(double))};
    }

    // Wrap the atoms of loaded frame `idx` (window-relative) around the centre of the orthorhombic
    // cell: x -= l_half; minimum image; x += l_half (reference :145-161).  Runs on the GPU.
    template <bool TRICLINIC>
    void pbc_wrap(ssize_t idx) {
        pbc_wrap_frames(idx, 1);
    }
    void pbc_wrap_frames(ssize_t first_idx, size_t nframes) {
        if (nframes == 0 || natoms == 0) return;
        sync_host();
        analisi_device::pbc_wrap(buffer_positions + first_idx * natoms * 3, nframes, natoms,
                                 buffer_boxes + first_idx * buffer_boxes_stride, static_cast<int>(buffer_boxes_stride));
        mark_window_changed();
    }

    // Squared minimum-image distance between atom i at itimestep and atom j at jtimestep with the
    // box of itimestep; x receives the minimum-image x(i)-x(j) (reference :168-194).  One pair is
    // evaluated by the GPU on the device copy of the window -- a probe, not a hot path.
    double d2_minImage(size_t i, size_t j, size_t itimestep, size_t jtimestep) {
        double x[3];
        return d2_minImage(i, j, itimestep, jtimestep, x);
    }
    double d2_minImage(size_t i, size_t j, size_t itimestep, size_t jtimestep, double *x) {
        double out[4];
        analisi_device::check(agofrt_traj_d2_pair(device_window(), i, j, itimestep, jtimestep, out), "agofrt_traj_d2_pair");
        x[0] = out[0];
        x[1] = out[1];
        x[2] = out[2];
        return out[3];
    }

    // Dense type ids: sorted distinct raw types -> 0..ntypes-1 (reference lib/src/basetrajectory.cpp:51-89)
    size_t get_ntypes() {
        if (ntypes == 0 && natoms > 0) {
            types.assign(buffer_type, buffer_type + natoms);
            std::sort(types.begin(), types.end());
            types.erase(std::unique(types.begin(), types.end()), types.end());
            min_type = static_cast<int>(types.front());
            max_type = static_cast<int>(types.back());
            type_map.clear();
            for (unsigned int k = 0; k < types.size(); ++k) type_map[static_cast<int>(types[k])] = k;
            for (ssize_t i = 0; i < natoms; ++i) buffer_type_id[i] = static_cast<int>(type_map.at(buffer_type[i]));
            ntypes = static_cast<ssize_t>(types.size());
        }
        return ntypes;
    }

    // Write frames [start_ts, stop_ts) as a pre-2020 LAMMPS binary with one chunk per frame, atom id =
    // internal index, type = dense type id (reference lib/src/basetrajectory.cpp:4-49).
    void dump_lammps_bin_traj(const std::string &fname, int start_ts, int stop_ts) {
        if (start_ts < 0 || start_ts >= n_timesteps)
            throw std::runtime_error("You must provide a starting timestep between 0 and the number of timesteps!");
        if (stop_ts <= 0) stop_ts = static_cast<int>(n_timesteps);
        std::ofstream out(fname, std::ofstream::binary);
        std::vector<double> rows(static_cast<size_t>(natoms) * kLammpsDoublesPerAtom);
        for (int t = start_ts; t < stop_ts; ++t) {
            LammpsFrameHeader head;
            head.timestep = t;
            head.natoms = natoms;
            head.triclinic = triclinic;
            const double *b = box(t);
            for (int k = 0; k < 6; ++k) head.box[k] = b[k];
            internal_to_lammps(head.box);
            if (triclinic)
                for (int k = 0; k < 3; ++k) head.xy_xz_yz[k] = b[6 + k];
            head.size_one = kLammpsDoublesPerAtom;
            head.nchunk = 1;
            head.write(out);
            const int n_data = static_cast<int>(natoms * kLammpsDoublesPerAtom);
            out.write(reinterpret_cast<const char *>(&n_data), sizeof(int));
            for (ssize_t a = 0; a < natoms; ++a) {
                double *r = &rows[a * kLammpsDoublesPerAtom];
                r[0] = static_cast<double>(a);
                r[1] = get_type(static_cast<unsigned int>(a));
                const double *p = positions(t, static_cast<int>(a)), *v = velocity(t, static_cast<int>(a));
                for (int k = 0; k < 3; ++k) {
                    r[2 + k] = p[k];
                    r[5 + k] = v ? v[k] : 0.0;
                }
            }
            out.write(reinterpret_cast<const char *>(rows.data()), rows.size() * sizeof(double));
        }
    }

    // ---- the device copy of the loaded window (this repository's addition) ----
    // Frames [current_timestep, current_timestep + loaded_timesteps) on every GPU of the process.
    agofrt_traj *device_window() {
        if (loaded_timesteps <= 0 || (!buffer_positions && host_current) || !buffer_boxes)
            throw std::runtime_error("No data is loaded!\n");
        get_ntypes();
        if (!dev_window.valid() || dev_window.capacity() < static_cast<size_t>(loaded_timesteps)) {
            dev_window.create(natoms, static_cast<int>(buffer_boxes_stride), buffer_type_id, static_cast<int>(ntypes),
                              loaded_timesteps);
            dev_uploaded_epoch = 0;
            ++dev_epoch;
        }
        if (dev_uploaded_epoch != host_epoch) {
            // once per box: the frames are dealt to the GPUs and exchanged device to device
            dev_window.upload_shared(current_timestep, loaded_timesteps, buffer_positions, buffer_boxes, false, nullptr);
            dev_uploaded_epoch = host_epoch;
        }
        return dev_window.handle();
    }
    // changes whenever a device handle is created or destroyed (Gofrt rebuilds its plan then; between such
    // events it only re-points the plan at the handle device_window() returns: the windows are double-buffered)
    uint64_t device_generation() const { return dev_epoch; }

    // The host copy of the positions may be BEHIND the device copy (a window that was wrapped on the GPUs straight
    // from the caller's arrays: g(r,t) never reads host positions).  Everything that hands out host positions calls
    // this first; the frames then come back from the device once.
    void sync_host() {
        if (!host_current) {
            static_cast<T *>(this)->materialise_host_positions();
            host_current = true;
        }
    }

protected:
    ~BaseTrajectory() = default;
    void mark_window_changed() { ++host_epoch; }
    void materialise_host_positions() {}   // containers whose host copy is always current
    // the window [current_timestep, +loaded_timesteps) goes to the GPUs straight from `pos` (read only, may be
    // pageable), wrapped there when asked; the host copy is left for sync_host() to fetch
    void upload_now(const double *pos, bool wrap) {
        get_ntypes();
        if (!dev_window.valid() || dev_window.capacity() < static_cast<size_t>(loaded_timesteps)) {
            dev_window.create(natoms, static_cast<int>(buffer_boxes_stride), buffer_type_id, static_cast<int>(ntypes),
                              loaded_timesteps);
            ++dev_epoch;
        }
        dev_window.upload_shared(current_timestep, loaded_timesteps, pos, buffer_boxes, wrap, nullptr);
        ++host_epoch;
        dev_uploaded_epoch = host_epoch;
        host_current = false;
    }
    void download_window(double *pos_out) { dev_window.download(current_timestep, loaded_timesteps, pos_out); }
    // rotation matrices of the loaded frames, device-resident next to positions and cells
    void upload_rotation(const double *q9) {
        if (dev_window.valid()) dev_window.set_rotation(current_timestep, loaded_timesteps, q9);
    }
    void download_rotation(size_t frame, double *q9) { dev_window.get_rotation(frame, q9); }
    // the same from the raw records of a LAMMPS dump (parsed on the GPUs); false: an atom changed type, nothing uploaded
    bool upload_records_now(const void *const *chunk_ptr, const int *chunk_atoms, const size_t *frame_chunk, bool wrap,
                            const int *slot_to_id, const int *slot_raw_type) {
        get_ntypes();
        if (!dev_window.valid() || dev_window.capacity() < static_cast<size_t>(loaded_timesteps)) {
            dev_window.create(natoms, static_cast<int>(buffer_boxes_stride), buffer_type_id, static_cast<int>(ntypes),
                              loaded_timesteps);
            ++dev_epoch;
        }
        if (!dev_window.ids_set()) dev_window.set_ids(slot_to_id, slot_raw_type);
        if (!dev_window.upload_records(current_timestep, loaded_timesteps, chunk_ptr, chunk_atoms, frame_chunk, buffer_boxes, wrap))
            return false;
        ++host_epoch;
        dev_uploaded_epoch = host_epoch;
        host_current = false;
        return true;
    }
    bool upload_next_window_records(size_t first, size_t n, const void *const *chunk_ptr, const int *chunk_atoms,
                                    const size_t *frame_chunk, const double *box_internal, bool wrap, const int *slot_to_id,
                                    const int *slot_raw_type) {
        if (!dev_window_next.valid() || dev_window_next.capacity() < n) {
            dev_window_next.create(natoms, static_cast<int>(buffer_boxes_stride), buffer_type_id, static_cast<int>(ntypes),
                                   std::max(n, dev_window.capacity()));
            next_created = true;
        }
        if (!dev_window_next.ids_set()) dev_window_next.set_ids(slot_to_id, slot_raw_type);
        return dev_window_next.upload_records(first, n, chunk_ptr, chunk_atoms, frame_chunk, box_internal, wrap);
    }
    bool host_current = true;

    // ---- read-ahead of the NEXT window onto the devices (derived containers with a background reader) ----
    // Gofrt has already used this trajectory on the GPUs: reading ahead may upload as well
    bool device_active() const { return dev_window.valid(); }
    // Called on the reader thread while the caller's thread may be inside Gofrt::calculate on the current
    // window: frames [first, first+n) go to the SECOND device window (wrapped there when asked; pos comes
    // back wrapped).  The C ABI allows exactly this concurrency (include/agofrt.h).
    void upload_next_window(size_t first, size_t n, double *pos_aos_inout, const double *box_internal, bool wrap) {
        if (!dev_window_next.valid() || dev_window_next.capacity() < n) {
            dev_window_next.create(natoms, static_cast<int>(buffer_boxes_stride), buffer_type_id, static_cast<int>(ntypes),
                                   std::max(n, dev_window.capacity()));
            next_created = true;
        }
        if (wrap)
            dev_window_next.upload_wrap(first, n, pos_aos_inout, box_internal);
        else
            dev_window_next.upload(first, n, pos_aos_inout, box_internal);
    }
    // Called on the caller's thread after the host buffers were swapped and mark_window_changed(): the device
    // already holds this window
    void adopt_next_window(bool host_copy_is_current = true) {
        dev_window.swap(dev_window_next);
        dev_uploaded_epoch = host_epoch;
        host_current = host_copy_is_current;
        if (next_created) {
            ++dev_epoch;
            next_created = false;
        }
    }

    double *buffer_positions = nullptr;
    double *buffer_velocity = nullptr;
    double *buffer_boxes = nullptr;      // internal format, buffer_boxes_stride doubles per frame
    size_t buffer_boxes_stride = 6;      // 6 orthorhombic, 9 triclinic
    BoxFormat box_format = BoxFormat::Invalid;
    int *buffer_type = nullptr;          // raw types
    int *buffer_type_id = nullptr;       // dense ids
    ssize_t natoms = 0, ntypes = 0, n_timesteps = 0, loaded_timesteps = 0, current_timestep = 0;
    int min_type = 0, max_type = 0;
    bool wrap_pbc = true, triclinic = false;
    std::vector<unsigned int> types;
    std::map<int, unsigned int> type_map;

private:
    analisi_device::Window dev_window, dev_window_next;
    uint64_t host_epoch = 1, dev_uploaded_epoch = 0, dev_epoch = 0;
    bool next_created = false;
};

#endif
