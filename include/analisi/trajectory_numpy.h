// trajectory_numpy.h -- Trajectory_numpy: a trajectory that lives in caller-provided arrays (python
// buffers), all frames resident.
//
// Mirrors the reference's lib/include/trajectory_numpy.h:7-36 / lib/src/trajectory_numpy.cpp:7-240 for
// what the g(r,t) path uses: the three box formats (Cell_vectors 3x3 per frame with the cell vectors as
// COLUMNS, Lammps_ortho 6, Lammps_triclinic 9), conversion to the internal box rows, QR rotation of a
// general cell into the LAMMPS lower-triangular frame (positions and velocities rotated with it,
// optionally keeping Q per frame), optional wrap around the cell centre, absolute frame accessors.
//
// Two constructors: the core one takes plain pointers and sizes (C++ callers, tests, the C-ABI spirit);
// when pybind11 is available (ANALISI_WITH_PYBIND11) the reference's signature taking pybind11::buffer
// objects validates the buffers with the reference's error texts and delegates to it.
//
// Differences that are this repository's design: with wrap on, the caller's arrays go to the GPUs as they are
// (pageable memory staged through page-locked slots, frames dealt to the GPUs of the process and exchanged device
// to device), are wrapped THERE, and the host only gets a wrapped copy when somebody asks for host positions
// (get_positions_copy, positions<>(), write_lammps_binary): g(r,t) never does; the per-type centre of mass arrays are computed on first use (g(r,t) never reads
// them); Lammps_triclinic input without wrap is copied (the reference leaves that buffer
// uninitialised, lib/src/trajectory_numpy.cpp:147-149 -- SURVEY.md section 8c).
#ifndef ANALISI_B200_TRAJECTORY_NUMPY_H
#define ANALISI_B200_TRAJECTORY_NUMPY_H

#include <memory>
#include <vector>

#include "analisi/basetrajectory.h"

#ifdef ANALISI_WITH_PYBIND11
#include "pybind11/pybind11.h"
#endif

class Trajectory_numpy : public BaseTrajectory<Trajectory_numpy> {
public:
    using BaseTrajectory<Trajectory_numpy>::BoxFormat;

    // pos, vel: [nts][natoms][3] float64 C-contiguous (vel may be NULL: treated as zeros);
    // types: [natoms] raw type numbers; box: [nts][3][3] | [nts][6] | [nts][9] by `format`.
    // The arrays must outlive the object when they are used in place (orthorhombic, wrap off).
    Trajectory_numpy(const double *pos, const double *vel, const int *types, const double *box, size_t nts,
                     size_t natoms, BoxFormat format = BoxFormat::Cell_vectors, bool pbc_wrap = false,
                     bool save_rotation_matrix = false);
#ifdef ANALISI_WITH_PYBIND11
    Trajectory_numpy(pybind11::buffer buffer_pos, pybind11::buffer buffer_vel, pybind11::buffer buffer_types,
                     pybind11::buffer buffer_box, BoxFormat matrix_box = BoxFormat::Cell_vectors, bool pbc_wrap = false,
                     bool save_rotation_matrix = false);
#endif
    ~Trajectory_numpy();

    template <bool SAFE = true>
    double *positions(const int &timestep, const int &atomo) {
        sync_host();
        return buffer_positions + static_cast<size_t>(natoms) * 3 * timestep + static_cast<size_t>(atomo) * 3;
    }
    template <bool SAFE = true>
    double *velocity(const int &timestep, const int &atomo) {
        if (!buffer_velocity) return nullptr;
        return buffer_velocity + static_cast<size_t>(natoms) * 3 * timestep + static_cast<size_t>(atomo) * 3;
    }
    template <bool SAFE = true>
    double *box(const int &timestep) {
        return buffer_boxes + static_cast<size_t>(timestep) * buffer_boxes_stride;
    }
    template <bool SAFE = true>
    double *positions_cm(const int &timestep, const int &tipo) {
        ensure_cm();
        return cm_pos.data() + (static_cast<size_t>(timestep) * ntypes + tipo) * 3;
    }
    template <bool SAFE = true>
    double *velocity_cm(const int &timestep, const int &tipo) {
        ensure_cm();
        return cm_vel.data() + (static_cast<size_t>(timestep) * ntypes + tipo) * 3;
    }
    double *box_last() { return buffer_boxes + (n_timesteps - 1) * buffer_boxes_stride; }
    // Q of frame t (9 doubles, column-major) or nullptr when no rotation was saved
    double *get_rotation_matrix(size_t t) { return rotation.empty() ? nullptr : rotation.data() + 9 * t; }
    // the device-resident copy of Q of frame t (this repository's addition; throws when no rotation was saved)
    void get_device_rotation_matrix(size_t t, double *q9) { download_rotation(t, q9); }

    // the wrapped frames come back from the GPUs the first time host positions are asked for
    void materialise_host_positions();

private:
    void init(const double *pos, const double *vel, const int *types, const double *box, size_t nts, size_t natoms_,
              BoxFormat format, bool wrap, bool save_rot);
    void ensure_cm();

    const double *in_pos = nullptr, *in_vel = nullptr;   // the caller's arrays (unwrapped, unrotated)
    bool wrapped_on_device = false;                        // the wrapped window lives on the GPUs; own_pos follows on demand
    analisi_device::PinnedBuffer own_pos;                  // wrapped / rotated copy (page-locked: it is uploaded)
    std::vector<double> own_vel, own_boxes, rotation, cm_pos, cm_vel, zero_vel;
    std::vector<int> raw_types, type_ids;
#ifdef ANALISI_WITH_PYBIND11
    std::vector<pybind11::buffer> keep;                    // the python objects behind in_pos / in_vel / ...
#endif
};

#endif
