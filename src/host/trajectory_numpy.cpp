// trajectory_numpy.cpp -- see include/analisi/trajectory_numpy.h
#include "analisi/trajectory_numpy.h"

#include <cstring>

#include "analisi/triclinic.h"

Trajectory_numpy::Trajectory_numpy(const double *pos, const double *vel, const int *types, const double *box,
                                   size_t nts, size_t natoms_, BoxFormat format, bool wrap, bool save_rot) {
    init(pos, vel, types, box, nts, natoms_, format, wrap, save_rot);
}

void Trajectory_numpy::init(const double *pos, const double *vel, const int *types, const double *box, size_t nts,
                            size_t natoms_, BoxFormat format, bool wrap, bool save_rot) {
    if (format != BoxFormat::Cell_vectors && format != BoxFormat::Lammps_ortho && format != BoxFormat::Lammps_triclinic)
        throw std::runtime_error("Invalid input cell format");
    loaded_timesteps = 0;
    current_timestep = 0;
    wrap_pbc = wrap;
    natoms = static_cast<ssize_t>(natoms_);
    n_timesteps = static_cast<ssize_t>(nts);
    in_pos = pos;
    in_vel = vel;
    const size_t nval = nts * natoms_ * 3;

    // ---- boxes -> internal rows; for general cell matrices one QR per run of identical cells ----
    struct CellRun {
        size_t first;
        TriclinicLammpsCell<double> cell;
    };
    std::vector<CellRun> runs;
    triclinic = false;
    if (format == BoxFormat::Cell_vectors) {
        for (size_t f = 0; f < nts; ++f) {
            const double *m = box + 9 * f;
            if (runs.empty() || !runs.back().cell.is_same_cell(m)) {
                runs.push_back({f, TriclinicLammpsCell<double>(m)});
                if (!runs.back().cell.isDiagonal()) triclinic = true;
            }
        }
        buffer_boxes_stride = triclinic ? 9 : 6;
        own_boxes.assign(nts * buffer_boxes_stride, 0.0);
        if (triclinic && save_rot) rotation.assign(nts * 9, 0.0);
        size_t r = 0;
        for (size_t f = 0; f < nts; ++f) {
            while (r + 1 < runs.size() && runs[r + 1].first <= f) ++r;
            runs[r].cell.set_lammps_cell(own_boxes.data() + f * buffer_boxes_stride, triclinic);
            if (!rotation.empty()) runs[r].cell.getQ(rotation.data() + 9 * f);
        }
    } else {
        triclinic = format == BoxFormat::Lammps_triclinic;
        buffer_boxes_stride = triclinic ? 9 : 6;
        own_boxes.assign(box, box + nts * buffer_boxes_stride);
        for (size_t f = 0; f < nts; ++f) lammps_to_internal(own_boxes.data() + f * buffer_boxes_stride);
        std::cerr << "Input format is lammps" << std::endl;
    }
    buffer_boxes = own_boxes.data();
    box_format = triclinic ? BoxFormat::Lammps_triclinic : BoxFormat::Lammps_ortho;

    // ---- positions / velocities: used in place unless they must be rotated (host copy) or wrapped (on the GPUs) ----
    const bool rotate = triclinic && format == BoxFormat::Cell_vectors;
    if (rotate) std::cerr << "Detected non orthorombic simulation cell. Using triclinic format" << std::endl;
    if (rotate) {
        own_pos.resize(nval > 0 ? nval : 1);
        if (nval) std::memcpy(own_pos.data(), pos, nval * sizeof(double));
        buffer_positions = own_pos.data();
    } else {
        buffer_positions = const_cast<double *>(pos);
    }
    if (!vel) {
        zero_vel.assign(nval, 0.0);
        in_vel = zero_vel.data();
    }
    if (rotate) {
        own_vel.assign(in_vel, in_vel + nval);
        buffer_velocity = own_vel.data();
        size_t r = 0;
        for (size_t f = 0; f < nts; ++f) {
            while (r + 1 < runs.size() && runs[r + 1].first <= f) ++r;
            const TriclinicLammpsCell<double> &c = runs[r].cell;
            double *p = buffer_positions + f * natoms_ * 3, *v = buffer_velocity + f * natoms_ * 3;
            for (size_t a = 0; a < natoms_; ++a) {
                c.rotate_vec(p + 3 * a);
                c.rotate_vec(v + 3 * a);
            }
        }
    } else {
        buffer_velocity = const_cast<double *>(in_vel);
    }

    // ---- types ----
    raw_types.assign(types, types + natoms_);
    type_ids.assign(natoms_, 0);
    buffer_type = raw_types.data();
    buffer_type_id = type_ids.data();
    get_ntypes();

    loaded_timesteps = n_timesteps;
    if (wrap && nts > 0 && natoms_ > 0) {
        // wrapped on the GPUs, straight from the (rotated copy of the) caller's array; the host copy follows on demand
        upload_now(buffer_positions, true);
        if (!rotate) buffer_positions = nullptr;   // the caller's array stays unwrapped: never hand it out as the window
        wrapped_on_device = true;
        if (!rotation.empty()) upload_rotation(rotation.data());   // Q per frame, next to positions and cells on the GPUs
    } else {
        mark_window_changed();
    }
}

void Trajectory_numpy::materialise_host_positions() {
    if (!wrapped_on_device) return;
    const size_t nval = static_cast<size_t>(n_timesteps) * natoms * 3;
    if (own_pos.size() < nval) own_pos.resize(nval > 0 ? nval : 1);
    download_window(own_pos.data());
    buffer_positions = own_pos.data();
}

// per-type centre of mass of the caller's (unwrapped, unrotated) arrays: running mean in atom order,
// as the reference does (lib/src/trajectory_numpy.cpp:201-223)
void Trajectory_numpy::ensure_cm() {
    if (!cm_pos.empty() || n_timesteps == 0 || ntypes == 0) return;
    auto run = [&](const double *a, std::vector<double> &cm) {
        cm.assign(static_cast<size_t>(n_timesteps) * ntypes * 3, 0.0);
        std::vector<int> cnt(ntypes);
        for (ssize_t f = 0; f < n_timesteps; ++f) {
            std::fill(cnt.begin(), cnt.end(), 0);
            double *c = cm.data() + static_cast<size_t>(f) * ntypes * 3;
            for (ssize_t i = 0; i < natoms; ++i) {
                const int k = type_ids[i];
                cnt[k]++;
                for (int d = 0; d < 3; ++d) c[3 * k + d] += (a[(f * natoms + i) * 3 + d] - c[3 * k + d]) / double(cnt[k]);
            }
        }
    };
    run(in_pos, cm_pos);
    run(in_vel, cm_vel);
}

Trajectory_numpy::~Trajectory_numpy() { buffer_positions = buffer_velocity = buffer_boxes = nullptr; }

#ifdef ANALISI_WITH_PYBIND11
namespace {
// C-contiguous, no padding between items (reference lib/include/buffer_utils.h:6-15)
template <class T>
bool dense(const pybind11::buffer_info &b) {
    ssize_t expect = sizeof(T);
    for (int d = static_cast<int>(b.ndim) - 1; d >= 0; --d) {
        if (b.strides[d] != expect) return false;
        expect *= b.shape[d];
    }
    return true;
}
}  // namespace

Trajectory_numpy::Trajectory_numpy(pybind11::buffer buffer_pos, pybind11::buffer buffer_vel, pybind11::buffer buffer_types,
                                   pybind11::buffer buffer_box, BoxFormat matrix_box, bool wrap, bool save_rot) {
    namespace py = pybind11;
    py::buffer_info p = buffer_pos.request(), v = buffer_vel.request(), t = buffer_types.request(), b = buffer_box.request();
    if (p.ndim != 3) throw std::runtime_error("Wrong number of dimension of position array (must be 3)");
    if (v.ndim != 3) throw std::runtime_error("Wrong number of dimension of velocities array (must be 3)");
    if (p.shape[2] != 3) throw std::runtime_error("Wrong number of cartesian components in the third dimension of positions array");
    if (v.shape[2] != 3) throw std::runtime_error("Wrong number of cartesian components in the third dimension of velocities array");
    for (int d = 0; d < 3; ++d)
        if (p.shape[d] != v.shape[d]) throw std::runtime_error("Shape of positions and velocities array is different");
    if (t.ndim != 1) throw std::runtime_error("Wrong number of dimension of types array (must be 1)");
    if (t.shape[0] != p.shape[1]) throw std::runtime_error("Wrong size of the type array");
    if (matrix_box == BoxFormat::Cell_vectors) {
        if (b.ndim != 3) throw std::runtime_error("Wrong number of dimensions of box array (must be 3) for cell matrix format");
        if (b.shape[0] != p.shape[0] || b.shape[1] != 3 || b.shape[2] != 3)
            throw std::runtime_error("Wrong shape of box array: must be (nsteps, 3, 3) for matrix format");
    } else {
        if (b.ndim != 2)
            throw std::runtime_error("Wrong number of dimensions of box array (must be 2 for lammps ortho/triclinic cell format)");
        if (b.shape[0] != p.shape[0]) throw std::runtime_error("Wrong shape of box array: first dimension must be nsteps");
        if (matrix_box == BoxFormat::Lammps_ortho && b.shape[1] != 6)
            throw std::runtime_error("Wrong shape of box array: must be (:, 6) for orthogonal lammps cell format");
        if (matrix_box == BoxFormat::Lammps_triclinic && b.shape[1] != 9)
            throw std::runtime_error("Wrong shape of box array: must be (:, 9) for triclinic lammps cell format");
    }
    if (b.format != py::format_descriptor<double>::format()) throw std::runtime_error("Format of box array should be double");
    if (t.format != py::format_descriptor<int>::format())
        throw std::runtime_error("Format of types array should be int (" + py::format_descriptor<int>::format() +
                                 ") while it was " + t.format);
    if (v.format != py::format_descriptor<double>::format()) throw std::runtime_error("Format of velocities array should be double");
    if (p.format != py::format_descriptor<double>::format()) throw std::runtime_error("Format of positions array should be double");
    if (!dense<double>(b)) throw std::runtime_error("Unsupported stride in box array");
    if (!dense<int>(t)) throw std::runtime_error("Unsupported stride in types array");
    if (!dense<double>(v)) throw std::runtime_error("Unsupported stride in vel array");
    if (!dense<double>(p)) throw std::runtime_error("Unsupported stride in pos array");
    keep = {buffer_pos, buffer_vel, buffer_types, buffer_box};
    init(static_cast<const double *>(p.ptr), static_cast<const double *>(v.ptr), static_cast<const int *>(t.ptr),
         static_cast<const double *>(b.ptr), static_cast<size_t>(p.shape[0]), static_cast<size_t>(p.shape[1]), matrix_box,
         wrap, save_rot);
}
#endif
