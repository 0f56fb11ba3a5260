// pyanalisi -- pybind11 module with the g(r,t) part of the reference's python interface.
//
// Same class and method names as the reference's pyanalisi/src/pyanalisi.cpp for this path:
//   Trajectory  (numpy buffers, reference :487-528)      Traj  (LAMMPS binary through mmap, :375-450)
//   Gofrt / Gofrt_lammps (reference :65-82, :535-538): ctor (traj, rmin, rmax, nbin, tmax, nthreads, skip,
//       every, debug), reset, getNumberOfExtraTimestepsNeeded, calculate, buffer protocol
//       (leff, ntypes*(ntypes+1), nbin) float64
//   BoxFormat enum, info(), has_mmap(), and the common trajectory methods (:283-372).
// The calculation itself runs on the GPUs through libagofrt.so; there is no CPU path, so calculate()
// raises RuntimeError on a machine without a B200.  Types are registered module_local so the module can
// share a process with the compiled reference's own module (the parity tests import both).  Additions (not in the reference): Gofrt.counts()
// (raw integer histogram) and Gofrt.last_stats().  The GIL is released while the device job runs.
#include <cstring>
#include <string>

#include "pybind11/numpy.h"
#include "pybind11/pybind11.h"
#include "pybind11/stl.h"

#include "analisi/blockaverage.h"
#include "analisi/gofrt.h"
#include "analisi/istogrammaatomiraggio.h"
#include "analisi/msd.h"
#include "analisi/neighbour.h"
#include "analisi/sphericalbase.h"
#include "analisi/trajectory.h"
#include "analisi/trajectory_numpy.h"

namespace py = pybind11;

namespace {

template <class T>
py::array_t<T> owned_copy(const T *src, std::vector<ssize_t> shape) {
    py::array_t<T> a(shape);
    size_t n = 1;
    for (ssize_t s : shape) n *= static_cast<size_t>(s);
    if (n) std::memcpy(a.mutable_data(), src, n * sizeof(T));
    return a;
}

template <class Tk, class C>
C &common_trajectory_methods(C &c) {
    c.def("write_lammps_binary", &Tk::dump_lammps_bin_traj,
          "file name, starting timestep, end timestep (if < 0 all the trajectory)")
        .def("get_positions_copy",
             [](Tk &t) {
                 double *p = t.positions_data();
                 if (!p || t.get_nloaded_timesteps() == 0) return py::array_t<double>();
                 return owned_copy<double>(p, {static_cast<ssize_t>(t.get_nloaded_timesteps()), static_cast<ssize_t>(t.get_natoms()), 3});
             })
        .def("get_velocities_copy",
             [](Tk &t) {
                 double *p = t.velocity_data();
                 if (!p || t.get_nloaded_timesteps() == 0) return py::array_t<double>();
                 return owned_copy<double>(p, {static_cast<ssize_t>(t.get_nloaded_timesteps()), static_cast<ssize_t>(t.get_natoms()), 3});
             })
        .def("get_box_copy",
             [](Tk &t) {
                 if (t.get_nloaded_timesteps() == 0) return py::array_t<double>();
                 double *p = t.box(static_cast<int>(t.get_current_timestep()));
                 if (!p) return py::array_t<double>();
                 return owned_copy<double>(p, {static_cast<ssize_t>(t.get_nloaded_timesteps()), static_cast<ssize_t>(t.get_box_stride())});
             })
        .def("get_type", &Tk::get_type, "get the type used for the internal representation")
        .def("get_ntypes", [](Tk &t) { return t.get_ntypes(); })
        .def("get_natoms", [](Tk &t) { return t.get_natoms(); })
        .def("get_type_ids",
             [](Tk &t) {
                 py::array_t<int> a(static_cast<ssize_t>(t.get_natoms()));
                 t.get_ntypes();
                 for (size_t i = 0; i < t.get_natoms(); ++i) a.mutable_data()[i] = static_cast<int>(t.get_type(static_cast<unsigned int>(i)));
                 return a;
             })
        .def("get_nloaded_timesteps", &Tk::get_nloaded_timesteps)
        .def("getNtimesteps", &Tk::get_ntimesteps, "returns estimated number of timesteps from the file size")
        .def("get_current_timestep", &Tk::get_current_timestep,
             "return the first timestep currently loaded in this object (meaningful for the lammps binary trajectory interface)")
        .def("getWrapPbc", &Tk::get_pbc_wrap, "return the pbc wrapping of the trajectory around the center of the cell flag")
        .def("is_triclinic", &Tk::is_triclinic)
        .def("minImage", [](Tk &t, size_t i, size_t j, size_t it, size_t jt) {
            py::array_t<double> out(4);
            double *x = out.mutable_data();
            x[3] = t.d2_minImage(i, j, it, jt, x);
            return out;
        });
    return c;
}

template <class T>
void define_gofrt(py::module &m, const std::string &suffix) {
    using G = Gofrt<double, T>;
    py::class_<G>(m, ("Gofrt" + suffix).c_str(), py::buffer_protocol(), py::module_local())
        .def(py::init<T *, double, double, unsigned int, unsigned int, unsigned int, unsigned int, unsigned int, bool>(),
             py::keep_alive<1, 2>(),
             "calculates g(r) and, in general, g(r,t).  Parameters: Trajectory instance, rmin, rmax, nbin, maximum time lag, "
             "number of threads (ignored: the GPUs own the parallelism), time skip, every time, debug flag")
        .def("reset", &G::reset)
        .def("getNumberOfExtraTimestepsNeeded", &G::nExtraTimesteps)
        .def("calculate", &G::calculate, py::call_guard<py::gil_scoped_release>())
        .def("get_columns_description", &G::get_columns_description)
        .def("setReportEdges", &G::set_report_edges,
             "addition: also count the pairs within 1 ulp of a bin edge in the next calculate() (slower kernel)")
        .def("edge_pairs", &G::edge_pairs, "pairs within 1 ulp of a bin edge found by the last calculate()")
        .def("counts",
             [](G &g) {
                 const std::vector<ssize_t> sh = g.get_shape();
                 if (g.counts().empty()) return py::array_t<uint64_t>();
                 return owned_copy<uint64_t>(g.counts().data(), sh);
             },
             "raw integer bin counts of the last calculate(): the histogram before the multiplication by incr")
        .def("last_stats",
             [](G &g) {
                 const agofrt_stats &s = g.last_stats();
                 py::dict d;
                 d["kernel_ms"] = s.kernel_ms;
                 d["total_ms"] = s.total_ms;
                 d["pair_evals"] = s.pair_evals;
                 d["pair_evals_total"] = s.pair_evals_total;
                 d["jobs"] = s.jobs;
                 d["jobs_fast"] = s.jobs_fast;
                 d["launches"] = s.launches;
                 d["ndev_local"] = s.ndev_local;
                 d["world"] = s.world;
                 d["incr"] = g.get_incr();
                 return d;
             })
        .def_buffer([](G &g) -> py::buffer_info {
            return py::buffer_info(g.access_vdata(), sizeof(double), py::format_descriptor<double>::format(),
                                   static_cast<ssize_t>(g.get_shape().size()), g.get_shape(), g.get_stride());
        });
}

// Addition (the reference keeps BlockAverage on the C++ side, used by its CLI only): the block-averaged
// g(r,t) -- mean and variance of the mean over n blocks (reference lib/include/blockaverage.h:83-221,
// calcoliblocchi.h:21-65) -- for python callers and for bench.py's MediaBlocchi workload.
template <class TR>
void define_block_average(py::module &m, const std::string &suffix) {
    using G = Gofrt<double, TR>;
    using BA = BlockAverageG<TR, G, double, double, unsigned int, unsigned int, unsigned int, unsigned int, unsigned int, bool>;
    auto as_array = [](G *g) {
        return owned_copy<double>(g->access_vdata(), g->get_shape());
    };
    py::class_<BA>(m, ("GofrtBlockAverage" + suffix).c_str(), py::module_local())
        .def(py::init<TR *, unsigned int>(), py::keep_alive<1, 2>(), "trajectory, number of blocks")
        .def("calculate",
             [](BA &b, double rmin, double rmax, unsigned int nbin, unsigned int tmax, unsigned int nthreads, unsigned int skip,
                unsigned int every, bool debug) { b.calculate(rmin, rmax, nbin, tmax, nthreads, skip, every, debug); },
             py::call_guard<py::gil_scoped_release>(),
             "rmin, rmax, nbin, maximum time lag, number of threads (ignored), time skip, every time, debug flag: the arguments "
             "of Gofrt; what `analisi -g nbin -F rmin rmax -S tmax -s skip -e every -B blocks` computes")
        .def("mean", [as_array](BA &b) { return as_array(b.media()); })
        .def("variance", [as_array](BA &b) { return as_array(b.varianza()); })
        .def("block_size", &BA::block_size)
        .def("last_block", [as_array](BA &b) { return as_array(b.puntatoreCalcolo()); },
             "the last block as the calculation object holds it (puntatoreCalcolo)")
        .def("get_columns_description", [](BA &b) { return b.puntatoreCalcolo()->get_columns_description(); })
        .def("stats", [](BA &b) {
            G *g = b.puntatoreCalcolo();
            py::dict d;
            d["blocks"] = g->total_calls();
            d["kernel_ms"] = g->total_kernel_ms();
            d["device_ms"] = g->total_device_ms();
            d["pair_evals"] = g->total_pair_evals();
            d["ndev"] = g->last_stats().ndev_local;
            return d;
        });
}

// MeanSquareDisplacement / MeanSquareDisplacement_lammps (reference pyanalisi/src/pyanalisi.cpp:123-151)
template <class TR>
void define_msd(py::module &m, const std::string &suffix) {
    using M = MSD<TR>;
    py::class_<M>(m, ("MeanSquareDisplacement" + suffix).c_str(), py::buffer_protocol(), py::module_local())
        .def(py::init<TR *, unsigned int, unsigned int, unsigned int, bool, bool, bool>(), py::keep_alive<1, 2>(),
             "Trajectory instance, time skip for the average computation, max time, number of threads (ignored), calculate "
             "center of mass MSD, calculate the atomic msd in the center of mass reference system of each specie, debug flag")
        .def("reset", &M::reset)
        .def("getNumberOfExtraTimestepsNeeded", &M::nExtraTimesteps)
        .def("calculate", &M::calculate, py::call_guard<py::gil_scoped_release>())
        .def_buffer([](M &g) -> py::buffer_info {
            return py::buffer_info(g.access_vdata(), sizeof(double), py::format_descriptor<double>::format(),
                                   static_cast<ssize_t>(g.get_shape().size()), g.get_shape(), g.get_stride());
        });
}

// Addition: the neighbour-count histogram (the reference only reaches it from its CLI, --neighbour)
template <class TR>
void define_neighbour_hist(py::module &m, const std::string &suffix) {
    using H = IstogrammaAtomiRaggioG<TR>;
    py::class_<H>(m, ("NeighbourHistogram" + suffix).c_str(), py::module_local())
        .def(py::init<TR *, double, unsigned int, unsigned int>(), py::keep_alive<1, 2>(), py::arg("traj"), py::arg("r"),
             py::arg("skip") = 1, py::arg("nthreads") = 0)
        .def("reset", &H::reset)
        .def("calculate", &H::calculate, py::call_guard<py::gil_scoped_release>())
        .def("get_hist", [](H &h, unsigned int type) {
            std::map<unsigned int, unsigned int> *hist = h.get_hist();
            if (!hist) throw std::runtime_error("reset() was not called");
            return hist[type];   // dict {number of neighbours: occurrences}
        });
}

// Neighbours / Neighbours_lammps (reference pyanalisi/src/pyanalisi.cpp:193-236): cutoff neighbour lists and SANN
template <class TR>
void define_neighbours(py::module &m, const std::string &suffix) {
    using N = Neighbours<TR, double>;
    auto rows = [](typename N::template NeighIterator<typename N::TType4> it) {
        py::array_t<double> a({static_cast<ssize_t>(it.size()), static_cast<ssize_t>(4)});
        if (it.size()) std::memcpy(a.mutable_data(), it.begin(), it.size() * 4 * sizeof(double));
        return a;
    };
    auto idxs = [](typename N::template NeighIterator<size_t> it) {
        py::array_t<size_t> a(static_cast<ssize_t>(it.size()));
        if (it.size()) std::memcpy(a.mutable_data(), it.begin(), it.size() * sizeof(size_t));
        return a;
    };
    py::class_<N>(m, ("Neighbours" + suffix).c_str(), py::module_local())
        .def(py::init<TR *, typename N::ListSpec>(), py::keep_alive<1, 2>(),
             "Trajectory instance, list of (max number of neighbours, cutoff**2, skin**2), one per atomic type")
        .def("calculate_neigh", &N::update_neigh, py::call_guard<py::gil_scoped_release>())
        .def("get_sann", [rows](N &n, size_t iatom, size_t jtype) { return rows(n.get_sann_r(iatom, jtype)); })
        .def("get_sann_idx", [idxs](N &n, size_t iatom, size_t jtype) { return idxs(n.get_sann(iatom, jtype)); })
        .def("get_neigh", [rows](N &n, size_t iatom, size_t jtype) { return rows(n.get_neigh_r(iatom, jtype)); })
        .def("get_neigh_idx", [idxs](N &n, size_t iatom, size_t jtype) { return idxs(n.get_neigh(iatom, jtype)); });
}

// Addition: SphericalBase::calc of one frame (the reference reaches it through SphericalCorrelations / Steinhardt only)
template <int L, class TR>
py::tuple sh_density_l(TR &t, size_t nbin, std::vector<std::pair<double, double>> rminmax, int timestep) {
    SphericalBase<L, double, TR> sb(&t, nbin, rminmax);
    const ssize_t n = static_cast<ssize_t>(t.get_natoms()), nt = static_cast<ssize_t>(t.get_ntypes()), nl = (L + 1) * (L + 1);
    py::array_t<double> result({n, nt, static_cast<ssize_t>(nbin), nl});
    py::array_t<int> counter({n, nt, static_cast<ssize_t>(nbin)});
    sb.calc(timestep, result.mutable_data(), nullptr, nullptr, counter.mutable_data(), nullptr);
    return py::make_tuple(result, counter);
}
template <class TR>
py::tuple sh_density(TR &t, int lmax, size_t nbin, std::vector<std::pair<double, double>> rminmax, int timestep) {
    switch (lmax) {
        case 0: return sh_density_l<0>(t, nbin, rminmax, timestep);
        case 1: return sh_density_l<1>(t, nbin, rminmax, timestep);
        case 2: return sh_density_l<2>(t, nbin, rminmax, timestep);
        case 3: return sh_density_l<3>(t, nbin, rminmax, timestep);
        case 4: return sh_density_l<4>(t, nbin, rminmax, timestep);
        case 5: return sh_density_l<5>(t, nbin, rminmax, timestep);
        case 6: return sh_density_l<6>(t, nbin, rminmax, timestep);
        case 7: return sh_density_l<7>(t, nbin, rminmax, timestep);
        case 8: return sh_density_l<8>(t, nbin, rminmax, timestep);
        case 9: return sh_density_l<9>(t, nbin, rminmax, timestep);
        case 10: return sh_density_l<10>(t, nbin, rminmax, timestep);
        default: throw std::runtime_error("lmax must be in [0, 10]");
    }
}

}  // namespace

PYBIND11_MODULE(pyanalisi, m) {
    m.doc() = "B200-native g(r,t) behind the pyanalisi interface of rikigigi/analisi";

    py::class_<Trajectory> traj(m, "Traj", py::buffer_protocol(), py::module_local());
    common_trajectory_methods<Trajectory>(traj)
        .def(py::init<std::string>(), "name of the binary file to open")
        .def("setWrapPbc", &Trajectory::set_pbc_wrap, "wrap all the atomic coordinates inside the simulation box read from the binary file")
        .def("setAccessWindowSize", [](Trajectory &t, int ts) { return static_cast<int>(t.set_data_access_block_size(ts)); },
             "sets the size of the read block. Must fit in memory.  Returns 1 on success")
        .def("setAccessStart", [](Trajectory &t, int ts) { return static_cast<int>(t.set_access_at(ts)); },
             "sets the first timestep to read, and reads the full block")
        .def("setLoadVelocities", &Trajectory::set_load_velocities,
             "addition: skip velocities and centres of mass when reading windows (g(r,t) does not use them)")
        .def("get_lammps_id",
             [](Trajectory &t) {
                 int *p = t.get_lammps_id();
                 py::array_t<int> a = owned_copy<int>(p, {static_cast<ssize_t>(t.get_natoms())});
                 delete[] p;
                 return a;
             })
        .def("get_lammps_type",
             [](Trajectory &t) {
                 int *p = t.get_lammps_type();
                 py::array_t<int> a = owned_copy<int>(p, {static_cast<ssize_t>(t.get_natoms())});
                 delete[] p;
                 return a;
             })
        .def_buffer([](Trajectory &g) -> py::buffer_info {
            return py::buffer_info(g.positions_data(), sizeof(double), py::format_descriptor<double>::format(),
                                   static_cast<ssize_t>(g.get_shape().size()), g.get_shape(), g.get_stride());
        });

    py::class_<Trajectory_numpy> trn(m, "Trajectory", py::module_local());
    common_trajectory_methods<Trajectory_numpy>(trn)
        .def(py::init<py::buffer, py::buffer, py::buffer, py::buffer, Trajectory_numpy::BoxFormat, bool, bool>(),
             py::keep_alive<1, 2>(), py::keep_alive<1, 3>(), py::keep_alive<1, 4>(), py::keep_alive<1, 5>(),
             "positions (double) (ntimesteps,natoms,3); velocities (double) (ntimesteps,natoms,3); types (int) (natoms); "
             "lattice vectors (double); format of lattice vectors (BoxFormat); wrap atoms inside the cell using pbc; "
             "save rotation matrix if triclinic format is used and a rotation is needed")
        .def("get_device_rotation_matrix", [](Trajectory_numpy &t, size_t frame) {
            py::array_t<double> q({3, 3});
            t.get_device_rotation_matrix(frame, q.mutable_data());
            return q;
        }, "addition: Q of one frame as the GPUs hold it (raises when no rotation matrix was saved)")
        .def("get_rotation_matrix", [](Trajectory_numpy &t) {
            double *q = t.get_rotation_matrix(0);
            if (!q) return py::array_t<double>();
            return owned_copy<double>(q, {static_cast<ssize_t>(t.get_ntimesteps()), 3, 3});
        });

    py::enum_<Trajectory_numpy::BoxFormat>(m, "BoxFormat", py::arithmetic(), py::module_local())
        .value("Invalid", Trajectory_numpy::BoxFormat::Invalid)
        .value("CellVectors", Trajectory_numpy::BoxFormat::Cell_vectors)
        .value("LammpsOrtho", Trajectory_numpy::BoxFormat::Lammps_ortho)
        .value("LammpsTriclinic", Trajectory_numpy::BoxFormat::Lammps_triclinic);

    define_gofrt<Trajectory>(m, "_lammps");
    define_gofrt<Trajectory_numpy>(m, "");
    define_msd<Trajectory>(m, "_lammps");
    define_msd<Trajectory_numpy>(m, "");
    define_neighbour_hist<Trajectory>(m, "_lammps");
    define_neighbour_hist<Trajectory_numpy>(m, "");
    define_neighbours<Trajectory>(m, "_lammps");
    define_neighbours<Trajectory_numpy>(m, "");
    m.def("spherical_harmonic_density", &sh_density<Trajectory_numpy>, py::arg("traj"), py::arg("lmax"), py::arg("nbin"),
          py::arg("rminmax"), py::arg("timestep"),
          "SphericalBase::calc of one frame: (result [natoms][ntypes][nbin][(lmax+1)^2], counter [natoms][ntypes][nbin])");
    m.def("spherical_harmonic_density_lammps", &sh_density<Trajectory>, py::arg("traj"), py::arg("lmax"), py::arg("nbin"),
          py::arg("rminmax"), py::arg("timestep"));
    define_block_average<Trajectory>(m, "_lammps");
    define_block_average<Trajectory_numpy>(m, "");

    m.def("info", []() -> std::string { return std::string("analisi g(r,t), B200-native: ") + agofrt_version(); });
    m.def("has_mmap", []() -> bool { return true; });
    // One process per GPU (torchrun / mpirun), this repository's replacement of the reference's MPI layer (Mp): rank 0
    // calls comm_unique_id(), the launcher broadcasts the 128 bytes, every process calls comm_join(id, rank, world).
    // Set ANALISI_DEVICES to the process's GPU before the first use of the module.
    m.def("comm_unique_id", []() { return py::bytes(analisi_device::Context::comm_unique_id()); });
    m.def("comm_join", [](const py::bytes &id, int rank, int world) {
        analisi_device::Context::instance().comm_join(std::string(id), rank, world);
    });
    m.def("devices_in_use", []() { return analisi_device::Context::instance().ndev(); });
    m.def("device_count", []() {
        int n = 0;
        agofrt_device_count(&n);
        return n;
    });
}
