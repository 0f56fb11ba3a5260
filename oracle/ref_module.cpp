// TEST INFRASTRUCTURE ONLY -- never imported by the product path.
//
// pybind11 harness around the UNMODIFIED reference sources.  The Makefile in this
// directory compiles this file together with the reference's own translation
// units, taken where they lie under /root/reference:
//     lib/src/gofrt.cpp  lib/src/basetrajectory.cpp  lib/src/trajectory.cpp
//     lib/src/trajectory_numpy.cpp  lib/src/cronometro.C  lib/src/neighbour.cpp  lib/src/sphericalbase.cpp
// into oracle/_ref/analisi_ref*.so.  Nothing of the reference is copied into
// this repository: this file only *includes* its headers and exposes them to
// the tests with the same class names the reference's own module uses
// (pyanalisi/src/pyanalisi.cpp:65-82, :375-550), plus a few probes
// (d2_min_image, block_average_gofrt) that the reference only reaches from C++.
//
// The module is the live oracle for everything the reference's golden files do
// not pin (triclinic min-image, every>1, ragged skip, block averages).

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pybind11/pybind11.h"
#include "pybind11/numpy.h"
#include "pybind11/stl.h"

#include "config.h"
#include "trajectory.h"
#include "trajectory_numpy.h"
#include "gofrt.h"
#include "blockaverage.h"
#include "neighbour.h"
#include "sphericalbase.h"

namespace py = pybind11;

// trajectory.cpp only uses FFTW as an aligned allocator (trajectory.cpp:362-371);
// the vendored FFTW is not built here, so provide the two symbols it needs.
extern "C" void *fftw_malloc(size_t n) {
    void *p = nullptr;
    if (n == 0) n = 64;
    if (posix_memalign(&p, 64, n) != 0) return nullptr;
    return p;
}
extern "C" void fftw_free(void *p) { free(p); }

namespace {

template <class T>
py::array_t<double> copy_positions(T &t) {
    const long nts = t.get_nloaded_timesteps();
    const long nat = t.get_natoms();
    py::array_t<double> out({nts, nat, 3L});
    std::memcpy(out.mutable_data(), t.positions(t.get_current_timestep(), 0),
                sizeof(double) * nts * nat * 3);
    return out;
}

template <class T>
py::array_t<double> copy_boxes(T &t) {
    const long nts = t.get_nloaded_timesteps();
    const long bs = t.get_box_stride();
    py::array_t<double> out({nts, bs});
    std::memcpy(out.mutable_data(), t.box(t.get_current_timestep()), sizeof(double) * nts * bs);
    return out;
}

template <class T>
py::array_t<int> copy_type_ids(T &t) {
    const long nat = t.get_natoms();
    t.get_ntypes();
    py::array_t<int> out(nat);
    for (long i = 0; i < nat; ++i) out.mutable_data()[i] = (int)t.get_type(i);
    return out;
}

template <class T>
py::tuple probe_d2(T &t, size_t i, size_t j, size_t it, size_t jt) {
    double x[3];
    double d2 = t.d2_minImage(i, j, it, jt, x);
    return py::make_tuple(x[0], x[1], x[2], d2);
}

// all N^2 (dx,dy,dz,d2) of one frame pair, the layout tests/src/test_trajectory.cpp:21-41 dumps
template <class T>
py::array_t<double> probe_d2_all(T &t, size_t it, size_t jt) {
    const long n = t.get_natoms();
    py::array_t<double> out({n, n, 4L});
    double *o = out.mutable_data();
    for (long i = 0; i < n; ++i)
        for (long j = 0; j < n; ++j)
            o[(i * n + j) * 4 + 3] = t.d2_minImage(i, j, it, jt, o + (i * n + j) * 4);
    return out;
}

template <class T, class C>
void common_traj(C &c) {
    c.def("get_positions_copy", &copy_positions<T>)
        .def("get_box_copy", &copy_boxes<T>)
        .def("get_type_ids", &copy_type_ids<T>)
        .def("get_ntypes", [](T &t) { return (long)t.get_ntypes(); })
        .def("get_natoms", [](T &t) { return (long)t.get_natoms(); })
        .def("get_ntimesteps", [](T &t) { return (long)t.get_ntimesteps(); })
        .def("is_triclinic", [](T &t) { return t.is_triclinic(); })
        .def("d2_min_image", &probe_d2<T>)
        .def("d2_min_image_all", &probe_d2_all<T>)
        .def("write_lammps_binary", &T::dump_lammps_bin_traj);
}

template <class T>
void bind_gofrt(py::module &m, const char *name) {
    using G = Gofrt<double, T>;
    py::class_<G>(m, name, py::buffer_protocol())
        .def(py::init<T *, double, double, unsigned int, unsigned int, unsigned int, unsigned int,
                      unsigned int, bool>(),
             py::keep_alive<1, 2>())
        .def("reset", &G::reset)
        .def("getNumberOfExtraTimestepsNeeded", &G::nExtraTimesteps)
        .def("calculate", &G::calculate)
        .def("get_columns_description", &G::get_columns_description)
        .def_buffer([](G &g) -> py::buffer_info {
            return py::buffer_info(g.access_vdata(), sizeof(double),
                                   py::format_descriptor<double>::format(), g.get_shape().size(),
                                   g.get_shape(), g.get_stride());
        });
}

// The CLI's g(r,t) branch (analisi/main.cpp:552-585) without the option parser and the printing.
py::tuple block_average_gofrt(const std::string &path, unsigned nblocks, double rmin, double rmax,
                              unsigned nbin, unsigned tmax, unsigned nthreads, unsigned skip,
                              unsigned every, bool dump, bool wrap) {
    Trajectory tr(path);
    tr.set_pbc_wrap(wrap);
    BlockAverage<Gofrt<double, Trajectory>, double, double, unsigned int, unsigned int, unsigned int,
                 unsigned int, unsigned int, bool>
        gofr(&tr, nblocks);
    gofr.calculate(rmin, rmax, nbin, tmax, nthreads, skip, every, dump);
    auto shape = gofr.media()->get_shape();
    const size_t len = gofr.media()->lunghezza();
    py::array_t<double> mean(shape), var(shape);
    std::memcpy(mean.mutable_data(), gofr.media()->access_vdata(), len * sizeof(double));
    std::memcpy(var.mutable_data(), gofr.varianza()->access_vdata(), len * sizeof(double));
    return py::make_tuple(mean, var, gofr.puntatoreCalcolo()->get_columns_description());
}

// Neighbours<Trajectory_numpy,double> (lib/include/neighbour.h, lib/src/neighbour.cpp): the lists of one frame as arrays.
//   spec: one (max neighbours, cutoff^2, skin^2) per type, as the reference's ListSpec
//   -> counts [natoms][ntypes], indices [natoms][ntypes][maxn] (-1 past the count), r [natoms][ntypes][maxn][4]
//      (distance, x, y, z; 0 past the count), sann_n [natoms][ntypes] (only meaningful after a sorted update)
py::tuple ref_neighbours(Trajectory_numpy &t, std::vector<std::tuple<size_t, double, double>> spec, size_t timestep, bool sort) {
    Neighbours<Trajectory_numpy, double> nn(&t, spec);
    nn.update_neigh(timestep, sort);
    const long n = t.get_natoms(), nt = t.get_ntypes();
    size_t maxn = 0;
    for (auto &s : spec) maxn = std::max(maxn, std::get<0>(s));
    py::array_t<long> counts({n, nt}), idx({n, nt, (long)maxn}), sann({n, nt});
    py::array_t<double> r({n, nt, (long)maxn, 4L});
    std::fill(idx.mutable_data(), idx.mutable_data() + idx.size(), -1L);
    std::fill(r.mutable_data(), r.mutable_data() + r.size(), 0.0);
    for (long i = 0; i < n; ++i)
        for (long jt = 0; jt < nt; ++jt) {
            auto it = nn.get_neigh(i, jt);
            auto ir = nn.get_neigh_r(i, jt);
            counts.mutable_at(i, jt) = (long)it.size();
            sann.mutable_at(i, jt) = sort ? (long)nn.get_sann_n(i, jt) : -1L;
            for (size_t k = 0; k < it.size(); ++k) {
                idx.mutable_at(i, jt, (long)k) = (long)it.begin()[k];
                for (int c = 0; c < 4; ++c) r.mutable_at(i, jt, (long)k, c) = ir.begin()[k][c];
            }
        }
    return py::make_tuple(counts, idx, r, sann);
}

// SphericalBase<L,double,Trajectory_numpy>::calc without neighbour list (lib/src/sphericalbase.cpp:19-68):
//   -> result [natoms][ntypes][nbin][(L+1)^2], counter [natoms][ntypes][nbin]
template <int L>
py::tuple ref_sh_density_l(Trajectory_numpy &t, size_t nbin, std::vector<std::pair<double, double>> rminmax, int timestep) {
    SphericalBase<L, double, Trajectory_numpy> sb(&t, nbin, rminmax);
    const long n = t.get_natoms(), nt = t.get_ntypes(), nl = (L + 1) * (L + 1);
    py::array_t<double> result({n, nt, (long)nbin, nl});
    py::array_t<int> counter({n, nt, (long)nbin});
    std::vector<double> workspace(nl), cheby(2 * (L + 1));
    sb.calc(timestep, result.mutable_data(), workspace.data(), cheby.data(), counter.mutable_data(), nullptr);
    return py::make_tuple(result, counter);
}
py::tuple ref_sh_density(Trajectory_numpy &t, int lmax, size_t nbin, std::vector<std::pair<double, double>> rminmax, int timestep) {
    switch (lmax) {
        case 2: return ref_sh_density_l<2>(t, nbin, rminmax, timestep);
        case 3: return ref_sh_density_l<3>(t, nbin, rminmax, timestep);
        case 4: return ref_sh_density_l<4>(t, nbin, rminmax, timestep);
        case 5: return ref_sh_density_l<5>(t, nbin, rminmax, timestep);
        case 6: return ref_sh_density_l<6>(t, nbin, rminmax, timestep);
        case 7: return ref_sh_density_l<7>(t, nbin, rminmax, timestep);
        case 8: return ref_sh_density_l<8>(t, nbin, rminmax, timestep);
        case 9: return ref_sh_density_l<9>(t, nbin, rminmax, timestep);
        case 10: return ref_sh_density_l<10>(t, nbin, rminmax, timestep);
        default: throw std::runtime_error("the reference instantiates SphericalBase for l = 2 .. 10");
    }
}

}  // namespace

PYBIND11_MODULE(analisi_ref, m) {
    m.doc() = "rikigigi/analisi reference (unmodified sources) -- g(r,t) path only; test oracle";

    py::enum_<Trajectory_numpy::BoxFormat>(m, "BoxFormat", py::arithmetic())
        .value("Invalid", Trajectory_numpy::BoxFormat::Invalid)
        .value("CellVectors", Trajectory_numpy::BoxFormat::Cell_vectors)
        .value("LammpsOrtho", Trajectory_numpy::BoxFormat::Lammps_ortho)
        .value("LammpsTriclinic", Trajectory_numpy::BoxFormat::Lammps_triclinic);

    {
        py::class_<Trajectory_numpy> c(m, "Trajectory");
        c.def(py::init<py::buffer, py::buffer, py::buffer, py::buffer, Trajectory_numpy::BoxFormat, bool,
                       bool>(),
              py::keep_alive<1, 2>(), py::keep_alive<1, 3>(), py::keep_alive<1, 4>(),
              py::keep_alive<1, 5>());
        c.def("get_rotation_matrix", [](Trajectory_numpy &t) {
            double *q = t.get_rotation_matrix(0);
            if (q == nullptr) return py::array_t<double>();
            const long nts = t.get_ntimesteps();
            py::array_t<double> out({nts, 3L, 3L});
            std::memcpy(out.mutable_data(), q, sizeof(double) * nts * 9);
            return out;
        });
        common_traj<Trajectory_numpy>(c);
    }
    {
        py::class_<Trajectory> c(m, "Traj");
        c.def(py::init<std::string>())
            .def("setWrapPbc", &Trajectory::set_pbc_wrap)
            .def("setAccessWindowSize",
                 [](Trajectory &t, int ts) { return (int)t.set_data_access_block_size(ts); })
            .def("setAccessStart", [](Trajectory &t, int ts) { return (int)t.set_access_at(ts); })
            .def("get_lammps_type", [](Trajectory &t) {
                int *p = t.get_lammps_type();
                py::array_t<int> out((long)t.get_natoms());
                std::memcpy(out.mutable_data(), p, sizeof(int) * t.get_natoms());
                delete[] p;
                return out;
            });
        common_traj<Trajectory>(c);
    }
    bind_gofrt<Trajectory_numpy>(m, "Gofrt");
    bind_gofrt<Trajectory>(m, "Gofrt_lammps");
    m.def("block_average_gofrt", &block_average_gofrt, py::arg("path"), py::arg("nblocks"),
          py::arg("rmin"), py::arg("rmax"), py::arg("nbin"), py::arg("tmax"), py::arg("nthreads"),
          py::arg("skip"), py::arg("every") = 1, py::arg("dump") = false, py::arg("wrap") = true);
    m.def("neighbours", &ref_neighbours, py::arg("traj"), py::arg("spec"), py::arg("timestep"), py::arg("sort"));
    m.def("sh_density", &ref_sh_density, py::arg("traj"), py::arg("lmax"), py::arg("nbin"), py::arg("rminmax"), py::arg("timestep"));
    m.def("info", []() -> std::string { return _info_msg; });
}
