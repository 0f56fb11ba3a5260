// sphericalbase.h -- SphericalBase<l,TFLOAT,T>: the density of the atoms around every atom, expanded in real spherical
// harmonics up to l, per partner type and radial bin, for one frame.
//
// Same public surface as the reference's lib/include/sphericalbase.h:6-57 / lib/src/sphericalbase.cpp:5-98: constructor
// (trajectory, number of radial bins, one (rmin, rmax) per ORDERED pair of types), calc(timestep, result, workspace,
// cheby, counter, nnl), get_single_atom_size / get_result_size, and the result layout
// [atom][type][bin][(l+1)^2] with the reference's ordering of the (l, m) components.  The pair loop and the spherical
// harmonics run on the GPU (agofrt_sh_density, include/agofrt.h), every operation rounded as the reference rounds it
// and the sums formed in the reference's order, so the result is the reference's bit for bit; workspace and cheby are
// accepted for compatibility and not used.  The variant that takes a neighbour list (nnl != nullptr: the SANN neighbours
// instead of radial bins, reference sphericalbase.cpp:69-96) is not built.
#ifndef ANALISI_B200_SPHERICALBASE_H
#define ANALISI_B200_SPHERICALBASE_H

#include <sstream>
#include <stdexcept>
#include <type_traits>
#include <utility>
#include <vector>

#include "analisi/device.h"
#include "analisi/neighbour.h"

template <int l, class TFLOAT, class T>
class SphericalBase {
    static_assert(std::is_same<TFLOAT, double>::value, "the device path computes in float64");
    static_assert(l >= 0 && l <= 10, "spherical harmonics up to l = 10 (the reference instantiates 2 .. 10)");

public:
    using Rminmax_t = std::vector<std::pair<TFLOAT, TFLOAT>>;
    using Neighbours_T = Neighbours<T, double>;
    SphericalBase(T *t, const size_t nbin, const Rminmax_t rminmax)
        : t(*t), natoms(t->get_natoms()), ntypes(static_cast<size_t>(t->get_ntypes())), nbin(nbin) {
        if (ntypes * ntypes != rminmax.size()) {
            std::stringstream ss;
            ss << "you must provide a radial range for each pair of atomic types, in total ntypes*ntypes pair of numbers. You provided "
               << rminmax.size() << " elements while ntypes is " << ntypes << " .";
            throw std::runtime_error(ss.str());
        }
        for (const auto &r : rminmax) {
            ranges.push_back(r.first);
            ranges.push_back(r.second);
        }
    }
    void calc(int timestep, TFLOAT *result, TFLOAT * /*workspace*/, TFLOAT * /*cheby*/, int *counter = nullptr,
              Neighbours_T *nnl = nullptr) const {
        if (nnl != nullptr)
            throw std::runtime_error("SphericalBase::calc with a neighbour list (SANN neighbours) is not built in this GPU port\n");
        analisi_device::check(agofrt_sh_density(t.device_window(), static_cast<size_t>(timestep), l, static_cast<unsigned>(nbin),
                                                ranges.data(), result, counter),
                              "agofrt_sh_density");
    }
    size_t get_single_atom_size() const { return (l + 1) * (l + 1) * nbin * ntypes; }
    size_t get_result_size() const { return get_single_atom_size() * natoms; }

private:
    T &t;
    const size_t natoms, ntypes, nbin;
    std::vector<double> ranges;
};

#endif
