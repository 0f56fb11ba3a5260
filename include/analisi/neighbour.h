// neighbour.h -- Neighbours<T,TType>: cutoff neighbour lists of one frame, per atom and partner type, with the
// minimum-image vector of every partner, and the SANN neighbour count on top of the sorted lists.
//
// Same public surface as the reference's lib/include/neighbour.h:10-100 / lib/src/neighbour.cpp (ListSpec = one
// (max neighbours, cutoff^2, skin^2) per type; update_neigh(timestep, sort); get_neigh / get_neigh_r iterators over the
// reference's own memory layout; get_sann_n / get_sann / get_sann_r).  The N^2 loop over d2_minImage runs on the GPU
// (agofrt_neighbours, include/agofrt.h): the lists arrive in this object's buffers in the reference's layout, in the
// reference's order.  get_sann_n is the reference's own short loop over one sorted list (host arithmetic on a few
// dozen numbers, lib/src/neighbour.cpp:79-94).
#ifndef ANALISI_B200_NEIGHBOUR_H
#define ANALISI_B200_NEIGHBOUR_H

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <tuple>
#include <vector>

#include "analisi/device.h"

template <class T, class TType = double>
class Neighbours {
    static_assert(sizeof(TType) == sizeof(double), "the device path computes in float64");

public:
    using ListSpec = std::vector<std::tuple<size_t, TType, TType>>;
    using TType4 = TType[4];
    template <class TN>
    class NeighIterator {
    public:
        NeighIterator(TN *idxs, const size_t len) : idxs_(idxs), len_(len) {}
        const TN *begin() { return idxs_; }
        TN *begin_w() { return idxs_; }
        const TN *end() { return idxs_ + len_; }
        size_t size() const { return len_; }

    private:
        TN *idxs_;
        size_t len_;
    };

    Neighbours(T *t, const ListSpec nneigh_cut2_skin2)
        : t_(*t), spec_(nneigh_cut2_skin2), natoms_(t->get_natoms()), ntypes_(t->get_ntypes()) {
        if (ntypes_ != spec_.size()) throw std::runtime_error("In neighbours list you must specify parameters for each atomic type");
        size_t words = 0, doubles = 0;
        for (const auto &s : spec_) {
            list_offset_.push_back(words);
            rpos_offset_.push_back(doubles);
            nneigh_.push_back(static_cast<uint64_t>(std::get<0>(s)));
            cutoff2_.push_back(static_cast<double>(std::get<1>(s)));
            words += (std::get<0>(s) + 1) * natoms_;
            doubles += std::get<0>(s) * natoms_ * 4;
        }
        static_assert(sizeof(size_t) == sizeof(uint64_t), "the lists travel as 64-bit words");
        list_.assign(words, 0);
        rpos_.assign(doubles > 0 ? doubles : 1, 0.0);
    }

    void update_neigh(const size_t timestep, bool sort) {
        const int rc = agofrt_neighbours(t_.device_window(), timestep, nneigh_.data(), cutoff2_.data(), sort ? 1 : 0,
                                         reinterpret_cast<uint64_t *>(list_.data()), rpos_.data());
        if (rc == AGOFRT_ERR_TOO_LARGE) throw std::runtime_error("Too many neighbours in shell!");   // the reference's exception
        analisi_device::check(rc, "agofrt_neighbours");
        sorted_ = sort;
    }
    NeighIterator<size_t> get_neigh(const size_t iatom, const size_t jtype) const {
        size_t *base = const_cast<size_t *>(list_.data()) + list_offset_[jtype] + iatom * (nneigh(jtype) + 1);
        return NeighIterator<size_t>{base + 1, base[0]};
    }
    NeighIterator<TType4> get_neigh_r(const size_t iatom, const size_t jtype) const {
        return NeighIterator<TType4>{reinterpret_cast<TType4 *>(const_cast<double *>(rpos_.data()) + rpos_offset_[jtype] + iatom * nneigh(jtype) * 4),
                                     list_[list_offset_[jtype] + iatom * (nneigh(jtype) + 1)]};
    }
    // solid-angle nearest neighbours on the sorted list: the smallest n >= 3 with sum_{k<n} r_k <= r_n (n - 2)
    size_t get_sann_n(const size_t iatom, const size_t jtype) const {
        auto shell = get_neigh_r(iatom, jtype);
        const size_t have = shell.size();
        if (have < 3) return 0;
        const TType4 *r = shell.begin();
        TType sum = 0.0;
        for (size_t k = 0; k < 3; ++k) sum += r[k][0];   // (the additions in the reference's order: 0 + r0 + r1 + r2 + ...)
        size_t n = 3;
        for (; n < have && !(sum <= r[n][0] * (n - 2)); ++n) sum += r[n][0];
        return n;
    }
    NeighIterator<size_t> get_sann(const size_t iatom, const size_t jtype) const {
        return NeighIterator<size_t>{get_neigh(iatom, jtype).begin_w(), get_sann_n(iatom, jtype)};
    }
    NeighIterator<TType4> get_sann_r(const size_t iatom, const size_t jtype) const {
        return NeighIterator<TType4>{get_neigh_r(iatom, jtype).begin_w(), get_sann_n(iatom, jtype)};
    }

private:
    size_t nneigh(size_t itype) const { return std::get<0>(spec_[itype]); }
    T &t_;
    const ListSpec spec_;
    const size_t natoms_, ntypes_;
    bool sorted_ = false;
    std::vector<size_t> list_, list_offset_, rpos_offset_;
    std::vector<uint64_t> nneigh_;
    std::vector<double> rpos_, cutoff2_;
};

#endif
