// istogrammaatomiraggio.h -- IstogrammaAtomiRaggio: histogram of the number of neighbours of each type
// within a radius (`analisi --neighbour r`).
//
// Same interface as the reference's lib/include/istogrammaatomiraggio.h:20-35 /
// lib/src/istogrammaatomiraggio.cpp:17-85: constructor (Trajectory*, r, skip, nthreads), reset(n),
// calculate(tstart) over the frames tstart, tstart+skip, ... < tstart+n, nExtraTimesteps() == 0, and
// get_hist() -> one std::map<count, occurrences> per type that ACCUMULATES over calculate() calls.
// The N^2 distance loop (the reference's lines 50-56, the same d2_minImage as g(r,t) at lag 0) runs on the
// GPUs through agofrt_neighbour_hist; nthreads is accepted and ignored.  SURVEY.md section 8f rank 2.
#ifndef ANALISI_B200_ISTOGRAMMAATOMIRAGGIO_H
#define ANALISI_B200_ISTOGRAMMAATOMIRAGGIO_H

#include <cstdint>
#include <map>
#include <vector>

#include "analisi/device.h"

template <class TR>
class IstogrammaAtomiRaggioG {
public:
    IstogrammaAtomiRaggioG(TR *t, double r, unsigned int skip = 1, unsigned int nthreads = 0)
        : r(r), skip(skip < 1 ? 1 : skip), nthreads(nthreads < 1 ? 1 : nthreads), traiettoria(t) {}
    ~IstogrammaAtomiRaggioG() { delete[] hist; }
    IstogrammaAtomiRaggioG(const IstogrammaAtomiRaggioG &) = delete;
    IstogrammaAtomiRaggioG &operator=(const IstogrammaAtomiRaggioG &) = delete;

    // empties the histograms (reference :22-29)
    void reset(const unsigned int numeroTimestepsPerBlocco) {
        numeroTimestepsBlocco = numeroTimestepsPerBlocco;
        ntypes = static_cast<unsigned int>(traiettoria->get_ntypes());
        natoms = static_cast<unsigned int>(traiettoria->get_natoms());
        delete[] hist;
        hist = new std::map<unsigned int, unsigned int>[ntypes];
    }
    unsigned int nExtraTimesteps(unsigned int) { return 0; }

    void calculate(unsigned int tstart) {
        if (!hist) reset(numeroTimestepsBlocco);
        dense.assign(static_cast<size_t>(ntypes) * (natoms + 1), 0);
        analisi_device::check(agofrt_neighbour_hist(traiettoria->device_window(), r, tstart, numeroTimestepsBlocco, skip,
                                                    dense.data(), &stats),
                              "agofrt_neighbour_hist");
        for (unsigned int ty = 0; ty < ntypes; ++ty)
            for (unsigned int c = 0; c <= natoms; ++c) {
                const uint64_t v = dense[static_cast<size_t>(ty) * (natoms + 1) + c];
                if (v) hist[ty][c] += static_cast<unsigned int>(v);
            }
    }
    std::map<unsigned int, unsigned int> *get_hist() { return hist; }
    const agofrt_stats &last_stats() const { return stats; }

private:
    double r;
    unsigned int skip, nthreads, ntypes = 0, natoms = 0, numeroTimestepsBlocco = 0;
    TR *traiettoria;
    std::map<unsigned int, unsigned int> *hist = nullptr;
    std::vector<uint64_t> dense;
    agofrt_stats stats{};
};

class Trajectory;
using IstogrammaAtomiRaggio = IstogrammaAtomiRaggioG<Trajectory>;

#endif
