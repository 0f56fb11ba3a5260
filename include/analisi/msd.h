// msd.h -- MSD<T,FPE>: mean square displacement per atomic type as a function of the time lag, optionally of the
// per-type centres of mass too (`analisi -q / -Q`, pyanalisi.MeanSquareDisplacement).
//
// Same public surface as the reference's lib/include/msd.h:26-60 / lib/src/msd.cpp:26-163: constructor
// (T*, skip, tmax, nthreads, centre-of-mass MSD, atoms in the frame of their type's centre of mass, debug), reset,
// nExtraTimesteps, calculate, get_shape/get_stride (leff, f_cm, ntypes), operator=, the VectorOp buffer, the two
// length checks of calc_init with the reference's messages, the msd.dump side effect of `debug`.
// calculate() shadows CalculateMultiThread's split over lags: the whole block is one device job (agofrt_msd) on the
// device-resident window.  SURVEY.md section 8f rank 3.  FPE (trap on NaN) has no device counterpart and is accepted
// as a template argument only.
#ifndef ANALISI_B200_MSD_H
#define ANALISI_B200_MSD_H

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "analisi/calculatemultithread.h"
#include "analisi/device.h"
#include "analisi/operazionisulista.h"

namespace MSD_Flags {
constexpr int FLAGS = CalculateMultiThread_Flags::PARALLEL_SPLIT_TIME | CalculateMultiThread_Flags::CALL_DEBUG_ROUTINE |
                      CalculateMultiThread_Flags::CALL_CALC_INIT;
}

template <class T, bool FPE = false>
class MSD : public VectorOp<MSD<T, FPE>>, public CalculateMultiThread<MSD<T, FPE>, MSD_Flags::FLAGS> {
public:
    using This = MSD<T, FPE>;
    using CMT = CalculateMultiThread<This, MSD_Flags::FLAGS>;

    MSD(T *t, unsigned int skip = 1, unsigned int tmax = 0, unsigned int nthreads = 0, bool calcola_msd_centro_di_massa = false,
        bool calcola_msd_nel_sistema_del_centro_di_massa = false, bool debug = false)
        : CMT(nthreads, skip, 0), traiettoria(t), lmax(tmax), f_cm(calcola_msd_centro_di_massa ? 2 : 1), ntypes(0),
          cm_msd(calcola_msd_centro_di_massa), cm_self(calcola_msd_nel_sistema_del_centro_di_massa), debug(debug) {}
    MSD(const This &) = delete;

    std::vector<ssize_t> get_shape() const {
        return {static_cast<ssize_t>(leff), static_cast<ssize_t>(f_cm), static_cast<ssize_t>(ntypes)};
    }
    std::vector<ssize_t> get_stride() const {
        return {static_cast<ssize_t>(ntypes * f_cm * sizeof(double)), static_cast<ssize_t>(ntypes * sizeof(double)),
                static_cast<ssize_t>(sizeof(double))};
    }
    unsigned int nExtraTimesteps(unsigned int n_b) {
        const size_t a = traiettoria->get_ntimesteps() / (n_b + 1) + 1;
        return static_cast<unsigned int>((a < lmax || lmax == 0) ? a : lmax);
    }
    void reset(const unsigned int numeroTimestepsPerBlocco) {
        leff = (numeroTimestepsPerBlocco < lmax || lmax == 0) ? numeroTimestepsPerBlocco : lmax;
        ntypes = traiettoria->get_ntypes();
        ntimesteps = numeroTimestepsPerBlocco;
        const unsigned int len = static_cast<unsigned int>(leff * ntypes * f_cm);
        if (len != data_length || !vdata) {
            delete[] vdata;
            data_length = len;
            vdata = new double[data_length];
        }
    }
    This &operator=(const This &destra) {
        VectorOp<This>::operator=(destra);
        return *this;
    }

    void calculate(size_t primo) {
        // reference msd.cpp:52-61
        if (static_cast<size_t>(leff) + static_cast<size_t>(ntimesteps) + primo > static_cast<size_t>(traiettoria->get_ntimesteps()))
            throw std::runtime_error(
                "trajectory is too short for this kind of calculation. Select a different starting timestep or lower the "
                "size of the average or the lenght of the time lag");
        if (static_cast<size_t>(leff) + static_cast<size_t>(ntimesteps) > traiettoria->get_nloaded_timesteps()) {
            std::stringstream ss;
            ss << "there are not enough loaded timesteps inside the trajectory object. I need at least " << leff + ntimesteps
               << " timesteps to do the requested calculation";
            throw std::runtime_error(ss.str());
        }
        if (data_length == 0) return;
        agofrt_traj *win = traiettoria->device_window();
        if (cm_msd || cm_self) {
            const size_t nfr = traiettoria->get_nloaded_timesteps();
            const ssize_t first = traiettoria->get_current_timestep();
            cm_buf.resize(nfr * ntypes * 3);
            for (size_t f = 0; f < nfr; ++f)
                for (size_t ty = 0; ty < ntypes; ++ty) {
                    const double *c = traiettoria->template positions_cm<false>(static_cast<int>(first + f), static_cast<int>(ty));
                    if (!c) throw std::runtime_error("the trajectory holds no centres of mass (velocities were not loaded)\n");
                    for (int k = 0; k < 3; ++k) cm_buf[(f * ntypes + ty) * 3 + k] = c[k];
                }
            analisi_device::check(agofrt_traj_set_cm(win, static_cast<size_t>(first), nfr, cm_buf.data()), "agofrt_traj_set_cm");
        }
        analisi_device::check(agofrt_msd(win, primo, static_cast<unsigned>(ntimesteps), static_cast<unsigned>(leff),
                                         static_cast<unsigned>(skip), cm_msd, cm_self, vdata, &stats),
                              "agofrt_msd");
        if (debug) {
            std::ofstream out("msd.dump", std::ios::app);
            for (size_t ts = 0; ts < static_cast<size_t>(leff); ts++) {
                out << ts;
                for (size_t k = 0; k < ntypes * f_cm; k++) out << " " << vdata[ntypes * ts * f_cm + k];
                out << "\n";
            }
            out << "\n\n";
        }
    }
    const agofrt_stats &last_stats() const { return stats; }

private:
    using VectorOp<This>::vdata;
    using VectorOp<This>::data_length;
    using CMT::leff;
    using CMT::nthreads;
    using CMT::ntimesteps;
    using CMT::skip;
    T *traiettoria;
    size_t lmax, f_cm, ntypes;
    bool cm_msd, cm_self, debug;
    std::vector<double> cm_buf;
    agofrt_stats stats{};
};

#endif
