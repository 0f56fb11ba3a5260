// blockaverage.h -- BlockAverageG<TR,T,Args...> ("MediaBlocchi"): split the trajectory in n_b blocks,
// run the calculation T on every block, accumulate mean and variance of the mean over blocks.
//
// Same interface and the same block geometry as the reference's lib/include/blockaverage.h:83-221:
//   s = (ntimesteps - nExtraTimesteps(n_b)) / n_b; window of s + nExtra frames at iblock*s;
//   per block reset(s), set_access_at(iblock*s), calculate(iblock*s), MediaVar::calculate.
// The reference optionally deals blocks to MPI ranks and gathers them on the root with
// MPI_Send/MPI_Recv (blockaverage.h:146-186, mp.h:35-41).  Here the parallelism sits one level
// lower: every block's (lag, origin, atom-tile) work units are sharded over all GPUs of the box inside
// T::calculate and the integer histograms are combined by one NCCL all-reduce (agofrt_block), so the
// block loop itself stays serial and MediaVar sees the blocks in order -- its floating-point result
// does not depend on the number of GPUs.
//
// Additive extension: TraiettoriaF is also specialised for Trajectory_numpy (all frames resident), which
// the reference leaves as abort() (blockaverage.h:32-37), so block averages work from python buffers too.
#ifndef ANALISI_B200_BLOCKAVERAGE_H
#define ANALISI_B200_BLOCKAVERAGE_H

#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "analisi/calcoliblocchi.h"
#include "analisi/cronometro.h"
#include "analisi/trajectory.h"

template <class TR>
class TraiettoriaF {
public:
    static void set_data_access_block_size(unsigned int, TR *) {}
    static void set_access_at(unsigned int, TR *) {}
    static void hint_stride(unsigned int, TR *) {}
    static unsigned int get_ntimesteps(TR *t) { return static_cast<unsigned int>(t->get_ntimesteps()); }
};

template <>
class TraiettoriaF<Trajectory> {
public:
    static void set_data_access_block_size(unsigned int s, Trajectory *t) { t->set_data_access_block_size(s); }
    static void set_access_at(unsigned int s, Trajectory *t) { t->set_access_at(s); }
    // this repository's addition: the block loop tells the reader how far apart its windows are, so the
    // read-ahead can start with the first block instead of guessing the stride from the first two
    static void hint_stride(unsigned int s, Trajectory *t) { t->set_access_stride_hint(s); }
    static unsigned int get_ntimesteps(Trajectory *t) { return static_cast<unsigned int>(t->get_ntimesteps()); }
};

template <class TR, class T, typename... Args>
class BlockAverageG {
public:
    BlockAverageG(TR *t, const unsigned int &numero_blocchi) : n_b(numero_blocchi), traiettoria(t) {}
    ~BlockAverageG() {
        delete Tmedio;
        delete Tvar;
        delete delta;
        delete tmp;
        delete calcolo;
    }
    BlockAverageG(const BlockAverageG &) = delete;
    BlockAverageG &operator=(const BlockAverageG &) = delete;

    // `calc` consumes the blocks: calcola_begin(s, calcolo), calculate(calcolo) per block, calcola_end(n_b)
    template <class Calcolo>
    void calcola_custom(Calcolo *calc, Args... arg) {
        delete calcolo;
        calcolo = new T(traiettoria, arg...);
        const unsigned int extra = calcolo->nExtraTimesteps(n_b);
        const int per_block = n_b > 0 ? (static_cast<int>(TraiettoriaF<TR>::get_ntimesteps(traiettoria)) - static_cast<int>(extra)) /
                                            static_cast<int>(n_b)
                                      : 0;
        if (per_block <= 0) {
            // (the reference aborts here, blockaverage.h:116-119; an exception keeps a python caller alive and the CLI
            // still exits with code 1)
            std::stringstream ss;
            ss << "Cannot divide the trajectory in " << n_b << " blocks!\n";
            throw std::runtime_error(ss.str());
        }
        s = static_cast<unsigned int>(per_block);
        ok = true;
        calcolo->reset(s);
        calc->calcola_begin(s, calcolo);
        // Many small blocks (this repository's addition): one window with the frames of all of them, whole blocks dealt
        // to the GPUs and computed without coming back to the host in between, then folded in block order -- the
        // reference's MPI branch deals blocks to ranks the same way (blockaverage.h:146-186).
        if constexpr (HasBlockBatch<T>::value) {
            if (calcolo->block_batch_wanted(n_b, s, extra)) {
                cronometro cron;
                cron.start();
                const bool debug_times = std::getenv("AGOFRT_DEBUG") != nullptr;
                auto lap = [&](const char *what) {
                    if (!debug_times) return;
                    cron.stop();
                    std::cerr << "[blocks] " << what << ": +" << cron.time_last() << "s\n";
                    cron.start();
                };
                TraiettoriaF<TR>::set_data_access_block_size(n_b * s + extra, traiettoria);
                lap("window buffers");
                TraiettoriaF<TR>::set_access_at(0, traiettoria);
                lap("window read");
                if (calcolo->calculate_blocks(0, s, n_b)) {
                    lap("blocks on the devices");
                    calc->calculate_blocks(calcolo, n_b);
                    calc->calcola_end(n_b);
                    if constexpr (!HasDeviceBlocksEnd<Calcolo>::value) calcolo->fetch_block_of_batch(n_b - 1);
                    lap("mean, variance, last block");
                    cron.stop();
                    std::cerr << "Time for " << n_b << " blocks of " << s << " steps, computed as one batch on the GPUs: " << cron.time()
                              << "s.\n";
                    return;
                }
            }
        }
        TraiettoriaF<TR>::set_data_access_block_size(s + extra, traiettoria);
        TraiettoriaF<TR>::hint_stride(s, traiettoria);
        cronometro cron;
        cron.set_expected(1.0 / double(n_b));
        cron.start();
        for (unsigned int iblock = 0; iblock < n_b; iblock++) {
            std::cerr << "beginning of block calculation " << iblock + 1 << std::endl;
            calcolo->reset(s);
            TraiettoriaF<TR>::set_access_at(iblock * s, traiettoria);
            calcolo->calculate(iblock * s);
            calc->calculate(calcolo);
            cron.stop();
            std::cerr << "Time for block " << iblock + 1 << " / " << n_b << ": " << cron.time_last()
                      << "s. Elapsed time and expected time to finish: " << cron.time() << "s " << cron.expected() << "s.\n";
        }
        calc->calcola_end(n_b);
    }

    void calculate(Args... arg) {
        delete Tmedio;
        delete Tvar;
        delete delta;
        delete tmp;
        delta = tmp = nullptr;
        Tmedio = new T(traiettoria, arg...);
        Tvar = new T(traiettoria, arg...);
        // Blocks born on the GPU are averaged there (MediaVarDevice: same rounded operations, bit-identical
        // result, no per-block read-back); ANALISI_DEVICE_BLOCKS=0 keeps the host MediaVar.
        if constexpr (HasDeviceBlocks<T>::value) {
            const char *e = std::getenv("ANALISI_DEVICE_BLOCKS");
            if (!e || std::atoi(e) != 0) {
                MediaVarDevice<T> media_var(Tmedio, Tvar);
                calcola_custom<MediaVarDevice<T>>(&media_var, arg...);
                return;
            }
        }
        delta = new T(traiettoria, arg...);
        tmp = new T(traiettoria, arg...);
        MediaVar<T> media_var(Tmedio, Tvar, delta, tmp);
        calcola_custom<MediaVar<T>>(&media_var, arg...);
    }

    T *media() {
        if (!ok) throw std::runtime_error("BlockAverage: calculate() has not been run\n");
        return Tmedio;
    }
    T *varianza() {
        if (!ok) throw std::runtime_error("BlockAverage: calculate() has not been run\n");
        return Tvar;
    }
    T *puntatoreCalcolo() {
        if (!ok) throw std::runtime_error("BlockAverage: calculate() has not been run\n");
        return calcolo;
    }
    unsigned int block_size() const { return s; }

private:
    unsigned int n_b, s = 0;
    T *Tmedio = nullptr, *Tvar = nullptr, *calcolo = nullptr, *delta = nullptr, *tmp = nullptr;
    TR *traiettoria;
    bool ok = false;
};

template <class T, typename... Args>
using BlockAverage = BlockAverageG<Trajectory, T, Args...>;

#endif
