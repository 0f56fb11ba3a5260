// calculatemultithread.h -- the parameter block of the reference's CRTP thread fan-out
// (lib/include/calculatemultithread.h:18-197).  In this repository the only calculation is Gofrt,
// whose calculate() is ONE device job per block (libagofrt), so no host threads are spawned: the
// class keeps the constructor, the coercions (0 -> 1) and the protected members the derived class
// and the callers rely on.  nthreads is accepted and ignored (the GPU owns the parallelism).
#ifndef ANALISI_B200_CALCULATEMULTITHREAD_H
#define ANALISI_B200_CALCULATEMULTITHREAD_H

#include <sys/types.h>
#include <cstddef>

namespace CalculateMultiThread_Flags {
constexpr int PARALLEL_SPLIT_AVERAGE = 0b00000001;
constexpr int PARALLEL_SPLIT_TIME = 0b00000010;
constexpr int PARALLEL_SPLIT_ATOM = 0b00000100;
constexpr int SERIAL_LOOP_AVERAGE = 0b00010000;
constexpr int SERIAL_LOOP_TIME = 0b00100000;
constexpr int CALL_INNER_JOIN_DATA = 0b01000000;
constexpr int CALL_DEBUG_ROUTINE = 0b10000000;
constexpr int CALL_CALC_INIT = 0b100000000;
}  // namespace CalculateMultiThread_Flags

template <class T, int FLAGS_T = CalculateMultiThread_Flags::PARALLEL_SPLIT_AVERAGE |
                                 CalculateMultiThread_Flags::CALL_INNER_JOIN_DATA>
class CalculateMultiThread {
public:
    CalculateMultiThread(const ssize_t nthreads = 0, const ssize_t skip = 0, const size_t natoms = 0,
                         const ssize_t every = 0)
        : nthreads(nthreads == 0 ? 1 : nthreads), skip(skip == 0 ? 1 : skip), ntimesteps(0),
          every(every == 0 ? 1 : every), leff(0), natoms(natoms) {}
    static constexpr int FLAGS = FLAGS_T;

protected:
    ssize_t nthreads, skip, ntimesteps, every, leff;
    size_t natoms;
};

#endif
