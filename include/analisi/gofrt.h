// gofrt.h -- Gofrt<TFLOAT,T>: the time-dependent pair distribution function g(r,t) ("distinct" and
// "self" part of the van Hove function), computed on the GPUs of the process.
//
// Same public surface as the reference's lib/include/gofrt.h:28-107 / lib/src/gofrt.cpp:22-169
// (constructor argument order, reset / nExtraTimesteps / calculate / get_shape / get_stride /
// get_columns_description / operator=, the VectorOp algebra and buffer, the error texts), so the
// callers -- BlockAverageG (blockaverage.h), the CLI branch (reference analisi/main.cpp:552-585)
// and the pybind11 class (reference pyanalisi/src/pyanalisi.cpp:65-82) -- work unchanged.
//
// What is different is HOW calculate(primo) gets its numbers.  The reference's
// CalculateMultiThread::calculate (lib/include/calculatemultithread.h:106-162) walks lags and
// origins on the host and spawns nthreads std::threads per (lag, origin), each running
// calc_single_th over a slice of atoms.  Here calculate() shadows that loop: the whole block is ONE
// device job (agofrt_block, include/agofrt.h) on the device-resident copy of the trajectory window,
// returning integer bin counts that are multiplied by `incr` (reference gofrt.cpp:91-92, :118).
// There is no host implementation of the pair loop in this class: without a GPU calculate() throws.
#ifndef ANALISI_B200_GOFRT_H
#define ANALISI_B200_GOFRT_H

#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "analisi/calculatemultithread.h"
#include "analisi/device.h"
#include "analisi/operazionisulista.h"

namespace Gofrt_Flags {
constexpr int FLAGS = CalculateMultiThread_Flags::PARALLEL_SPLIT_ATOM | CalculateMultiThread_Flags::SERIAL_LOOP_TIME |
                      CalculateMultiThread_Flags::SERIAL_LOOP_AVERAGE | CalculateMultiThread_Flags::CALL_DEBUG_ROUTINE |
                      CalculateMultiThread_Flags::CALL_CALC_INIT;
}

template <class TFLOAT, class T>
class Gofrt : public VectorOp<Gofrt<TFLOAT, T>, TFLOAT>,
              public CalculateMultiThread<Gofrt<TFLOAT, T>, Gofrt_Flags::FLAGS> {
    static_assert(std::is_same<TFLOAT, double>::value,
                  "the device path computes in float64 (the reference only instantiates Gofrt<double,...>)");

public:
    using This = Gofrt<TFLOAT, T>;
    using CalculateMultiThread_T = CalculateMultiThread<This, Gofrt_Flags::FLAGS>;
    using CalculateMultiThread_T::FLAGS;
    using VectorOp_T = VectorOp<This, TFLOAT>;
    using VectorOp_T::azzera;

    // argument order as in the reference (gofrt.h:37-45): ..., tmax, nthreads, skip, every, debug.
    // nthreads is accepted for compatibility; the GPUs own the parallelism.
    Gofrt(T *t, TFLOAT rmin, TFLOAT rmax, unsigned int nbin, unsigned int tmax = 0, unsigned int nthreads = 0,
          unsigned int skip = 1, unsigned int every = 1, bool debug = false)
        : CalculateMultiThread_T(nthreads, skip, t->get_natoms(), every), rmin(rmin), rmax(rmax), incr(1),
          debug(debug), traiettoria(t), nbin(nbin), lmax(tmax) {
        // ANALISI_EDGE_PAIRS=1 (the CLI's --edge-pairs): report the pairs within 1 ulp of a bin edge per block
        if (const char *e = std::getenv("ANALISI_EDGE_PAIRS")) report_edges = std::atoi(e) != 0;
        // ANALISI_KERNEL_OPTIONS=<AGOFRT_OPT_* bits>: kernel selection for A/B measurements (e.g. 256 = never the
        // small-system kernel); the counts do not depend on it
        if (const char *e = std::getenv("ANALISI_KERNEL_OPTIONS")) kernel_options = static_cast<unsigned>(std::strtoul(e, nullptr, 0));
    }
    ~Gofrt() { drop_plan(); }
    Gofrt(const This &) = delete;

    // deep copy of the result buffer only (reference gofrt.cpp:157-160)
    This &operator=(const This &destra) {
        VectorOp_T::operator=(destra);
        return *this;
    }

    // frames a block needs beyond its own ntimesteps (reference gofrt.cpp:37-39)
    unsigned int nExtraTimesteps(unsigned int n_b) {
        const unsigned int a = static_cast<unsigned int>(traiettoria->get_ntimesteps() / (n_b + 1) + 1);
        return (a < lmax || lmax == 0) ? a : lmax;
    }

    // sizes for a block of n averaged timesteps; the buffer is not zeroed until calculate()
    // (reference gofrt.cpp:41-63)
    void reset(const unsigned int n) {
        const unsigned int nt = static_cast<unsigned int>(traiettoria->get_ntypes());
        std::stringstream d;
        d << "# The first column is the time difference in timesteps, then you have the bin index. Every column after is "
             "followed by the variance. Then you have the following: "
          << std::endl;
        for (unsigned int a = 0; a < nt; a++)
            for (unsigned int b = a; b < nt; b++) {
                d << "#g(" << a << ", " << b << "), different atom index: " << pair_slot(a, b) * 2 + 3 << std::endl;
                d << "#g(" << a << ", " << b << "), same atom index: " << (pair_slot(a, b) + nt * (nt + 1) / 2) * 2 + 3
                  << std::endl;
            }
        d << "# same atom index means that the atom is tracked around and the average self-spread is shown with larger "
             "time differences."
          << std::endl;
        d << "# different atom index is something that for t=0 is the traditional g(r) " << std::endl;
        c_descr = d.str();
        leff = (n < lmax || lmax == 0) ? n : lmax;
        ntimesteps = n;
        const unsigned int len = static_cast<unsigned int>(leff) * nt * (nt + 1) * nbin;
        if (len != data_length || !vdata) {
            delete[] vdata;
            data_length = len;
            vdata = new TFLOAT[data_length];
        }
    }

    std::vector<ssize_t> get_shape() {
        const ssize_t nt = static_cast<ssize_t>(traiettoria->get_ntypes());
        return {static_cast<ssize_t>(leff), nt * (nt + 1), static_cast<ssize_t>(nbin)};
    }
    std::vector<ssize_t> get_stride() {
        const ssize_t nt = static_cast<ssize_t>(traiettoria->get_ntypes());
        return {static_cast<ssize_t>(nt * (nt + 1) * nbin * sizeof(TFLOAT)), static_cast<ssize_t>(nbin * sizeof(TFLOAT)),
                static_cast<ssize_t>(sizeof(TFLOAT))};
    }
    std::string get_columns_description() { return c_descr; }

    // One block: lags 0, every, ... < leff; origins primo, primo+skip, ... < primo+ntimesteps; all
    // N^2 ordered pairs of (frame origin, frame origin+lag).  Shadows CalculateMultiThread::calculate
    // (reference calculatemultithread.h:106-162 + gofrt.cpp:73-155).
    void calculate(size_t primo) {
        if (static_cast<size_t>(leff) + static_cast<size_t>(ntimesteps) + primo >
            static_cast<size_t>(traiettoria->get_ntimesteps()) + 1)
            throw std::runtime_error(
                "trajectory is too short for this kind of calculation. Select a different starting timestep or lower the "
                "size of the average or the lenght of the time lag");
        // blocks kept on the device (MediaVarDevice): this object's buffer is only filled by fetch_block()
        const bool on_device = keep_on_device && !debug;
        if (!on_device) azzera();
        incr = (ntimesteps / skip > 0) ? 1.0 / static_cast<int>(ntimesteps / skip) : 1.0;
        if (data_length > 0) {
            agofrt_traj *win = traiettoria->device_window();
            if (!plan || plan_generation != traiettoria->device_generation()) {
                drop_plan();
                analisi_device::check(agofrt_plan_create(&plan, win, rmin, rmax, nbin), "agofrt_plan_create");
                plan_generation = traiettoria->device_generation();
            }
            // the trajectory double-buffers its device windows (read-ahead): follow the current one
            analisi_device::check(agofrt_plan_retarget(plan, win), "agofrt_plan_retarget");
            // block averages on the device (MediaVarDevice, calcoliblocchi.h): the counts stay on the GPU and this
            // object's buffer is filled once, after the last block (fetch_block).  The debug dump needs every block.
            if (!on_device) counts_buf.resize(data_length);
            edge_count = 0;
            const unsigned options = kernel_options | (report_edges ? AGOFRT_OPT_EDGES : AGOFRT_OPT_DEFAULT) |
                                     (on_device ? AGOFRT_OPT_ON_DEVICE : AGOFRT_OPT_DEFAULT);
            analisi_device::check(agofrt_block(plan, primo, static_cast<unsigned>(ntimesteps), static_cast<unsigned>(leff),
                                               static_cast<unsigned>(skip), static_cast<unsigned>(every), options,
                                               on_device ? nullptr : counts_buf.data(),
                                               report_edges ? &edge_count : nullptr, &stats),
                                  "agofrt_block");
            if (report_edges)
                std::cerr << "pairs within 1 ulp of a bin edge in this block: " << edge_count << " of "
                          << stats.pair_evals_total << " pair evaluations\n";
            sum_kernel_ms += stats.kernel_ms;
            sum_total_ms += stats.total_ms;
            sum_pair_evals += stats.pair_evals_total;
            ++ncalls;
            // the reference adds incr once per counted pair; count*incr is that sum with one rounding
            if (!on_device)
                for (unsigned int k = 0; k < data_length; ++k) vdata[k] = static_cast<TFLOAT>(counts_buf[k]) * incr;
        }
        if (debug) dump_block();
    }

    // this repository's addition -- block averages on the device (MediaVarDevice in calcoliblocchi.h drives these):
    // while keep_on_device is set calculate() leaves the block's counts on the GPU (this object's buffer is not
    // touched), device_plan() is the plan that holds them, and fetch_block() brings the last block into the buffer
    // (what calculate() would have left there).
    void set_keep_on_device(bool on) { keep_on_device = on; }
    agofrt_plan *device_plan() { return plan; }
    void fetch_block() {
        if (!plan || data_length == 0) return;
        counts_buf.resize(data_length);
        analisi_device::check(agofrt_plan_last_counts(plan, counts_buf.data(), data_length), "agofrt_plan_last_counts");
        for (unsigned int k = 0; k < data_length; ++k) vdata[k] = static_cast<TFLOAT>(counts_buf[k]) * incr;
    }

    // this repository's addition -- many small blocks at once (BlockAverageG::calcola_custom): whole blocks are dealt to
    // the GPUs (block b to device b mod n, the reference's round-robin of blocks over MPI ranks,
    // lib/include/blockaverage.h:146-186) and nothing returns to the host between them.  Worth it where one block is a
    // millisecond of kernel: systems that take the small-system kernel, and a window with the frames of ALL the blocks
    // that is small.  ANALISI_BLOCK_BATCH=0 turns it off.
    bool block_batch_wanted(unsigned int n_b, unsigned int s, unsigned int extra) const {
        // ANALISI_BLOCK_BATCH=1 / 0 forces it on / off.  Default: only with several GPUs -- on one GPU the block-by-block
        // loop already hides reading and upload of block b+1 behind block b (read-ahead), and measured faster on C1
        bool forced = false;
        if (const char *e = std::getenv("ANALISI_BLOCK_BATCH")) {
            if (std::atoi(e) == 0) return false;
            forced = true;
        }
        if (!forced && analisi_device::Context::instance().ndev() < 2) return false;
        if (debug || report_edges || n_b < 2) return false;
        const double frames = static_cast<double>(n_b) * s + extra;
        const unsigned int nt = static_cast<unsigned int>(traiettoria->get_ntypes());
        const double result_words = static_cast<double>(n_b) * ((s < lmax || lmax == 0) ? s : lmax) * nt * (nt + 1) * nbin;
        return traiettoria->get_natoms() <= 248 && frames * traiettoria->get_natoms() * 24.0 <= 512e6 && result_words * 8.0 <= 1.5e9;
    }
    // blocks b = 0 .. nblocks-1: reset(ntimesteps) (already done by the caller); calculate(primo0 + b*stride).  The counts
    // stay on the devices: MediaVarDevice folds them there, fetch_block_of_batch(b) brings one into this object's buffer.
    // Returns false (nothing done) when the blocks cannot be batched after all: the caller runs them one by one.
    bool calculate_blocks(size_t primo0, size_t stride, unsigned int nblocks) {
        if (nblocks == 0 || data_length == 0) return false;
        const size_t last = primo0 + static_cast<size_t>(nblocks - 1) * stride;
        if (static_cast<size_t>(leff) + static_cast<size_t>(ntimesteps) + last > static_cast<size_t>(traiettoria->get_ntimesteps()) + 1)
            throw std::runtime_error(
                "trajectory is too short for this kind of calculation. Select a different starting timestep or lower the "
                "size of the average or the lenght of the time lag");
        incr = (ntimesteps / skip > 0) ? 1.0 / static_cast<int>(ntimesteps / skip) : 1.0;
        agofrt_traj *win = traiettoria->device_window();
        if (!plan || plan_generation != traiettoria->device_generation()) {
            drop_plan();
            analisi_device::check(agofrt_plan_create(&plan, win, rmin, rmax, nbin), "agofrt_plan_create");
            plan_generation = traiettoria->device_generation();
        }
        analisi_device::check(agofrt_plan_retarget(plan, win), "agofrt_plan_retarget");
        const int rc = agofrt_blocks(plan, primo0, stride, nblocks, static_cast<unsigned>(ntimesteps), static_cast<unsigned>(leff),
                                     static_cast<unsigned>(skip), static_cast<unsigned>(every), kernel_options, &stats);
        if (rc == AGOFRT_ERR_ARG || rc == AGOFRT_ERR_TOO_LARGE) return false;   // not batchable: blocks one by one
        analisi_device::check(rc, "agofrt_blocks");
        sum_kernel_ms += stats.kernel_ms;
        sum_total_ms += stats.total_ms;
        sum_pair_evals += stats.pair_evals_total;
        ncalls += nblocks;
        return true;
    }
    void fetch_block_of_batch(unsigned int b) {
        counts_buf.resize(data_length);
        analisi_device::check(agofrt_plan_block_counts(plan, b, counts_buf.data(), data_length), "agofrt_plan_block_counts");
        for (unsigned int k = 0; k < data_length; ++k) vdata[k] = static_cast<TFLOAT>(counts_buf[k]) * incr;
    }

    // this repository's additions: the raw integer counts of the last calculate() and its device timings
    const std::vector<uint64_t> &counts() const { return counts_buf; }
    const agofrt_stats &last_stats() const { return stats; }
    TFLOAT get_incr() const { return incr; }
    // Pairs whose squared distance is a bin threshold or the double just below one: a 1-ulp change of d2 would move
    // them to the neighbouring bin (or in / out of the range).  Counting them selects the kernel that compares
    // against the plain threshold table (slower); off by default.
    void set_report_edges(bool on) { report_edges = on; }
    uint64_t edge_pairs() const { return edge_count; }
    // ... and the sums over every calculate() of this object (BlockAverageG runs all blocks on one object)
    double total_kernel_ms() const { return sum_kernel_ms; }
    double total_device_ms() const { return sum_total_ms; }
    uint64_t total_pair_evals() const { return sum_pair_evals; }
    unsigned int total_calls() const { return ncalls; }

private:
    using VectorOp_T::data_length;
    using VectorOp_T::vdata;
    using CalculateMultiThread_T::every;
    using CalculateMultiThread_T::leff;
    using CalculateMultiThread_T::nthreads;
    using CalculateMultiThread_T::ntimesteps;
    using CalculateMultiThread_T::skip;

    // slot of the unordered type pair: P - (hi+1)(hi+2)/2 + lo  (reference gofrt.h:86-104)
    unsigned int pair_slot(unsigned int a, unsigned int b) const {
        const unsigned int nt = static_cast<unsigned int>(traiettoria->get_ntypes());
        const unsigned int lo = a < b ? a : b, hi = a < b ? b : a;
        return nt * (nt + 1) / 2 - (hi + 1) * (hi + 2) / 2 + lo;
    }

    // -d / debug: append the block as text, "lag bin v0 v1 ..." (reference gofrt.cpp:137-153)
    void dump_block() {
        std::ofstream out("gofrt.dump", std::ios::app);
        const unsigned int nt = static_cast<unsigned int>(traiettoria->get_ntypes());
        const unsigned int ncol = nt * (nt + 1);
        for (unsigned int ts = 0; ts < static_cast<unsigned int>(leff); ts++)
            for (unsigned int r = 0; r < nbin; r++) {
                out << ts << " " << r;
                for (unsigned int c = 0; c < ncol; c++) out << " " << vdata[(ts * ncol + c) * nbin + r];
                out << "\n";
            }
        out << "\n\n";
    }

    void drop_plan() {
        if (plan) agofrt_plan_destroy(plan);
        plan = nullptr;
    }

    TFLOAT rmin, rmax, incr;
    bool debug;
    T *traiettoria;
    unsigned int nbin, lmax;
    std::string c_descr;
    agofrt_plan *plan = nullptr;
    uint64_t plan_generation = 0;
    std::vector<uint64_t> counts_buf;
    agofrt_stats stats{};
    bool report_edges = false;
    bool keep_on_device = false;
    unsigned kernel_options = 0;
    uint64_t edge_count = 0;
    double sum_kernel_ms = 0, sum_total_ms = 0;
    uint64_t sum_pair_evals = 0;
    unsigned int ncalls = 0;
};

#endif
