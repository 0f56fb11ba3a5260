// calcoliblocchi.h -- MediaVar: Welford mean / variance-of-the-mean over blocks, written with
// the VectorOp algebra of whole result objects.  Same interface and the same sequence of rounded
// operations as the reference's lib/include/calcoliblocchi.h:21-65 (so block averages are
// bit-identical given bit-identical blocks); MediaVarCovar (Green-Kubo only) is out of scope.
//
// MediaVarDevice (this repository's addition) is the same consumer of blocks for
// calculations whose blocks are born on the GPU (Gofrt): the blocks are not read back one by one and pushed
// through eight whole-vector VectorOp passes on the host; a kernel folds each block's integer counts into mean
// and variance accumulators on the device (agofrt_blockavg_*, include/agofrt.h) with the same per-element
// sequence of rounded operations, so the two classes give bit-identical results.
#ifndef ANALISI_B200_CALCOLIBLOCCHI_H
#define ANALISI_B200_CALCOLIBLOCCHI_H

#include <type_traits>

#include "analisi/device.h"

template <class T>
class MediaVar {
public:
    MediaVar(T *mean, T *var, T *delta, T *tmp) : mean_(mean), var_(var), delta_(delta), tmp_(tmp), seen_(0) {}

    void calcola_begin(unsigned int s, T * /*calc*/) {
        mean_->reset(s);
        mean_->azzera();
        var_->reset(s);
        var_->azzera();
        delta_->reset(s);
        tmp_->reset(s);
        seen_ = 0;
    }

    // one more block x:  delta = x - mean;  mean += delta/(k+1);  var += (x - mean) * delta
    void calculate(T *x) {
        *delta_ = *x;
        *delta_ -= *mean_;
        *tmp_ = *delta_;
        *tmp_ /= static_cast<double>(seen_ + 1);
        *mean_ += *tmp_;
        *tmp_ = *x;
        *tmp_ -= *mean_;
        *tmp_ *= *delta_;
        *var_ += *tmp_;
        ++seen_;
    }

    // every block of x's last calculate_blocks(), in block order (the blocks were computed on the GPUs in one go;
    // each is fetched into x and folded in as above)
    template <class U = T>
    void calculate_blocks(U *x, unsigned int n_b) {
        for (unsigned int b = 0; b < n_b; ++b) {
            x->fetch_block_of_batch(b);
            calculate(x);
        }
    }

    // variance of the mean: sum / ((n_b - 1) n_b)
    void calcola_end(unsigned int n_b) { *var_ /= static_cast<double>((n_b - 1) * n_b); }

private:
    T *mean_, *var_, *delta_, *tmp_;
    unsigned int seen_;
};

// calculations that can leave their blocks on the device: set_keep_on_device / device_plan / fetch_block / get_incr
template <class T, class = void>
struct HasDeviceBlocks : std::false_type {};
template <class T>
struct HasDeviceBlocks<T, std::void_t<decltype(&T::set_keep_on_device), decltype(&T::device_plan), decltype(&T::fetch_block)>>
    : std::true_type {};

// calculations that can run a batch of blocks on the devices: block_batch_wanted / calculate_blocks / fetch_block_of_batch
template <class T, class = void>
struct HasBlockBatch : std::false_type {};
template <class T>
struct HasBlockBatch<T, std::void_t<decltype(&T::block_batch_wanted), decltype(&T::calculate_blocks), decltype(&T::fetch_block_of_batch)>>
    : std::true_type {};

// block consumers whose calcola_end() already leaves the last block in the calculation object (MediaVarDevice)
template <class C, class = void>
struct HasDeviceBlocksEnd : std::false_type {};
template <class C>
struct HasDeviceBlocksEnd<C, std::void_t<typename C::fetches_last_block>> : std::true_type {};

template <class T>
class MediaVarDevice {
public:
    using fetches_last_block = void;
    MediaVarDevice(T *mean, T *var) : mean_(mean), var_(var) {}
    ~MediaVarDevice() {
        if (calc_) calc_->set_keep_on_device(false);
        if (acc_) agofrt_blockavg_destroy(acc_);
    }
    MediaVarDevice(const MediaVarDevice &) = delete;
    MediaVarDevice &operator=(const MediaVarDevice &) = delete;

    void calcola_begin(unsigned int s, T *calc) {
        mean_->reset(s);
        mean_->azzera();
        var_->reset(s);
        var_->azzera();
        calc_ = calc;
        calc_->set_keep_on_device(true);
        if (!acc_)
            analisi_device::check(agofrt_blockavg_create(&acc_, analisi_device::Context::instance().handle()),
                                  "agofrt_blockavg_create");
        analisi_device::check(agofrt_blockavg_begin(acc_, mean_->lunghezza()), "agofrt_blockavg_begin");
    }

    // the block calc->calculate(primo) has just left on the device
    void calculate(T *calc) {
        if (mean_->lunghezza() == 0) return;
        analisi_device::check(agofrt_blockavg_push(acc_, calc->device_plan(), calc->get_incr()), "agofrt_blockavg_push");
    }

    // every block of calc's last calculate_blocks(), in block order
    void calculate_blocks(T *calc, unsigned int) {
        if (mean_->lunghezza() == 0) return;
        analisi_device::check(agofrt_blockavg_push_blocks(acc_, calc->device_plan(), calc->get_incr()), "agofrt_blockavg_push_blocks");
    }

    void calcola_end(unsigned int n_b) {
        analisi_device::check(agofrt_blockavg_end(acc_, n_b, mean_->access_vdata(), var_->access_vdata()),
                              "agofrt_blockavg_end");
        // leave the calculation object as the host path does: holding its last block
        calc_->set_keep_on_device(false);
        calc_->fetch_block();
        calc_ = nullptr;
    }

private:
    T *mean_, *var_;
    T *calc_ = nullptr;
    agofrt_blockavg *acc_ = nullptr;
};

#endif
