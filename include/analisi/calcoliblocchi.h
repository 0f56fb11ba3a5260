// calcoliblocchi.h -- MediaVar: Welford mean / variance-of-the-mean over blocks, written with
// the VectorOp algebra of whole result objects.  Same interface and the same sequence of rounded
// operations as the reference's lib/include/calcoliblocchi.h:21-65 (so block averages are
// bit-identical given bit-identical blocks); MediaVarCovar (Green-Kubo only) is out of scope.
#ifndef ANALISI_B200_CALCOLIBLOCCHI_H
#define ANALISI_B200_CALCOLIBLOCCHI_H

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

    // variance of the mean: sum / ((n_b - 1) n_b)
    void calcola_end(unsigned int n_b) { *var_ /= static_cast<double>((n_b - 1) * n_b); }

private:
    T *mean_, *var_, *delta_, *tmp_;
    unsigned int seen_;
};

#endif
