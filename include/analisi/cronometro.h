// cronometro.h -- wall-clock stopwatch with a remaining-time estimate, used by BlockAverageG for its
// per-block stderr line.  Same member names as the reference's lib/include/cronometro.h:17-36
// (start / stop / reset / time / time_last / set_expected / expected); std::chrono::steady_clock.
#ifndef ANALISI_B200_CRONOMETRO_H
#define ANALISI_B200_CRONOMETRO_H

#include <chrono>

class cronometro {
public:
    cronometro() = default;
    void start() { t0 = clock::now(); }
    // adds the time since the last start(); restarts the lap so that consecutive stop() calls measure laps
    void stop(unsigned int = 0) {
        const clock::time_point now = clock::now();
        last = std::chrono::duration<double>(now - t0).count();
        total += last;
        t0 = now;
        if (estimate) {
            done += fraction;
            remaining = done > 0 ? total * (1.0 - done) / done : 0.0;
        }
    }
    void reset() {
        total = last = remaining = done = 0;
    }
    double time() const { return total; }
    double time_last() const { return last; }
    // `f` = the fraction of the whole work finished by every stop()
    void set_expected(double f) {
        estimate = true;
        fraction = f;
        done = 0;
    }
    void unset_expected() { estimate = false; }
    double expected() const { return remaining; }

private:
    using clock = std::chrono::steady_clock;
    clock::time_point t0 = clock::now();
    double total = 0, last = 0, remaining = 0, fraction = 0, done = 0;
    bool estimate = false;
};

#endif
