// operazionisulista.h -- VectorOp: the result buffer + elementwise algebra every calculation
// derives from.  Mirrors the public/protected surface of the reference's
// lib/include/operazionisulista.h:25-148 (same names, same error behaviour) so MediaVar and the
// pybind buffer protocol work unchanged; the implementation is this repository's own.
#ifndef ANALISI_B200_OPERAZIONISULISTA_H
#define ANALISI_B200_OPERAZIONISULISTA_H

#include <stdexcept>

template <class T, class TFLOAT = double>
class VectorOp {
public:
    // deep copy; reallocates when the lengths differ (reference :28-40)
    VectorOp<T, TFLOAT> &operator=(const VectorOp<T, TFLOAT> &rhs) {
        if (this == &rhs) return *this;
        if (data_length != rhs.data_length) {
            delete[] vdata;
            data_length = rhs.data_length;
            vdata = new TFLOAT[data_length];
        }
        for (unsigned int i = 0; i < data_length; ++i) vdata[i] = rhs.vdata[i];
        return *this;
    }

    T &operator+=(const T &rhs) { return zip(rhs, [](TFLOAT &a, TFLOAT b) { a += b; }); }
    T &operator-=(const T &rhs) { return zip(rhs, [](TFLOAT &a, TFLOAT b) { a -= b; }); }
    T &operator*=(const T &rhs) { return zip(rhs, [](TFLOAT &a, TFLOAT b) { a *= b; }); }
    T &operator/=(const T &rhs) { return zip(rhs, [](TFLOAT &a, TFLOAT b) { a /= b; }); }

    T &operator+=(const TFLOAT &s) { return each([&](TFLOAT &a) { a += s; }); }
    T &operator-=(const TFLOAT &s) { return each([&](TFLOAT &a) { a -= s; }); }
    T &operator*=(const TFLOAT &s) { return each([&](TFLOAT &a) { a *= s; }); }
    T &operator/=(const TFLOAT &s) { return each([&](TFLOAT &a) { a /= s; }); }

    unsigned int lunghezza() const { return data_length; }
    TFLOAT elemento(unsigned int i) const {
        if (i >= data_length) throw std::runtime_error("Out of range index");
        return vdata[i];
    }
    TFLOAT *access_vdata() { return vdata; }
    void azzera() {
        for (unsigned int i = 0; i < data_length; ++i) vdata[i] = 0;
    }
    void azzera(int start, int stop) {
        for (int i = start; i < stop; ++i) vdata[i] = 0;
    }

protected:
    VectorOp() : vdata(nullptr), data_length(0) {}
    VectorOp(const VectorOp<T, TFLOAT> &other) : vdata(nullptr), data_length(0) { operator=(other); }
    ~VectorOp() { delete[] vdata; }

    TFLOAT *vdata;
    unsigned int data_length;

private:
    template <class F>
    T &zip(const T &rhs, F f) {
        if (rhs.lunghezza() != data_length)
            throw std::runtime_error("Trying to operate on VectorOp of different sizes!");
        for (unsigned int i = 0; i < data_length; ++i) f(vdata[i], rhs.elemento(i));
        return static_cast<T &>(*this);
    }
    template <class F>
    T &each(F f) {
        for (unsigned int i = 0; i < data_length; ++i) f(vdata[i]);
        return static_cast<T &>(*this);
    }
};

#endif
