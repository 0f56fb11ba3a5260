// CPU unit checks of the host-side classes that need no GPU (compiled and run by tests/test_host_units.py):
// VectorOp algebra, MediaVar (Welford over blocks), BlockAverageG's block geometry and call order with a
// mock calculation, CalculateMultiThread's coercions, the box permutation, TriclinicLammpsCell's QR.
// Prints one JSON object; the python side compares it with numpy and with the oracle's MediaVar.
#include <cmath>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "analisi/blockaverage.h"
#include "analisi/calcoliblocchi.h"
#include "analisi/calculatemultithread.h"
#include "analisi/operazionisulista.h"
#include "analisi/trajectory_numpy.h"
#include "analisi/triclinic.h"
#include "analisi/gofrt.h"
#include "analisi/msd.h"

// A calculation with the interface BlockAverageG and MediaVar expect (what Gofrt offers), computed on the host:
// element k of block `primo` is a fixed function of (primo, k) -- no trajectory access, no device.
struct Mock : public VectorOp<Mock, double>, public CalculateMultiThread<Mock> {
    Mock(Trajectory_numpy *t, unsigned int len, unsigned int extra) : CalculateMultiThread<Mock>(1, 1, t->get_natoms(), 1), len(len), extra(extra) {}
    void reset(unsigned int n) {
        ntimesteps = n;
        if (data_length != len) {
            delete[] vdata;
            data_length = len;
            vdata = new double[len];
        }
        resets.push_back(n);
    }
    unsigned int nExtraTimesteps(unsigned int) { return extra; }
    void calculate(size_t primo) {
        calls.push_back(primo);
        for (unsigned int k = 0; k < len; ++k) vdata[k] = value(primo, k);
    }
    static double value(size_t primo, unsigned int k) { return std::sin(0.37 * primo + k) * 3.0 + 0.1 * k + 1.0 / (1.0 + primo); }
    unsigned int len, extra;
    static std::vector<size_t> calls;
    static std::vector<unsigned int> resets;
};
std::vector<size_t> Mock::calls;
std::vector<unsigned int> Mock::resets;

static void print_vec(const char *name, const std::vector<double> &v, bool last = false) {
    std::printf("\"%s\": [", name);
    for (size_t i = 0; i < v.size(); ++i) std::printf("%s%.17g", i ? ", " : "", v[i]);
    std::printf("]%s\n", last ? "" : ",");
}

// BlockAverageG::calculate folds the blocks on the device only for calculations that can keep them there
static_assert(HasDeviceBlocks<Gofrt<double, Trajectory>>::value && HasDeviceBlocks<Gofrt<double, Trajectory_numpy>>::value,
              "Gofrt offers set_keep_on_device / device_plan / fetch_block");
static_assert(!HasDeviceBlocks<MSD<Trajectory>>::value, "MSD blocks are averaged by the host MediaVar");

int main() {
    std::printf("{\n");
    // ---- VectorOp ----
    struct V : public VectorOp<V, double> {
        explicit V(unsigned n) {
            data_length = n;
            vdata = new double[n];
        }
    };
    V a(4), b(4), c(3);
    for (int i = 0; i < 4; ++i) {
        a.access_vdata()[i] = i + 1.0;
        b.access_vdata()[i] = 0.5 * (i + 1);
    }
    a += b;        // 1.5 3 4.5 6
    a *= b;        // .75 3 6.75 12
    a -= 0.25;     // .5 2.75 6.5 11.75
    a /= 2.0;
    a /= b;        // elementwise
    std::vector<double> va(a.access_vdata(), a.access_vdata() + 4);
    print_vec("vectorop", va);
    bool threw = false;
    try {
        a += c;
    } catch (const std::runtime_error &) {
        threw = true;
    }
    bool threw2 = false;
    try {
        a.elemento(4);
    } catch (const std::runtime_error &) {
        threw2 = true;
    }
    V d(1);
    d = a;   // deep copy with reallocation
    std::printf("\"vectorop_size_mismatch_throws\": %s, \"vectorop_range_throws\": %s, \"vectorop_copy_len\": %u,\n", threw ? "true" : "false",
                threw2 ? "true" : "false", d.lunghezza());

    // ---- CalculateMultiThread coercions (0 -> 1) ----
    struct C : public CalculateMultiThread<C> {
        C() : CalculateMultiThread<C>(0, 0, 7, 0) {}
        long long th() { return nthreads; }
        long long sk() { return skip; }
        long long ev() { return every; }
    } cm;
    std::printf("\"cmt\": [%lld, %lld, %lld],\n", cm.th(), cm.sk(), cm.ev());

    // ---- box permutation round trip (reference tests/src/test_lammps2020.cpp:50-58) ----
    double bx[6] = {1, 2, 3, 4, 5, 6}, by[6] = {1, 2, 3, 4, 5, 6};
    BaseTrajectory<Trajectory_numpy>::lammps_to_internal(bx);
    std::vector<double> internal(bx, bx + 6);
    BaseTrajectory<Trajectory_numpy>::internal_to_lammps(bx);
    bool same = true;
    for (int i = 0; i < 6; ++i) same = same && bx[i] == by[i];
    print_vec("internal_box", internal);
    std::printf("\"box_round_trip\": %s,\n", same ? "true" : "false");

    // ---- TriclinicLammpsCell: M = Q R, Q orthogonal, R upper triangular with a non-negative diagonal ----
    std::mt19937 rng(5);
    std::normal_distribution<double> g(0.0, 1.0);
    double worst_qr = 0, worst_orth = 0, worst_rot = 0;
    bool tri_ok = true;
    for (int rep = 0; rep < 50; ++rep) {
        double M[9];
        for (double &x : M) x = g(rng) * 3;
        TriclinicLammpsCell<double> cell(M);
        double Q[9], box[9];
        cell.getQ(Q);   // column-major
        cell.set_lammps_cell(box, true);
        const double R[3][3] = {{2 * box[3], box[6], box[7]}, {0, 2 * box[4], box[8]}, {0, 0, 2 * box[5]}};
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc) {
                double qr = 0, qq = 0;
                for (int k = 0; k < 3; ++k) {
                    qr += Q[r + 3 * k] * R[k][cc];
                    qq += Q[k + 3 * r] * Q[k + 3 * cc];
                }
                worst_qr = std::fmax(worst_qr, std::fabs(qr - M[3 * r + cc]));
                worst_orth = std::fmax(worst_orth, std::fabs(qq - (r == cc ? 1.0 : 0.0)));
            }
        tri_ok = tri_ok && box[3] >= 0 && box[4] >= 0 && box[5] >= 0 && !cell.isDiagonal();
        // rotate_vec maps the cell vectors (columns of M) onto the columns of R
        for (int cc = 0; cc < 3; ++cc) {
            double v[3] = {M[cc], M[3 + cc], M[6 + cc]};
            cell.rotate_vec(v);
            for (int r = 0; r < 3; ++r) worst_rot = std::fmax(worst_rot, std::fabs(v[r] - R[r][cc]));
        }
    }
    double D[9] = {4, 0, 0, 0, 5, 0, 0, 0, 6};
    TriclinicLammpsCell<double> diag(D);
    std::printf("\"qr_residual\": %.3g, \"q_orthogonality\": %.3g, \"rotate_residual\": %.3g, \"qr_signs_ok\": %s, \"diag_detected\": %s,\n", worst_qr,
                worst_orth, worst_rot, tri_ok ? "true" : "false", diag.isDiagonal() ? "true" : "false");

    // ---- BlockAverageG + MediaVar with the mock calculation over a numpy-style trajectory of 47 frames ----
    const size_t nts = 47, nat = 3;
    std::vector<double> pos(nts * nat * 3, 0.5), boxes(nts * 6);
    for (size_t f = 0; f < nts; ++f) {
        const double row[6] = {0, 1, 0, 1, 0, 1};
        for (int k = 0; k < 6; ++k) boxes[f * 6 + k] = row[k];
    }
    std::vector<int> types = {1, 1, 2};
    Trajectory_numpy traj(pos.data(), nullptr, types.data(), boxes.data(), nts, nat, Trajectory_numpy::BoxFormat::Lammps_ortho, false, false);
    const unsigned int n_b = 5, len = 6, extra = 4;
    BlockAverageG<Trajectory_numpy, Mock, unsigned int, unsigned int> ba(&traj, n_b);
    ba.calculate(len, extra);
    std::vector<double> mean(ba.media()->access_vdata(), ba.media()->access_vdata() + len);
    std::vector<double> var(ba.varianza()->access_vdata(), ba.varianza()->access_vdata() + len);
    std::vector<double> blocks;
    for (size_t ib = 0; ib < n_b; ++ib)
        for (unsigned int k = 0; k < len; ++k) blocks.push_back(Mock::value(ib * ba.block_size(), k));
    print_vec("ba_mean", mean);
    print_vec("ba_var", var);
    print_vec("ba_blocks", blocks);
    std::printf("\"ba_block_size\": %u, \"ba_calls\": [", ba.block_size());
    for (size_t i = 0; i < Mock::calls.size(); ++i) std::printf("%s%zu", i ? ", " : "", Mock::calls[i]);
    std::printf("], \"ntypes\": %zu, \"type_ids\": [%u, %u, %u]\n}\n", traj.get_ntypes(), traj.get_type(0), traj.get_type(1), traj.get_type(2));
    return 0;
}
