// analisi (g(r,t) branch) -- command line front end of the B200-native Gofrt.
//
// Mirrors what `analisi -i <lammps binary> -g <nbin> -F <rmin> <rmax> [-S lmax] [-s skip] [-e every]
// [-B blocks] [-N threads] [-d]` does in the reference (analisi/main.cpp:126-245 option handling,
// :552-585 the g(r,t) branch): open the trajectory, wrap it, block-average Gofrt over -B blocks, print
// the column description and "lag bin mean var ..." rows on stdout (progress and timings on stderr),
// exit code 1 with the message of any std::exception.
//
// The reference parses its ~45 options with boost::program_options; this front end only serves the
// g(r,t) calculation, so it carries a small parser for the options that branch reads (same short and
// long names, `-g 200`, `-g200`, `--gofrt 200`, `--gofrt=200`, multitoken `-F a b`).  Every other
// calculation of the reference is out of scope and is refused with a message.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "analisi/blockaverage.h"
#include "analisi/gofrt.h"
#include "analisi/istogrammaatomiraggio.h"
#include "analisi/msd.h"
#include "analisi/trajectory.h"

namespace {

struct Options {
    std::string input;
    unsigned int gofrt = 0;
    std::vector<double> factors;
    int stop_acf = 0, skip = 1, every = 1, blocknumber = 20, nthreads = 2;
    bool dump = false, help = false, edges = false;
    double neighbour_r = 0.0;
    bool msd = false, msd_cm = false, msd_self = false;
};

const char *kUsage =
    "Program to analyze molecular dynamics trajectories: g(r,t) on NVIDIA B200 GPUs, with block averages.\n\n"
    "Allowed options:\n"
    "  -i [ --input ] arg            input file in binary LAMMPS format: id type xu yu zu vx vy vz\n"
    "  -h [ --help ]                 help message\n"
    "  -g [ --gofrt ] arg (=0)       calculate the distinctive and non distinctive part of the van Hove\n"
    "                                correlation function; the argument is the number of bins of every histogram.\n"
    "                                With -S 1 you get only the traditional g(r). Needs -F rmin rmax\n"
    "  -F [ --factors ] arg          for the g(r) calculation: the interval of distances of the histogram\n"
    "  -S [ --stop ] arg (=0)        maximum number of time lags (0: as many as a block allows)\n"
    "  -s [ --skip ] arg (=1)        distance between consecutive time origins of the average\n"
    "  -e [ --every ] arg (=1)       distance between consecutive time lags\n"
    "  -B [ --blocknumber ] arg (=20) number of blocks for averages and variances (and for reading the trajectory)\n"
    "  -N [ --thread ] arg           accepted for compatibility (the GPUs own the parallelism)\n"
    "  -d [ --dump-block ]           append the histogram of each block to ./gofrt.dump\n"
    "  -q [ --mean-square-displacement ]     compute atomic mean square displacement\n"
    "  -Q [ --mean-square-displacement-cm ]  ... and the square displacement of the centre of mass of each atomic type\n"
    "  --mean-square-displacement-self       ... in the reference system of the centre of mass of the atom's type\n"
    "  --neighbour arg (=0)          calculate the histogram of the neighbours up to the specified distance\n"
    "  --edge-pairs                  (addition) report on stderr, per block, the pairs within 1 ulp of a bin edge\n"
    "Environment: ANALISI_DEVICES=0,1,... selects the GPUs (default: all visible).\n";

struct Spec {
    char shortname;
    const char *longname;
    int nargs;   // 0 switch, 1 one value, -1 multitoken
};
const Spec kSpecs[] = {{'i', "input", 1},   {'h', "help", 0},        {'g', "gofrt", 1},  {'F', "factors", -1},
                       {'S', "stop", 1},    {'s', "skip", 1},        {'e', "every", 1},  {'B', "blocknumber", 1},
                       {'N', "thread", 1},  {'d', "dump-block", 0}, {'\1', "edge-pairs", 0}, {'\2', "neighbour", 1},
                       {'q', "mean-square-displacement", 0}, {'Q', "mean-square-displacement-cm", 0},
                       {'\3', "mean-square-displacement-self", 0}};

// options of the reference that select or tune calculations this front end does not provide
const char *kForeign = "lVvMHaDzukYIEACf";
const char *kForeignLong[] = {"loginput", "vibrational-spectrum", "velocity-histogram", "histogram-minmax",
                              "heat-transport-coefficient", "headers", "dt", "covariance", "subtract-mean",
                              "subtract-mean-start", "subBlock", "kk", "kk-range", "binary-convert",
                              "binary-convert-gromacs", "spherical-harmonics-correlation", "buffer-size", "lt",
                              "cut", "write-mass-currents", "fpe", "test-debug"};

bool looks_like_option(const char *a) {
    if (a[0] != '-' || a[1] == 0) return false;
    if (a[1] == '-') return true;
    // "-3.5" is a value, "-S" an option
    return !(a[1] == '.' || (a[1] >= '0' && a[1] <= '9'));
}

double to_double(const std::string &s, const char *name) {
    char *end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end) throw std::runtime_error(std::string("the argument ('") + s + "') for option '--" + name + "' is invalid");
    return v;
}
long to_long(const std::string &s, const char *name) {
    char *end = nullptr;
    const long v = std::strtol(s.c_str(), &end, 10);
    if (end == s.c_str() || *end) throw std::runtime_error(std::string("the argument ('") + s + "') for option '--" + name + "' is invalid");
    return v;
}

void assign(Options &o, const Spec &sp, const std::vector<std::string> &vals) {
    switch (sp.shortname) {
        case 'i': o.input = vals[0]; break;
        case 'h': o.help = true; break;
        case 'g': {
            const long v = to_long(vals[0], sp.longname);
            if (v < 0) throw std::runtime_error("the argument for option '--gofrt' is invalid");
            o.gofrt = static_cast<unsigned int>(v);
            break;
        }
        case 'F':
            for (const std::string &s : vals) o.factors.push_back(to_double(s, sp.longname));
            break;
        case 'S': o.stop_acf = static_cast<int>(to_long(vals[0], sp.longname)); break;
        case 's': o.skip = static_cast<int>(to_long(vals[0], sp.longname)); break;
        case 'e': o.every = static_cast<int>(to_long(vals[0], sp.longname)); break;
        case 'B': o.blocknumber = static_cast<int>(to_long(vals[0], sp.longname)); break;
        case 'N': o.nthreads = static_cast<int>(to_long(vals[0], sp.longname)); break;
        case 'd': o.dump = true; break;
        case '\1': o.edges = true; break;
        case '\2': o.neighbour_r = to_double(vals[0], sp.longname); break;
        case 'q': o.msd = true; break;
        case 'Q': o.msd_cm = true; break;
        case '\3': o.msd_self = true; break;
    }
}

Options parse(int argc, char **argv) {
    Options o;
    for (int i = 1; i < argc;) {
        const std::string a = argv[i];
        const Spec *sp = nullptr;
        std::vector<std::string> vals;
        if (a.rfind("--", 0) == 0) {
            const size_t eq = a.find('=');
            const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            for (const Spec &s : kSpecs)
                if (name == s.longname) sp = &s;
            if (!sp) {
                for (const char *f : kForeignLong)
                    if (name == f)
                        throw std::runtime_error("option '--" + name + "' belongs to a calculation this build does not provide (it has g(r,t), the neighbour histogram and the MSD)\n");
                throw std::runtime_error("unrecognised option '" + a + "'");
            }
            if (eq != std::string::npos) vals.push_back(a.substr(eq + 1));
            ++i;
        } else if (a.size() >= 2 && a[0] == '-') {
            for (const Spec &s : kSpecs)
                if (a[1] == s.shortname) sp = &s;
            if (!sp) {
                if (std::strchr(kForeign, a[1]))
                    throw std::runtime_error("option '" + a + "' belongs to a calculation this build does not provide (it has g(r,t), the neighbour histogram and the MSD)\n");
                throw std::runtime_error("unrecognised option '" + a + "'");
            }
            if (a.size() > 2) {
                if (sp->nargs == 0) throw std::runtime_error("option '" + a + "' does not take any arguments");
                vals.push_back(a.substr(2));
            }
            ++i;
        } else {
            throw std::runtime_error("too many positional options have been specified on the command line");
        }
        if (sp->nargs != 0) {
            while (i < argc && !looks_like_option(argv[i]) && (sp->nargs == -1 || vals.empty())) vals.push_back(argv[i++]);
            if (vals.empty()) throw std::runtime_error(std::string("the required argument for option '--") + sp->longname + "' is missing");
        }
        assign(o, *sp, vals);
    }
    return o;
}

}  // namespace

int main(int argc, char **argv) {
    std::cerr << "analisi (B200-native g(r,t)) -- " << agofrt_version() << std::endl;
    std::cerr << "arguments (enclosed by '') were:";
    for (int i = 0; i < argc; ++i) std::cerr << " '" << argv[i] << "'";
    std::cerr << std::endl;

    Options o;
    // -N defaults to OMP_NUM_THREADS like the reference (main.cpp:145-158); the value is not used
    if (const char *e = std::getenv("OMP_NUM_THREADS")) {
        const int n = std::atoi(e);
        if (n > 0) o.nthreads = n;
    }
    try {
        const int default_threads = o.nthreads;
        o = parse(argc, argv);
        if (o.nthreads <= 0) o.nthreads = default_threads;
        if (argc <= 1 || o.help || o.skip <= 0 || o.stop_acf < 0 || o.neighbour_r < 0) {
            std::cout << kUsage << "\n";
            return o.help ? 0 : 1;
        }
    } catch (const std::exception &e) {
        std::cerr << e.what() << "\n";
        std::cerr << kUsage << "\n";
        return 1;
    }

    try {
        // branch precedence as in the reference (analisi/main.cpp:520, :552, :620): mean square displacement first
        if (!(o.msd || o.msd_cm || o.msd_self) && o.gofrt > 0) {
            if (o.factors.size() != 2) throw std::runtime_error("You have to specify the distance range with the option -F.\n");
            std::cerr << "Calculation of g(r,t) -- distinctive and non distinctive part of the van Hove function...\n";
            if (o.edges) setenv("ANALISI_EDGE_PAIRS", "1", 1);
            Trajectory tr(o.input);
            tr.set_load_velocities(false);   // g(r,t) reads positions only
            tr.set_pbc_wrap(true);           // the minimum image needs wrapped coordinates (reference main.cpp:558)

            BlockAverage<Gofrt<double, Trajectory>, double, double, unsigned int, unsigned int, unsigned int, unsigned int,
                         unsigned int, bool>
                gofr(&tr, static_cast<unsigned int>(o.blocknumber));
            gofr.calculate(o.factors[0], o.factors[1], o.gofrt, static_cast<unsigned int>(o.stop_acf),
                           static_cast<unsigned int>(o.nthreads), static_cast<unsigned int>(o.skip),
                           static_cast<unsigned int>(o.every), o.dump);

            const unsigned int ncol = static_cast<unsigned int>(tr.get_ntypes() * (tr.get_ntypes() + 1));
            const unsigned int nlag = gofr.media()->lunghezza() / o.gofrt / ncol;
            const unsigned int every = o.every > 0 ? static_cast<unsigned int>(o.every) : 1u;
            std::cout << gofr.puntatoreCalcolo()->get_columns_description();
            for (unsigned int t = 0; t < nlag; t += every) {
                for (unsigned int r = 0; r < o.gofrt; r++) {
                    std::cout << t << " " << r;
                    for (unsigned int c = 0; c < ncol; c++) {
                        const unsigned int k = (t * ncol + c) * o.gofrt + r;
                        std::cout << " " << gofr.media()->elemento(k) << " " << gofr.varianza()->elemento(k);
                    }
                    std::cout << "\n";
                }
                std::cout << "\n\n";
            }
        } else if (o.msd || o.msd_cm || o.msd_self) {
            // reference analisi/main.cpp:520-549
            std::cerr << "Mean square displacement calculation ";
            const unsigned int f_cm = o.msd_cm ? 2 : 1;
            if (o.msd_cm)
                std::cerr << "of the center of mass and of the atoms is beginning...\n";
            else
                std::cerr << " of the atoms is beginning...\n"
                          << (o.msd_self ? "In the reference system of each atomic type center of mass...\n" : "In the cell coordinate system...\n");
            Trajectory test(o.input);   // unwrapped, velocities (hence centres of mass) loaded, as in the reference
            BlockAverage<MSD<Trajectory>, unsigned int, unsigned int, unsigned int, bool, bool, bool> Msd(&test, static_cast<unsigned int>(o.blocknumber));
            Msd.calculate(static_cast<unsigned int>(o.skip), static_cast<unsigned int>(o.stop_acf), static_cast<unsigned int>(o.nthreads),
                          o.msd_cm, o.msd_self, o.dump);
            const unsigned int nt = static_cast<unsigned int>(test.get_ntypes());
            for (unsigned int i = 0; i < Msd.media()->lunghezza() / nt / f_cm; i++) {
                for (unsigned int j = 0; j < nt * f_cm; j++)
                    std::cout << Msd.media()->elemento(i * nt * f_cm + j) << " " << Msd.varianza()->elemento(i * nt * f_cm + j) << " ";
                std::cout << "\n";
            }
        } else if (o.neighbour_r > 0) {
            // reference analisi/main.cpp:620-642: histogram of the number of neighbours within r, per type
            std::cerr << "Beginning of calculation of neighbour histogram\n";
            Trajectory test(o.input);
            test.set_load_velocities(false);
            test.set_pbc_wrap(true);
            IstogrammaAtomiRaggio h(&test, o.neighbour_r, static_cast<unsigned int>(o.skip), static_cast<unsigned int>(o.nthreads));
            const unsigned int nt = static_cast<unsigned int>(test.get_ntimesteps());
            const unsigned int nb = static_cast<unsigned int>(o.blocknumber);
            if (nb == 0 || (nt - 1) / nb == 0) throw std::runtime_error("Cannot divide the trajectory in that many blocks!\n");
            const unsigned int s = (nt - 1) / nb;
            h.reset(s);
            test.set_data_access_block_size(s);
            test.set_access_stride_hint(s);
            for (unsigned int i = 0; i < nb; i++) {
                const unsigned int t = s * i;
                test.set_access_at(t);
                h.calculate(t);
            }
            std::map<unsigned int, unsigned int> *hist = h.get_hist();
            for (unsigned int i = 0; i < test.get_ntypes(); i++) {
                std::cout << "\"" << i << "\"\n";
                for (auto it = hist[i].begin(); it != hist[i].end(); ++it) std::cout << it->first << " " << it->second << "\n";
                std::cout << "\n\n";
            }
        } else {
            throw std::runtime_error("Nothing to do: choose -g <nbin> -F <rmin> <rmax>, --neighbour <r>, -q or -Q.\n");
        }
    } catch (const std::exception &e) {
        std::cerr << e.what() << "\n";
        return 1;
    }
    return 0;
}
