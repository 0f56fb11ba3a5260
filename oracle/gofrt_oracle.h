/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's g(r,t) path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product (libagofrt.so and everything above it) never links, imports or calls it, and has
 * no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
 *   - the reference's golden vectors that really pin the path:
 *       tests/test_gofrt/test_gofr.csv, tests/test_notebook/test_gofr.csv,
 *       tests/data/cli/pair_corr_{t,no_t}, tests/cpp_regression_data/{min_image,pbc_1,pbc_2}
 *     (small fixtures derived from them are committed under tests/golden/ by
 *      tests/golden/make_golden.py), and
 *   - the unmodified reference compiled from /root/reference by oracle/Makefile (oracle/_ref),
 *     for what no golden file pins: triclinic min-image, every>1, ragged skip, block averages.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef GOFRT_ORACLE_H
#define GOFRT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* A window of a trajectory in the reference's internal layout
 * (lib/include/basetrajectory.h:310-323): AoS positions, internal box rows
 * [xlo,ylo,zlo,lx/2,ly/2,lz/2(,xy,xz,yz)], dense type ids. */
typedef struct {
    size_t natoms;
    size_t nframes;       /* frames present in pos/box (the loaded window)            */
    size_t first_frame;   /* absolute index of pos[0] (Trajectory::current_timestep)  */
    size_t total_frames;  /* get_ntimesteps(): used only by the length check          */
    int ntypes;
    int triclinic;        /* box row stride is 9 if set, else 6                        */
    const double *pos;    /* [nframes][natoms][3] */
    const double *box;    /* [nframes][6|9]       */
    const int *type_id;   /* [natoms], values in [0,ntypes) */
} gofrt_oracle_traj;

typedef struct {
    double rmin, rmax;
    unsigned nbin;
    unsigned lmax;     /* "tmax" ctor argument; 0 = no limit           */
    unsigned skip;     /* 0 is coerced to 1 (calculatemultithread.h:44) */
    unsigned every;    /* 0 is coerced to 1 (calculatemultithread.h:45) */
} gofrt_oracle_params;

enum { GOFRT_ORACLE_OK = 0, GOFRT_ORACLE_TOO_SHORT = -1, GOFRT_ORACLE_BAD_ARG = -2 };

/* basetrajectory.h:94-105 / :109-120 */
void gofrt_oracle_lammps_to_internal(double *c);
void gofrt_oracle_internal_to_lammps(double *c);
/* basetrajectory.h:224-268 */
void gofrt_oracle_min_image(double *delta, const double *l_half, const double *xy_xz_yz, int triclinic);
/* basetrajectory.h:200-219; x receives the min-image vector */
double gofrt_oracle_d2(const double *xi, const double *xj, const double *l_half, const double *xy_xz_yz,
                       int triclinic, double *x);
/* basetrajectory.h:145-161, one frame in place */
void gofrt_oracle_pbc_wrap(double *pos_frame, size_t natoms, const double *box_row, int triclinic);
/* basetrajectory.cpp:51-89: sorted distinct raw types -> dense ids; returns ntypes */
int gofrt_oracle_type_ids(const int *raw_types, size_t natoms, int *type_id_out);
/* gofrt.h:86-104 */
unsigned gofrt_oracle_itype(unsigned ntypes, unsigned type1, unsigned type2);
/* gofrt.cpp:37-39 */
unsigned gofrt_oracle_nextra(size_t total_frames, unsigned n_b, unsigned lmax);
/* gofrt.cpp:55 */
unsigned gofrt_oracle_leff(unsigned ntimesteps, unsigned lmax);
/* the reference's binning expression, gofrt.cpp:114-117; returns idx clipped to [-1, nbin] */
int gofrt_oracle_bin(double d2, double rmin, double dr, unsigned nbin);

/* Integer bin counts of one calculate(primo) after reset(ntimesteps)
 * (gofrt.cpp:73-122 + calculatemultithread.h:80-162).
 *   counts     [leff][ntypes*(ntypes+1)][nbin], zeroed here
 *   edge_pairs optional (may be NULL): number of accepted pairs whose d2 is a bin threshold T[k]>0
 *              or its predecessor double, i.e. pairs a 1-ulp change of d2 would move to another bin
 *   nthreads   worker threads used to go faster (split over atoms i); the counts do not depend on it */
int gofrt_oracle_counts(const gofrt_oracle_traj *tr, const gofrt_oracle_params *p, size_t primo,
                        unsigned ntimesteps, uint64_t *counts, uint64_t *edge_pairs, unsigned nthreads);

/* The same calculation accumulating `incr` in double exactly as the reference does with
 * `ref_nthreads` threads (thread-private partial sums over the atom split of
 * calculatemultithread.h:50-104, merged in thread order, gofrt.cpp:126-132): bitwise equal to the
 * reference's vdata for the same thread count. */
int gofrt_oracle_vdata(const gofrt_oracle_traj *tr, const gofrt_oracle_params *p, size_t primo,
                       unsigned ntimesteps, double *vdata, unsigned ref_nthreads);

/* MediaVar (calcoliblocchi.h:25-61): Welford mean / variance-of-the-mean over n_b blocks of
 * `len` values laid out blocks[n_b][len]. */
void gofrt_oracle_mediavar(const double *blocks, unsigned n_b, size_t len, double *mean, double *var);

/* Next row of the scope table: the neighbour-count histogram of
 * IstogrammaAtomiRaggio::calculate (lib/src/istogrammaatomiraggio.cpp:31-85, driven by `analisi --neighbour r`,
 * analisi/main.cpp:620-642).  For every frame tstart, tstart+skip, ... < tstart+ntimesteps and every atom i:
 * cont[type(j)]++ for every j (j == i included) with d2_minImage(i,j,frame,frame) < r*r, then
 * hist[type][cont[type]]++ for every type.  hist is [ntypes][natoms+1], ADDED to (the reference's maps
 * accumulate over calculate() calls).  Pinned by the reference's golden text tests/data/cli/neighbours. */
int gofrt_oracle_neighbour_hist(const gofrt_oracle_traj *tr, double r, size_t tstart, unsigned ntimesteps,
                                unsigned skip, uint64_t *hist, unsigned nthreads);

/* Scope table rank 3: mean square displacement, MSD<T,FPE>::calc_single_th (lib/src/msd.cpp:63-125) driven by
 * CalculateMultiThread (PARALLEL_SPLIT_TIME: lags are split over threads, every lag is computed by ONE thread in
 * the order origins-then-atoms, so the result does not depend on the thread count).
 *   cm [nframes][ntypes][3]: per-type centres of mass of the window frames (positions_cm), needed when cm_msd or
 *                            cm_self is set (may be NULL otherwise)
 *   out [leff][f_cm][ntypes], f_cm = cm_msd ? 2 : 1: the running means vdata[t][..] exactly as the reference forms them
 * Per-type centre of mass of one frame = running mean over the atoms in index order
 * (lib/src/trajectory_numpy.cpp:201-223; lib/src/trajectory.cpp:648-657 uses the file order of the atoms). */
void gofrt_oracle_cm(const double *pos_frame, const int *type_id, size_t natoms, int ntypes, double *cm_out);
int gofrt_oracle_msd(const gofrt_oracle_traj *tr, const double *cm, size_t primo, unsigned ntimesteps, unsigned lmax,
                     unsigned skip, int cm_msd, int cm_self, double *out);

#ifdef __cplusplus
}
#endif
#endif
