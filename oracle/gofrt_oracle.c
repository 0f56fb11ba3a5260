/* TEST INFRASTRUCTURE ONLY -- see gofrt_oracle.h.  Parity status: PINNED (goldens + oracle/_ref).
 *
 * Plain-C restatement of the reference's g(r,t) path.  Compile with -ffp-contract=off and without
 * -ffast-math / -march (oracle/Makefile): the reference object has no FMA and rounds every
 * operation separately, and so must this file.
 */
#include "gofrt_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- box permutation: lib/include/basetrajectory.h:94-105 ---------------------------------- */
void gofrt_oracle_lammps_to_internal(double *c) {
    /* [xlo,xhi,ylo,yhi,zlo,zhi] -> [xlo,ylo,zlo,(xhi-xlo)/2,(yhi-ylo)/2,(zhi-zlo)/2] */
    const double xlo = c[0], xhi = c[1], ylo = c[2], yhi = c[3], zlo = c[4], zhi = c[5];
    c[0] = xlo;
    c[1] = ylo;
    c[2] = zlo;
    c[3] = (xhi - xlo) / 2;
    c[4] = (yhi - ylo) / 2;
    c[5] = (zhi - zlo) / 2;
}

/* ---- lib/include/basetrajectory.h:109-120 ---------------------------------------------------- */
void gofrt_oracle_internal_to_lammps(double *c) {
    const double xlo = c[0], ylo = c[1], zlo = c[2], hx = c[3], hy = c[4], hz = c[5];
    c[0] = xlo;
    c[1] = xlo + hx * 2;
    c[2] = ylo;
    c[3] = hy * 2 + ylo;
    c[4] = zlo;
    c[5] = hz * 2 + zlo;
}

/* ---- minimum image: lib/include/basetrajectory.h:224-268 ------------------------------------
 * Sequential z -> y -> x; each dimension is a `while`, the tilt factors are added to the lower
 * dimensions inside the loop body, and every += / -= rounds on its own. */
void gofrt_oracle_min_image(double *delta, const double *l_half, const double *xy_xz_yz, int triclinic) {
    double xy = 0.0, xz = 0.0, yz = 0.0;
    if (triclinic) {
        xy = xy_xz_yz[0];
        xz = xy_xz_yz[1];
        yz = xy_xz_yz[2];
    }
    while (fabs(delta[2]) > l_half[2]) {
        if (delta[2] < 0.0) {
            delta[2] += l_half[2] * 2;
            if (triclinic) {
                delta[1] += yz;
                delta[0] += xz;
            }
        } else {
            delta[2] -= l_half[2] * 2;
            if (triclinic) {
                delta[1] -= yz;
                delta[0] -= xz;
            }
        }
    }
    while (fabs(delta[1]) > l_half[1]) {
        if (delta[1] < 0.0) {
            delta[1] += l_half[1] * 2;
            if (triclinic) delta[0] += xy;
        } else {
            delta[1] -= l_half[1] * 2;
            if (triclinic) delta[0] -= xy;
        }
    }
    while (fabs(delta[0]) > l_half[0]) {
        if (delta[0] < 0.0)
            delta[0] += l_half[0] * 2;
        else
            delta[0] -= l_half[0] * 2;
    }
}

/* ---- lib/include/basetrajectory.h:200-219 ---------------------------------------------------- */
double gofrt_oracle_d2(const double *xi, const double *xj, const double *l_half, const double *xy_xz_yz,
                       int triclinic, double *x) {
    double d2 = 0.0;
    for (int k = 0; k < 3; ++k) x[k] = xi[k] - xj[k];
    gofrt_oracle_min_image(x, l_half, xy_xz_yz, triclinic);
    for (int k = 0; k < 3; ++k) d2 += x[k] * x[k];
    return d2;
}

/* ---- lib/include/basetrajectory.h:145-161 ----------------------------------------------------
 * The wrap is around the half edges (box row +3), it ignores xlo/ylo/zlo. */
void gofrt_oracle_pbc_wrap(double *pos_frame, size_t natoms, const double *box_row, int triclinic) {
    const double mid[3] = {box_row[3], box_row[4], box_row[5]};
    for (size_t a = 0; a < natoms; ++a) {
        double *xa = pos_frame + 3 * a;
        for (int k = 0; k < 3; ++k) xa[k] = xa[k] - mid[k];
        gofrt_oracle_min_image(xa, box_row + 3, box_row + 6, triclinic);
        for (int k = 0; k < 3; ++k) xa[k] = xa[k] + mid[k];
    }
}

/* ---- lib/src/basetrajectory.cpp:51-89 -------------------------------------------------------- */
static int cmp_int(const void *a, const void *b) {
    const int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}
int gofrt_oracle_type_ids(const int *raw, size_t natoms, int *ids) {
    if (natoms == 0) return 0;
    int *sorted = (int *)malloc(natoms * sizeof(int));
    memcpy(sorted, raw, natoms * sizeof(int));
    qsort(sorted, natoms, sizeof(int), cmp_int);
    size_t nt = 0;
    for (size_t i = 0; i < natoms; ++i)
        if (i == 0 || sorted[i] != sorted[i - 1]) sorted[nt++] = sorted[i];
    for (size_t i = 0; i < natoms; ++i) {
        size_t lo = 0, hi = nt;
        while (hi - lo > 1) {
            size_t m = (lo + hi) / 2;
            if (sorted[m] <= raw[i]) lo = m; else hi = m;
        }
        ids[i] = (int)lo;
    }
    free(sorted);
    return (int)nt;
}

/* ---- lib/include/gofrt.h:86-104 -------------------------------------------------------------- */
unsigned gofrt_oracle_itype(unsigned ntypes, unsigned type1, unsigned type2) {
    if (type2 < type1) {
        unsigned t = type2;
        type2 = type1;
        type1 = t;
    }
    return ntypes * (ntypes + 1) / 2 - (type2 + 1) * (type2 + 2) / 2 + type1;
}

/* ---- lib/src/gofrt.cpp:37-39 ----------------------------------------------------------------- */
unsigned gofrt_oracle_nextra(size_t total_frames, unsigned n_b, unsigned lmax) {
    const unsigned a = (unsigned)(total_frames / (n_b + 1) + 1);
    return (a < lmax || lmax == 0) ? a : lmax;
}

/* ---- lib/src/gofrt.cpp:55 -------------------------------------------------------------------- */
unsigned gofrt_oracle_leff(unsigned ntimesteps, unsigned lmax) {
    return (ntimesteps < lmax || lmax == 0) ? ntimesteps : lmax;
}

/* ---- lib/src/gofrt.cpp:114-117 ---------------------------------------------------------------
 * sqrt in double, subtract and divide in double, ROUND TO FLOAT, floorf, then int. */
int gofrt_oracle_bin(double d2, double rmin, double dr, unsigned nbin) {
    const double d = sqrt(d2);
    const float q = (float)((d - rmin) / dr);
    const float f = floorf(q);
    if (!(f >= 0.0f)) return -1;              /* negative or NaN: rejected by idx>=0 */
    if (f >= (float)nbin) return (int)nbin;   /* rejected by idx<nbin */
    return (int)f;
}

/* ---- the pair loop --------------------------------------------------------------------------- */
typedef struct {
    const gofrt_oracle_traj *tr;
    const gofrt_oracle_params *p;
    size_t primo;
    unsigned ntimesteps, leff, skip, every;
    double rmin2, rmax2, dr, incr;
    size_t a0, a1;        /* atom range of this worker                         */
    uint64_t *counts;     /* integer mode: private [leff][2P][nbin], or NULL   */
    double *acc;          /* float mode: private [leff][2P][nbin], or NULL     */
    uint64_t edges;
} worker_t;

static inline const double *frame_pos(const gofrt_oracle_traj *tr, size_t t) {
    /* Trajectory::positions<false>: window-relative (lib/include/trajectory.h:74) */
    return tr->pos + (t - tr->first_frame) * tr->natoms * 3;
}
static inline const double *frame_box(const gofrt_oracle_traj *tr, size_t t) {
    return tr->box + (t - tr->first_frame) * (tr->triclinic ? 9 : 6);
}

static double next_up(double x) { return nextafter(x, INFINITY); }
static double next_down(double x) { return nextafter(x, -INFINITY); }

/* lib/src/gofrt.cpp:95-122 for atoms [a0,a1), looped as calculatemultithread.h:114-115 */
static void *pair_worker(void *arg) {
    worker_t *w = (worker_t *)arg;
    const gofrt_oracle_traj *tr = w->tr;
    const unsigned nt = (unsigned)tr->ntypes, nbin = w->p->nbin;
    const unsigned P = nt * (nt + 1) / 2;
    const size_t N = tr->natoms;
    const double rmin = w->p->rmin;
    for (unsigned t = 0; t < w->leff; t += w->every) {
        for (unsigned im = 0; im < w->ntimesteps; im += w->skip) {
            const size_t fi = w->primo + im, fj = w->primo + im + t;
            const double *pi = frame_pos(tr, fi), *pj = frame_pos(tr, fj);
            const double *bx = frame_box(tr, fi); /* the box of frame fi for both atoms */
            for (size_t i = w->a0; i < w->a1; ++i) {
                for (size_t j = 0; j < N; ++j) {
                    double x[3];
                    const double d2 = gofrt_oracle_d2(pi + 3 * i, pj + 3 * j, bx + 3, bx + 6, tr->triclinic, x);
                    if (d2 > w->rmax2 || d2 < w->rmin2) continue;
                    unsigned slot = gofrt_oracle_itype(nt, (unsigned)tr->type_id[i], (unsigned)tr->type_id[j]);
                    if (i == j) slot += P;
                    const int idx = gofrt_oracle_bin(d2, rmin, w->dr, nbin);
                    if (w->counts && d2 > 0.0) {
                        /* 1-ulp neighbours fall in another bin <=> d2 is a threshold or its predecessor */
                        const int up = gofrt_oracle_bin(next_up(d2), rmin, w->dr, nbin);
                        const int dn = gofrt_oracle_bin(next_down(d2), rmin, w->dr, nbin);
                        if (up != idx || dn != idx) w->edges++;
                    }
                    if (idx < (int)nbin && idx >= 0) {
                        const size_t k = ((size_t)t * 2 * P + slot) * nbin + (size_t)idx;
                        if (w->counts) w->counts[k] += 1;
                        if (w->acc) w->acc[k] += w->incr;
                    }
                }
            }
        }
    }
    return NULL;
}

static int setup(worker_t *w, const gofrt_oracle_traj *tr, const gofrt_oracle_params *p, size_t primo,
                 unsigned ntimesteps) {
    if (!tr || !p || p->nbin == 0 || tr->ntypes <= 0) return GOFRT_ORACLE_BAD_ARG;
    memset(w, 0, sizeof(*w));
    w->tr = tr;
    w->p = p;
    w->primo = primo;
    w->ntimesteps = ntimesteps;
    w->leff = gofrt_oracle_leff(ntimesteps, p->lmax);
    w->skip = p->skip ? p->skip : 1;
    w->every = p->every ? p->every : 1;
    /* gofrt.cpp:27-29 */
    w->dr = (p->rmax - p->rmin) / p->nbin;
    w->rmax2 = p->rmax * p->rmax;
    w->rmin2 = p->rmin * p->rmin;
    /* gofrt.cpp:81-83 */
    if ((size_t)w->leff + ntimesteps + primo > tr->total_frames + 1) return GOFRT_ORACLE_TOO_SHORT;
    /* the loop reads frames up to primo + (last origin) + (last lag) from the loaded window */
    if (ntimesteps > 0 && w->leff > 0) {
        const size_t last = primo + (size_t)((ntimesteps - 1) / w->skip) * w->skip +
                            (size_t)((w->leff - 1) / w->every) * w->every;
        if (primo < tr->first_frame || last >= tr->first_frame + tr->nframes) return GOFRT_ORACLE_BAD_ARG;
    }
    /* gofrt.cpp:91-92 */
    if (ntimesteps / w->skip > 0)
        w->incr = 1.0 / (int)(ntimesteps / w->skip);
    else
        w->incr = 1;
    return GOFRT_ORACLE_OK;
}

int gofrt_oracle_counts(const gofrt_oracle_traj *tr, const gofrt_oracle_params *p, size_t primo,
                        unsigned ntimesteps, uint64_t *counts, uint64_t *edge_pairs, unsigned nthreads) {
    worker_t proto;
    const int rc = setup(&proto, tr, p, primo, ntimesteps);
    if (rc != GOFRT_ORACLE_OK) return rc;
    const unsigned nt = (unsigned)tr->ntypes;
    const size_t len = (size_t)proto.leff * nt * (nt + 1) * p->nbin;
    memset(counts, 0, len * sizeof(uint64_t));
    if (edge_pairs) *edge_pairs = 0;
    if (nthreads == 0) nthreads = 1;
    if (nthreads > tr->natoms && tr->natoms > 0) nthreads = (unsigned)tr->natoms;

    worker_t *ws = (worker_t *)calloc(nthreads, sizeof(worker_t));
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    const size_t per = tr->natoms / nthreads;
    for (unsigned k = 0; k < nthreads; ++k) {
        ws[k] = proto;
        ws[k].a0 = k * per;
        ws[k].a1 = (k == nthreads - 1) ? tr->natoms : (k + 1) * per;
        ws[k].counts = (k == 0) ? counts : (uint64_t *)calloc(len ? len : 1, sizeof(uint64_t));
    }
    for (unsigned k = 1; k < nthreads; ++k) pthread_create(&th[k], NULL, pair_worker, &ws[k]);
    pair_worker(&ws[0]);
    for (unsigned k = 1; k < nthreads; ++k) pthread_join(th[k], NULL);
    for (unsigned k = 0; k < nthreads; ++k) {
        if (k > 0) {
            for (size_t i = 0; i < len; ++i) counts[i] += ws[k].counts[i];
            free(ws[k].counts);
        }
        if (edge_pairs) *edge_pairs += ws[k].edges;
    }
    free(ws);
    free(th);
    return GOFRT_ORACLE_OK;
}

int gofrt_oracle_vdata(const gofrt_oracle_traj *tr, const gofrt_oracle_params *p, size_t primo,
                       unsigned ntimesteps, double *vdata, unsigned ref_nthreads) {
    worker_t proto;
    const int rc = setup(&proto, tr, p, primo, ntimesteps);
    if (rc != GOFRT_ORACLE_OK) return rc;
    const unsigned nt = (unsigned)tr->ntypes;
    const size_t len = (size_t)proto.leff * nt * (nt + 1) * p->nbin;
    if (ref_nthreads == 0) ref_nthreads = 1; /* calculatemultithread.h:43, gofrt.cpp:75-78 */
    for (size_t i = 0; i < len; ++i) vdata[i] = 0; /* azzera(), gofrt.cpp:87 */

    /* calculatemultithread.h:50-104 with PARALLEL_SPLIT_ATOM: npassith = natoms/nthreads,
     * the last thread runs to natoms.  Thread 0 accumulates straight into vdata, the others into
     * th_data (gofrt.cpp:96-100).  Here every worker really is a thread; the per-thread order of
     * the additions is the reference's (t outer, origin inner). */
    worker_t *ws = (worker_t *)calloc(ref_nthreads, sizeof(worker_t));
    pthread_t *th = (pthread_t *)calloc(ref_nthreads, sizeof(pthread_t));
    const size_t per = tr->natoms / ref_nthreads;
    for (unsigned k = 0; k < ref_nthreads; ++k) {
        ws[k] = proto;
        ws[k].a0 = k * per;
        ws[k].a1 = (k == ref_nthreads - 1) ? tr->natoms : (k + 1) * per;
        ws[k].acc = (k == 0) ? vdata : (double *)calloc(len ? len : 1, sizeof(double));
    }
    for (unsigned k = 1; k < ref_nthreads; ++k) pthread_create(&th[k], NULL, pair_worker, &ws[k]);
    pair_worker(&ws[0]);
    for (unsigned k = 1; k < ref_nthreads; ++k) pthread_join(th[k], NULL);
    /* calc_end, gofrt.cpp:126-132: vdata += th_data[ith] in thread order */
    for (unsigned k = 1; k < ref_nthreads; ++k) {
        for (size_t i = 0; i < len; ++i) vdata[i] += ws[k].acc[i];
        free(ws[k].acc);
    }
    free(ws);
    free(th);
    return GOFRT_ORACLE_OK;
}

/* ---- lib/include/calcoliblocchi.h:25-61 with the VectorOp algebra of operazionisulista.h ------ */
void gofrt_oracle_mediavar(const double *blocks, unsigned n_b, size_t len, double *mean, double *var) {
    for (size_t i = 0; i < len; ++i) mean[i] = var[i] = 0.0;
    for (unsigned ib = 0; ib < n_b; ++ib) {
        const double *x = blocks + (size_t)ib * len;
        const double div = (double)(ib + 1);
        for (size_t i = 0; i < len; ++i) {
            const double delta = x[i] - mean[i];
            const double step = delta / div;
            mean[i] += step;
            double tmp = x[i] - mean[i];
            tmp *= delta;
            var[i] += tmp;
        }
    }
    const double norm = (double)((n_b - 1) * n_b);
    for (size_t i = 0; i < len; ++i) var[i] /= norm;
}


/* ---- lib/src/istogrammaatomiraggio.cpp:31-85 --------------------------------------------------
 * hist[type][count] += 1 per atom and frame; threads split the atoms as the reference does (the
 * result does not depend on the split: integer counts). */
typedef struct {
    const gofrt_oracle_traj *tr;
    double r2;
    size_t tstart;
    unsigned ntimesteps, skip;
    size_t a0, a1;
    uint64_t *hist; /* private [ntypes][natoms+1] */
} nb_worker_t;

static void *nb_worker(void *arg) {
    nb_worker_t *w = (nb_worker_t *)arg;
    const gofrt_oracle_traj *tr = w->tr;
    const size_t N = tr->natoms, stride = tr->triclinic ? 9 : 6;
    const int nt = tr->ntypes;
    unsigned *cont = (unsigned *)calloc((size_t)nt, sizeof(unsigned));
    for (size_t f = w->tstart; f < w->tstart + w->ntimesteps; f += w->skip) {
        const size_t rel = f - tr->first_frame;
        const double *p = tr->pos + rel * N * 3;
        const double *bx = tr->box + rel * stride;
        for (size_t i = w->a0; i < w->a1; ++i) {
            for (int k = 0; k < nt; ++k) cont[k] = 0;
            for (size_t j = 0; j < N; ++j) {
                double x[3];
                if (gofrt_oracle_d2(p + 3 * i, p + 3 * j, bx + 3, bx + 6, tr->triclinic, x) < w->r2) cont[tr->type_id[j]]++;
            }
            for (int k = 0; k < nt; ++k) w->hist[(size_t)k * (N + 1) + cont[k]] += 1;
        }
    }
    free(cont);
    return NULL;
}

int gofrt_oracle_neighbour_hist(const gofrt_oracle_traj *tr, double r, size_t tstart, unsigned ntimesteps,
                                unsigned skip, uint64_t *hist, unsigned nthreads) {
    if (!tr || !hist || tr->ntypes <= 0) return GOFRT_ORACLE_BAD_ARG;
    if (skip < 1) skip = 1; /* istogrammaatomiraggio.cpp:19 */
    if (nthreads < 1) nthreads = 1;
    if (ntimesteps == 0 || tr->natoms == 0) return GOFRT_ORACLE_OK;
    const size_t last = tstart + ((size_t)(ntimesteps - 1) / skip) * skip;
    if (tstart < tr->first_frame || last >= tr->first_frame + tr->nframes) return GOFRT_ORACLE_TOO_SHORT;
    const size_t N = tr->natoms, len = (size_t)tr->ntypes * (N + 1);
    if (nthreads > N) nthreads = (unsigned)N;
    nb_worker_t *w = (nb_worker_t *)calloc(nthreads, sizeof(nb_worker_t));
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    const size_t per = N / nthreads;
    for (unsigned k = 0; k < nthreads; ++k) {
        w[k].tr = tr;
        w[k].r2 = r * r; /* istogrammaatomiraggio.cpp:17 */
        w[k].tstart = tstart;
        w[k].ntimesteps = ntimesteps;
        w[k].skip = skip;
        w[k].a0 = per * k;
        w[k].a1 = (k != nthreads - 1) ? per * (k + 1) : N;
        w[k].hist = (uint64_t *)calloc(len, sizeof(uint64_t));
        pthread_create(&th[k], NULL, nb_worker, &w[k]);
    }
    for (unsigned k = 0; k < nthreads; ++k) {
        pthread_join(th[k], NULL);
        for (size_t q = 0; q < len; ++q) hist[q] += w[k].hist[q];
        free(w[k].hist);
    }
    free(w);
    free(th);
    return GOFRT_ORACLE_OK;
}


/* ---- lib/src/trajectory_numpy.cpp:201-223 ----------------------------------------------------- */
void gofrt_oracle_cm(const double *pos_frame, const int *type_id, size_t natoms, int ntypes, double *cm_out) {
    unsigned *cont = (unsigned *)calloc((size_t)ntypes, sizeof(unsigned));
    for (int k = 0; k < 3 * ntypes; ++k) cm_out[k] = 0.0;
    for (size_t a = 0; a < natoms; ++a) {
        const int ty = type_id[a];
        cont[ty]++;
        for (int c = 0; c < 3; ++c) cm_out[3 * ty + c] += (pos_frame[3 * a + c] - cm_out[3 * ty + c]) / (double)cont[ty];
    }
    free(cont);
}

/* ---- lib/src/msd.cpp:41-125 --------------------------------------------------------------------
 * pow(x,2) is x*x exactly (one rounding); the sums are formed left to right as written there. */
int gofrt_oracle_msd(const gofrt_oracle_traj *tr, const double *cm, size_t primo, unsigned ntimesteps, unsigned lmax,
                     unsigned skip, int cm_msd, int cm_self, double *out) {
    if (!tr || !out || tr->ntypes <= 0) return GOFRT_ORACLE_BAD_ARG;
    if ((cm_msd || cm_self) && !cm) return GOFRT_ORACLE_BAD_ARG;
    if (skip < 1) skip = 1;
    const unsigned leff = (ntimesteps < lmax || lmax == 0) ? ntimesteps : lmax; /* msd.cpp:42 */
    const size_t N = tr->natoms, nt = (size_t)tr->ntypes, f_cm = cm_msd ? 2 : 1;
    if ((size_t)leff + ntimesteps + primo > tr->total_frames) return GOFRT_ORACLE_TOO_SHORT; /* msd.cpp:54 */
    if (leff == 0) return GOFRT_ORACLE_OK;
    if (primo < tr->first_frame || primo + (size_t)(ntimesteps - 1) / skip * skip + (leff - 1) >= tr->first_frame + tr->nframes)
        return GOFRT_ORACLE_TOO_SHORT;
    uint64_t *cont = (uint64_t *)calloc(nt * f_cm, sizeof(uint64_t));
    for (size_t t = 0; t < leff; ++t) {
        double *v = out + nt * t * f_cm;
        for (size_t i = 0; i < nt * f_cm; ++i) {
            v[i] = 0.0;
            cont[i] = 0;
        }
        for (size_t im = 0; im < ntimesteps; im += skip) {
            const size_t fa = primo + im - tr->first_frame, fb = fa + t;
            const double *pa = tr->pos + fa * N * 3, *pb = tr->pos + fb * N * 3;
            const double *ca = cm ? cm + fa * nt * 3 : NULL, *cb = cm ? cm + fb * nt * 3 : NULL;
            for (size_t a = 0; a < N; ++a) {
                const size_t ty = (size_t)tr->type_id[a];
                double d[3];
                for (int c = 0; c < 3; ++c) {
                    d[c] = pa[3 * a + c] - pb[3 * a + c];
                    if (cm_self) d[c] = d[c] - (ca[3 * ty + c] - cb[3 * ty + c]);
                }
                const double delta = (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) - v[ty];
                v[ty] += delta / (double)(++cont[ty]);
            }
            if (cm_msd) {
                for (size_t ty = 0; ty < nt; ++ty) {
                    double d[3];
                    for (int c = 0; c < 3; ++c) d[c] = ca[3 * ty + c] - cb[3 * ty + c];
                    const double delta = (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) - v[nt + ty];
                    v[nt + ty] += delta / (double)(++cont[nt + ty]);
                }
            }
        }
    }
    free(cont);
    return GOFRT_ORACLE_OK;
}
