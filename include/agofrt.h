/* agofrt.h -- C ABI of libagofrt.so: B200-native (sm_100a) g(r,t) for rikigigi/analisi.
 *
 * This is the drop-in boundary.  The only callers are the host C++ classes of this repository that
 * mirror the reference's API (include/analisi/...: Gofrt, BlockAverageG, Trajectory,
 * Trajectory_numpy) and, for tests and bench.py, a ctypes binding.  Plain pointers and sizes; no
 * C++ or torch types; nothing throws across it.  There is NO CPU fallback behind it: every entry
 * point that needs a GPU returns AGOFRT_ERR_CUDA when none is usable.
 *
 * What each group replaces in the reference (paths relative to the reference tree):
 *
 *   agofrt_traj_*      the window buffers of BaseTrajectory (lib/include/basetrajectory.h:310-323)
 *                      as filled by Trajectory::set_access_at (lib/src/trajectory.cpp:450-690) or
 *                      the Trajectory_numpy ctor (lib/src/trajectory_numpy.cpp:7-199): positions
 *                      [frame][atom][3] float64, one internal box row per frame
 *                      [xlo,ylo,zlo,lx/2,ly/2,lz/2(,xy,xz,yz)], dense type ids
 *                      (lib/src/basetrajectory.cpp:51-89).
 *   agofrt_plan_*      the Gofrt constructor state: rmin, rmax, nbin and the derived dr, rmin2,
 *                      rmax2 (lib/src/gofrt.cpp:22-31).
 *   agofrt_block       one CalculateMultiThread::calculate(primo) of a Gofrt after reset(ntimesteps)
 *                      (lib/include/calculatemultithread.h:106-162 driving
 *                      Gofrt::calc_init / calc_single_th / calc_end, lib/src/gofrt.cpp:73-155, which
 *                      call BaseTrajectory::d2_minImage, lib/include/basetrajectory.h:168-268).
 *                      It returns INTEGER bin counts; the host shim multiplies by `incr`
 *                      (lib/src/gofrt.cpp:91-92).
 *   agofrt_comm_*      the block exchange of Mp::send_to_root / recv_root
 *                      (lib/include/mp.h:35-41, used by lib/include/blockaverage.h:146-186),
 *                      replaced by one NCCL all-reduce of the integer histograms.
 *   agofrt_pbc_wrap    BaseTrajectory::pbc_wrap<TRICLINIC> (lib/include/basetrajectory.h:145-161).
 *
 * Conventions: every function returns 0 (AGOFRT_OK) or a negative agofrt_status; the message of the
 * last failure on the calling thread is agofrt_last_error().  Host pointers passed in are only read
 * (or written, for outputs) during the call; no ownership is transferred.  All calls on one context
 * must come from one host thread at a time, with one exception: agofrt_traj_upload / agofrt_traj_upload_wrap
 * on one window may run on a second host thread while the first is inside agofrt_block on ANOTHER window.
 */
#ifndef AGOFRT_H
#define AGOFRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AGOFRT_API __attribute__((visibility("default")))
#else
#define AGOFRT_API
#endif

typedef enum {
    AGOFRT_OK = 0,
    AGOFRT_ERR_ARG = -1,       /* invalid argument                                              */
    AGOFRT_ERR_CUDA = -2,      /* CUDA runtime / driver failure, or no usable device            */
    AGOFRT_ERR_WINDOW = -3,    /* the block needs frames that are not in the uploaded window    */
    AGOFRT_ERR_NCCL = -4,      /* NCCL missing or failed                                        */
    AGOFRT_ERR_NONFINITE = -5, /* an infinite coordinate or box entry (the reference would spin) */
    AGOFRT_ERR_TOO_LARGE = -6, /* histogram does not fit the shared-memory budget               */
    AGOFRT_ERR_INTERNAL = -7,
    AGOFRT_ERR_RETYPED = -8    /* agofrt_traj_upload_records: an atom's type differs from the table's (the reference
                                  warns and goes on, lib/src/trajectory.cpp:640-646: read the window on the host) */
} agofrt_status;

typedef struct agofrt_ctx agofrt_ctx;    /* a set of GPUs of this process (+ optional peers)   */
typedef struct agofrt_traj agofrt_traj;  /* device-resident trajectory window                  */
typedef struct agofrt_plan agofrt_plan;  /* Gofrt parameters + threshold table on the devices  */

/* ---- library ------------------------------------------------------------------------------- */
AGOFRT_API const char *agofrt_version(void);
AGOFRT_API const char *agofrt_last_error(void);
AGOFRT_API int agofrt_device_count(int *count);

/* Pinned host memory for window buffers (the reference uses fftw_malloc only as an aligned
 * allocator, lib/src/trajectory.cpp:362-371).  Buffers of 1 MiB .. 256 MiB given back with agofrt_host_free are kept
 * page-locked (at most 8, 512 MiB in all) and handed to the next request of the same size. */
AGOFRT_API int agofrt_host_alloc(void **ptr, size_t bytes);
AGOFRT_API int agofrt_host_free(void *ptr);

/* ---- context ------------------------------------------------------------------------------- */
/* devices == NULL, ndev == 0: device 0 only.  ndev == -1: every visible device. */
AGOFRT_API int agofrt_ctx_create(agofrt_ctx **ctx, const int *devices, int ndev);
AGOFRT_API int agofrt_ctx_destroy(agofrt_ctx *ctx);
AGOFRT_API int agofrt_ctx_ndev(const agofrt_ctx *ctx);

#define AGOFRT_COMM_ID_BYTES 128
/* Multi-process operation (one process per GPU, e.g. under torchrun): rank 0 calls
 * agofrt_comm_unique_id, the bytes are broadcast by whatever launcher plumbing exists, and every
 * process calls agofrt_comm_join with the rank of its first device and the total device count.
 * Afterwards agofrt_block shards its work over all `world` devices and all-reduces the counts.
 * A context with several local devices and no join is its own single-process communicator. */
AGOFRT_API int agofrt_comm_unique_id(char id[AGOFRT_COMM_ID_BYTES]);
AGOFRT_API int agofrt_comm_join(agofrt_ctx *ctx, const char id[AGOFRT_COMM_ID_BYTES], int first_rank, int world);
/* Shard geometry without any communicator (tests; "weak scaling" replicas): the context then
 * computes only its share of the work units and returns PARTIAL counts. */
AGOFRT_API int agofrt_ctx_set_shard(agofrt_ctx *ctx, int first_rank, int world);

/* The contiguous share [begin,end) of `units` equal work units that device `rank` of `world`
 * computes (pure host arithmetic; agofrt_block uses exactly this). */
AGOFRT_API int agofrt_shard_range(uint64_t units, int rank, int world, uint64_t *begin, uint64_t *end);
/* The geometry of agofrt_blocks: what rank `rank` of `world` takes of block `block` of `nblocks`, as the part
 * [*part_a, *part_b) / world of the block's work units (0, 0: nothing; 0, world: the whole block).  Rank r owns the
 * stretch [r, r + 1) * nblocks / world of the batch: whole blocks when nblocks is a multiple of world (the reference's
 * dealing of blocks to MPI ranks, lib/include/blockaverage.h:146-186), whole blocks and parts of blocks otherwise. */
AGOFRT_API int agofrt_block_share(unsigned nblocks, int rank, int world, unsigned block, unsigned *part_a, unsigned *part_b);

/* ---- trajectory window --------------------------------------------------------------------- */
/* type_id[natoms] are dense ids in [0,ntypes) (BaseTrajectory::get_type).  box_stride is 6
 * (orthorhombic) or 9 (triclinic, BaseTrajectory::is_triclinic).  max_frames bounds the windows
 * that will be uploaded (Trajectory::set_data_access_block_size). */
AGOFRT_API int agofrt_traj_create(agofrt_traj **traj, agofrt_ctx *ctx, size_t natoms, int box_stride,
                                  const int *type_id, int ntypes, size_t max_frames);
AGOFRT_API int agofrt_traj_destroy(agofrt_traj *traj);
/* Replace the device window by frames [first_frame, first_frame+nframes): pos_aos is
 * [nframes][natoms][3], box_internal [nframes][box_stride].  The device keeps the atoms in a
 * type-major, spatially sorted, padded SoA layout (a permutation: counts do not depend on it). */
AGOFRT_API int agofrt_traj_upload(agofrt_traj *traj, size_t first_frame, size_t nframes,
                                  const double *pos_aos, const double *box_internal);
/* The same with BaseTrajectory::pbc_wrap (lib/include/basetrajectory.h:145-161) applied ON THE DEVICE before the
 * layout change: pos_aos_inout holds the unwrapped frames on entry and the wrapped ones on return (what
 * Trajectory::set_access_at leaves in its window when wrapping is on, lib/src/trajectory.cpp:664-670).
 * Uploads run on a stream of their own: a second window of the same context may be uploaded by another host
 * thread while agofrt_block works on the first one (read-ahead of the next block). */
AGOFRT_API int agofrt_traj_upload_wrap(agofrt_traj *traj, size_t first_frame, size_t nframes,
                                       double *pos_aos_inout, const double *box_internal);
/* The general upload.  flags:
 *   AGOFRT_UP_WRAP       BaseTrajectory::pbc_wrap on the device before the layout change (pos_aos is only read);
 *   AGOFRT_UP_WRITEBACK  (with WRAP) the wrapped frames are written to pos_wrapped_out [nframes][natoms][3], which may be
 *                        pos_aos itself -- what agofrt_traj_upload_wrap does;
 *   AGOFRT_UP_SHARED     with a communicator (several local devices, or agofrt_comm_join): the frames of the window are
 *                        dealt to the devices, every device copies / wraps / lays out its share, and the shares are
 *                        exchanged device to device (NCCL over NVLink): the window crosses PCIe once per box, not once
 *                        per GPU.  A collective: under agofrt_comm_join every process calls it with the same arguments;
 *                        a process only needs valid host data for its own share (agofrt_shard_range over nframes) and
 *                        for frame 0, and with WRITEBACK only its share comes back wrapped.
 * Pageable host memory is staged through page-locked slots by several host threads. */
enum { AGOFRT_UP_WRAP = 1, AGOFRT_UP_WRITEBACK = 2, AGOFRT_UP_SHARED = 4 };
AGOFRT_API int agofrt_traj_upload_ex(agofrt_traj *traj, size_t first_frame, size_t nframes, const double *pos_aos,
                                     const double *box_internal, unsigned flags, double *pos_wrapped_out);
/* The window straight from the records of a LAMMPS binary dump: the frame loop of Trajectory::set_access_at
 * (lib/src/trajectory.cpp:593-674) on the device.  Per atom the dump holds 8 doubles, id type x y z vx vy vz
 * (lib/include/lammps_struct.h:105-125), in any order and split over any number of chunks per frame.
 *   agofrt_traj_set_ids         the id -> slot map built from the first frame (slot_to_id[natoms], ids compact enough
 *                               for a flat table) and the raw type of every slot;
 *   agofrt_traj_upload_records  frame f of the window = chunks [frame_chunk[f], frame_chunk[f+1]) of chunk_ptr /
 *                               chunk_atoms (records, e.g. addresses inside the mmap'd file); box_internal as for
 *                               agofrt_traj_upload.  The raw bytes are staged through page-locked slots and copied as
 *                               they are; a kernel resolves every record's id, checks its type and scatters x y z to the
 *                               atom's slot; then wrap / layout / bounds / exchange as in agofrt_traj_upload_ex (same
 *                               flags; AGOFRT_UP_WRITEBACK also without WRAP: pos_out receives the parsed frames).
 * An unknown id is AGOFRT_ERR_ARG, a changed type AGOFRT_ERR_RETYPED (callers then read that window on the host, as the
 * reference does, with its warning).  Velocities are not extracted (g(r,t) never reads them). */
AGOFRT_API int agofrt_traj_set_ids(agofrt_traj *traj, const int *slot_to_id, const int *slot_raw_type);
AGOFRT_API int agofrt_traj_upload_records(agofrt_traj *traj, size_t first_frame, size_t nframes, const void *const *chunk_ptr,
                                          const int *chunk_atoms, const size_t *frame_chunk, const double *box_internal,
                                          unsigned flags, double *pos_out);
/* Frames of the device window back in the caller's atom order (wrapped if they were uploaded with AGOFRT_UP_WRAP): the
 * host classes materialise their host copy with it the first time an accessor needs one. */
AGOFRT_API int agofrt_traj_download(agofrt_traj *traj, size_t first_frame, size_t nframes, double *pos_aos);
/* Per-frame rotation matrices Q (9 doubles, as Trajectory_numpy::get_rotation_matrix hands them out; reference
 * lib/src/trajectory_numpy.cpp:120,131) kept on the devices next to positions and cells. */
AGOFRT_API int agofrt_traj_set_rotation(agofrt_traj *traj, size_t first_frame, size_t nframes, const double *q);
AGOFRT_API int agofrt_traj_get_rotation(agofrt_traj *traj, size_t frame, double *q9);
/* Read one frame back in the caller's atom order (tests: the layout round-trips bit-exactly). */
AGOFRT_API int agofrt_traj_download_frame(agofrt_traj *traj, size_t frame, double *pos_aos);
/* In-place BaseTrajectory::pbc_wrap on a host buffer through the GPU (frames with their own box
 * rows).  Same arithmetic as the reference: x-=l_half; minImage; x+=l_half. */
AGOFRT_API int agofrt_pbc_wrap(agofrt_ctx *ctx, double *pos_aos, size_t nframes, size_t natoms,
                               const double *box_internal, int box_stride);
/* All N^2 (dx,dy,dz,d2) of BaseTrajectory::d2_minImage(i,j,frame_i,frame_j,x) for small N
 * (the probe tests/src/test_trajectory.cpp:21-41 dumps); out is [natoms][natoms][4]. */
AGOFRT_API int agofrt_traj_d2_all(agofrt_traj *traj, size_t frame_i, size_t frame_j, double *out);
/* One BaseTrajectory::d2_minImage(i,j,frame_i,frame_j,x) (lib/include/basetrajectory.h:168-194) on the
 * device copy of the window, atoms in the caller's numbering: out4 = dx,dy,dz,d2.  A probe for the host
 * classes' d2_minImage accessor, not a hot path. */
AGOFRT_API int agofrt_traj_d2_pair(agofrt_traj *traj, size_t atom_i, size_t atom_j, size_t frame_i,
                                   size_t frame_j, double *out4);

/* ---- plan ---------------------------------------------------------------------------------- */
AGOFRT_API int agofrt_plan_create(agofrt_plan **plan, agofrt_traj *traj, double rmin, double rmax,
                                  unsigned nbin);
AGOFRT_API int agofrt_plan_destroy(agofrt_plan *plan);
/* Point the plan at another window of the same context with the same number of types (double-buffered windows). */
AGOFRT_API int agofrt_plan_retarget(agofrt_plan *plan, agofrt_traj *traj);
/* thresholds[nbin+1]: thresholds[k] = the smallest d2 (>=0) whose reference bin index
 * (int)floorf((sqrt(d2)-rmin)/dr) is >= k (+inf if none). */
AGOFRT_API int agofrt_plan_thresholds(const agofrt_plan *plan, double *thresholds);
/* Which float shortcuts of the binning passed their validation on the device for this plan (each may be NULL):
 * the safe-zone guess, its two-floor form (dense windows; needs -rmin/dr to be an integer), and the guard bins a
 * shared-memory histogram row carries below bin 0.  The counts never depend on them. */
AGOFRT_API int agofrt_plan_info(const agofrt_plan *plan, int *safe_zone_ok, int *two_floor_ok, int *guard_bins);

enum {
    AGOFRT_OPT_DEFAULT = 0,
    AGOFRT_OPT_EDGES = 1,          /* also count the pairs within 1 ulp of a bin edge            */
    AGOFRT_OPT_FORCE_GENERAL = 2,  /* never take the single-pass minimum-image kernel            */
    AGOFRT_OPT_NO_AGGREGATE = 4,   /* plain shared atomics instead of __match_any_sync merging   */
    AGOFRT_OPT_AGGREGATE = 8,      /* force warp-aggregated shared atomics                       */
    AGOFRT_OPT_NO_SAFE = 16,       /* bracket every guess against the exact thresholds (no safe-zone shortcut) */
    AGOFRT_OPT_DENSE = 32,         /* force the kernel without the group filter (dense in-range workloads)      */
    AGOFRT_OPT_SPARSE = 64,        /* force the group-filtered kernel meant for sparse in-range workloads      */
    AGOFRT_OPT_NO_UBOX = 128,      /* never pass a constant box as kernel parameter (uniform operands)         */
    AGOFRT_OPT_NO_SMALL = 256,     /* never take the small-system kernel (job ranges per CTA, batches of jobs per warp) */
    AGOFRT_OPT_ON_DEVICE = 512,    /* leave the counts on the device for agofrt_blockavg_push: counts_out may be NULL
                                      and is not written */
    AGOFRT_OPT_SMALL = 1024,       /* take the small-system kernel for up to 512 device slots (default: up to 256, where
                                      it beats the tile kernel) */
    AGOFRT_OPT_SAFE2 = 2048,       /* dense windows: the two-floor form of the safe-zone binning (two round-down FFMAs give the
                                      histogram word and the near-an-edge flag); exact like the others, measured slower */
    AGOFRT_OPT_SKEW = 4096,        /* with AGOFRT_OPT_SAFE2: offset half of the warps by one binning run (A/B measurements) */
    AGOFRT_OPT_EXPLICIT_JOBS = 8192 /* always build and upload the (lag, origin) job list (default: derived on the device when regular) */
};

typedef struct {
    double kernel_ms;          /* device time of the pair kernels, CUDA events on their stream, max over local devices */
    double total_ms;           /* device time of the whole call (zeroing, kernels, all-reduce, read-back) */
    uint64_t pair_evals;       /* pair evaluations this context performed (N^2 per (lag, origin))  */
    uint64_t pair_evals_total; /* ... the whole job over all shards                                */
    uint64_t jobs;             /* (lag, origin) pairs in the whole job                             */
    uint64_t jobs_fast;        /* ... of which proven single-pass minimum image                    */
    uint32_t launches;         /* kernels launched by this call on this context                    */
    uint32_t ndev_local;
    uint32_t world;
    uint32_t kernel_modes;     /* bit m set: a pair kernel of binning mode m ran (0 thresholds, 1 aggregated, 2 edges,
                                  3 safe-zone, 4 safe-zone dense, 5 two-floor dense); bit 8: the small-system kernel ran */
} agofrt_stats;

/* counts_out [leff][ntypes*(ntypes+1)][nbin] (host, uint64): the number of ordered pairs (i,j)
 * per lag, type-pair slot (Gofrt::get_itype; +P for i==j) and bin, summed over the origins
 * primo, primo+skip, ... < primo+ntimesteps and over the lags 0, every, ... < leff.
 * edge_pairs_out (optional): pairs whose d2 is a bin threshold or the double just below one.
 * stats (optional). */
AGOFRT_API int agofrt_block(agofrt_plan *plan, size_t primo, unsigned ntimesteps, unsigned leff,
                            unsigned skip, unsigned every, unsigned options, uint64_t *counts_out,
                            uint64_t *edge_pairs_out, agofrt_stats *stats);

/* Many small blocks at once (BlockAverageG on systems of a few dozen atoms, where one block is a millisecond of kernel
 * and sharding its work units over several GPUs buys nothing): block b = reset(ntimesteps); calculate(primo0 + b*stride),
 * WHOLE blocks dealt to the devices (a contiguous run of blocks per device -- the reference deals blocks to MPI ranks,
 * lib/include/blockaverage.h:146-186; when the blocks do not divide among the devices a device also takes PART of the
 * work units of a block), one host thread per local device issuing its blocks, no host synchronisation between blocks,
 * then every device receives every block (one NCCL all-gather, or an all-reduce of the zero-filled batch when blocks
 * were split).  The window must hold the frames of all the blocks; every block must have a regular job list (the
 * single-pass minimum image proven for its frame range, or AGOFRT_OPT_FORCE_GENERAL), else AGOFRT_ERR_ARG: run them one by
 * one.  The counts stay on the devices: agofrt_plan_block_counts reads one block back, agofrt_blockavg_push_blocks folds
 * them all, in block order, into a device-resident mean / variance; the last block is also what agofrt_plan_last_counts
 * returns.  Under agofrt_comm_join a collective. */
AGOFRT_API int agofrt_blocks(agofrt_plan *plan, size_t primo0, size_t stride, unsigned nblocks, unsigned ntimesteps,
                             unsigned leff, unsigned skip, unsigned every, unsigned options, agofrt_stats *stats);
AGOFRT_API int agofrt_plan_block_counts(agofrt_plan *plan, unsigned block, uint64_t *counts_out, size_t len);

/* ---- block averages on the device (MediaVar) ------------------------------------------------- */
/* MediaVar<T> (lib/include/calcoliblocchi.h:21-65) for T = Gofrt, on the counts agofrt_block leaves on the device:
 * mean and variance-of-the-mean over blocks with the reference's own sequence of rounded operations per element
 *     x = count * incr;  delta = x - mean;  mean += delta / (k+1);  var += (x - mean) * delta      (block k = 0, 1, ...)
 * and var /= (n_b - 1) * n_b at the end, so the result is bit-identical to the host MediaVar fed with the same
 * blocks in the same order -- without reading every block back and without the eight whole-vector VectorOp passes
 * per block of BlockAverageG::calcola_custom (lib/include/blockaverage.h:128-144).  The accumulator belongs to a
 * context and lives on its first device (after the all-reduce every device holds the same counts).
 *   begin = calcola_begin: len = leff * ntypes*(ntypes+1) * nbin elements, zeroed, k = 0;
 *   push  = calculate on the counts of the plan's LAST agofrt_block (which must have produced len words);
 *   end   = calcola_end + read-back into mean_out / var_out (host, [len] doubles each; either may be NULL).
 * agofrt_plan_last_counts copies the counts of the plan's last agofrt_block to the host (what counts_out would
 * have received), for callers that ran the block with AGOFRT_OPT_ON_DEVICE and want one block after all. */
typedef struct agofrt_blockavg agofrt_blockavg;
AGOFRT_API int agofrt_blockavg_create(agofrt_blockavg **acc, agofrt_ctx *ctx);
AGOFRT_API int agofrt_blockavg_destroy(agofrt_blockavg *acc);
AGOFRT_API int agofrt_blockavg_begin(agofrt_blockavg *acc, size_t len);
AGOFRT_API int agofrt_blockavg_push(agofrt_blockavg *acc, agofrt_plan *plan, double incr);
/* every block of the plan's last agofrt_blocks, in block order (one launch; the same rounded operations per element) */
AGOFRT_API int agofrt_blockavg_push_blocks(agofrt_blockavg *acc, agofrt_plan *plan, double incr);
AGOFRT_API int agofrt_blockavg_end(agofrt_blockavg *acc, unsigned n_b, double *mean_out, double *var_out);
AGOFRT_API int agofrt_plan_last_counts(agofrt_plan *plan, uint64_t *counts_out, size_t len);

/* ---- next row of the scope table: the neighbour-count histogram ------------------------------ */
/* IstogrammaAtomiRaggio::calculate (lib/src/istogrammaatomiraggio.cpp:31-85, `analisi --neighbour r`): for the
 * frames tstart, tstart+skip, ... < tstart+ntimesteps of the uploaded window and every atom i, count the atoms j
 * (j == i included) of each type with d2_minImage(i,j,frame,frame) < r*r, then hist[type][count] += 1.
 * hist_inout is [ntypes][natoms+1] (host, uint64) and is ADDED to, as the reference's maps accumulate over calls.
 * Same minimum-image arithmetic and multi-GPU sharding (frames x atom tiles, one all-reduce) as agofrt_block. */
AGOFRT_API int agofrt_neighbour_hist(agofrt_traj *traj, double r, size_t tstart, unsigned ntimesteps, unsigned skip,
                                     uint64_t *hist_inout, agofrt_stats *stats);

/* Mean square displacement: MSD<T>::calculate(primo) after reset(ntimesteps) (lib/src/msd.cpp:41-125; `analisi -q/-Q`,
 * pyanalisi.MeanSquareDisplacement).  out is [leff][f_cm][ntypes] with f_cm = cm_msd ? 2 : 1: row 0 the per-type MSD
 * of the atoms over the origins primo, primo+skip, ... < primo+ntimesteps (in the frame of the type's centre of mass
 * when cm_self), row 1 the MSD of the per-type centres of mass.  The coordinates are used as uploaded (no minimum
 * image, as in the reference); cm_msd / cm_self need agofrt_traj_set_cm for the uploaded window
 * (cm [nframes][ntypes][3] = BaseTrajectory::positions_cm).  The atom rows are a sum / count where the reference keeps
 * a running mean (equal to rounding); the centre-of-mass rows replay the reference's running mean exactly. */
AGOFRT_API int agofrt_traj_set_cm(agofrt_traj *traj, size_t first_frame, size_t nframes, const double *cm);
AGOFRT_API int agofrt_msd(agofrt_traj *traj, size_t primo, unsigned ntimesteps, unsigned leff, unsigned skip,
                          int cm_msd, int cm_self, double *out, agofrt_stats *stats);

/* ---- the other pair loops over d2_minImage: neighbour lists and spherical-harmonic densities of one frame ---------- */
/* Neighbours<T,double>::update_neigh(frame, sort) (lib/src/neighbour.cpp:8-76): for every atom and every type the list of
 * the atoms of that type within the type's cutoff, in ascending atom index (sort: in ascending distance), with the
 * minimum-image vector xi - xj of each.  nneigh[t] / cutoff2[t]: capacity and squared cutoff for type t (the reference's
 * ListSpec).  Output in the reference's own layout: list_out = per type t a block of natoms * (nneigh[t] + 1) words, the
 * count first; rpos_out = per type t a block of natoms * nneigh[t] * 4 doubles (r, x, y, z).  A list that overflows is
 * AGOFRT_ERR_TOO_LARGE ("Too many neighbours in shell!", the reference's exception).  The literal minimum image; one
 * thread walks the partners of its atom in the reference's order. */
AGOFRT_API int agofrt_neighbours(agofrt_traj *traj, size_t frame, const uint64_t *nneigh, const double *cutoff2, int sort,
                                 uint64_t *list_out, double *rpos_out);
/* SphericalBase<lmax,double,T>::calc(frame, ...) without neighbour list (lib/src/sphericalbase.cpp:19-68): for every atom
 * i, type and radial bin the sum over the atoms j of that type in that bin of the real spherical harmonics Y_lm of the
 * direction xi - xj, l = 0 .. lmax <= 10 (SpecialFunctions::SphericalHarmonics, lib/include/specialfunctions.h:335-388:
 * the same recursions, every operation rounded on its own, summed in ascending j).  rminmax [ntypes*ntypes][2]: radial
 * range of the ordered type pair (type of i, type of j); result_out [natoms][ntypes][nbin][(lmax+1)^2] in the
 * reference's layout (l = lmax .. 0, negative m's first), counter_out [natoms][ntypes][nbin] or NULL. */
AGOFRT_API int agofrt_sh_density(agofrt_traj *traj, size_t frame, int lmax, unsigned nbin, const double *rminmax,
                                 double *result_out, int *counter_out);

/* ---- measurement --------------------------------------------------------------------------- */
/* Sustained FP64 FMA issue rate of one device (DFMA chains, CUDA events): the roofline
 * denominator of SURVEY.md section 8(d).  Runs for about `seconds`. */
AGOFRT_API int agofrt_fp64_peak(agofrt_ctx *ctx, int local_device, double seconds, double *dfma_per_second);

#ifdef __cplusplus
}
#endif
#endif /* AGOFRT_H */
