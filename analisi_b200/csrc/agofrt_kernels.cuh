// agofrt_kernels.cuh -- device side of libagofrt.so (sm_100a only).
//
// The pair kernel restates, on the GPU, the body of Gofrt::calc_single_th
// (reference lib/src/gofrt.cpp:95-122) on top of BaseTrajectory::d2_minImage_triclinic /
// minImage_triclinic (reference lib/include/basetrajectory.h:200-268), with the SAME double
// arithmetic: every subtraction, addition and product is rounded on its own (__dsub_rn / __dadd_rn
// / __dmul_rn, never contracted to FMA), the minimum image adds +-2*l_half (and the tilt factors)
// one image at a time in the order z, y, x, and d2 = ((0+dx*dx)+dy*dy)+dz*dz.
//
// The bin index (int)floorf((sqrt(d2)-rmin)/dr) is a monotone step function of d2, so it is
// looked up in a table of exact d2 thresholds computed on the host with the reference expression
// (agofrt_cabi.cu: build_thresholds).  The kernel only needs a float guess of the bin and two
// FP64 compares to land on exactly the reference's bin: no FP64 sqrt or divide on the device.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace agofrt {

#ifndef AGOFRT_IPT
#define AGOFRT_IPT 2
#endif
#ifndef AGOFRT_JU
#define AGOFRT_JU 8
#endif
#ifndef AGOFRT_MINBLOCKS
#define AGOFRT_MINBLOCKS 2
#endif
constexpr int kThreads = 256;               // threads per CTA
constexpr int kMinBlocks = AGOFRT_MINBLOCKS; // CTAs per SM the register budget is sized for
constexpr int kIPT = AGOFRT_IPT;            // i atoms held in registers per thread
constexpr int kTileI = kThreads * kIPT;     // i atoms per work unit
constexpr int kTileJ = 512;                 // j atoms per shared-memory stage
constexpr int kStages = 2;                  // bulk-copy stages
constexpr int kPadGroup = 8;                // every type group is padded to a multiple of this
constexpr int kJU = AGOFRT_JU;              // j atoms per inner step (one LDS.128 per coordinate per two)
constexpr int kWrapCap = 1 << 20;           // images the general minimum image may add per dimension
constexpr int kSmallMax = 512;              // systems of up to this many slots CAN go to pair_small_kernel
constexpr int kSmallDefault = 256;          // ... and up to this many do by default

struct Job {
    int fi;    // window-relative frame of the i atoms (its box is used)
    int fj;    // window-relative frame of the j atoms
    int tout;  // output lag row
};

struct PairParams {
    const double *pos;        // [frames][3][npad]  (x row, y row, z row per frame; ghosts are NaN)
    const double *box;        // [frames][6]: lx/2, ly/2, lz/2, xy, xz, yz
    const int *type_pad;      // [npad] dense type of every slot (ghosts: the type of their group)
    const int *type_start;    // [ntypes+1] first slot of every type group
    const Job *jobs;
    unsigned long long *ghist;  // [leff][2P][nbin]
    unsigned long long *edges;  // [1] (EDGES variant)
    unsigned int *counter;      // work-unit ticket
    unsigned int *error_flag;   // set when the general minimum image hits kWrapCap
    const double *thr;          // [nbin+1] thresholds folded with the rmin2/rmax2 test
    const double *thr_full;     // [nbin+1] plain thresholds (EDGES variant)
    double rmin2, rmax2;
    double ubox[12];            // UBOX variants: lx/2, ly/2, lz/2, xy, xz, yz, -lx, -ly, -lz, -xy, -xz, -yz of the one box of the window
    unsigned unit_begin, unit_end;    // pair_kernel: work units (job, i tile, j chunk) of this launch; pair_small_kernel: its JOBS
    int npad, ntypes, nbin;
    int n_itiles, n_jchunks, jchunk;  // jchunk is a multiple of kTileJ; pair_small_kernel: n_itiles = jobs a warp works on
                                      // at a time, n_jchunks = slots between their j frames in the warp's stage slice,
                                      // jchunk = how many jobs at the head of the job list are lag 0
    int small_nb, small_w16;          // pair_small_kernel: histograms (lags) a CTA holds at a time; cost of a lag-0 job in
                                      // sixteenths of another job's
    float inv_dr, c0;                 // bin guess = floor(sqrtf(d2) * inv_dr + c0)
    float c0h, lim, qmax;             // MODE_SAFE: c0 - 0.5, 0.5 - eps, nbin + 0.25 (clamp of the bin coordinate)
    int glo;                          // guard bins below bin 0 in every shared-memory histogram row
    int nhi;                          // bins from bin 0 up in every shared-memory row (>= nbin; only the first nbin are merged)
    int skew;                         // MODE_SAFE2: half of the warps start every tile with a dummy binning run (phase_skew)
    float inv_lo, inv_hi, bias0, smax;  // MODE_SAFE2: inv_dr * (1 -+ 2^-20), 1.5*2^23 + c0 (c0 an integer), clamp of sqrt(d2)
    unsigned hlo, hspan, hhi;         // candidate tests on the high word of d2 (hhi = hlo + hspan)
    // implicit jobs (imp != 0; jobs is then not read): job k = lag (k / imp_norig) * imp_every, origin frame
    // imp_f0 + (k % imp_norig) * imp_skip
    int imp, imp_f0, imp_norig, imp_skip, imp_every;
};

size_t pair_kernel_smem_bytes(int ntypes, int nbin, int nhi, int glo, bool edges);
size_t pair_small_kernel_smem_bytes(int ntypes, int nbin, int nhi, int glo, bool edges, int jobs_per_batch, int job_stride,
                                    int nhist);

// variant = TRI | FAST<<1 | MODE<<2 | UBOX<<5 | SMALL<<6, MODE: 0 thresholds, 1 thresholds + warp aggregation, 2 edges, 3 safe-zone,
// 4 safe-zone without the group filter (dense in-range workloads: nearly every group holds an in-range pair)
enum { kModeThr = 0, kModeAgg = 1, kModeEdges = 2, kModeSafe = 3, kModeSafeDense = 4, kModeSafe2 = 5 };  // dense, safe2: FAST only; safe2: tile kernel only
cudaError_t launch_pair_kernel(int variant, int grid, size_t smem, cudaStream_t stream, const PairParams &p);
cudaError_t prepare_pair_kernels(size_t max_smem_optin);

// pos_aos [nframes][natoms][3] -> pos_soa [nframes][3][npad] through perm[npad] (-1 = ghost -> NaN)
cudaError_t launch_gather_soa(const double *pos_aos, const int *perm, int natoms, int npad, int nframes,
                              double *pos_soa, cudaStream_t stream);
// inverse: frames back to the caller's order
cudaError_t launch_scatter_aos(const double *pos_soa, const int *perm, int natoms, int npad, int nframes,
                               double *pos_aos, cudaStream_t stream);
// raw LAMMPS dump records [nframes][natoms][8] (id type x y z vx vy vz, file order) -> AoS positions [nframes][natoms][3]
// in the caller's atom order through the id -> slot table; flags[4] counts unknown ids, flags[5] changed types
cudaError_t launch_parse_records(const double *raw, int natoms, int nframes, const int *id_table, int table_len,
                                 const int *slot_type, double *aos, unsigned int *flags, cudaStream_t stream);
// box rows [nframes][stride] -> [nframes][6] (lx/2, ly/2, lz/2, xy, xz, yz)
cudaError_t launch_pack_box(const double *box_internal, int stride, int nframes, double *box6,
                            cudaStream_t stream);
// per frame: min x,y,z, max x,y,z (NaN ignored), and a flag if any coordinate is +-inf
cudaError_t launch_frame_bounds(const double *pos_soa, const int *perm, int npad, int nframes, double *bounds6,
                                unsigned int *flags /* [0] inf seen, [2] NaN in a real atom */, cudaStream_t stream);
// BaseTrajectory::pbc_wrap on AoS positions
cudaError_t launch_pbc_wrap(double *pos_aos, int natoms, int nframes, const double *box_internal,
                            int stride, unsigned int *error_flag, cudaStream_t stream);
// all N^2 (dx,dy,dz,d2), caller's atom order
cudaError_t launch_d2_all(const double *pos_i, const double *pos_j, const double *box6, int triclinic,
                          const int *perm, int natoms, int npad, double *out, unsigned int *error_flag,
                          cudaStream_t stream);
// one pair, by device slots: out4 = dx,dy,dz,d2
cudaError_t launch_d2_pair(const double *pos_i, const double *pos_j, const double *box6, int triclinic, int slot_i,
                           int slot_j, int npad, double *out4, unsigned int *error_flag, cudaStream_t stream);
// neighbour-count histogram (IstogrammaAtomiRaggio)
struct NeighbourParams {
    const double *pos;         // [frames][3][npad]
    const double *box;         // [frames][6]
    const int *perm;           // [npad] slot -> atom, -1 = ghost
    const int *type_start;     // [ntypes+1]
    const int *frames;         // window-relative frames to visit
    unsigned long long *hist;  // [ntypes][hist_stride], hist_stride = natoms + 1
    unsigned int *counts;      // [listed frames][ntypes][npad] neighbours found so far (zeroed by the caller)
    unsigned int *error_flag;
    double r2;
    unsigned unit_begin, unit_end;   // units = (frame index, i tile, j chunk)
    int npad, ntypes, n_itiles, n_jchunks, jchunk;
    unsigned long long hist_stride;
};
cudaError_t launch_neighbour_kernel(bool triclinic, bool fast, int grid, cudaStream_t stream, const NeighbourParams &p);
// hist[type][counts[frame][type][slot]] += 1 for the real atoms of the (frame, i tile) pairs [v_begin, v_end)
cudaError_t launch_neighbour_finish(int grid, cudaStream_t stream, const NeighbourParams &p, unsigned int v_begin, unsigned int v_end);
int neighbour_tile_atoms();

// neighbour lists (Neighbours::update_neigh) and spherical-harmonic densities (SphericalBase::calc) of one frame
constexpr int kLsMaxTypes = 8;   // atom types these two kernels handle
constexpr int kShMaxL = 10;      // the reference instantiates l = 2 .. 10
struct NeighListParams {
    const double *pos;          // [frames][3][npad]
    const double *box;          // [frames][6]
    const int *atom_slot;       // [natoms] atom (caller's numbering) -> device slot
    const int *atom_type;       // [natoms] dense type
    unsigned long long *list;   // the reference's layout: per type t a block of natoms * (nneigh[t] + 1) words, count first
    double *rpos;               // per type t a block of natoms * nneigh[t] * 4 doubles: r, x, y, z
    unsigned int *flags;        // [0] a list overflowed ("Too many neighbours in shell!"), [1] minimum image did not converge
    unsigned long long nneigh[kLsMaxTypes], list_offset[kLsMaxTypes], rpos_offset[kLsMaxTypes];
    double cutoff2[kLsMaxTypes];
    int natoms, npad, ntypes, frame, triclinic, sort;
};
cudaError_t launch_neigh_list(const NeighListParams &p, cudaStream_t stream);
struct ShDensityParams {
    const double *pos, *box;
    const int *atom_slot, *atom_type;
    const double *rmin, *dr;    // [ntypes*ntypes] per ordered type pair (type of i, type of j)
    const double *coeff;        // [(lmax+1)^2] real spherical harmonics coefficients, [l][m]
    double *result;             // [natoms][ntypes][nbin][(lmax+1)^2], zeroed by the caller
    int *counter;               // [natoms][ntypes][nbin] or NULL
    unsigned int *flags;
    int natoms, npad, ntypes, frame, triclinic, lmax, nbin;
};
cudaError_t launch_sh_density(const ShDensityParams &p, cudaStream_t stream);

// mean square displacement (MSD)
struct MsdParams {
    const double *pos;        // [frames][3][npad]
    const double *cm;         // [frames][ntypes][3] (cm_self / cm_msd), else nullptr
    const int *tile_type;     // [ntiles]
    const int *tile_start;    // [ntiles] first slot
    const int *tile_count;    // [ntiles] real atoms in the tile (<= 256)
    const int *type_count;    // [ntypes] atoms of each type
    double *partial;          // [leff][ntiles]
    double *out;              // [leff][f_cm][ntypes]
    int npad, ntypes, ntiles, leff;
    int f0, ntimesteps, skip; // window-relative first origin, averaged steps, origin stride
    int cm_msd, cm_self;
};
cudaError_t launch_msd(const MsdParams &p, cudaStream_t stream);
int msd_tile_atoms();

// MediaVar::calculate on the device: block `block_index` (0-based) with x = counts * incr folded into mean / var
cudaError_t launch_blockavg_push(const unsigned long long *counts, double incr, unsigned block_index, double *mean,
                                 double *var, size_t len, int sm_count, cudaStream_t stream);

// ... for nblocks blocks [nblocks][len] in block order, first of them block `first_index`
cudaError_t launch_blockavg_push_blocks(const unsigned long long *counts, unsigned nblocks, double incr, unsigned first_index,
                                        double *mean, double *var, size_t len, int sm_count, cudaStream_t stream);

// MODE_SAFE validation: bad += number of probes whose unflagged float guess differs from expected[]
cudaError_t launch_validate_safe(const double *probes, const int *expected, int n, float inv_dr, float c0h, float lim,
                                 float qmax, int nbin, int glo, unsigned int *bad, cudaStream_t stream);
// MODE_SAFE2 validation: the same for the two-floor guess
cudaError_t launch_validate_safe2(const double *probes, const int *expected, int n, float inv_lo, float inv_hi, float bias0,
                                  float smax, int nbin, int glo, unsigned int *bad, cudaStream_t stream);
// DFMA chains; *count_per_launch receives the number of lane-level DFMAs one launch executes
cudaError_t launch_dfma_peak(double *sink, int blocks, int iters, cudaStream_t stream,
                             unsigned long long *count_per_launch);

}  // namespace agofrt
