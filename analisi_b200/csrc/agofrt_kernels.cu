// agofrt_kernels.cu -- hand-written sm_100a kernels of libagofrt.so.  See agofrt_kernels.cuh.
#include "agofrt_kernels.cuh"

#include <cstdlib>

#include <cmath>

namespace agofrt {

// ---------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + 1-D bulk copy (TMA engine, SASS UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// explicit shared-window accesses (32-bit addresses: no generic-pointer arithmetic in the hot loop)
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// ---------------------------------------------------------------------------------------------
// minimum image, literal form (reference lib/include/basetrajectory.h:224-268)
// ---------------------------------------------------------------------------------------------
struct BoxRegs {
    double lhx, lhy, lhz;   // half edges
    double xy, xz, yz;      // tilt factors
    double nxy, nxz, nyz;   // ... and their negatives (single-pass triclinic form)
};

template <bool TRI>
__device__ __forceinline__ bool min_image_general(double &dx, double &dy, double &dz, const BoxRegs &b) {
    int it = 0;
    const double Lz = __dmul_rn(b.lhz, 2.0), Ly = __dmul_rn(b.lhy, 2.0), Lx = __dmul_rn(b.lhx, 2.0);
    while (fabs(dz) > b.lhz) {
        if (dz < 0.0) {
            dz = __dadd_rn(dz, Lz);
            if (TRI) {
                dy = __dadd_rn(dy, b.yz);
                dx = __dadd_rn(dx, b.xz);
            }
        } else {
            dz = __dsub_rn(dz, Lz);
            if (TRI) {
                dy = __dsub_rn(dy, b.yz);
                dx = __dsub_rn(dx, b.xz);
            }
        }
        if (++it > kWrapCap) return false;
    }
    it = 0;
    while (fabs(dy) > b.lhy) {
        if (dy < 0.0) {
            dy = __dadd_rn(dy, Ly);
            if (TRI) dx = __dadd_rn(dx, b.xy);
        } else {
            dy = __dsub_rn(dy, Ly);
            if (TRI) dx = __dsub_rn(dx, b.xy);
        }
        if (++it > kWrapCap) return false;
    }
    it = 0;
    while (fabs(dx) > b.lhx) {
        if (dx < 0.0)
            dx = __dadd_rn(dx, Lx);
        else
            dx = __dsub_rn(dx, Lx);
        if (++it > kWrapCap) return false;
    }
    return true;
}

// Single-pass form, used only for (lag, origin) jobs whose coordinate bounds PROVE that one image
// per dimension is enough (agofrt_cabi.cu: job_is_single_pass).
//
// The reference's step for a component x with |x| > l_half is  x -= sign(x) * 2*l_half  (and the
// same signed subtraction of the tilt factors from the lower components).  A component whose lower
// neighbours do not depend on its sign (orthorhombic cells; the last, x, component of triclinic
// ones) only feeds its square, so |x| - 2*l_half serves as well: (-a)^2 == a^2 exactly.
// The wrap of a component whose sign matters (z and y of a triclinic cell: the tilt corrections of the
// lower components follow it) is  x = fma(m, -2*l_half, x)  with m = +-1.0 or 0.0 from the compare:
// m*c is exact, so the fused operation rounds once, exactly like the reference's "x -= 2*l_half"
// (m = 1) or leaves x untouched (m = 0), and the same m serves the tilt corrections.
// +-1.0 with the sign of s when p, else 0.0: one LOP3 (sign | exponent of 1.0) and one select, each
// reading a single register
__device__ __forceinline__ double mask_signed(bool p, double s) {
    const int one = (__double2hiint(s) & 0x80000000) | 0x3FF00000;
    return __hiloint2double(p ? one : 0, 0);
}

// A component that only feeds its square (all three of an orthorhombic cell; the last, x, of a triclinic
// one):  if |x| > l_half then x = |x| - 2*l_half  -- one DSETP and one PREDICATED DADD, nothing on the
// other pipes (no select, no mask register).  When the predicate is false x keeps its sign; (-a)^2 == a^2.
// Written as a branch over one instruction: ptxas if-converts it to "@P DADD" (a "@p add" in PTX comes
// back as DADD + 2 FSEL, two more issue slots).
__device__ __forceinline__ void wrap_abs(double &x, double lh, double nL) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f64 a;\n\t"
        "abs.f64 a, %0;\n\t"
        "setp.gt.f64 p, a, %1;\n\t"
        "@!p bra.uni WRAP_SKIP%=;\n\t"
        "add.rn.f64 %0, a, %2;\n\t"
        "WRAP_SKIP%=:\n\t}"
        : "+d"(x)
        : "d"(lh), "d"(nL));
}

// Triclinic components whose SIGN feeds the tilt corrections of the lower ones (z, then y).  The reference's step
//   if |v| > l_half:  (v, lower...) -= sign(v) * (2*l_half, tilt...)
// is odd in the whole displacement vector (IEEE addition and the compares are sign-symmetric), and only squares of
// the final components are used.  So instead of carrying sign(v) into the corrections (the r1 form built a +-1.0
// multiplier with two LOP3 and a SEL per component and applied it with DFMAs), the LOWER components are multiplied
// by sign(v) first -- one LOP3 on the high word each, exact -- and the step becomes the one-sided
//   if |v| > l_half:  v = |v| - 2*l_half;  lower -= tilt
// i.e. predicated DADDs with uniform constants.  The vector that comes out is the reference's times +-1 component
// by component: the same squares, the same d2, bit for bit.
__device__ __forceinline__ void flip_by_sign(double &x, double s) {   // x *= sign(s)  (xor of the sign bits)
    asm volatile(
        "{\n\t.reg .b32 xl, xh, sl, sh;\n\t"
        "mov.b64 {xl, xh}, %0;\n\t"
        "mov.b64 {sl, sh}, %1;\n\t"
        "lop3.b32 xh, xh, sh, 0x80000000, 0x78;\n\t"   // xh ^ (sh & 0x80000000)
        "mov.b64 %0, {xl, xh};\n\t}"
        : "+d"(x)
        : "d"(s));
}
// if |v| > lh: v = |v| + nL; a += na; b += nb   (nL = -2*lh, na / nb = minus the tilt factors)
__device__ __forceinline__ void wrap_abs3(double &v, double &a, double &b, double lh, double nL, double na, double nb) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f64 t;\n\t"
        "abs.f64 t, %0;\n\t"
        "setp.gt.f64 p, t, %3;\n\t"
        "@!p bra WRAP3_SKIP%=;\n\t"
        "add.rn.f64 %0, t, %4;\n\t"
        "add.rn.f64 %1, %1, %5;\n\t"
        "add.rn.f64 %2, %2, %6;\n\t"
        "WRAP3_SKIP%=:\n\t}"
        : "+d"(v), "+d"(a), "+d"(b)
        : "d"(lh), "d"(nL), "d"(na), "d"(nb));
}
__device__ __forceinline__ void wrap_abs2(double &v, double &a, double lh, double nL, double na) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f64 t;\n\t"
        "abs.f64 t, %0;\n\t"
        "setp.gt.f64 p, t, %2;\n\t"
        "@!p bra WRAP2_SKIP%=;\n\t"
        "add.rn.f64 %0, t, %3;\n\t"
        "add.rn.f64 %1, %1, %4;\n\t"
        "WRAP2_SKIP%=:\n\t}"
        : "+d"(v), "+d"(a)
        : "d"(lh), "d"(nL), "d"(na));
}

template <bool TRI>
__device__ __forceinline__ void min_image_single(double &dx, double &dy, double &dz, const BoxRegs &b,
                                                 double nLx, double nLy, double nLz) {  // nL = -(2*l_half)
    if (TRI) {
        // reference: z<0 ? (z+=L, y+=yz, x+=xz) : (z-=L, y-=yz, x-=xz)  ==  v -= sign(z) * (L, yz, xz)
#ifdef AGOFRT_OLD_TRIWRAP
        {
            const double m = mask_signed(fabs(dz) > b.lhz, dz);
            dz = __fma_rn(m, nLz, dz);
            dy = __fma_rn(m, -b.yz, dy);
            dx = __fma_rn(m, -b.xz, dx);
        }
        {
            const double m = mask_signed(fabs(dy) > b.lhy, dy);
            dy = __fma_rn(m, nLy, dy);
            dx = __fma_rn(m, -b.xy, dx);
        }
#else
        flip_by_sign(dy, dz);
        flip_by_sign(dx, dz);
        wrap_abs3(dz, dy, dx, b.lhz, nLz, b.nyz, b.nxz);
        flip_by_sign(dx, dy);
        wrap_abs2(dy, dx, b.lhy, nLy, b.nxy);
#endif
        wrap_abs(dx, b.lhx, nLx);   // last component: only its square is used
    } else {
        wrap_abs(dz, b.lhz, nLz);
        wrap_abs(dy, b.lhy, nLy);
        wrap_abs(dx, b.lhx, nLx);
    }
}

__device__ __forceinline__ double d2_of(double dx, double dy, double dz) {
    // reference lib/include/basetrajectory.h:215-217: d2=0; d2+=x*x for x,y,z in this order
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// ---------------------------------------------------------------------------------------------
// binning
// ---------------------------------------------------------------------------------------------
// MODE_THR   float guess of the bin + bracket test against the exact thresholds (always exact)
// MODE_AGG   the same, the shared atomic merged across the warp with __match_any_sync
// MODE_EDGES plain thresholds + explicit range test + count of pairs within 1 ulp of a bin edge
// MODE_SAFE  float guess accepted without looking at the thresholds when it is farther than `eps`
//            bins from a bin edge (eps bounds the float error, validated on the device when the plan
//            is made); the few pairs closer than eps to an edge go through the exact bracket search
//   MODE_SAFE2 the dense path, second form (see group_dense2): the bin index AND its row come out of ONE
//            round-down FFMA as the word index of the shared histogram, a second FFMA with a slightly larger
//            slope says whether the guess sits within eps of an edge, and one unsigned compare of that word
//            against the end of the row predicates the shared atomic (far pairs, NaN ghosts and +inf never touch
//            memory).  Needs an integer c0 (rmin a multiple of dr, e.g. 0).
enum { MODE_THR = 0, MODE_AGG = 1, MODE_EDGES = 2, MODE_SAFE = 3, MODE_SAFE_DENSE = 4, MODE_SAFE2 = 5 };

// float guess of the bin from the bits of d2 (no FP64 conversion instruction): rebias the
// exponent, keep 23 mantissa bits, MUFU sqrt, one FFMA.  Returns guess+1 clamped to [0, nbin+2]:
// the index into the padded pair table below.
__device__ __forceinline__ float d2_as_float(double d2) {
    const int hi = __double2hiint(d2);
    const uint32_t lo = static_cast<uint32_t>(__double2loint(d2));
    int hh = max(hi, 0x38000000) - 0x38000000;   // exponent 1023-127 = 896; tiny and zero -> 0
    hh = min(hh, 0x0FEFFFFF);
    return __uint_as_float(__funnelshift_l(lo, static_cast<uint32_t>(hh), 3));
}
__device__ __forceinline__ float sqrt_approx(float f) {
    float s;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(f));
    return s;
}
__device__ __forceinline__ unsigned int bin_guess1(double d2, float inv_dr, float c0, unsigned int nbin2) {
    const int g = __float2int_rd(fmaf(sqrt_approx(d2_as_float(d2)), inv_dr, c0));
    return min(static_cast<unsigned int>(g + 1), nbin2);
}

template <bool AGG>
__device__ __forceinline__ void hist_add(unsigned int *hist, unsigned int idx) {
    if (AGG) {
        const unsigned int lane = threadIdx.x & 31u;
        const unsigned int peers = __match_any_sync(__activemask(), idx);
        if (static_cast<unsigned int>(__ffs(peers) - 1) == lane) atomicAdd(hist + idx, __popc(peers));
    } else {
        atomicAdd(hist + idx, 1u);
    }
}

// The shared-memory threshold table is stored as pairs: thr2[g+1] = {T[g], T[g+1]} for
// g = -1 .. nbin+1, with {+inf,+inf} for the slots outside [0,nbin) -- one LDS.128 brackets a guess,
// and a guess outside the histogram can never pass the bracket test.
// Rare path: exact bracket search (the guess missed, or sits within eps of an edge, or d2 is outside).
template <bool AGG>
__device__ __noinline__ void bin_pair_slow(double d2, const double2 *__restrict__ thr2, int nbin, float inv_dr,
                                           float c0, unsigned int *hist, unsigned int row) {
    // thr2[1].x = T[0], thr2[nbin].y = T[nbin]
    if (!(d2 >= thr2[1].x) || !(d2 < thr2[nbin].y)) return;  // not counted by the reference
    int g = static_cast<int>(bin_guess1(d2, inv_dr, c0, static_cast<unsigned int>(nbin) + 2u)) - 1;
    g = min(max(g, 0), nbin - 1);
    while (d2 < thr2[g + 1].x) --g;
    while (d2 >= thr2[g + 1].y) ++g;
    hist_add<AGG>(hist, row + static_cast<unsigned int>(g));
}

// EDGES path (tests / reporting): plain thresholds, explicit range test, and the count of pairs
// whose d2 is a threshold or the double just below one.
__device__ __forceinline__ void bin_pair_edges(double d2, const double *__restrict__ thrf, int nbin, float inv_dr,
                                               float c0, double rmin2, double rmax2, unsigned int *hist,
                                               unsigned int row, unsigned long long &edges) {
    if (d2 > rmax2 || d2 < rmin2 || d2 != d2) return;  // reference lib/src/gofrt.cpp:104
    int g = static_cast<int>(bin_guess1(d2, inv_dr, c0, static_cast<unsigned int>(nbin) + 2u)) - 1;
    g = min(max(g, 0), nbin - 1);
    // idx = (number of k in 0..nbin with thrf[k] <= d2) - 1, in [-1, nbin]
    while (g >= 0 && d2 < thrf[g]) --g;
    while (g < nbin && d2 >= thrf[g + 1]) ++g;
    if (d2 > 0.0) {
        const double up = __longlong_as_double(__double_as_longlong(d2) + 1);  // next double above (d2 > 0)
        bool e = false;
        if (g >= 0 && d2 == thrf[g]) e = true;
        if (g < nbin && up == thrf[g + 1]) e = true;
        if (e) ++edges;
    }
    if (g >= 0 && g < nbin) atomicAdd(hist + row + static_cast<unsigned int>(g), 1u);
}

// MODE_THR fast path for one pair: bracket test and predicated shared atomic; returns nonzero if
// the bracket test failed (the caller decides whether the pair needs the exact search).
__device__ __forceinline__ unsigned int bin_pair_thr(double v, uint32_t thr2_addr, uint32_t row_addr_m4,
                                                     float inv_dr, float c0, unsigned int nbin2) {
    const unsigned int gi = bin_guess1(v, inv_dr, c0, nbin2);
    const double2 t = lds_f64x2(thr2_addr + gi * 16u);
    unsigned int miss;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ge.f64 p, %1, %2;\n\t"
        "setp.lt.and.f64 p, %1, %3, p;\n\t"
        "@p red.shared.add.u32 [%4], 1;\n\t"
        "selp.u32 %0, 0, 1, p;\n\t}"
        : "=r"(miss)
        : "d"(v), "d"(t.x), "d"(t.y), "r"(row_addr_m4 + gi * 4u)
        : "memory");
    return miss;
}

// MODE_SAFE fast path.  qh = sqrtf(d2)*inv_dr + (c0 - 0.5) is the bin coordinate minus one half; adding
// 1.5*2^23 rounds it to the nearest integer n (the guessed bin) in the low mantissa bits, and
// |qh - n| < 0.5 - eps says the guess is more than eps bins away from both edges of bin n.
//
// The fast path is UNCONDITIONAL: every pair increments hist[row + n], whatever n is.  Rows of the
// shared histogram carry guard bins -- `glo` below bin 0 (qh >= c0 - 0.5, so n >= -glo) and one above
// (min.f32 clamps qh to nbin + 0.25, which also catches NaN ghosts and far pairs) -- so out-of-range
// pairs land in words that are never merged.  No range compare, no address select, no predicated atomic
// (ptxas turns those into branches).  What the fast path cannot decide is left to a group-level test:
// the largest |qh - n| of the group (FMNMX3, half an instruction per pair) against lim.  If it fails --
// about 3 pairs in 10^4 -- the group's pairs are re-examined and each pair within eps of an edge is
// CORRECTED: its fast-path increment is taken back and the exact bin (bracket search in the threshold
// table) is incremented instead.  Sums are integers modulo 2^32 and a CTA merges its rows only at
// barriers, so the order of +1 / -1 does not matter.
struct SafeGuess {
    float dl;   // qh - n (signed)
    int n;      // guessed bin, in [-glo, nbin]
};
__device__ __forceinline__ SafeGuess safe_guess(double v, float inv_dr, float c0h, float qmax) {
    float f;
    asm("cvt.rz.f32.f64 %0, %1;" : "=f"(f) : "d"(v));   // +inf for huge, NaN stays NaN; sqrt.ftz flushes denormals
    const float s = sqrt_approx(f);
    float q, r;
    asm("{\n\t.reg .f32 t;\n\t"
        "fma.rn.f32 t, %2, %3, %4;\n\t"
        "min.f32 %0, t, %5;\n\t"                 // NaN -> qmax
        "add.rn.f32 %1, %0, 0f4B400000;\n\t}"    // 1.5 * 2^23: r = 1.5*2^23 + n, n = rint(q)
        : "=f"(q), "=f"(r)
        : "f"(s), "f"(inv_dr), "f"(c0h), "f"(qmax));
    SafeGuess g;
    g.dl = __fsub_rn(q, __fadd_rn(r, -12582912.0f));
    g.n = __float_as_int(r) - 0x4B400000;
    return g;
}

// one pair: returns qh - n; row_addr_adj = shared address of hist[row] - 4 * 0x4B400000
__device__ __forceinline__ float bin_pair_safe(double v, uint32_t row_addr_adj, float inv_dr, float c0h, float qmax) {
    float f;
    asm("cvt.rz.f32.f64 %0, %1;" : "=f"(f) : "d"(v));
    const float s = sqrt_approx(f);
    float dl;
    asm volatile(
        "{\n\t.reg .f32 q, r, n;\n\t.reg .b32 ri, ad;\n\t"
        "fma.rn.f32 q, %1, %2, %3;\n\t"
        "min.f32 q, q, %4;\n\t"
        "add.rn.f32 r, q, 0f4B400000;\n\t"
        "add.rn.f32 n, r, 0fCB400000;\n\t"
        "sub.rn.f32 %0, q, n;\n\t"
        "mov.b32 ri, r;\n\t"
        "mad.lo.u32 ad, ri, 4, %5;\n\t"
        "red.shared.add.u32 [ad], 1;\n\t}"
        : "=f"(dl)
        : "f"(s), "f"(inv_dr), "f"(c0h), "f"(qmax), "r"(row_addr_adj)
        : "memory");
    return dl;
}

__device__ __forceinline__ float max3abs(float a, float b, float c) {
    float d;
    asm("{\n\t.reg .f32 x, y;\n\tabs.f32 x, %2;\n\tabs.f32 y, %3;\n\tmax.f32 %0, %1, x, y;\n\t}"
        : "=f"(d)
        : "f"(a), "f"(b), "f"(c));
    return d;
}

// Rare path of MODE_SAFE: the pair's guess is within eps of a bin edge (or is NaN-like): take the
// fast-path increment back and count the pair where the exact table says.
__device__ __noinline__ void bin_pair_fix(double d2, const double2 *__restrict__ thr2, int nbin, float inv_dr, float c0,
                                          float c0h, float qmax, float lim, unsigned int *hist_row) {
    const SafeGuess sg = safe_guess(d2, inv_dr, c0h, qmax);
    if (fabsf(sg.dl) < lim) return;   // this pair of the group was safe
    int g = -1;
    if ((d2 >= thr2[1].x) && (d2 < thr2[nbin].y)) {
        g = static_cast<int>(bin_guess1(d2, inv_dr, c0, static_cast<unsigned int>(nbin) + 2u)) - 1;
        g = min(max(g, 0), nbin - 1);
        while (d2 < thr2[g + 1].x) --g;
        while (d2 >= thr2[g + 1].y) ++g;
    }
    if (g == sg.n) return;
    atomicAdd(hist_row + sg.n, 0xffffffffu);   // -1 modulo 2^32 (sg.n may be a guard bin)
    if (g >= 0) atomicAdd(hist_row + g, 1u);
}

// ---------------------------------------------------------------------------------------------
// the pair kernel
// ---------------------------------------------------------------------------------------------
struct SmemLayout {
    size_t thr2, thr_full, stage, hist, dump, rowtab, tstart, bars, sched, total;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// every histogram row: glo guard bins, nhi bins (the first nbin of them are merged; nhi = nbin today), one guard bin
__host__ __device__ inline int row_stride(int nhi, int glo) { return glo + nhi + 1; }

// stage_slots: doubles per coordinate row of the stage area -- kStages tiles of the tile kernel, or the warps' job
// slices of pair_small_kernel (small_stage_slots); nhist: histograms kept at a time (pair_small_kernel: one per lag)
__host__ __device__ inline SmemLayout smem_layout(int ntypes, int nbin, int nhi, int glo, bool edges,
                                                   size_t stage_slots = static_cast<size_t>(kStages) * kTileJ, int nhist = 1) {
    SmemLayout L;
    size_t o = 0;
    L.stage = o;
    o += stage_slots * 3 * sizeof(double);
    L.thr2 = o;
    o += static_cast<size_t>(nbin + 3) * 2 * sizeof(double);
    L.thr_full = o;
    if (edges) o += static_cast<size_t>(nbin + 1) * sizeof(double);
    L.bars = align_up(o, 8);
    o = L.bars + kStages * sizeof(uint64_t);
    L.hist = o;
    o += static_cast<size_t>(nhist) * ntypes * (ntypes + 1) * row_stride(nhi, glo) * sizeof(unsigned int);
    L.dump = o;
    o += 32 * sizeof(unsigned int);
    L.rowtab = o;
    o += static_cast<size_t>(ntypes) * ntypes * sizeof(unsigned int);
    L.tstart = o;
    o += static_cast<size_t>(ntypes + 1) * sizeof(int);
    L.sched = o;
    o += 4 * sizeof(unsigned int);
    L.total = align_up(o, 16);
    return L;
}

size_t pair_kernel_smem_bytes(int ntypes, int nbin, int nhi, int glo, bool edges) {
    return smem_layout(ntypes, nbin, nhi, glo, edges).total;
}

// pair_small_kernel: every warp of a CTA stages `jobs_per_batch` j frames of `job_stride` slots each
__host__ __device__ inline size_t small_stage_slots(int jobs_per_batch, int job_stride) {
    return static_cast<size_t>(kThreads / 32) * jobs_per_batch * job_stride;
}
size_t pair_small_kernel_smem_bytes(int ntypes, int nbin, int nhi, int glo, bool edges, int jobs_per_batch, int job_stride,
                                    int nhist) {
    return smem_layout(ntypes, nbin, nhi, glo, edges, small_stage_slots(jobs_per_batch, job_stride), nhist).total;
}

constexpr int kWarpsPerCta = kThreads / 32;

// the (lag, origin) job with this index: read from the list, or derived (implicit jobs)
__device__ __forceinline__ Job job_at(const PairParams &p, unsigned int k) {
    if (!p.imp) return p.jobs[k];
    const unsigned int lag = k / static_cast<unsigned int>(p.imp_norig), o = k - lag * static_cast<unsigned int>(p.imp_norig);
    Job j;
    j.fi = p.imp_f0 + static_cast<int>(o) * p.imp_skip;
    j.tout = static_cast<int>(lag) * p.imp_every;
    j.fj = j.fi + j.tout;
    return j;
}

struct PairConst {
    BoxRegs box;
    double nLx, nLy, nLz;   // -(2*l_half), exact
    uint32_t thr2_addr;     // shared-window address of thr2[0]
    uint32_t hist_addr;     // shared-window address of hist[0]
};

// One group = kIPT i atoms (registers) x kJU j atoms (shared memory): all d2 first (straight-line,
// independent FP64 chains), then one cheap test for "some pair of the group may be in range", then
// the binning of the group's pairs without per-pair branches.
//   DIAG: the j atoms may include one of this thread's own i atoms (i == j goes to the "self" rows).
//   ZSELF (pair_small_kernel, lag 0): i == j is skipped; the caller counts those pairs.
// All kIPT x kJU squared distances of one group (straight-line, independent FP64 chains).
//   jrow_bytes: bytes between the x, y and z rows of the staged j coordinates (a literal in pair_kernel)
template <bool TRI, bool FAST>
__device__ __forceinline__ void group_distances(const PairConst &c, const double (&xi)[kIPT], const double (&yi)[kIPT],
                                                const double (&zi)[kIPT], uint32_t sx_addr, uint32_t jrow_bytes, int jrel,
                                                double (&d2)[kIPT][kJU], bool &wrap_ok) {
    double xj[kJU], yj[kJU], zj[kJU];
#pragma unroll
    for (int q = 0; q < kJU; q += 2) {
        const uint32_t a = sx_addr + static_cast<uint32_t>(jrel + q) * 8u;
        const double2 vx = lds_f64x2(a);
        const double2 vy = lds_f64x2(a + jrow_bytes);
        const double2 vz = lds_f64x2(a + 2u * jrow_bytes);
        xj[q] = vx.x;
        xj[q + 1] = vx.y;
        yj[q] = vy.x;
        yj[q + 1] = vy.y;
        zj[q] = vz.x;
        zj[q + 1] = vz.y;
    }
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
#pragma unroll
        for (int q = 0; q < kJU; ++q) {
            // reference lib/include/basetrajectory.h:207-209: x = xi - xj
            double dx = __dsub_rn(xi[k], xj[q]);
            double dy = __dsub_rn(yi[k], yj[q]);
            double dz = __dsub_rn(zi[k], zj[q]);
            if (FAST) {
                min_image_single<TRI>(dx, dy, dz, c.box, c.nLx, c.nLy, c.nLz);
            } else {
                wrap_ok &= min_image_general<TRI>(dx, dy, dz, c.box);
            }
            d2[k][q] = d2_of(dx, dy, dz);
        }
    }
}

// Safe-zone binning of one group: straight-line, one group-level test for the rare correction.
template <bool DIAG>
__device__ __forceinline__ void group_bin_safe(const PairParams &p, const PairConst &c, const double (&d2)[kIPT][kJU],
                                               const int (&ii)[kIPT], const unsigned int (&row)[kIPT], int j,
                                               const double2 *s_thr2, unsigned int *s_hist, unsigned int self_off) {
    float dl[kIPT * kJU];
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
#pragma unroll
        for (int q = 0; q < kJU; ++q) {
            const unsigned int r = row[k] + ((DIAG && ii[k] == j + q) ? self_off : 0u);
            // byte address of hist[r + n] = hist_addr + 4*(r + bits(rounded) - 0x4B400000)
            const uint32_t adj = c.hist_addr + 4u * r - 4u * 0x4B400000u;
            dl[k * kJU + q] = bin_pair_safe(d2[k][q], adj, p.inv_dr, p.c0h, p.qmax);
        }
    }
    float m = 0.0f;
#pragma unroll
    for (int t = 0; t + 1 < kIPT * kJU; t += 2) m = max3abs(m, dl[t], dl[t + 1]);
    if (kIPT * kJU % 2) m = fmaxf(m, fabsf(dl[kIPT * kJU - 1]));
    if (!(m < p.lim)) {
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
#pragma unroll
            for (int q = 0; q < kJU; ++q) {   // fully unrolled: the distances stay in registers
                if (!(fabsf(dl[k * kJU + q]) < p.lim)) {
                    const unsigned int r = row[k] + ((DIAG && ii[k] == j + q) ? self_off : 0u);
                    bin_pair_fix(d2[k][q], s_thr2, p.nbin, p.inv_dr, p.c0, p.c0h, p.qmax, p.lim, s_hist + r);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// MODE_SAFE2: distance and binning of one group, interleaved
// ---------------------------------------------------------------------------------------------
// Per pair:  f = (float)d2 (toward zero), s = sqrt.approx(f), and then
//     ra = fma.rm(s, inv_lo, bias)      rb = fma.rm(s, inv_hi, bias)
// with bias = 1.5*2^23 + c0 + (word index of bin 0 of the pair's histogram row), an integer: in [2^23, 2^24) floats
// are the integers, so the round-down FFMA gives  1.5*2^23 + row + floor(s*inv + c0)  EXACTLY and the low mantissa
// bits of ra are the word of the shared histogram to increment -- no clamp, no float->int conversion, no separate
// row add.  inv_lo / inv_hi are inv_dr * (1 -+ 2^-20): the two floors differ exactly when an integer lies between the
// two products, i.e. when the guess is within a relative 2^-20 of a bin edge (the float error of s and of the
// reference's own rounding is below 2^-21.4; validated on the device for every plan, as MODE_SAFE is).  The group ORs
// ra ^ rb; a nonzero OR (3 groups in 10^3) sends the group to group_fix2, which recomputes it and corrects the pairs
// that are near an edge.  The instruction streams of "distance of pair t" (14 / 17 FP64 instructions) and "binning of
// pair t-2" (7 instructions on the other pipes) are emitted alternately, so that every warp always has work for both
// the FP64 pipe and the rest of the issue slots (in r1 a warp ran 224 FP64 instructions, then 190 others: four warps
// per scheduler were too few to keep the FP64 pipe fed).
struct IAtoms {
    double x[kIPT], y[kIPT], z[kIPT];
    float bias[kIPT];
};

__device__ __forceinline__ float cvt_rz(double v) {
    float f;
    asm volatile("cvt.rz.f32.f64 %0, %1;" : "=f"(f) : "d"(v));
    return f;
}
__device__ __forceinline__ float sqrt_approx_v(float f) {
    float s;
    asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(f));
    return s;
}
__device__ __forceinline__ void safe2_floors(float s, float inv_lo, float inv_hi, float bias, float &ra, float &rb) {
    asm("fma.rm.f32 %0, %1, %2, %3;" : "=f"(ra) : "f"(s), "f"(inv_lo), "f"(bias));
    asm("fma.rm.f32 %0, %1, %2, %3;" : "=f"(rb) : "f"(s), "f"(inv_hi), "f"(bias));
}
// kaddr = shared address of hist[0] - 4 * 0x4B400000 (mod 2^32).  s comes in clamped to smax (see safe2_root), so the
// word is at most row + nbin: the guard bin that ends every row.  (The OR into acc stays outside the asm block: with
// acc as a read-write operand of a lop3 inside it, ptxas 12.9 crashes on the unrolled group.  And the atomic is
// unconditional: ptxas turns a predicated red.shared into a branch around it.)
__device__ __forceinline__ void safe2_bin(float s, float inv_lo, float inv_hi, float bias, uint32_t kaddr, unsigned int &acc) {
    unsigned int x;
    asm volatile(
        "{\n\t.reg .f32 ra, rb;\n\t.reg .b32 ia, ib, ad;\n\t"
        "fma.rm.f32 ra, %1, %2, %4;\n\t"
        "fma.rm.f32 rb, %1, %3, %4;\n\t"
        "mov.b32 ia, ra;\n\t"
        "mov.b32 ib, rb;\n\t"
        "xor.b32 %0, ia, ib;\n\t"
        "shl.b32 ad, ia, 2;\n\t"
        "add.u32 ad, ad, %5;\n\t"
        "red.shared.add.u32 [ad], 1;\n\t}"
        : "=r"(x)
        : "f"(s), "f"(inv_lo), "f"(inv_hi), "f"(bias), "r"(kaddr)
        : "memory");
    acc |= x;
}
// s = min(sqrt.approx((float)d2), smax): smax = (nbin + 0.5 - c0) * dr maps to the middle of the guard bin past the
// last real one for both slopes; min.f32 returns the other operand for NaN (ghost slots), so far pairs, +inf and NaN
// all count there and are never flagged as near an edge
__device__ __forceinline__ float safe2_root(float f, float smax) {
    float s;
    asm volatile("{\n\t.reg .f32 t;\n\tsqrt.approx.ftz.f32 t, %1;\n\tmin.f32 %0, t, %2;\n\t}" : "=f"(s) : "f"(f), "f"(smax));
    return s;
}

template <bool TRI>
__device__ __forceinline__ double pair_d2_single(const PairConst &c, double xi, double yi, double zi, double xj, double yj,
                                                 double zj) {
    double dx = __dsub_rn(xi, xj);
    double dy = __dsub_rn(yi, yj);
    double dz = __dsub_rn(zi, zj);
    min_image_single<TRI>(dx, dy, dz, c.box, c.nLx, c.nLy, c.nLz);
    return d2_of(dx, dy, dz);
}

template <bool TRI>
__device__ __forceinline__ unsigned int group_dense2(const PairParams &p, const PairConst &c, const double (&xi)[kIPT],
                                                     const double (&yi)[kIPT], const double (&zi)[kIPT],
                                                     const float (&bias)[kIPT], uint32_t kaddr, uint32_t sx_addr,
                                                     uint32_t jrow_bytes, int jrel) {
    double xj[kJU], yj[kJU], zj[kJU];
#pragma unroll
    for (int q = 0; q < kJU; q += 2) {
        const uint32_t a = sx_addr + static_cast<uint32_t>(jrel + q) * 8u;
        const double2 vx = lds_f64x2(a);
        const double2 vy = lds_f64x2(a + jrow_bytes);
        const double2 vz = lds_f64x2(a + 2u * jrow_bytes);
        xj[q] = vx.x;
        xj[q + 1] = vx.y;
        yj[q] = vy.x;
        yj[q + 1] = vy.y;
        zj[q] = vz.x;
        zj[q + 1] = vz.y;
    }
    constexpr int NP = kIPT * kJU;
    constexpr int LAG_S = 1, LAG_B = 2;   // pairs between the distance, the root and the binning of a pair
    float fv[NP];
    unsigned int acc = 0;
#pragma unroll
    for (int t = 0; t < NP + LAG_B; ++t) {
        if (t < NP) {
            const int q = t / kIPT, k = t % kIPT;
            fv[t] = cvt_rz(pair_d2_single<TRI>(c, xi[k], yi[k], zi[k], xj[q], yj[q], zj[q]));
        }
        if (t >= LAG_S && t - LAG_S < NP) fv[t - LAG_S] = safe2_root(fv[t - LAG_S], p.smax);
        if (t >= LAG_B && t - LAG_B < NP) safe2_bin(fv[t - LAG_B], p.inv_lo, p.inv_hi, bias[(t - LAG_B) % kIPT], kaddr, acc);
    }
    return acc;
}

// Phase skew (MODE_SAFE2).  A group is a run of FP64 instructions (the distances) followed by a run that keeps the
// XU pipe busy (conversion and root of every pair: 8 cycles each) -- and the warps of a scheduler, released together by
// the tile barrier and served round-robin, walk through the two runs in step: FP64 pipe and XU pipe take turns instead
// of working side by side (r1/r2 profiles: 40 cycles per pair where the FP64 pipe alone needs 28).  Half of the warps
// of every scheduler therefore start each tile with one binning run on dummy values (they count in the guard bin that
// ends a row): from then on these warps bin while the others compute distances, and the offset keeps itself up, because
// two warps in the same run compete for one pipe and two warps in different runs do not.
__device__ __forceinline__ void phase_skew(const PairParams &p, float bias, uint32_t kaddr, int seed) {
    unsigned int acc = 0;
#pragma unroll
    for (int t = 0; t < kIPT * kJU; ++t) {
        const float f = cvt_rz(__hiloint2double(0x7fe00000 - (seed << 4) - t, 0));   // huge, finite, different every time
        safe2_bin(safe2_root(f, p.smax), p.inv_lo, p.inv_hi, bias, kaddr, acc);
    }
    if (acc) atomicExch(p.error_flag, 2u);   // cannot happen: the clamped root is the same for both slopes
}

// Rare path of MODE_SAFE2: some pair of the group is within eps of a bin edge.  Recompute the group (same instructions,
// same bits) and for every such pair take the fast-path increment back (if there was one) and count the pair where
// the exact threshold table says (or nowhere).  Counters are integers modulo 2^32 and a CTA merges its rows only at barriers.
template <bool TRI>
__device__ __noinline__ void group_fix2(PairConst c, IAtoms ia, float inv_lo, float inv_hi, float bias0, float smax, uint32_t sx_addr,
                                        uint32_t jrow_bytes, int jrel, const double2 *__restrict__ thr2, int nbin,
                                        float inv_dr, float c0, unsigned int *s_hist) {
    for (int q = 0; q < kJU; ++q) {
        const uint32_t a = sx_addr + static_cast<uint32_t>(jrel + q) * 8u;
        double xj, yj, zj;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(xj) : "r"(a));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(yj) : "r"(a + jrow_bytes));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(zj) : "r"(a + 2u * jrow_bytes));
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            const double d2 = pair_d2_single<TRI>(c, ia.x[k], ia.y[k], ia.z[k], xj, yj, zj);
            const float s = safe2_root(cvt_rz(d2), smax);
            float ra, rb;
            safe2_floors(s, inv_lo, inv_hi, ia.bias[k], ra, rb);
            if (__float_as_int(ra) == __float_as_int(rb)) continue;
            const int word = __float_as_int(ra) - 0x4B400000;                        // what the fast path incremented
            const int row = static_cast<int>(__fsub_rn(ia.bias[k], bias0));         // exact: both are integers
            int g = -1;
            if ((d2 >= thr2[1].x) && (d2 < thr2[nbin].y)) {
                g = static_cast<int>(bin_guess1(d2, inv_dr, c0, static_cast<unsigned int>(nbin) + 2u)) - 1;
                g = min(max(g, 0), nbin - 1);
                while (d2 < thr2[g + 1].x) --g;
                while (d2 >= thr2[g + 1].y) ++g;
            }
            if (g >= 0 && row + g == word) continue;
            atomicAdd(s_hist + word, 0xffffffffu);   // (may be the guard bin that ends the row: never merged)
            if (g >= 0) atomicAdd(s_hist + row + g, 1u);
        }
    }
}

template <bool TRI, bool FAST, int MODE, bool DIAG, bool ZSELF = false>
__device__ __forceinline__ void process_group(const PairParams &p, const PairConst &c, const double (&xi)[kIPT],
                                              const double (&yi)[kIPT], const double (&zi)[kIPT],
                                              const int (&ii)[kIPT], const unsigned int (&row)[kIPT],
                                              uint32_t sx_addr, uint32_t jrow_bytes, int jrel, int j, const double2 *s_thr2,
                                              const double *s_thrf, unsigned int *s_hist, unsigned int self_off,
                                              unsigned long long &edges, bool &wrap_ok) {
    double d2[kIPT][kJU];
    group_distances<TRI, FAST>(c, xi, yi, zi, sx_addr, jrow_bytes, jrel, d2, wrap_ok);
    if (ZSELF) {
        // the pair of an atom with itself is left out here (NaN: out of range for every mode, never "near an edge")
        // and counted by the caller
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
#pragma unroll
            for (int q = 0; q < kJU; ++q)
                if (ii[k] == j + q) d2[k][q] = __longlong_as_double(0x7ff8000000000000ll);
        }
    }
    if (MODE == MODE_EDGES) {
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
#pragma unroll
            for (int q = 0; q < kJU; ++q) {
                const unsigned int r = row[k] + ((DIAG && ii[k] == j + q) ? self_off : 0u);
                bin_pair_edges(d2[k][q], s_thrf, p.nbin, p.inv_dr, p.c0, p.rmin2, p.rmax2, s_hist, r, edges);
            }
        }
        return;
    }
    // group filter on the high words (d2 >= 0 or NaN: the unsigned high word is monotone in d2);
    // only the upper end is tested here, the exact range test is in the binning
    unsigned int m = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
#pragma unroll
        for (int q = 0; q < kJU; ++q) m = min(m, static_cast<unsigned int>(__double2hiint(d2[k][q])));
    }
    if (MODE != MODE_SAFE_DENSE && m > p.hhi) return;

    if (MODE == MODE_SAFE || MODE == MODE_SAFE_DENSE) {
        group_bin_safe<DIAG>(p, c, d2, ii, row, j, s_thr2, s_hist, self_off);
        return;
    }
    unsigned int mask = 0;
    if (MODE == MODE_THR) {
        const unsigned int nbin2 = static_cast<unsigned int>(p.nbin) + 2u;
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
#pragma unroll
            for (int q = 0; q < kJU; ++q) {
                const unsigned int r = row[k] + ((DIAG && ii[k] == j + q) ? self_off : 0u);
                const unsigned int miss =
                    bin_pair_thr(d2[k][q], c.thr2_addr, c.hist_addr + 4u * r - 4u, p.inv_dr, p.c0, nbin2);
                const unsigned int h = static_cast<unsigned int>(__double2hiint(d2[k][q]));
                if (miss && (h - p.hlo <= p.hspan)) mask |= 1u << (k * kJU + q);
            }
        }
    } else {  // MODE_AGG: every candidate pair through the generic path with warp aggregation
        const unsigned int nbin2 = static_cast<unsigned int>(p.nbin) + 2u;
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
#pragma unroll
            for (int q = 0; q < kJU; ++q) {
                const double v = d2[k][q];
                const unsigned int h = static_cast<unsigned int>(__double2hiint(v));
                if (h - p.hlo <= p.hspan) {
                    const unsigned int r = row[k] + ((DIAG && ii[k] == j + q) ? self_off : 0u);
                    const unsigned int gi = bin_guess1(v, p.inv_dr, p.c0, nbin2);
                    const double2 t = s_thr2[gi];
                    if ((v >= t.x) && (v < t.y))
                        hist_add<true>(s_hist, r + gi - 1u);
                    else
                        bin_pair_slow<true>(v, s_thr2, p.nbin, p.inv_dr, p.c0, s_hist, r);
                }
            }
        }
    }
    if (mask) {
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
#pragma unroll
            for (int q = 0; q < kJU; ++q) {
                if (mask & (1u << (k * kJU + q))) {
                    const unsigned int r = row[k] + ((DIAG && ii[k] == j + q) ? self_off : 0u);
                    bin_pair_slow<false>(d2[k][q], s_thr2, p.nbin, p.inv_dr, p.c0, s_hist, r);
                }
            }
        }
    }
}

// Shared-memory tables every pair kernel starts from: zeroed histogram rows, the threshold pairs, the
// (type_i, type_j) -> row table of Gofrt::get_itype, and the first slot of every type group.
template <bool EDGES>
__device__ __forceinline__ void cta_tables(const PairParams &p, int tid, int P, int rstride, unsigned int *s_hist,
                                           double2 *s_thr2, double *s_thrf, unsigned int *s_rowtab, int *s_tstart) {
    const int nt = p.ntypes, nbin = p.nbin;
    for (int k = tid; k < 2 * P * rstride; k += kThreads) s_hist[k] = 0u;
    for (int k = tid; k < nbin + 3; k += kThreads) {
        // slot k <-> bin g = k-1
        const int g = k - 1;
        double2 t;
        if (g >= 0 && g < nbin) {
            t.x = p.thr[g];
            t.y = p.thr[g + 1];
        } else {
            t.x = t.y = INFINITY;
        }
        s_thr2[k] = t;
    }
    if (EDGES)
        for (int k = tid; k <= nbin; k += kThreads) s_thrf[k] = p.thr_full[k];
    for (int k = tid; k < nt * nt; k += kThreads) {
        // Gofrt::get_itype, reference lib/include/gofrt.h:86-104
        int a = k / nt, b = k % nt;
        if (b < a) {
            const int c = a;
            a = b;
            b = c;
        }
        s_rowtab[k] = static_cast<unsigned int>((P - (b + 1) * (b + 2) / 2 + a) * rstride + p.glo);
    }
    for (int k = tid; k <= nt; k += kThreads) s_tstart[k] = p.type_start[k];
}

// Merge the CTA's shared-memory rows (guard bins left out) into lag row `lag` of the global histogram and
// zero them.  Called between barriers.
__device__ __forceinline__ void cta_flush(const PairParams &p, int tid, int lag, int hlen, int rstride,
                                          unsigned int *s_hist) {
    unsigned long long *g = p.ghist + static_cast<size_t>(lag) * hlen;
    for (int k = tid; k < hlen; k += kThreads) {
        unsigned int *w = s_hist + (k / p.nbin) * rstride + p.glo + (k % p.nbin);
        const unsigned int v = *w;
        if (v) {
            atomicAdd(g + k, static_cast<unsigned long long>(v));
            *w = 0u;
        }
    }
}

// UBOX: every frame of the window has the same box, passed as a kernel parameter.  The half edges,
// the -2*l_half constants and the tilt factors then reach the FP64 instructions as constant-bank /
// uniform-register operands instead of per-thread registers.  B200's register file delivers one even
// and one odd 32-bit register per cycle, so a DFMA with three distinct register pairs holds the issue
// port for 3 cycles and a DSETP with two for 2 (tools/pipe_probe2.cu); with uniform constants they cost
// 2 and 1, which frees issue cycles for the integer / FP32 instructions of the binning.
template <bool TRI, bool FAST, int MODE, bool UBOX>
__global__ void __launch_bounds__(kThreads, kMinBlocks) pair_kernel(const PairParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr bool EDGES = MODE == MODE_EDGES;
    const SmemLayout L = smem_layout(p.ntypes, p.nbin, p.nhi, p.glo, EDGES);
    double *s_stage = reinterpret_cast<double *>(smem + L.stage);
    double2 *s_thr2 = reinterpret_cast<double2 *>(smem + L.thr2);
    double *s_thrf = reinterpret_cast<double *>(smem + L.thr_full);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.bars);
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(smem + L.hist);
    unsigned int *s_rowtab = reinterpret_cast<unsigned int *>(smem + L.rowtab);
    int *s_tstart = reinterpret_cast<int *>(smem + L.tstart);
    unsigned int *s_sched = reinterpret_cast<unsigned int *>(smem + L.sched);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nt = p.ntypes, nbin = p.nbin;
    const int P = nt * (nt + 1) / 2;
    const int hlen = 2 * P * nbin;                 // words of one lag in the global histogram
    const int rstride = row_stride(p.nhi, p.glo);   // words of one row in shared memory (with its guard bins)
    const unsigned int self_off = static_cast<unsigned int>(P * rstride);

    cta_tables<EDGES>(p, tid, P, rstride, s_hist, s_thr2, s_thrf, s_rowtab, s_tstart);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&s_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    PairConst c;
    c.thr2_addr = smem_u32(s_thr2);
    c.hist_addr = smem_u32(s_hist);
    const uint32_t stage_addr = smem_u32(s_stage);

    int cur_t = -1;
    unsigned long long acc_pairs = 0;
    unsigned int gt = 0;  // tiles consumed by this CTA so far (stage = gt & 1, parity = (gt >> 1) & 1)
    unsigned long long edges = 0;
    bool wrap_ok = true;

    for (;;) {
        if (tid == 0) s_sched[0] = p.unit_begin + atomicAdd(p.counter, 1u);
        __syncthreads();
        const unsigned int u = s_sched[0];
        __syncthreads();
        if (u >= p.unit_end) break;

        const int jc = static_cast<int>(u % static_cast<unsigned int>(p.n_jchunks));
        const unsigned int u2 = u / static_cast<unsigned int>(p.n_jchunks);
        const int itile = static_cast<int>(u2 % static_cast<unsigned int>(p.n_itiles));
        const Job job = job_at(p, u2 / static_cast<unsigned int>(p.n_itiles));

        const int jbeg = jc * p.jchunk;
        const int jend = min(jbeg + p.jchunk, p.npad);
        const unsigned long long unit_pairs = static_cast<unsigned long long>(kTileI) * (jend - jbeg);

        if (job.tout != cur_t || acc_pairs + unit_pairs > 0xF0000000ull) {
            if (cur_t >= 0) {
                cta_flush(p, tid, cur_t, hlen, rstride, s_hist);
                __syncthreads();
            }
            cur_t = job.tout;
            acc_pairs = 0;
        }
        acc_pairs += unit_pairs;

        // ---- the box of frame fi (used for both atoms, reference basetrajectory.h:190-191) ----
        if (UBOX) {
            c.box.lhx = p.ubox[0];
            c.box.lhy = p.ubox[1];
            c.box.lhz = p.ubox[2];
            c.box.xy = p.ubox[3];
            c.box.xz = p.ubox[4];
            c.box.yz = p.ubox[5];
            // the predicated DADDs of wrap_abs want -2*l_half in REGISTERS: as a constant-bank operand ptxas
            // re-loads it (a predicated LDC.64) next to every wrap.  The xor with a run-time zero keeps it
            // from folding the value back into the constant bank.
            const int zero = p.npad >> 31;
            c.nLx = __hiloint2double(__double2hiint(p.ubox[6]) ^ zero, __double2loint(p.ubox[6]));
            c.nLy = __hiloint2double(__double2hiint(p.ubox[7]) ^ zero, __double2loint(p.ubox[7]));
            c.nLz = __hiloint2double(__double2hiint(p.ubox[8]) ^ zero, __double2loint(p.ubox[8]));
        c.box.nxy = __hiloint2double(__double2hiint(p.ubox[9]) ^ zero, __double2loint(p.ubox[9]));
        c.box.nxz = __hiloint2double(__double2hiint(p.ubox[10]) ^ zero, __double2loint(p.ubox[10]));
        c.box.nyz = __hiloint2double(__double2hiint(p.ubox[11]) ^ zero, __double2loint(p.ubox[11]));
            c.box.nxy = __hiloint2double(__double2hiint(p.ubox[9]) ^ zero, __double2loint(p.ubox[9]));
            c.box.nxz = __hiloint2double(__double2hiint(p.ubox[10]) ^ zero, __double2loint(p.ubox[10]));
            c.box.nyz = __hiloint2double(__double2hiint(p.ubox[11]) ^ zero, __double2loint(p.ubox[11]));
        } else {
            const double *bx = p.box + static_cast<size_t>(job.fi) * 6;
            c.box.lhx = __ldg(bx + 0);
            c.box.lhy = __ldg(bx + 1);
            c.box.lhz = __ldg(bx + 2);
            c.box.xy = __ldg(bx + 3);
            c.box.xz = __ldg(bx + 4);
            c.box.yz = __ldg(bx + 5);
            c.box.nxy = -c.box.xy;
            c.box.nxz = -c.box.xz;
            c.box.nyz = -c.box.yz;
            c.nLx = __dmul_rn(c.box.lhx, -2.0);
            c.nLy = __dmul_rn(c.box.lhy, -2.0);
            c.nLz = __dmul_rn(c.box.lhz, -2.0);
        }

        // ---- this thread's i atoms (frame fi) ----
        double xi[kIPT], yi[kIPT], zi[kIPT];
        int ii[kIPT], ti[kIPT];
        const double *pi = p.pos + static_cast<size_t>(job.fi) * 3 * p.npad;
        const int wi0 = itile * kTileI + warp * (32 * kIPT);  // this warp's i atoms: [wi0, wi0 + 32*kIPT)
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            const int idx = wi0 + k * 32 + lane;
            ii[k] = idx;
            if (idx < p.npad) {
                xi[k] = pi[idx];
                yi[k] = pi[p.npad + idx];
                zi[k] = pi[2 * static_cast<size_t>(p.npad) + idx];
                ti[k] = p.type_pad[idx];
            } else {
                xi[k] = yi[k] = zi[k] = __longlong_as_double(0x7ff8000000000000ll);
                ti[k] = 0;
            }
        }

        // ---- j tiles of frame fj through the bulk-copy pipeline ----
        const double *pj = p.pos + static_cast<size_t>(job.fj) * 3 * p.npad;
        const int ntile = (jend - jbeg + kTileJ - 1) / kTileJ;
        auto issue = [&](int tl, unsigned int g) {
            const int j0 = jbeg + tl * kTileJ;
            const int cnt = min(kTileJ, jend - j0);
            const unsigned int st = g & 1u;
            double *dst = s_stage + static_cast<size_t>(st) * 3 * kTileJ;
            const uint32_t bytes = static_cast<uint32_t>(cnt) * 8u;
            mbar_expect_tx(&s_bar[st], 3u * bytes);
            bulk_g2s(dst, pj + j0, bytes, &s_bar[st]);
            bulk_g2s(dst + kTileJ, pj + p.npad + j0, bytes, &s_bar[st]);
            bulk_g2s(dst + 2 * kTileJ, pj + 2 * static_cast<size_t>(p.npad) + j0, bytes, &s_bar[st]);
        };
        if (tid == 0) issue(0, gt);

        for (int tl = 0; tl < ntile; ++tl, ++gt) {
            __syncthreads();  // every thread is done with the stage tile tl+1 will overwrite
            if (tid == 0 && tl + 1 < ntile) issue(tl + 1, gt + 1);
            mbar_wait(&s_bar[gt & 1u], (gt >> 1) & 1u);

            const uint32_t sx_addr = stage_addr + (gt & 1u) * (3u * kTileJ * 8u);
            const int j0 = jbeg + tl * kTileJ;
            const int j1 = min(j0 + kTileJ, jend);
            if (MODE == MODE_SAFE2 && (warp & (kWarpsPerCta / 2)) && p.skew)
                phase_skew(p, __fadd_rn(p.bias0, static_cast<float>(s_rowtab[0])), c.hist_addr - 4u * 0x4B400000u, lane);

            for (int ty = 0; ty < nt; ++ty) {
                const int lo = max(j0, s_tstart[ty]);
                const int hi = min(j1, s_tstart[ty + 1]);
                if (lo >= hi) continue;
                unsigned int row[kIPT];
#pragma unroll
                for (int k = 0; k < kIPT; ++k) row[k] = s_rowtab[ti[k] * nt + ty];
                // [lo,hi) = before | overlap with this warp's own atoms | after   (all multiples of kPadGroup)
                const int da = min(max(wi0, lo), hi);
                const int db = min(max(wi0 + 32 * kIPT, lo), hi);
                if (MODE == MODE_SAFE2) {
                    // fast groups: every group but those that hold one of this warp's own atoms (i == j goes to another
                    // row: the diagonal segment takes the clamped form of MODE_SAFE_DENSE, which shares the rows)
                    float bias[kIPT];
#pragma unroll
                    for (int k = 0; k < kIPT; ++k) bias[k] = __fadd_rn(p.bias0, static_cast<float>(row[k]));
                    const uint32_t kaddr = c.hist_addr - 4u * 0x4B400000u;
                    auto fast = [&](int j) {
                        const unsigned int near = group_dense2<TRI>(p, c, xi, yi, zi, bias, kaddr, sx_addr, kTileJ * 8u, j - j0);
                        if (near) {
                            IAtoms ia;
#pragma unroll
                            for (int k = 0; k < kIPT; ++k) {
                                ia.x[k] = xi[k];
                                ia.y[k] = yi[k];
                                ia.z[k] = zi[k];
                                ia.bias[k] = bias[k];
                            }
                            group_fix2<TRI>(c, ia, p.inv_lo, p.inv_hi, p.bias0, p.smax, sx_addr, kTileJ * 8u, j - j0, s_thr2, p.nbin,
                                            p.inv_dr, p.c0, s_hist);
                        }
                    };
#pragma unroll 1
                    for (int j = lo; j < da; j += kJU) fast(j);
#pragma unroll 1
                    for (int j = da; j < db; j += kJU)
                        process_group<TRI, FAST, MODE_SAFE_DENSE, true>(p, c, xi, yi, zi, ii, row, sx_addr, kTileJ * 8u, j - j0, j,
                                                                        s_thr2, s_thrf, s_hist, self_off, edges, wrap_ok);
#pragma unroll 1
                    for (int j = db; j < hi; j += kJU) fast(j);
                    continue;
                }
#pragma unroll 1
                for (int j = lo; j < da; j += kJU)
                    process_group<TRI, FAST, MODE, false>(p, c, xi, yi, zi, ii, row, sx_addr, kTileJ * 8u, j - j0, j, s_thr2, s_thrf,
                                                          s_hist, self_off, edges, wrap_ok);
#pragma unroll 1
                for (int j = da; j < db; j += kJU)
                    process_group<TRI, FAST, MODE, true>(p, c, xi, yi, zi, ii, row, sx_addr, kTileJ * 8u, j - j0, j, s_thr2, s_thrf,
                                                         s_hist, self_off, edges, wrap_ok);
#pragma unroll 1
                for (int j = db; j < hi; j += kJU)
                    process_group<TRI, FAST, MODE, false>(p, c, xi, yi, zi, ii, row, sx_addr, kTileJ * 8u, j - j0, j, s_thr2, s_thrf,
                                                          s_hist, self_off, edges, wrap_ok);
            }
        }
    }

    // ---- final merge of the CTA's histogram ----
    __syncthreads();
    if (cur_t >= 0) {
        unsigned long long *g = p.ghist + static_cast<size_t>(cur_t) * hlen;
        for (int k = tid; k < hlen; k += kThreads) {
            const unsigned int v = s_hist[(k / nbin) * rstride + p.glo + (k % nbin)];
            if (v) atomicAdd(g + k, static_cast<unsigned long long>(v));
        }
    }
    if (EDGES) {
        for (int o = 16; o > 0; o >>= 1) edges += __shfl_xor_sync(0xffffffffu, edges, o);
        if (lane == 0 && edges) atomicAdd(p.edges, edges);
    }
    if (!wrap_ok) atomicExch(p.error_flag, 1u);
}

// ---------------------------------------------------------------------------------------------
// the pair kernel for SMALL systems (npad <= kSmallMax = 512 slots: the 56 atoms of the reference's bundled
// tests/data/lammps.bin, ab-initio cells of a few hundred atoms)
//
// pair_kernel gives a whole CTA (512 i slots) to one (lag, origin) job; with a few dozen atoms nine lanes in
// ten hold no atom, and because consecutive tickets go to different CTAs nearly every job ends with a merge
// of the CTA's histogram into the global row of its lag.  Here
//   * the jobs of the launch (lag-major) are cut into one contiguous range per CTA, of equal COST: all jobs cost the
//     same except those of lag 0, where every atom meets itself at distance 0 -- which the safe-zone binning sees as
//     "on a bin edge" (r = 0 is an edge of the extended grid whenever rmin / dr is an integer) and sends through its
//     correction path (2.4x the time, measured on C1); a warp that holds such a job leaves the self pairs out of the
//     main pass (ZSELF) and the remaining difference is a weight (small_w16 / 16).  No tickets, no tail;
//   * a CTA keeps NB histograms in shared memory, one per lag, and walks its range NB lags at a time (C1: a CTA's
//     483 jobs touch two lags of 756 jobs each -- one pass, one merge into the global rows at the end);
//   * the jobs of a pass are split over the warps by cost, and a warp works on a BATCH of JB jobs at a time with
//     the i slots of the batch PACKED over its lanes: slot s = (job s / npad, atom s % npad), 64 slots per round, two
//     consecutive slots (always the same job: npad is even) per thread.  56 atoms x 8 jobs fill 7 rounds of 64
//     lanes completely, where one job per warp leaves 8 of 64 slots empty;
//   * the warp copies the JB j frames of the batch into its own slice of the stage area with 16-byte cp.async
//     (one exposed round trip per batch), and every thread walks the j atoms of ITS job through process_group --
//     same arithmetic, same binning modes as the tile kernel -- adding to the histogram of ITS job's lag; the j
//     addresses of a warp differ only by the job (at most a few distinct addresses per LDS).
// Warps never wait for each other inside a pass.
// ---------------------------------------------------------------------------------------------
constexpr int kWarps = kThreads / 32;
constexpr int kSubTile = 32 * kIPT;   // i slots per warp and round
static_assert(kIPT == 2, "pair_small_kernel packs two consecutive slots per thread");

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// cost of the jobs [0, k) of a lag-major list whose first n0 jobs are lag 0 (in sixteenths of a job), and its inverse
// (the number of whole jobs that fit into cost x)
struct SmallCost {
    unsigned long long n0, w16;
    __device__ __forceinline__ unsigned long long upto(unsigned long long k) const {
        return k < n0 ? k * w16 : n0 * w16 + (k - n0) * 16ull;
    }
    __device__ __forceinline__ unsigned int jobs_in(unsigned long long x) const {
        return static_cast<unsigned int>(x < n0 * w16 ? x / w16 : n0 + (x - n0 * w16) / 16ull);
    }
    // boundary `part` of `parts` equal-cost pieces of the jobs [a, b)
    __device__ __forceinline__ unsigned int cut(unsigned int a, unsigned int b, unsigned int part, unsigned int parts) const {
        if (part >= parts) return b;
        const unsigned long long ca = upto(a), cb = upto(b);
        return max(a, min(b, jobs_in(ca + (cb - ca) * part / parts)));
    }
};

template <bool TRI, bool FAST, int MODE, bool UBOX>
__global__ void __launch_bounds__(kThreads, kMinBlocks) pair_small_kernel(const PairParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr bool EDGES = MODE == MODE_EDGES;
    const int JB = p.n_itiles, js = p.n_jchunks;   // jobs per batch, slots between the j frames of a batch
    const int NB = p.small_nb;                      // histograms (lags) a CTA holds at a time
    const SmemLayout L = smem_layout(p.ntypes, p.nbin, p.nhi, p.glo, EDGES, small_stage_slots(JB, js), NB);
    double *s_stage = reinterpret_cast<double *>(smem + L.stage);
    double2 *s_thr2 = reinterpret_cast<double2 *>(smem + L.thr2);
    double *s_thrf = reinterpret_cast<double *>(smem + L.thr_full);
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(smem + L.hist);
    unsigned int *s_rowtab = reinterpret_cast<unsigned int *>(smem + L.rowtab);
    int *s_tstart = reinterpret_cast<int *>(smem + L.tstart);
    unsigned int *s_sched = reinterpret_cast<unsigned int *>(smem + L.sched);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nt = p.ntypes, nbin = p.nbin, npad = p.npad;
    const int P = nt * (nt + 1) / 2;
    const int hlen = 2 * P * nbin;
    const int rstride = row_stride(p.nhi, p.glo);
    const unsigned int self_off = static_cast<unsigned int>(P * rstride);
    const int hwords = 2 * P * rstride;             // words of one histogram

    cta_tables<EDGES>(p, tid, P, rstride, s_hist, s_thr2, s_thrf, s_rowtab, s_tstart);
    for (int k = hwords + tid; k < NB * hwords; k += kThreads) s_hist[k] = 0u;
    __syncthreads();

    PairConst c;
    c.thr2_addr = smem_u32(s_thr2);
    c.hist_addr = smem_u32(s_hist);

    // the warp's slice of the stage area: [3 coordinates][JB jobs][js slots]
    const int crow = JB * js;
    double *area = s_stage + static_cast<size_t>(warp) * 3 * crow;
    const uint32_t area_addr = smem_u32(area);
    const uint32_t jrow_bytes = static_cast<uint32_t>(crow) * 8u;

    if (UBOX) {
        c.box.lhx = p.ubox[0];
        c.box.lhy = p.ubox[1];
        c.box.lhz = p.ubox[2];
        c.box.xy = p.ubox[3];
        c.box.xz = p.ubox[4];
        c.box.yz = p.ubox[5];
        const int zero = p.npad >> 31;   // see pair_kernel: keeps -2*l_half in registers
        c.nLx = __hiloint2double(__double2hiint(p.ubox[6]) ^ zero, __double2loint(p.ubox[6]));
        c.nLy = __hiloint2double(__double2hiint(p.ubox[7]) ^ zero, __double2loint(p.ubox[7]));
        c.nLz = __hiloint2double(__double2hiint(p.ubox[8]) ^ zero, __double2loint(p.ubox[8]));
        c.box.nxy = __hiloint2double(__double2hiint(p.ubox[9]) ^ zero, __double2loint(p.ubox[9]));
        c.box.nxz = __hiloint2double(__double2hiint(p.ubox[10]) ^ zero, __double2loint(p.ubox[10]));
        c.box.nyz = __hiloint2double(__double2hiint(p.ubox[11]) ^ zero, __double2loint(p.ubox[11]));
    }

    unsigned long long edges = 0;
    bool wrap_ok = true;

    // this CTA's share of the jobs [unit_begin, unit_end) of the launch
    SmallCost cost;
    cost.n0 = static_cast<unsigned long long>(p.jchunk);
    cost.w16 = static_cast<unsigned long long>(p.small_w16);
    const unsigned int cta_a = cost.cut(p.unit_begin, p.unit_end, blockIdx.x, gridDim.x);
    const unsigned int cta_b = cost.cut(p.unit_begin, p.unit_end, blockIdx.x + 1, gridDim.x);
    // a pass ends before its 32-bit shared-memory counters could wrap (every pair adds at most one)
    const unsigned int pass_cap = 0x7fffffffu / static_cast<unsigned int>(npad * npad);
    const unsigned int norig = static_cast<unsigned int>(p.imp_norig), every = static_cast<unsigned int>(p.imp_every);

    for (unsigned int s0 = cta_a; s0 < cta_b;) {
        // ---- the pass [s0, s1): jobs of the lags l0 .. l0 + NB - 1 (lag = index * every) ----
        unsigned int s1, l0;
        if (p.imp) {
            l0 = s0 / norig;
            s1 = min(cta_b, (l0 + static_cast<unsigned int>(NB)) * norig);
        } else {
            l0 = static_cast<unsigned int>(p.jobs[s0].tout) / every;
            if (tid == 0) s_sched[1] = cta_b;
            __syncthreads();
            for (unsigned int k = s0 + 1u + tid; k < cta_b; k += kThreads)
                if (static_cast<unsigned int>(p.jobs[k].tout) / every >= l0 + static_cast<unsigned int>(NB)) {
                    atomicMin(&s_sched[1], k);
                    break;
                }
            __syncthreads();
            s1 = s_sched[1];
        }
        s1 = min(s1, s0 + pass_cap);
        const unsigned int wa = cost.cut(s0, s1, warp, kWarps), wb = cost.cut(s0, s1, warp + 1, kWarps);

        for (unsigned int b0 = wa; b0 < wb; b0 += JB) {
            const int nj = min(JB, static_cast<int>(wb - b0));
            // lane l holds job l of the batch (lanes past the batch repeat its last job; never used)
            const Job mine = job_at(p, b0 + static_cast<unsigned int>(min(lane, nj - 1)));
            const int mybuf = static_cast<int>(static_cast<unsigned int>(mine.tout) / every - l0);   // its histogram
            __syncwarp();   // every lane is done with the previous batch's j frames
            for (int jl = 0; jl < nj; ++jl) {
                const int fj = __shfl_sync(0xffffffffu, mine.fj, jl);
                const double *src = p.pos + static_cast<size_t>(fj) * 3 * npad;
                const uint32_t dst = area_addr + static_cast<uint32_t>(jl * js) * 8u;
                for (int q = 2 * lane; q < npad; q += 64) {
                    cp_async16(dst + static_cast<uint32_t>(q) * 8u, src + q);
                    cp_async16(dst + jrow_bytes + static_cast<uint32_t>(q) * 8u, src + npad + q);
                    cp_async16(dst + 2u * jrow_bytes + static_cast<uint32_t>(q) * 8u, src + 2 * npad + q);
                }
            }
            // the packed slot of this thread in round 0: (job jl, atoms atom, atom + 1)
            int jl = 0, atom = 2 * lane;
            while (atom >= npad) {
                atom -= npad;
                ++jl;
            }
            cp_async_wait_all();
            __syncwarp();

            const int total = nj * npad;
            for (int base = 0; base < total; base += kSubTile) {
                const bool valid = jl < nj;
                const int jv = valid ? jl : 0, av = valid ? atom : 0;
                const int fi = __shfl_sync(0xffffffffu, mine.fi, jv);
                unsigned int *hist = s_hist + __shfl_sync(0xffffffffu, mybuf, jv) * hwords;
                c.hist_addr = smem_u32(hist);
                if (!UBOX) {
                    const double *bx = p.box + static_cast<size_t>(fi) * 6;
                    c.box.lhx = __ldg(bx + 0);
                    c.box.lhy = __ldg(bx + 1);
                    c.box.lhz = __ldg(bx + 2);
                    c.box.xy = __ldg(bx + 3);
                    c.box.xz = __ldg(bx + 4);
                    c.box.yz = __ldg(bx + 5);
                    c.box.nxy = -c.box.xy;
                    c.box.nxz = -c.box.xz;
                    c.box.nyz = -c.box.yz;
                    c.nLx = __dmul_rn(c.box.lhx, -2.0);
                    c.nLy = __dmul_rn(c.box.lhy, -2.0);
                    c.nLz = __dmul_rn(c.box.lhz, -2.0);
                }
                double xi[kIPT], yi[kIPT], zi[kIPT];
                int ii[kIPT], ti[kIPT];
                {
                    const double *pi = p.pos + static_cast<size_t>(fi) * 3 * npad + av;
                    const double2 vx = *reinterpret_cast<const double2 *>(pi);
                    const double2 vy = *reinterpret_cast<const double2 *>(pi + npad);
                    const double2 vz = *reinterpret_cast<const double2 *>(pi + 2 * npad);
                    const int2 vt = *reinterpret_cast<const int2 *>(p.type_pad + av);
                    const double ghost = __longlong_as_double(0x7ff8000000000000ll);
                    xi[0] = valid ? vx.x : ghost;
                    xi[1] = valid ? vx.y : ghost;
                    yi[0] = valid ? vy.x : ghost;
                    yi[1] = valid ? vy.y : ghost;
                    zi[0] = valid ? vz.x : ghost;
                    zi[1] = valid ? vz.y : ghost;
                    ti[0] = vt.x;
                    ti[1] = vt.y;
                    ii[0] = av;
                    ii[1] = av + 1;
                }
                const uint32_t jaddr = area_addr + static_cast<uint32_t>(jv * js) * 8u;

                // Lag 0 (fi == fj): every atom meets itself at distance 0, which the safe-zone binning sees as "on a bin
                // edge" (r = 0 is an edge of the extended grid whenever rmin / dr is an integer) and sends through its
                // correction path -- 2.4x the time of any other job (measured, r2w).  A warp that holds such a job leaves the
                // self pairs out of the main pass instead (three instructions per pair, on 1 job in 189 of C1).
                constexpr bool kSafe = MODE == MODE_SAFE || MODE == MODE_SAFE_DENSE;
                const int tout = __shfl_sync(0xffffffffu, mine.tout, jv);   // (every lane takes part: no short-circuit)
                const bool zself = kSafe && __any_sync(0xffffffffu, valid && tout == 0);
                for (int ty = 0; ty < nt; ++ty) {
                    const int lo = s_tstart[ty], hi = s_tstart[ty + 1];
                    if (lo >= hi) continue;
                    unsigned int row[kIPT];
#pragma unroll
                    for (int k = 0; k < kIPT; ++k) row[k] = s_rowtab[ti[k] * nt + ty];
                    // every pair, the atom with itself included, goes to the "different atoms" row here: with a few
                    // dozen atoms EVERY group holds some of the warp's own atoms, and a per-pair i == j row select
                    // would cost three instructions on each of the N^2 pairs; the N self pairs are moved below
                    if (!zself) {
#pragma unroll 1
                        for (int j = lo; j < hi; j += kJU)
                            process_group<TRI, FAST, MODE, false>(p, c, xi, yi, zi, ii, row, jaddr, jrow_bytes, j, j, s_thr2,
                                                                  s_thrf, hist, self_off, edges, wrap_ok);
                    } else {
#pragma unroll 1
                        for (int j = lo; j < hi; j += kJU)
                            process_group<TRI, FAST, MODE, false, true>(p, c, xi, yi, zi, ii, row, jaddr, jrow_bytes, j, j,
                                                                        s_thr2, s_thrf, hist, self_off, edges, wrap_ok);
                    }
                }
                // the self pairs (i, frame fi) - (i, frame fj): from the "different atoms" row of the atom's type to its
                // "same atom" row (or only into the latter, when the main pass has left them out), at the exact bin of
                // the threshold table (where the pass above -- fast path plus correction, or a bracket search -- has
                // counted them)
                if (valid) {
                    const double *sj = area + static_cast<size_t>(jv) * js + av;
#pragma unroll
                    for (int k = 0; k < kIPT; ++k) {
                        double dx = __dsub_rn(xi[k], sj[k]);
                        double dy = __dsub_rn(yi[k], sj[crow + k]);
                        double dz = __dsub_rn(zi[k], sj[2 * crow + k]);
                        if (FAST)
                            min_image_single<TRI>(dx, dy, dz, c.box, c.nLx, c.nLy, c.nLz);
                        else
                            wrap_ok &= min_image_general<TRI>(dx, dy, dz, c.box);
                        const double d2 = d2_of(dx, dy, dz);
                        if ((d2 >= s_thr2[1].x) && (d2 < s_thr2[nbin].y)) {   // false for the NaN of a ghost slot
                            int g = static_cast<int>(bin_guess1(d2, p.inv_dr, p.c0, static_cast<unsigned int>(nbin) + 2u)) - 1;
                            g = min(max(g, 0), nbin - 1);
                            while (d2 < s_thr2[g + 1].x) --g;
                            while (d2 >= s_thr2[g + 1].y) ++g;
                            unsigned int *w = hist + s_rowtab[ti[k] * nt + ti[k]] + g;
                            if (!zself) atomicAdd(w, 0xffffffffu);
                            atomicAdd(w + self_off, 1u);
                        }
                    }
                }
                // the thread's slot of the next round
                atom += kSubTile;
                while (atom >= npad) {
                    atom -= npad;
                    ++jl;
                }
            }
        }
        // ---- the histograms of the pass into the global rows of their lags ----
        __syncthreads();
        const unsigned int l_last = p.imp ? (s1 - 1u) / norig : static_cast<unsigned int>(p.jobs[s1 - 1u].tout) / every;
        for (unsigned int l = l0; l <= l_last; ++l)
            cta_flush(p, tid, static_cast<int>(l * every), hlen, rstride, s_hist + (l - l0) * hwords);
        __syncthreads();
        s0 = s1;
    }
    if (EDGES) {
        for (int o = 16; o > 0; o >>= 1) edges += __shfl_xor_sync(0xffffffffu, edges, o);
        if (lane == 0 && edges) atomicAdd(p.edges, edges);
    }
    if (!wrap_ok) atomicExch(p.error_flag, 1u);
}

// variant = TRI | FAST<<1 | MODE<<2 | UBOX<<5 | SMALL<<6   (UBOX only with FAST and MODE in {THR, SAFE, SAFE_DENSE})
template <int V>
static cudaError_t launch_variant(int grid, size_t smem, cudaStream_t stream, const PairParams &p) {
    if constexpr ((V & 64) != 0)
        pair_small_kernel<(V & 1) != 0, (V & 2) != 0, ((V >> 2) & 7), (V & 32) != 0><<<grid, kThreads, smem, stream>>>(p);
    else
        pair_kernel<(V & 1) != 0, (V & 2) != 0, ((V >> 2) & 7), (V & 32) != 0><<<grid, kThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int V>
static cudaError_t prepare_variant(size_t max_smem) {
    if constexpr ((V & 64) != 0)
        return cudaFuncSetAttribute(pair_small_kernel<(V & 1) != 0, (V & 2) != 0, ((V >> 2) & 7), (V & 32) != 0>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(max_smem));
    else
        return cudaFuncSetAttribute(pair_kernel<(V & 1) != 0, (V & 2) != 0, ((V >> 2) & 7), (V & 32) != 0>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(max_smem));
}

#define AGOFRT_VARIANTS_OF(X, S)                                                                              \
    X(S + 0) X(S + 1) X(S + 2) X(S + 3) X(S + 4) X(S + 5) X(S + 6) X(S + 7) X(S + 8) X(S + 9) X(S + 10)     \
    X(S + 11) X(S + 12) X(S + 13) X(S + 14) X(S + 15) X(S + 18) X(S + 19)                                   \
    X(S + 32 + 2) X(S + 32 + 3) X(S + 32 + 14) X(S + 32 + 15) X(S + 32 + 18) X(S + 32 + 19)
#define AGOFRT_VARIANTS(X) AGOFRT_VARIANTS_OF(X, 0) AGOFRT_VARIANTS_OF(X, 64) X(22) X(23) X(32 + 22) X(32 + 23)

cudaError_t launch_pair_kernel(int variant, int grid, size_t smem, cudaStream_t stream, const PairParams &p) {
    switch (variant) {
#define AGOFRT_CASE(V) \
    case V: return launch_variant<V>(grid, smem, stream, p);
        AGOFRT_VARIANTS(AGOFRT_CASE)
#undef AGOFRT_CASE
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t prepare_pair_kernels(size_t max_smem) {
    cudaError_t e;
#define AGOFRT_PREP(V)                \
    e = prepare_variant<V>(max_smem); \
    if (e != cudaSuccess) return e;
    AGOFRT_VARIANTS(AGOFRT_PREP)
#undef AGOFRT_PREP
    return cudaSuccess;
}

// Device-side validation of MODE_SAFE's float guess (run once per plan): for every probe value the
// guess must either be flagged "within eps of an edge" or equal the exact bin expected[k].
__global__ void validate_safe_kernel(const double *__restrict__ probes, const int *__restrict__ expected, int n,
                                     float inv_dr, float c0h, float lim, float qmax, int nbin, int glo,
                                     unsigned int *bad) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const SafeGuess sg = safe_guess(probes[k], inv_dr, c0h, qmax);
    const int e = expected[k];  // exact bin, or -1 when the reference does not count the pair
    if (sg.n < -glo || sg.n > nbin) atomicAdd(bad, 1u);   // would leave the row and its guard bins
    if (fabsf(sg.dl) < lim) {
        // an unflagged guess is final: it must be the exact bin, or a guard bin when the pair is not counted
        const bool counted = sg.n >= 0 && sg.n < nbin;
        const bool should = e >= 0 && e < nbin;
        if (counted != should || (counted && sg.n != e)) atomicAdd(bad, 1u);
    }
}

cudaError_t launch_validate_safe(const double *probes, const int *expected, int n, float inv_dr, float c0h, float lim,
                                 float qmax, int nbin, int glo, unsigned int *bad, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    validate_safe_kernel<<<(n + 255) / 256, 256, 0, stream>>>(probes, expected, n, inv_dr, c0h, lim, qmax, nbin, glo, bad);
    return cudaGetLastError();
}

__global__ void validate_safe2_kernel(const double *__restrict__ probes, const int *__restrict__ expected, int n,
                                      float inv_lo, float inv_hi, float bias0, float smax, int nbin, int glo,
                                      unsigned int *bad) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double d2 = probes[k];   // every probe counts here: far values, +inf and NaN must end in the guard bin
    float ra, rb;
    safe2_floors(safe2_root(cvt_rz(d2), smax), inv_lo, inv_hi, bias0, ra, rb);
    const int g = __float_as_int(ra) - 0x4B400000;   // guessed bin (bias0 = 1.5*2^23 + c0, row 0)
    if (g < -glo || g > nbin) atomicAdd(bad, 1u);   // would leave the row and its guard bins
    if (__float_as_int(ra) == __float_as_int(rb)) {
        const int e = expected[k];
        const bool counted = g >= 0 && g < nbin, should = e >= 0 && e < nbin;
        if (counted != should || (counted && g != e)) atomicAdd(bad, 1u);
    }
}

cudaError_t launch_validate_safe2(const double *probes, const int *expected, int n, float inv_lo, float inv_hi, float bias0,
                                  float smax, int nbin, int glo, unsigned int *bad, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    validate_safe2_kernel<<<(n + 255) / 256, 256, 0, stream>>>(probes, expected, n, inv_lo, inv_hi, bias0, smax, nbin, glo, bad);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// layout kernels
// ---------------------------------------------------------------------------------------------
__global__ void gather_soa_kernel(const double *__restrict__ pos_aos, const int *__restrict__ perm, int natoms,
                                  int npad, int nframes, double *__restrict__ pos_soa) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (k >= npad || f >= nframes) return;
    const int src = perm[k];
    double x, y, z;
    if (src >= 0) {
        const double *a = pos_aos + (static_cast<size_t>(f) * natoms + src) * 3;
        x = a[0];
        y = a[1];
        z = a[2];
    } else {
        x = y = z = __longlong_as_double(0x7ff8000000000000ll);
    }
    double *o = pos_soa + static_cast<size_t>(f) * 3 * npad;
    o[k] = x;
    o[npad + k] = y;
    o[2 * static_cast<size_t>(npad) + k] = z;
}

cudaError_t launch_gather_soa(const double *pos_aos, const int *perm, int natoms, int npad, int nframes,
                              double *pos_soa, cudaStream_t stream) {
    if (nframes <= 0) return cudaSuccess;
    dim3 grid((npad + 255) / 256, nframes);
    gather_soa_kernel<<<grid, 256, 0, stream>>>(pos_aos, perm, natoms, npad, nframes, pos_soa);
    return cudaGetLastError();
}

__global__ void scatter_aos_kernel(const double *__restrict__ soa, const int *__restrict__ perm, int natoms, int npad,
                                   int nframes, double *__restrict__ aos) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (k >= npad || f >= nframes) return;
    const int dst = perm[k];
    if (dst < 0) return;
    const double *s = soa + static_cast<size_t>(f) * 3 * npad;
    double *a = aos + (static_cast<size_t>(f) * natoms + dst) * 3;
    a[0] = s[k];
    a[1] = s[npad + k];
    a[2] = s[2 * static_cast<size_t>(npad) + k];
}

cudaError_t launch_scatter_aos(const double *pos_soa, const int *perm, int natoms, int npad, int nframes, double *pos_aos,
                               cudaStream_t stream) {
    if (nframes <= 0 || npad <= 0) return cudaSuccess;
    dim3 grid((npad + 255) / 256, nframes);
    scatter_aos_kernel<<<grid, 256, 0, stream>>>(pos_soa, perm, natoms, npad, nframes, pos_aos);
    return cudaGetLastError();
}

// The per-atom part of Trajectory::set_access_at's frame loop (reference lib/src/trajectory.cpp:629-646) for raw dump
// records: id -> slot through the flat table, type check, x y z to the atom's slot of the AoS staging frame.
// raw: [nframes][natoms][8] doubles (id type x y z vx vy vz) in file order; aos: [nframes][natoms][3].
__global__ void parse_records_kernel(const double *__restrict__ raw, int natoms, int nframes, const int *__restrict__ id_table,
                                     int table_len, const int *__restrict__ slot_type, double *__restrict__ aos,
                                     unsigned int *flags) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (r >= natoms || f >= nframes) return;
    const double *rec = raw + (static_cast<size_t>(f) * natoms + r) * 8;
    // two 32-byte halves of the record: one sector each
    const double4 a = *reinterpret_cast<const double4 *>(rec);       // id type x y
    const double z = rec[4];
    const long long id = llrint(a.x);   // std::round of an integral double: the same integer
    int slot = -1;
    if (id >= 0 && id < table_len) slot = id_table[id];
    if (slot < 0) {
        atomicAdd(flags + 4, 1u);
        return;
    }
    if (static_cast<int>(llrint(a.y)) != slot_type[slot]) atomicAdd(flags + 5, 1u);
    double *o = aos + (static_cast<size_t>(f) * natoms + slot) * 3;
    o[0] = a.z;
    o[1] = a.w;
    o[2] = z;
}

cudaError_t launch_parse_records(const double *raw, int natoms, int nframes, const int *id_table, int table_len,
                                 const int *slot_type, double *aos, unsigned int *flags, cudaStream_t stream) {
    if (nframes <= 0 || natoms <= 0) return cudaSuccess;
    dim3 grid((natoms + 255) / 256, nframes);
    parse_records_kernel<<<grid, 256, 0, stream>>>(raw, natoms, nframes, id_table, table_len, slot_type, aos, flags);
    return cudaGetLastError();
}

__global__ void pack_box_kernel(const double *__restrict__ in, int stride, int nframes, double *__restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const double *r = in + static_cast<size_t>(f) * stride;
    double *o = out + static_cast<size_t>(f) * 6;
    o[0] = r[3];
    o[1] = r[4];
    o[2] = r[5];
    o[3] = stride == 9 ? r[6] : 0.0;
    o[4] = stride == 9 ? r[7] : 0.0;
    o[5] = stride == 9 ? r[8] : 0.0;
}

cudaError_t launch_pack_box(const double *box_internal, int stride, int nframes, double *box6, cudaStream_t stream) {
    if (nframes <= 0) return cudaSuccess;
    pack_box_kernel<<<(nframes + 127) / 128, 128, 0, stream>>>(box_internal, stride, nframes, box6);
    return cudaGetLastError();
}

__global__ void frame_bounds_kernel(const double *__restrict__ pos_soa, const int *__restrict__ perm, int npad,
                                    double *__restrict__ bounds, unsigned int *inf_flag) {
    const int f = blockIdx.x, c = blockIdx.y;
    const double *row = pos_soa + (static_cast<size_t>(f) * 3 + c) * npad;
    double lo = INFINITY, hi = -INFINITY;
    bool inf = false, nan = false;
    for (int k = threadIdx.x; k < npad; k += blockDim.x) {
        const double v = row[k];
        if (isinf(v)) inf = true;
        if (v != v && perm[k] >= 0) nan = true;  // a NaN that is not a ghost slot
        lo = fmin(lo, v);  // fmin/fmax drop NaN (ghost slots, NaN input)
        hi = fmax(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ double slo[32], shi[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        slo[w] = lo;
        shi[w] = hi;
    }
    if (inf) atomicExch(inf_flag, 1u);
    if (nan) atomicExch(inf_flag + 2, 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int k = 1; k < nw; ++k) {
            lo = fmin(lo, slo[k]);
            hi = fmax(hi, shi[k]);
        }
        bounds[static_cast<size_t>(f) * 6 + c] = lo;
        bounds[static_cast<size_t>(f) * 6 + 3 + c] = hi;
    }
}

cudaError_t launch_frame_bounds(const double *pos_soa, const int *perm, int npad, int nframes, double *bounds6,
                                unsigned int *inf_flag, cudaStream_t stream) {
    if (nframes <= 0) return cudaSuccess;
    dim3 grid(nframes, 3);
    frame_bounds_kernel<<<grid, 256, 0, stream>>>(pos_soa, perm, npad, bounds6, inf_flag);
    return cudaGetLastError();
}

// BaseTrajectory::pbc_wrap, reference lib/include/basetrajectory.h:145-161
__global__ void pbc_wrap_kernel(double *pos_aos, int natoms, int nframes, const double *__restrict__ box, int stride,
                                unsigned int *error_flag) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (a >= natoms || f >= nframes) return;
    const double *r = box + static_cast<size_t>(f) * stride;
    BoxRegs b;
    b.lhx = r[3];
    b.lhy = r[4];
    b.lhz = r[5];
    b.xy = stride == 9 ? r[6] : 0.0;
    b.xz = stride == 9 ? r[7] : 0.0;
    b.yz = stride == 9 ? r[8] : 0.0;
    double *x = pos_aos + (static_cast<size_t>(f) * natoms + a) * 3;
    double dx = __dsub_rn(x[0], b.lhx), dy = __dsub_rn(x[1], b.lhy), dz = __dsub_rn(x[2], b.lhz);
    bool ok;
    if (stride == 9)
        ok = min_image_general<true>(dx, dy, dz, b);
    else
        ok = min_image_general<false>(dx, dy, dz, b);
    x[0] = __dadd_rn(dx, b.lhx);
    x[1] = __dadd_rn(dy, b.lhy);
    x[2] = __dadd_rn(dz, b.lhz);
    if (!ok) atomicExch(error_flag, 1u);
}

cudaError_t launch_pbc_wrap(double *pos_aos, int natoms, int nframes, const double *box_internal, int stride,
                            unsigned int *error_flag, cudaStream_t stream) {
    if (nframes <= 0 || natoms <= 0) return cudaSuccess;
    dim3 grid((natoms + 255) / 256, nframes);
    pbc_wrap_kernel<<<grid, 256, 0, stream>>>(pos_aos, natoms, nframes, box_internal, stride, error_flag);
    return cudaGetLastError();
}

// BaseTrajectory::d2_minImage(i,j,it,jt,x) for every ordered pair (tests)
__global__ void d2_all_kernel(const double *__restrict__ pi, const double *__restrict__ pj,
                              const double *__restrict__ box6, int triclinic, const int *__restrict__ perm, int natoms,
                              int npad, double *__restrict__ out, unsigned int *error_flag) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= npad || i >= npad) return;
    const int oi = perm[i], oj = perm[j];
    if (oi < 0 || oj < 0) return;
    BoxRegs b;
    b.lhx = box6[0];
    b.lhy = box6[1];
    b.lhz = box6[2];
    b.xy = box6[3];
    b.xz = box6[4];
    b.yz = box6[5];
    double dx = __dsub_rn(pi[i], pj[j]);
    double dy = __dsub_rn(pi[npad + i], pj[npad + j]);
    double dz = __dsub_rn(pi[2 * static_cast<size_t>(npad) + i], pj[2 * static_cast<size_t>(npad) + j]);
    const bool ok = triclinic ? min_image_general<true>(dx, dy, dz, b) : min_image_general<false>(dx, dy, dz, b);
    double *o = out + (static_cast<size_t>(oi) * natoms + oj) * 4;
    o[0] = dx;
    o[1] = dy;
    o[2] = dz;
    o[3] = d2_of(dx, dy, dz);
    if (!ok) atomicExch(error_flag, 1u);
}

cudaError_t launch_d2_all(const double *pos_i, const double *pos_j, const double *box6, int triclinic, const int *perm,
                          int natoms, int npad, double *out, unsigned int *error_flag, cudaStream_t stream) {
    dim3 grid((npad + 127) / 128, npad);
    d2_all_kernel<<<grid, 128, 0, stream>>>(pos_i, pos_j, box6, triclinic, perm, natoms, npad, out, error_flag);
    return cudaGetLastError();
}

// BaseTrajectory::d2_minImage(i,j,it,jt,x) of ONE pair given by device slots (host-class probe)
__global__ void d2_pair_kernel(const double *__restrict__ pi, const double *__restrict__ pj,
                               const double *__restrict__ box6, int triclinic, int si, int sj, int npad,
                               double *__restrict__ out, unsigned int *error_flag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    BoxRegs b;
    b.lhx = box6[0];
    b.lhy = box6[1];
    b.lhz = box6[2];
    b.xy = box6[3];
    b.xz = box6[4];
    b.yz = box6[5];
    double dx = __dsub_rn(pi[si], pj[sj]);
    double dy = __dsub_rn(pi[npad + si], pj[npad + sj]);
    double dz = __dsub_rn(pi[2 * static_cast<size_t>(npad) + si], pj[2 * static_cast<size_t>(npad) + sj]);
    const bool ok = triclinic ? min_image_general<true>(dx, dy, dz, b) : min_image_general<false>(dx, dy, dz, b);
    out[0] = dx;
    out[1] = dy;
    out[2] = dz;
    out[3] = d2_of(dx, dy, dz);
    if (!ok) atomicExch(error_flag, 1u);
}

cudaError_t launch_d2_pair(const double *pos_i, const double *pos_j, const double *box6, int triclinic, int slot_i,
                           int slot_j, int npad, double *out4, unsigned int *error_flag, cudaStream_t stream) {
    d2_pair_kernel<<<1, 32, 0, stream>>>(pos_i, pos_j, box6, triclinic, slot_i, slot_j, npad, out4, error_flag);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Neighbour-count histogram (IstogrammaAtomiRaggio::calculate, reference lib/src/istogrammaatomiraggio.cpp:31-85):
// for every listed frame and every atom i, the number of atoms j (j == i included) of each type with
// d2_minImage(i,j,frame,frame) < r2, then hist[type][count] += 1.  Same distance arithmetic as the pair kernel;
// the device layout is type-major, so the count of one type is a plain loop over that type's slot range.
// One work unit = (frame, tile of kNbThreads*kNbIPT i atoms); j coordinates go through shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kNbThreads = 256;
constexpr int kNbIPT = 2;
constexpr int kNbTileJ = 512;

// Two steps, so that the j atoms of one (frame, i tile) can be spread over several work units (few frames of a large
// system: 192 (frame, tile) units would leave most of the 148 SMs idle):
//   neighbour_kernel        unit = (frame, tile of 512 i atoms, chunk of j atoms): per i atom and j type the number of j
//                           of the chunk within r, added to counts[frame][type][slot] (u32, atomic: chunks meet there)
//   neighbour_finish_kernel hist[type][counts[frame][type][slot]] += 1 for every real atom
template <bool TRI, bool FAST>
__global__ void __launch_bounds__(kNbThreads) neighbour_kernel(const NeighbourParams p) {
    __shared__ __align__(16) double sj[3][kNbTileJ];
    const int tid = threadIdx.x;
    unsigned int wrap_bad = 0;
    const unsigned int per_frame = static_cast<unsigned int>(p.n_itiles) * static_cast<unsigned int>(p.n_jchunks);
    for (unsigned int u = p.unit_begin + blockIdx.x; u < p.unit_end; u += gridDim.x) {
        const unsigned int fidx = u / per_frame, rest = u - fidx * per_frame;
        const int itile = static_cast<int>(rest / static_cast<unsigned int>(p.n_jchunks));
        const int jc = static_cast<int>(rest % static_cast<unsigned int>(p.n_jchunks));
        const int f = p.frames[fidx];
        const double *pf = p.pos + static_cast<size_t>(f) * 3 * p.npad;
        const double *bx = p.box + static_cast<size_t>(f) * 6;
        BoxRegs b;
        b.lhx = bx[0];
        b.lhy = bx[1];
        b.lhz = bx[2];
        b.xy = bx[3];
        b.xz = bx[4];
        b.yz = bx[5];
        b.nxy = -b.xy;
        b.nxz = -b.xz;
        b.nyz = -b.yz;
        const double nLx = __dmul_rn(b.lhx, -2.0), nLy = __dmul_rn(b.lhy, -2.0), nLz = __dmul_rn(b.lhz, -2.0);
        double xi[kNbIPT], yi[kNbIPT], zi[kNbIPT];
        int slot[kNbIPT];
#pragma unroll
        for (int k = 0; k < kNbIPT; ++k) {
            slot[k] = itile * (kNbThreads * kNbIPT) + k * kNbThreads + tid;
            if (slot[k] < p.npad) {
                xi[k] = pf[slot[k]];
                yi[k] = pf[p.npad + slot[k]];
                zi[k] = pf[2 * static_cast<size_t>(p.npad) + slot[k]];
            } else {
                xi[k] = yi[k] = zi[k] = __longlong_as_double(0x7ff8000000000000ll);
            }
        }
        const int cbeg = jc * p.jchunk, cend = min(cbeg + p.jchunk, p.npad);
        for (int ty = 0; ty < p.ntypes; ++ty) {
            const int jb = max(p.type_start[ty], cbeg), je = min(p.type_start[ty + 1], cend);
            if (jb >= je) continue;
            unsigned int cnt[kNbIPT];
#pragma unroll
            for (int k = 0; k < kNbIPT; ++k) cnt[k] = 0;
            for (int j0 = jb; j0 < je; j0 += kNbTileJ) {
                const int n = min(kNbTileJ, je - j0);
                __syncthreads();
                for (int q = tid; q < n; q += kNbThreads) {
                    sj[0][q] = pf[j0 + q];
                    sj[1][q] = pf[p.npad + j0 + q];
                    sj[2][q] = pf[2 * static_cast<size_t>(p.npad) + j0 + q];
                }
                __syncthreads();
#pragma unroll 4
                for (int q = 0; q < n; ++q) {
                    const double xj = sj[0][q], yj = sj[1][q], zj = sj[2][q];
#pragma unroll
                    for (int k = 0; k < kNbIPT; ++k) {
                        double dx = __dsub_rn(xi[k], xj), dy = __dsub_rn(yi[k], yj), dz = __dsub_rn(zi[k], zj);
                        if (FAST) {
                            min_image_single<TRI>(dx, dy, dz, b, nLx, nLy, nLz);
                        } else {
                            if (!min_image_general<TRI>(dx, dy, dz, b)) wrap_bad = 1;
                        }
                        // NaN (ghost slots, NaN coordinates) compares false: never a neighbour, as in the reference
                        cnt[k] += d2_of(dx, dy, dz) < p.r2 ? 1u : 0u;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kNbIPT; ++k)
                if (slot[k] < p.npad && cnt[k])
                    atomicAdd(p.counts + (static_cast<size_t>(fidx) * p.ntypes + ty) * p.npad + slot[k], cnt[k]);
        }
    }
    if (wrap_bad) atomicExch(p.error_flag, 1u);
}

// the (frame, i tile) pairs [v_begin, v_end) of this device: their counts are complete (all j chunks ran here)
__global__ void neighbour_finish_kernel(const NeighbourParams p, unsigned int v_begin, unsigned int v_end) {
    const int tile = kNbThreads * kNbIPT;
    const size_t n = static_cast<size_t>(v_end - v_begin) * p.ntypes * tile;
    for (size_t k = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int s = static_cast<int>(k % tile);
        const int ty = static_cast<int>((k / tile) % p.ntypes);
        const unsigned int v = v_begin + static_cast<unsigned int>(k / (static_cast<size_t>(tile) * p.ntypes));
        const unsigned int fidx = v / static_cast<unsigned int>(p.n_itiles), itile = v % static_cast<unsigned int>(p.n_itiles);
        const int slot = static_cast<int>(itile) * tile + s;
        if (slot >= p.npad || p.perm[slot] < 0) continue;   // past the end, or a ghost slot
        const unsigned int c = p.counts[(static_cast<size_t>(fidx) * p.ntypes + ty) * p.npad + slot];
        atomicAdd(p.hist + static_cast<size_t>(ty) * p.hist_stride + c, 1ull);
    }
}

cudaError_t launch_neighbour_kernel(bool triclinic, bool fast, int grid, cudaStream_t stream, const NeighbourParams &p) {
    if (grid <= 0) return cudaSuccess;
    if (triclinic) {
        if (fast)
            neighbour_kernel<true, true><<<grid, kNbThreads, 0, stream>>>(p);
        else
            neighbour_kernel<true, false><<<grid, kNbThreads, 0, stream>>>(p);
    } else {
        if (fast)
            neighbour_kernel<false, true><<<grid, kNbThreads, 0, stream>>>(p);
        else
            neighbour_kernel<false, false><<<grid, kNbThreads, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t launch_neighbour_finish(int grid, cudaStream_t stream, const NeighbourParams &p, unsigned int v_begin, unsigned int v_end) {
    if (grid <= 0 || v_end <= v_begin) return cudaSuccess;
    neighbour_finish_kernel<<<grid, 256, 0, stream>>>(p, v_begin, v_end);
    return cudaGetLastError();
}

int neighbour_tile_atoms() { return kNbThreads * kNbIPT; }

// ---------------------------------------------------------------------------------------------
// Neighbour lists and spherical-harmonic densities: the other two pair loops over d2_minImage of the reference
// (Neighbours::update_neigh, lib/src/neighbour.cpp:8-62; SphericalBase::calc, lib/src/sphericalbase.cpp:19-68).
// Both need the minimum-image VECTOR xi - xj with its signs, so they take the literal minimum image; both are
// order-dependent per atom (a list is filled in the order the partners come, a density is a floating-point sum over
// the partners), so ONE THREAD WALKS ALL PARTNERS OF ITS ATOM IN THE REFERENCE'S ORDER (ascending atom index); the
// partners' coordinates are staged through shared memory, every thread of the block reading the same one.
// With 10^5 atoms that is 3000 warps of 10^5 pair evaluations each: enough to fill the FP64 pipes.
// ---------------------------------------------------------------------------------------------
constexpr int kLsThreads = 128;
constexpr int kLsTile = 256;

__global__ void __launch_bounds__(kLsThreads) neigh_list_kernel(const NeighListParams p) {
    __shared__ double sj[3][kLsTile];
    __shared__ int stype[kLsTile];
    const int i = blockIdx.x * kLsThreads + threadIdx.x;
    const bool live = i < p.natoms;
    const double *pf = p.pos + static_cast<size_t>(p.frame) * 3 * p.npad;
    const double *bx = p.box + static_cast<size_t>(p.frame) * 6;
    BoxRegs b;
    b.lhx = bx[0];
    b.lhy = bx[1];
    b.lhz = bx[2];
    b.xy = bx[3];
    b.xz = bx[4];
    b.yz = bx[5];
    b.nxy = b.nxz = b.nyz = 0.0;
    double xi = 0, yi = 0, zi = 0;
    int ti = 0;
    if (live) {
        const int s = p.atom_slot[i];
        xi = pf[s];
        yi = pf[p.npad + s];
        zi = pf[2 * static_cast<size_t>(p.npad) + s];
        ti = p.atom_type[i];
    }
    const double cut_i = live ? p.cutoff2[ti] : 0.0;
    unsigned long long cnt[kLsMaxTypes];
#pragma unroll
    for (int k = 0; k < kLsMaxTypes; ++k) cnt[k] = 0;
    bool bad = false, wrap_bad = false;
    for (int j0 = 0; j0 < p.natoms; j0 += kLsTile) {
        const int n = min(kLsTile, p.natoms - j0);
        __syncthreads();
        for (int q = threadIdx.x; q < n; q += kLsThreads) {
            const int s = p.atom_slot[j0 + q];
            sj[0][q] = pf[s];
            sj[1][q] = pf[p.npad + s];
            sj[2][q] = pf[2 * static_cast<size_t>(p.npad) + s];
            stype[q] = p.atom_type[j0 + q];
        }
        __syncthreads();
        if (!live) continue;
        for (int q = 0; q < n; ++q) {
            const int j = j0 + q;
            if (j == i) continue;
            // the reference evaluates the pair once, as (smaller index) - (larger index), and hands the larger one the
            // negated vector (neighbour.cpp:18, :35-38); the minimum image is odd, so that IS minImage(xi - xj)
            double dx = __dsub_rn(xi, sj[0][q]), dy = __dsub_rn(yi, sj[1][q]), dz = __dsub_rn(zi, sj[2][q]);
            if (p.triclinic) {
                if (!min_image_general<true>(dx, dy, dz, b)) wrap_bad = true;
            } else {
                if (!min_image_general<false>(dx, dy, dz, b)) wrap_bad = true;
            }
            const double d2 = d2_of(dx, dy, dz);
            const int tj = stype[q];
            const double cut_j = p.cutoff2[tj];
            if (d2 <= cut_j || d2 <= cut_i) {
                const unsigned long long n_here = cnt[tj];
                if (n_here >= p.nneigh[tj]) {
                    bad = true;   // "Too many neighbours in shell!"
                } else {
                    unsigned long long *li = p.list + p.list_offset[tj] + static_cast<size_t>(i) * (p.nneigh[tj] + 1) + 1;
                    double *ri = p.rpos + p.rpos_offset[tj] + (static_cast<size_t>(i) * p.nneigh[tj] + n_here) * 4;
                    li[n_here] = static_cast<unsigned long long>(j);
                    ri[0] = __dsqrt_rn(d2);
                    ri[1] = dx;
                    ri[2] = dy;
                    ri[3] = dz;
                    if (d2 <= cut_j) cnt[tj] = n_here + 1;
                }
            }
        }
    }
    if (live) {
        for (int tj = 0; tj < p.ntypes; ++tj) {
            unsigned long long *li = p.list + p.list_offset[tj] + static_cast<size_t>(i) * (p.nneigh[tj] + 1);
            li[0] = cnt[tj];
            if (p.sort && cnt[tj] > 1) {
                // ascending distance (the reference: std::sort on the distances, neighbour.cpp:47-69); insertion sort, the
                // lists hold a few dozen entries
                double *r = p.rpos + p.rpos_offset[tj] + static_cast<size_t>(i) * p.nneigh[tj] * 4;
                for (unsigned long long a = 1; a < cnt[tj]; ++a) {
                    const double r0 = r[a * 4], r1 = r[a * 4 + 1], r2 = r[a * 4 + 2], r3 = r[a * 4 + 3];
                    const unsigned long long id = li[1 + a];
                    unsigned long long c = a;
                    while (c > 0 && r[(c - 1) * 4] > r0) {
                        r[c * 4] = r[(c - 1) * 4];
                        r[c * 4 + 1] = r[(c - 1) * 4 + 1];
                        r[c * 4 + 2] = r[(c - 1) * 4 + 2];
                        r[c * 4 + 3] = r[(c - 1) * 4 + 3];
                        li[1 + c] = li[c];
                        --c;
                    }
                    r[c * 4] = r0;
                    r[c * 4 + 1] = r1;
                    r[c * 4 + 2] = r2;
                    r[c * 4 + 3] = r3;
                    li[1 + c] = id;
                }
            }
        }
    }
    if (bad) atomicExch(p.flags, 1u);
    if (wrap_bad) atomicExch(p.flags + 1, 1u);
}

cudaError_t launch_neigh_list(const NeighListParams &p, cudaStream_t stream) {
    if (p.natoms <= 0) return cudaSuccess;
    neigh_list_kernel<<<(p.natoms + kLsThreads - 1) / kLsThreads, kLsThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

// Real spherical harmonics of the direction (x, y, z) for l = 0 .. lmax, the operations of
// SpecialFunctions::SphericalHarmonics<lmax,double,true,true>::calc (reference lib/include/specialfunctions.h:244-388)
// one by one, each rounded on its own: cos(theta) = z/r, sin / cos(phi) = y, x over sqrt(x^2+y^2); associated Legendre
// polynomials by the l,l / l+1,l / l+1,m recursions; cos(m phi), sin(m phi) by the Chebyshev recursion; then
// (P * coeff) * trig.  out: the reference's dynamic layout, blocks l = lmax .. 0 of (l negative m's, then m = 0 .. l).
// plm is a scratch of (lmax+1)^2 doubles indexed [l][m].
__device__ void real_spherical_harmonics(int lmax, double x, double y, double z, const double *__restrict__ coeff,
                                         double *__restrict__ plm, double *__restrict__ out) {
    const double rxy = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
    const double r = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    const double cost = __ddiv_rn(z, r), sinp = __ddiv_rn(y, rxy), cosp = __ddiv_rn(x, rxy);
    const int L1 = lmax + 1;
    // P^m_l: the diagonal first, then from each diagonal element the column m upwards in l
    plm[0] = 1.0;
    const double sq = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(cost, cost)));
    for (int l = 1; l <= lmax; ++l)   // P^l_l = (-(2l-1) * sqrt(1-x^2)) * P^{l-1}_{l-1}
        plm[l * L1 + l] = __dmul_rn(__dmul_rn(static_cast<double>(-(2 * l - 1)), sq), plm[(l - 1) * L1 + (l - 1)]);
    for (int m = 0; m < lmax; ++m) {
        // P^m_{m+1} = (P^m_m * x) * (2m+1)
        plm[(m + 1) * L1 + m] = __dmul_rn(__dmul_rn(plm[m * L1 + m], cost), static_cast<double>(2 * m + 1));
        for (int l = m + 2; l <= lmax; ++l)   // P^m_l = ((P^m_{l-1} * x) * (2l-1) - P^m_{l-2} * (l+m-1)) / (l-m)
            plm[l * L1 + m] = __ddiv_rn(__dsub_rn(__dmul_rn(__dmul_rn(plm[(l - 1) * L1 + m], cost), static_cast<double>(2 * l - 1)),
                                                  __dmul_rn(plm[(l - 2) * L1 + m], static_cast<double>(l + m - 1))),
                                        static_cast<double>(l - m));
    }
    // cos(m phi), sin(m phi): c_m = (2 cosp) * c_{m-1} - c_{m-2}
    double cm[kShMaxL + 1], sm[kShMaxL + 1];
    cm[0] = 1.0;
    sm[0] = 0.0;
    if (lmax >= 1) {
        cm[1] = cosp;
        sm[1] = sinp;
    }
    const double two_c = __dmul_rn(2.0, cosp);
    for (int m = 2; m <= lmax; ++m) {
        cm[m] = __dsub_rn(__dmul_rn(two_c, cm[m - 1]), cm[m - 2]);
        sm[m] = __dsub_rn(__dmul_rn(two_c, sm[m - 1]), sm[m - 2]);
    }
    // (P * coeff) * trig, into the reference's layout
    int base = 0;
    for (int l = lmax; l >= 0; --l) {
        double *minus = out + base, *plus = out + base + l;
        for (int m = 0; m <= l; ++m) plus[m] = __dmul_rn(__dmul_rn(plm[l * L1 + m], coeff[l * L1 + m]), cm[m]);
        for (int k = 0; k < l; ++k) {   // val_minus[k] is m = -(k+1): the same P and coefficient as m = k+1, times sin
            const int m = k + 1;
            minus[k] = __dmul_rn(__dmul_rn(plm[l * L1 + m], coeff[l * L1 + m]), sm[m]);
        }
        base += 2 * l + 1;
    }
}

__global__ void __launch_bounds__(kLsThreads) sh_density_kernel(const ShDensityParams p) {
    __shared__ double sj[3][kLsTile];
    __shared__ int stype[kLsTile];
    const int i = blockIdx.x * kLsThreads + threadIdx.x;
    const bool live = i < p.natoms;
    const int nl = (p.lmax + 1) * (p.lmax + 1);
    const double *pf = p.pos + static_cast<size_t>(p.frame) * 3 * p.npad;
    const double *bx = p.box + static_cast<size_t>(p.frame) * 6;
    BoxRegs b;
    b.lhx = bx[0];
    b.lhy = bx[1];
    b.lhz = bx[2];
    b.xy = bx[3];
    b.xz = bx[4];
    b.yz = bx[5];
    b.nxy = b.nxz = b.nyz = 0.0;
    double xi = 0, yi = 0, zi = 0;
    int ti = 0;
    if (live) {
        const int s = p.atom_slot[i];
        xi = pf[s];
        yi = pf[p.npad + s];
        zi = pf[2 * static_cast<size_t>(p.npad) + s];
        ti = p.atom_type[i];
    }
    double plm[(kShMaxL + 1) * (kShMaxL + 1)], ylm[(kShMaxL + 1) * (kShMaxL + 1)];
    bool wrap_bad = false;
    for (int j0 = 0; j0 < p.natoms; j0 += kLsTile) {
        const int n = min(kLsTile, p.natoms - j0);
        __syncthreads();
        for (int q = threadIdx.x; q < n; q += kLsThreads) {
            const int s = p.atom_slot[j0 + q];
            sj[0][q] = pf[s];
            sj[1][q] = pf[p.npad + s];
            sj[2][q] = pf[2 * static_cast<size_t>(p.npad) + s];
            stype[q] = p.atom_type[j0 + q];
        }
        __syncthreads();
        if (!live) continue;
        for (int q = 0; q < n; ++q) {
            const int j = j0 + q;
            if (j == i) continue;
            double dx = __dsub_rn(xi, sj[0][q]), dy = __dsub_rn(yi, sj[1][q]), dz = __dsub_rn(zi, sj[2][q]);
            if (p.triclinic) {
                if (!min_image_general<true>(dx, dy, dz, b)) wrap_bad = true;
            } else {
                if (!min_image_general<false>(dx, dy, dz, b)) wrap_bad = true;
            }
            const double d = __dsqrt_rn(d2_of(dx, dy, dz));
            const int tj = stype[q];
            const int pair = p.ntypes * ti + tj;
            // int idx = (int) floorf((d - rmin) / dr): the double quotient is rounded to float first (sphericalbase.cpp:51)
            const float qf = __double2float_rn(__ddiv_rn(__dsub_rn(d, p.rmin[pair]), p.dr[pair]));
            const float fl = floorf(qf);
            if (!(fl >= 0.0f) || !(fl < static_cast<float>(p.nbin))) continue;   // also NaN
            const int idx = static_cast<int>(fl);
            real_spherical_harmonics(p.lmax, dx, dy, dz, p.coeff, plm, ylm);
            const size_t cell = (static_cast<size_t>(p.nbin) * (static_cast<size_t>(p.ntypes) * i + tj) + idx);
            double *res = p.result + cell * nl;
            for (int ll = 0; ll < nl; ++ll) res[ll] = __dadd_rn(res[ll], ylm[ll]);
            if (p.counter) p.counter[cell] += 1;
        }
    }
    if (wrap_bad) atomicExch(p.flags + 1, 1u);
}

cudaError_t launch_sh_density(const ShDensityParams &p, cudaStream_t stream) {
    if (p.natoms <= 0) return cudaSuccess;
    sh_density_kernel<<<(p.natoms + kLsThreads - 1) / kLsThreads, kLsThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Mean square displacement (MSD<T>::calc_single_th, reference lib/src/msd.cpp:63-125): for lag t and type ty the
// mean over origins and atoms of |x_i(o) - x_i(o+t)|^2 (optionally minus the displacement of the type's centre of
// mass).  No minimum image: the reference works on the coordinates as stored.  HBM/L2-bound: 48 bytes read for
// 9 FP64 operations per (atom, lag, origin).
//   msd_partial_kernel: block = (tile of 256 slots of ONE type, lag); every thread walks the origins of its atom
//                       (coalesced rows of the SoA window), then a fixed-order tree sum -> partial[lag][tile]
//   msd_finish_kernel:  one thread per (lag, type): partials in tile order / count; and the centre-of-mass MSD as
//                       the reference's own sequential running mean over the origins (bit-identical)
// The atom part is a sum divided by a count where the reference keeps a running mean: equal to rounding
// (tests: 1e-12 relative), deterministic from run to run.
// ---------------------------------------------------------------------------------------------
constexpr int kMsdThreads = 256;

constexpr int kMsdLags = 16;   // lags one block walks together

// block = (chunk of kMsdLags lags, tile of 256 atoms of one type).  A thread keeps one running sum per lag of the chunk and
// walks the origins once: the frame of the origin is read once for the whole chunk, and the frames origin + lag of
// neighbouring origins overlap (origin stride < chunk), so they come from L1/L2 -- a window larger than L2 streams from
// HBM about (2 + stride) / kMsdLags... times per chunk instead of twice per LAG.  The sums of one lag are formed in the
// same order as before (origins in order per thread, then the fixed tree): the same partials, bit for bit.
__global__ void __launch_bounds__(kMsdThreads) msd_partial_kernel(const MsdParams p) {
    __shared__ double red[kMsdThreads];
    const int t0 = blockIdx.x * kMsdLags, tid = threadIdx.x;
    const int nl = min(kMsdLags, p.leff - t0);
    for (int tile = blockIdx.y; tile < p.ntiles; tile += gridDim.y) {
        const int ty = p.tile_type[tile];
        const int slot = p.tile_start[tile] + tid;
        const bool live = tid < p.tile_count[tile];
        double acc[kMsdLags];
#pragma unroll
        for (int l = 0; l < kMsdLags; ++l) acc[l] = 0.0;
        if (live) {
            const size_t row = static_cast<size_t>(p.npad);
            for (int im = 0; im < p.ntimesteps; im += p.skip) {
                const size_t fa = static_cast<size_t>(p.f0 + im);
                const double *pa = p.pos + fa * 3 * row;
                const double xa = pa[slot], ya = pa[row + slot], za = pa[2 * row + slot];
                double cax = 0, cay = 0, caz = 0;
                if (p.cm_self) {
                    const double *ca = p.cm + (fa * p.ntypes + ty) * 3;
                    cax = ca[0];
                    cay = ca[1];
                    caz = ca[2];
                }
#pragma unroll
                for (int l = 0; l < kMsdLags; ++l) {
                    if (l < nl) {
                        const size_t fb = fa + t0 + l;
                        const double *pb = p.pos + fb * 3 * row;
                        double dx = __dsub_rn(xa, pb[slot]);
                        double dy = __dsub_rn(ya, pb[row + slot]);
                        double dz = __dsub_rn(za, pb[2 * row + slot]);
                        if (p.cm_self) {
                            const double *cb = p.cm + (fb * p.ntypes + ty) * 3;
                            dx = __dsub_rn(dx, __dsub_rn(cax, cb[0]));
                            dy = __dsub_rn(dy, __dsub_rn(cay, cb[1]));
                            dz = __dsub_rn(dz, __dsub_rn(caz, cb[2]));
                        }
                        acc[l] = __dadd_rn(acc[l], __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
                    }
                }
            }
        }
#pragma unroll 1
        for (int l = 0; l < nl; ++l) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < kMsdLags; ++q)
                if (q == l) v = acc[q];   // (static indexing keeps acc in registers)
            red[tid] = v;
            __syncthreads();
            for (int s = kMsdThreads / 2; s > 0; s >>= 1) {
                if (tid < s) red[tid] = __dadd_rn(red[tid], red[tid + s]);
                __syncthreads();
            }
            if (tid == 0) p.partial[static_cast<size_t>(t0 + l) * p.ntiles + tile] = red[0];
            __syncthreads();
        }
    }
}

// The same sums when every frame is an origin (skip == 1) and no centre of mass is subtracted: the later frames of
// consecutive origins overlap in all but one frame, so a thread keeps the kMsdLags later frames of its atom in a
// REGISTER RING and reads two frames per origin (the origin itself and the one later frame that enters the ring)
// instead of kMsdLags + 1 -- 6 loads for 144 FP64 operations: the kernel is bound by the FP64 pipe, not by L1.  The
// origin loop is unrolled by the ring length, so every ring index is a literal; the loads of origin o + 4 are issued
// before the arithmetic of origin o.  Every running sum sees the origins in the same order as in msd_partial_kernel:
// the same partials, bit for bit.
__global__ void __launch_bounds__(kMsdThreads) msd_ring_kernel(const MsdParams p) {
    __shared__ double red[kMsdThreads];
    constexpr int R = kMsdLags;
    const int t0 = blockIdx.x * R, tid = threadIdx.x;
    const int nl = min(R, p.leff - t0);
    const int fmax = p.f0 + p.ntimesteps - 1 + p.leff - 1;   // the last frame any lag of the block touches
    for (int tile = blockIdx.y; tile < p.ntiles; tile += gridDim.y) {
        const int slot = p.tile_start[tile] + tid;
        const bool live = tid < p.tile_count[tile];
        double acc[R];
#pragma unroll
        for (int l = 0; l < R; ++l) acc[l] = 0.0;
        if (live) {
            const size_t row = static_cast<size_t>(p.npad);
            const double *base = p.pos + slot;
            // frames past fmax are only ever asked for by the lags >= nl of the last chunk, whose sums are dropped
            auto frame = [&](int f) { return base + static_cast<size_t>(min(f, fmax)) * 3 * row; };
            double rx[R], ry[R], rz[R];
#pragma unroll
            for (int l = 0; l < R - 1; ++l) {
                const double *q = frame(p.f0 + t0 + l);
                rx[l] = q[0];
                ry[l] = q[row];
                rz[l] = q[2 * row];
            }
            // in flight: the origin frame and the entering later frame of the next D origins (an L2 round trip is
            // longer than the arithmetic of one origin: 144 FP64 instructions)
            constexpr int D = 4;
            static_assert(R % D == 0, "the stage of an origin must be a literal");
            double pxa[D], pya[D], pza[D], pxn[D], pyn[D], pzn[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const double *qa = frame(p.f0 + d), *qn = frame(p.f0 + d + t0 + R - 1);
                pxa[d] = qa[0];
                pya[d] = qa[row];
                pza[d] = qa[2 * row];
                pxn[d] = qn[0];
                pyn[d] = qn[row];
                pzn[d] = qn[2 * row];
            }
            for (int im0 = 0; im0 < p.ntimesteps; im0 += R) {
#pragma unroll
                for (int u = 0; u < R; ++u) {
                    const int im = im0 + u;
                    if (im < p.ntimesteps) {
                        rx[(u + R - 1) % R] = pxn[u % D];
                        ry[(u + R - 1) % R] = pyn[u % D];
                        rz[(u + R - 1) % R] = pzn[u % D];
                        const double x0 = pxa[u % D], y0 = pya[u % D], z0 = pza[u % D];
                        // the loads of origin im + D (clamped: past the last origin they are not used).  (An additional
                        // prefetch.global.L2 of the frames 32 origins ahead made it slower: 15.8 against 13.1 ms, r2o.)
                        const double *qa1 = frame(p.f0 + im + D), *qn1 = frame(p.f0 + im + D + t0 + R - 1);
                        pxa[u % D] = qa1[0];
                        pya[u % D] = qa1[row];
                        pza[u % D] = qa1[2 * row];
                        pxn[u % D] = qn1[0];
                        pyn[u % D] = qn1[row];
                        pzn[u % D] = qn1[2 * row];
#pragma unroll
                        for (int l = 0; l < R; ++l) {
                            const int sl = (u + l) % R;
                            const double dx = __dsub_rn(x0, rx[sl]);
                            const double dy = __dsub_rn(y0, ry[sl]);
                            const double dz = __dsub_rn(z0, rz[sl]);
                            acc[l] = __dadd_rn(acc[l], __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
                        }
                    }
                }
            }
        }
#pragma unroll 1
        for (int l = 0; l < nl; ++l) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (q == l) v = acc[q];   // (static indexing keeps acc in registers)
            red[tid] = v;
            __syncthreads();
            for (int s = kMsdThreads / 2; s > 0; s >>= 1) {
                if (tid < s) red[tid] = __dadd_rn(red[tid], red[tid + s]);
                __syncthreads();
            }
            if (tid == 0) p.partial[static_cast<size_t>(t0 + l) * p.ntiles + tile] = red[0];
            __syncthreads();
        }
    }
}

__global__ void msd_finish_kernel(const MsdParams p) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p.leff * p.ntypes) return;
    const int t = k / p.ntypes, ty = k % p.ntypes;
    const int f_cm = p.cm_msd ? 2 : 1;
    double sum = 0.0;
    for (int tile = 0; tile < p.ntiles; ++tile)
        if (p.tile_type[tile] == ty) sum = __dadd_rn(sum, p.partial[static_cast<size_t>(t) * p.ntiles + tile]);
    const int norig = (p.ntimesteps + p.skip - 1) / p.skip;
    const double count = static_cast<double>(static_cast<long long>(norig) * p.type_count[ty]);
    p.out[(static_cast<size_t>(t) * f_cm) * p.ntypes + ty] = count > 0 ? __ddiv_rn(sum, count) : 0.0;
    if (p.cm_msd) {
        double v = 0.0;
        unsigned long long cont = 0;
        for (int im = 0; im < p.ntimesteps; im += p.skip) {
            const size_t fa = static_cast<size_t>(p.f0 + im), fb = fa + t;
            const double *ca = p.cm + (fa * p.ntypes + ty) * 3, *cb = p.cm + (fb * p.ntypes + ty) * 3;
            const double dx = __dsub_rn(ca[0], cb[0]), dy = __dsub_rn(ca[1], cb[1]), dz = __dsub_rn(ca[2], cb[2]);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            const double delta = __dsub_rn(d2, v);
            v = __dadd_rn(v, __ddiv_rn(delta, static_cast<double>(++cont)));   // reference msd.cpp:111-117
        }
        p.out[(static_cast<size_t>(t) * f_cm + 1) * p.ntypes + ty] = v;
    }
}

cudaError_t launch_msd(const MsdParams &p, cudaStream_t stream) {
    if (p.leff <= 0 || p.ntypes <= 0) return cudaSuccess;
    if (p.ntiles > 0) {
        dim3 grid((p.leff + kMsdLags - 1) / kMsdLags, p.ntiles < 65535 ? p.ntiles : 65535);
        static const bool ring_ok = !(getenv("AGOFRT_MSD_RING") && atoi(getenv("AGOFRT_MSD_RING")) == 0);
        if (p.skip == 1 && !p.cm_self && ring_ok)
            msd_ring_kernel<<<grid, kMsdThreads, 0, stream>>>(p);
        else
            msd_partial_kernel<<<grid, kMsdThreads, 0, stream>>>(p);
    }
    const int n = p.leff * p.ntypes;
    msd_finish_kernel<<<(n + 127) / 128, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

int msd_tile_atoms() { return kMsdThreads; }

// ---------------------------------------------------------------------------------------------
// Block averages on the device: MediaVar<T>::calculate (reference lib/include/calcoliblocchi.h:35-56) per
// element, on the integer counts of the block that has just been computed.  The reference runs it as eight
// whole-vector VectorOp passes on the host (delta = x; delta -= mean; tmp = delta; tmp /= k+1; mean += tmp;
// tmp = x; tmp -= mean; tmp *= delta; var += tmp); per element that is the sequence below, every operation
// rounded on its own (no FMA), so mean and var are bit-identical to the host's.  HBM-bound: 40 bytes per element
// (8 read for the count, 16 read and 16 written for mean and var).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blockavg_push_kernel(const unsigned long long *__restrict__ counts, double incr,
                                                            double kp1, double *__restrict__ mean,
                                                            double *__restrict__ var, size_t len) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t k = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < len; k += stride) {
        const double x = __dmul_rn(__ull2double_rn(counts[k]), incr);   // Gofrt: vdata = count * incr
        const double m0 = mean[k];
        const double delta = __dsub_rn(x, m0);
        const double m1 = __dadd_rn(m0, __ddiv_rn(delta, kp1));
        mean[k] = m1;
        var[k] = __dadd_rn(var[k], __dmul_rn(__dsub_rn(x, m1), delta));
    }
}

cudaError_t launch_blockavg_push(const unsigned long long *counts, double incr, unsigned block_index, double *mean,
                                 double *var, size_t len, int sm_count, cudaStream_t stream) {
    if (len == 0) return cudaSuccess;
    const size_t want = (len + 255) / 256;
    const int grid = static_cast<int>(want < static_cast<size_t>(8 * sm_count) ? want : static_cast<size_t>(8 * sm_count));
    blockavg_push_kernel<<<grid, 256, 0, stream>>>(counts, incr, static_cast<double>(block_index + 1u), mean, var, len);
    return cudaGetLastError();
}

// The same for a batch of blocks (agofrt_blocks): every element walks the blocks IN ORDER, so the sequence of rounded
// operations per element is the one of nblocks single pushes -- in one launch.
__global__ void __launch_bounds__(256) blockavg_push_blocks_kernel(const unsigned long long *__restrict__ counts, unsigned nblocks,
                                                                   double incr, unsigned first_index, double *__restrict__ mean,
                                                                   double *__restrict__ var, size_t len) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t k = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < len; k += stride) {
        double m = mean[k], v = var[k];
        for (unsigned b = 0; b < nblocks; ++b) {
            const double x = __dmul_rn(__ull2double_rn(counts[static_cast<size_t>(b) * len + k]), incr);
            const double delta = __dsub_rn(x, m);
            m = __dadd_rn(m, __ddiv_rn(delta, static_cast<double>(first_index + b + 1u)));
            v = __dadd_rn(v, __dmul_rn(__dsub_rn(x, m), delta));
        }
        mean[k] = m;
        var[k] = v;
    }
}

cudaError_t launch_blockavg_push_blocks(const unsigned long long *counts, unsigned nblocks, double incr, unsigned first_index,
                                        double *mean, double *var, size_t len, int sm_count, cudaStream_t stream) {
    if (len == 0 || nblocks == 0) return cudaSuccess;
    const size_t want = (len + 255) / 256;
    const int grid = static_cast<int>(want < static_cast<size_t>(8 * sm_count) ? want : static_cast<size_t>(8 * sm_count));
    blockavg_push_blocks_kernel<<<grid, 256, 0, stream>>>(counts, nblocks, incr, first_index, mean, var, len);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// FP64 issue-rate microbenchmark: 8 independent DFMA chains per thread
// ---------------------------------------------------------------------------------------------
constexpr int kPeakChains = 8;
constexpr int kPeakUnroll = 16;

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *sink, int iters) {
    double a[kPeakChains];
    const double b = 1.0000000001, c = 1e-9 * (threadIdx.x + 1);
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) a[k] = 1.0 + 0.001 * k + 1e-6 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kPeakUnroll; ++u) {
#pragma unroll
            for (int k = 0; k < kPeakChains; ++k) a[k] = fma(a[k], b, c);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) s += a[k];
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

cudaError_t launch_dfma_peak(double *sink, int blocks, int iters, cudaStream_t stream,
                             unsigned long long *count_per_launch) {
    dfma_peak_kernel<<<blocks, 256, 0, stream>>>(sink, iters);
    if (count_per_launch)
        *count_per_launch = static_cast<unsigned long long>(blocks) * 256ull * iters * kPeakChains * kPeakUnroll;
    return cudaGetLastError();
}

}  // namespace agofrt
