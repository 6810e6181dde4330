// k3d_kernels.cu -- sm_100a kernels and C ABI for the krige3d grid-estimation path.
//
// Two kernels do, per grid node, what the reference's node loop does
// (/root/reference/pygeostatistics/krige3d.py:302-403):
//   k3d_search_kernel   super-block neighbour search     super_block.py:163-317 (repairs D3-D5)
//   k3d_solve_kernel    kriging system assembly          krige3d.py:470-583, 748-801, 838-875
//                       solve + estimate + variance      krige3d.py:590-614
// The neighbour lists travel between them through global memory (HBM is not the bound:
// 128 B per node against 6.5 TB/s), which lets each kernel have its own shape: the search
// is bounded by shared memory (12 warps per SM on config 3), the register-heavy solve runs
// 16-20.  The same kernels estimate at listed locations (cross-validation / jackknife,
// krige3d.py:310-332, 359-363) when given a K3dPoints.
//
// Mapping to the machine (DESIGN.md has the full account):
//   * one CTA per "tile" = the grid nodes that fall in one super block (or in a brick of
//     super blocks when those hold few nodes); all of them search the same super blocks;
//   * search: the tile's candidate samples are listed once per CTA, ordered by their
//     distance from the centre of the tile (counting sort) and staged in shared memory;
//     ONE THREAD PER NODE then streams them nearest-first -- all lanes of a warp read the
//     same candidate -- evaluates the reference's exact (FMA-free) anisotropic distance,
//     keeps the nearest per octant (or the ndmax nearest) in small shared-memory lists and
//     stops as soon as the triangle inequality rules out what is left;
//   * solve: one warp per node, each warp walking a boustrophedon path so that consecutive
//     nodes are neighbours and the sample-sample covariances of the previous node are
//     re-used; the covariances the nodes in flight need are dealt out over the warps of the
//     CTA in rounds of 32, evaluated from per-structure pre-scaled coordinates (no division);
//   * the (n+mdt) system is assembled in shared memory in a mirrored packed-triangular
//     layout, bordered by the right-hand side and the data vector, so that a single
//     register-resident LDL^T sweep (no pivoting: covariance block is SPD, the constraint
//     Schur complement is definite) leaves the kriging variance and the estimate in the two
//     border entries: no triangular solves, no weights materialised.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (see build.py).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "k3d.h"

#define K3D_VERSION_STR "pygeostatistics_b200 k3d 0.2 (sm_100a)"
#define K3D_EPS 2.220446049250313e-16 /* 7./3-4./3-1, krige3d.py:795,853 */
#define FULL 0xffffffffu
// A kriging system whose elimination meets a pivot of magnitude <= K3D_SING_TOL * maxcov is
// reported singular (K3D_ST_SINGULAR, NaN).  The reference hands such systems (duplicate
// samples, collinear drift columns) to LAPACK, which returns garbage weights of order 1e16
// with an "ill-conditioned matrix" warning unless the pivot is exactly zero (repair D16;
// the oracle applies the same threshold to its pivoted LU).
#define K3D_SING_TOL 1e-12
#ifndef K3D_SMALL_CTAS
#define K3D_SMALL_CTAS 3 /* resident CTAs targeted by the small-system (<= 20 rows) kernel */
#endif
#define K3D_SBINS 64  /* search: bins (in squared distance from the tile centre) of the candidates */
#define K3D_SCHUNK 16 /* search: candidates between two termination tests */
// The warps of a CTA move through the phases of a node together (search+fill /
// covariances / solve): warps that run the same loop share instruction-cache lines, which
// measured as the difference between 2.5 and 0.2 "no instruction" stall cycles per issue.
#define K3D_PHASE_BARRIER() __syncthreads()

// -DK3D_DEBUG: every index into the shared-memory carve-up that is computed from data (list
// rows, stage places, round-table slots, matrix positions) is checked, and a violation traps
// the kernel with a message.  compute-sanitizer is not available on the pool; the parity suite
// is run against this build instead (tests/test_gpu_debug_build.py).
#ifdef K3D_DEBUG
#define K3D_CHECK(cond, what, a, b)                                                        \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            printf("k3d debug: %s (%d, %d) block %d thread %d line %d\n", what, (int)(a), \
                   (int)(b), (int)blockIdx.x, (int)threadIdx.x, __LINE__);                 \
            __trap();                                                                      \
        }                                                                                  \
    } while (0)
#else
#define K3D_CHECK(cond, what, a, b) do { } while (0)
#endif

#ifdef K3D_PHASE_CLOCKS /* diagnostic build: cycles a warp spends in each phase of the solve kernel */
__device__ unsigned long long g_phase_clk[8];
#define K3D_CLK(i) do { const long long c_ = clock64(); if (lane == 0) pc_[i] += c_ - pt_; pt_ = c_; } while (0)
#else
#define K3D_CLK(i) do { } while (0)
#endif

namespace {

thread_local char g_err[512] = "";
thread_local K3dLaunchInfo g_launch = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.f, 0.f};
thread_local cudaEvent_t g_ev[3] = {nullptr, nullptr, nullptr};
thread_local int g_ev_dev = -1;   // device the events belong to
thread_local bool g_ev_pending = false;

int fail(int code, const char *fmt, const char *detail = "")
{
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

// destroys this thread's timing events (created on g_ev_dev; destroying an event does not
// need that device to be current)
void drop_events()
{
    for (int i = 0; i < 3; i++)
        if (g_ev[i]) { cudaEventDestroy(g_ev[i]); g_ev[i] = nullptr; }
    g_ev_dev = -1;
    g_ev_pending = false;
}

#define CUDA_TRY(call)                                                         \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess)                                                 \
            return fail(K3D_ECUDA, #call ": %s", cudaGetErrorString(e_));      \
    } while (0)

// ---------------------------------------------------------------------------
// kernel configuration (computed on the host, identical formulas on both sides)
// ---------------------------------------------------------------------------
// The path runs as two kernels per slab: a SEARCH kernel (many light warps per SM) that
// writes every node's neighbour list, and a SOLVE kernel (register-heavy warps) that turns
// the lists into estimates.  Both work tile by tile (tile = grid nodes of one super block).
struct SCfg {       // search kernel
    int warps;      // warps per CTA
    int cs;         // candidates a tile may list before the cull (temporaries, overlaid)
    int cst;        // staged candidates per CTA: those within reach of the tile, sorted
    int cap;        // entries per selection list: noct (octant search) or ndmax
    int nlist;      // selection lists per node: 8 octants, or 1
    int pw_bytes;   // per-warp shared bytes
    int smem_bytes; // total dynamic shared bytes
};
struct BCfg {       // solve kernel
    int warps;      // warps per CTA
    int rtot;       // ndmax + mdt + 2 rows in the bordered system
    int ntri;       // rtot*(rtot+1)/2 packed entries
    int ndp;        // ndmax rounded up to a multiple of 4
    int nst;        // number of variogram structures
    int ndb;        // block discretisation points
    int urow;       // doubles in one pivot-row buffer (register solve) or rtot (generic)
    int reuse;      // 1: keep the sample-sample covariance block between adjacent nodes
    int nsplit;     // block kriging: work items per sample of the right-hand side
    int napad;      // ndmax rounded up to a multiple of 32
    int ustride;    // doubles per point in structure space: 4 per structure (x, y, z, pad) + 2,
                    // which spreads consecutive records over all shared-memory banks
    int has_ext;    // 1: external-drift values of the neighbours are kept (ktype 2/3)
    int need_xyz;   // 1: neighbour coordinates are kept (drift monomials, block kriging)
    int pw_bytes;   // per-warp shared bytes
    int smem_bytes; // total dynamic shared bytes
};

// Variogram structures prepared on the host for the kernel: the rotation matrix of
// every structure is pre-divided by its range, so that |rs * d|^2 = (h/a)^2 and no
// division is left in the covariance (power structures keep the unscaled matrix).
struct KPrep {
    double rs[K3D_MAXNST][9];
    double cc[K3D_MAXNST];
    double aa[K3D_MAXNST];
    double eps0s;   // K3D_EPS / a0^2: the "zero distance" test in scaled units
    double epsblk;  // block kriging: a sample and a block point whose structure-0 distance
                    // h2 is at least this are more than sqrt(K3D_EPS) apart (4 EPS |rs0|_F^2)
    int it[K3D_MAXNST];
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

// shared memory carve-up (bytes from the dynamic smem base)
// search kernel
//   CTA : [stage: records of (x, y, z f64, sample index i32, bin i32) x cst+4]
//         [run table i32 x 3*threads][bin histogram u16 x warps*SBINS][bin start u16 x SBINS+1]
//         [reduction scratch f64 x 8*6]
//   warp: [keys f64 x E*32][gidx i32 x E*32][tau f64 x nlist*32][meta u16 x nlist*32]
//         E = nlist*cap entries per node, lane-interleaved: entry e of lane l at e*32 + l
//   The staging temporaries (the candidates before the sort: sample index i32, rank u16,
//   bin u8 each) overlay the per-warp regions, which are initialised afterwards.
__host__ __device__ inline int s_off_chunk(const SCfg &c) { return (c.cst + 4) * 32; }
__host__ __device__ inline int s_off_whist(const SCfg &c) { return s_off_chunk(c) + c.warps * 32 * 12; }
__host__ __device__ inline int s_off_bstart(const SCfg &c) { return s_off_whist(c) + c.warps * K3D_SBINS * 2; }
__host__ __device__ inline int s_off_red(const SCfg &c) { return align16(s_off_bstart(c) + (K3D_SBINS + 1) * 2); }
__host__ __device__ inline int s_off_warp(const SCfg &c) { return align16(s_off_red(c) + 8 * 6 * 8); }
__host__ __device__ inline int spw_bytes(const SCfg &c)
{
    const int E = c.nlist * c.cap;
    return align16(E * 320 + (c.nlist == 8 ? 0 : 8 * 256) + 8 * 64);
}
__host__ __device__ inline int s_total_bytes(const SCfg &c)
{
    const int lists = c.warps * spw_bytes(c), tmp = align16(7 * c.cs);
    return s_off_warp(c) + (lists > tmp ? lists : tmp);
}
// solve kernel
//   CTA : [tri LUT u16 x ntri][block points per structure f64 x 3*nst*ndb]
//         [round table u16 x warps*rounds_per_warp]
//   warp: [P f64 x ntri]                        pristine bordered system (persists)
//         [va(,px,py,pz)(,ve) f64 x ndp]        neighbours by matrix position (coordinates:
//                                               drift / block kriging only; ve: ktype 2/3)
//         [ut f64 x ustride*(ndp+ndb)]          their coordinates in structure space (one
//                                               record of nst x (x, y, z, pad) + 2 pad per
//                                               point), then the node's block points (ndp..)
//         [pid,newid i32 x ndp][newpos u8 x ndp] sample id by position / fill list
//         [pivot rows f64 x 2*urow]              the node's structure-space record (unode,
//                                               ustride doubles, set-up phase only) shares it
//         [part f64 x nsplit*napad]             block kriging: partial right-hand sides
__host__ __device__ inline int b_off_udb(const BCfg &c) { return align16(c.ntri * 2); } // 16-byte aligned records
__host__ __device__ inline int b_rounds_per_warp(const BCfg &c)
{   // most rounds of 32 covariance items one node can publish.  Pair rounds: the whole
    // triangle 32 pairs at a time (full rebuild), or one round per newcomer (partial update,
    // taken when 2 * newcomers <= na: up to ndp / 2 rounds, which exceeds the triangle's
    // count below ndp = 32); then the right-hand side rounds.
    const int tri_rounds = (c.ndp * (c.ndp - 1) / 2 + 31) >> 5, new_rounds = c.ndp >> 1;
    return (tri_rounds > new_rounds ? tri_rounds : new_rounds) + c.nsplit * (c.napad >> 5);
}
__host__ __device__ inline int b_off_rtab(const BCfg &c) { return align16(b_off_udb(c) + c.ustride * c.ndb * 8); }
__host__ __device__ inline int b_off_warp(const BCfg &c) { return align16(b_off_rtab(c) + c.warps * b_rounds_per_warp(c) * 2); }
__host__ __device__ inline int bpw_off_pos(const BCfg &c) { return align16(c.ntri * 8); }
__host__ __device__ inline int bpw_off_ut(const BCfg &c) { return bpw_off_pos(c) + (1 + 3 * c.need_xyz + c.has_ext) * c.ndp * 8; }
__host__ __device__ inline int bpw_off_ints(const BCfg &c) { return bpw_off_ut(c) + c.ustride * (c.ndp + c.ndb) * 8; }
__host__ __device__ inline int bpw_off_rows(const BCfg &c) { return align16(bpw_off_ints(c) + 9 * c.ndp); }
__host__ __device__ inline int bpw_off_part(const BCfg &c)
{
    const int rows = 2 * c.urow > c.ustride ? 2 * c.urow : c.ustride;
    return bpw_off_rows(c) + rows * 8;
}
__host__ __device__ inline int bpw_bytes(const BCfg &c)
{
    return align16(bpw_off_part(c) + (c.ndb > 1 ? c.nsplit * c.napad * 8 : 0));
}

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int pk(int r, int c) { return ((r * (r + 1)) >> 1) + c; } // r >= c

// super_block.py:349-378: the search distance, evaluated exactly as the reference
// does -- no fused multiply-add, (r0*dx + r1*dy) + r2*dz, squares summed from 0.0.
__device__ __forceinline__ double sqdist_exact(double dx, double dy, double dz,
                                               const double *__restrict__ r)
{
    double c0 = __dadd_rn(__dadd_rn(__dmul_rn(r[0], dx), __dmul_rn(r[1], dy)), __dmul_rn(r[2], dz));
    double c1 = __dadd_rn(__dadd_rn(__dmul_rn(r[3], dx), __dmul_rn(r[4], dy)), __dmul_rn(r[5], dz));
    double c2 = __dadd_rn(__dadd_rn(__dmul_rn(r[6], dx), __dmul_rn(r[7], dy)), __dmul_rn(r[8], dz));
    double s = __dmul_rn(c0, c0); // 0.0 + c0*c0 is exact
    s = __dadd_rn(s, __dmul_rn(c1, c1));
    s = __dadd_rn(s, __dmul_rn(c2, c2));
    return s;
}

// explicit shared-window accesses for the hottest loop (32-bit addresses, no generic-pointer
// arithmetic in the instruction stream)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds_f64x2(uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_u16(uint32_t a)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return (int)v;
}
__device__ __forceinline__ int lds_s32(uint32_t a)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// polynomial and reduction constants of exp_neg, kept in the constant bank so that every
// DFMA takes its coefficient as an operand
__constant__ double k_expc[10] = {
    46.166241308446828384,       // 32 / ln 2
    -0.021660849390173098,       // -ln2/32 high part (20 trailing zero bits)
    -2.325192846878874e-12,      // -ln2/32 low part
    1.98412698412698412698e-04,  // 1/5040
    1.38888888888888888889e-03,  // 1/720
    8.33333333333333333333e-03,  // 1/120
    4.16666666666666666667e-02,  // 1/24
    1.66666666666666666667e-01,  // 1/6
    6755399441055744.0,          // 1.5 * 2^52
    0.0};

// exp(x) for x <= 0 to ~1 ulp: x = (32 k + j) ln2/32 + r, exp(x) = 2^k * 2^(j/32) * e^r with
// |r| <= ln2/64 (degree-6 polynomial) and 2^(j/32) from a 32-entry shared-memory table.
// About half the instructions of the library exp; no special cases are needed on this
// domain (results below 2^-1000 are flushed to zero).
__device__ __forceinline__ double exp_neg(double x, const double *__restrict__ tab)
{
    const double magic = k_expc[8];
    double kd = fma(x, k_expc[0], magic);
    const int ki = __double2loint(kd);
    kd -= magic;
    double r = fma(kd, k_expc[1], x);
    r = fma(kd, k_expc[2], r);
    double p = k_expc[3];
    p = fma(p, r, k_expc[4]);
    p = fma(p, r, k_expc[5]);
    p = fma(p, r, k_expc[6]);
    p = fma(p, r, k_expc[7]);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = p * r; // e^r - 1
    const double t = tab[ki & 31];
    double y = fma(t, p, t);
    const int q = ki >> 5; // floor division: j = ki & 31 >= 0
    y = __hiloint2double(__double2hiint(y) + (q << 20), __double2loint(y));
    return x < -690.0 ? 0.0 : y;
}

// sqrt(a) for a > 0 to ~1 ulp: hardware reciprocal-sqrt seed, one coupled Newton step and a
// final residual correction, without the library's special-case path.
__device__ __forceinline__ double sqrt_pos(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = a * y, h = 0.5 * y;
    const double r = fma(-h, g, 0.5); // one coupled Newton step: ~2^-40
    g = fma(g, r, g);
    h = fma(h, r, h);
    g = fma(fma(-g, g, a), h, g);     // g += (a - g^2) / (2 g): full precision
    return a < 1e-290 ? 0.0 : g;
}

// Nested covariance (krige3d.py:838-875, cova3) from the per-structure squared scaled
// distances h2[s] = (h/a)^2 (h^2 for a power structure).  Inlined at exactly ONE call
// site (the assembly work loop): the transcendental code is large and the instruction
// cache is small.
// one structure's term of cova3 (krige3d.py:856-874) at squared scaled distance h2
__device__ __forceinline__ double cova_term(const K3dParams &P, const KPrep &Q, int s, double h2,
                                            const double *__restrict__ exptab)
{
    const double c = Q.cc[s];
    const int it = Q.it[s];
    if (it == 1) {
        const double hr = sqrt_pos(h2);
        const double v = c * (1.0 - hr * (1.5 - 0.5 * h2));
        return h2 < 1.0 ? v : 0.0;
    } else if (it == 2) {
        return c * exp_neg(-3.0 * sqrt_pos(h2), exptab);
    } else if (it == 3) {
        return c * exp_neg(-3.0 * h2, exptab);
    } else if (it == 4) {
        return P.maxcov - c * pow(sqrt(h2), Q.aa[s]);
    }
    return c * cos(sqrt(h2) * 3.141592653589793);
}

// The same term at N distances at once (the block points of one right-hand-side item; the
// two records a round pairs with a position): one dispatch on the model for the N, whose
// evaluations are independent work.
template <int N>
__device__ __forceinline__ void cova_termN(const K3dParams &P, const KPrep &Q, int s,
                                           const double (&h2)[N], double (&acc)[N],
                                           const double *__restrict__ exptab)
{
    const double c = Q.cc[s];
    const int it = Q.it[s];
    if (it == 1) {
#pragma unroll
        for (int j = 0; j < N; j++) {
            const double hr = sqrt_pos(h2[j]);
            const double v = c * (1.0 - hr * (1.5 - 0.5 * h2[j]));
            acc[j] += h2[j] < 1.0 ? v : 0.0;
        }
    } else if (it == 2) {
#pragma unroll
        for (int j = 0; j < N; j++) acc[j] += c * exp_neg(-3.0 * sqrt_pos(h2[j]), exptab);
    } else if (it == 3) {
#pragma unroll
        for (int j = 0; j < N; j++) acc[j] += c * exp_neg(-3.0 * h2[j], exptab);
    } else {
        for (int j = 0; j < N; j++) acc[j] += cova_term(P, Q, s, h2[j], exptab);
    }
}

// u = rs[s] * (x,y,z) for every structure, written as record idx (4 doubles per structure:
// x, y, z, pad; records `stride` doubles apart)
__device__ __forceinline__ void transform_point(const K3dParams &P, const KPrep &Q, double x,
                                                double y, double z, double *u, int stride, int idx)
{
    double *o = u + idx * stride;
    for (int s = 0; s < P.nst; s++) {
        const double *r = Q.rs[s];
        o[4 * s + 0] = r[0] * x + r[1] * y + r[2] * z;
        o[4 * s + 1] = r[3] * x + r[4] * y + r[5] * z;
        o[4 * s + 2] = r[6] * x + r[7] * y + r[8] * z;
    }
}
// out-of-line copy for the rarely executed call sites
__device__ __noinline__ void transform_point_cold(const K3dParams &P, const KPrep &Q, double x,
                                                  double y, double z, double *u, int stride, int idx)
{
    transform_point(P, Q, x, y, z, u, stride, idx);
}

// 1/d to (better than) double precision without the library's special-case slow path:
// hardware seed (MUFU.RCP64H, relative error e ~ 2^-23) and one cubic step
// y (1 + e + e^2), which leaves e^3.  d == 0 gives a non-finite result; the caller tracks the
// magnitude of every pivot and reports tiny ones as a singular system.
__device__ __forceinline__ double rcp_fast(double d)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}

// predicated shared stores without a branch around them (the owner lanes of a pivot row)
__device__ __forceinline__ void sts_f64x2_if(int pred, uint32_t a, double x, double y)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %0, 0;\n\t@p st.shared.v2.f64 [%1], {%2, %3};\n\t}"
                 ::"r"(pred), "r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sts_f64_if(int pred, uint32_t a, double x)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %0, 0;\n\t@p st.shared.f64 [%1], %2;\n\t}"
                 ::"r"(pred), "r"(a), "d"(x) : "memory");
}

// Register-resident LDL^T of the bordered system.  Entry (r, c), c <= r, lives in lane
// (r & 3) * 8 + (c & 7), register m[r >> 2][c >> 3].  A pivot row is handed to the other
// lanes through a shared row buffer stored PERMUTED: column c at
//     rowpos(c) = (c & 3) * 12 + ((c >> 2) & 1) * 6 + (c >> 3)
// so that everything a lane needs is contiguous and 16-byte aligned: the columns
// lj, lj + 8, ... it multiplies into (its own registers' columns, which is also what an
// owner lane writes) start at rowpos(lj), and the rows li, li + 8, ... / li + 4, li + 12, ...
// it updates are the runs of lanes lj = li and lj = li + 4.  One pivot costs 128-bit shared
// accesses only (c3: 3 stores + 8 loads against 5 + 15 with the plain row), the owner lanes'
// stores are predicated instead of branched around, and the pivots are entered through ONE
// switch on the system size instead of a test per pivot.
template <int AMAX> struct RegSolve {
    static constexpr int BMAX = ((4 * (AMAX - 1) + 3) >> 3) + 1;
    static constexpr int UROW = 48; // doubles per row buffer (4 x 12)
    static_assert(BMAX <= 6, "row buffer groups hold 6 columns");

    static __host__ __device__ constexpr int rowpos(int c) { return (c & 3) * 12 + ((c >> 2) & 1) * 6 + (c >> 3); }

    // rank-1 update of row blocks 0..A-1 with the pivot row at shared address `row`
    template <int A>
    static __device__ __forceinline__ void update(double (&m)[AMAX][BMAX], uint32_t row,
                                                  uint32_t abi, uint32_t abj, double ninv)
    {
        if constexpr (A > 0) {
            constexpr int B = ((4 * (A - 1) + 3) >> 3) + 1;
            double u[B + 1];
#pragma unroll
            for (int b = 0; b < B; b += 2) {
                if (b + 1 < B) {
                    const double2 t = lds_f64x2(row + abj + 8 * b);
                    u[b] = t.x * ninv;
                    u[b + 1] = t.y * ninv;
                } else {
                    u[b] = lds_f64(row + abj + 8 * b) * ninv;
                }
            }
#pragma unroll
            for (int par = 0; par < 2; par++) {      // rows li + 8 e (+ 4): a = 2 e + par
#pragma unroll
                for (int e = 0; 2 * e + par < A; e += 2) {
                    const int a0 = 2 * e + par, a1 = a0 + 2;
                    double w0, w1 = 0.0;
                    if (a1 < A) {
                        const double2 t = lds_f64x2(row + abi + 8 * (6 * par + e));
                        w0 = t.x;
                        w1 = t.y;
                    } else {
                        w0 = lds_f64(row + abi + 8 * (6 * par + e));
                    }
#pragma unroll
                    for (int b = 0; b < B; b++)
                        if (8 * b <= 4 * a0 + 3) m[a0][b] = fma(w0, u[b], m[a0][b]);
                    if (a1 < A) {
#pragma unroll
                        for (int b = 0; b < B; b++)
                            if (8 * b <= 4 * a1 + 3) m[a1][b] = fma(w1, u[b], m[a1][b]);
                    }
                }
            }
        }
    }

    // Pivot M: its row lives in register block AP = M >> 2 of the lanes with li == M & 3.
    // The update covers blocks 0..AP (0..AP-1 when M is the first row of its block: the
    // block's other rows are eliminated already); rows of block AP beyond the pivot then
    // hold don't-care values, which keeps one code path per pivot.
    template <int M>
    static __device__ __forceinline__ void pivot(double (&m)[AMAX][BMAX], uint32_t rowbuf, int li,
                                                 uint32_t abi, uint32_t abj, unsigned &minhi)
    {
        if constexpr (M >= 2 && M < 4 * AMAX) {
            constexpr int AP = M >> 2, q = M & 3;
            constexpr int BW = ((4 * AP + 3) >> 3) + 1; // column blocks of row block AP
            const uint32_t row = rowbuf + (M & 1) * (UROW * 8);
            const int own = li == q;
#pragma unroll
            for (int b = 0; b < BW; b += 2) {
                if (b + 1 < BW) sts_f64x2_if(own, row + abj + 8 * b, m[AP][b], m[AP][b + 1]);
                else sts_f64_if(own, row + abj + 8 * b, m[AP][b]);
            }
            __syncwarp();
            const double d = lds_f64(row + 8 * rowpos(M));
            minhi = min(minhi, (unsigned)__double2hiint(d) & 0x7fffffffu);
            const double ninv = -rcp_fast(d);
            update<(q == 0 ? AP : AP + 1)>(m, row, abi, abj, ninv);
        }
    }

    // Eliminates pivots M = R .. 2; leaves var = entry (1,1) and est = -entry (1,0).
    // minhi: smallest |pivot| seen (high word of the double, which orders like the value).
    static __device__ __forceinline__ void run(const double *Pm, int R, double *rowbuf_g, int lane,
                                               double &est, double &var, unsigned &minhi)
    {
        const int li = lane >> 3, lj = lane & 7;
        const uint32_t abi = (uint32_t)(li * 12) * 8, abj = (uint32_t)((lj & 3) * 12 + (lj >> 2) * 6) * 8;
        const uint32_t rowbuf = smem_u32(rowbuf_g);
        double m[AMAX][BMAX];
#pragma unroll
        for (int a = 0; a < AMAX; a++)
#pragma unroll
            for (int b = 0; b < BMAX; b++)
                if (8 * b <= 4 * a + 3) {
                    int r = li + 4 * a, c = lj + 8 * b;
                    m[a][b] = (c <= r && r <= R) ? Pm[pk(r, c)] : 0.0;
                }
        minhi = 0x7fffffffu;
#define K3D_PIV(M) case M: pivot<M>(m, rowbuf, li, abi, abj, minhi); /* falls through */
        switch (R) { // R <= 4 * AMAX - 1 (amax_for)
        K3D_PIV(43) K3D_PIV(42) K3D_PIV(41) K3D_PIV(40) K3D_PIV(39) K3D_PIV(38) K3D_PIV(37) K3D_PIV(36)
        K3D_PIV(35) K3D_PIV(34) K3D_PIV(33) K3D_PIV(32) K3D_PIV(31) K3D_PIV(30) K3D_PIV(29) K3D_PIV(28)
        K3D_PIV(27) K3D_PIV(26) K3D_PIV(25) K3D_PIV(24) K3D_PIV(23) K3D_PIV(22) K3D_PIV(21) K3D_PIV(20)
        K3D_PIV(19) K3D_PIV(18) K3D_PIV(17) K3D_PIV(16) K3D_PIV(15) K3D_PIV(14) K3D_PIV(13) K3D_PIV(12)
        K3D_PIV(11) K3D_PIV(10) K3D_PIV(9) K3D_PIV(8) K3D_PIV(7) K3D_PIV(6) K3D_PIV(5) K3D_PIV(4)
        K3D_PIV(3) K3D_PIV(2)
        default: break;
        }
#undef K3D_PIV
        var = __shfl_sync(FULL, m[0][0], 9);   // (r,c) = (1,1): lane 1*8+1
        est = -__shfl_sync(FULL, m[0][0], 8);  // (r,c) = (1,0): lane 1*8+0
    }
};

// Generic fallback for systems too large for the register solver: the same elimination
// on a scratch copy of the packed triangle in shared memory (flat rank-1 updates).
__device__ __forceinline__ void smem_solve(double *Pm, const uint16_t *tri, int R, double *wv,
                                           int lane, double &est, double &var, unsigned &minhi)
{
    minhi = 0x7fffffffu;
    for (int M = R; M >= 2; M--) {
        const int T = (M * (M + 1)) >> 1;
        const double d = Pm[T + M];
        minhi = min(minhi, (unsigned)__double2hiint(d) & 0x7fffffffu);
        const double inv = 1.0 / d;
        for (int j = lane; j < M; j += 32) wv[j] = Pm[T + j] * inv;
        __syncwarp();
        const double *u = Pm + T;
        for (int e = lane; e < T; e += 32) {
            int rc = tri[e];
            Pm[e] = fma(-wv[rc >> 8], u[rc & 255], Pm[e]);
        }
        __syncwarp();
    }
    var = Pm[2];
    est = -Pm[1];
}

// sgsim.py:945-964 octant classification (repair D5); arguments are NODE - SAMPLE
// (the sign the distance code already has), so every test is mirrored.
__device__ __forceinline__ int octant_of_neg(double ndx, double ndy, double ndz)
{
    int iq;
    if (ndz <= 0.0) {           // dz >= 0
        iq = 3;
        if (ndx >= 0.0 && ndy < 0.0) iq = 0;   // dx <= 0 and dy > 0
        if (ndx < 0.0 && ndy <= 0.0) iq = 1;   // dx > 0 and dy >= 0
        if (ndx > 0.0 && ndy >= 0.0) iq = 2;   // dx < 0 and dy <= 0
    } else {
        iq = 7;
        if (ndx >= 0.0 && ndy < 0.0) iq = 4;
        if (ndx < 0.0 && ndy <= 0.0) iq = 5;
        if (ndx > 0.0 && ndy >= 0.0) iq = 6;
    }
    return iq;
}

// krige3d.py:359-362: |dx| + |dy| + |dz| < eps (arguments: location - sample)
__device__ __forceinline__ bool coincident(double ndx, double ndy, double ndz)
{
    return __dadd_rn(__dadd_rn(fabs(ndx), fabs(ndy)), fabs(ndz)) < K3D_EPS;
}

// ---------------------------------------------------------------------------
// tile geometry shared by both kernels
// ---------------------------------------------------------------------------
struct Tile {
    int sbx, sby, sbz;  // super block
    int x0, y0, z0;     // first node (z clipped to the slab)
    int tnx, tny, tnz;  // nodes per axis
    int tnodes;         // grid nodes of the tile, or listed locations of the super block
    int zsb0;           // first node plane of the super block (unclipped)
    int first;          // locations mode: index of the tile's first location
};

// Estimation at listed locations instead of grid nodes (cross-validation / jackknife,
// krige3d.py:310-332): device copy of K3dPoints; x == NULL means grid mode.
struct PtSet {
    const double *x, *y, *z, *ext;
    const int32_t *start;
    int exclude;
};

// (the grid covers the tiles tile0 .. tile0 + gridDim.x - 1: the super-block layers that can
// meet the slab)
__device__ __forceinline__ Tile tile_of_block(const K3dParams &P, const K3dInputs &in, int iz0,
                                              int iz1, const PtSet &pts, int tile0)
{
    // A tile is a BRICK of brick[0] x brick[1] x brick[2] super blocks (1 x 1 x 1 unless the
    // caller asks for more: super blocks of a few nodes make CTAs too small); the template
    // runs the caller passes are then those of the whole brick, relative to its first block.
    Tile T;
    const int bx = P.brick[0] > 1 ? P.brick[0] : 1, by = P.brick[1] > 1 ? P.brick[1] : 1,
              bz = P.brick[2] > 1 ? P.brick[2] : 1;
    const int nbx = (P.nxsup + bx - 1) / bx, nby = (P.nysup + by - 1) / by;
    int tile = blockIdx.x + tile0;
    const int tile_id = tile;
    T.sbx = (tile % nbx) * bx;
    tile /= nbx;
    T.sby = (tile % nby) * by;
    T.sbz = (tile / nby) * bz;
    const int sbx1 = min(T.sbx + bx, P.nxsup), sby1 = min(T.sby + by, P.nysup),
              sbz1 = min(T.sbz + bz, P.nzsup);
    T.x0 = in.nodex0[T.sbx];
    T.y0 = in.nodey0[T.sby];
    T.zsb0 = in.nodez0[T.sbz];
    T.first = 0;
    if (pts.x) { // the tile's "nodes" are the locations listed for this super block
        T.first = pts.start[tile_id];
        T.tnodes = pts.start[tile_id + 1] - T.first;
        T.z0 = T.zsb0;
        T.tnx = T.tny = T.tnz = 1;
        return T;
    }
    int z1 = in.nodez0[sbz1];
    T.z0 = T.zsb0 > iz0 ? T.zsb0 : iz0;
    z1 = z1 < iz1 ? z1 : iz1;
    T.tnx = in.nodex0[sbx1] - T.x0;
    T.tny = in.nodey0[sby1] - T.y0;
    T.tnz = z1 - T.z0;
    T.tnodes = (T.tnx > 0 && T.tny > 0 && T.tnz > 0) ? T.tnx * T.tny * T.tnz : 0;
    return T;
}

// location of the tile's t-th node and its index in the output arrays
struct NodeLoc {
    double x, y, z;
    size_t o;
};
__device__ __forceinline__ void snake_node(const Tile &T, int t, int &lx, int &ly, int &lz);
__device__ __forceinline__ NodeLoc node_of_tile(const K3dParams &P, const Tile &T, const PtSet &pts,
                                                int t, size_t slab_off)
{
    NodeLoc L;
    if (pts.x) {
        L.o = (size_t)(T.first + t);
        L.x = pts.x[L.o]; L.y = pts.y[L.o]; L.z = pts.z[L.o];
        return L;
    }
    int lx, ly, lz;
    snake_node(T, t, lx, ly, lz);
    const int ix = T.x0 + lx, iy = T.y0 + ly, iz = T.z0 + lz;
    L.o = (((size_t)iz * P.ny + iy) * P.nx + ix) - slab_off;
    // krige3d.py:307-309: xmn + ix*xsiz, multiply then add, unfused
    L.x = __dadd_rn(P.xmn, __dmul_rn((double)ix, P.xsiz));
    L.y = __dadd_rn(P.ymn, __dmul_rn((double)iy, P.ysiz));
    L.z = __dadd_rn(P.zmn, __dmul_rn((double)iz, P.zsiz));
    return L;
}

// node t of the tile's boustrophedon path: consecutive t are grid neighbours
__device__ __forceinline__ void snake_node(const Tile &T, int t, int &lx, int &ly, int &lz)
{
    lz = t / (T.tnx * T.tny);
    const int rem = t - lz * (T.tnx * T.tny);
    const int lyr = rem / T.tnx, lxr = rem - lyr * T.tnx;
    ly = (lz & 1) ? T.tny - 1 - lyr : lyr;
    lx = ((lz * T.tny + lyr) & 1) ? T.tnx - 1 - lxr : lxr;
}

// ---------------------------------------------------------------------------
// SEARCH kernel: super-block neighbour search (super_block.py:163-317, repairs D3-D5)
//   nbr[(node - slab)*ndmax + i]  neighbours of the node (sorted-sample indices); ordered
//                                 by (distance, index) and -1 padded when `ordered`
//   cnt[node - slab]              number of neighbours, min(nclose, ndmax)
// ---------------------------------------------------------------------------
// One CTA per tile, ONE THREAD PER NODE.  All nodes of a tile search the same super blocks,
// so the candidate samples are gathered once per CTA -- and ordered by their distance from
// the centre of the tile (64 bins in squared distance, a stable counting sort).  A thread
// then streams the candidates in that order (every lane of a warp reads the same candidate:
// shared-memory broadcasts), evaluates the reference's distance exactly as it does
// (sqdist_exact) and keeps, per octant, the `noct` nearest seen so far (or the `ndmax`
// nearest overall) in a small lane-interleaved list with its farthest entry tracked.
// Because the stream arrives nearest-first, the lists converge early, and the thread stops
// as soon as the triangle inequality rules out everything that is left:
//     sqrt(h(centre, s)) - sqrt(h(centre, node)) > sqrt(max over lists of the farthest kept)
// Exactness: which samples are kept is decided on the reference's own distance values with
// ties broken by sample index, the order of the stream only decides when the thread may
// stop; the bound is applied with a 1e-9 safety margin on the safe side.
// PTS: estimation at listed locations (K3dPoints) instead of grid nodes -- a separate
// instantiation, so that the grid kernel carries none of the coincident-sample logic.
// OCT: octant search (noct > 0).
template <bool PTS, bool OCT>
__global__ void __launch_bounds__(256, 2)
k3d_search_kernel(const __grid_constant__ K3dParams P, const K3dInputs in, const int iz0,
                  const int iz1, const int tile0, int32_t *__restrict__ nbr,
                  int32_t *__restrict__ cnt, int32_t *__restrict__ ncand_out, const int ordered,
                  const __grid_constant__ SCfg cfg, const __grid_constant__ PtSet pts)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = cfg.warps * 32;

    double *st4 = (double *)smem; // staged candidates: x, y, z, (sample index | bin << 32)
    int *chunk = (int *)(smem + s_off_chunk(cfg)); // [3][NT]: lo, cnt, off of a batch of runs
    unsigned short *whist = (unsigned short *)(smem + s_off_whist(cfg));
    unsigned short *bstart = (unsigned short *)(smem + s_off_bstart(cfg));
    double *red = (double *)(smem + s_off_red(cfg));
    unsigned char *wreg = smem + s_off_warp(cfg);
    // staging temporaries (the candidates before the sort) overlay the per-warp regions
    int *gtmp = (int *)wreg;
    unsigned short *rtmp = (unsigned short *)(wreg + 4 * cfg.cs);
    unsigned char *btmp = wreg + 6 * cfg.cs;

    __shared__ int s_warpsum[8];

    const PtSet nopts = {nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    const PtSet &geo = PTS ? pts : nopts; // (folds to grid mode at compile time)
    const Tile T = tile_of_block(P, in, iz0, iz1, geo, tile0);
    if (T.tnodes <= 0)
        return;
    const int sbx = T.sbx, sby = T.sby, sbz = T.sbz;
    const unsigned ltm = (1u << lane) - 1u;
    const double *rs = P.rotmat + 9 * P.nst;

    // ---- 1. the tile's candidates: template runs clipped to the grid, ascending index ----
    int total = 0;
    for (int r0 = 0; r0 < in.nruns; r0 += NT) {
        const int r = r0 + tid;
        int lo = 0, n = 0;
        if (r < in.nruns) {
            const int16_t *rn = in.runs + 4 * r;
            const int by = sby + rn[2], bz = sbz + rn[3];
            if (by >= 0 && by < P.nysup && bz >= 0 && bz < P.nzsup) {
                int bx0 = sbx + rn[0], bx1 = sbx + rn[1];
                bx0 = bx0 < 0 ? 0 : bx0;
                bx1 = bx1 >= P.nxsup ? P.nxsup - 1 : bx1;
                if (bx0 <= bx1) {
                    const int b = (bz * P.nysup + by) * P.nxsup;
                    lo = in.sbstart[b + bx0];
                    n = in.sbstart[b + bx1 + 1] - lo;
                }
            }
        }
        int incl = n; // CTA-wide exclusive scan of n
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warpsum[warp] = incl;
        __syncthreads();
        int woff = 0, ctot = 0;
        for (int w = 0; w < cfg.warps; w++) {
            const int v = s_warpsum[w];
            if (w < warp) woff += v;
            ctot += v;
        }
        chunk[tid] = lo;
        chunk[NT + tid] = n;
        chunk[2 * NT + tid] = total + woff + incl - n;
        total += ctot;
        __syncthreads();
        if (total <= cfg.cs) {
            int nr = in.nruns - r0;
            nr = nr < NT ? nr : NT;
            for (int q = warp; q < nr; q += cfg.warps) {
                const int qlo = chunk[q], qcnt = chunk[NT + q], qoff = chunk[2 * NT + q];
                for (int i = lane; i < qcnt; i += 32) {
                    K3D_CHECK(qoff + i >= 0 && qoff + i < cfg.cs, "candidate list index", qoff + i, cfg.cs);
                    gtmp[qoff + i] = qlo + i;
                }
            }
        }
        __syncthreads();
    }
    bool staged = total <= cfg.cs; // else: every thread walks the runs in global memory

    // ---- 2. centre of the tile's nodes and their largest distance from it ------------------
    double cx, cy, cz, ex, ey, ez;
    if (PTS) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int t = tid; t < T.tnodes; t += NT) {
            const size_t q = (size_t)(T.first + t);
            lo[0] = fmin(lo[0], pts.x[q]); hi[0] = fmax(hi[0], pts.x[q]);
            lo[1] = fmin(lo[1], pts.y[q]); hi[1] = fmax(hi[1], pts.y[q]);
            lo[2] = fmin(lo[2], pts.z[q]); hi[2] = fmax(hi[2], pts.z[q]);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                lo[k] = fmin(lo[k], __shfl_xor_sync(FULL, lo[k], d));
                hi[k] = fmax(hi[k], __shfl_xor_sync(FULL, hi[k], d));
            }
            if (lane == 0) { red[warp * 6 + k] = lo[k]; red[warp * 6 + 3 + k] = hi[k]; }
        }
        __syncthreads();
        for (int w = 0; w < cfg.warps; w++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                lo[k] = fmin(lo[k], red[w * 6 + k]);
                hi[k] = fmax(hi[k], red[w * 6 + 3 + k]);
            }
        cx = 0.5 * (lo[0] + hi[0]); cy = 0.5 * (lo[1] + hi[1]); cz = 0.5 * (lo[2] + hi[2]);
        // (the 1e-9 margins below also cover the rounding of these halves)
        ex = 0.5 * (hi[0] - lo[0]); ey = 0.5 * (hi[1] - lo[1]); ez = 0.5 * (hi[2] - lo[2]);
    } else {
        cx = P.xmn + (T.x0 + 0.5 * (T.tnx - 1)) * P.xsiz;
        cy = P.ymn + (T.y0 + 0.5 * (T.tny - 1)) * P.ysiz;
        cz = P.zmn + (T.z0 + 0.5 * (T.tnz - 1)) * P.zsiz;
        ex = 0.5 * (T.tnx - 1) * fabs(P.xsiz); ey = 0.5 * (T.tny - 1) * fabs(P.ysiz);
        ez = 0.5 * (T.tnz - 1) * fabs(P.zsiz);
    }
    double rho2 = sqdist_exact(ex, ey, ez, rs); // corners come in +- pairs: four sign patterns
    rho2 = fmax(rho2, sqdist_exact(ex, ey, -ez, rs));
    rho2 = fmax(rho2, sqdist_exact(ex, -ey, ez, rs));
    rho2 = fmax(rho2, sqdist_exact(ex, -ey, -ez, rs));
    // no node of the tile can reach a sample farther from the centre than this
    const double reach = (sqrt(P.radsqd) + sqrt(rho2)) * (1.0 + 1e-9);
    const double hcmax = reach * reach;
    const double bscale = (double)K3D_SBINS / hcmax;
    const double binw = hcmax / (double)K3D_SBINS * (1.0 - 1e-12); // bin b starts at >= b * binw

    // ---- 3. order the candidates by distance from the centre (stable counting sort) -------
    // Warp w bins a contiguous piece of the candidate list, 32 at a time in list order; a
    // candidate's place is (bin, warp, round, lane), which is its place in a stable sort.
    if (staged) {
        const int per_w = ((total + NT - 1) / NT) * 32;
        const int wbeg = min(total, warp * per_w), wend = min(total, wbeg + per_w);
        for (int b = lane; b < K3D_SBINS; b += 32) whist[warp * K3D_SBINS + b] = 0;
        __syncwarp();
        for (int i0 = wbeg; i0 < wend; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < wend;
            int bin = 255; // not a candidate of any node of the tile
            if (valid) {
                const int g = gtmp[i];
                const double hc = sqdist_exact(cx - in.sx[g], cy - in.sy[g], cz - in.sz[g], rs);
                if (!(hc > hcmax)) bin = min(K3D_SBINS - 1, (int)(hc * bscale));
            }
            const unsigned peers = __match_any_sync(FULL, bin);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (bin != 255 && lane == leader) {
                base = whist[warp * K3D_SBINS + bin];
                whist[warp * K3D_SBINS + bin] = (unsigned short)(base + __popc(peers));
            }
            base = __shfl_sync(FULL, base, leader);
            if (valid) {
                btmp[i] = (unsigned char)bin;
                rtmp[i] = (unsigned short)(base + __popc(peers & ltm));
            }
            __syncwarp();
        }
        __syncthreads();
        for (int b = tid; b < K3D_SBINS; b += NT) { // exclusive prefix over the warps of a bin
            int run = 0;
            for (int w = 0; w < cfg.warps; w++) {
                const int v = whist[w * K3D_SBINS + b];
                whist[w * K3D_SBINS + b] = (unsigned short)run;
                run += v;
            }
            bstart[b] = (unsigned short)run;
        }
        __syncthreads();
        if (warp == 0) { // exclusive prefix over the bins, two per lane
            const int a = bstart[2 * lane], b2 = bstart[2 * lane + 1];
            int incl = a + b2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += v;
            }
            __syncwarp();
            bstart[2 * lane] = (unsigned short)(incl - a - b2);
            bstart[2 * lane + 1] = (unsigned short)(incl - b2);
            if (lane == 31) bstart[K3D_SBINS] = (unsigned short)incl;
        }
        __syncthreads();
        staged = (int)bstart[K3D_SBINS] <= cfg.cst; // what is within reach must fit the stage
        for (int i0 = wbeg; staged && i0 < wend; i0 += 32) {
            const int i = i0 + lane;
            if (i < wend) {
                const int bin = btmp[i];
                if (bin != 255) {
                    const int pos = bstart[bin] + whist[warp * K3D_SBINS + bin] + rtmp[i];
                    K3D_CHECK(pos >= 0 && pos < cfg.cst, "stage place", pos, cfg.cst);
                    const int g = gtmp[i];
                    double2 *rec = reinterpret_cast<double2 *>(st4 + 4 * pos);
                    rec[0] = make_double2(in.sx[g], in.sy[g]);
                    rec[1] = make_double2(in.sz[g], __hiloint2double(bin, g));
                }
            }
        }
        if (staged && tid < 4) { // the stream is read four records at a time: pad with far-away samples
            double2 *rec = reinterpret_cast<double2 *>(st4 + 4 * ((int)bstart[K3D_SBINS] + tid));
            rec[0] = make_double2(1e150, 1e150);
            rec[1] = make_double2(1e150, __hiloint2double(255, 0));
        }
        __syncthreads();
    }
    const int nsorted = staged ? (int)bstart[K3D_SBINS] : 0;

    // ---- 4. one thread per node: stream the candidates, keep the nearest ------------------
    // An entry is a distance (f64) and a 16-bit reference: the candidate's place in the
    // sorted stage (bit 15: the sample at the location itself).  A tile that walks global
    // memory has no stage: its references are sample indices (bit 31: the flag) whose upper
    // halves go where the stage would have been.
    const int cap = cfg.cap, nlist = OCT ? 8 : 1, E = nlist * cap;
    const bool cnt_reg = OCT && cap <= 15;   // octant counts: 4 bits each in one register
    double *keys = (double *)(wreg + warp * cfg.pw_bytes);                        // [E][32] distances
    unsigned short *ref = (unsigned short *)((unsigned char *)keys + E * 256);    // [E][32] references
    unsigned short *refhi = (unsigned short *)smem + warp * (E * 32);             // (no stage) upper halves
    double *tau = (double *)((unsigned char *)keys + E * 320);         // [8][32] bounds (see spw_bytes)
    unsigned short *count = (unsigned short *)((unsigned char *)tau + (OCT ? 0 : 8 * 256)); // [8][32]
    unsigned *outl = (unsigned *)keys; // [ndmax][33] output staging, once the keys are dead
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const size_t slab_off = (size_t)iz0 * P.nx * P.ny;
    const int excl = PTS ? pts.exclude : 0;
    const unsigned FLAG = staged ? 0x8000u : 0x80000000u;
    auto ref_get = [&](int row) -> unsigned {
        unsigned v = ref[row * 32 + lane];
        if (!staged) v |= (unsigned)refhi[row * 32 + lane] << 16;
        return v;
    };
    auto ref_put = [&](int row, unsigned v) {
        K3D_CHECK(row >= 0 && row < E, "list row", row, E);
        K3D_CHECK(!staged || (v & 0x7fffu) < (unsigned)nsorted, "stage reference", v, nsorted);
        ref[row * 32 + lane] = (unsigned short)v;
        if (!staged) refhi[row * 32 + lane] = (unsigned short)(v >> 16);
    };
    // sample index behind a reference (ties on distance go to the lower index)
    auto gidx_of = [&](unsigned v) -> int {
        v &= ~FLAG;
        return staged ? __double2loint(st4[4 * v + 3]) : (int)v;
    };

    for (int base = warp * 32; base < T.tnodes; base += NT) {
        const int t = base + lane;
        const bool active = t < T.tnodes;
        double xloc = 0.0, yloc = 0.0, zloc = 0.0;
        size_t o = 0;
        if (active) {
            if (PTS) {
                o = (size_t)(T.first + t);
                xloc = pts.x[o]; yloc = pts.y[o]; zloc = pts.z[o];
            } else {
                const int lx = t % T.tnx, q = t / T.tnx;
                const int ly = q % T.tny, lz = q / T.tny;
                const int ix = T.x0 + lx, iy = T.y0 + ly, iz = T.z0 + lz;
                o = (((size_t)iz * P.ny + iy) * P.nx + ix) - slab_off;
                // krige3d.py:307-309: xmn + ix*xsiz, multiply then add, unfused
                xloc = __dadd_rn(P.xmn, __dmul_rn((double)ix, P.xsiz));
                yloc = __dadd_rn(P.ymn, __dmul_rn((double)iy, P.ysiz));
                zloc = __dadd_rn(P.zmn, __dmul_rn((double)iz, P.zsiz));
            }
        }
        unsigned cnts = 0; // (cnt_reg) entries kept per octant
        if (OCT && !cnt_reg)
            for (int l = 0; l < 8; l++) count[l * 32 + lane] = 0;
        const double rho = sqrt(sqdist_exact(xloc - cx, yloc - cy, zloc - cz, rs));
        // candidates whose squared distance from the centre exceeds `limit` cannot be kept
        auto limit_of = [&](double tmax) {
            const double s = sqrt(fmin(tmax, P.radsqd)) + rho;
            return s * s * (1.0 + 1e-9);
        };
        double limit = limit_of(INF);
        bool done = !active, changed = false;
        int tested = 0;

        // Octant search: eight lists of `cap` entries kept sorted by (distance, index); a full
        // list's last entry is its bound, a list that is not full accepts anything in the
        // radius.  Entries farther than a newcomer move up one place, the farthest falls off
        // a full list; the stream arrives roughly nearest first, so the newcomer usually lands
        // at or next to the end.
        // Without octants the one list of ndmax entries is kept UNSORTED in `ng` interleaved
        // groups (rows g, g + ng, ...) whose farthest entries are tracked (tau[g], count[g] =
        // its row); the farthest of all (gtau, vrow: registers) is what a newcomer replaces,
        // after which one group and the ng group maxima are looked at again: 4 + 8 loads
        // instead of 32.
        auto oct_count = [&](int l) -> int {
            return cnt_reg ? (int)((cnts >> (4 * l)) & 15u) : (int)count[l * 32 + lane];
        };
        auto oct_bound = [&](int l) -> double {
            return oct_count(l) == cap ? keys[(l * cap + cap - 1) * 32 + lane] : INF;
        };
        double gtau = INF; // (no octants) bound of the list: farthest kept once it is full
        int gn = 0, vrow = 0;
        const int ng = cap >= 32 ? 8 : cap >= 12 ? 4 : cap >= 6 ? 2 : 1;
        auto farthest_of_group = [&](int g) {
            int bi = g;
            double bk = keys[g * 32 + lane];
            for (int r = g + ng; r < cap; r += ng) {
                const double k = keys[r * 32 + lane];
                bool gt = k > bk;
                if (k == bk) gt = gidx_of(ref_get(r)) > gidx_of(ref_get(bi));
                if (gt) { bk = k; bi = r; }
            }
            tau[g * 32 + lane] = bk;
            count[g * 32 + lane] = (unsigned short)bi;
        };
        auto farthest_of_all = [&]() {
            int bg = 0;
            double bk = tau[lane];
            for (int g = 1; g < ng; g++) {
                const double k = tau[g * 32 + lane];
                bool gt = k > bk;
                if (k == bk)
                    gt = gidx_of(ref_get(count[g * 32 + lane])) > gidx_of(ref_get(count[bg * 32 + lane]));
                if (gt) { bk = k; bg = g; }
            }
            gtau = bk;
            vrow = count[bg * 32 + lane];
            changed = true;
        };
        // a candidate: octant l, squared distance h, reference v (its sample index: g)
        auto keep = [&](int l, double h, unsigned v, int g) {
            if constexpr (!OCT) {
                if (h > gtau) return;
                if (gn < cap) {
                    keys[gn * 32 + lane] = h;
                    ref_put(gn, v);
                    if (++gn < cap) return;
                    for (int gr = 0; gr < ng; gr++) farthest_of_group(gr); // full: the bounds start
                } else { // it must beat the farthest kept, (gtau, index) lexicographically
                    if (h == gtau && g > gidx_of(ref_get(vrow))) return;
                    keys[vrow * 32 + lane] = h;
                    ref_put(vrow, v);
                    farthest_of_group(vrow % ng);
                }
                farthest_of_all();
                return;
            }
            int n = oct_count(l), e;
            const int row0 = l * cap;
            if (n == cap) { // full: it must beat the farthest kept, (bound, index) lexicographically
                e = cap - 1;
                const double tl = keys[(row0 + e) * 32 + lane];
                if (h > tl) return;
                if (h == tl && g > gidx_of(ref_get(row0 + e))) return;
                changed = true; // the bound of this list moves
            } else {
                e = n++;
                if (cnt_reg) cnts += 1u << (4 * l);
                else count[l * 32 + lane] = (unsigned short)n;
                if (n == cap) changed = true; // a full list bounds what can still be kept
            }
            while (e > 0) {
                const double k = keys[(row0 + e - 1) * 32 + lane];
                if (k < h) break;
                const unsigned vk = ref_get(row0 + e - 1);
                if (k == h && gidx_of(vk) < g) break;
                keys[(row0 + e) * 32 + lane] = k;
                ref_put(row0 + e, vk);
                e--;
            }
            keys[(row0 + e) * 32 + lane] = h;
            ref_put(row0 + e, v);
        };
        // squared search distance of a sample from the node, its octant and whether it may be
        // used at all; point1 = node, point2 = sample (super_block.py:301-305)
        auto measure = [&](double sx_, double sy_, double sz_, double &h, int &l, bool &self) {
            const double dx = __dsub_rn(xloc, sx_), dy = __dsub_rn(yloc, sy_), dz = __dsub_rn(zloc, sz_);
            h = sqdist_exact(dx, dy, dz, rs);
            l = OCT ? octant_of_neg(dx, dy, dz) : 0;
            self = false;
            // super_block.py:306, and the list's bound as it stands (keep() looks again)
            bool ok = !(h > P.radsqd) && !(h > (OCT ? oct_bound(l) : gtau));
            if (PTS && excl && coincident(dx, dy, dz)) { // krige3d.py:359-363
                if (OCT) self = true;  // takes its place in the octant, dropped afterwards
                else ok = false;       // never a candidate
            }
            return ok;
        };

        if (staged) {
            for (int c0 = 0; c0 < nsorted; c0 += K3D_SCHUNK) {
                if (!done) {
                    if (changed) { // a bound moved: what is the farthest anything can still be?
                        double tmax = OCT ? oct_bound(0) : gtau;
                        if (OCT)
                            for (int l = 1; l < 8; l++) tmax = fmax(tmax, oct_bound(l));
                        limit = limit_of(tmax);
                        changed = false;
                    }
                    done = (double)__double2hiint(st4[4 * c0 + 3]) * binw > limit;
                }
                if (__all_sync(FULL, done)) break;
                if (!done) {
                    // four candidates at a time: their distances are independent work, which
                    // hides the latency of the dependent fp64 chain of each
                    const int c1 = min(c0 + K3D_SCHUNK, nsorted);
                    for (int c = c0; c < c1; c += 4) {
                        double h[4];
                        int l[4], g[4];
                        bool ok[4], self[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const double2 *rec = reinterpret_cast<const double2 *>(st4 + 4 * (c + j));
                            const double2 a = rec[0], b = rec[1];
                            g[j] = __double2loint(b.y);
                            ok[j] = measure(a.x, a.y, b.x, h[j], l[j], self[j]);
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (ok[j]) keep(l[j], h[j], (unsigned)(c + j) | (self[j] ? FLAG : 0u), g[j]);
                    }
                    tested += c1 - c0;
                }
            }
        } else { // tile too dense for the stage: walk the runs in global memory
            for (int r = 0; r < in.nruns; r++) {
                const int16_t *rn = in.runs + 4 * r;
                const int by = sby + rn[2], bz = sbz + rn[3];
                if (by < 0 || by >= P.nysup || bz < 0 || bz >= P.nzsup) continue;
                int bx0 = sbx + rn[0], bx1 = sbx + rn[1];
                bx0 = bx0 < 0 ? 0 : bx0;
                bx1 = bx1 >= P.nxsup ? P.nxsup - 1 : bx1;
                if (bx0 > bx1) continue;
                const int b = (bz * P.nysup + by) * P.nxsup;
                const int lo = in.sbstart[b + bx0], hi = in.sbstart[b + bx1 + 1];
                if (active)
                    for (int i = lo; i < hi; i++) {
                        double h;
                        int l;
                        bool self;
                        if (measure(in.sx[i], in.sy[i], in.sz[i], h, l, self))
                            keep(l, h, (unsigned)i | (self ? FLAG : 0u), i);
                    }
                tested += hi - lo;
            }
        }

        // ---- 5. what the reference keeps (super_block.py:185-224, krige3d.py:356-375) -----
        // The references become sample indices (32 bits, in the rows of `outl` once the keys
        // are dead); kept entries are compacted to rows 0.. of the lane's column first.
        int tot = 0;
        if (active) {
            if (OCT) {
                for (int l = 0; l < 8; l++) {
                    const int n = oct_count(l);
                    for (int k = 0; k < n; k++) {
                        const int row = l * cap + k;
                        const unsigned v = ref_get(row);
                        if (v & FLAG) continue; // the sample at the location itself
                        if (tot != row) {
                            ref_put(tot, v);
                            keys[tot * 32 + lane] = keys[row * 32 + lane];
                        }
                        tot++;
                    }
                }
            } else {
                tot = gn;
            }
        }
        const int na = tot < P.ndmax ? tot : P.ndmax;
        if ((ordered || tot > P.ndmax) && tot > 1) {
            // rows 0..na-1 become the na nearest in (distance, index) order
            for (int j = 0; j < na; j++) {
                int bi = j;
                double bk = keys[j * 32 + lane];
                for (int e = j + 1; e < tot; e++) {
                    const double k = keys[e * 32 + lane];
                    bool lt = k < bk;
                    if (k == bk) lt = gidx_of(ref_get(e)) < gidx_of(ref_get(bi));
                    if (lt) { bk = k; bi = e; }
                }
                if (bi != j) {
                    const unsigned vj = ref_get(j), vb = ref_get(bi);
                    keys[bi * 32 + lane] = keys[j * 32 + lane];
                    ref_put(bi, vj);
                    keys[j * 32 + lane] = bk;
                    ref_put(j, vb);
                }
            }
        }
        int gout[2] = {0, 0}; // (rows are handed over through registers two at a time: the
                              //  staging overlays the keys of every lane of the warp)
        __syncwarp(); // the keys are dead: the output staging overlays them
        for (int e0 = 0; e0 < na; e0 += 2) {
            gout[0] = gidx_of(ref_get(e0));
            if (e0 + 1 < na) gout[1] = gidx_of(ref_get(e0 + 1));
            K3D_CHECK(((e0 + 1) * 33 + lane) * 4 < E * 256, "output staging", e0, E);
            outl[e0 * 33 + lane] = (unsigned)gout[0];
            if (e0 + 1 < na) outl[(e0 + 1) * 33 + lane] = (unsigned)gout[1];
        }
        __syncwarp();
        const int nact = min(32, T.tnodes - base);
        const int pieces = P.ndmax >> 2; // 16-byte pieces of a node's row
        if ((P.ndmax & 3) == 0 && (pieces & (pieces - 1)) == 0 && pieces <= 16) {
            // 32 / pieces nodes per step, one int4 per lane: coalesced 128-bit stores
            const int nps = 32 / pieces, jn = lane / pieces, e0 = 4 * (lane % pieces);
            for (int j0 = 0; j0 < nact; j0 += nps) {
                const int j = j0 + jn, src = j < 32 ? j : 31;
                const unsigned long long oj = __shfl_sync(FULL, (unsigned long long)o, src);
                const int nj = __shfl_sync(FULL, na, src);
                if (j < nact && e0 < (ordered ? P.ndmax : nj)) {
                    int4 v;
                    v.x = e0 < nj ? (int)outl[e0 * 33 + j] : -1;
                    v.y = e0 + 1 < nj ? (int)outl[(e0 + 1) * 33 + j] : -1;
                    v.z = e0 + 2 < nj ? (int)outl[(e0 + 2) * 33 + j] : -1;
                    v.w = e0 + 3 < nj ? (int)outl[(e0 + 3) * 33 + j] : -1;
                    *reinterpret_cast<int4 *>(nbr + oj * P.ndmax + e0) = v;
                }
            }
        } else {
            for (int j = 0; j < nact; j++) { // one coalesced row per node
                const unsigned long long oj = __shfl_sync(FULL, (unsigned long long)o, j);
                const int nj = __shfl_sync(FULL, na, j);
                const int lim = ordered ? P.ndmax : nj;
                for (int e = lane; e < lim; e += 32)
                    nbr[oj * P.ndmax + e] = e < nj ? (int)outl[e * 33 + j] : -1;
            }
        }
        if (active) {
            cnt[o] = na;
            if (ncand_out) ncand_out[o] = tested;
        }
        __syncwarp(); // the next pass re-initialises the lists
    }
}

// ---------------------------------------------------------------------------
// SOLVE kernel: kriging system assembly (krige3d.py:470-583, 748-801, 838-875), solve and
// estimate (krige3d.py:590-614) from the neighbour lists of the search kernel
// ---------------------------------------------------------------------------
// BLOCK: block kriging (ndb > 1), whose right-hand side averages over the block points
// (one round per part of the discretisation, partial sums added up by the owner warp).
// registers per thread: 96 (20 warps per SM) up to 36 rows; the 44-row solver keeps 35 doubles
// of the matrix per lane and spills at 96 (c4: 126 ms against 99.5 ms), so it runs 16 warps of 128
#define K3D_BIG_THREADS(AMAX) ((AMAX) == 11 ? 512 : 640)
template <int AMAX, bool BLOCK>
__global__ void __launch_bounds__(AMAX == 5 ? 256 : K3D_BIG_THREADS(AMAX), AMAX == 5 ? K3D_SMALL_CTAS : 1)
k3d_solve_kernel(const __grid_constant__ K3dParams P, const __grid_constant__ KPrep Q,
                 const K3dInputs in, const K3dOutputs out, const int iz0, const int iz1,
                 const int tile0, const int32_t *__restrict__ nbr, int32_t *__restrict__ cnt,
                 const __grid_constant__ BCfg cfg, const __grid_constant__ PtSet pts)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = cfg.warps * 32;
    const int ndp = cfg.ndp;
    constexpr bool PAIRED = AMAX == 11 || BLOCK; // covariance rounds pair two new records (see below)

    uint16_t *tri = (uint16_t *)smem;
    double *udb = (double *)(smem + b_off_udb(cfg)); // [3*nst][ndb] block points, structure space
    const int pwb = bpw_bytes(cfg), o_warp0 = b_off_warp(cfg), o_pos = bpw_off_pos(cfg),
              o_ut = bpw_off_ut(cfg), o_ints = bpw_off_ints(cfg), o_part = bpw_off_part(cfg);
    unsigned char *wbase = smem + o_warp0 + warp * pwb;
    double *Pm = (double *)wbase;
    // [va][px,py,pz: need_xyz only][ve: has_ext only]
    double *va = (double *)(wbase + o_pos);
    double *px = va + ndp, *py = px + ndp, *pz = py + ndp, *ve = va + (1 + 3 * cfg.need_xyz) * ndp;
    double *ut = (double *)(wbase + o_ut);
    int *pid = (int *)(wbase + o_ints);
    int *newid = pid + ndp;
    unsigned char *newpos = (unsigned char *)(newid + ndp);
    const int ustride = cfg.ustride; // records of ut: neighbours by position, then block points
    double *rows = (double *)(wbase + bpw_off_rows(cfg));
    double *unode = rows; // set-up phase only; the pivot rows are live during the solve
    double *part = (double *)(wbase + o_part);

    __shared__ double s_exptab[32];  // 2^(j/32)
    __shared__ int s_job[20];        // per warp: R | na << 8 | pair rounds << 16 | full << 24
    __shared__ int s_nrounds[2];     // rounds of 32 covariance items published (by parity of tt)
    unsigned short *s_rtab = (unsigned short *)(smem + b_off_rtab(cfg)); // (warp << 8) | round
    __shared__ double s_jloc[20][3]; // per warp: node location

    const Tile T = tile_of_block(P, in, iz0, iz1, pts, tile0);
    if (T.tnodes <= 0)
        return;
    // structure-space origin of the tile: first node of the super block (slab independent)
    const double cx0 = P.xmn + T.x0 * P.xsiz, cy0 = P.ymn + T.y0 * P.ysiz,
                 cz0 = P.zmn + T.zsb0 * P.zsiz;

    if (tid < 32) s_exptab[tid] = exp2((double)tid * 0.03125);
    if (tid < 2) s_nrounds[tid] = 0;
    // ---- triangular index LUT: e -> (row<<8 | col), row >= col --------------
    for (int e = tid; e < cfg.ntri; e += nthreads) {
        int r = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
        while (pk(r + 1, 0) <= e) r++;
        while (pk(r, 0) > e) r--;
        tri[e] = (uint16_t)((r << 8) | (e - pk(r, 0)));
    }
    // ---- block discretisation offsets in every structure's coordinates -------
    for (int j = tid; j < P.ndb; j += nthreads)
        transform_point_cold(P, Q, in.dbxyz[j], in.dbxyz[P.ndb + j], in.dbxyz[2 * P.ndb + j],
                             udb, ustride, j);
    __syncthreads();

    const double *db = in.dbxyz;
    const size_t slab_off = (size_t)iz0 * P.nx * P.ny;
    const unsigned ltm = (1u << lane) - 1u;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    const unsigned singhi = (unsigned)__double2hiint(K3D_SING_TOL * P.maxcov); // |pivot| test, high word
    const int mdt = P.mdt;
    // block kriging: a sample's ndb block-point covariances are split into nsplit work items
    // of a few points so that the pooled rounds stay evenly sized; the partial sums are added
    // up in a fixed order by the owning warp (results do not depend on warp scheduling)
    const int nsplit = cfg.nsplit, napad_max = cfg.napad;
    const bool has_ext = (P.ktype == 2 || P.ktype == 3);

    // ---- each warp walks a contiguous piece of the tile's boustrophedon path ------------
    const int per = (T.tnodes + cfg.warps - 1) / cfg.warps;
    const int t_begin = warp * per, t_end = (t_begin + per < T.tnodes) ? t_begin + per : T.tnodes;
    int nprev = -1; // neighbours of the previous node whose covariances are still in Pm

#ifdef K3D_PHASE_CLOCKS
    long long pc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt_ = clock64();
#endif
    for (int tt = 0; tt < per; tt++) {
        const int t = t_begin + tt;
        const bool active = t < t_end;
        const NodeLoc L = node_of_tile(P, T, pts, active ? t : 0, slab_off);
        const size_t o = L.o;
        const double xloc = L.x, yloc = L.y, zloc = L.z;
        double est = NaN, var = NaN;
        int status = K3D_ST_OK;
        int na = 0;
        double extest = 0.0, resce = 0.0, skmean = P.skmean;

        bool live = active;
        if (has_ext && active) { // krige3d.py:334-344
            extest = pts.x ? pts.ext[o] : in.extgrid[o];
            if (extest < P.tmin || extest >= P.tmax) {
                status = K3D_ST_EXT_TRIMMED;
                live = false;
            }
            resce = P.maxcov / fmax(extest, 0.0001);
            if (P.ktype == 2) skmean = extest;
        }
        if (live) {
            na = cnt[o];
        }

        bool solve = false;
        if (!live) {
            nprev = -1;
            if (active) { // the reference returns before searching: report no neighbours
                if (lane == 0) cnt[o] = 0;
                if (out.nbr_idx)
                    for (int i = lane; i < P.ndmax; i += 32) out.nbr_idx[o * P.ndmax + i] = -1;
            }
        } else if (na < P.ndmin || na == 0) { // krige3d.py:377, repair D15
            status = K3D_ST_NOT_ENOUGH_DATA;
            nprev = -1;
        } else if (na <= mdt && !(na == 1 && mdt == 1 && (P.flags & K3D_FLAG_OK_ONE_SAMPLE))) {
            status = K3D_ST_NOT_ENOUGH_DRIFT; // krige3d.py:383
            nprev = -1;
        } else {
            solve = true;
        }
        const int N = na + mdt, R = N + 1;
        int nnew = na;    // newcomers among the neighbours
        bool full = true; // the whole system is rebuilt (no reuse of the previous node's)
        if (solve) {
            // ---- 1. map the neighbours to matrix positions ---------------------------
            // A neighbour kept from the previous node keeps its row/column, so the
            // sample-sample covariances already in Pm stay valid; newcomers take the
            // positions that were freed.  (newpos[], newid[]) lists what must be filled.
            if (cfg.reuse && nprev == na && na <= 32) {
                int pos = -1;
                const int myid = lane < na ? nbr[o * P.ndmax + lane] : -1;
                // (four positions per shared-memory load; entries beyond na are stale)
                const int4 *pid4 = reinterpret_cast<const int4 *>(pid);
                for (int p = 0; p < na; p += 4) {
                    const int4 v = pid4[p >> 2];
                    if (v.x == myid) pos = p;
                    if (v.y == myid && p + 1 < na) pos = p + 1;
                    if (v.z == myid && p + 2 < na) pos = p + 2;
                    if (v.w == myid && p + 3 < na) pos = p + 3;
                }
                if (lane >= na) pos = -2;
                const unsigned used = __reduce_or_sync(FULL, pos >= 0 ? (1u << pos) : 0u);
                const unsigned unm = __ballot_sync(FULL, pos == -1);
                const int cntnew = __popc(unm);
                if (2 * cntnew <= na) {
                    full = false;
                    nnew = cntnew;
                    if (pos == -1) {
                        unsigned freem = ~used & (na == 32 ? 0xffffffffu : ((1u << na) - 1u));
                        const int r = __popc(unm & ltm);
                        for (int k = 0; k < r; k++) freem &= freem - 1u;
                        K3D_CHECK(r < ndp && freem != 0u, "free position", r, na);
                        newpos[r] = (unsigned char)(__ffs(freem) - 1);
                        newid[r] = myid;
                    }
                }
            }
            if (full) { // positions = list order
                for (int i = lane; i < na; i += 32) {
                    newpos[i] = (unsigned char)i;
                    newid[i] = nbr[o * P.ndmax + i];
                }
            }
            __syncwarp();
            for (int k = lane; k < nnew; k += 32) {
                const int pos = newpos[k], g = newid[k];
                K3D_CHECK(pos >= 0 && pos < na && g >= 0 && g < (int)in.nd, "newcomer position / sample", pos, g);
                pid[pos] = g;
                const double x = in.sx[g], y = in.sy[g], z = in.sz[g];
                if (cfg.need_xyz) { px[pos] = x; py[pos] = y; pz[pos] = z; }
                const double v = in.sv[g], e = has_ext ? in.ssec[g] : 0.0;
                va[pos] = (P.ktype == 0) ? v - P.skmean : (P.ktype == 2 ? v - e : v);
                if (has_ext) ve[pos] = e;
                transform_point(P, Q, x - cx0, y - cy0, z - cz0, ut, ustride, pos);
            }
            // the node and its block points in every structure's coordinates (records
            // ndp..): lane 4*s + k computes component k of structure s
            double un = 0.0;
            if (lane < 4 * P.nst && (lane & 3) < 3) {
                const double *r = Q.rs[lane >> 2] + 3 * (lane & 3);
                un = r[0] * (xloc - cx0) + r[1] * (yloc - cy0) + r[2] * (zloc - cz0);
            }
            if constexpr (BLOCK) {
                if (lane < ustride) unode[lane] = un;
                __syncwarp();
                for (int k = lane; k < ustride * P.ndb; k += 32) {
                    const int j = k / ustride, row = k - j * ustride;
                    ut[(ndp + j) * ustride + row] = unode[row] + udb[j * ustride + row];
                }
            } else {
                if (lane < ustride) ut[ndp * ustride + lane] = un + udb[lane];
            }
            __syncwarp();

            // publish this warp's covariance job for the pooled phase: one job word and one
            // round-table entry per round of 32 work items (lane = a position of the system).
            // Partial update: one round per newcomer, then the right-hand side (lane = sample);
            // or (PAIRED: the instantiations with registers to spare -- the 44-row solver at
            // 128 registers, block kriging -- where it measured 3 % faster; on the 36-row
            // solver at 96 registers it spills and measured 3 % slower) the records that are
            // new to the system -- the newcomers and, for point kriging, the node itself --
            // go TWO per round: a lane evaluates both against its position with one load of
            // the position's record and one dispatch on the model.  Full rebuild: the whole triangle 32 pairs
            // at a time, then the right-hand side (lane = sample).  Block kriging: one
            // right-hand-side round per part of the block discretisation.
            {
                const bool dbl = PAIRED && !full;
                const int nrec = nnew + (BLOCK ? 0 : 1);
                const int npr = full ? (((na * (na - 1)) >> 1) + 31) >> 5 : dbl ? (nrec + 1) >> 1 : nnew;
                const int nr = npr + ((dbl && !BLOCK) ? 0 : nsplit * ((na + 31) >> 5));
                int base = 0;
                if (lane == 0) {
                    s_job[warp] = R | (na << 8) | (npr << 16) | (full ? 1 << 24 : 0) | (dbl ? 1 << 25 : 0) |
                                  ((dbl ? nnew : 0) << 26);
                    if constexpr (BLOCK) {
                        s_jloc[warp][0] = xloc; s_jloc[warp][1] = yloc; s_jloc[warp][2] = zloc;
                    }
                    base = atomicAdd(&s_nrounds[tt & 1], nr);
                }
                base = __shfl_sync(FULL, base, 0);
                K3D_CHECK(base + nr <= cfg.warps * b_rounds_per_warp(cfg), "round table", base + nr,
                          cfg.warps * b_rounds_per_warp(cfg));
                for (int r = lane; r < nr; r += 32) s_rtab[base + r] = (unsigned short)((warp << 8) | r);
            }
        }
        if (tid == 0) s_nrounds[(tt + 1) & 1] = 0; // next node's counter: idle during this phase
        K3D_CLK(0); // match / fill / publish
        K3D_PHASE_BARRIER();
        K3D_CLK(1); // barrier 1

        // ---- 2. covariances of every node in flight, pooled over the CTA ---------------
        // One work loop evaluates every covariance the nodes need (krige3d.py:748-801,
        // 838-875); the rounds of all warps are dealt out evenly, so that no warp waits for
        // another's longer list.  One item = the covariance between two records of the job's
        // structure-space table: two samples (pair rounds), or a sample and the node's block
        // points (right-hand side rounds; record ndp.. = the node).
        {
            const int nrounds = s_nrounds[tt & 1];
            const uint32_t a_rtab = smem_u32(s_rtab), a_w0 = smem_u32(smem + o_warp0),
                           a_job = smem_u32(&s_job[0]), a_tri = smem_u32(tri);
            const int o_newpos = o_ints + 8 * ndp;
            for (int e = warp; e < nrounds; e += cfg.warps) {
                const int ent = lds_u16(a_rtab + 2 * e);
                const int jw = ent >> 8, r = ent & 255;
                const uint32_t jb = a_w0 + jw * pwb;
                const int job = lds_s32(a_job + 4 * jw);
                const int jR = job & 255, jna = (job >> 8) & 255, npr = (job >> 16) & 255;
                if (PAIRED && r < npr && ((job >> 25) & 1)) {
                    // two records (newcomers, then the node) against the lane's position
                    const int jnew = (job >> 26) & 63, nrec = jnew + (BLOCK ? 0 : 1);
                    const int k0 = 2 * r, k1 = k0 + 1 < nrec ? k0 + 1 : k0;
                    int pn0 = ndp, pn1 = ndp; // (the node's record follows the neighbours')
                    if (k0 < jnew) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pn0) : "r"(jb + o_newpos + k0));
                    if (k1 < jnew) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pn1) : "r"(jb + o_newpos + k1));
                    const int q = min(lane, jna - 1); // idle lanes repeat the last position
                    K3D_CHECK(pn0 <= ndp && pn1 <= ndp && q >= 0, "paired records", pn0, pn1);
                    const uint32_t aq = jb + o_ut + q * (ustride << 3), ar0 = jb + o_ut + pn0 * (ustride << 3),
                                   ar1 = jb + o_ut + pn1 * (ustride << 3);
                    double c2[2] = {0.0, 0.0}, h20[2] = {1.0, 1.0};
#pragma unroll 1
                    for (int st = 0; st < P.nst; st++) {
                        const double2 v = lds_f64x2(aq + 32 * st), u0 = lds_f64x2(ar0 + 32 * st),
                                      u1 = lds_f64x2(ar1 + 32 * st);
                        const double vz = lds_f64(aq + 32 * st + 16), u0z = lds_f64(ar0 + 32 * st + 16),
                                     u1z = lds_f64(ar1 + 32 * st + 16);
                        double h2[2];
                        {
                            const double dx = u0.x - v.x, dy = u0.y - v.y, dz = u0z - vz;
                            h2[0] = dx * dx + dy * dy + dz * dz;
                        }
                        {
                            const double dx = u1.x - v.x, dy = u1.y - v.y, dz = u1z - vz;
                            h2[1] = dx * dx + dy * dy + dz * dz;
                        }
                        if (st == 0) { h20[0] = h2[0]; h20[1] = h2[1]; }
                        cova_termN<2>(P, Q, st, h2, c2, s_exptab);
                    }
                    // krige3d.py:853-855; a newcomer pairs into the triangle, the node into the b row
                    const int e0 = pn0 == ndp ? pk(jR - q, 1) : pk(jR - min(pn0, q), jR - max(pn0, q));
                    K3D_CHECK(e0 >= 0 && e0 < cfg.ntri, "packed matrix entry", e0, cfg.ntri);
                    sts_f64(jb + (e0 << 3), h20[0] < Q.eps0s ? P.maxcov : c2[0]);
                    if (k1 != k0) {
                        const int e1 = pn1 == ndp ? pk(jR - q, 1) : pk(jR - min(pn1, q), jR - max(pn1, q));
                        K3D_CHECK(e1 >= 0 && e1 < cfg.ntri, "packed matrix entry", e1, cfg.ntri);
                        sts_f64(jb + (e1 << 3), h20[1] < Q.eps0s ? P.maxcov : c2[1]);
                    }
                    continue;
                }
                // idle lanes repeat the last item of their kind (same value, same place)
                const bool rhs = r >= npr;
                int p1, p2r, prt = 0, nsub = 1;
                if (rhs) {
                    int rr = r - npr;
                    if constexpr (BLOCK) { // part prt of the block points, 32 samples per round
                        const int ngr = (jna + 31) >> 5;
                        prt = ngr == 1 ? rr : rr / ngr;
                        rr -= prt * ngr;
                        const int sub0 = (prt * P.ndb) / nsplit;
                        nsub = ((prt + 1) * P.ndb) / nsplit - sub0;
                        p2r = ndp + sub0;
                    } else {
                        p2r = ndp;
                    }
                    p1 = min((rr << 5) + lane, jna - 1);
                } else if (job >> 24 & 1) { // full rebuild: the triangle, 32 pairs at a time
                    const int w = min((r << 5) + lane, ((jna * (jna - 1)) >> 1) - 1);
                    const int rc = lds_u16(a_tri + 2 * w);
                    p1 = rc & 255;
                    p2r = (rc >> 8) + 1;
                } else { // newcomer r against the lane's position
                    int pn;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pn) : "r"(jb + o_newpos + r));
                    const int q = min(lane, jna - 1);
                    p1 = min(pn, q);
                    p2r = max(pn, q);
                }
                const uint32_t a1 = jb + o_ut + p1 * (ustride << 3);
                uint32_t a2 = jb + o_ut + p2r * (ustride << 3);
                double relx = 0.0, rely = 0.0, relz = 0.0; // sample - node (block kriging)
                if constexpr (BLOCK) {
                    if (rhs) {
                        const double *jpx = (const double *)(smem + o_warp0 + jw * pwb + o_pos) + ndp;
                        relx = jpx[p1] - s_jloc[jw][0];
                        rely = jpx[ndp + p1] - s_jloc[jw][1];
                        relz = jpx[2 * ndp + p1] - s_jloc[jw][2];
                    }
                }
                double acc = 0.0;
                bool done_block = false;
                if constexpr (BLOCK) {
                    if (rhs) { // a sample against its part of the block points, four at a time:
                               // one model dispatch and one load of the sample's record per four
                        done_block = true;
#pragma unroll 1
                        for (int k0 = 0; k0 < nsub; k0 += 4, a2 += 4 * (ustride << 3)) {
                            const int kn = nsub - k0 < 4 ? nsub - k0 : 4;
                            double c4[4] = {0.0, 0.0, 0.0, 0.0}, h20[4] = {1.0, 1.0, 1.0, 1.0};
                            uint32_t b1 = a1, b2 = a2;
#pragma unroll 1
                            for (int st = 0; st < P.nst; st++, b1 += 32, b2 += 32) {
                                const double2 u = lds_f64x2(b1);
                                const double uz = lds_f64(b1 + 16);
                                double h2[4];
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    // (points past the part repeat its last one; not added up)
                                    const uint32_t bj = b2 + (j < kn ? j : kn - 1) * (ustride << 3);
                                    const double2 v = lds_f64x2(bj);
                                    const double vz = lds_f64(bj + 16);
                                    const double dx = u.x - v.x, dy = u.y - v.y, dz = uz - vz;
                                    h2[j] = dx * dx + dy * dy + dz * dz;
                                }
                                if (st == 0) {
#pragma unroll
                                    for (int j = 0; j < 4; j++) h20[j] = h2[j];
                                }
                                cova_termN<4>(P, Q, st, h2, c4, s_exptab);
                            }
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                if (j < kn) {
                                    double c = h20[j] < Q.eps0s ? P.maxcov : c4[j]; // krige3d.py:853-855
                                    // krige3d.py:792-796: a block point on the sample loses the nugget
                                    // (only looked at when the structure-0 distance does not rule it out)
                                    if (h20[j] < Q.epsblk) {
                                        const int sub = p2r - ndp + k0 + j;
                                        const double dx = relx - db[sub], dy = rely - db[P.ndb + sub],
                                                     dz = relz - db[2 * P.ndb + sub];
                                        if (dx * dx + dy * dy + dz * dz < K3D_EPS) c -= P.c0;
                                    }
                                    acc += c;
                                }
                            }
                        }
                    }
                }
                if (!done_block) {
                    double c = 0.0;
                    bool zero = false;
                    uint32_t b1 = a1, b2 = a2;
#pragma unroll 1
                    for (int st = 0; st < P.nst; st++, b1 += 32, b2 += 32) {
                        const double2 u = lds_f64x2(b1), v = lds_f64x2(b2);
                        const double uz = lds_f64(b1 + 16), vz = lds_f64(b2 + 16);
                        const double dx = u.x - v.x, dy = u.y - v.y, dz = uz - vz;
                        const double h2 = dx * dx + dy * dy + dz * dz;
                        if (st == 0) zero = h2 < Q.eps0s;
                        c += cova_term(P, Q, st, h2, s_exptab);
                    }
                    if (zero) c = P.maxcov; // krige3d.py:853-855
                    acc = c;
                }
                K3D_CHECK(jw >= 0 && jw < cfg.warps && p1 >= 0 && p1 < ndp, "round owner / position", jw, p1);
                if (BLOCK && rhs) { // partial sum of the block average, added up by the owner
                    K3D_CHECK(prt * napad_max + p1 < nsplit * napad_max, "partial sum slot", prt, p1);
                    sts_f64(jb + o_part + ((prt * napad_max + p1) << 3), acc);
                } else {
                    K3D_CHECK(pk(jR - p1, rhs ? 1 : jR - p2r) < cfg.ntri && jR - p1 >= (rhs ? 1 : jR - p2r),
                              "packed matrix entry", jR - p1, rhs ? 1 : jR - p2r);
                    sts_f64(jb + (pk(jR - p1, rhs ? 1 : jR - p2r) << 3), acc);
                }
            }
        }
        K3D_CLK(2); // covariance rounds
        K3D_PHASE_BARRIER();
        K3D_CLK(3); // barrier 2
        if (solve) {
            if (na == 1) { // krige3d.py:423-468 (mdt == 0 here, R == 2)
                double cb = Pm[pk(R, 1)];
                if constexpr (BLOCK) {
                    cb = 0.0;
                    for (int q = 0; q < nsplit; q++) cb += part[q * napad_max];
                    cb /= (double)P.ndb;
                }
                const double sw = cb / P.cbb;
                const double v0 = in.sv[pid[0]];
                est = sw * v0 + (1.0 - sw) * skmean;
                var = P.cbb - sw * cb;
                if (mdt == 1) { // kb2d's ordinary-kriging rule, krige2d.py:360-363
                    est = v0;
                    var = P.cbb - 2.0 * cb + P.maxcov;
                }
                nprev = -1;
            } else {
                // Border of the system (krige3d.py:764-767, 493-583 with repair D6).  What
                // only depends on the neighbours' positions -- diagonal, data vector,
                // unbiasedness row, constraint block -- stays in Pm from node to node and is
                // written on a full rebuild (newcomers refresh their data-vector entry);
                // what depends on the node -- drift monomials, external drift, the block
                // right-hand side -- is rebuilt every time.
                const bool drift = mdt > 1;
                if (full) {
                    for (int i = lane; i < na; i += 32) {
                        const int ri = R - i;
                        Pm[pk(ri, ri)] = P.maxcov; // cova3(p, p)
                        Pm[pk(ri, 0)] = va[i];     // v row
                        if (mdt > 0) Pm[pk(ri, R - na)] = P.unbias;
                    }
                    if (lane == 0) {
                        for (int q1 = na; q1 < N; q1++)
                            for (int q2 = q1; q2 < N; q2++)
                                Pm[pk(R - q1, R - q2)] = 0.0;
                        int q = na;
                        if (mdt > 0) {
                            Pm[pk(R - q, 1)] = P.unbias; Pm[pk(R - q, 0)] = 0.0; q++;
                            for (int l = 0; l < K3D_MAXDRIFT; l++)
                                if (P.idrift[l]) { Pm[pk(R - q, 1)] = P.bv[l]; Pm[pk(R - q, 0)] = 0.0; q++; }
                            if (P.ktype == 3) { Pm[pk(R - q, 0)] = 0.0; q++; }
                        }
                        Pm[2] = P.cbb; // (b,b): block covariance
                        Pm[1] = 0.0;   // (v,b): becomes -estimate
                        Pm[0] = 0.0;   // (v,v): unused
                    }
                } else {
                    for (int k = lane; k < nnew; k += 32) {
                        const int pos = newpos[k];
                        Pm[pk(R - pos, 0)] = va[pos];
                    }
                }
                if (BLOCK || P.itrend || drift) {
                    for (int i = lane; i < na; i += 32) {
                        const int ri = R - i;
                        if constexpr (BLOCK) { // b row from the partial sums (krige3d.py:797-798)
                            double cb = 0.0;
                            for (int q = 0; q < nsplit; q++) cb += part[q * napad_max + i];
                            Pm[pk(ri, 1)] = cb / (double)P.ndb;
                        }
                        if (P.itrend) Pm[pk(ri, 1)] = 0.0; // repair D10: kriging the trend
                        if (drift) {
                            const double x = px[i] - xloc, y = py[i] - yloc, z = pz[i] - zloc;
                            int q = na + 1;
#pragma unroll
                            for (int l = 0; l < K3D_MAXDRIFT; l++) {
                                if (P.idrift[l]) {
                                    double f = l == 0 ? x : l == 1 ? y : l == 2 ? z
                                             : l == 3 ? x * x : l == 4 ? y * y : l == 5 ? z * z
                                             : l == 6 ? x * y : l == 7 ? x * z : y * z;
                                    Pm[pk(ri, R - q)] = f * P.resc;
                                    q++;
                                }
                            }
                            if (P.ktype == 3) {
                                Pm[pk(ri, R - q)] = ve[i] * resce;
                                q++;
                            }
                        }
                    }
                }
                if (P.ktype == 3 && lane == 0) Pm[pk(R - (N - 1), 1)] = extest * resce; // last constraint
                __syncwarp();

                K3D_CLK(4); // border
                // ---- 3. LDL^T elimination of the N pivots ---------------------------
                unsigned minhi;
                if constexpr (AMAX > 0) {
                    RegSolve<AMAX>::run(Pm, R, rows, lane, est, var, minhi);
                    nprev = na;
                } else {
                    smem_solve(Pm, tri, R, rows, lane, est, var, minhi); // destroys Pm
                    nprev = -1;
                }
                // krige3d.py:593 "Singular matrix" (repair D16): a pivot below 1e-12 maxcov,
                // or a non-finite result
                if (minhi <= singhi || !isfinite(est) || !isfinite(var)) {
                    status = K3D_ST_SINGULAR;
                    est = var = NaN;
                } else {
                    // var = cbb - b' A^-1 b (krige3d.py:601); est = v' A^-1 b (:603-609)
                    if (P.ktype == 0 || P.ktype == 2) est += skmean; // krige3d.py:611
                }
                __syncwarp();
            }
        }
        K3D_CLK(5); // elimination
        if (lane == 0 && active) {
            out.est[o] = est;
            out.var[o] = var;
            if (out.status) out.status[o] = (uint8_t)status;
        }
        __syncwarp();
    }
#ifdef K3D_PHASE_CLOCKS
    if (lane == 0)
        for (int i = 0; i < 8; i++) atomicAdd(&g_phase_clk[i], (unsigned long long)pc_[i]);
#endif
}

// ---------------------------------------------------------------------------
// template builder (func_pickup, super_block.py:226-265)
// ---------------------------------------------------------------------------
struct Rot9 { double r[9]; };

__global__ void k3d_template_kernel(int nxsup, int nysup, int nzsup, double sxs, double sys,
                                    double szs, const __grid_constant__ Rot9 rot, double radsqd,
                                    uint8_t *mask)
{
    const int nx2 = 2 * nxsup - 1, ny2 = 2 * nysup - 1, nz2 = 2 * nzsup - 1;
    const long long n = (long long)nx2 * ny2 * nz2;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n;
         g += (long long)gridDim.x * blockDim.x) {
        int k = (int)(g % nz2), j = (int)((g / nz2) % ny2), i = (int)(g / ((long long)nz2 * ny2));
        double xo = __dmul_rn((double)(i - (nxsup - 1)), sxs);
        double yo = __dmul_rn((double)(j - (nysup - 1)), sys);
        double zo = __dmul_rn((double)(k - (nzsup - 1)), szs);
        double shortest = 1.7976931348623157e308;
        // (i1-i2)*0.5 takes the values -1, 0, 1: 27 distinct corner differences
        for (int a = -1; a <= 1; a++)
            for (int b = -1; b <= 1; b++)
                for (int c = -1; c <= 1; c++) {
                    double xd = __dadd_rn(__dmul_rn((double)a, sxs), xo);
                    double yd = __dadd_rn(__dmul_rn((double)b, sys), yo);
                    double zd = __dadd_rn(__dmul_rn((double)c, szs), zo);
                    // point1 = (0,0,0), point2 = (xd,yd,zd): d = 0 - xd
                    double h = sqdist_exact(__dsub_rn(0.0, xd), __dsub_rn(0.0, yd),
                                            __dsub_rn(0.0, zd), rot.r);
                    shortest = h < shortest ? h : shortest;
                }
        mask[g] = shortest <= radsqd ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------
// fp64 pipe probe
// ---------------------------------------------------------------------------
__global__ void k3d_dfma_probe(double *sink, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

// The same pipe with the operand pattern of the elimination's rank-1 update: 5 x 5 accumulators,
// every DFMA reading three distinct register pairs (m += w[a] * u[b]).
__global__ void k3d_dfma_rank1_probe(double *sink, int iters)
{
    double m[5][5], w[5], u[5];
#pragma unroll
    for (int a = 0; a < 5; a++) {
        w[a] = 1.0 + (threadIdx.x + a) * 1e-9;
        u[a] = 1e-7 * (1 + a);
#pragma unroll
        for (int b = 0; b < 5; b++) m[a][b] = a + b;
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int a = 0; a < 5; a++)
#pragma unroll
            for (int b = 0; b < 5; b++) m[a][b] = fma(w[a], u[b], m[a][b]);
#pragma unroll
        for (int a = 0; a < 5; a++) { w[a] += 1e-12; u[a] -= 1e-13; } // keeps the loop from folding
    }
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < 5; a++)
#pragma unroll
        for (int b = 0; b < 5; b++) s += m[a][b];
    if (s == 123.456) sink[0] = s;
}

// ---------------------------------------------------------------------------
// self test of the in-line math (exp_neg, sqrt_pos, rcp_fast) against the CUDA library
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{   // splitmix64
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(unsigned long long z) { return (double)(z >> 11) * 0x1.0p-53; }
__device__ __forceinline__ unsigned long long ulp_dist(double a, double b)
{   // both finite and of the same sign on the tested domains
    const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
    return (unsigned long long)(ia > ib ? ia - ib : ib - ia);
}

// out[0..2]: largest distance in units in the last place between exp_neg / sqrt_pos / rcp_fast
// and exp / sqrt / IEEE division; out[3]: number of cut-off violations (exp_neg below -690
// must return 0 or agree to 1e-300 absolute; sqrt_pos below 1e-290 must return 0)
__global__ void k3d_selftest_kernel(long long n, unsigned long long *out)
{
    __shared__ double tab[32];
    if (threadIdx.x < 32) tab[threadIdx.x] = exp2((double)threadIdx.x * 0.03125);
    __syncthreads();
    unsigned long long e0 = 0, e1 = 0, e2 = 0, bad = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long h = mix64((unsigned long long)i);
        // exp_neg on [-690, 0]: uniform, plus a share concentrated near 0 and near the cut-off
        double x = -690.0 * u01(h);
        if ((i & 7) == 1) x = -u01(mix64(h)) * 1e-3;
        if ((i & 7) == 2) x = -689.0 - u01(mix64(h));
        e0 = max(e0, ulp_dist(exp_neg(x, tab), exp(x)));
        const double xc = -690.0 - 60.0 * u01(mix64(h ^ 1));   // beyond the cut-off
        const double yc = exp_neg(xc, tab);
        if (!(yc == 0.0 || fabs(yc - exp(xc)) <= 1e-300)) bad++;
        // sqrt_pos on [1e-290, 1e6], log-uniform (the covariance sees (h/a)^2)
        const double a = exp(log(1e-290) + (log(1e6) - log(1e-290)) * u01(mix64(h ^ 2)));
        e1 = max(e1, ulp_dist(sqrt_pos(a), sqrt(a)));
        if (sqrt_pos(1e-291 * u01(mix64(h ^ 3))) != 0.0) bad++;
        // rcp_fast on +-[1e-150, 1e150], log-uniform (pivots of the elimination)
        double d = exp(log(1e-150) + (log(1e150) - log(1e-150)) * u01(mix64(h ^ 4)));
        if (h & 1) d = -d;
        e2 = max(e2, ulp_dist(rcp_fast(d), 1.0 / d));
    }
    atomicMax(&out[0], e0);
    atomicMax(&out[1], e1);
    atomicMax(&out[2], e2);
    atomicAdd(&out[3], bad);
}

int roundup32(int x) { return (x + 31) & ~31; }

// register-solver size class for a bordered system of `rtot` rows (0 = generic path)
int amax_for(int rtot)
{
    if (rtot <= 20) return 5;
    if (rtot <= 36) return 9;
    if (rtot <= 44) return 11;
    return 0;
}

void make_prep(const K3dParams *p, KPrep *q)
{
    memset(q, 0, sizeof(*q));
    for (int s = 0; s < p->nst; s++) {
        const bool power = p->it[s] == 4;
        const double inva = power ? 1.0 : 1.0 / p->aa[s];
        for (int k = 0; k < 9; k++) q->rs[s][k] = p->rotmat[9 * s + k] * inva;
        q->cc[s] = p->cc[s];
        q->aa[s] = p->aa[s];
        q->it[s] = p->it[s];
    }
    const double a0 = p->it[0] == 4 ? 1.0 : p->aa[0];
    q->eps0s = K3D_EPS / (a0 * a0);
    double fro = 0.0; // h2 = |rs0 d|^2 <= |rs0|_F^2 |d|^2
    for (int k = 0; k < 9; k++) fro += q->rs[0][k] * q->rs[0][k];
    q->epsblk = 4.0 * K3D_EPS * fro;
}

// ---- search kernel plan: warps per CTA and stage capacity ---------------------------------
// `mean_cand` is the expected number of candidate samples per tile, `tile_nodes` the mean
// number of grid nodes per tile.  The per-node selection lists (12 B per kept neighbour, in
// shared memory) set the number of resident warps; the plan takes the CTA width that keeps
// the most warps on an SM without giving a tile more warps than it has nodes for.
int make_scfg(const K3dParams *p, double mean_cand, double tile_nodes, int smem_max, SCfg *c)
{
    memset(c, 0, sizeof(*c));
    c->nlist = p->noct > 0 ? 8 : 1;
    c->cap = p->noct > 0 ? p->noct : p->ndmax;
    if (c->cap > 255) return K3D_ECAPACITY; // list counters are bytes
    c->pw_bytes = spw_bytes(*c);
    // the stage holds what is within reach of the tile (typically 60 % of what the template
    // lists); a tile that needs more walks global memory instead (tested)
    int want = roundup32((int)(1.25 * mean_cand) + 64);
    if (want < 128) want = 128;
    if (want > 4096) want = 4096;
    int floor_cst = roundup32((int)(0.9 * mean_cand) + 32);
    if (floor_cst < 96) floor_cst = 96;
    if (floor_cst > want) floor_cst = want;
    int wneed = (int)ceil(tile_nodes / 32.0); // warps a tile can keep busy at once
    if (wneed < 1) wneed = 1;
    if (wneed > 8) wneed = 8;
    int force_warps = 0; // tuning knob: K3D_SEARCH_WARPS
    if (const char *e = getenv("K3D_SEARCH_WARPS")) force_warps = atoi(e);
    double best = -1.0;
    SCfg pick = *c;
    for (int pass = 0; pass < 2 && best < 0.0; pass++) {
        for (int warps = 1; warps <= 8; warps++) {
            if (force_warps ? warps != force_warps : warps > wneed) continue;
            c->warps = warps;
            c->cst = pass == 0 ? want : floor_cst;
            // (room for the upper halves of the references of a tile that is not staged)
            if (c->cst + 4 < 2 * warps * c->nlist * c->cap) c->cst = roundup32(2 * warps * c->nlist * c->cap);
            // the unsorted candidate list overlays the selection lists: 7 bytes each
            c->cs = roundup32((int)(2.5 * mean_cand) + 64);
            if (c->cs < warps * c->pw_bytes / 7) c->cs = (warps * c->pw_bytes / 7) & ~31;
            if (c->cs > 8192) c->cs = 8192;
            const int bytes = s_total_bytes(*c);
            // the driver reserves 1 KB per CTA; 256 B covers the static shared memory
            if (bytes + 256 > smem_max) continue;
            int ctas = (smem_max + 1024) / (bytes + 256 + 1024);
            if (ctas > 32) ctas = 32;
            if (ctas * warps > 16) ctas = 16 / warps > 0 ? 16 / warps : 1; // ~125 registers per thread
            // resident warps, discounted by the share of a tile's warp slots that idle
            const double wp = ceil(tile_nodes / 32.0), passes = ceil(wp / warps);
            const double score = ctas * warps * ((tile_nodes / 32.0) / (passes * warps)) + 1e-3 * warps;
            if (score > best) {
                best = score;
                pick = *c;
                pick.smem_bytes = bytes;
            }
        }
    }
    if (best < 0.0) return K3D_ECAPACITY;
    *c = pick;
    return K3D_OK;
}

// ---- solve kernel plan: warps per CTA and CTAs per SM -------------------------------------
int make_bcfg(const K3dParams *p, double tile_nodes, int smem_max, int static_smem, BCfg *c)
{
    memset(c, 0, sizeof(*c));
    c->rtot = p->ndmax + p->mdt + 2;
    c->ntri = c->rtot * (c->rtot + 1) / 2;
    c->ndp = (p->ndmax + 3) & ~3;
    c->nst = p->nst;
    c->ndb = p->ndb;
    c->has_ext = (p->ktype == 2 || p->ktype == 3) ? 1 : 0;
    c->need_xyz = (p->mdt > 1 || p->ndb > 1) ? 1 : 0;
    const int amax = amax_for(c->rtot);
    c->urow = amax > 0 ? 48 : c->rtot; // RegSolve<>::UROW (permuted row: 4 groups of 12)
    c->reuse = amax > 0 ? 1 : 0;
    c->napad = (p->ndmax + 31) & ~31;
    c->nsplit = 1;
    if (p->ndb > 1) { // right-hand side work items of <= 4 block points, at most 8 per sample
        int ppi = 4; // tuning knob: K3D_BLOCK_PTS = block points per work item
        if (const char *e = getenv("K3D_BLOCK_PTS")) ppi = atoi(e) > 0 ? atoi(e) : 4;
        c->nsplit = (p->ndb + ppi - 1) / ppi;
        if (c->nsplit > 8) c->nsplit = 8;
    }
    // The kernel is latency bound, so more warps per SM help, but every warp starts its walk
    // with a full-cost node (no covariance reuse), so short walks hurt: score = warps per SM *
    // per / (per + 4) with per = nodes per warp and tile.  Registers cap an SM at 20 warps
    // (102 per thread; 24 warps of 85 for the small-system kernel, whose CTAs hold <= 8 warps).
    const bool small = amax == 5;
    const int big_warps = amax == 11 ? 16 : 20; // K3D_BIG_THREADS
    const int max_cta_warps = small ? 8 : big_warps, max_sm_warps = small ? 24 : big_warps;
    static const int shapes[][2] = {{2, 10}, {1, 20}, {1, 18}, {2, 8}, {1, 16}, {1, 14}, {2, 6},
                                    {1, 12}, {1, 10}, {2, 4}, {1, 8}, {1, 6}, {1, 4},
                                    {3, 8}, {3, 6}, {3, 5}, {4, 4}, {5, 3}, {8, 2}, {6, 2}, {4, 2},
                                    {3, 4}, {12, 2}, {6, 4}, {4, 5}, {5, 4}, {10, 2}, {6, 3}, {7, 2}, {2, 9}, {2, 7}, {3, 7}};
    double best = -1.0;
    BCfg pick = *c;
    int force_ctas = 0, force_warps = 0; // tuning knob: K3D_SOLVE_SHAPE="ctas,warps"
    if (const char *e = getenv("K3D_SOLVE_SHAPE")) sscanf(e, "%d,%d", &force_ctas, &force_warps);
    // Record stride of the structure-space table: 4 doubles per structure, plus 2 when that
    // costs no occupancy (point kriging) -- a stride of 2 mod 4 doubles spreads the 16-byte
    // loads of 8 consecutive records over all banks (the kernel sits at 70 % of the
    // shared-memory wavefront peak; the plain stride makes them 4-way conflicts).
    int force_pad = -1; // tuning knob: K3D_FORCE_PAD=0|1
    if (const char *e = getenv("K3D_FORCE_PAD")) force_pad = atoi(e);
    for (int pad = 2; pad >= 0; pad -= 2) {
    if (force_pad >= 0 ? (pad != 2 * force_pad) : (p->ndb > 1 && pad)) continue;
    c->ustride = 4 * p->nst + pad;
    for (unsigned i = 0; i < sizeof(shapes) / sizeof(shapes[0]); i++) {
        const int ctas = shapes[i][0], warps = shapes[i][1];
        if (warps > max_cta_warps || ctas * warps > max_sm_warps) continue;
        if (force_warps && (ctas != force_ctas || warps != force_warps)) continue;
        // the driver reserves 1 KB per CTA, and the kernel has its static shared memory
        const int budget = (smem_max + 1024) / ctas - 1024 - static_smem;
        c->warps = warps;
        const int bytes = b_off_warp(*c) + warps * bpw_bytes(*c);
        if (bytes > budget) continue;
        // Measured on c3 / c4 / c5: resident warps count most; a walk
        // longer than ~13 nodes adds little; a single CTA per SM loses ~10 % (nothing runs
        // while it waits at its phase barriers); CTAs of 2-3 warps lose 10-15 % to the
        // imbalance of a small covariance pool.
        const double per = tile_nodes / warps;
        const double score = ctas * warps * per / (per + 1.5) * (ctas == 1 ? 0.88 : 1.0) *
                             (warps <= 2 ? 0.85 : warps == 3 ? 0.8 : warps == 4 ? 0.97 : 1.0);
        if (score > best) {
            best = score;
            c->pw_bytes = bpw_bytes(*c);
            c->smem_bytes = bytes;
            pick = *c;
        }
    }
    }
    if (best <= 0.0) return K3D_ECAPACITY;
    *c = pick;
    return K3D_OK;
}

// Stream-ordered scratch for the neighbour lists comes from one memory pool per device
// that keeps what it has been given (release threshold = max): the default pool hands
// freed memory back to the OS at the next synchronisation, which costs more than the
// search kernel itself when 2 GB are re-allocated on every call.
std::mutex g_pool_mu;
cudaMemPool_t g_pools[64] = {};

int scratch_pool(int dev, cudaMemPool_t *pool)
{
    std::mutex &mu = g_pool_mu;
    cudaMemPool_t *pools = g_pools;
    if (dev < 0 || dev >= 64) return fail(K3D_EINVAL, "device index out of range");
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        CUDA_TRY(cudaMemPoolCreate(&pools[dev], &props));
        uint64_t keep = UINT64_MAX;
        CUDA_TRY(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *pool = pools[dev];
    return K3D_OK;
}

// The kernels' dynamic shared memory limit is raised ONCE per kernel and device, to the
// device maximum (a per-call value would race between host threads with different plans).
int raise_smem_limit(const void *func, int dev, int smem_max)
{
    static std::mutex mu;
    static const void *done[64][16] = {};
    if (dev < 0 || dev >= 64) return fail(K3D_EINVAL, "device index out of range");
    std::lock_guard<std::mutex> lock(mu);
    int i = 0;
    for (; i < 16 && done[dev][i]; i++)
        if (done[dev][i] == func) return K3D_OK;
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, func)); // static + dynamic must fit the opt-in maximum
    CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  smem_max - (int)fa.sharedSizeBytes));
    if (i < 16) done[dev][i] = func;
    return K3D_OK;
}

template <int AMAX, bool BLOCK>
int launch_solve2(const K3dParams *p, const KPrep *q, const K3dInputs *in, const K3dOutputs *out,
                  int iz0, int iz1, const int32_t *nbr, int32_t *cnt, const BCfg &c,
                  int tile0, long long ntiles, cudaStream_t st, const PtSet &ps)
{
    int dev = 0, smem_max = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int rc = raise_smem_limit((const void *)k3d_solve_kernel<AMAX, BLOCK>, dev, smem_max);
    if (rc) return rc;
    k3d_solve_kernel<AMAX, BLOCK><<<(unsigned)ntiles, c.warps * 32, c.smem_bytes, st>>>(
        *p, *q, *in, *out, iz0, iz1, tile0, nbr, cnt, c, ps);
    CUDA_TRY(cudaGetLastError());
    return K3D_OK;
}

template <int AMAX>
int launch_solve(const K3dParams *p, const KPrep *q, const K3dInputs *in, const K3dOutputs *out,
                 int iz0, int iz1, const int32_t *nbr, int32_t *cnt, const BCfg &c,
                 int tile0, long long ntiles, cudaStream_t st, const PtSet &ps)
{
    return p->ndb > 1 ? launch_solve2<AMAX, true>(p, q, in, out, iz0, iz1, nbr, cnt, c, tile0, ntiles, st, ps)
                      : launch_solve2<AMAX, false>(p, q, in, out, iz0, iz1, nbr, cnt, c, tile0, ntiles, st, ps);
}

// static shared memory of the solve kernel instantiation a run will use
template <int AMAX>
size_t solve_static_smem_of(bool block)
{
    cudaFuncAttributes fa;
    const void *f = block ? (const void *)k3d_solve_kernel<AMAX, true> : (const void *)k3d_solve_kernel<AMAX, false>;
    return cudaFuncGetAttributes(&fa, f) == cudaSuccess ? fa.sharedSizeBytes : 1536;
}
int solve_static_smem(int amax, bool block)
{
    switch (amax) {
    case 5: return (int)solve_static_smem_of<5>(block);
    case 9: return (int)solve_static_smem_of<9>(block);
    case 11: return (int)solve_static_smem_of<11>(block);
    default: return (int)solve_static_smem_of<0>(block);
    }
}

template <bool PTS, bool OCT>
int launch_search(const K3dParams *p, const K3dInputs *in, int iz0, int iz1, int tile0,
                  long long ntiles, int32_t *nbr, int32_t *cnt, int32_t *ncand, int ordered,
                  const SCfg &c, cudaStream_t st, const PtSet &ps)
{
    int dev = 0, smem_max = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int rc = raise_smem_limit((const void *)k3d_search_kernel<PTS, OCT>, dev, smem_max);
    if (rc) return rc;
    k3d_search_kernel<PTS, OCT><<<(unsigned)ntiles, c.warps * 32, c.smem_bytes, st>>>(
        *p, *in, iz0, iz1, tile0, nbr, cnt, ncand, ordered, c, ps);
    CUDA_TRY(cudaGetLastError());
    return K3D_OK;
}

} // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

const char *k3d_version(void) { return K3D_VERSION_STR; }
const char *k3d_last_error(void) { return g_err; }

int k3d_last_launch(K3dLaunchInfo *info)
{
    if (!info) return fail(K3D_EINVAL, "info is NULL");
#ifdef K3D_PHASE_CLOCKS
    {
        unsigned long long h[8], z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_phase_clk, sizeof(h));
        cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z));
        double tot = 0;
        for (int i = 0; i < 6; i++) tot += (double)h[i];
        static const char *nm[6] = {"fill", "barrier1", "cov", "barrier2", "border", "ldl+out"};
        for (int i = 0; i < 6; i++) fprintf(stderr, "phase %-9s %5.1f %%\n", nm[i], 100.0 * h[i] / (tot > 0 ? tot : 1));
    }
#endif
    if (g_ev_pending) {
        g_ev_pending = false;
        if (cudaEventSynchronize(g_ev[2]) == cudaSuccess) {
            cudaEventElapsedTime(&g_launch.search_ms, g_ev[0], g_ev[1]);
            cudaEventElapsedTime(&g_launch.solve_ms, g_ev[1], g_ev[2]);
        }
    }
    *info = g_launch;
    return K3D_OK;
}

static int validate(const K3dParams *p, const K3dInputs *in, int32_t iz0, int32_t iz1,
                    const K3dOutputs *out, bool points = false)
{
    if (!p || !in || !out) return fail(K3D_EINVAL, "NULL argument");
    if (p->nx < 1 || p->ny < 1 || p->nz < 1) return fail(K3D_EINVAL, "empty grid");
    if (iz0 < 0 || iz1 > p->nz || iz0 > iz1) return fail(K3D_EINVAL, "bad slab [iz0, iz1)");
    if (p->ndmax < 1 || p->ndmax > K3D_MAXNDMAX) return fail(K3D_EINVAL, "ndmax outside 1..K3D_MAXNDMAX");
    if (p->nst < 1 || p->nst > K3D_MAXNST) return fail(K3D_EINVAL, "nst outside 1..K3D_MAXNST");
    if (p->ndb < 1 || p->ndb > K3D_MAXDIS) return fail(K3D_EINVAL, "ndb outside 1..K3D_MAXDIS");
    if (p->ktype < 0 || p->ktype > 3) return fail(K3D_EINVAL, "ikrige outside 0..3");
    if (p->noct < 0 || 8 * p->noct > 512) return fail(K3D_EINVAL, "noct outside 0..64");
    for (int i = 0; i < p->nst; i++) {
        if (p->it[i] < 1 || p->it[i] > 5) return fail(K3D_EINVAL, "variogram type outside 1..5");
        if (p->it[i] != 4 && !(p->aa[i] > 0.0)) return fail(K3D_EINVAL, "variogram range must be positive");
    }
    if (!(p->radsqd > 0.0)) return fail(K3D_EINVAL, "search radius must be positive");
    if (p->mdt < 0 || p->mdt > K3D_MAXDRIFT + 2) return fail(K3D_EINVAL, "bad mdt");
    for (int i = 0; i < 3; i++)
        if (p->brick[i] < 0 || p->brick[i] > 8 || (points && p->brick[i] > 1))
            return fail(K3D_EINVAL, "brick outside 0..8 (and 1 for listed locations)");
    if (!in->sx || !in->sy || !in->sz || !in->sv || !in->sbstart || !in->runs ||
        !in->nodex0 || !in->nodey0 || !in->nodez0 || !in->dbxyz)
        return fail(K3D_EINVAL, "NULL input array");
    if ((p->ktype == 2 || p->ktype == 3) && (!in->ssec || (!in->extgrid && !points)))
        return fail(K3D_EINVAL, "ikrige 2/3 needs ssec and extgrid");
    if (!out->est || !out->var) return fail(K3D_EINVAL, "NULL output array");
    if ((int64_t)p->nxsup * p->nysup * p->nzsup > 0x7fffffff || in->nd > 0x7fffffff)
        return fail(K3D_EINVAL, "too many super blocks / samples for int32 indexing");
    return K3D_OK;
}

// Both entry points: grid nodes of planes [iz0, iz1), or (pts != NULL) the listed locations
static int krige_run(const K3dParams *p, const K3dInputs *in, int32_t iz0, int32_t iz1,
                     const K3dOutputs *out, const K3dPoints *pts, void *stream)
{
    int rc = validate(p, in, iz0, iz1, out, pts != nullptr);
    if (rc) return rc;
    if (iz0 == iz1) return K3D_OK;
    PtSet ps = {nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    if (pts) {
        if (pts->npts < 0 || pts->npts > 0x7fffffff) return fail(K3D_EINVAL, "bad number of locations");
        if (pts->npts == 0) return K3D_OK;
        if (!pts->x || !pts->y || !pts->z || !pts->start) return fail(K3D_EINVAL, "NULL location array");
        if ((p->ktype == 2 || p->ktype == 3) && !pts->ext)
            return fail(K3D_EINVAL, "ikrige 2/3 needs the external value of every location");
        ps.x = pts->x; ps.y = pts->y; ps.z = pts->z; ps.ext = pts->ext; ps.start = pts->start;
        ps.exclude = pts->exclude_coincident ? 1 : 0;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, smem_max = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int bkx = p->brick[0] > 1 ? p->brick[0] : 1, bky = p->brick[1] > 1 ? p->brick[1] : 1,
              bkz = p->brick[2] > 1 ? p->brick[2] : 1;
    const int nbx = (p->nxsup + bkx - 1) / bkx, nby = (p->nysup + bky - 1) / bky,
              nbz = (p->nzsup + bkz - 1) / bkz;
    const double nsb = (double)p->nxsup * p->nysup * p->nzsup, ntl_all = (double)nbx * nby * nbz;
    // (ntmpl counts the super blocks a TILE searches: the brick's template when bricks are used)
    const double mean_cand = in->ntmpl > 0 ? (double)in->nd / nsb * in->ntmpl : 512.0;
    const double tile_nodes = pts ? (double)pts->npts / nsb : (double)p->nx * p->ny * p->nz / ntl_all;
    SCfg sc;
    BCfg bc;
    if (make_scfg(p, mean_cand, tile_nodes, smem_max, &sc) != K3D_OK ||
        make_bcfg(p, tile_nodes, smem_max, solve_static_smem(amax_for(p->ndmax + p->mdt + 2), p->ndb > 1), &bc) != K3D_OK)
        return fail(K3D_ECAPACITY, "ndmax/drift/discretisation exceed the shared memory of an SM");
    // one CTA per super block; CTAs whose super block does not meet the planes exit at once
    const long long ntiles = (long long)nbx * nby * nbz;
    if (ntiles > 0x7fffffffLL) return fail(K3D_EINVAL, "too many tiles");
    KPrep q;
    make_prep(p, &q);

    // neighbour lists travel from the search kernel to the solve kernel through global
    // memory: the caller's arrays when given, else stream-ordered scratch, at most ~4 GB at
    // a time (the slab is then processed in equal runs of z-planes)
    // (locations mode: one "plane" holding every location, one run)
    const size_t plane = pts ? (size_t)pts->npts : (size_t)p->nx * p->ny;
    if (pts) { iz0 = 0; iz1 = 1; }
    const size_t per_plane = plane * ((out->nbr_idx ? 0 : (size_t)p->ndmax * 4) + (out->nbr_cnt ? 0 : 4));
    int zstep = iz1 - iz0;
    if (per_plane > 0) {
        size_t cap = (size_t)4 << 30;
        if (const char *e = getenv("K3D_SCRATCH_CAP_MB")) // tests: force several runs of planes
            if (atoll(e) > 0) cap = (size_t)atoll(e) << 20;
        size_t fit = cap / per_plane;
        if (fit < 1) fit = 1;
        if (!pts && (size_t)zstep > fit) {
            const int nruns = (int)(((size_t)zstep + fit - 1) / fit);
            zstep = (zstep + nruns - 1) / nruns;
        }
    }
    if (g_ev_dev != dev) { // per-thread timing events for k3d_last_launch
        drop_events();
        for (int i = 0; i < 3; i++) CUDA_TRY(cudaEventCreate(&g_ev[i]));
        g_ev_dev = dev;
    }
    int32_t *snbr = nullptr, *scnt = nullptr;
    cudaMemPool_t pool = nullptr;
    if (!out->nbr_idx || !out->nbr_cnt) {
        rc = scratch_pool(dev, &pool);
        if (rc) return rc;
    }
    if (!out->nbr_idx &&
        cudaMallocFromPoolAsync((void **)&snbr, (size_t)zstep * plane * p->ndmax * 4, pool, st) != cudaSuccess) {
        cudaGetLastError();
        return fail(K3D_ENOMEM, "neighbour-list scratch allocation failed");
    }
    if (!out->nbr_cnt &&
        cudaMallocFromPoolAsync((void **)&scnt, (size_t)zstep * plane * 4, pool, st) != cudaSuccess) {
        cudaGetLastError();
        if (snbr) cudaFreeAsync(snbr, st);
        return fail(K3D_ENOMEM, "neighbour-count scratch allocation failed");
    }
    for (int za = iz0; za < iz1 && rc == K3D_OK; za += zstep) {
        const int zb = za + zstep < iz1 ? za + zstep : iz1;
        const size_t shift = (size_t)(za - iz0) * plane; // nodes before this run in the slab
        K3dInputs ci = *in;
        if (ci.extgrid) ci.extgrid += shift;
        K3dOutputs co = *out;
        co.est += shift;
        co.var += shift;
        if (co.nbr_idx) co.nbr_idx += shift * p->ndmax;
        if (co.nbr_cnt) co.nbr_cnt += shift;
        if (co.status) co.status += shift;
        if (co.ncand) co.ncand += shift;
        int32_t *nbr = co.nbr_idx ? co.nbr_idx : snbr;
        int32_t *cnt = co.nbr_cnt ? co.nbr_cnt : scnt;
        // only the super-block layers that can meet planes [za, zb) are launched (one layer
        // of slack on each side: the exact layer of a plane is in the device-side table)
        int tile0 = 0;
        long long ntl = ntiles;
        if (!pts) {
            int l0 = (int)((double)za * p->nzsup / p->nz) - 1;
            int l1 = (int)((double)(zb - 1) * p->nzsup / p->nz) + 1;
            l0 = l0 < 0 ? 0 : l0;
            l1 = l1 >= p->nzsup ? p->nzsup - 1 : l1;
            l0 /= bkz; // (layers of bricks)
            l1 /= bkz;
            tile0 = l0 * nbx * nby;
            ntl = (long long)(l1 - l0 + 1) * nbx * nby;
        }
        const int sz0 = pts ? 0 : za, sz1 = pts ? p->nz : zb, ordered = co.nbr_idx ? 1 : 0;
        cudaEventRecord(g_ev[0], st);
        if (pts)
            rc = p->noct > 0 ? launch_search<true, true>(p, &ci, sz0, sz1, tile0, ntl, nbr, cnt, co.ncand, ordered, sc, st, ps)
                             : launch_search<true, false>(p, &ci, sz0, sz1, tile0, ntl, nbr, cnt, co.ncand, ordered, sc, st, ps);
        else
            rc = p->noct > 0 ? launch_search<false, true>(p, &ci, sz0, sz1, tile0, ntl, nbr, cnt, co.ncand, ordered, sc, st, ps)
                             : launch_search<false, false>(p, &ci, sz0, sz1, tile0, ntl, nbr, cnt, co.ncand, ordered, sc, st, ps);
        if (rc) break;
        cudaEventRecord(g_ev[1], st);
        switch (amax_for(bc.rtot)) {
        case 5: rc = launch_solve<5>(p, &q, &ci, &co, sz0, sz1, nbr, cnt, bc, tile0, ntl, st, ps); break;
        case 9: rc = launch_solve<9>(p, &q, &ci, &co, sz0, sz1, nbr, cnt, bc, tile0, ntl, st, ps); break;
        case 11: rc = launch_solve<11>(p, &q, &ci, &co, sz0, sz1, nbr, cnt, bc, tile0, ntl, st, ps); break;
        default: rc = launch_solve<0>(p, &q, &ci, &co, sz0, sz1, nbr, cnt, bc, tile0, ntl, st, ps); break;
        }
        cudaEventRecord(g_ev[2], st);
        g_ev_pending = true;
        g_launch.launches += 2;
    }
    if (snbr) cudaFreeAsync(snbr, st);
    if (scnt) cudaFreeAsync(scnt, st);
    if (rc) return rc;
    g_launch.grid = (int)ntiles;
    g_launch.block = bc.warps * 32;
    g_launch.smem_bytes = bc.smem_bytes;
    g_launch.warps_per_cta = bc.warps;
    g_launch.stage_capacity = sc.cst;
    g_launch.list_capacity = sc.nlist * sc.cap;
    g_launch.search_block = sc.warps * 32;
    g_launch.search_smem_bytes = sc.smem_bytes;
    return K3D_OK;
}

int k3d_krige_slab(const K3dParams *p, const K3dInputs *in, int32_t iz0, int32_t iz1,
                   const K3dOutputs *out, void *stream)
{
    return krige_run(p, in, iz0, iz1, out, nullptr, stream);
}

int k3d_krige_points(const K3dParams *p, const K3dInputs *in, const K3dPoints *pts,
                     const K3dOutputs *out, void *stream)
{
    if (!pts) return fail(K3D_EINVAL, "NULL argument");
    return krige_run(p, in, 0, p ? p->nz : 0, out, pts, stream);
}

int k3d_build_template(int32_t nxsup, int32_t nysup, int32_t nzsup, double xsizsup,
                       double ysizsup, double zsizsup, const double *rot9_host, double radsqd,
                       uint8_t *mask, void *stream)
{
    if (!rot9_host || !mask || nxsup < 1 || nysup < 1 || nzsup < 1)
        return fail(K3D_EINVAL, "bad template arguments");
    Rot9 r;
    memcpy(r.r, rot9_host, sizeof(r.r));
    long long n = (long long)(2 * nxsup - 1) * (2 * nysup - 1) * (2 * nzsup - 1);
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k3d_template_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(nxsup, nysup, nzsup, xsizsup,
                                                                 ysizsup, zsizsup, r, radsqd, mask);
    CUDA_TRY(cudaGetLastError());
    return K3D_OK;
}

int k3d_measure_fp64_peak(double *tflops, double *ms, void *stream)
{
    if (!tflops) return fail(K3D_EINVAL, "tflops is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    double *sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 1 << 16, blocks = 148 * 8, threads = 256;
    k3d_dfma_probe<<<blocks, threads, 0, st>>>(sink, 1024); // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0, st));
        k3d_dfma_probe<<<blocks, threads, 0, st>>>(sink, iters);
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, e0, e1));
        if (t < best) best = t;
    }
    CUDA_TRY(cudaGetLastError());
    double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms) *ms = best;
    if (getenv("K3D_PROBE_RANK1")) { // diagnostic: the update's operand pattern (35 flop-instr / iteration)
        const int it2 = 1 << 14;
        k3d_dfma_rank1_probe<<<blocks, threads, 0, st>>>(sink, 256);
        float b2 = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0, st);
            k3d_dfma_rank1_probe<<<blocks, threads, 0, st>>>(sink, it2);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float t = 0;
            cudaEventElapsedTime(&t, e0, e1);
            if (t < b2) b2 = t;
        }
        fprintf(stderr, "k3d rank-1 DFMA pattern: %.2f TFLOP/s counting the 25 FMAs (35 fp64 instructions per iteration: %.2f T-instr/s x2)\n",
                2.0 * 25.0 * it2 * blocks * threads / (b2 * 1e-3) / 1e12,
                2.0 * 35.0 * it2 * blocks * threads / (b2 * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return K3D_OK;
}

int k3d_release(void)
{
    drop_events();
    std::lock_guard<std::mutex> lock(g_pool_mu);
    int rc = K3D_OK;
    for (int dev = 0; dev < 64; dev++)
        if (g_pools[dev] && cudaMemPoolTrimTo(g_pools[dev], 0) != cudaSuccess) {
            cudaGetLastError();
            rc = fail(K3D_ECUDA, "cudaMemPoolTrimTo failed");
        }
    return rc;
}

int k3d_selftest_math(int64_t n, uint64_t *max_ulp, uint64_t *violations, void *stream)
{
    if (n < 1 || !max_ulp || !violations) return fail(K3D_EINVAL, "bad self-test arguments");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *d = nullptr, h[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMalloc(&d, sizeof(h)));
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(h), st);
    if (e == cudaSuccess) {
        k3d_selftest_kernel<<<148 * 4, 256, 0, st>>>((long long)n, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return fail(K3D_ECUDA, "math self test: %s", cudaGetErrorString(e));
    max_ulp[0] = h[0]; max_ulp[1] = h[1]; max_ulp[2] = h[2];
    *violations = h[3];
    return K3D_OK;
}

// Host-buffer variant of both entry points: one device arena, copies in, the kernels,
// copies out, one synchronisation.
static int krige_host(const K3dParams *p, const K3dInputs *in, int32_t iz0, int32_t iz1,
                      const K3dOutputs *out, const K3dPoints *pts)
{
    int rc = validate(p, in, iz0, iz1, out, pts != nullptr);
    if (rc) return rc;
    if (pts && (pts->npts < 0 || pts->npts > 0x7fffffff)) return fail(K3D_EINVAL, "bad number of locations");
    if (pts && pts->npts > 0 && (!pts->x || !pts->y || !pts->z || !pts->start))
        return fail(K3D_EINVAL, "NULL location array");
    const size_t nd = (size_t)in->nd;
    const size_t nsb = (size_t)p->nxsup * p->nysup * p->nzsup;
    const size_t nodes = pts ? (size_t)pts->npts : (size_t)(iz1 - iz0) * p->nx * p->ny;
    const bool ext = (p->ktype == 2 || p->ktype == 3);
    if (pts && ext && nodes > 0 && !pts->ext)
        return fail(K3D_EINVAL, "ikrige 2/3 needs the external value of every location");
    cudaStream_t st = nullptr;
    CUDA_TRY(cudaStreamCreate(&st));
    // one device arena for everything
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_sx = take(nd * 8), o_sy = take(nd * 8), o_sz = take(nd * 8), o_sv = take(nd * 8);
    size_t o_sec = take(ext ? nd * 8 : 0), o_sb = take((nsb + 1) * 4), o_runs = take((size_t)in->nruns * 8);
    size_t o_nx = take((p->nxsup + 1) * 4), o_ny = take((p->nysup + 1) * 4), o_nz = take((p->nzsup + 1) * 4);
    size_t o_db = take((size_t)p->ndb * 24), o_ext = take(ext ? nodes * 8 : 0);
    size_t o_px = take(pts ? nodes * 8 : 0), o_py = take(pts ? nodes * 8 : 0), o_pz = take(pts ? nodes * 8 : 0);
    size_t o_ps = take(pts ? (nsb + 1) * 4 : 0);
    size_t o_est = take(nodes * 8), o_var = take(nodes * 8);
    size_t o_ni = take(out->nbr_idx ? nodes * p->ndmax * 4 : 0), o_nc = take(out->nbr_cnt ? nodes * 4 : 0);
    size_t o_st = take(out->status ? nodes : 0), o_ca = take(out->ncand ? nodes * 4 : 0);
    unsigned char *d = nullptr;
    if (cudaMalloc(&d, off ? off : 256) != cudaSuccess) {
        cudaGetLastError(); // (the failed allocation must not surface in the next call)
        cudaStreamDestroy(st);
        return fail(K3D_ENOMEM, "device allocation failed");
    }
    rc = K3D_OK;
#define H2D(dst, src, bytes)                                                              \
    if (rc == K3D_OK && (bytes) > 0 &&                                                    \
        cudaMemcpyAsync(d + (dst), (src), (bytes), cudaMemcpyHostToDevice, st) != cudaSuccess) \
        rc = fail(K3D_ECUDA, "host to device copy failed");
    H2D(o_sx, in->sx, nd * 8) H2D(o_sy, in->sy, nd * 8) H2D(o_sz, in->sz, nd * 8) H2D(o_sv, in->sv, nd * 8)
    if (ext) { H2D(o_sec, in->ssec, nd * 8) H2D(o_ext, pts ? pts->ext : in->extgrid, nodes * 8) }
    H2D(o_sb, in->sbstart, (nsb + 1) * 4) H2D(o_runs, in->runs, (size_t)in->nruns * 8)
    H2D(o_nx, in->nodex0, (p->nxsup + 1) * 4) H2D(o_ny, in->nodey0, (p->nysup + 1) * 4)
    H2D(o_nz, in->nodez0, (p->nzsup + 1) * 4) H2D(o_db, in->dbxyz, (size_t)p->ndb * 24)
    if (pts) {
        H2D(o_px, pts->x, nodes * 8) H2D(o_py, pts->y, nodes * 8) H2D(o_pz, pts->z, nodes * 8)
        H2D(o_ps, pts->start, (nsb + 1) * 4)
    }
#undef H2D
    if (rc == K3D_OK) {
        K3dInputs di = *in;
        di.sx = (double *)(d + o_sx); di.sy = (double *)(d + o_sy); di.sz = (double *)(d + o_sz);
        di.sv = (double *)(d + o_sv); di.ssec = ext ? (double *)(d + o_sec) : nullptr;
        di.sbstart = (int32_t *)(d + o_sb); di.runs = (int16_t *)(d + o_runs);
        di.nodex0 = (int32_t *)(d + o_nx); di.nodey0 = (int32_t *)(d + o_ny); di.nodez0 = (int32_t *)(d + o_nz);
        di.dbxyz = (double *)(d + o_db); di.extgrid = (ext && !pts) ? (double *)(d + o_ext) : nullptr;
        K3dOutputs dout;
        dout.est = (double *)(d + o_est); dout.var = (double *)(d + o_var);
        dout.nbr_idx = out->nbr_idx ? (int32_t *)(d + o_ni) : nullptr;
        dout.nbr_cnt = out->nbr_cnt ? (int32_t *)(d + o_nc) : nullptr;
        dout.status = out->status ? (uint8_t *)(d + o_st) : nullptr;
        dout.ncand = out->ncand ? (int32_t *)(d + o_ca) : nullptr;
        if (pts) {
            K3dPoints dp = *pts;
            dp.x = (double *)(d + o_px); dp.y = (double *)(d + o_py); dp.z = (double *)(d + o_pz);
            dp.ext = ext ? (double *)(d + o_ext) : nullptr;
            dp.start = (int32_t *)(d + o_ps);
            rc = k3d_krige_points(p, &di, &dp, &dout, st);
        } else {
            rc = k3d_krige_slab(p, &di, iz0, iz1, &dout, st);
        }
    }
#define D2H(dst, src, bytes)                                                              \
    if (rc == K3D_OK && (dst) && (bytes) > 0 &&                                           \
        cudaMemcpyAsync((dst), d + (src), (bytes), cudaMemcpyDeviceToHost, st) != cudaSuccess) \
        rc = fail(K3D_ECUDA, "device to host copy failed");
    D2H(out->est, o_est, nodes * 8) D2H(out->var, o_var, nodes * 8)
    D2H(out->nbr_idx, o_ni, nodes * p->ndmax * 4) D2H(out->nbr_cnt, o_nc, nodes * 4)
    D2H(out->status, o_st, nodes) D2H(out->ncand, o_ca, nodes * 4)
#undef D2H
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == K3D_OK)
        rc = fail(K3D_ECUDA, "kernel execution failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
    cudaStreamDestroy(st);
    return rc;
}

int k3d_krige_slab_host(const K3dParams *p, const K3dInputs *in, int32_t iz0, int32_t iz1,
                        const K3dOutputs *out)
{
    return krige_host(p, in, iz0, iz1, out, nullptr);
}

int k3d_krige_points_host(const K3dParams *p, const K3dInputs *in, const K3dPoints *pts,
                          const K3dOutputs *out)
{
    if (!pts) return fail(K3D_EINVAL, "NULL argument");
    return krige_host(p, in, 0, p ? p->nz : 0, out, pts);
}

} // extern "C"
