/*
 * k3d.h -- C ABI of the B200-native krige3d grid-estimation path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * The reference has no FFI of its own for this path (it is a Python class,
 * /root/reference/pygeostatistics/krige3d.py:27); each entry point below names the
 * reference code it replaces.  INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative K3D_E* code; nothing throws;
 *     k3d_last_error() returns a thread-local, human readable message;
 *   - "device" pointers are CUDA device pointers on the current device, owned by the
 *     caller; "host" pointers are ordinary host memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); the
 *     device-pointer entry points are asynchronous on it;
 *   - all floating point data are IEEE fp64; sample indices are int32 positions in
 *     the super-block-sorted sample arrays.
 */
#ifndef K3D_H
#define K3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define K3D_MAXNST 4    /* krige3d.py:184  MAXNST */
#define K3D_MAXDRIFT 9  /* krige3d.py:185  MAXDT  */
#define K3D_MAXNDMAX 64 /* largest ndmax this build accepts */
#define K3D_MAXDIS 128  /* largest nxdis*nydis*nzdis this build accepts */

enum {
    K3D_OK = 0,
    K3D_EINVAL = -1,      /* bad argument / unsupported parameter combination */
    K3D_ECUDA = -2,       /* a CUDA runtime call failed */
    K3D_ENOMEM = -3,      /* host or device allocation failed */
    K3D_ECAPACITY = -4    /* configuration does not fit the kernel's shared memory */
};

/* K3dParams.flags */
enum {
    /* kb2d's rule for a single neighbour under ordinary kriging (krige2d.py:360-363):
       estimate = the datum, variance = cbb - 2 cb + cov(0); krige3d rejects the node
       (krige3d.py:383) */
    K3D_FLAG_OK_ONE_SAMPLE = 1
};

/* per-node status codes written to K3dOutputs.status (krige3d.py:377-389, 590-596) */
enum {
    K3D_ST_OK = 0,
    K3D_ST_NOT_ENOUGH_DATA = 1,   /* na < ndmin, or na == 0            krige3d.py:377 */
    K3D_ST_NOT_ENOUGH_DRIFT = 2,  /* 1 <= na <= mdt                    krige3d.py:383 */
    K3D_ST_SINGULAR = 3,          /* zero pivot ("Singular matrix")    krige3d.py:593 */
    K3D_ST_EXT_TRIMMED = 4        /* external value outside [tmin,tmax) krige3d.py:339 */
};

/*
 * Run parameters: what Krige3d._read_params / _preprocess / _set_rotation /
 * _max_covariance / _rescaling / block_covariance / mdt leave on the object
 * (krige3d.py:57-130, 161-253, 414-421, 616-662, 711-721).
 */
typedef struct K3dParams {
    /* estimation grid, x fastest (krige3d.py:85-93, 302-309) */
    int32_t nx, ny, nz;
    int32_t _pad0;
    double xmn, ymn, zmn;
    double xsiz, ysiz, zsiz;
    /* search (krige3d.py:99-111, 192) */
    int32_t ndmin, ndmax, noct;
    int32_t flags;          /* K3D_FLAG_* (0 for krige3d) */
    double radsqd;
    /* kriging type (krige3d.py:113-119): 0 SK, 1 OK/UK, 2 non-stationary SK, 3 KED */
    int32_t ktype;
    int32_t itrend;
    int32_t idrift[K3D_MAXDRIFT]; /* already forced to 0 for ktype 0/2 (krige3d.py:653) */
    int32_t mdt;                  /* number of constraint rows (krige3d.py:636-662) */
    double skmean;
    double tmin, tmax;
    /* variogram (krige3d.py:121-130) */
    int32_t nst;
    int32_t it[K3D_MAXNST];
    int32_t _pad2;
    double c0;
    double cc[K3D_MAXNST];
    double aa[K3D_MAXNST];
    double rotmat[(K3D_MAXNST + 1) * 9]; /* row major; [nst] is the search matrix */
    double maxcov, resc, unbias, cbb;    /* krige3d.py:270, 414-421, 616-634, 711-721 */
    double bv[K3D_MAXDRIFT];             /* drift means x resc (krige3d.py:273-282) */
    /* super-block network (super_block.py:98-108) */
    int32_t nxsup, nysup, nzsup;
    /* block discretisation: ndb joint points (krige3d.py:692-707) */
    int32_t ndb;
    /* work decomposition, not semantics: a CTA works on a brick of brick[0] x brick[1] x
       brick[2] super blocks (0 or 1 = one super block).  K3dInputs.runs / ntmpl must then
       describe the union of the templates of the brick's blocks, relative to its first
       block.  Useful when super blocks hold only a few grid nodes; results do not change
       (every sample within the search radius is in the template of its node's block). */
    int32_t brick[3];
    int32_t _pad3;
} K3dParams;

/*
 * Device-resident inputs.  Replaces the state SuperBlockSearcher.setup()/pickup()
 * build (super_block.py:87-161) plus the sorted record array (krige3d.py:690).
 */
typedef struct K3dInputs {
    int64_t nd;            /* number of samples */
    const double *sx, *sy, *sz; /* sample coordinates, sorted by super block */
    const double *sv;      /* primary variable, same order */
    const double *ssec;    /* secondary variable (ktype 2/3) or NULL */
    const int32_t *sbstart;/* [nxsup*nysup*nzsup + 1] first sample of each super block
                              (exclusive prefix of super_block.py:138's nisb) */
    /* search template as runs of x-contiguous super-block offsets, ascending in
       (dz, dy, dx): run r covers offsets (dx0[r]..dx1[r], dy[r], dz[r]).  The union
       of the runs is exactly the reference template (super_block.py:226-265). */
    int32_t nruns;
    int32_t ntmpl;         /* number of template offsets the runs cover (sizing hint, 0 = unknown) */
    const int16_t *runs;   /* [nruns*4] = dx0, dx1, dy, dz */
    /* node -> super block index per axis (super_block.py:319-346) and its inverse:
       nodes lutx0[b] .. lutx0[b+1]-1 fall in super-block column b */
    const int32_t *nodex0, *nodey0, *nodez0; /* [nxsup+1], [nysup+1], [nzsup+1] */
    const double *dbxyz;   /* [3*ndb] block discretisation offsets x.., y.., z.. */
    const double *extgrid; /* external drift value per node of the slab, index
                              (node - iz0*nx*ny), or NULL */
} K3dInputs;

/* Device-resident outputs for the slab iz0..iz1-1; element (node - iz0*nx*ny). */
typedef struct K3dOutputs {
    double *est;       /* estimate,           NaN = unestimated (krige3d.py:299)   */
    double *var;       /* kriging variance    (krige3d.py:300)                     */
    int32_t *nbr_idx;  /* [nodes*ndmax] neighbours used, nearest first, or NULL    */
    int32_t *nbr_cnt;  /* [nodes] number of neighbours used, or NULL               */
    uint8_t *status;   /* [nodes] K3D_ST_* or NULL                                 */
    int32_t *ncand;    /* [nodes] samples distance-tested by the search, or NULL   */
} K3dOutputs;

/* Library / build identification. */
const char *k3d_version(void);
/* Message for the last failing call on this thread. */
const char *k3d_last_error(void);

/*
 * Hot path.  Kriges every node of z-planes [iz0, iz1) on the current device:
 * super-block neighbour search (super_block.py:163-317), kriging-system assembly
 * (krige3d.py:470-583, 748-801, 838-875), solve and estimate (krige3d.py:590-614),
 * i.e. the body of the node loop krige3d.py:302-403.  Asynchronous on `stream`.
 */
int k3d_krige_slab(const K3dParams *p, const K3dInputs *in, int32_t iz0, int32_t iz1,
                   const K3dOutputs *out, void *stream);

/*
 * Estimation at listed locations instead of grid nodes: cross-validation (`option` 1,
 * the locations are the data themselves) and jackknife (`option` 2, the records of
 * `jackfl`), krige3d.py:310-332 + 359-363 (the reference reads no file there and crashes;
 * semantics per its own acceptance loop and gslib_help/kt3d_help.md:20-27).
 * The locations are grouped by the super block they fall in (super_block.py:319-346),
 * stable within a block; results are indexed by position in these arrays.
 */
typedef struct K3dPoints {
    int64_t npts;
    const double *x, *y, *z;  /* device: locations, sorted by super block */
    const double *ext;        /* device: external-drift value at each location (ktype 2/3) or NULL */
    const int32_t *start;     /* device: [nxsup*nysup*nzsup + 1] first location of each super block */
    int32_t exclude_coincident; /* 1: samples with |dx|+|dy|+|dz| < eps of the location are
                                   not used (krige3d.py:359-363, option != 0) */
    int32_t _pad;
} K3dPoints;

/* Same path as k3d_krige_slab for the locations of `pts`; K3dInputs.extgrid is not read,
   every K3dOutputs array has pts->npts (x ndmax) elements.  Asynchronous on `stream`. */
int k3d_krige_points(const K3dParams *p, const K3dInputs *in, const K3dPoints *pts,
                     const K3dOutputs *out, void *stream);

/*
 * Builds the search template mask on the device: mask[((i*ny2)+j)*nz2+k] = 1 when
 * offset (i-(nxsup-1), j-(nysup-1), k-(nzsup-1)) must be searched; n?2 = 2*n?sup-1.
 * Replaces func_pickup (super_block.py:226-265).  `mask` is a device pointer.
 */
int k3d_build_template(int32_t nxsup, int32_t nysup, int32_t nzsup, double xsizsup,
                       double ysizsup, double zsizsup, const double *rot9_host,
                       double radsqd, uint8_t *mask, void *stream);

/*
 * Same hot path with HOST buffers (the call a ctypes/cffi binding from the reference
 * would make): copies the inputs to the device, kriges planes [iz0, iz1), copies
 * est/var (and the optional outputs) back, synchronises.  All pointers in `in`/`out`
 * are host pointers here.
 */
int k3d_krige_slab_host(const K3dParams *p, const K3dInputs *in, int32_t iz0, int32_t iz1,
                        const K3dOutputs *out);

/* k3d_krige_points with HOST buffers (every pointer of `in`, `pts`, `out` is a host pointer). */
int k3d_krige_points_host(const K3dParams *p, const K3dInputs *in, const K3dPoints *pts,
                          const K3dOutputs *out);

/* Kernel launch facts for the last k3d_krige_slab on this thread (bench / tests).
   `launches` counts kernels launched so far: a slab takes one search and one solve kernel
   per run of z-planes (one run unless the neighbour scratch would exceed 4 GB). */
typedef struct K3dLaunchInfo {
    int32_t grid, block, smem_bytes, warps_per_cta; /* solve kernel */
    int32_t stage_capacity, list_capacity, launches;
    int32_t search_block, search_smem_bytes;        /* search kernel */
    int32_t _pad;
    float search_ms, solve_ms; /* device time of the two kernels of the last run of planes
                                  (CUDA events on the caller's stream; k3d_last_launch waits
                                  for them) */
} K3dLaunchInfo;
int k3d_last_launch(K3dLaunchInfo *info);

/* Gives back what the library holds between calls on behalf of the calling thread and
   process: the per-thread timing events of k3d_last_launch and the unused memory of the
   per-device scratch pools the neighbour lists travel through (up to ~4 GB per device,
   kept between calls because re-allocating it costs more than the search kernel).
   Safe to call at any time between runs; the next call re-creates what it needs. */
int k3d_release(void);

/* Self test of the in-line device math the covariance and the elimination use in place of
   the CUDA library (exp for arguments <= 0, sqrt, reciprocal; the reference calls libm
   through Numba, krige3d.py:838-875): evaluates `n` pseudo-random points per function on
   the device and returns the largest distance, in units in the last place, from exp(),
   sqrt() and IEEE division in max_ulp[0..2], and the number of cut-off violations
   (exp below -690 -> 0, sqrt below 1e-290 -> 0) in *violations.  Synchronises `stream`. */
int k3d_selftest_math(int64_t n, uint64_t *max_ulp, uint64_t *violations, void *stream);

/* fp64 pipe peak probe: runs a dependent-chain-free DFMA loop on all SMs and returns
   the measured TFLOP/s (2 flops per DFMA).  Used only by bench.py for the roofline. */
int k3d_measure_fp64_peak(double *tflops, double *ms, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* K3D_H */
