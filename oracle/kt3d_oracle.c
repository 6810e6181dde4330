/*
 * kt3d_oracle.c -- CPU restatement of the reference's krige3d grid-estimation path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call it.  Nothing under pygeostatistics_b200/ links or imports it.
 *
 * It restates, in plain scalar C (compile with -ffp-contract=off so that, like the
 * reference's Numba/LLVM build, no multiply-add is ever fused), what
 * /root/reference/pygeostatistics/krige3d.py and super_block.py compute per grid
 * node.  Each function cites the reference lines it follows.  Where the reference
 * as written cannot run the minimal
 * repair listed there is applied and marked "repair Dn".
 *
 * Pinning: oracle/ is checked by tests/test_oracle_golden.py against the vectors in
 * tests/golden/ that tests/golden/make_golden.py produced by running the UNMODIFIED
 * reference in the build container (config 1 as shipped and with ikrige=1, the
 * reference's jitted kernels on seeded inputs, and five small synthetic full runs).
 * The reference's own test-suite holds no vector for this path.
 *
 * The dense solve: the reference calls scipy.linalg.solve (krige3d.py:592), i.e.
 * LAPACK *gesv = LU with partial (row) pivoting.  SciPy/LAPACK are third-party,
 * unpinned by the reference (setup.py:32-37); oracle_solve() restates the published
 * unblocked algorithm (dgetf2 + forward/back substitution).  LAPACK returns info>0 (SciPy
 * raises LinAlgError, krige3d.py:593) only for an exactly zero pivot; the oracle treats a
 * pivot <= ORC_SING_TOL x the largest matrix entry as zero (repair D16, see oracle_solve).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXNST 4
#define ORC_MAXDRIFT 9
#define ORC_SING_TOL 1e-12
#define ORC_EPS (7. / 3 - 4. / 3 - 1) /* krige3d.py:795,853 : 2.220446049250313e-16 */

typedef struct {
    /* grid, krige3d.py:85-93 */
    int32_t nx, ny, nz;
    double xmn, ymn, zmn, xsiz, ysiz, zsiz;
    /* search, krige3d.py:99-111 */
    int32_t ndmin, ndmax, noct;
    double radsqd;
    /* kriging type, krige3d.py:113-119 */
    int32_t ktype;          /* 0 SK, 1 OK/UK, 2 non-stationary SK, 3 KED */
    int32_t idrift[ORC_MAXDRIFT];
    int32_t itrend;
    double skmean;
    double tmin, tmax;
    /* variogram, krige3d.py:121-130 */
    int32_t nst;
    int32_t it[ORC_MAXNST];
    double c0;
    double cc[ORC_MAXNST];
    double aa[ORC_MAXNST];
    double rotmat[(ORC_MAXNST + 1) * 9]; /* [nst] = search matrix, krige3d.py:211-253 */
    /* derived scalars, krige3d.py:270,414-421,711-721 */
    double maxcov, resc, unbias, cbb;
    double bv[ORC_MAXDRIFT];
    /* super-block network, super_block.py:98-108 */
    int32_t nxsup, nysup, nzsup;
    /* block discretisation (joint list of ndb points, repair D9) */
    int32_t ndb;
    /* bit 0: kb2d's one-sample rule under ordinary kriging (krige2d.py:360-363) */
    int32_t flags;
} OrcParams;

/* super_block.py:349-378 and krige3d.py:803-836 (identical bodies) */
static double orc_sqdist(double x1, double y1, double z1, double x2, double y2,
                         double z2, const double *r)
{
    double dx = x1 - x2, dy = y1 - y2, dz = z1 - z2;
    double s = 0.0;
    for (int i = 0; i < 3; i++) {
        double cont = r[3 * i + 0] * dx + r[3 * i + 1] * dy + r[3 * i + 2] * dz;
        s += cont * cont;
    }
    return s;
}

/* krige3d.py:838-875 */
double orc_cova3(const OrcParams *p, double x1, double y1, double z1, double x2,
                 double y2, double z2)
{
    double hsqd = orc_sqdist(x1, y1, z1, x2, y2, z2, p->rotmat);
    if (hsqd < ORC_EPS)
        return p->maxcov;
    double cova = 0.0;
    for (int ist = 0; ist < p->nst; ist++) {
        if (ist != 0)
            hsqd = orc_sqdist(x1, y1, z1, x2, y2, z2, p->rotmat + 9 * ist);
        double h = sqrt(hsqd);
        double a = p->aa[ist], c = p->cc[ist];
        switch (p->it[ist]) {
        case 1: {
            double hr = h / a;
            if (hr < 1)
                cova += c * (1 - hr * (1.5 - 0.5 * hr * hr));
            break;
        }
        case 2:
            cova += c * exp(-3.0 * h / a);
            break;
        case 3:
            cova += c * exp(-3.0 * (h / a) * (h / a));
            break;
        case 4:
            cova += p->maxcov - c * pow(h, a);
            break;
        case 5:
            cova += c * cos(h / a * M_PI);
            break;
        default:
            break;
        }
    }
    return cova;
}

/* super_block.py:226-265 : template of super-block offsets, 64 corner pairs each */
int orc_pickup(int nxsup, int nysup, int nzsup, double xsizsup, double ysizsup,
               double zsizsup, const double *rot, double radsqd, int32_t *ox,
               int32_t *oy, int32_t *oz)
{
    int n = 0;
    static const int pm[2] = {-1, 1};
    for (int i = -(nxsup - 1); i < nxsup; i++)
        for (int j = -(nysup - 1); j < nysup; j++)
            for (int k = -(nzsup - 1); k < nzsup; k++) {
                double xo = i * xsizsup, yo = j * ysizsup, zo = k * zsizsup;
                double shortest = 1.7976931348623157e308;
                for (int a1 = 0; a1 < 2; a1++)
                for (int b1 = 0; b1 < 2; b1++)
                for (int c1 = 0; c1 < 2; c1++)
                for (int a2 = 0; a2 < 2; a2++)
                for (int b2 = 0; b2 < 2; b2++)
                for (int c2 = 0; c2 < 2; c2++) {
                    double xdis = (pm[a1] - pm[a2]) * 0.5 * xsizsup + xo;
                    double ydis = (pm[b1] - pm[b2]) * 0.5 * ysizsup + yo;
                    double zdis = (pm[c1] - pm[c2]) * 0.5 * zsizsup + zo;
                    double h = orc_sqdist(0, 0, 0, xdis, ydis, zdis, rot);
                    if (h < shortest)
                        shortest = h;
                }
                if (shortest <= radsqd) {
                    if (ox) {
                        ox[n] = i;
                        oy[n] = j;
                        oz[n] = k;
                    }
                    n++;
                }
            }
    return n;
}

typedef struct {
    double h;
    int64_t i;
} OrcCand;

static int orc_cmp(const void *a, const void *b)
{
    const OrcCand *p = (const OrcCand *)a, *q = (const OrcCand *)b;
    if (p->h < q->h) return -1;
    if (p->h > q->h) return 1;
    return (p->i > q->i) - (p->i < q->i); /* repair D4: ties -> lower sorted index */
}

/*
 * super_block.py:267-317 with repairs D3 (cursor advances on every sample) and D4
 * (sort key is the squared distance), then the octant filter super_block.py:185-224
 * with repair D5 (classification per sgsim.py:945-964).  Returns nclose; cand[] holds
 * the kept candidates in ascending (hsqd, index) order.
 */
static int orc_search(const OrcParams *p, double xloc, double yloc, double zloc,
                      int ix, int iy, int iz, int ntmpl, const int32_t *tx,
                      const int32_t *ty, const int32_t *tz, const int64_t *nisb,
                      const double *sx, const double *sy, const double *sz,
                      OrcCand *cand)
{
    const double *rot = p->rotmat + 9 * p->nst;
    int nclose = 0;
    for (int t = 0; t < ntmpl; t++) {
        int bx = ix + tx[t], by = iy + ty[t], bz = iz + tz[t];
        if (bx < 0 || bx >= p->nxsup || by < 0 || by >= p->nysup || bz < 0 ||
            bz >= p->nzsup)
            continue;
        int64_t ii = bx + (int64_t)by * p->nxsup + (int64_t)bz * p->nxsup * p->nysup;
        int64_t lo = ii == 0 ? 0 : nisb[ii - 1], hi = nisb[ii];
        for (int64_t i = lo; i < hi; i++) {
            double h = orc_sqdist(xloc, yloc, zloc, sx[i], sy[i], sz[i], rot);
            if (h > p->radsqd)
                continue;
            cand[nclose].h = h;
            cand[nclose].i = i;
            nclose++;
        }
    }
    qsort(cand, nclose, sizeof(OrcCand), orc_cmp);
    if (p->noct <= 0)
        return nclose;
    int inoct[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int na = 0, nt = 8 * p->noct;
    for (int j = 0; j < nclose && na < nt; j++) {
        int64_t i = cand[j].i;
        double dx = sx[i] - xloc, dy = sy[i] - yloc, dz = sz[i] - zloc;
        int iq;
        if (dz >= 0) {
            iq = 3;
            if (dx <= 0 && dy > 0) iq = 0;
            if (dx > 0 && dy >= 0) iq = 1;
            if (dx < 0 && dy <= 0) iq = 2;
        } else {
            iq = 7;
            if (dx <= 0 && dy > 0) iq = 4;
            if (dx > 0 && dy >= 0) iq = 5;
            if (dx < 0 && dy <= 0) iq = 6;
        }
        inoct[iq]++;
        if (inoct[iq] <= p->noct)
            cand[na++] = cand[j];
    }
    return na;
}

/* LAPACK dgetf2/dgetrs restated: LU with partial pivoting, row-major a[n*n]. */
static int oracle_solve(int n, double *a, double *b)
{
    /* Repair D16: LAPACK reports a singular matrix only for an exactly zero pivot; the
     * systems that ARE singular (duplicate samples, collinear drift columns) reach it with
     * pivots of order 1e-16 and come back as garbage weights of order 1e16 (SciPy adds an
     * "ill-conditioned matrix" warning, the reference prints nothing).  A pivot at or below
     * ORC_SING_TOL times the largest matrix entry is treated as zero. */
    double amax = 0.0;
    for (int i = 0; i < n * n; i++)
        if (fabs(a[i]) > amax)
            amax = fabs(a[i]);
    for (int k = 0; k < n; k++) {
        int piv = k;
        double big = fabs(a[k * n + k]);
        for (int r = k + 1; r < n; r++) {
            double v = fabs(a[r * n + k]);
            if (v > big) {
                big = v;
                piv = r;
            }
        }
        if (big <= ORC_SING_TOL * amax || big != big)
            return 1;
        if (piv != k) {
            for (int c = 0; c < n; c++) {
                double t = a[k * n + c];
                a[k * n + c] = a[piv * n + c];
                a[piv * n + c] = t;
            }
            double t = b[k];
            b[k] = b[piv];
            b[piv] = t;
        }
        double rinv = 1.0 / a[k * n + k];
        for (int r = k + 1; r < n; r++) {
            double l = a[r * n + k] * rinv;
            a[r * n + k] = l;
            if (l != 0.0) {
                for (int c = k + 1; c < n; c++)
                    a[r * n + c] -= l * a[k * n + c];
                b[r] -= l * b[k];
            }
        }
    }
    for (int k = n - 1; k >= 0; k--) {
        double s = b[k];
        for (int c = k + 1; c < n; c++)
            s -= a[k * n + c] * b[c];
        b[k] = s / a[k * n + k];
    }
    return 0;
}

/* krige3d.py:636-662 */
static int orc_mdt(const OrcParams *p)
{
    if (p->ktype == 0 || p->ktype == 2)
        return 0;
    int m = 1;
    for (int i = 0; i < ORC_MAXDRIFT; i++)
        m += p->idrift[i] ? 1 : 0;
    if (p->ktype == 3)
        m += 1;
    return m;
}

/* krige3d.py:770-801 for one sample: block-averaged covariance to the node */
static double orc_cb(const OrcParams *p, double xa, double ya, double za,
                     const double *xdb, const double *ydb, const double *zdb)
{
    if (p->ndb <= 1)
        return orc_cova3(p, xa, ya, za, xdb[0], ydb[0], zdb[0]);
    double cb = 0;
    for (int j = 0; j < p->ndb; j++) {
        cb += orc_cova3(p, xa, ya, za, xdb[j], ydb[j], zdb[j]);
        double dx = xa - xdb[j], dy = ya - ydb[j], dz = za - zdb[j];
        if (dx * dx + dy * dy + dz * dz < ORC_EPS)
            cb -= p->c0;
    }
    return cb / p->ndb;
}

/* krige3d.py:616-634 (ndb joint points, repair D9) */
double orc_block_covariance(const OrcParams *p, const double *xdb, const double *ydb,
                            const double *zdb)
{
    if (p->ndb <= 1)
        return p->unbias;
    double s = 0;
    for (int i = 0; i < p->ndb; i++)
        for (int j = 0; j < p->ndb; j++) {
            double c = orc_cova3(p, xdb[i], ydb[i], zdb[i], xdb[j], ydb[j], zdb[j]);
            if (i == j)
                c -= p->c0;
            s += c;
        }
    return s / ((double)p->ndb * p->ndb);
}

/* fills the n x n covariance block and the OK row, krige3d.py:748-768 */
void orc_left_side(const OrcParams *p, int na, int neq, const double *xa,
                   const double *ya, const double *za, double *left)
{
    for (int i = 0; i < neq * neq; i++)
        left[i] = NAN;
    for (int i = 0; i < na; i++)
        for (int j = 0; j < na; j++) {
            if (isnan(left[j * neq + i]))
                left[i * neq + j] = orc_cova3(p, xa[i], ya[i], za[i], xa[j], ya[j], za[j]);
            else
                left[i * neq + j] = left[j * neq + i];
        }
    if (neq > na) {
        for (int i = 0; i < na; i++) {
            left[na * neq + i] = p->unbias;
            left[i * neq + na] = p->unbias;
        }
        left[na * neq + na] = 0;
    }
}

void orc_right_side(const OrcParams *p, int na, int neq, const double *xa,
                    const double *ya, const double *za, const double *xdb,
                    const double *ydb, const double *zdb, double *right)
{
    for (int i = 0; i < na; i++)
        right[i] = orc_cb(p, xa[i], ya[i], za[i], xdb, ydb, zdb);
    for (int i = na; i < neq; i++)
        right[i] = p->unbias;
}

/*
 * One grid node: krige3d.py:302-403 (driver loop body), :423-468 (_one_sample),
 * :470-614 (_many_samples) with repairs D6 (drift rows), D7/D8 (external variable),
 * D10 (itrend), D15 (na==0 -> unestimated).
 */
/*
 * Body of the loop krige3d.py:302-403 for ONE location: (xloc, yloc, zloc) in super block
 * (isx, isy, isz), external value extest (ktype 2/3).  koption != 0 (cross-validation /
 * jackknife): samples at the location itself are not used (krige3d.py:359-363).
 */
static void orc_location(const OrcParams *p, double xloc, double yloc, double zloc, int isx,
                         int isy, int isz, double extest, int koption, int ntmpl,
                         const int32_t *tx, const int32_t *ty, const int32_t *tz,
                         const int64_t *nisb, const double *sx, const double *sy,
                         const double *sz, const double *sv, const double *ssec,
                         const double *xdb, const double *ydb, const double *zdb,
                         OrcCand *cand, double *work, double *est_out, double *var_out,
                         int32_t *nbr_idx, int32_t *nbr_cnt)
{
    double est = NAN, estv = NAN;
    double resce = 0, skmean = p->skmean;
    int nused = 0;
    if (nbr_cnt)
        *nbr_cnt = 0;

    if (p->ktype == 2 || p->ktype == 3) { /* krige3d.py:334-344, repair D7 */
        if (extest < p->tmin || extest >= p->tmax)
            goto done;
        resce = p->maxcov / (extest > 0.0001 ? extest : 0.0001);
    }
    int nclose = orc_search(p, xloc, yloc, zloc, isx, isy, isz, ntmpl,
                            tx, ty, tz, nisb, sx, sy, sz, cand);
    int na = 0; /* krige3d.py:356-375: the first ndmax accepted samples */
    for (int i = 0; i < nclose && na < p->ndmax; i++) {
        int64_t s = cand[i].i;
        if (koption != 0 &&
            fabs(sx[s] - xloc) + fabs(sy[s] - yloc) + fabs(sz[s] - zloc) < ORC_EPS)
            continue;
        cand[na++] = cand[i];
    }
    nused = na;
    int mdt = orc_mdt(p);
    if (nbr_cnt)
        *nbr_cnt = na;
    if (nbr_idx)
        for (int i = 0; i < na; i++)
            nbr_idx[i] = (int32_t)cand[i].i;
    if (na < p->ndmin) /* krige3d.py:377 */
        goto done;
    if (na >= 1 && na <= mdt && !(na == 1 && mdt == 1 && (p->flags & 1))) /* krige3d.py:383 */
        goto done;
    if (na == 0) /* repair D15 */
        goto done;

    int neq = na + mdt;
    double *xa = work, *ya = xa + na, *za = ya + na, *vra = za + na, *vea = vra + na;
    double *right = vea + na, *rcopy = right + neq, *left = rcopy + neq;
    for (int i = 0; i < na; i++) {
        int64_t s = cand[i].i;
        xa[i] = sx[s] - xloc;
        ya[i] = sy[s] - yloc;
        za[i] = sz[s] - zloc;
        vra[i] = sv[s];
        vea[i] = (p->ktype == 2 || p->ktype == 3) ? ssec[s] : 0.0; /* repair D8 */
    }
    if (p->ktype == 2)
        skmean = extest; /* krige3d.py:458,598 */

    if (na <= 1) { /* krige3d.py:423-468 (only reachable with mdt == 0) */
        double cb = orc_cb(p, xa[0], ya[0], za[0], xdb, ydb, zdb);
        double s = cb / p->cbb;
        est = s * vra[0] + (1 - s) * skmean;
        estv = p->cbb - s * cb;
        if (mdt == 1) { /* krige2d.py:360-363 (flag bit 0) */
            est = vra[0];
            estv = p->cbb - 2.0 * cb + p->maxcov;
        }
        goto done;
    }

    orc_left_side(p, na, neq, xa, ya, za, left);
    orc_right_side(p, na, neq, xa, ya, za, xdb, ydb, zdb, right);
    /* drift rows, krige3d.py:493-583 with repair D6 */
    int im = na;
    for (int l = 0; l < ORC_MAXDRIFT + 1; l++) {
        int on = (l < ORC_MAXDRIFT) ? (mdt > 0 && p->idrift[l]) : (p->ktype == 3);
        if (!on)
            continue;
        im++;
        for (int i = 0; i < na; i++) {
            double f;
            switch (l) {
            case 0: f = xa[i] * p->resc; break;
            case 1: f = ya[i] * p->resc; break;
            case 2: f = za[i] * p->resc; break;
            case 3: f = xa[i] * xa[i] * p->resc; break;
            case 4: f = ya[i] * ya[i] * p->resc; break;
            case 5: f = za[i] * za[i] * p->resc; break;
            case 6: f = xa[i] * ya[i] * p->resc; break;
            case 7: f = xa[i] * za[i] * p->resc; break;
            case 8: f = ya[i] * za[i] * p->resc; break;
            default: f = vea[i] * resce; break;
            }
            left[im * neq + i] = f;
            left[i * neq + im] = f;
        }
        right[im] = (l < ORC_MAXDRIFT) ? p->bv[l] : extest * resce;
    }
    for (int i = na; i < neq; i++) /* constraint block is zero */
        for (int j = na; j < neq; j++)
            left[i * neq + j] = 0.0;
    if (p->itrend) /* repair D10: kriging the trend zeroes the data covariances */
        for (int i = 0; i < na; i++)
            right[i] = 0.0;

    memcpy(rcopy, right, sizeof(double) * neq);
    if (oracle_solve(neq, left, rcopy)) /* krige3d.py:590-596 */
        goto done;
    /* krige3d.py:598-614 */
    double sum = 0;
    for (int i = 0; i < neq; i++)
        sum += rcopy[i] * right[i];
    estv = p->cbb - sum;
    est = 0;
    for (int i = 0; i < na; i++) {
        if (p->ktype == 0)
            est += rcopy[i] * (vra[i] - skmean);
        else if (p->ktype == 2)
            est += rcopy[i] * (vra[i] - vea[i]);
        else
            est += rcopy[i] * vra[i];
    }
    if (p->ktype == 0 || p->ktype == 2)
        est += skmean;
done:
    (void)nused;
    *est_out = est;
    *var_out = estv;
}

/* grid node `index` (x fastest, krige3d.py:303-309) */
static void orc_node(const OrcParams *p, int64_t index, int ntmpl, const int32_t *tx,
                     const int32_t *ty, const int32_t *tz, const int64_t *nisb,
                     const int32_t *lutx, const int32_t *luty, const int32_t *lutz,
                     const double *sx, const double *sy, const double *sz,
                     const double *sv, const double *ssec, const double *extgrid,
                     const double *xdb, const double *ydb, const double *zdb,
                     OrcCand *cand, double *work, double *est_out, double *var_out,
                     int32_t *nbr_idx, int32_t *nbr_cnt)
{
    int64_t nxy = (int64_t)p->nx * p->ny;
    int iz = (int)(index / nxy);
    int iy = (int)((index - iz * nxy) / p->nx);
    int ix = (int)(index - iz * nxy - (int64_t)iy * p->nx);
    double xloc = p->xmn + ix * p->xsiz;
    double yloc = p->ymn + iy * p->ysiz;
    double zloc = p->zmn + iz * p->zsiz;
    double extest = (p->ktype == 2 || p->ktype == 3) ? extgrid[index] : 0.0;
    orc_location(p, xloc, yloc, zloc, lutx[ix], luty[iy], lutz[iz], extest, 0, ntmpl, tx, ty,
                 tz, nisb, sx, sy, sz, sv, ssec, xdb, ydb, zdb, cand, work, est_out, var_out,
                 nbr_idx, nbr_cnt);
}

/*
 * Kriges nodes [node0, node1) of the grid (x fastest, krige3d.py:302-309) or, when
 * node_list != NULL, the nnodes flat indices it holds.  est/var/nbr are indexed by
 * position in that range/list.  nbr_idx may be NULL.  Returns 0.
 */
int kt3d_oracle_run(const OrcParams *p, int64_t nd, const double *sx, const double *sy,
                    const double *sz, const double *sv, const double *ssec,
                    const int64_t *nisb, int ntmpl, const int32_t *tx,
                    const int32_t *ty, const int32_t *tz, const int32_t *lutx,
                    const int32_t *luty, const int32_t *lutz, const double *extgrid,
                    const double *xdb, const double *ydb, const double *zdb,
                    int64_t node0, int64_t node1, const int64_t *node_list,
                    int64_t nnodes, double *est, double *var, int32_t *nbr_idx,
                    int32_t *nbr_cnt, int nthreads)
{
    int64_t count = node_list ? nnodes : node1 - node0;
    int neqmax = p->ndmax + ORC_MAXDRIFT + 2;
#ifdef _OPENMP
    if (nthreads > 0)
        omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        OrcCand *cand = (OrcCand *)malloc(sizeof(OrcCand) * (size_t)(nd > 0 ? nd : 1));
        double *work = (double *)malloc(sizeof(double) *
                                        (size_t)(7 * neqmax + neqmax * neqmax));
#pragma omp for schedule(dynamic, 64)
        for (int64_t k = 0; k < count; k++) {
            int64_t index = node_list ? node_list[k] : node0 + k;
            orc_node(p, index, ntmpl, tx, ty, tz, nisb, lutx, luty, lutz, sx, sy, sz,
                     sv, ssec, extgrid, xdb, ydb, zdb, cand, work, est + k, var + k,
                     nbr_idx ? nbr_idx + k * (int64_t)p->ndmax : NULL,
                     nbr_cnt ? nbr_cnt + k : NULL);
        }
        free(cand);
        free(work);
    }
    return 0;
}

/*
 * Cross-validation / jackknife (krige3d.py:310-332, repair D11): kriges the npts listed
 * locations; sb[3*k..] is the super block of location k (getindx, super_block.py:319-346),
 * ext[k] its external-drift value (ktype 2/3).  Results indexed by k.
 */
int kt3d_oracle_run_points(const OrcParams *p, int64_t nd, const double *sx, const double *sy,
                           const double *sz, const double *sv, const double *ssec,
                           const int64_t *nisb, int ntmpl, const int32_t *tx,
                           const int32_t *ty, const int32_t *tz, const double *xdb,
                           const double *ydb, const double *zdb, int64_t npts,
                           const double *px, const double *py, const double *pz,
                           const double *ext, const int32_t *sb, int koption, double *est,
                           double *var, int32_t *nbr_idx, int32_t *nbr_cnt, int nthreads)
{
    int neqmax = p->ndmax + ORC_MAXDRIFT + 2;
#ifdef _OPENMP
    if (nthreads > 0)
        omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        OrcCand *cand = (OrcCand *)malloc(sizeof(OrcCand) * (size_t)(nd > 0 ? nd : 1));
        double *work = (double *)malloc(sizeof(double) *
                                        (size_t)(7 * neqmax + neqmax * neqmax));
#pragma omp for schedule(dynamic, 64)
        for (int64_t k = 0; k < npts; k++)
            orc_location(p, px[k], py[k], pz[k], sb[3 * k], sb[3 * k + 1], sb[3 * k + 2],
                         ext ? ext[k] : 0.0, koption, ntmpl, tx, ty, tz, nisb, sx, sy, sz, sv,
                         ssec, xdb, ydb, zdb, cand, work, est + k, var + k,
                         nbr_idx ? nbr_idx + k * (int64_t)p->ndmax : NULL,
                         nbr_cnt ? nbr_cnt + k : NULL);
        free(cand);
        free(work);
    }
    return 0;
}

int kt3d_oracle_sizeof_params(void) { return (int)sizeof(OrcParams); }
