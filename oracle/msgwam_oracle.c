/*
 * msgwam_oracle.c -- CPU restatement of the python-msgwam hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under python-msgwam_b200/ may import,
 * link or call this file.  It is used by tests/, by __graft_entry__.smoke()
 * and by bench.py's cpu_baseline / --impl reference legs, as the checker
 * and the timed CPU baseline -- never as the product path.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit (or to
 * the stated ulp bound where libm trig is involved) against the unmodified
 * Python reference imported from /root/reference/lib/libprop.py in
 * tests/test_oracle_vs_reference.py, and against the committed fixtures in
 * tests/golden/ (made by tests/golden/make_golden.py from that same reference).
 *
 * The reference is numpy float64 code; each numpy elementwise operation rounds
 * once, left to right, with no fused multiply-add.  This file therefore must be
 * compiled with  -ffp-contract=off  (see oracle/Makefile) and every expression
 * below keeps the reference's association.  Scalars that the reference derives
 * with Python-float arithmetic (bvf**2, 2*ROT_EARTH, kappa**2*.5, sin(phi0) ...)
 * are computed by the *caller* with those same Python expressions and arrive
 * here inside orc_params, so that no libm pow() difference can creep in.
 *
 * Reference line numbers (L:nnn) are /root/reference/lib/libprop.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    double n2;          /* model_config['bvf'] ** 2                     L:380,383 */
    double two_rot;     /* 2 * ROT_EARTH                                L:382     */
    double rad_earth;   /* RAD_EARTH                                    L:3       */
    double c8rot2;      /* 8 * ROT_EARTH**2                             L:491     */
    double f0;          /* 2 * ROT_EARTH * np.sin(phi0)                 L:535,589 */
    double f0sq;        /* f0 ** 2   (numpy scalar power)               L:383,601 */
    double k2half;      /* kappa**2 * .5                                L:601     */
    double dz_grid;     /* np.diff(grid[:2])[0]                         L:349,662 */
    double dz_grids;    /* np.diff(grids[:2])[0]  (deposit grid = grids) L:123    */
    int32_t ngrid;      /* len(grid); len(grids) = ngrid-1                        */
    int32_t hprop;      /* HPROP_GLOBAL                                 L:5       */
    int32_t saturate_online; /* model_config['saturate_online']         L:633     */
    int32_t nthreads;   /* 1 = reference order; >1 = OpenMP timing mode           */
    /* EXTENSION (not in the reference, parity unpinned -- DESIGN.md section 9): height-dependent buoyancy
       frequency.  bvf_prof = N on the staggered grid bvf_grids (ngrid-1 points) or NULL for the reference's
       scalar.  N(z) = np.interp(z, grids, bvf), N^2(z) = N(z)**2 wherever the reference uses bvf**2, evaluated
       at the position argument the reference already passes to cg_rr / cg_lambda / cg_phi. */
    const double *bvf_prof;
    const double *bvf_grids;
} orc_params;

#define INVALID_CELL (-99999)

/* ------------------------------------------------------------------------ */
/* np.interp(x, xp, fp) for float64, default left/right.                     */
/* numpy/_core/src/multiarray/compiled_base.c: arr_interp + binary search.   */
/* Call sites: L:355-358, L:400, L:424, L:595.                               */
static inline double interp1(double x, const double *xp, const double *fp, long m)
{
    if (isnan(x)) return x;
    if (m == 1) return fp[0];               /* numpy: single point -> constant (x<xp: lval, >: rval, ==: fp[0]) */
    if (x < xp[0]) return fp[0];
    if (x > xp[m - 1]) return fp[m - 1];
    /* j with xp[j] <= x < xp[j+1]; x == xp[m-1] -> j = m-1 */
    long lo = 0, hi = m;                    /* invariant: xp[lo] <= x, (hi==m or x < xp[hi]) */
    while (hi - lo > 1) {
        long mid = lo + ((hi - lo) >> 1);
        if (x >= xp[mid]) lo = mid; else hi = mid;
    }
    long j = lo;
    if (j == m - 1) return fp[j];
    if (xp[j] == x) return fp[j];
    {
        const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
        double r = slope * (x - xp[j]) + fp[j];
        if (isnan(r)) {
            r = slope * (x - xp[j + 1]) + fp[j + 1];
            if (isnan(r) && fp[j] == fp[j + 1]) r = fp[j];
        }
        return r;
    }
}

void orc_interp(const double *x, long n, const double *xp, const double *fp, long m, double *out)
{
    for (long i = 0; i < n; ++i) out[i] = interp1(x[i], xp, fp, m);
}

/* N^2 at height z: the reference's scalar, or (extension) the square of the interpolated profile */
static inline double n2_at(const orc_params *P, double z)
{
    if (!P->bvf_prof) return P->n2;
    const double nn = interp1(z, P->bvf_grids, P->bvf_prof, P->ngrid - 1);
    return nn * nn;
}

/* ------------------------------------------------------------------------ */
/* omega(kk, ll, mm, phi)  L:369-383; f and f**2 supplied by the caller so    */
/* both the array-phi (ff*ff) and the scalar-phi0 (pow) forms are covered.    */
static inline double omega1(double kk, double ll, double mm, double f2, double n2)
{
    const double kh2 = kk * kk + ll * ll;
    const double m2 = mm * mm;
    return sqrt((n2 * kh2 + f2 * m2) / (kh2 + m2));
}

/* cg_rr(kk, ll, mm, lam, phi, rr)  L:434-448 (lam, rr are ignored there) */
static inline double cg_rr1(double kk, double ll, double mm, double ff, double n2)
{
    const double vk = kk * kk + ll * ll + mm * mm;
    const double f2 = ff * ff;
    const double om = omega1(kk, ll, mm, f2, n2);
    return (-mm) * (om * om - f2) / om / vk;
}

void orc_omega(long n, const double *kk, const double *ll, const double *mm, const double *phi,
               const double *rr /* extension only, may be NULL */, const orc_params *P, double *out)
{
    for (long i = 0; i < n; ++i) {
        const double ff = P->two_rot * sin(phi[i]);
        out[i] = omega1(kk[i], ll[i], mm[i], ff * ff, rr ? n2_at(P, rr[i]) : P->n2);
    }
}

/* omega with a scalar latitude: f**2 is a numpy-scalar power computed by the caller (L:597) */
void orc_omega_scalar_phi(long n, const double *kk, const double *ll, const double *mm, double f2,
                          const double *rr /* extension only, may be NULL */, const orc_params *P, double *out)
{
    for (long i = 0; i < n; ++i) out[i] = omega1(kk[i], ll[i], mm[i], f2, rr ? n2_at(P, rr[i]) : P->n2);
}

void orc_cg_rr(long n, const double *kk, const double *ll, const double *mm, const double *phi,
               const double *rr /* extension only, may be NULL */, const orc_params *P, double *out)
{
    for (long i = 0; i < n; ++i)
        out[i] = cg_rr1(kk[i], ll[i], mm[i], P->two_rot * sin(phi[i]), rr ? n2_at(P, rr[i]) : P->n2);
}

/* ------------------------------------------------------------------------ */
/* wave_projection  L:92-221.  `grid` is whatever grid the caller passes      */
/* (the staggered grid on the hot path, L:657; the full grid in diagnostics). */
/* Cell indices: L:123-135.                                                   */
static inline void cell_range(double rr_low, double rr_up, double dz, long nzmax, long *nlow_o, long *nup_o)
{
    long nlow = (long)(rr_low / dz);            /* astype(int): truncation */
    long nup = (long)(rr_up / dz + 1.);
    const int ood = ((nlow >= nzmax) && (nup >= nzmax)) || ((nlow <= 0) && (nup <= 0));
    if (ood) { *nlow_o = INVALID_CELL; *nup_o = INVALID_CELL; return; }
    if (nup < 0) nup = 0;
    if (nup >= nzmax) nup = nzmax;
    if (nlow < 0) nlow = 0;
    if (nlow >= nzmax) nlow = nzmax;
    *nlow_o = nlow; *nup_o = nup;
}

/* out layout: var 0 -> (2, ng-1); var 1,2 -> (ng-1,); var 3 -> (ng,); var 4 -> (2, ng) */
void orc_wave_projection(int var, long n,
                         const double *dens, const double *phi,
                         const double *rr_low, const double *rr_up,
                         const double *kk, const double *ll,
                         const double *mm_low, const double *mm_up,
                         const double *dkk, const double *dll, const double *dmm,
                         const double *grid, long ng, const orc_params *P, double *out)
{
    const double dz = grid[1] - grid[0];
    const long nzmax = ng - 2;
    const long ncell_out = ng - 1;
    size_t outlen = (var == 0) ? 2 * (size_t)ncell_out : (var == 4) ? 2 * (size_t)ng : (var == 3) ? (size_t)ng : (size_t)ncell_out;
    memset(out, 0, outlen * sizeof(double));

    if (var == 3 || var == 4) {
        /* interface fluxes, L:199-219: np.sum over the selected rays.  numpy's sum is
           pairwise; the oracle accumulates sequentially -- this branch is compared to
           the reference with a tolerance, not bit-for-bit. */
        for (long i = 0; i < n; ++i) {
            long nlow, nup;
            cell_range(rr_low[i], rr_up[i], dz, nzmax, &nlow, &nup);
            if (nlow == INVALID_CELL) continue;
            const double psv = fabs(dkk[i] * dll[i] * dmm[i]);
            const double cgr = cg_rr1(kk[i], ll[i], .5 * (mm_low[i] + mm_up[i]), P->two_rot * sin(phi[i]),
                                      n2_at(P, .5 * (rr_low[i] + rr_up[i])));
            for (long nb = 1; nb < ng - 1; ++nb) {
                if (nlow < nb && nup > nb) {
                    if (var == 3) out[nb] += (cgr * dens[i]) * psv;
                    else { out[nb] += (cgr * kk[i] * dens[i]) * psv; out[ng + nb] += (cgr * ll[i] * dens[i]) * psv; }
                }
            }
        }
        return;
    }

    for (long i = 0; i < n; ++i) {              /* L:151 / 169 / 186 */
        long nlow, nup;
        cell_range(rr_low[i], rr_up[i], dz, nzmax, &nlow, &nup);
        if (nlow == INVALID_CELL) continue;     /* L:153 */
        const double psv = fabs(dkk[i] * dll[i] * dmm[i]);      /* L:137 */
        double v0, v1 = 0.0;
        if (var == 2) v0 = dens[i];             /* L:184 */
        else {
            const double cgr = cg_rr1(kk[i], ll[i], .5 * (mm_low[i] + mm_up[i]),
                                      P->two_rot * sin(phi[i]), n2_at(P, .5 * (rr_low[i] + rr_up[i])));     /* L:139-144 */
            if (var == 0) { v0 = cgr * kk[i] * dens[i]; v1 = cgr * ll[i] * dens[i]; }   /* L:148-149 */
            else v0 = cgr * dens[i];            /* L:167 */
        }
        for (long c = nlow; c < nup; ++c) {     /* L:156-163 */
            const double zmin = (grid[c] > rr_low[i]) ? grid[c] : rr_low[i];
            const double zmax = (grid[c + 1] < rr_up[i]) ? grid[c + 1] : rr_up[i];
            const double w = fabs(zmax - zmin) / dz;
            out[c] += w * psv * v0;
            if (var == 0) out[ncell_out + c] += w * psv * v1;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* saturation  L:561-615.  omh uses mm_center and the global phi0 (L:597).   */
void orc_saturation(double dt, long n, const double *dens, const double *rr_c, const double *rr_c_st,
                    const double *drr, const double *drr_st, const double *kk, const double *ll,
                    const double *mm_c, const double *mm_c_st,
                    const double *dkk, const double *dll, const double *rr_mm_area,
                    const double *grids, const double *rhobar, const orc_params *P,
                    int direct, double *out)
{
    const long G = P->ngrid - 1;
    for (long i = 0; i < n; ++i) {
        const double rr_final = rr_c[i] + rr_c_st[i] * dt;
        const double drr_final = drr[i] + drr_st[i] * dt;
        const double mm_final = mm_c[i] + mm_c_st[i] * dt;
        const double dmm_final = rr_mm_area[i] / drr_final;
        const double rho = interp1(rr_final, grids, rhobar, G);
        const double omh = omega1(kk[i], ll[i], mm_c[i], P->f0sq, n2_at(P, rr_c[i]));   /* ext: N at rr_center */
        const double psv = dkk[i] * dll[i] * dmm_final;
        const double maxd = P->k2half * rho * omh * n2_at(P, rr_final) / (mm_final * mm_final) / (omh * omh - P->f0sq);  /* ext: N at rr_final */
        const int hit = maxd < dens[i] * psv;
        if (direct) out[i] = hit ? maxd : dens[i];
        else out[i] = hit ? (maxd - dens[i]) / dt : 0.0;
    }
}

/* ------------------------------------------------------------------------ */
/* rhs_default  L:618-676.  state = [dens,lam,phi,rr,drr,kk,ll,mm,dmm],       */
/* statics = [dkk,dll,rr_mm_area]; tend = 9 ray tendencies; du,dv length G.   */
/* scratch: 2*(ng) doubles for pm_flux + 4*(G-1) for du_dz/dv_dz tables.      */
typedef struct { const double *dens, *lam, *phi, *rr, *drr, *kk, *ll, *mm, *dmm; } ray_in;

static void deposit_var0(long i0, long i1, const ray_in *s, const double *dkk, const double *dll,
                         const double *grids, long G, const orc_params *P, double *proj /* (2,G-1) */)
{
    const double dz = P->dz_grids;
    const long nzmax = G - 2, nc = G - 1;
    for (long i = i0; i < i1; ++i) {
        const double hd = .5 * s->drr[i], hm = .5 * s->dmm[i];
        const double rl = s->rr[i] - hd, ru = s->rr[i] + hd;         /* L:655 */
        const double ml = s->mm[i] - hm, mu = s->mm[i] + hm;         /* L:656 */
        long nlow, nup;
        cell_range(rl, ru, dz, nzmax, &nlow, &nup);
        if (nlow == INVALID_CELL) continue;
        const double psv = fabs(dkk[i] * dll[i] * s->dmm[i]);
        const double cgr = cg_rr1(s->kk[i], s->ll[i], .5 * (ml + mu), P->two_rot * sin(s->phi[i]), n2_at(P, .5 * (rl + ru)));
        const double v0 = cgr * s->kk[i] * s->dens[i], v1 = cgr * s->ll[i] * s->dens[i];
        for (long c = nlow; c < nup; ++c) {
            const double zmin = (grids[c] > rl) ? grids[c] : rl;
            const double zmax = (grids[c + 1] < ru) ? grids[c + 1] : ru;
            const double w = fabs(zmax - zmin) / dz;
            proj[c] += w * psv * v0;
            proj[nc + c] += w * psv * v1;
        }
    }
}

void orc_rhs_default(double dt, long n, const double *const state[9], const double *uu, const double *vv,
                     const double *const statics[3], const double *grid, const double *grids,
                     const double *rhobar, const double *pg /* (2,G) */, const orc_params *P,
                     double *const tend[9], double *du, double *dv, double *proj_out /* (2,G-1) or NULL */)
{
    const long ng = P->ngrid, G = ng - 1, nc = G - 1;
    const ray_in s = { state[0], state[1], state[2], state[3], state[4], state[5], state[6], state[7], state[8] };
    const double *dkk = statics[0], *dll = statics[1], *area = statics[2];
    const double dzg = P->dz_grid, R = P->rad_earth;

    /* gradients(): L:349-353; tables on grid[1:-1] (G-1 points) */
    double *dudz = (double *)malloc(sizeof(double) * 3 * (size_t)(nc > 0 ? nc : 1));
    double *dvdz = dudz + nc, *dndz = dvdz + nc;
    for (long j = 0; j < nc; ++j) { dudz[j] = (uu[j + 1] - uu[j]) / dzg; dvdz[j] = (vv[j + 1] - vv[j]) / dzg; }
    if (P->bvf_prof) for (long j = 0; j < nc; ++j) dndz[j] = (P->bvf_prof[j + 1] - P->bvf_prof[j]) / dzg;   /* ext */
    const double *xg = grid + 1;

    #pragma omp parallel for schedule(static) if (P->nthreads > 1) num_threads(P->nthreads > 1 ? P->nthreads : 1)
    for (long i = 0; i < n; ++i) {
        const double kk = s.kk[i], ll = s.ll[i], mm = s.mm[i], phi = s.phi[i], rr = s.rr[i];
        const double sphi = sin(phi);
        const double ff = P->two_rot * sphi;
        const double f2 = ff * ff;
        const double vk = kk * kk + ll * ll + mm * mm;
        const double n2 = n2_at(P, rr);                               /* ext: N^2 at the ray centre */
        const double om = omega1(kk, ll, mm, f2, n2);
        double cgr_up, cgr_down;                                      /* L:635-636 */
        if (P->bvf_prof) {
            cgr_up = cg_rr1(kk, ll, mm, ff, n2_at(P, rr + .5 * s.drr[i]));
            cgr_down = cg_rr1(kk, ll, mm, ff, n2_at(P, rr - .5 * s.drr[i]));
        } else {
            cgr_up = cgr_down = (-mm) * (om * om - f2) / om / vk;     /* cg_rr ignores rr: up == down */
        }
        const double cgr = P->bvf_prof ? cg_rr1(kk, ll, mm, ff, n2) : cgr_up;   /* cg_rr at the centre (dk_dt, dl_dt) */
        const double du_ray = interp1(rr, xg, dudz, nc);              /* L:355 */
        const double dv_ray = interp1(rr, xg, dvdz, nc);              /* L:356 */
        double cgl, cgp;                                              /* cg_lambda, cg_phi L:386-431 */
        if (P->hprop) {
            const double uu_ray = interp1(rr, grids, uu, G);
            const double vv_ray = interp1(rr, grids, vv, G);
            cgl = kk / om / vk * (n2 - om * om) + uu_ray;
            cgp = ll / om / vk * (n2 - om * om) + vv_ray;
        } else { cgl = 0.0; cgp = 0.0; }
        const double rad = R + rr;
        const double cphi = cos(phi);
        tend[1][i] = cgl / rad / cphi;                                /* dlam_st L:638 */
        tend[2][i] = cgp / rad;                                       /* dphi_st L:639 */
        const double drr_st = .5 * (cgr_down + cgr_up);               /* L:640 */
        const double ddrr_st = cgr_up - cgr_down;                     /* L:641 */
        tend[3][i] = drr_st;
        tend[4][i] = ddrr_st;
        if (P->hprop) {
            const double tphi = tan(phi);
            const double g_lam = (kk * 0.0 + ll * 0.0) / rad / cphi;  /* L:465 with vel[1:3,0] == 0 */
            tend[5][i] = kk / rad * (tphi * cgp - cgr) - g_lam;       /* dk_dt L:468-469 */
            const double g_phi = (kk * 0.0 + ll * 0.0) / rad;         /* L:489 */
            const double df2 = P->c8rot2 * sphi * cphi * 1;           /* L:491 */
            tend[6][i] = -(ll * cgr + kk * tphi * cgl + mm * mm / 2 / om / vk * df2) / rad - g_phi;   /* L:494-497 */
        } else { tend[5][i] = 0.0; tend[6][i] = 0.0; }
        const double g_rr = kk * du_ray + ll * dv_ray;                /* L:517 */
        double dmm_st = (kk * cgl + ll * cgp) / rad - g_rr;           /* L:519-520 */
        if (P->bvf_prof) {                                            /* ext: - N N' (k^2 + l^2) / om / |k|^2 */
            const double nr = interp1(rr, grids, P->bvf_prof, G), dnr = interp1(rr, xg, dndz, nc);
            dmm_st = dmm_st - nr * dnr * (kk * kk + ll * ll) / om / vk;
        }
        tend[7][i] = dmm_st;
        tend[8][i] = s.dmm[i] / s.drr[i] * ddrr_st;                   /* L:645 */
        {   /* saturation(dt, dens, rr, drr_st, drr, ddrr_st, kk, ll, mm, dmm_st)  L:647-651 */
            const double rr_final = rr + drr_st * dt;
            const double drr_final = s.drr[i] + ddrr_st * dt;
            const double mm_final = mm + dmm_st * dt;
            const double dmm_final = area[i] / drr_final;
            const double rho = interp1(rr_final, grids, rhobar, G);
            const double omh = omega1(kk, ll, mm, P->f0sq, n2);
            const double psv = dkk[i] * dll[i] * dmm_final;
            const double maxd = P->k2half * rho * omh * n2_at(P, rr_final) / (mm_final * mm_final) / (omh * omh - P->f0sq);
            const double st = (maxd < s.dens[i] * psv) ? (maxd - s.dens[i]) / dt : 0.0;
            tend[0][i] = (double)(P->saturate_online ? 1 : 0) * st;  /* bool * ndarray */
        }
    }

    /* deposition, L:653-660 */
    double *flux = (double *)calloc(2 * (size_t)ng + 2 * (size_t)(nc > 0 ? nc : 1), sizeof(double));
    double *proj = flux + 2 * ng;
#ifdef _OPENMP
    if (P->nthreads > 1) {
        const int T = P->nthreads;
        double *part = (double *)calloc((size_t)T * 2 * (size_t)nc, sizeof(double));
        #pragma omp parallel num_threads(T)
        {
            const int t = omp_get_thread_num();
            const long i0 = n * t / T, i1 = n * (t + 1) / T;
            deposit_var0(i0, i1, &s, dkk, dll, grids, G, P, part + (size_t)t * 2 * nc);
        }
        for (int t = 0; t < T; ++t) for (long c = 0; c < 2 * nc; ++c) proj[c] += part[(size_t)t * 2 * nc + c];
        free(part);
    } else
#endif
    deposit_var0(0, n, &s, dkk, dll, grids, G, P, proj);
    if (proj_out) memcpy(proj_out, proj, sizeof(double) * 2 * (size_t)nc);

    for (int c = 0; c < 2; ++c) {
        double *F = flux + (size_t)c * ng;
        for (long j = 0; j < nc; ++j) F[1 + j] = proj[(size_t)c * nc + j];
        F[0] = F[1];                      /* L:659 */
        F[ng - 1] = F[ng - 2];            /* L:660 */
    }
    for (long j = 0; j < G; ++j) {        /* L:663-666, du_dt L:523-539, dv_dt L:542-558 */
        const double g0 = (flux[j + 1] - flux[j]) / dzg;
        const double g1 = (flux[ng + j + 1] - flux[ng + j]) / dzg;
        const double rinv = 1.0 / rhobar[j];                          /* rhobar**-1 == np.reciprocal */
        du[j] = P->f0 * vv[j] - rinv * (pg[j] + g0);
        dv[j] = (-P->f0) * uu[j] - rinv * (pg[G + j] + g1);
    }
    free(flux);
    free(dudz);
}

/* ------------------------------------------------------------------------ */
/* RK3  L:680-700 (Williamson low-storage), rhs = rhs_default.               */
/* state_io: 9 ray arrays of length n, updated in place (caller passes copies */
/* to keep the reference's "inputs are not mutated" behaviour); uu,vv idem.   */
void orc_rk3(double dt, long n, double *const state[9], double *uu, double *vv,
             const double *const statics[3], const double *grid, const double *grids,
             const double *rhobar, const double *pg, const orc_params *P)
{
    const long G = P->ngrid - 1;
    double *buf = (double *)malloc(sizeof(double) * (18 * (size_t)(n > 0 ? n : 1) + 4 * (size_t)G));
    double *tend[9], *qq[9];
    for (int f = 0; f < 9; ++f) { tend[f] = buf + (size_t)f * n; qq[f] = buf + (size_t)(9 + f) * n; }
    double *du = buf + 18 * (size_t)n, *dv = du + G, *qu = dv + G, *qv = qu + G;
    const double a[3] = { 0.0, 5 / 9., 153 / 128. };
    const double b[3] = { 0.0, 15 / 16., 8 / 15. };
    for (int stage = 0; stage < 3; ++stage) {
        const double *cst[9];
        for (int f = 0; f < 9; ++f) cst[f] = state[f];
        orc_rhs_default(dt, n, cst, uu, vv, statics, grid, grids, rhobar, pg, P, tend, du, dv, NULL);
        if (P->nthreads > 1) {
            /* timing mode: one parallel sweep over the rays updates all nine fields (same arithmetic) */
            const double as = a[stage], bs = b[stage];
            #pragma omp parallel for schedule(static) num_threads(P->nthreads)
            for (long i = 0; i < n; ++i)
                for (int f = 0; f < 9; ++f) {
                    if (stage == 0) { qq[f][i] = dt * tend[f][i]; state[f][i] = state[f][i] + qq[f][i] / 3; }
                    else { qq[f][i] = dt * tend[f][i] - as * qq[f][i]; state[f][i] = state[f][i] + bs * qq[f][i]; }
                }
        } else
        for (int f = 0; f < 9; ++f) {
            double *x = state[f], *q = qq[f]; const double *t = tend[f];
            if (stage == 0) {
                for (long i = 0; i < n; ++i) { q[i] = dt * t[i]; x[i] = x[i] + q[i] / 3; }     /* L:693-694 */
            } else {
                const double as = a[stage], bs = b[stage];
                for (long i = 0; i < n; ++i) { q[i] = dt * t[i] - as * q[i]; x[i] = x[i] + bs * q[i]; }   /* L:695-698 */
            }
        }
        for (long j = 0; j < G; ++j) {
            if (stage == 0) { qu[j] = dt * du[j]; uu[j] = uu[j] + qu[j] / 3; qv[j] = dt * dv[j]; vv[j] = vv[j] + qv[j] / 3; }
            else {
                qu[j] = dt * du[j] - a[stage] * qu[j]; uu[j] = uu[j] + b[stage] * qu[j];
                qv[j] = dt * dv[j] - a[stage] * qv[j]; vv[j] = vv[j] + b[stage] * qv[j];
            }
        }
    }
    free(buf);
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
