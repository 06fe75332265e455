// general.cu -- every branch of the reference right-hand side, one stage at a time, plus the
// step-adjacent functions of the drop-in boundary (wave_projection var 0..4, saturation, the point
// functions and the low-storage update).  These kernels serve
//   * rhs_default called directly / plugged into RK3 by the user (L:618-676),
//   * RK3 when HPROP_GLOBAL or saturate_online is on (all nine ray fields then change),
//   * the diagnostics around the step (R:183-188, 213-231).
// The BASELINE configurations run through column_step.cu; this file is the complete, slower path.
#include "common.cuh"
#include "deposit.cuh"

namespace {

using namespace mw;

constexpr int NT = 256;

// gradients() (L:349-358) without tables: du_dz/dv_dz on grid[1:-1] evaluated on the fly, which is
// bit-identical to building the table first.
__device__ __forceinline__ void shear_interp(double x, const double *__restrict__ grid, const double *__restrict__ uu,
                                             const double *__restrict__ vv, int G, double dzg, double rdz,
                                             double &du_ray, double &dv_ray)
{
    const int nc = G - 1;
    const double *xg = grid + 1;
    if (x != x) { du_ray = x; dv_ray = x; return; }
    int j; bool flat;
    if (x <= xg[0]) { j = 0; flat = true; }
    else if (x >= xg[nc - 1]) { j = nc - 1; flat = true; }
    else { j = interp_locate(x, xg, nc, rdz); flat = false; }
    // divisions by the loop-invariant dz_grid in the exact invariant-divisor form (rdz = RN(1 / dzg); common.cuh)
    const double fu0 = div_inv_safe(sub(uu[j + 1], uu[j]), dzg, rdz), fv0 = div_inv_safe(sub(vv[j + 1], vv[j]), dzg, rdz);
    const double dx = flat ? 0.0 : sub(x, xg[j]);
    if (flat || dx == 0.0) { du_ray = fu0; dv_ray = fv0; return; }
    const double fu1 = div_inv_safe(sub(uu[j + 2], uu[j + 1]), dzg, rdz), fv1 = div_inv_safe(sub(vv[j + 2], vv[j + 1]), dzg, rdz);
    const double w = sub(xg[j + 1], xg[j]);
    const double nu = sub(fu1, fu0), nv = sub(fv1, fv0);
    du_ray = add(mul(w == dzg ? div_inv_safe(nu, dzg, rdz) : dvd(nu, w), dx), fu0);
    dv_ray = add(mul(w == dzg ? div_inv_safe(nv, dzg, rdz) : dvd(nv, w), dx), fv0);
}

struct RhsArgs {
    msgwam_params_t p;
    msgwam_rays_t r;
    int64_t n;
    const double *grid, *grids, *rhobar, *uu, *vv, *bvf;
    double *tend[9];
};

// rhs_default's nine ray tendencies (L:629-651), all branches, for ray i: x[9] receives the state, t[9] the tendencies
// need_sat = false (the fused RK stage with saturate_online off) skips the saturation threshold, whose result the
// reference multiplies by False (L:647): the tendency is then +0.0 instead of -0.0 where the threshold is exceeded,
// which no state can tell apart.
// ff_out / cgr_out, if given, receive 2 Omega sin(phi) and cg_rr at the ray centre for the caller's deposit.
// The divisions that share a divisor -- by RAD_EARTH + rr, om, |k|^2, cos(phi) -- go through SharedDiv (one refined
// reciprocal per divisor, quotients bit-identical to the IEEE division).
__device__ __forceinline__ void ray_rhs(const RhsArgs &a, int64_t i, double x[9], double t[9], bool need_sat = true,
                                        double *ff_out = nullptr, double *cgr_out = nullptr)
{
    const msgwam_params_t &p = a.p;
    const int G = p.G;
    const double lam = a.r.lam ? a.r.lam[i] : 0.0;
    const double dens = a.r.dens[i], phi = a.r.phi[i], rr = a.r.rr[i], drr = a.r.drr[i];
    const double kk = a.r.kk[i], ll = a.r.ll[i], mm = a.r.mm[i], dmm = a.r.dmm[i];
    // HPROP off: phi does not move, so an ensemble's derived static ff = 2 Omega sin(phi) (msgwam_derive_statics, the same
    // expression) stands in for the two trigonometric calls per ray and stage; lam_st, phi_st are then exact zeros
    // (L:638-639: zeros / (RAD + rr) / cos(phi))
    const bool use_ff = !p.hprop && a.r.ff != nullptr;
    double sphi = 0.0, cphi = 1.0, ff;
    if (use_ff) ff = a.r.ff[i];
    else { sphi = sin(phi); cphi = cos(phi); ff = mul(p.two_rot, sphi); }
    const double f2 = mul(ff, ff);
    const double kh2 = add(mul(kk, kk), mul(ll, ll)), m2 = mul(mm, mm);
    const double vk = add(kh2, m2);
    const double n2 = n2_at(a.bvf, a.grids, G, p.inv_dz_grids, p.n2, rr);     // ext: N^2 at the ray centre
    double om;                                                                // L:383
    const double cgr = cg_rr_fast(kh2, mm, f2, n2, &om);                      // L:448 (at the centre); same roundings
    double cgr_up = cgr, cgr_down = cgr;                                      // cg_rr ignores rr: up == down (L:635-636)
    if (a.bvf != nullptr) {                                                   // ext: N at the two edges
        const double hd = mul(.5, drr);
        cgr_up = cg_rr_fast(kh2, mm, f2, n2_at(a.bvf, a.grids, G, p.inv_dz_grids, p.n2, add(rr, hd)));
        cgr_down = cg_rr_fast(kh2, mm, f2, n2_at(a.bvf, a.grids, G, p.inv_dz_grids, p.n2, sub(rr, hd)));
    }
    double du_ray, dv_ray;
    shear_interp(rr, a.grid, a.uu, a.vv, G, p.dz_grid, p.inv_dz_grid, du_ray, dv_ray);
    if (ff_out) *ff_out = ff;
    if (cgr_out) *cgr_out = cgr;
    double cgl = 0.0, cgp = 0.0;
    const SharedDiv by_om(om), by_vk(vk);
    if (p.hprop) {                                                            // L:400-405, 424-429
        const double uu_ray = interp1_dz(rr, a.grids, a.uu, G, p.dz_grids, p.inv_dz_grids);
        const double vv_ray = interp1_dz(rr, a.grids, a.vv, G, p.dz_grids, p.inv_dz_grids);
        const double nd = sub(n2, mul(om, om));
        cgl = add(mul(by_vk(by_om(kk)), nd), uu_ray);
        cgp = add(mul(by_vk(by_om(ll)), nd), vv_ray);
    }
    const double rad = add(p.rad_earth, rr);
    const SharedDiv by_rad(rad), by_cphi(cphi);
    const double drr_st = mul(.5, add(cgr_down, cgr_up));                     // L:640
    const double ddrr_st = sub(cgr_up, cgr_down);                             // L:641
    double dkk_st = 0.0, dll_st = 0.0;
    if (p.hprop) {
        const double tphi = tan(phi);
        const double zero = add(mul(kk, 0.0), mul(ll, 0.0));
        const double g_lam = by_cphi(by_rad(zero));                           // L:465
        dkk_st = sub(mul(by_rad(kk), sub(mul(tphi, cgp), cgr)), g_lam);       // L:468-469
        const double g_phi = by_rad(zero);                                    // L:489
        const double df2 = mul(mul(mul(p.c8rot2, sphi), cphi), 1.0);          // L:491
        const double t3 = mul(by_vk(by_om(mul(m2, .5))), df2);                // m2 / 2 == m2 * .5, exactly
        const double sum = add(add(mul(ll, cgr), mul(mul(kk, tphi), cgl)), t3);
        dll_st = sub(by_rad(-sum), g_phi);                                    // L:494-497
    }
    const double g_rr = add(mul(kk, du_ray), mul(ll, dv_ray));                // L:517
    double dmm_st = sub(by_rad(add(mul(kk, cgl), mul(ll, cgp))), g_rr);       // L:519-520
    if (a.bvf != nullptr) {                                                   // ext: - N N' (k^2 + l^2) / om / |k|^2
        const double nr = interp1(rr, a.grids, a.bvf, G, p.inv_dz_grids);
        double dnr, unused;
        shear_interp(rr, a.grid, a.bvf, a.bvf, G, p.dz_grid, p.inv_dz_grid, dnr, unused);
        dmm_st = sub(dmm_st, by_vk(by_om(mul(mul(nr, dnr), kh2))));
    }
    double st = 0.0;
    if (need_sat) {
        double maxd;
        const bool hit = saturation_limit(p, p.dt, dens, rr, drr_st, drr, ddrr_st, kk, ll, mm, dmm_st,
                                          mul(a.r.dkk[i], a.r.dll[i]), a.r.rr_mm_area[i], a.grids, a.rhobar, a.bvf, maxd);
        st = hit ? dvd(sub(maxd, dens), p.dt) : 0.0;                          // L:612-615
    }
    x[0] = dens; x[1] = lam; x[2] = phi; x[3] = rr; x[4] = drr; x[5] = kk; x[6] = ll; x[7] = mm; x[8] = dmm;
    t[0] = mul(p.saturate_online ? 1.0 : 0.0, st);                            // L:647
    t[1] = use_ff ? 0.0 : by_cphi(by_rad(cgl));                               // L:638
    t[2] = use_ff ? 0.0 : by_rad(cgp);                                        // L:639
    t[3] = drr_st;
    t[4] = ddrr_st;
    t[5] = dkk_st;
    t[6] = dll_st;
    t[7] = dmm_st;
    t[8] = mul(dvd(dmm, drr), ddrr_st);                                       // L:645
}

__global__ void __launch_bounds__(NT) rhs_rays_kernel(const RhsArgs a)
{
    for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * NT) {
        double x[9], t[9];
        ray_rhs(a, i, x, t);
#pragma unroll
        for (int f = 0; f < 9; ++f) a.tend[f][i] = t[f];
    }
}

// ---- wave_projection, var = 0, 1, 2 ----------------------------------------------------------
struct ProjArgs {
    msgwam_params_t p;
    int var, centered, use_smem;
    int64_t n;
    // centered == 0: edges given (the reference signature); centered == 1: a,b = centre, extent
    const double *dens, *phi, *ra, *rb, *kk, *ll, *ma, *mb, *dkk, *dll, *dmm;
    const double *grid;
    int ng;
    double dz, rdz;
    const double *bvf, *bvf_grids;      // extension: N on the staggered grid (p.G points) or NULL
    double *out;
};

// dynamic shared memory: warp windows (always) | grid copy | histogram (when they fit)
__global__ void __launch_bounds__(NT) project_kernel(const ProjArgs a)
{
    extern __shared__ double sm[];
    const int nc = a.ng - 1;
    const int ncomp = a.var == 0 ? 2 : 1;
    double *wins = sm + ((reinterpret_cast<uintptr_t>(sm) & 8) ? 1 : 0);
    double *rest = wins + (NT / 32) * WIN_DOUBLES;
    double *h0, *h1;
    const double *g;
    Window win;
    window_init(win, wins + (size_t)(threadIdx.x >> 5) * WIN_DOUBLES);
    if (a.use_smem) {
        double *gs = rest; h0 = rest + a.ng; h1 = h0 + nc;
        for (int j = threadIdx.x; j < a.ng; j += NT) gs[j] = a.grid[j];
        for (int j = threadIdx.x; j < 2 * nc; j += NT) h0[j] = 0.0;
        g = gs;
    } else {
        // grids too large for shared memory: accumulate straight into the output.  The second
        // component of single-component variants is identically zero and lands on the first.
        g = a.grid; h0 = a.out; h1 = a.out + (ncomp == 2 ? nc : 0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (NT / 32);
    const int64_t gw = (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    const int64_t per = (((a.n + nwarps - 1) / nwarps) + 31) & ~(int64_t)31;
    const int64_t begin = gw * per;
    const int64_t end = (begin + per < a.n) ? begin + per : a.n;
    for (int64_t base = begin; base < end; base += 32) {
        const int64_t i = base + lane;
        const bool live = i < end;
        double rl = 0.0, ru = 0.0, psv = 0.0, v0 = 0.0, v1 = 0.0;
        int nlow = 0, nup = 0;
        bool ok = false;
        if (live) {
            double ml, mu;
            if (a.centered) {
                const double hd = mul(.5, a.rb[i]), hm = mul(.5, a.mb[i]);
                rl = sub(a.ra[i], hd); ru = add(a.ra[i], hd);
                ml = sub(a.ma[i], hm); mu = add(a.ma[i], hm);
            } else { rl = a.ra[i]; ru = a.rb[i]; ml = a.ma[i]; mu = a.mb[i]; }
            ok = cell_range(rl, ru, a.dz, a.rdz, a.ng - 2, nlow, nup);
            if (ok) {
                psv = fabs(mul(mul(a.dkk[i], a.dll[i]), a.dmm[i]));              // L:137
                const double dens = a.dens[i];
                if (a.var == 2) v0 = dens;                                       // L:184
                else {
                    const double kk = a.kk[i], ll = a.ll[i];
                    const double ff = mul(a.p.two_rot, sin(a.phi[i]));
                    const double n2 = n2_at(a.bvf, a.bvf_grids, a.p.G, a.p.inv_dz_grids, a.p.n2, mul(.5, add(rl, ru)));
                    const double cgr = cg_rr_from(add(mul(kk, kk), mul(ll, ll)), mul(.5, add(ml, mu)), mul(ff, ff), n2);
                    if (a.var == 0) { v0 = mul(mul(cgr, kk), dens); v1 = mul(mul(cgr, ll), dens); }   // L:148-149
                    else v0 = mul(cgr, dens);                                    // L:167
                }
            }
        }
        deposit_cells(ok, nlow, nup, rl, ru, psv, v0, v1, a.dz, a.rdz, g, win, h0, h1, SplitTargets{h0, h1, nullptr});
    }
    window_flush(win, h0, h1);
    if (a.use_smem) {
        __syncthreads();
        for (int j = threadIdx.x; j < ncomp * nc; j += NT) {
            const double v = h0[j];
            if (v != 0.0) atomicAdd(a.out + j, v);
        }
    }
}

// var = 3, 4: fluxes through the interfaces strictly inside the ray volume's cell range (L:199-219)
__global__ void __launch_bounds__(NT) project_iface_kernel(const ProjArgs a)
{
    for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * NT) {
        int nlow, nup;
        if (!cell_range(a.ra[i], a.rb[i], a.dz, a.rdz, a.ng - 2, nlow, nup)) continue;
        const double psv = fabs(mul(mul(a.dkk[i], a.dll[i]), a.dmm[i]));
        const double kk = a.kk[i], ll = a.ll[i], dens = a.dens[i];
        const double ff = mul(a.p.two_rot, sin(a.phi[i]));
        const double n2 = n2_at(a.bvf, a.bvf_grids, a.p.G, a.p.inv_dz_grids, a.p.n2, mul(.5, add(a.ra[i], a.rb[i])));
        const double cgr = cg_rr_from(add(mul(kk, kk), mul(ll, ll)), mul(.5, add(a.ma[i], a.mb[i])), mul(ff, ff), n2);
        const int b0 = max(nlow + 1, 1), b1 = min(nup - 1, a.ng - 2);
        if (a.var == 3) {
            const double t = mul(mul(cgr, dens), psv);
            for (int nb = b0; nb <= b1; ++nb) atomicAdd(a.out + nb, t);
        } else {
            const double t0 = mul(mul(mul(cgr, kk), dens), psv), t1 = mul(mul(mul(cgr, ll), dens), psv);
            for (int nb = b0; nb <= b1; ++nb) { atomicAdd(a.out + nb, t0); atomicAdd(a.out + a.ng + nb, t1); }
        }
    }
}

// ---- du_dt, dv_dt from the deposit (L:653-666, 523-558) -------------------------------------
__global__ void grid_tendency_kernel(msgwam_params_t p, const double *__restrict__ rhobar, const double *__restrict__ pg,
                                     const double *__restrict__ uu, const double *__restrict__ vv,
                                     const double *__restrict__ D, double *__restrict__ du, double *__restrict__ dv)
{
    const int G = p.G, nc = G - 1;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < G; j += gridDim.x * blockDim.x) {
        const int i0 = min(max(j - 1, 0), nc - 1), i1 = min(j, nc - 1);
        const double g0 = dvd(sub(D[i1], D[i0]), p.dz_grid);
        const double g1 = dvd(sub(D[nc + i1], D[nc + i0]), p.dz_grid);
        const double rinv = dvd(1.0, rhobar[j]);
        du[j] = sub(mul(p.f0, vv[j]), mul(rinv, add(pg[j], g0)));
        dv[j] = sub(mul(-p.f0, uu[j]), mul(rinv, add(pg[G + j], g1)));
    }
}

__global__ void mean_flow_tendency_kernel(int which, double f0, int G, const double *__restrict__ wind,
                                          const double *__restrict__ fg, const double *__restrict__ rhobar,
                                          const double *__restrict__ pg, double *__restrict__ out)
{
    const double f = which == 0 ? f0 : -f0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < G; j += gridDim.x * blockDim.x)
        out[j] = sub(mul(f, wind[j]), mul(dvd(1.0, rhobar[j]), add(pg[j], fg[j])));
}

// ---- low-storage RK update (L:693-698) -------------------------------------------------------
__global__ void rk_update_kernel(int stage, double dt, const double *__restrict__ t, double *__restrict__ q,
                                 const double *x, double *xo, int64_t n)
{
    const double as = stage == 1 ? 5 / 9. : 153 / 128., bs = stage == 1 ? 15 / 16. : 8 / 15.;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (stage == 0) {
            const double qq = mul(dt, t[i]);
            q[i] = qq; xo[i] = add(x[i], dvd(qq, 3.0));
        } else {
            const double qq = sub(mul(dt, t[i]), mul(as, q[i]));
            q[i] = qq; xo[i] = add(x[i], mul(bs, qq));
        }
    }
}

// ---- one fused RK stage for any mode (HPROP, online saturation, N(z) profile) ------------------------------
// rhs_default on state x_s (ray_rhs), its deposit (wave_projection var = 0 of the same state, L:653-658) and the
// low-storage update of all nine ray slots (L:693-698) in one sweep: 9 + 3 + 9 fields read, 9 + 9 written, instead
// of the rhs / projection / 9 x rk_update kernels (~1.9 KB/ray-step of traffic and ~45 launches per step).
#ifndef MSGWAM_STAGE_CTAS
#define MSGWAM_STAGE_CTAS 3
#endif
struct StageArgs {
    RhsArgs r;                  // state in, statics, grid fields, uu_s, vv_s
    int stage;
    double *q[9];               // low-storage register, in/out (written only at stage 0)
    double *xo[9];              // state out (may alias the state in)
    double *proj;               // (2, G-1) deposit, accumulated (zero on entry)
    int use_smem;
};

// dynamic shared memory: warp windows | grids copy | CTA histogram (when they fit), as in project_kernel
__global__ void __launch_bounds__(NT, MSGWAM_STAGE_CTAS) stage_rays_kernel(const StageArgs a)
{
    extern __shared__ double sm[];
    const msgwam_params_t &p = a.r.p;
    const int ng = p.G, nc = ng - 1;
    double *wins = sm + ((reinterpret_cast<uintptr_t>(sm) & 8) ? 1 : 0);
    double *rest = wins + (NT / 32) * WIN_DOUBLES;
    double *h0, *h1;
    const double *g;
    Window win;
    window_init(win, wins + (size_t)(threadIdx.x >> 5) * WIN_DOUBLES);
    if (a.use_smem) {
        double *gsm = rest; h0 = rest + ng; h1 = h0 + nc;
        for (int j = threadIdx.x; j < ng; j += NT) gsm[j] = a.r.grids[j];
        for (int j = threadIdx.x; j < 2 * nc; j += NT) h0[j] = 0.0;
        g = gsm;
    } else { g = a.r.grids; h0 = a.proj; h1 = a.proj + nc; }
    __syncthreads();
    const double as = a.stage == 1 ? 5 / 9. : 153 / 128., bs = a.stage == 1 ? 15 / 16. : 8 / 15.;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (NT / 32);
    const int64_t gw = (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    const int64_t per = (((a.r.n + nwarps - 1) / nwarps) + 31) & ~(int64_t)31;
    const int64_t begin = gw * per;
    const int64_t end = (begin + per < a.r.n) ? begin + per : a.r.n;
    for (int64_t base = begin; base < end; base += 32) {
        const int64_t i = base + lane;
        const bool live = i < end;
        double rl = 0.0, ru = 0.0, psv = 0.0, v0 = 0.0, v1 = 0.0;
        int nlow = 0, nup = 0;
        bool ok = false;
        double x[9], t[9];
        if (i + 32 < end) {
            // the next iteration's lines into L2: the sweep waited on its own loads (ncu: 8.9 stall cycles per issue
            // on the long scoreboard with 24 warps per SM)
            const msgwam_rays_t &rs = a.r.r;
            const double *in[12] = {rs.dens, rs.lam, rs.phi, rs.rr, rs.drr, rs.kk, rs.ll, rs.mm, rs.dmm, rs.dkk, rs.dll, rs.rr_mm_area};
#pragma unroll
            for (int f = 0; f < 12; ++f) if (in[f]) asm volatile("prefetch.global.L2 [%0];" ::"l"(in[f] + i + 32));
            if (a.stage != 0) {
#pragma unroll
                for (int f = 0; f < 9; ++f) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.q[f] + i + 32));
            }
        }
        if (live) {
            double ff, cgr_c;
            ray_rhs(a.r, i, x, t, p.saturate_online != 0, &ff, &cgr_c);
            // wave_projection(var = 0) of the same state, called as L:654-658
            const double hd = mul(.5, x[4]), hm = mul(.5, x[8]);
            rl = sub(x[3], hd); ru = add(x[3], hd);
            const double ml = sub(x[7], hm), mu = add(x[7], hm);
            ok = cell_range(rl, ru, p.dz_grids, p.inv_dz_grids, ng - 2, nlow, nup);
            if (ok) {
                psv = fabs(mul(mul(a.r.r.dkk[i], a.r.r.dll[i]), x[8]));           // L:137
                // cg_rr at (.5 (m_low + m_up), .5 (r_low + r_up)): almost always bit-identical to the centre's arguments,
                // then the right-hand side's value is reused (ff = 2 Omega sin(phi) is the same expression there)
                const double mid = mul(.5, add(ml, mu)), zc = mul(.5, add(rl, ru));
                double cgr = cgr_c;
                if (mid != x[7] || (a.r.bvf != nullptr && zc != x[3])) {
                    const double n2 = n2_at(a.r.bvf, a.r.grids, ng, p.inv_dz_grids, p.n2, zc);
                    cgr = cg_rr_fast(add(mul(x[5], x[5]), mul(x[6], x[6])), mid, mul(ff, ff), n2);
                }
                v0 = mul(mul(cgr, x[5]), x[0]); v1 = mul(mul(cgr, x[6]), x[0]);   // L:148-149
            }
        }
        deposit_cells(ok, nlow, nup, rl, ru, psv, v0, v1, p.dz_grids, p.inv_dz_grids, g, win, h0, h1, SplitTargets{h0, h1, nullptr});
        if (live) {
            // slots whose tendency is identically zero in this mode (L:638-645 with HPROP off: lam, phi, kk, ll; a scalar N:
            // drr, dmm; saturate_online off: dens) keep their value: no low-storage traffic for them (their register stays
            // the exact zero it would hold), and no store at all when the step runs in place
            const double *const xin[9] = {a.r.r.dens, a.r.r.lam, a.r.r.phi, a.r.r.rr, a.r.r.drr, a.r.r.kk, a.r.r.ll, a.r.r.mm, a.r.r.dmm};
#pragma unroll
            for (int f = 0; f < 9; ++f) {                                         // L:693-698
                const bool moves = (f == 3 || f == 7) ? true : (f == 0) ? p.saturate_online != 0 : (f == 4 || f == 8) ? a.r.bvf != nullptr : p.hprop != 0;
                if (!moves) {
                    if (a.xo[f] != xin[f]) a.xo[f][i] = x[f];
                    continue;
                }
                double qq;
                if (a.stage == 0) { qq = mul(p.dt, t[f]); a.xo[f][i] = add(x[f], div_inv_safe(qq, 3.0, 1.0 / 3.0)); }   // qq / 3, exact
                else { qq = sub(mul(p.dt, t[f]), mul(as, a.q[f][i])); a.xo[f][i] = add(x[f], mul(bs, qq)); }
                a.q[f][i] = qq;
            }
        }
    }
    window_flush(win, h0, h1);
    if (a.use_smem) {
        __syncthreads();
        for (int j = threadIdx.x; j < 2 * nc; j += NT) {
            const double v = h0[j];
            if (v != 0.0) atomicAdd(a.proj + j, v);
        }
    }
}

// the mean-flow half of the stage: du_st, dv_st from the (reduced) deposit (L:653-666, 523-558), the low-storage
// update of uu, vv, and the deposit zeroed for the next stage.  One CTA.
__global__ void __launch_bounds__(1024) stage_grid_kernel(int stage, msgwam_params_t p, const double *__restrict__ rhobar,
                                                          const double *__restrict__ pg, const double *uu, const double *vv,
                                                          double *D, double *qu, double *qv, double *uo, double *vo)
{
    const int G = p.G, nc = G - 1;
    const double as = stage == 1 ? 5 / 9. : 153 / 128., bs = stage == 1 ? 15 / 16. : 8 / 15.;
    for (int j = threadIdx.x; j < G; j += blockDim.x) {
        const int i0 = min(max(j - 1, 0), nc - 1), i1 = min(j, nc - 1);
        const double g0 = dvd(sub(D[i1], D[i0]), p.dz_grid);
        const double g1 = dvd(sub(D[nc + i1], D[nc + i0]), p.dz_grid);
        const double rinv = dvd(1.0, rhobar[j]);
        const double u = uu[j], v = vv[j];
        const double du = sub(mul(p.f0, v), mul(rinv, add(pg[j], g0)));
        const double dv = sub(mul(-p.f0, u), mul(rinv, add(pg[G + j], g1)));
        double q0, q1;
        if (stage == 0) { q0 = mul(p.dt, du); q1 = mul(p.dt, dv); uo[j] = add(u, dvd(q0, 3.0)); vo[j] = add(v, dvd(q1, 3.0)); }
        else {
            q0 = sub(mul(p.dt, du), mul(as, qu[j])); q1 = sub(mul(p.dt, dv), mul(as, qv[j]));
            uo[j] = add(u, mul(bs, q0)); vo[j] = add(v, mul(bs, q1));
        }
        qu[j] = q0; qv[j] = q1;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * nc; j += blockDim.x) D[j] = 0.0;
}

// ---- saturation (L:561-615) ---------------------------------------------------------------------
struct SatArgs {
    msgwam_params_t p;
    int64_t n; int direct;
    const double *dens, *rr, *rr_st, *drr, *drr_st, *kk, *ll, *mm, *mm_st, *dkk, *dll, *area, *grids, *rhobar, *bvf;
    double *out;
};

__global__ void __launch_bounds__(NT) saturation_kernel(const SatArgs a)
{
    for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * NT) {
        double maxd;
        const double dens = a.dens[i];
        const bool hit = saturation_limit(a.p, a.p.dt, dens, a.rr[i], a.rr_st[i], a.drr[i], a.drr_st[i], a.kk[i], a.ll[i],
                                          a.mm[i], a.mm_st[i], mul(a.dkk[i], a.dll[i]), a.area[i], a.grids, a.rhobar, a.bvf, maxd);
        if (a.direct) a.out[i] = hit ? maxd : dens;                               // L:606-610
        else a.out[i] = hit ? dvd(sub(maxd, dens), a.p.dt) : 0.0;                 // L:612-615
    }
}

// ---- the driver's post-step clamp (R:182-188): saturation(dt, dens_prop, rr_old, (rr_new - rr_old) / 1, drr_old,
// (drr_new - drr_old) / dt, kk_new, ll_new, mm_old, (mm_new - mm_old) / dt, direct=True) -- bug for bug, including
// the `/ 1` of the position increment.  One kernel between two RK3 steps keeps a whole run on the device.
struct SatStepArgs {
    msgwam_params_t p;
    int64_t n;
    const double *dens, *rr0, *rr1, *drr0, *drr1, *kk, *ll, *mm0, *mm1, *dkk, *dll, *area, *grids, *rhobar, *bvf;
    double *out;
    double *rr_commit, *mm_commit;   // optional: where rr1, mm1 are copied (may alias rr0, mm0)
};

__global__ void __launch_bounds__(NT) saturation_step_kernel(const SatStepArgs a)
{
    for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * NT) {
        const double dens = a.dens[i], rr0 = a.rr0[i], drr0 = a.drr0[i], mm0 = a.mm0[i];
        const double rr1 = a.rr1[i], mm1 = a.mm1[i];
        if (a.rr_commit) a.rr_commit[i] = rr1;                                   // this thread has read rr0[i], mm0[i]
        if (a.mm_commit) a.mm_commit[i] = mm1;
        const double rr_st = sub(rr1, rr0);                                      // R:184: `/ 1` (not `/ dt`), exact
        const double drr_st = dvd(sub(a.drr1[i], drr0), a.p.dt);                 // R:185
        const double mm_st = dvd(sub(mm1, mm0), a.p.dt);                         // R:187
        double maxd;
        const bool hit = saturation_limit(a.p, a.p.dt, dens, rr0, rr_st, drr0, drr_st, a.kk[i], a.ll[i], mm0, mm_st,
                                          mul(a.dkk[i], a.dll[i]), a.area[i], a.grids, a.rhobar, a.bvf, maxd);
        a.out[i] = hit ? maxd : dens;                                            // L:606-610
    }
}

// ---- point functions ----------------------------------------------------------------------------
struct PwArgs {
    msgwam_params_t p;
    int op; int64_t n;
    const double *kk, *ll, *mm, *phi, *rr, *grid, *grids, *uu, *vv, *bvf;
    double f, f2;
    double *out;
};

__global__ void __launch_bounds__(NT) pointwise_kernel(const PwArgs a)
{
    const msgwam_params_t &p = a.p;
    const int G = p.G;
    for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * NT) {
        if (a.op == MSGWAM_OP_GRADIENTS) {
            const double rr = a.rr[i];
            double du_ray, dv_ray;
            shear_interp(rr, a.grid, a.uu, a.vv, G, p.dz_grid, p.inv_dz_grid, du_ray, dv_ray);
            a.out[i] = interp1(rr, a.grids, a.uu, G, p.inv_dz_grids);             // L:357
            a.out[a.n + i] = interp1(rr, a.grids, a.vv, G, p.inv_dz_grids);       // L:358
            a.out[2 * a.n + i] = du_ray;                                          // L:355
            a.out[3 * a.n + i] = dv_ray;                                          // L:356
            continue;
        }
        const double kk = a.kk[i], ll = a.ll[i], mm = a.mm[i];
        const double kh2 = add(mul(kk, kk), mul(ll, ll)), m2 = mul(mm, mm), vk = add(kh2, m2);
        double sphi = 0.0, cphi = 1.0, f2;
        if (a.op == MSGWAM_OP_OMEGA_F) f2 = a.f2;
        else {
            const double phi = a.phi[i];
            sphi = sin(phi); cphi = cos(phi);
            const double ff = mul(p.two_rot, sphi);
            f2 = mul(ff, ff);
        }
        const double n2 = (a.bvf != nullptr) ? n2_at(a.bvf, a.grids, G, p.inv_dz_grids, p.n2, a.rr[i]) : p.n2;   // ext
        const double om = omega_from(kh2, m2, f2, n2);
        if (a.op == MSGWAM_OP_OMEGA || a.op == MSGWAM_OP_OMEGA_F) { a.out[i] = om; continue; }
        const double cgr = dvd(dvd(mul(-mm, sub(mul(om, om), f2)), om), vk);
        if (a.op == MSGWAM_OP_CG_RR) { a.out[i] = cgr; continue; }
        const double rr = a.rr[i];
        double cgl = 0.0, cgp = 0.0;
        if (p.hprop) {
            const double nd = sub(n2, mul(om, om));
            cgl = add(mul(dvd(dvd(kk, om), vk), nd), interp1(rr, a.grids, a.uu, G, p.inv_dz_grids));
            cgp = add(mul(dvd(dvd(ll, om), vk), nd), interp1(rr, a.grids, a.vv, G, p.inv_dz_grids));
        }
        if (a.op == MSGWAM_OP_CG_LAMBDA) { a.out[i] = cgl; continue; }
        if (a.op == MSGWAM_OP_CG_PHI) { a.out[i] = cgp; continue; }
        const double rad = add(p.rad_earth, rr);
        if (a.op == MSGWAM_OP_DM_DT) {
            double du_ray, dv_ray;
            shear_interp(rr, a.grid, a.uu, a.vv, G, p.dz_grid, p.inv_dz_grid, du_ray, dv_ray);
            double dm = sub(dvd(add(mul(kk, cgl), mul(ll, cgp)), rad), add(mul(kk, du_ray), mul(ll, dv_ray)));
            if (a.bvf != nullptr) {                                               // ext
                const double nr = interp1(rr, a.grids, a.bvf, G, p.inv_dz_grids);
                double dnr, unused;
                shear_interp(rr, a.grid, a.bvf, a.bvf, G, p.dz_grid, p.inv_dz_grid, dnr, unused);
                dm = sub(dm, dvd(dvd(mul(mul(nr, dnr), kh2), om), vk));
            }
            a.out[i] = dm;
            continue;
        }
        if (!p.hprop) { a.out[i] = 0.0; continue; }                               // L:470-471, 498-499
        const double tphi = tan(a.phi[i]);
        const double zero = add(mul(kk, 0.0), mul(ll, 0.0));
        if (a.op == MSGWAM_OP_DK_DT) {
            a.out[i] = sub(mul(dvd(kk, rad), sub(mul(tphi, cgp), cgr)), dvd(dvd(zero, rad), cphi));
        } else {   // DL_DT
            const double df2 = mul(mul(mul(p.c8rot2, sphi), cphi), 1.0);
            const double t3 = mul(dvd(dvd(dvd(m2, 2.0), om), vk), df2);
            const double sum = add(add(mul(ll, cgr), mul(mul(kk, tphi), cgl)), t3);
            a.out[i] = sub(dvd(-sum, rad), dvd(zero, rad));
        }
    }
}

int g_sms = 0, g_smem = 0;
int props()
{
    return msgwam_device_info(&g_sms, &g_smem);      // of the current device (cached per device in column_step.cu)
}

inline int grid_for(int64_t n, int threads, int per_sm)
{
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)g_sms * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

int launch_project(ProjArgs &q, int64_t n, cudaStream_t s)
{
    const size_t win_bytes = ((size_t)(NT / 32) * WIN_DOUBLES + 2) * sizeof(double);
    const size_t full = win_bytes + (size_t)(q.ng + 2 * (q.ng - 1)) * sizeof(double);
    q.use_smem = full <= (size_t)g_smem;
    const size_t bytes = q.use_smem ? full : win_bytes;
    static bool configured_dev[MW_MAX_DEVICES] = {};       // cudaFuncSetAttribute is per device
    bool &configured = configured_dev[mw_current_device()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    // few, fat CTAs: each one carries a private histogram that is merged with global atomics at the end
    project_kernel<<<grid_for(n, NT * 8, 2), NT, bytes, s>>>(q);
    return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

int msgwam_rhs_rays(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n, const msgwam_grid_t *grid,
                    const double *d_uu, const double *d_vv, double *const d_tend[9], double *d_proj, void *stream)
{
    if (!p || !rays || !grid || !d_uu || !d_vv || !d_tend || !d_proj || n < 0) return MSGWAM_E_BADARG;
    if (p->G < 3) return MSGWAM_E_GRID_SIZE;
    int rc = props();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (n > 0) {
        RhsArgs a{};
        a.p = *p; a.r = *rays; a.n = n;
        a.grid = grid->grid; a.grids = grid->grids; a.rhobar = grid->rhobar; a.uu = d_uu; a.vv = d_vv; a.bvf = grid->bvf;
        for (int f = 0; f < 9; ++f) { if (!d_tend[f]) return MSGWAM_E_BADARG; a.tend[f] = d_tend[f]; }
        rhs_rays_kernel<<<grid_for(n, NT, 8), NT, 0, s>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    // wave_projection(dens, lam, phi, rr -/+ .5*drr, kk, ll, mm -/+ .5*dmm, dkk, dll, dmm, grids)  L:654-658
    ProjArgs q{};
    q.p = *p; q.var = 0; q.centered = 1; q.n = n;
    q.dens = rays->dens; q.phi = rays->phi; q.ra = rays->rr; q.rb = rays->drr; q.kk = rays->kk; q.ll = rays->ll;
    q.ma = rays->mm; q.mb = rays->dmm; q.dkk = rays->dkk; q.dll = rays->dll; q.dmm = rays->dmm;
    q.grid = grid->grids; q.ng = p->G; q.dz = p->dz_grids; q.rdz = p->inv_dz_grids; q.out = d_proj;
    q.bvf = grid->bvf; q.bvf_grids = grid->grids;
    if (n > 0) {
        return launch_project(q, n, s);
    }
    return 0;
}

int msgwam_rk_stage_rays(int32_t stage, const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                         const msgwam_grid_t *grid, const double *d_uu, const double *d_vv, double *const d_q[9],
                         double *const d_x_out[9], double *d_proj, void *stream)
{
    if (!p || !rays || !grid || !d_uu || !d_vv || !d_q || !d_x_out || !d_proj || n < 0 || stage < 0 || stage > 2)
        return MSGWAM_E_BADARG;
    if (p->G < 3) return MSGWAM_E_GRID_SIZE;
    if (n == 0) return 0;
    int rc = props();
    if (rc) return rc;
    if (!rays->dens || !rays->lam || !rays->phi || !rays->rr || !rays->drr || !rays->kk || !rays->ll || !rays->mm || !rays->dmm ||
        !rays->dkk || !rays->dll || !rays->rr_mm_area || !grid->grid || !grid->grids || !grid->rhobar)
        return MSGWAM_E_BADARG;
    StageArgs a{};
    a.r.p = *p; a.r.r = *rays; a.r.n = n;
    a.r.grid = grid->grid; a.r.grids = grid->grids; a.r.rhobar = grid->rhobar; a.r.uu = d_uu; a.r.vv = d_vv; a.r.bvf = grid->bvf;
    a.stage = stage; a.proj = d_proj;
    for (int f = 0; f < 9; ++f) {
        if (!d_q[f] || !d_x_out[f]) return MSGWAM_E_BADARG;
        a.q[f] = d_q[f]; a.xo[f] = d_x_out[f];
    }
    const size_t win_bytes = ((size_t)(NT / 32) * WIN_DOUBLES + 2) * sizeof(double);
    const size_t full = win_bytes + (size_t)(p->G + 2 * (p->G - 1)) * sizeof(double);
    a.use_smem = full <= (size_t)g_smem;
    static bool configured_dev[MW_MAX_DEVICES] = {};
    bool &configured = configured_dev[mw_current_device()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(stage_rays_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    stage_rays_kernel<<<grid_for(n, NT * 8, MSGWAM_STAGE_CTAS), NT, a.use_smem ? full : win_bytes, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

int msgwam_rk_stage_grid(int32_t stage, const msgwam_params_t *p, const msgwam_grid_t *grid, const double *d_uu,
                         const double *d_vv, double *d_proj, double *d_qu, double *d_qv, double *d_uu_out, double *d_vv_out,
                         void *stream)
{
    if (!p || !grid || !d_uu || !d_vv || !d_proj || !d_qu || !d_qv || !d_uu_out || !d_vv_out || stage < 0 || stage > 2 ||
        !grid->rhobar || !grid->pg)
        return MSGWAM_E_BADARG;
    if (p->G < 3) return MSGWAM_E_GRID_SIZE;
    stage_grid_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(stage, *p, grid->rhobar, grid->pg, d_uu, d_vv, d_proj, d_qu, d_qv,
                                                           d_uu_out, d_vv_out);
    return (int)cudaGetLastError();
}

int msgwam_grid_tendency(const msgwam_params_t *p, const msgwam_grid_t *grid, const double *d_uu, const double *d_vv,
                         const double *d_proj, double *d_du, double *d_dv, void *stream)
{
    if (!p || !grid || !d_uu || !d_vv || !d_proj || !d_du || !d_dv) return MSGWAM_E_BADARG;
    if (p->G < 3) return MSGWAM_E_GRID_SIZE;
    grid_tendency_kernel<<<(p->G + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, grid->rhobar, grid->pg, d_uu, d_vv,
                                                                               d_proj, d_du, d_dv);
    return (int)cudaGetLastError();
}

int msgwam_mean_flow_tendency(int32_t which, double f0, int32_t G, const double *d_wind, const double *d_flux_gradient,
                              const double *d_rhobar, const double *d_pg, double *d_out, void *stream)
{
    if (which < 0 || which > 1 || G < 0) return MSGWAM_E_BADARG;
    if (G == 0) return 0;
    if (!d_wind || !d_flux_gradient || !d_rhobar || !d_pg || !d_out) return MSGWAM_E_BADARG;
    mean_flow_tendency_kernel<<<(G + 255) / 256, 256, 0, (cudaStream_t)stream>>>(which, f0, G, d_wind, d_flux_gradient,
                                                                                 d_rhobar, d_pg, d_out);
    return (int)cudaGetLastError();
}

int msgwam_rk_update(int32_t stage, double dt, const double *d_tend, double *d_q, const double *d_x, double *d_x_out,
                     int64_t n, void *stream)
{
    if (stage < 0 || stage > 2 || n < 0 || (n > 0 && (!d_tend || !d_q || !d_x || !d_x_out))) return MSGWAM_E_BADARG;
    if (n == 0) return 0;
    int rc = props();
    if (rc) return rc;
    rk_update_kernel<<<grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(stage, dt, d_tend, d_q, d_x, d_x_out, n);
    return (int)cudaGetLastError();
}

int msgwam_wave_projection(int32_t var, const msgwam_params_t *p, int64_t n, const double *d_dens, const double *d_phi,
                           const double *d_rr_low, const double *d_rr_up, const double *d_kk, const double *d_ll,
                           const double *d_mm_low, const double *d_mm_up, const double *d_dkk, const double *d_dll,
                           const double *d_dmm, const double *d_grid, int32_t ng, double dz, double inv_dz,
                           const double *d_bvf, const double *d_bvf_grids, double *d_out, void *stream)
{
    if (!p || var < 0 || var > 4 || n < 0 || ng < 3 || !d_grid || !d_out) return MSGWAM_E_BADARG;
    if (n > 0 && (!d_dens || !d_phi || !d_rr_low || !d_rr_up || !d_kk || !d_ll || !d_mm_low || !d_mm_up || !d_dkk ||
                  !d_dll || !d_dmm))
        return MSGWAM_E_BADARG;
    int rc = props();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t outlen = var == 0 ? 2 * (size_t)(ng - 1) : var == 4 ? 2 * (size_t)ng : var == 3 ? (size_t)ng : (size_t)(ng - 1);
    cudaError_t e = cudaMemsetAsync(d_out, 0, outlen * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    if (n == 0) return 0;
    ProjArgs q{};
    q.p = *p; q.var = var; q.centered = 0; q.n = n;
    q.dens = d_dens; q.phi = d_phi; q.ra = d_rr_low; q.rb = d_rr_up; q.kk = d_kk; q.ll = d_ll;
    q.ma = d_mm_low; q.mb = d_mm_up; q.dkk = d_dkk; q.dll = d_dll; q.dmm = d_dmm;
    q.grid = d_grid; q.ng = ng; q.out = d_out;
    q.bvf = d_bvf; q.bvf_grids = d_bvf_grids;
    if ((d_bvf == nullptr) != (d_bvf_grids == nullptr)) return MSGWAM_E_BADARG;
    q.dz = dz; q.rdz = inv_dz;
    if (var >= 3) {
        project_iface_kernel<<<grid_for(n, NT, 8), NT, 0, s>>>(q);
        return (int)cudaGetLastError();
    }
    return launch_project(q, n, s);
}

int msgwam_saturation(const msgwam_params_t *p, int64_t n, int32_t direct, const double *d_dens, const double *d_rr,
                      const double *d_rr_st, const double *d_drr, const double *d_drr_st, const double *d_kk,
                      const double *d_ll, const double *d_mm, const double *d_mm_st, const double *d_dkk,
                      const double *d_dll, const double *d_area, const double *d_grids, const double *d_rhobar,
                      const double *d_bvf, double *d_out, void *stream)
{
    if (!p || n < 0) return MSGWAM_E_BADARG;
    if (n == 0) return 0;
    if (!d_dens || !d_rr || !d_rr_st || !d_drr || !d_drr_st || !d_kk || !d_ll || !d_mm || !d_mm_st || !d_dkk || !d_dll ||
        !d_area || !d_grids || !d_rhobar || !d_out)
        return MSGWAM_E_BADARG;
    int rc = props();
    if (rc) return rc;
    SatArgs a{*p, n, direct, d_dens, d_rr, d_rr_st, d_drr, d_drr_st, d_kk, d_ll, d_mm, d_mm_st, d_dkk, d_dll, d_area,
              d_grids, d_rhobar, d_bvf, d_out};
    saturation_kernel<<<grid_for(n, NT, 8), NT, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

int msgwam_saturation_step(const msgwam_params_t *p, int64_t n, const double *d_dens, const double *d_rr_old,
                           const double *d_rr_new, const double *d_drr_old, const double *d_drr_new, const double *d_kk,
                           const double *d_ll, const double *d_mm_old, const double *d_mm_new, const double *d_dkk,
                           const double *d_dll, const double *d_area, const double *d_grids, const double *d_rhobar,
                           const double *d_bvf, double *d_dens_out, void *stream)
{
    return msgwam_saturation_step_commit(p, n, d_dens, d_rr_old, d_rr_new, d_drr_old, d_drr_new, d_kk, d_ll, d_mm_old, d_mm_new,
                                         d_dkk, d_dll, d_area, d_grids, d_rhobar, d_bvf, d_dens_out, nullptr, nullptr, stream);
}

// The clamp moves 11 fields in and 3 out per ray (112 B): ~17 us per 1e6 rays at HBM speed, measured ~25 us.
int msgwam_saturation_step_commit(const msgwam_params_t *p, int64_t n, const double *d_dens, const double *d_rr_old,
                                  const double *d_rr_new, const double *d_drr_old, const double *d_drr_new, const double *d_kk,
                                  const double *d_ll, const double *d_mm_old, const double *d_mm_new, const double *d_dkk,
                                  const double *d_dll, const double *d_area, const double *d_grids, const double *d_rhobar,
                                  const double *d_bvf, double *d_dens_out, double *d_rr_commit, double *d_mm_commit, void *stream)
{
    if (!p || n < 0) return MSGWAM_E_BADARG;
    if (n == 0) return 0;
    if (!d_dens || !d_rr_old || !d_rr_new || !d_drr_old || !d_drr_new || !d_kk || !d_ll || !d_mm_old || !d_mm_new || !d_dkk ||
        !d_dll || !d_area || !d_grids || !d_rhobar || !d_dens_out)
        return MSGWAM_E_BADARG;
    int rc = props();
    if (rc) return rc;
    SatStepArgs a{*p, n, d_dens, d_rr_old, d_rr_new, d_drr_old, d_drr_new, d_kk, d_ll, d_mm_old, d_mm_new, d_dkk, d_dll,
                  d_area, d_grids, d_rhobar, d_bvf, d_dens_out, d_rr_commit, d_mm_commit};
    saturation_step_kernel<<<grid_for(n, NT, 8), NT, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

int msgwam_pointwise(int32_t op, const msgwam_params_t *p, int64_t n, const double *d_kk, const double *d_ll,
                     const double *d_mm, const double *d_phi, const double *d_rr, double f, double f2,
                     const msgwam_grid_t *grid, const double *d_uu, const double *d_vv, double *d_out, void *stream)
{
    if (!p || op < 0 || op > MSGWAM_OP_GRADIENTS || n < 0) return MSGWAM_E_BADARG;
    if (n == 0) return 0;
    if (!d_out) return MSGWAM_E_BADARG;
    const bool needs_wave = op != MSGWAM_OP_GRADIENTS;
    const bool needs_phi = needs_wave && op != MSGWAM_OP_OMEGA_F;
    const bool needs_pos = op >= MSGWAM_OP_CG_LAMBDA;
    if ((needs_wave && (!d_kk || !d_ll || !d_mm)) || (needs_phi && !d_phi)) return MSGWAM_E_BADARG;
    if (needs_pos && (!d_rr || !grid || !grid->grid || !grid->grids || !d_uu || !d_vv || p->G < 3)) return MSGWAM_E_BADARG;
    int rc = props();
    if (rc) return rc;
    PwArgs a{};
    a.p = *p; a.op = op; a.n = n; a.kk = d_kk; a.ll = d_ll; a.mm = d_mm; a.phi = d_phi; a.rr = d_rr;
    if (grid) { a.grid = grid->grid; a.grids = grid->grids; a.bvf = grid->bvf; }
    if (a.bvf && (!d_rr || !a.grids || !a.grid || p->G < 3)) return MSGWAM_E_BADARG;     // the profile needs a position
    a.uu = d_uu; a.vv = d_vv; a.f = f; a.f2 = f2; a.out = d_out;
    pointwise_kernel<<<grid_for(n, NT, 8), NT, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

}  // extern "C"
