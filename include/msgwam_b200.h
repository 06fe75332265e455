/*
 * msgwam_b200.h -- C ABI of the B200-native replacement for python-msgwam's hot path
 * (Runge-Kutta stepping of the ray-volume equations + pseudo-momentum-flux deposition).
 *
 * The reference has no FFI: its "plugin API" is the Python module namespace of
 * /root/reference/lib/libprop.py (imported as `lprop` at /root/reference/raytracer.py:2).
 * Every entry point below names the reference callable(s) it replaces (L:nnn =
 * lib/libprop.py line, R:nnn = raytracer.py line).  The ctypes binding a maintainer
 * adds on the reference side is shown in INTEGRATION.md; the shipped binding is
 * python-msgwam_b200/msgwam_b200/_cabi.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  All floating point data is IEEE binary64.
 *   - Pointers named d_* / inside msgwam_rays_t / msgwam_grid_t are DEVICE pointers owned by the
 *     caller (torch tensors in the shipped host code); h_* are HOST pointers.
 *   - No hidden allocations: scratch memory is passed in; its size comes from the *_bytes helpers.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Nothing synchronises
 *     the stream unless stated.
 *   - Return value: 0 = ok, > 0 = cudaError_t of the failing runtime call / launch,
 *     < 0 = MSGWAM_E_* argument error.  msgwam_error_string() explains either.
 *   - Scalars that the reference derives with Python-float arithmetic (bvf**2, 2*ROT_EARTH,
 *     2*ROT_EARTH*sin(phi0), kappa**2*.5, np.diff(grid[:2])[0] ...) are derived by the host
 *     binding with the same expressions and arrive in msgwam_params_t, so that no libm/pow
 *     difference can enter (see _cabi.snapshot_params()).
 */
#ifndef MSGWAM_B200_H
#define MSGWAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSGWAM_ABI_VERSION 5   /* 5: msgwam_rays_t.bounds (fixed-point CTA histogram); N(z) host entry point; peer timeout */

#define MSGWAM_E_BADARG      (-1)   /* null pointer / negative size / inconsistent arguments   */
#define MSGWAM_E_GRID_SIZE   (-2)   /* G too small (< 3) or too large for the fused column kernels */
#define MSGWAM_E_UNSUPPORTED (-3)   /* mode not handled by this entry point                     */
#define MSGWAM_E_TIMEOUT     (-4)   /* a bounded device-side wait ran out (grid-wide arrival of the mean-flow slices, or a peer
                                       in the NVLink all-reduce); the step's results are invalid          */

/* model_config / module globals read on the hot path (SURVEY.md section 5, L:3-11, 380-383, 534, 582-584, 633) */
typedef struct msgwam_params {
    double dt;          /* time step passed to RK3 / rhs_default / saturation                     */
    double n2;          /* model_config['bvf'] ** 2                                L:383           */
    double two_rot;     /* 2 * ROT_EARTH                                           L:382           */
    double rad_earth;   /* RAD_EARTH                                               L:3             */
    double c8rot2;      /* 8 * ROT_EARTH**2                                        L:491           */
    double f0;          /* 2 * ROT_EARTH * np.sin(model_config['phi0'])            L:535, 554, 589 */
    double f0sq;        /* f0 ** 2                                                 L:597, 601      */
    double k2half;      /* model_config['kappa']**2 * .5                           L:601           */
    double dz_grid;     /* np.diff(grid[:2])[0]                                    L:349, 662      */
    double dz_grids;    /* np.diff(grids[:2])[0]  (the deposit receives `grids`)   L:123, 657      */
    double inv_dz_grid; /* 1.0 / dz_grid  (correctly rounded; used by exact-division-by-invariant)  */
    double inv_dz_grids;/* 1.0 / dz_grids                                                           */
    int32_t G;          /* len(grids) == len(grid) - 1                                              */
    int32_t hprop;      /* HPROP_GLOBAL                                            L:5             */
    int32_t saturate_online; /* model_config['saturate_online']                    L:633           */
    int32_t reserved;
} msgwam_params_t;

/* Structure-of-arrays ray store: the 9 state slots of the reference's state vector
 * (R:160-172, L:629), the per-ray statics (L:630-632) and two derived statics. */
typedef struct msgwam_rays {
    const double *dens, *lam, *phi, *rr, *drr, *kk, *ll, *mm, *dmm;
    const double *dkk, *dll, *rr_mm_area;
    const double *ff;    /* 2*ROT_EARTH*sin(phi)   (msgwam_derive_statics; column mode only) */
    const double *pkl;   /* dkk*dll                (msgwam_derive_statics; column mode only) */
    double *stage1;      /* column mode only: 3*n doubles of scratch.  Pass A leaves the stage-1 increments and the
                            group velocity of state r1 there (dt*cg_rr(r0) | dt*dm_dt(r0) | cg_rr(r1)) so that
                            pass B starts at RK stage 2 instead of recomputing stage 1.                           */
    double *bounds;      /* column mode only, may be NULL: 16 doubles of per-ensemble state that live next to a ray store
                            which is advanced IN PLACE.  [0..5] = for the deposits D0, D1, D2 of the previous step and
                            each of their two flux components, the max over CTAs of the sum of |contribution| over the
                            CTA's rays (D0 and D1, the two deposits of the first sweep, share their scales -- the larger
                            of their bounds counts -- and are measured together, half of the sum to each);
                            [6..11] = the same, being gathered by the running step; [12] = 1.0 when [0..5] are
                            valid (zero the 16 doubles when the store is created or edited from outside, or call
                            msgwam_column_bounds).  With valid bounds the CTA histogram of the deposit accumulates in
                            64-bit fixed point with native integer atomics (deposit.cuh); NULL or invalid: fp64
                            compare-and-swap atomics.  Stale bounds cost precision and speed, never correctness: a thread
                            adds to the histogram only while the running sum of its scaled contributions is below 2^62 /
                            (threads per CTA), so no accumulator can overflow; past that, and for non-finite rays,
                            contributions go to the global deposit in fp64, and so do non-zero contributions of a
                            component whose bound was exactly zero.                                                  */
} msgwam_rays_t;

/* Background profiles on the 1-D mean-flow grid (L:6-9). */
typedef struct msgwam_grid {
    const double *grid;     /* (G+1,) cell edges                 lprop.grid              */
    const double *grids;    /* (G,)   staggered grid             lprop.grids             */
    const double *rhobar;   /* (G,)                              lprop.rhobar            */
    const double *pg;       /* (2,G)                             lprop.pressure_gradient */
    const double *bvf;      /* (G,) or NULL.  EXTENSION, not in the reference: buoyancy
                               frequency N on `grids`; N^2(z) = np.interp(z, grids, bvf)**2 replaces the scalar
                               params.n2 at the position argument of cg_rr / cg_lambda / cg_phi, and dm_dt gains
                               -N N'(k^2+l^2)/om/|k|^2.  Only the general (stage-by-stage) entry points use it. */
} msgwam_grid_t;

int         msgwam_abi_version(void);
const char *msgwam_error_string(int code);
/* number of SMs / max dynamic shared memory of device 0 as seen by the library (diagnostics) */
int         msgwam_device_info(int *sm_count, int *max_smem_optin);

/* ---- derived statics ------------------------------------------------------------------------
 * ff = 2*ROT_EARTH*sin(phi) (L:382, 446) and pkl = dkk*dll (first product of L:137), once per
 * upload; phi is constant while HPROP_GLOBAL is False, dkk/dll are statics. */
int msgwam_derive_statics(const double *d_phi, const double *d_dkk, const double *d_dll,
                          double *d_ff, double *d_pkl, int64_t n, double two_rot, void *stream);

/* ---- fused column-mode RK3 step  (replaces RK3 L:680-700 with rhs = rhs_default L:618-676,
 *      for HPROP_GLOBAL == False and saturate_online == False: only rr and mm change) --------
 *
 * The mean flow is part of the RK state, so stage s+1 needs the deposit of ALL rays at stage s:
 *   pass A : deposit D(r0); stage 1 with u0; deposit D(r1); stage-1 increments and cg_rr(r1) -> rays->stage1
 *            (reads 9 fields, writes 3)
 *   [multi-GPU: all-reduce D0|D1 here -- they are contiguous: 4*(G-1) doubles]
 *   pass B : prologue = the mean-flow half of RK stages 1-2 on the reduced D0, D1, distributed over the CTAs
 *            (u1, u2 and the shear tables the sweep interpolates); then per ray r1 = r0 + stage-1 increment
 *            (from rays->stage1), stage 2 with u1, deposit D(r2), stage 3 with u2; write rr, mm
 *            (reads 12 fields, writes 2)
 *   [multi-GPU: all-reduce D2: 2*(G-1) doubles]
 *   finish : u3, v3 from u2 and D2; zeroes the deposit buffers for the next step
 * d_work: msgwam_column_work_doubles(G) doubles, zero-initialised by the caller once; layout
 * D0 (2,G-1) | D1 (2,G-1) | D2 (2,G-1) | shear tables and saved mean-flow stage | counters (internal).
 * rr_out/mm_out may alias rays->rr / rays->mm.  msgwam_column_pass_b must follow msgwam_column_pass_a on the same
 * stream (pass A arms the counter pass B's CTAs meet on and fills rays->stage1).
 */
int64_t msgwam_column_work_doubles(int32_t G);
/* largest G (= len(grids)) the fused column kernels accept on this device: the shear tables of the three
 * RK stages must fit in shared memory.  Larger grids take msgwam_rhs_rays stage by stage. */
int32_t msgwam_column_max_levels(void);
int msgwam_column_pass_a(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                         const msgwam_grid_t *grid, const double *d_uu, const double *d_vv,
                         double *d_work, void *stream);
int msgwam_column_pass_b(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                         const msgwam_grid_t *grid, const double *d_uu, const double *d_vv,
                         double *d_work, double *d_rr_out, double *d_mm_out, void *stream);
int msgwam_column_finish(const msgwam_params_t *p, const msgwam_grid_t *grid,
                         const double *d_uu, const double *d_vv, double *d_work,
                         double *d_uu_out, double *d_vv_out, void *stream);
/* Multi-GPU without a separate collective: the one-CTA chain / finish kernels all-reduce the deposit
 * themselves with one-shot pushes over NVLink peer memory.  inbox[r] is rank r's inbox buffer
 * (msgwam_p2p_inbox_doubles(G, world) doubles, zero-initialised once, allocated in symmetric / IPC memory)
 * as mapped into THIS process; epoch must be identical on all ranks and increase by one per reduction
 * (pass_b_p2p and finish_p2p each perform one).  Every value travels as a 16-byte cell {low word, flag, high word,
 * flag} with flag = epoch | 2^31, so a receiver polls the data itself (2 parities x world slots x 4 (G - 1) cells per
 * inbox).  A peer that does not answer within ~20 s sets the error word returned by msgwam_column_error()
 * instead of hanging the GPU. */
#define MSGWAM_MAX_PEERS 16
typedef struct msgwam_peers {
    int32_t world, rank;
    uint64_t epoch;
    void *inbox[MSGWAM_MAX_PEERS];
} msgwam_peers_t;
int64_t msgwam_p2p_inbox_doubles(int32_t G, int32_t world);
int msgwam_column_pass_b_p2p(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                             const msgwam_grid_t *grid, const double *d_uu, const double *d_vv,
                             double *d_work, double *d_rr_out, double *d_mm_out,
                             const msgwam_peers_t *peers, void *stream);
int msgwam_column_finish_p2p(const msgwam_params_t *p, const msgwam_grid_t *grid,
                             const double *d_uu, const double *d_vv, double *d_work,
                             double *d_uu_out, double *d_vv_out, const msgwam_peers_t *peers, void *stream);
/* Multi-GPU, fully fused: the same two launches as msgwam_column_step (epochs peers->epoch and peers->epoch + 1:
 * advance the epoch by TWO per call).  The last CTA of pass A pushes this GPU's D0 | D1 to every inbox and retires;
 * each CTA of pass B polls its own inbox for the cells its slice of the mean-flow chain reads and sums them in rank
 * order.  The last CTA of pass B pushes D2, collects the sums and runs the finish.  d_work's D0 | D1 hold this
 * GPU's partial sums only. */
int msgwam_column_step_p2p(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                           const msgwam_grid_t *grid, const double *d_uu, const double *d_vv,
                           double *d_work, double *d_rr_out, double *d_mm_out,
                           double *d_uu_out, double *d_vv_out, const msgwam_peers_t *peers, void *stream);
/* N(z) EXTENSION (no counterpart in the reference): the same two launches with a buoyancy-frequency profile grid->bvf
 * (N on grids).  N^2 differs between the centre and the edges of a ray volume, so rr, drr, mm and dmm all evolve;
 * rays->stage1 must hold 7 * n doubles.  G <= msgwam_column_nz_max_levels().  With `peers` (several GPUs) the
 * all-reduces of the deposit run inside the sweeps, epochs peers->epoch and peers->epoch + 1, as in
 * msgwam_column_step_p2p. */
int msgwam_column_step_nz(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                          const msgwam_grid_t *grid, const double *d_uu, const double *d_vv, double *d_work,
                          double *d_rr_out, double *d_drr_out, double *d_mm_out, double *d_dmm_out,
                          double *d_uu_out, double *d_vv_out, const msgwam_peers_t *peers /* NULL: one GPU */,
                          void *stream);
int32_t msgwam_column_nz_max_levels(void);
/* offset (in doubles) inside d_work of the error word set by a timed-out peer exchange (0.0 = ok) */
int64_t msgwam_column_error_offset(int32_t G);
/* FROZEN-BACKGROUND MODE "M2" -- an extension under its own name, NOT the reference's RK3 + rhs_default (whose mean
 * flow is part of the RK state, L:629, 665-674): all three RK stages of a ray in registers with uu, vv frozen over the
 * step; one deposit, of the state at the END of the step; uu += dt * du_dt(vv, dF/dz), vv += dt * dv_dt(uu, dF/dz) once
 * per step (L:523-558, 653-663).  The ray half is exactly the reference's RK3 (L:680-700) with model_config['rhs'] (L:691)
 * set to rhs_default with du_st = dv_st = 0.  One launch per step; constant N, HPROP off, saturate_online off.
 * peers: NULL on one GPU, else one reduction (epoch peers->epoch) of the deposit in the tail of the sweep. */
int msgwam_column_step_frozen(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                              const msgwam_grid_t *grid, const double *d_uu, const double *d_vv, double *d_work,
                              double *d_rr_out, double *d_mm_out, double *d_uu_out, double *d_vv_out,
                              const msgwam_peers_t *peers, void *stream);
/* The reference driver's loop body (R:175-188) as one call -- the RK3 step (msgwam_column_step / _p2p / _nz) with the
 * post-step clamp  dens <- saturation(dt, dens, rr_old, (rr_new - rr_old) / 1, drr_old, (drr_new - drr_old) / dt, kk,
 * ll, mm_old, (mm_new - mm_old) / dt, direct=True)  (L:561-610, bug for bug including R:184's `/ 1`) fused into the end
 * of pass B, where both ends of the step are in registers.  d_dens_out may alias rays->dens (then only clamped rays
 * are written); rays->rr_mm_area is read; peers: NULL on one GPU. */
int msgwam_column_advance(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                          const msgwam_grid_t *grid, const double *d_uu, const double *d_vv, double *d_work,
                          double *d_rr_out, double *d_mm_out, double *d_dens_out,
                          double *d_uu_out, double *d_vv_out, const msgwam_peers_t *peers, void *stream);
int msgwam_column_advance_nz(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                             const msgwam_grid_t *grid, const double *d_uu, const double *d_vv, double *d_work,
                             double *d_rr_out, double *d_drr_out, double *d_mm_out, double *d_dmm_out,
                             double *d_dens_out, double *d_uu_out, double *d_vv_out,
                             const msgwam_peers_t *peers /* NULL: one GPU */, void *stream);
/* Deposit bounds of a ray store whose bounds are unknown (a new store, a store edited from outside): one cheap sweep
 * sets rays->bounds[0..5] to the bounds of wave_projection(var = 0) at the current state (L:137-149: max over CTAs of the
 * sums of |dkk dll dmm cg k dens| and |dkk dll dmm cg l dens| over the CTA's rays), [6..11] to zero and [12] to 1.  The column step that
 * follows then accumulates its deposits in fixed point (see msgwam_rays_t.bounds) and measures the bounds for the
 * step after it.  grid->bvf != NULL: for msgwam_column_step_nz. */
int msgwam_column_bounds(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                         const msgwam_grid_t *grid, void *stream);
/* bound of the device-side polls of the peer exchange in seconds (default 120): a rank that is merely late must
 * never reach it; when it fires the error word is set and the host binding raises at its next synchronisation */
int msgwam_set_peer_timeout(double seconds);
/* measurement hook: while `event` (a cudaEvent_t) is set, every fused column step records it between its two
 * launches so that the sweeps can be timed separately; NULL switches it off */
int msgwam_debug_mid_event(void *event);
/* test hook: launch every sweep with mult (1..8) CTAs per SM instead of one.  A CTA of a sweep fills an SM, so with
 * mult > 1 most CTAs of a grid are not resident while the first ones run -- what kernels of other streams or MPS
 * clients do to a grid -- and the step must complete with the same results: no kernel of this library waits on a
 * CTA that has not started (the mean-flow chain hands its slices out by ticket).  1 = product configuration. */
int msgwam_debug_grid_mult(int mult);

/* single GPU: two launches -- the mean-flow chain and the finish run as the tails of the sweeps, in the last
 * CTA to retire */
int msgwam_column_step(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                       const msgwam_grid_t *grid, const double *d_uu, const double *d_vv,
                       double *d_work, double *d_rr_out, double *d_mm_out,
                       double *d_uu_out, double *d_vv_out, void *stream);

/* test hook: cg_rr exactly as the fused column kernels evaluate it (shared reciprocals, one range check),
 * so that tests can compare it bit for bit with the library route used by msgwam_pointwise */
int msgwam_debug_cg_rr_fast(const double *d_kk, const double *d_ll, const double *d_mm, const double *d_ff,
                            double n2, double *d_out, int64_t n, void *stream);

/* ---- general single-stage right-hand side  (replaces rhs_default L:618-676, every branch:
 *      HPROP on/off, saturate_online on/off).  d_tend[9] receive the nine ray tendencies in
 *      state-vector order, d_proj (2,G-1) the deposit of wave_projection(var=0) (must be zero
 *      on entry), then msgwam_grid_tendency turns the deposit into du_st, dv_st (L:659-666). */
int msgwam_rhs_rays(const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                    const msgwam_grid_t *grid, const double *d_uu, const double *d_vv,
                    double *const d_tend[9], double *d_proj, void *stream);
int msgwam_grid_tendency(const msgwam_params_t *p, const msgwam_grid_t *grid,
                         const double *d_uu, const double *d_vv, const double *d_proj,
                         double *d_du, double *d_dv, void *stream);
/* du_dt (which = 0, L:523-539: f0*wind - (pg + flux_gradient)/rhobar) and dv_dt (which = 1, L:542-558:
 * -f0*wind - ...) from an already differentiated flux; d_pg is the matching row of pressure_gradient. */
int msgwam_mean_flow_tendency(int32_t which, double f0, int32_t G, const double *d_wind,
                              const double *d_flux_gradient, const double *d_rhobar, const double *d_pg,
                              double *d_out, void *stream);
/* low-storage RK update of one array (L:693-698): stage 0: q = dt*t, x_out = x + q/3;
 * stage 1,2: q = dt*t - a*q, x_out = x + b*q. */
int msgwam_rk_update(int32_t stage, double dt, const double *d_tend, double *d_q,
                     const double *d_x, double *d_x_out, int64_t n, void *stream);

/* ---- one fused RK stage for any mode (HPROP on/off, saturate_online on/off, N(z) profile) --------------------
 * msgwam_rk_stage_rays: rhs_default on the state in `rays` (L:618-651), its deposit wave_projection(var=0) of the same
 * state (L:653-658, accumulated into d_proj, which must be zero on entry) and the low-storage update of the nine ray
 * slots (L:693-698: stage 0: q = dt*k, x += q/3; stage 1, 2: q = dt*k - a*q, x += b*q) in one sweep.  d_x_out[f] may
 * alias the state in `rays`.  [multi-GPU: all-reduce d_proj here]  msgwam_rk_stage_grid: du_st, dv_st from the deposit
 * (L:659-666), the same update for uu, vv, and d_proj zeroed for the next stage.  Three such pairs are one RK3. */
int msgwam_rk_stage_rays(int32_t stage, const msgwam_params_t *p, const msgwam_rays_t *rays, int64_t n,
                         const msgwam_grid_t *grid, const double *d_uu, const double *d_vv,
                         double *const d_q[9], double *const d_x_out[9], double *d_proj, void *stream);
int msgwam_rk_stage_grid(int32_t stage, const msgwam_params_t *p, const msgwam_grid_t *grid,
                         const double *d_uu, const double *d_vv, double *d_proj,
                         double *d_qu, double *d_qv, double *d_uu_out, double *d_vv_out, void *stream);

/* ---- deposition  (replaces wave_projection L:92-221, var = 0..4, any uniform grid) ----------
 * out sizes: var 0 -> 2*(ng-1); 1,2 -> ng-1; 3 -> ng; 4 -> 2*ng doubles, zeroed by the call.
 * dz = np.diff(grid[:2])[0] of the grid passed (L:123) and inv_dz = 1.0/dz, derived by the host. */
/* d_bvf/d_bvf_grids: NULL, or the N(z) extension's profile and its abscissa (grids) -- see msgwam_grid_t.bvf */
int msgwam_wave_projection(int32_t var, const msgwam_params_t *p, int64_t n,
                           const double *d_dens, const double *d_phi,
                           const double *d_rr_low, const double *d_rr_up,
                           const double *d_kk, const double *d_ll,
                           const double *d_mm_low, const double *d_mm_up,
                           const double *d_dkk, const double *d_dll, const double *d_dmm,
                           const double *d_grid, int32_t ng, double dz, double inv_dz,
                           const double *d_bvf, const double *d_bvf_grids,
                           double *d_out, void *stream);

/* ---- saturation  (replaces saturation L:561-615; direct = 0 tendency, 1 clamp) --------------*/
int msgwam_saturation(const msgwam_params_t *p, int64_t n, int32_t direct,
                      const double *d_dens, const double *d_rr, const double *d_rr_st,
                      const double *d_drr, const double *d_drr_st,
                      const double *d_kk, const double *d_ll,
                      const double *d_mm, const double *d_mm_st,
                      const double *d_dkk, const double *d_dll, const double *d_area,
                      const double *d_grids, const double *d_rhobar, const double *d_bvf /* NULL or N on grids */,
                      double *d_out, void *stream);

/* The driver's post-step clamp, /root/reference/raytracer.py:182-188: saturation(direct=True) on the propagated wave
 * action with the step's finite-difference tendencies ((rr_new - rr_old) / 1 -- sic --, (drr_new - drr_old) / dt,
 * (mm_new - mm_old) / dt), in one kernel, so that consecutive RK3 steps need not leave the device.  p->dt = dt.
 * d_dens_out may alias d_dens. */
int msgwam_saturation_step(const msgwam_params_t *p, int64_t n, const double *d_dens,
                           const double *d_rr_old, const double *d_rr_new,
                           const double *d_drr_old, const double *d_drr_new,
                           const double *d_kk, const double *d_ll,
                           const double *d_mm_old, const double *d_mm_new,
                           const double *d_dkk, const double *d_dll, const double *d_area,
                           const double *d_grids, const double *d_rhobar, const double *d_bvf /* NULL or N on grids */,
                           double *d_dens_out, void *stream);
/* The same clamp for a driver that stepped OUT of place (rr_new, mm_new in scratch): the kernel also copies rr_new to
 * d_rr_commit and mm_new to d_mm_commit (each may be NULL, and may alias d_rr_old / d_mm_old -- the ray store), which
 * replaces the three copies of the old state a driver loop would otherwise make before every in-place step. */
int msgwam_saturation_step_commit(const msgwam_params_t *p, int64_t n, const double *d_dens,
                                  const double *d_rr_old, const double *d_rr_new,
                                  const double *d_drr_old, const double *d_drr_new,
                                  const double *d_kk, const double *d_ll,
                                  const double *d_mm_old, const double *d_mm_new,
                                  const double *d_dkk, const double *d_dll, const double *d_area,
                                  const double *d_grids, const double *d_rhobar, const double *d_bvf /* NULL or N on grids */,
                                  double *d_dens_out, double *d_rr_commit, double *d_mm_commit, void *stream);

/* ---- point functions (replace omega L:369, cg_rr L:434, cg_lambda L:386, cg_phi L:410,
 *      dk_dt L:451, dl_dt L:474, dm_dt L:502, gradients L:328) ------------------------------
 * op selects the function; unused inputs may be NULL.  For MSGWAM_OP_OMEGA_F the latitude is the
 * scalar `f`/`f2` pair (2*ROT*sin(phi), its square) instead of d_phi.  gradients writes
 * 4 arrays (uu_ray, vv_ray, du_dz_ray, dv_dz_ray) of n doubles into d_out. */
enum {
    MSGWAM_OP_OMEGA = 0, MSGWAM_OP_OMEGA_F = 1, MSGWAM_OP_CG_RR = 2, MSGWAM_OP_CG_LAMBDA = 3,
    MSGWAM_OP_CG_PHI = 4, MSGWAM_OP_DK_DT = 5, MSGWAM_OP_DL_DT = 6, MSGWAM_OP_DM_DT = 7,
    MSGWAM_OP_GRADIENTS = 8
};
int msgwam_pointwise(int32_t op, const msgwam_params_t *p, int64_t n,
                     const double *d_kk, const double *d_ll, const double *d_mm,
                     const double *d_phi, const double *d_rr /* also read by omega / cg_rr when grid->bvf is set */,
                     double f, double f2,
                     const msgwam_grid_t *grid, const double *d_uu, const double *d_vv,
                     double *d_out, void *stream);

/* ---- stream compaction of the ray store (ray deletion; SURVEY.md 7.3-6, no reference code) --
 * keep[i] != 0 keeps ray i.  nfields arrays are compacted stably from d_in[f] to d_out[f]
 * (d_out[f] != d_in[f]).  d_count receives the number of survivors (device int64).
 * d_scratch: msgwam_compact_scratch_bytes(n) bytes. */
int64_t msgwam_compact_scratch_bytes(int64_t n);
int msgwam_flag_rays(const msgwam_params_t *p, int64_t n, const double *d_rr, const double *d_drr,
                     const double *d_mm, double m_crit, uint8_t *d_keep, void *stream);
int msgwam_compact(int64_t n, const uint8_t *d_keep, int32_t nfields,
                   const double *const d_in[], double *const d_out[], int64_t *d_count,
                   void *d_scratch, void *stream);

/* ---- host-buffer entry point: the call the reference-facing shim makes for numpy inputs -----
 * Copies the step's inputs host->device, runs the column step, copies rr, mm, uu, vv back.
 * All h_* are host pointers (pinned or pageable); d_stage is a device scratch of
 * msgwam_host_stage_doubles(n, G) doubles; d_work as in msgwam_column_step.
 * h_dkk == h_dll == NULL reuses the statics a previous call with the same n left in d_stage (they are
 * per-run constants, L:722-726).  Synchronises `stream` before returning (the outputs are host memory). */
int64_t msgwam_host_stage_doubles(int64_t n, int32_t G);
int msgwam_rk3_column_host(const msgwam_params_t *p, int64_t n,
                           const double *const h_state[9], const double *h_dkk, const double *h_dll,
                           const double *h_uu, const double *h_vv,
                           const double *h_grid, const double *h_grids, const double *h_rhobar,
                           const double *h_pg,
                           double *h_rr_out, double *h_mm_out, double *h_uu_out, double *h_vv_out,
                           double *d_stage, double *d_work, void *stream);
/* the same call with the N(z) extension (h_bvf: N on grids, G values): rr, drr, mm, dmm come
 * back; d_stage holds msgwam_host_stage_doubles_nz(n, G) doubles */
int64_t msgwam_host_stage_doubles_nz(int64_t n, int32_t G);
int msgwam_rk3_column_nz_host(const msgwam_params_t *p, int64_t n,
                              const double *const h_state[9], const double *h_dkk, const double *h_dll,
                              const double *h_uu, const double *h_vv,
                              const double *h_grid, const double *h_grids, const double *h_rhobar,
                              const double *h_pg, const double *h_bvf,
                              double *h_rr_out, double *h_drr_out, double *h_mm_out, double *h_dmm_out,
                              double *h_uu_out, double *h_vv_out,
                              double *d_stage, double *d_work, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MSGWAM_B200_H */
