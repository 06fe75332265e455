// host_path.cu -- the host-buffer entry point and the ABI bookkeeping.
//
// msgwam_rk3_column_host is what the reference-facing shim calls when the caller hands numpy arrays
// to RK3 (R:160-175): every input of the step is copied host->device, the fused column step runs,
// and the changed slots (rr, mm, uu, vv) are copied back.  lam, phi, dens, drr, kk, ll, dmm keep their
// values in column mode (their tendencies are exactly zero, L:638-645 with HPROP off), so the shim
// returns copies of the inputs for those slots and they are never moved over PCIe.
#include "common.cuh"
#include <string.h>
#include <sched.h>
#include <omp.h>

extern "C" {

int msgwam_abi_version(void) { return MSGWAM_ABI_VERSION; }

const char *msgwam_error_string(int code)
{
    switch (code) {
    case 0: return "ok";
    case MSGWAM_E_BADARG: return "msgwam: bad argument (null pointer, negative size or inconsistent sizes)";
    case MSGWAM_E_GRID_SIZE: return "msgwam: grid size unsupported (G < 3, or shear tables exceed shared memory)";
    case MSGWAM_E_TIMEOUT: return "msgwam: a bounded device-side wait timed out (mean-flow slices or a peer of the all-reduce); results invalid";
    case MSGWAM_E_UNSUPPORTED: return "msgwam: mode not supported by this entry point (column kernels need HPROP off and saturate_online off)";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "msgwam: unknown error";
    }
}

static inline int64_t pad32(int64_t n) { return (n + 31) & ~(int64_t)31; }

// ---- uploads from PAGEABLE host arrays (what an unmodified driver script passes: rows of its history arrays, R:160-172)
// cudaMemcpyAsync from pageable memory is a staged, blocking copy that the driver runs on one thread (~10-15 GB/s).
// Large pageable arrays are instead copied by several host threads into one of two page-locked chunks, each of which
// goes up with an asynchronous copy while the threads fill the other one: the upload then runs at the speed of the
// host's memory system, up to the PCIe rate.  Page-locked (or registered) arrays are copied directly.
namespace {
constexpr size_t STAGE_CHUNK = (size_t)32 << 20;       // bytes per page-locked chunk
struct Stager {
    char *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int next = 0, threads = 1;
    bool ok = false;
    bool init()
    {
        if (ok) return true;
        for (int k = 0; k < 2; ++k) {
            if (cudaHostAlloc(reinterpret_cast<void **>(&buf[k]), STAGE_CHUNK, cudaHostAllocPortable) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        cpu_set_t set;
        int n = 1;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);     // not OMP_NUM_THREADS: torchrun sets it to 1
        threads = n < 1 ? 1 : (n > 8 ? 8 : n);
        ok = true;
        return true;
    }
} g_stager;

bool is_pageable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

int upload(double *dst, const double *src, int64_t cnt, cudaStream_t s)
{
    if (cnt <= 0) return 0;
    if (!src) return MSGWAM_E_BADARG;
    const size_t bytes = (size_t)cnt * sizeof(double);
    if (bytes < ((size_t)4 << 20) || !is_pageable(src) || !g_stager.init())
        return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    const char *from = reinterpret_cast<const char *>(src);
    char *to = reinterpret_cast<char *>(dst);
    for (size_t off = 0; off < bytes; off += STAGE_CHUNK) {
        const size_t len = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
        const int k = g_stager.next;
        g_stager.next ^= 1;
        cudaError_t e = cudaEventSynchronize(g_stager.ev[k]);          // the chunk's previous upload has left it
        if (e != cudaSuccess) return (int)e;
        const int T = g_stager.threads;
        const size_t per = ((len + T - 1) / T + 63) & ~(size_t)63;
#pragma omp parallel for num_threads(T) schedule(static)
        for (int t = 0; t < T; ++t) {
            const size_t b = (size_t)t * per;
            if (b < len) memcpy(g_stager.buf[k] + b, from + off + b, b + per <= len ? per : len - b);
        }
        e = cudaMemcpyAsync(to + off, g_stager.buf[k], len, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return (int)e;
        e = cudaEventRecord(g_stager.ev[k], s);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}
}  // namespace

// Page-locked staging for the grid-sized arguments of msgwam_rk3_column_host (one block per process, portable across
// the process's devices, grown on demand;
// every call ends with a stream synchronisation, so no copy is in flight when it is reallocated).
static double *g_pin = nullptr;
static size_t g_pin_cap = 0;
static double *pinned_block(size_t doubles)
{
    if (doubles > g_pin_cap) {
        if (g_pin) cudaFreeHost(g_pin);
        g_pin = nullptr; g_pin_cap = 0;
        const size_t cap = doubles < 65536 ? 65536 : doubles;
        if (cudaHostAlloc(reinterpret_cast<void **>(&g_pin), cap * sizeof(double), cudaHostAllocPortable) != cudaSuccess) {
            g_pin = nullptr;
            return nullptr;
        }
        g_pin_cap = cap;
    }
    return g_pin;
}

// ray-sized doubles of the staging slab: 10 uploaded ray fields + ff + pkl, then the outputs and the stage-1 hand-over
// (constant N: rr, mm + 3; N(z) profile: rr, drr, mm, dmm + 7)
static inline int64_t ray_slots(bool nz) { return nz ? 12 + 4 + 7 : 12 + 2 + 3; }

int64_t msgwam_host_stage_doubles(int64_t n, int32_t G)
{
    if (n < 0 || G < 3) return 0;
    // then grid (G+1), grids, rhobar, pg (2G), uu, vv, uu_out, vv_out, (bvf: unused here, keeps one layout), deposit bounds
    return ray_slots(false) * pad32(n) + pad32(G + 1) + 9 * pad32(G) + 32;
}

int64_t msgwam_host_stage_doubles_nz(int64_t n, int32_t G)
{
    if (n < 0 || G < 3) return 0;
    return ray_slots(true) * pad32(n) + pad32(G + 1) + 9 * pad32(G) + 32;
}

// h_bvf != NULL: the N(z) extension (msgwam_column_step_nz; rr, drr, mm, dmm come back), else the reference's scalar N
static int rk3_column_host_impl(const msgwam_params_t *p, int64_t n, const double *const h_state[9], const double *h_dkk,
                                const double *h_dll, const double *h_uu, const double *h_vv, const double *h_grid,
                                const double *h_grids, const double *h_rhobar, const double *h_pg, const double *h_bvf,
                                double *h_rr_out, double *h_drr_out, double *h_mm_out, double *h_dmm_out, double *h_uu_out,
                                double *h_vv_out, double *d_stage, double *d_work, void *stream)
{
    const bool nz = h_bvf != nullptr;
    if (!p || n < 0 || !h_state || !h_uu || !h_vv || !h_grid || !h_grids || !h_rhobar || !h_pg || !h_uu_out || !h_vv_out ||
        !d_stage || !d_work)
        return MSGWAM_E_BADARG;
    if (p->G < 3) return MSGWAM_E_GRID_SIZE;
    if (n > 0 && (!h_rr_out || !h_mm_out || (nz && (!h_drr_out || !h_dmm_out)))) return MSGWAM_E_BADARG;
    if ((h_dkk == nullptr) != (h_dll == nullptr)) return MSGWAM_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t np = pad32(n), G = p->G, gp = pad32(G);
    double *d = d_stage;
    // state order: dens, lam, phi, rr, drr, kk, ll, mm, dmm  (lam is not needed on the device)
    double *d_dens = d, *d_phi = d + np, *d_rr = d + 2 * np, *d_drr = d + 3 * np, *d_kk = d + 4 * np, *d_ll = d + 5 * np,
           *d_mm = d + 6 * np, *d_dmm = d + 7 * np, *d_dkk = d + 8 * np, *d_dll = d + 9 * np, *d_ff = d + 10 * np,
           *d_pkl = d + 11 * np, *d_out = d + 12 * np;
    double *d_rro = d_out, *d_mmo = d_out + np, *d_drro = d_out + 2 * np, *d_dmmo = d_out + 3 * np;   // the last two: N(z) only
    double *d_st1 = d_out + (nz ? 4 : 2) * np;
    double *g = d + ray_slots(nz) * np;
    double *d_grid = g, *d_grids = g + pad32(G + 1), *d_rho = d_grids + gp, *d_pg = d_rho + gp, *d_uu = d_pg + 2 * gp,
           *d_vv = d_uu + gp, *d_bvf = d_vv + gp, *d_uuo = d_bvf + gp, *d_vvo = d_uuo + gp, *d_bounds = d_vvo + gp;
    cudaError_t e;
#define MW_H2D(dst, src, cnt)                                                                          \
    do {                                                                                               \
        const int rc_up = upload((dst), (src), (cnt), s);     /* staged through page-locked chunks if pageable */ \
        if (rc_up) return rc_up;                                                                       \
    } while (0)
    // The grid-sized inputs are ordinary (pageable) numpy arrays: six cudaMemcpyAsync calls from pageable memory are
    // six staged, blocking copies (~100 us in all).  They are gathered into one page-locked block laid out like the
    // device region and go up with ONE asynchronous copy; uu, vv and the error word come back the same way.
    const int64_t gblock = pad32(G + 1) + 7 * gp;                 // grid | grids | rhobar | pg (2) | uu | vv | bvf
    double *pin = pinned_block((size_t)(gblock + 2 * gp + 8));
    if (!pin) return (int)cudaErrorMemoryAllocation;
    {
        double *q = pin;
        memcpy(q, h_grid, (size_t)(G + 1) * sizeof(double)); q += pad32(G + 1);
        memcpy(q, h_grids, (size_t)G * sizeof(double)); q += gp;
        memcpy(q, h_rhobar, (size_t)G * sizeof(double)); q += gp;
        memcpy(q, h_pg, (size_t)(2 * G) * sizeof(double)); q += 2 * gp;                  // (2, G) contiguous, as the kernels index it
        memcpy(q, h_uu, (size_t)G * sizeof(double)); q += gp;
        memcpy(q, h_vv, (size_t)G * sizeof(double)); q += gp;
        if (nz) memcpy(q, h_bvf, (size_t)G * sizeof(double));
    }
    e = cudaMemcpyAsync(d_grid, pin, (size_t)gblock * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return (int)e;
    MW_H2D(d_phi, h_state[2], n);
    if (h_dkk) { MW_H2D(d_dkk, h_dkk, n); MW_H2D(d_dll, h_dll, n); }   // NULL: reuse the statics left in d_stage
    int rc = msgwam_derive_statics(d_phi, d_dkk, d_dll, d_ff, d_pkl, n, p->two_rot, stream);
    if (rc) return rc;
    MW_H2D(d_dens, h_state[0], n); MW_H2D(d_rr, h_state[3], n); MW_H2D(d_drr, h_state[4], n); MW_H2D(d_kk, h_state[5], n);
    MW_H2D(d_ll, h_state[6], n); MW_H2D(d_mm, h_state[7], n); MW_H2D(d_dmm, h_state[8], n);
#undef MW_H2D
    msgwam_rays_t r{};
    r.dens = d_dens; r.phi = d_phi; r.rr = d_rr; r.drr = d_drr; r.kk = d_kk; r.ll = d_ll; r.mm = d_mm; r.dmm = d_dmm;
    r.dkk = d_dkk; r.dll = d_dll; r.ff = d_ff; r.pkl = d_pkl; r.stage1 = d_st1; r.bounds = d_bounds;
    msgwam_grid_t gr{d_grid, d_grids, d_rho, d_pg, nz ? d_bvf : nullptr};
    // a fresh state every call: one cheap sweep bounds its deposits so that the step's CTA histograms run in fixed point
    rc = msgwam_column_bounds(p, &r, n, &gr, stream);
    if (rc) return rc;
    if (nz) rc = msgwam_column_step_nz(p, &r, n, &gr, d_uu, d_vv, d_work, d_rro, d_drro, d_mmo, d_dmmo, d_uuo, d_vvo, nullptr, stream);
    else rc = msgwam_column_step(p, &r, n, &gr, d_uu, d_vv, d_work, d_rro, d_mmo, d_uuo, d_vvo, stream);
    if (rc) return rc;
    if (n > 0) {
        double *const hs[4] = {h_rr_out, h_mm_out, h_drr_out, h_dmm_out};
        double *const ds[4] = {d_rro, d_mmo, d_drro, d_dmmo};
        for (int k = 0; k < (nz ? 4 : 2); ++k) {
            e = cudaMemcpyAsync(hs[k], ds[k], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) return (int)e;
        }
    }
    double *pin_out = pin + gblock;                               // uu_out | vv_out | error word
    e = cudaMemcpyAsync(pin_out, d_uuo, (size_t)(2 * gp) * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return (int)e;
    // the error word of the bounded device-side waits travels back with the results
    double *d_err = d_work + msgwam_column_error_offset(p->G);
    e = cudaMemcpyAsync(pin_out + 2 * gp, d_err, sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return (int)e;
    memcpy(h_uu_out, pin_out, (size_t)G * sizeof(double));
    memcpy(h_vv_out, pin_out + gp, (size_t)G * sizeof(double));
    const double err_word = pin_out[2 * gp];
    if (err_word != 0.0) {
        cudaMemsetAsync(d_err, 0, sizeof(double), s);
        return MSGWAM_E_TIMEOUT;
    }
    return 0;
}

int msgwam_rk3_column_host(const msgwam_params_t *p, int64_t n, const double *const h_state[9], const double *h_dkk,
                           const double *h_dll, const double *h_uu, const double *h_vv, const double *h_grid,
                           const double *h_grids, const double *h_rhobar, const double *h_pg, double *h_rr_out,
                           double *h_mm_out, double *h_uu_out, double *h_vv_out, double *d_stage, double *d_work,
                           void *stream)
{
    return rk3_column_host_impl(p, n, h_state, h_dkk, h_dll, h_uu, h_vv, h_grid, h_grids, h_rhobar, h_pg, nullptr, h_rr_out,
                                nullptr, h_mm_out, nullptr, h_uu_out, h_vv_out, d_stage, d_work, stream);
}

int msgwam_rk3_column_nz_host(const msgwam_params_t *p, int64_t n, const double *const h_state[9], const double *h_dkk,
                              const double *h_dll, const double *h_uu, const double *h_vv, const double *h_grid,
                              const double *h_grids, const double *h_rhobar, const double *h_pg, const double *h_bvf,
                              double *h_rr_out, double *h_drr_out, double *h_mm_out, double *h_dmm_out, double *h_uu_out,
                              double *h_vv_out, double *d_stage, double *d_work, void *stream)
{
    if (!h_bvf) return MSGWAM_E_BADARG;
    return rk3_column_host_impl(p, n, h_state, h_dkk, h_dll, h_uu, h_vv, h_grid, h_grids, h_rhobar, h_pg, h_bvf, h_rr_out,
                                h_drr_out, h_mm_out, h_dmm_out, h_uu_out, h_vv_out, d_stage, d_work, stream);
}

}  // extern "C"
