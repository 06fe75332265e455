"""numpy restatement of the N(z) EXTENSION -- there is no such code in the reference.

Specification (E1-E9 in DESIGN.md): `bvf` is an array of N on `grids`; wherever the reference uses the scalar
`bvf ** 2` it becomes `np.interp(z, grids, bvf) ** 2` at the position argument the reference already passes
(cg_rr / cg_lambda / cg_phi / the cg_rr call inside wave_projection); dm_dt gains
`- N N' (k^2 + l^2) / om / |k|^2` with `N' = np.interp(rr, grid[1:-1], diff(bvf)/dz)`.
Used to pin the C oracle's extension branch; the reduction of a constant profile to the reference's scalar
is pinned separately against the golden fixtures.  Small cases only (the deposit is a Python loop).
"""
import numpy as np

RAD, ROT = 6378e3, 7.2921e-5


class ExtReference:
    def __init__(self, cfg):
        self.c = cfg
        self.grid, self.grids = np.asarray(cfg["grid"], float), np.asarray(cfg["grids"], float)
        self.bvf = np.asarray(cfg["bvf"], float)
        self.f0 = 2 * ROT * np.sin(cfg["phi0"])

    def n_at(self, z):
        return np.interp(z, self.grids, self.bvf)

    def omega(self, kk, ll, mm, ff, z):
        return np.sqrt((self.n_at(z) ** 2 * (kk ** 2 + ll ** 2) + ff ** 2 * mm ** 2) / (kk ** 2 + ll ** 2 + mm ** 2))

    def cg_rr(self, kk, ll, mm, phi, z):
        ff = 2 * ROT * np.sin(phi)
        om = self.omega(kk, ll, mm, ff, z)
        return - mm * (om ** 2 - ff ** 2) / om / (kk ** 2 + ll ** 2 + mm ** 2)

    def projection(self, dens, phi, rl, ru, kk, ll, ml, mu, dkk, dll, dmm):
        g = self.grids
        dz = np.diff(g[:2])[0]
        nlow = (rl / dz).astype(int); nup = (ru / dz + 1.).astype(int)
        nzmax = len(g) - 2
        ood = ((nlow >= nzmax) & (nup >= nzmax)) | ((nlow <= 0) & (nup <= 0))
        nlow = np.clip(nlow, 0, nzmax); nup = np.clip(nup, 0, nzmax)
        psv = abs(dkk * dll * dmm)
        cgr = self.cg_rr(kk, ll, .5 * (ml + mu), phi, .5 * (rl + ru))
        v0, v1 = cgr * kk * dens, cgr * ll * dens
        out = np.zeros((2, len(g) - 1))
        for i in range(len(dens)):
            if ood[i]:
                continue
            for c in range(nlow[i], nup[i]):
                w = np.abs(min(g[c + 1], ru[i]) - max(g[c], rl[i])) / dz
                out[0, c] += w * psv[i] * v0[i]
                out[1, c] += w * psv[i] * v1[i]
        return out

    def rhs(self, dt, var):
        c = self.c
        dens, lam, phi, rr, drr, kk, ll, mm, dmm, uu, vv = [np.asarray(a, float) for a in var]
        grid, grids = self.grid, self.grids
        dz = np.diff(grid[:2])[0]
        ff = 2 * ROT * np.sin(phi)
        vk = kk ** 2 + ll ** 2 + mm ** 2
        om = self.omega(kk, ll, mm, ff, rr)
        n2 = self.n_at(rr) ** 2
        cup, cdn = self.cg_rr(kk, ll, mm, phi, rr + .5 * drr), self.cg_rr(kk, ll, mm, phi, rr - .5 * drr)
        du = np.interp(rr, grid[1:-1], (uu[1:] - uu[:-1]) / dz)
        dv = np.interp(rr, grid[1:-1], (vv[1:] - vv[:-1]) / dz)
        if c["hprop"]:
            cgl = kk / om / vk * (n2 - om ** 2) + np.interp(rr, grids, uu)
            cgp = ll / om / vk * (n2 - om ** 2) + np.interp(rr, grids, vv)
        else:
            cgl = np.zeros(kk.shape); cgp = np.zeros(kk.shape)
        t = [None] * 11
        t[1] = cgl / (RAD + rr) / np.cos(phi)
        t[2] = cgp / (RAD + rr)
        t[3] = .5 * (cdn + cup)
        t[4] = cup - cdn
        if c["hprop"]:
            cgr = self.cg_rr(kk, ll, mm, phi, rr)
            zero = (kk * 0. + ll * 0.)
            t[5] = kk / (RAD + rr) * (np.tan(phi) * cgp - cgr) - zero / (RAD + rr) / np.cos(phi)
            df2 = 8 * ROT ** 2 * np.sin(phi) * np.cos(phi) * 1
            t[6] = - (ll * cgr + kk * np.tan(phi) * cgl + mm ** 2 / 2 / om / vk * df2) / (RAD + rr) - zero / (RAD + rr)
        else:
            t[5] = np.zeros(kk.shape); t[6] = np.zeros(kk.shape)
        t[7] = (kk * cgl + ll * cgp) / (RAD + rr) - (kk * du + ll * dv)
        dn = np.interp(rr, grid[1:-1], (self.bvf[1:] - self.bvf[:-1]) / dz)
        t[7] = t[7] - self.n_at(rr) * dn * (kk ** 2 + ll ** 2) / om / vk
        t[8] = dmm / drr * t[4]
        # saturation (tendency form) with N at rr_final / rr_center
        rr_f, drr_f, mm_f = rr + t[3] * dt, drr + t[4] * dt, mm + t[7] * dt
        dmm_f = c["rr_mm_area"] / drr_f
        rho = np.interp(rr_f, grids, c["rhobar"])
        omh = np.sqrt((n2 * (kk ** 2 + ll ** 2) + self.f0 ** 2 * mm ** 2) / vk)
        maxd = c["kappa"] ** 2 * .5 * rho * omh * self.n_at(rr_f) ** 2 / mm_f ** 2 / (omh ** 2 - self.f0 ** 2)
        st = np.where(maxd < dens * (c["dkk"] * c["dll"] * dmm_f), (maxd - dens) / dt, 0.0)
        t[0] = c["saturate_online"] * st
        proj = self.projection(dens, phi, rr - .5 * drr, rr + .5 * drr, kk, ll, mm - .5 * dmm, mm + .5 * dmm,
                               c["dkk"], c["dll"], dmm)
        flux = np.zeros((2, len(grid)))
        flux[:, 1:-1] = proj
        flux[:, 0] = flux[:, 1]; flux[:, -1] = flux[:, -2]
        grad = (flux[:, 1:] - flux[:, :-1]) / dz
        rho_g, pg = np.asarray(c["rhobar"], float), np.asarray(c["pressure_gradient"], float)
        t[9] = self.f0 * vv - rho_g ** -1 * (pg[0] + grad[0])
        t[10] = -self.f0 * uu - rho_g ** -1 * (pg[1] + grad[1])
        out = np.empty(11, dtype=object)
        for i in range(11):
            out[i] = t[i]
        return out

    def RK3(self, dt, var):
        qq = dt * self.rhs(dt, var)
        var = var + qq / 3
        qq = dt * self.rhs(dt, var) - 5 / 9 * qq
        var = var + 15 / 16 * qq
        qq = dt * self.rhs(dt, var) - 153 / 128 * qq
        return var + 8 / 15 * qq
