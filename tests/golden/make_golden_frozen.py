"""Golden fixture of the frozen-background mode "M2" (an EXTENSION: the reference has no such stepper), composed ONLY of
functions of the unmodified Python reference, through the reference's own plug-in point model_config['rhs'] (L:691):

    rays : lprop.RK3(dt, var) with model_config['rhs'] = rhs_default whose du_st, dv_st are replaced by zeros
    flow : once per step  pm_flux[:, 1:-1] = wave_projection(new rays, grids, var=0); edge copies; diff / dz (L:653-663);
           uu += dt * du_dt(vv, grad[0]); vv += dt * dv_dt(uu, grad[1])                                 (L:523-558)

Run in the build container (where /root/reference exists):   python tests/golden/make_golden_frozen.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200"))

from _reference import load_reference  # noqa: E402
from msgwam_b200 import scenarios  # noqa: E402
from make_golden import pack_scenario  # noqa: E402


def frozen_step(ref, dt, var):
    """one step of the frozen-background mode out of reference functions (model_config['rhs'] must be rhs_frozen)"""
    var = ref.RK3(dt, var)
    dens, lam, phi, rr, drr, kk, ll, mm, dmm, uu, vv = var
    pm_flux = np.zeros((2, len(ref.grid)))
    pm_flux[:, 1:-1] = ref.wave_projection(dens, lam, phi, rr - .5 * drr, rr + .5 * drr, kk, ll, mm - .5 * dmm, mm + .5 * dmm,
                                           ref.statics['dkk'], ref.statics['dll'], dmm, ref.grids)
    pm_flux[:, 0] = pm_flux[:, 1]
    pm_flux[:, -1] = pm_flux[:, -2]
    dz = np.diff(ref.grid[:2])[0]
    grad = (pm_flux[:, 1:] - pm_flux[:, :-1]) / dz
    out = np.empty(11, dtype=object)
    for i in range(9):
        out[i] = var[i]
    out[9] = uu + dt * ref.du_dt(vv, grad[0])
    out[10] = vv + dt * ref.dv_dt(uu, grad[1])
    return out


def install_frozen(ref, sc):
    sc.install(ref)

    def rhs_frozen(dt, var_in):
        t = ref.rhs_default(dt, var_in)
        t[9] = np.zeros(np.shape(var_in[9]))
        t[10] = np.zeros(np.shape(var_in[10]))
        return t
    ref.set_model_setup(rhs=rhs_frozen)


def main():
    ref = load_reference()
    assert ref is not None
    sc = scenarios.column_ensemble(1201, seed=9, ngrid=201, sheared=True, amplitude=0.3)
    install_frozen(ref, sc)
    out = pack_scenario(sc)
    var = sc.var()
    for step in (1, 2, 3):
        var = frozen_step(ref, sc.dt, var)
        for i, nm in enumerate(scenarios.STATE_NAMES + ("uu", "vv")):
            out["step%d_%s" % (step, nm)] = np.asarray(var[i], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "frozen_col.npz"), **out)
    print("frozen_col.npz: uu change over 3 steps", float(np.max(np.abs(out["step3_uu"] - sc.uu))))


if __name__ == "__main__":
    main()
