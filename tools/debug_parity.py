import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-msgwam_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import msgwam_b200.libprop as lprop
from msgwam_b200 import scenarios
import oracle
from helpers import max_rel, field_rel, FIELDS

for (sheared, shuffled, n, ngrid) in [(False, False, 100003, 1001), (True, False, 100003, 1001), (True, True, 50021, 401)]:
    sc = scenarios.column_ensemble(n, seed=11, ngrid=ngrid, sheared=sheared, shuffled=shuffled, amplitude=0.3)
    sc.install(lprop)
    orc = oracle.Oracle(sc.oracle_cfg())
    dens, lam, phi, rr, drr, kk, ll, mm, dmm = sc.state
    args = (dens, lam, phi, rr - .5*drr, rr + .5*drr, kk, ll, mm - .5*dmm, mm + .5*dmm, sc.dkk, sc.dll, dmm, sc.grids)
    pg = lprop.wave_projection(*args, var=0); po = orc.wave_projection(*args, var=0)
    print("case", sheared, shuffled, n, "projection field_rel", field_rel(pg[0], po[0]), field_rel(pg[1], po[1]), "max|D|", np.abs(po).max(),
          "mean|terms|~", np.abs(po).mean())
    rg = lprop.rhs_default(sc.dt, sc.var()); ro = orc.rhs_default(sc.dt, sc.var())
    print("   rhs: ", {nm: (max_rel(rg[i], ro[i]) if i < 9 else field_rel(rg[i], ro[i])) for i, nm in enumerate(FIELDS)})
    vg = lprop.RK3(sc.dt, sc.var()); vo = orc.RK3(sc.dt, sc.var())
    print("   RK3 numpy: ", {nm: (max_rel(vg[i], vo[i]) if i < 9 else field_rel(vg[i], vo[i])) for i, nm in enumerate(FIELDS) if nm in ("rr","mm","uu","vv")})
    k = int(np.argmax(np.abs(vg[7]-vo[7])/np.abs(vo[7])))
    print("   worst ray", k, "mm0", mm[k], "mm got", vg[7][k], "want", vo[7][k], "dm", vo[7][k]-mm[k], "rr", rr[k], "->", vo[3][k])
    # general-path RK3 (stage by stage) through a plugin wrapper
    lprop.set_model_setup(rhs=lambda dt, v: lprop.rhs_default(dt, v))
    vp = lprop.RK3(sc.dt, sc.var())
    lprop.set_model_setup(rhs=lprop.rhs_default)
    print("   RK3 staged: ", {nm: (max_rel(vp[i], vo[i]) if i < 9 else field_rel(vp[i], vo[i])) for i, nm in enumerate(FIELDS) if nm in ("rr","mm","uu","vv")})
