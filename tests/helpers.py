"""Shared helpers for the parity tests."""
import numpy as np

from msgwam_b200 import scenarios

FIELDS = scenarios.STATE_NAMES + ("uu", "vv")


def scenario_from_npz(d, name="golden"):
    state = [d[k] for k in scenarios.STATE_NAMES]
    model = dict(bvf=float(d["bvf"]), phi0=float(d["phi0"]), kappa=float(d["kappa"]),
                 saturate_online=bool(d["saturate_online"]))
    return scenarios.Scenario(name, float(d["dt"]), state, d["uu"], d["vv"], d["dkk"], d["dll"], d["rr_mm_area"],
                              d["grid"], d["grids"], d["rhobar"], d["pressure_gradient"], model, hprop=bool(d["hprop"]))


def max_rel(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) elementwise; 0 where both are exactly equal."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    diff = np.abs(a - b)
    den = np.maximum(np.abs(b), floor)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(diff == 0, 0.0, diff / den)
    return float(np.max(r)) if r.size else 0.0


def field_rel(a, b):
    """max |a-b| / max|b|  -- the norm used for deposited grid fields (summation-order noise)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b)) if b.size else 0.0
    if s == 0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / s)
