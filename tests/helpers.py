import numpy as np


def match_pairs(pg, po):
    """Align GPU and oracle pair records by (tag_i, tag_j); returns index arrays."""
    kg = {(int(a), int(b)): k for k, (a, b) in enumerate(zip(pg["tag_i"], pg["tag_j"]))}
    ko = {(int(a), int(b)): k for k, (a, b) in enumerate(zip(po["tag_i"], po["tag_j"]))}
    assert set(kg) == set(ko), "pair lists differ: only-gpu %s only-oracle %s" % (
        sorted(set(kg) - set(ko))[:5], sorted(set(ko) - set(kg))[:5])
    keys = sorted(kg)
    return np.array([kg[k] for k in keys]), np.array([ko[k] for k in keys])


def pair_rel_errors(pg, po, rscale=1.0):
    """Max relative errors of V, F, torque over matched pairs.

    Tolerance convention (north_star: <= 1e-10 relative on per-pair overlap volume, force, torque):
    V relative to V; F relative to |F|; torque relative to max(|tau|, |F|*rscale) (a torque
    component can vanish by symmetry while the lever-arm scale |F| R does not)."""
    ig, io = match_pairs(pg, po)
    Vg, Vo = pg["V"][ig], po["V"][io]
    assert np.array_equal(Vg > 0, Vo > 0), "contact / no-contact decision differs"
    m = Vo > 0
    if not m.any():
        return dict(V=0.0, F=0.0, tau=0.0, centroid=0.0, ncontact=0)
    eV = np.max(np.abs(Vg[m] - Vo[m]) / Vo[m])
    Fg, Fo = pg["F"][ig][m], po["F"][io][m]
    fn = np.linalg.norm(Fo, axis=1)
    eF = np.max(np.linalg.norm(Fg - Fo, axis=1) / fn)
    et = 0.0
    for key in ("tau_i", "tau_j"):
        tg, to = pg[key][ig][m], po[key][io][m]
        sc = np.maximum(np.linalg.norm(to, axis=1), fn * rscale)
        et = max(et, np.max(np.linalg.norm(tg - to, axis=1) / sc))
    xg, xo = pg["centroid"][ig][m], po["centroid"][io][m]
    dxc = np.linalg.norm(xg - xo, axis=1) / rscale
    ex = np.max(dxc)
    # the centroid is a first moment divided by V: for a grazing contact (V -> 0) its rounding error grows like 1/V while
    # everything it feeds (torques of the dissipative terms) is multiplied by forces ~ V; weight by V / median V
    Vm = np.median(Vo[m])
    exw = np.max(dxc * np.minimum(1.0, Vo[m] / Vm))
    return dict(V=eV, F=eF, tau=et, centroid=ex, centroid_weighted=exw, ncontact=int(m.sum()))
