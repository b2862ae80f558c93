"""CPU prediction (numpy emulation of the tensor-core arithmetic, profiles/emulate_tc16.py) of what the conditioning-triggered
fp32 fix-up does to the RAW pdf error per golden material: share of rows recomputed and p99 of the rest, per threshold."""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "profiles"))
import emulate_tc16 as E
from oracle import bsdf_oracle as O
from oracle import c_oracle as C
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
    flow, base, z = O.load_material_npz(path)
    T = int(z["T"]); wi, x0 = z["wi"], z["x0"]
    f64 = flow.astype(np.float64)
    for rev in (False, True):
        start = z["wo_eval"] if rev else x0
        wi_ = z["wi_eval"] if rev else wi
        if rev:
            pref, w = C.pdf(flow, base, start, wi_, T, with_mindet=True)
        else:
            _, pref, w = C.sample(flow, base, wi_, T, start, with_mindet=True)
        xt, Rt = O._euler(f64, start.astype(np.float64), wi_.astype(np.float64), T, rev)
        x, R = E.run(flow, wi_, start, T, rev, "h2tan")
        rel = np.abs(R / Rt - 1.0)
        if rev:  # include base density at end point error
            lp = (O.base_logprob_disk if flow.domain == O.DISK else O.base_logprob_spherical)
            b64 = base.astype(np.float64)
            rel = np.abs(np.exp(lp(b64, x.astype(np.float64), wi_.astype(np.float64)) - lp(b64, xt, wi_.astype(np.float64))) * R / Rt - 1.0)
        s = f"{os.path.basename(path)[:-4]:40s} {'pdf' if rev else 'smp'} raw p99 {np.quantile(rel,0.99):.2e}"
        for thr in (0.1, 0.25, 0.5, 0.75):
            fl = w < thr
            r2 = np.where(fl, 0.0, rel)
            s += f" | w<{thr}: flag {fl.mean()*100:.2f}% p99 {np.quantile(r2,0.99):.2e} >.5:{(r2>0.5).mean()*100:.3f}%"
        print(s)
