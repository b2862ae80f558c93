#!/usr/bin/env python
"""CPU half of the fix-up flag study: per-row conditioning features (float64 restatement of the flow's step Jacobians) against
the measured RAW errors of the tensor-core launch (gpurun_out/flag_dump.npz from profiles/flag_dump.py).  For every candidate
feature: the share of rows that must be flagged (in the feature's order) for the unflagged rest to meet the pdf bar, per
material, the feature value at that point, and what ONE global threshold costs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bsdf_oracle as O     # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "bsdf_diffusion_sampling_b200"))


def load_packs():
    import importlib.util, types, json, struct
    mats = {}
    for kind in ("disk", "spherical", "bsdf"):
        path = os.path.join(ROOT, "oracle", "_ref", f"all_{kind}.bsdfpack")
        buf = np.fromfile(path, dtype=np.uint8)
        jl = struct.unpack("<I", bytes(buf[8:12]))[0]
        index = json.loads(bytes(buf[12:12 + jl]).decode())
        base_off = (12 + jl + 63) // 64 * 64
        for rec in index["materials"]:
            arr = lambda off, cnt: np.frombuffer(buf, dtype="<f4", count=cnt, offset=base_off + off)
            layers = [arr(f["offset"], int(np.prod(f["shape"]))).reshape(f["shape"]).astype(np.float64) for f in rec["flow"]]
            b = arr(rec["base"]["offset"], 308).astype(np.float64)
            mats[f"{kind}/{rec['name']}"] = (kind, rec["T"], O.FlowWeights(layers),
                                             O.BaseWeights(b[:224].reshape(16, 14), b[224:240], b[240:304].reshape(4, 16), b[304:308]))
    return mats


def features(flow, base, x, wi, T, reverse):
    pe = O.positional_encoding(wi, 5)
    n = x.shape[0]
    mind = np.full(n, np.inf); amp = np.ones(n); EB = np.zeros(n); EBmax = np.zeros(n); S = np.zeros(n); EBw = np.zeros(n)
    ampsig = np.ones(n)
    inv_t = 1.0 / T
    for t in range(T):
        alpha = (1 - t / T) if reverse else (t / T)
        d, du, dv = O.flow_velocity(flow, x, alpha, pe)
        sgn = -1.0 if reverse else 1.0
        g00, g01, g10, g11 = sgn * inv_t * du[:, 0], sgn * inv_t * dv[:, 0], sgn * inv_t * du[:, 1], sgn * inv_t * dv[:, 1]
        j00, j01, j10, j11 = 1 + g00, g01, g10, 1 + g11
        det = j00 * j11 - j01 * j10
        ad = np.abs(det)
        mind = np.minimum(mind, ad)
        fro = j00 ** 2 + j01 ** 2 + j10 ** 2 + j11 ** 2
        amp *= np.maximum(1.0, fro - 1.0)
        smax2 = 0.5 * (fro + np.sqrt(np.maximum(fro * fro - 4 * det * det, 0)))
        eb = (np.abs(j11 * g00) + np.abs(j00 * g11) + np.abs(j10 * g01) + np.abs(j01 * g10)) / ad
        EBw += eb * np.sqrt(ampsig)           # weighted by the amplification accumulated so far (state error grows with it)
        ampsig *= np.maximum(1.0, smax2)
        EB += eb
        EBmax = np.maximum(EBmax, eb)
        S += np.sqrt(g00 ** 2 + g01 ** 2 + g10 ** 2 + g11 ** 2) / ad
        x = x + sgn * inv_t * d
    f = {"mind": mind, "amp": amp, "EB": EB, "EBmax": EBmax, "S": S, "EBw": EBw, "ampsig": ampsig}
    if reverse:
        p = O.base_forward(base, wi)
        if flow.domain == O.DISK:
            g0 = (x[:, 0] - p[:, 0]) * np.exp(-2 * p[:, 2]); g1 = (x[:, 1] - p[:, 1]) * np.exp(-2 * p[:, 3])
        else:
            sc = np.exp(p[:, 1]) + 1e-3
            kappa = O.softplus(p[:, 3]) + 1e-3
            g0 = (x[:, 0] - p[:, 0]) / sc ** 2; g1 = kappa * np.sin(x[:, 1] - p[:, 2])
        f["gn"] = np.sqrt(g0 ** 2 + g1 ** 2)
    return f


def need(err, score, bar=5e-2, out_bar=1e-3, med_bar=5e-3):
    """smallest k such that flagging the k highest-score rows leaves p99(err) <= bar, median(err) <= med_bar and
    P(err > 0.5) <= out_bar; -> (k / n, score at the cut)"""
    n = err.size
    order = np.argsort(-score, kind="stable")
    e = err[order]
    # after flagging the first k rows their error is 0: p99 over n rows = the value with 1 % of n rows above it
    allow = int(np.floor(0.01 * n))
    # count of rows with err > bar among the unflagged must be <= allow  (and err > .5 count <= out_bar n)
    bad = np.cumsum((e > bar)[::-1])[::-1]            # bad[k] = # rows > bar among rows k..n-1
    bad5 = np.cumsum((e > 0.5)[::-1])[::-1]
    badm = np.cumsum((e > med_bar)[::-1])[::-1]
    ok = (bad <= allow) & (bad5 <= int(out_bar * n)) & (badm <= n // 2 - 1)
    k = int(np.argmax(ok)) if ok.any() else n
    return k / n, (score[order][k] if k < n else -np.inf)


def main():
    z = np.load(os.path.join(ROOT, "gpurun_out", "flag_dump.npz"))
    mats = load_packs()
    rows = []
    for key, (kind, T, flow, base) in mats.items():
        wi, x0, x32 = (z[f"{key}/{k}"].astype(np.float64) for k in ("wi", "x0", "x32"))
        rs, rp, dx = (z[f"{key}/{k}"].astype(np.float64) for k in ("rs", "rp", "dx"))
        fs = features(flow, base, x0, wi, T, False)
        fp = features(flow, base, x32, wi, T, True)
        rows.append((key, kind, T, rs, rp, dx, fs, fp))
    cands = {
        "cur": lambda f, rev: 1.0 / (np.minimum(1, 5 * f["mind"]) * np.minimum(1, 16 / np.sqrt(f["amp"])) * (np.minimum(1, 25 / np.maximum(f["gn"], 1e-9)) if rev else 1.0)),
        "1/mind": lambda f, rev: 1.0 / f["mind"],
        "EB": lambda f, rev: f["EB"],
        "EBmax": lambda f, rev: f["EBmax"],
        "S": lambda f, rev: f["S"],
        "EBw": lambda f, rev: f["EBw"],
        "EB*sqrt(ampsig)": lambda f, rev: f["EB"] * np.sqrt(f["ampsig"]),
        "EB+gn/8": lambda f, rev: f["EB"] + (f["gn"] / 8 if rev else 0),
        "EBw+gn/8": lambda f, rev: f["EBw"] + (f["gn"] / 8 if rev else 0),
        "oracle": None,
    }
    verbose = "-v" in sys.argv
    for mode in ("sample", "pdf"):
        print(f"==== {mode}: share of rows to flag (in each feature's order) for the rest to meet p99 <= 5e-2, and the cut value")
        print("%-46s %9s " % ("material", "raw p99") + " ".join("%16s" % c for c in cands))
        cuts = {c: [] for c in cands}
        for key, kind, T, rs, rp, dx, fs, fp in rows:
            err = rs if mode == "sample" else rp
            f = fs if mode == "sample" else fp
            line = "%-46s %9.2e " % (key, np.quantile(err, 0.99))
            anyk = False
            for c, fn in cands.items():
                score = err if fn is None else fn(f, mode == "pdf")
                k, cut = need(err, score)
                cuts[c].append((key, kind, k, cut))
                anyk |= k > 0
                line += " %6.2f%% %8.3g" % (100 * k, cut)
            if verbose or anyk:
                print(line)
        print("---- ONE threshold per feature and plugin kind (the lowest cut any material of the kind needs) -> share flagged")
        for c, fn in cands.items():
            if fn is None:
                continue
            line = "%-18s" % c
            for kd in ("disk", "spherical", "bsdf"):
                finite = [cut for _, kind, k, cut in cuts[c] if kind == kd and k > 0 and np.isfinite(cut)]
                thr = min(finite) if finite else np.inf
                v = [(fn(fs if mode == "sample" else fp, mode == "pdf") >= thr).mean() for key, kind, T, rs, rp, dx, fs, fp in rows if kind == kd]
                line += f" | {kd}: thr {thr:8.3g} mean {100*np.mean(v):5.2f}% max {100*np.max(v):5.2f}%"
            print(line)


if __name__ == "__main__":
    main()
