#!/usr/bin/env python
"""numpy emulation of the tensor-core kernel's ARITHMETIC (not its code): where fp16 roundings happen in the
shipped formulation ("f32tan": fp32 tangent accumulators, fp32 activation math) and in the half2 formulation
("h2tan": fp16 tangent accumulators, s2 and the tangent products in half2).  Used on the CPU to predict the
error statistics against the goldens before spending GPU time.  tanh.approx is modelled as exact tanh rounded
to 17 bits of relative precision: PTX documents 2^-11, but on sm_100a the instruction measures max 1e-5 / rms 4e-6
relative error (profiles/microbench/tanh_err_r2.txt) -- the 11-bit model of round 1 overstated its share of the error.

    python profiles/emulate_tc16.py
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bsdf_oracle as O  # noqa: E402

f16 = np.float16
f32 = np.float32


def r16(a):
    return a.astype(f16).astype(f32)


def split(w):
    hi = r16(w)
    lo = r16(w - hi)
    return hi, lo


def tanh_approx(z):
    t = np.tanh(z.astype(np.float64))
    m, e = np.frexp(t)
    return np.ldexp(np.round(m * 131072.0) / 131072.0, e).astype(f32)


def mm(a, w):          # fp32 accumulate of exactly representable fp16 products
    return (a.astype(np.float64) @ w.astype(np.float64).T).astype(f32)


def mm16(a, w):        # fp16 accumulator: K chunks of 16, rounded to fp16 after each chunk
    acc = np.zeros((a.shape[0], w.shape[0]), f32)
    for k in range(0, a.shape[1], 16):
        acc = r16(acc.astype(np.float64) + a[:, k:k + 16].astype(np.float64) @ w[:, k:k + 16].astype(np.float64).T)
    return acc


def act(zh, du, dv, mode):
    t = tanh_approx(zh)
    h = r16(zh + zh * t)                                     # fp32 fma, packed to fp16 (the next A operand)
    if mode == "f32tan":
        hf = (zh + zh * t).astype(f32)
        s2 = (1.0 + t) + hf * (1.0 - t)
        return h, r16(s2 * du), r16(s2 * dv)
    t16 = r16(t)
    a, b = r16(1.0 - t16), r16(1.0 + t16)
    s2 = r16(h.astype(np.float64) * a + b)                   # one half2 fma
    return h, r16(s2 * du), r16(s2 * dv)


def velocity(W, x, alpha, pe16, domain, mode):
    n = x.shape[0]
    H = W[0].shape[0]
    hi0, lo0 = split(W[0])
    if domain == O.DISK:
        st = np.concatenate([x, np.full((n, 1), alpha, f32)], 1)
        su = np.zeros((n, 3), f32); su[:, 0] = 1
        sv = np.zeros((n, 3), f32); sv[:, 1] = 1
        ns = 3
    else:
        s, c = np.sin(x[:, 1:2]), np.cos(x[:, 1:2])
        st = np.concatenate([x[:, 0:1], s, c, np.full((n, 1), alpha, f32)], 1).astype(f32)
        su = np.zeros((n, 4), f32); su[:, 0] = 1
        sv = np.concatenate([np.zeros((n, 1), f32), c, -s, np.zeros((n, 1), f32)], 1).astype(f32)
        ns = 4
    st_hi = r16(st); st_lo = r16(st - st_hi)
    inp_hi = np.concatenate([st_hi, pe16], 1)
    w = hi0 + lo0                                            # exactly what hi.A + lo.A accumulates (fp32 acc)
    zh = 0.5 * (mm(inp_hi, w) + mm(st_lo, w[:, :ns]))
    tmm = mm if mode == "f32tan" else mm16
    if domain == O.DISK:
        du = 0.5 * tmm(np.pad(su, ((0, 0), (0, 13))), np.pad(hi0[:, :3], ((0, 0), (0, 13))))
        dv = 0.5 * tmm(np.pad(sv, ((0, 0), (0, 13))), np.pad(hi0[:, :3], ((0, 0), (0, 13))))
    else:
        sv_hi = r16(sv); sv_lo = r16(sv - sv_hi)
        du = 0.5 * tmm(np.pad(su, ((0, 0), (0, 12))), np.pad(hi0[:, :4], ((0, 0), (0, 12))))
        dv = 0.5 * tmm(np.pad(np.concatenate([sv_hi, sv_lo], 1), ((0, 0), (0, 8))),
                       np.pad(np.concatenate([hi0[:, :4], hi0[:, :4]], 1), ((0, 0), (0, 8))))
    h, u, v = act(zh, du, dv, mode)
    for Wk in W[1:-1]:
        hi, lo = split(0.5 * Wk)
        zh = mm(h, hi + lo)
        du, dv = tmm(u, hi), tmm(v, hi)
        h, u, v = act(zh, du, dv, mode)
    hi, lo = split(W[-1])
    return mm(h, hi + lo), mm(u, hi), mm(v, hi)               # output round: fp32 accumulators in both modes


def run(flow, wi, x0, T, reverse, mode):
    W = [np.asarray(w, f32) for w in flow.layers]
    pe16 = r16(O.positional_encoding(wi.astype(f32), 5))
    x = x0.astype(f32).copy()
    R = np.ones(x.shape[0], f32)
    inv_t = f32(1.0 / T)
    sg = f32(-1.0 if reverse else 1.0)
    for t in range(T):
        alpha = f32((1 - t / T) if reverse else t / T)
        d, du, dv = velocity(W, x, alpha, pe16, flow.domain, mode)
        det = (1 + sg * inv_t * du[:, 0]) * (1 + sg * inv_t * dv[:, 1]) - (sg * inv_t * dv[:, 0]) * (sg * inv_t * du[:, 1])
        R = R * det if reverse else R / det
        x = x + sg * inv_t * d
    return x, R


def q(a):
    a = a[np.isfinite(a)]
    return f"med {np.median(a):.2e} p99 {np.quantile(a, 0.99):.2e}"


def main():
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        flow, base, z = O.load_material_npz(path)
        T = int(z["T"])
        wi, x0 = z["wi"], z["x0"]
        f64 = flow.astype(np.float64)
        xt, Rt = O._euler(f64, x0.astype(np.float64), wi.astype(np.float64), T, False)
        for mode in ("f32tan", "h2tan"):
            x, R = run(flow, wi, x0, T, False, mode)
            rel = np.abs(R / Rt - 1.0)
            print(f"{os.path.basename(path)[:-4]:44s} {mode:7s} |dx| {q(np.abs(x - xt).ravel())}   R rel {q(rel)}")


if __name__ == "__main__":
    main()
