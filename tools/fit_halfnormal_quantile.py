#!/usr/bin/env python
"""Derive the fp32 polynomial coefficients used by the embed kernel for the half-normal quantile

    g(v) = sqrt(2) * erfinv(v) = Phi^-1(1/2 + v/2),      v = (2m+1) * 2^-24,  m in [0, 2^23)

in exactly the arithmetic the kernel uses (see csrc/gswm_math.cuh):

    f = 1 + m*2^-23            (bit pattern 0x3F800000 | m)
    v = f - (1 - 2^-24)        (exact)
    q = (K - K*v) * f          (K = 2^c: one FFMA + one FMUL;  q ~ K * (1 - v^2))
    X = lg2(q)                 (MUFU.LG2)              ->  X = c + lg2((1-v)*f)
    central (X >= X_SPLIT):    g = v * P(X)
    tail    (X <  X_SPLIT):    s = sqrt(c - X) ;  g = Q(s - S0)

The fit is our own (weighted least squares on Chebyshev nodes then a discrete Remez exchange, in
float64, against scipy.special.ndtri), not a transcription of any published coefficient table.
Writes csrc/gswm_coeffs.inc and prints the emulated-fp32 error statistics.
"""
import argparse
import os

import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import ndtri

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "a-watermark-for-diffusion-models_b200", "csrc", "gswm_coeffs.inc")


def target(v):
    return ndtri(0.5 + 0.5 * np.asarray(v, dtype=np.float64))


def lawson_fit(xs, ys, deg, lo, hi, iters=40):
    """Near-minimax fit of ys(xs) in RELATIVE error by a degree-`deg` polynomial.

    Chebyshev-weighted least squares gives the starting error curve; a discrete Remez exchange on
    the sorted samples then levels it.  The best iterate (sup norm over ALL samples) is returned as
    float64 monomial coefficients in xs, highest power first, together with that sup norm."""
    tt = np.clip((2 * xs - (lo + hi)) / (hi - lo), -1.0, 1.0)
    A = C.chebvander(tt, deg) / ys[:, None]      # relative error: A c - 1
    n = deg + 2
    # Chebyshev-weighted LS: weight each sample by local sample spacing / sqrt(1 - t^2)
    dx = np.gradient(tt)
    wts = np.sqrt(np.abs(dx) / np.sqrt(np.maximum(1 - tt * tt, 1e-6)))
    c, *_ = np.linalg.lstsq(A * wts[:, None], wts, rcond=None)
    best = (np.abs(A @ c - 1.0).max(), c.copy())
    for _ in range(iters):
        err = A @ c - 1.0
        sgn = np.sign(err)
        sgn[sgn == 0] = 1
        cuts = np.flatnonzero(np.diff(sgn)) + 1
        starts = np.concatenate([[0], cuts])
        stops = np.concatenate([cuts, [len(err)]])
        ext = np.array([a + np.argmax(np.abs(err[a:b])) for a, b in zip(starts, stops)])
        if len(ext) < n:
            break
        while len(ext) > n:      # drop the weaker end; alternation is preserved
            ext = ext[1:] if abs(err[ext[0]]) < abs(err[ext[-1]]) else ext[:-1]
        M = np.hstack([A[ext], ((-1.0) ** np.arange(n))[:, None]])
        try:
            sol = np.linalg.solve(M, np.ones(n))
        except np.linalg.LinAlgError:
            break
        c = sol[:-1]
        m = np.abs(A @ c - 1.0).max()
        if m < best[0]:
            improved = best[0] - m > 1e-3 * m
            best = (m, c.copy())
            if not improved:
                break
    cheb = best[1]
    # chebyshev in tt -> monomial in xs
    p_tt = C.cheb2poly(cheb)                      # low -> high in tt
    a = 2.0 / (hi - lo)
    b0 = -(lo + hi) / (hi - lo)
    poly = np.poly1d([0.0])
    base = np.poly1d([a, b0])
    for k, ck in enumerate(p_tt):
        poly = poly + ck * base ** k
    return np.asarray(poly.coeffs, dtype=np.float64), best[0]


def f32(x):
    return np.asarray(x, dtype=np.float32)


def horner32(coeffs, x):
    """fp32 Horner with fused multiply-add emulated through float64 (exact product, one rounding)."""
    p = np.full(x.shape, np.float32(coeffs[0]), dtype=np.float32)
    xd = x.astype(np.float64)
    for c in coeffs[1:]:
        p = (p.astype(np.float64) * xd + np.float64(np.float32(c))).astype(np.float32)
    return p


def kernel_arith(m, c_shift):
    """Emulate the kernel's fp32 front end for integer m (uint32 array)."""
    f = (np.uint32(0x3F800000) | m.astype(np.uint32)).view(np.float32)
    v = (f.astype(np.float64) - (1.0 - 2.0 ** -24)).astype(np.float32)
    K = 2.0 ** c_shift
    t = (K - K * v.astype(np.float64)).astype(np.float32)       # FFMA, one rounding
    q = (t.astype(np.float64) * f.astype(np.float64)).astype(np.float32)
    X = np.log2(q.astype(np.float64)).astype(np.float32)        # MUFU.LG2 modelled as correctly rounded
    return f, v, X


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--deg-central", type=int, default=8)
    ap.add_argument("--deg-tail", type=int, default=8)
    ap.add_argument("--w-split", type=float, default=5.0, help="central/tail split in w = -ln(1-v^2)")
    ap.add_argument("--deg64", type=int, default=16, help="degree of the float64-path polynomials")
    ap.add_argument("--no-write", action="store_true")
    ap.add_argument("--out", default=OUT, help="where to write the coefficient header (default: csrc/gswm_coeffs.inc)")
    ap.add_argument("--exhaustive", action="store_true", help="evaluate all 2^23 inputs (slow-ish)")
    args = ap.parse_args()

    x_split_raw = -args.w_split / np.log(2.0)          # lg2(1-v^2) at the split
    c_shift = float(np.round(-x_split_raw / 2.0))       # K = 2^c centres the central interval roughly
    x_lo, x_hi = x_split_raw + c_shift, c_shift         # X range for the central branch
    X_SPLIT = np.float32(x_lo)

    # ---- sample points: all structure comes from the kernel's own front end -----------------
    rng = np.random.RandomState(0)
    m_all = np.unique(np.concatenate([
        rng.randint(0, 1 << 23, size=400000).astype(np.uint32),
        np.arange(0, 4096, dtype=np.uint32),
        (1 << 23) - 1 - np.arange(0, 200000, dtype=np.uint32),
        ((1 << 23) - 1 - (np.logspace(0, 6.9, 20000)).astype(np.uint32)),
    ]))
    f, v, X = kernel_arith(m_all, c_shift)
    g = target(v.astype(np.float64))
    central = X >= X_SPLIT

    # central: P(X) = g / v on X in [x_lo, x_hi]; fit against the *exact* X (float64) of each sample
    K = 2.0 ** c_shift
    vd = v.astype(np.float64)
    Xd = np.log2((K - K * vd) * f.astype(np.float64))
    cm = central
    order = np.argsort(Xd[cm])
    xs, ys = Xd[cm][order], (g[cm] / vd[cm])[order]
    # thin to Chebyshev-like density so the ends are represented
    cc, e_c = lawson_fit(xs, ys, args.deg_central, x_lo - 1e-3, x_hi + 1e-6)

    # tail: Q(s - S0) = g, s = sqrt(c - X)
    tm = ~central
    s_all = np.sqrt(c_shift - Xd[tm])
    s_lo, s_hi = np.sqrt(c_shift - x_lo) - 1e-3, np.sqrt(c_shift - Xd.min()) + 1e-3
    S0 = float(np.float32(0.5 * (s_lo + s_hi)))
    order = np.argsort(s_all)
    ct, e_t = lawson_fit(s_all[order] - S0, g[tm][order], args.deg_tail, s_lo - S0, s_hi - S0)

    print(f"c_shift={c_shift}  X_SPLIT={float(X_SPLIT)!r}  S0={S0!r}")
    print(f"float64 fit error (relative, sup over samples): central {e_c:.3e}  tail {e_t:.3e}")
    print(f"tail probability per element: {1.0 - np.sqrt(1 - np.exp(-args.w_split)):.5f}")

    # ---- emulate the kernel in fp32 ------------------------------------------------------
    def emulate(m):
        f, v, X = kernel_arith(m, c_shift)
        P = horner32(cc, X)
        gc = (v.astype(np.float64) * P.astype(np.float64)).astype(np.float32)
        s = np.sqrt(np.maximum(np.float64(c_shift) - X.astype(np.float64), 0)).astype(np.float32)
        sm = (s.astype(np.float64) - S0).astype(np.float32)
        gt = horner32(ct, sm)
        return np.where(X >= X_SPLIT, gc, gt), v, X

    def report(m, label):
        ge, v, X = emulate(m)
        ref = target(v.astype(np.float64))
        rel = np.abs(ge.astype(np.float64) - ref) / ref
        cen = X >= X_SPLIT
        print(f"{label}: n={m.size}  max rel err central {rel[cen].max():.3e}  "
              f"tail {rel[~cen].max() if (~cen).any() else 0:.3e}  "
              f"(fp32 rounding of the exact answer alone is <= 5.96e-8)")
        return rel.max()

    report(m_all, "samples   ")

    # The kernel enters its rare tail-patch block on an INTEGER test that is evaluated before the MUFU: some element of the
    # float4 has bits(f) > FGUARD_BITS.  It must never miss an element with X < X_SPLIT, so the guard sits 64 grid points
    # below the first m whose emulated X falls under X_SPLIT (X moves 1.4e-4 per grid point there, MUFU.LG2's error is
    # ~2e-7), and the implication is checked over all 2^23 values of m.
    m_every = np.arange(1 << 23, dtype=np.uint32)
    _, _, X_every = kernel_arith(m_every, c_shift)
    m_first_tail = int(m_every[X_every < X_SPLIT].min())
    m_guard = m_first_tail - 64
    assert (X_every[:m_guard + 1] >= X_SPLIT).all() and (X_every[m_first_tail + 64:] < X_SPLIT).all()
    print(f"integer guard: first tail m={m_first_tail}, guard m={m_guard} (fbits 0x{0x3F800000 | m_guard:08X}); "
          f"X(2^23-1)={float(X_every[-1])!r}, X(2^23-2)={float(X_every[-2])!r}")

    # ---- the outermost grid cell, refined (uniforms v3): m = 2^23 - 1 is subdivided by a 28-bit m2, p = P(|Z| > z) / 2 =
    # (m2 + 1/2) 2^-52, z in [5.29, 8.21].  Kernel arithmetic: mf = float(m2) + 0.5f; s = sqrt(52 - lg2(mf)); z = R(s - S1).
    m2 = np.unique(np.concatenate([np.arange(0, 70000), (1 << 28) - 1 - np.arange(0, 70000),
                                   np.random.RandomState(28).randint(0, 1 << 28, size=300000),
                                   np.logspace(0, 8.42, 100000).astype(np.int64)]))
    m2 = m2[m2 < (1 << 28)]
    z_far = -ndtri((m2.astype(np.float64) + 0.5) * 2.0 ** -52)
    s_far = np.sqrt(52.0 - np.log2(m2 + 0.5))
    f_lo, f_hi = s_far.min() - 1e-3, s_far.max() + 1e-3
    S1 = float(np.float32(0.5 * (f_lo + f_hi)))
    o = np.argsort(s_far)
    cf, e_f = lawson_fit(s_far[o] - S1, z_far[o], 5, f_lo - S1, f_hi - S1)
    mf = (m2.astype(np.float32) + np.float32(0.5)).astype(np.float32)
    s32 = np.sqrt((np.float32(52.0) - np.log2(mf.astype(np.float64)).astype(np.float32)).astype(np.float64)).astype(np.float32)
    z32 = horner32(cf, (s32.astype(np.float64) - S1).astype(np.float32))
    e_far = (np.abs(z32.astype(np.float64) - z_far) / z_far).max()
    print(f"refined outermost cell (degree 5 in sqrt(52 - lg2 m2)): fit error {e_f:.3e}, emulated fp32 max rel err {e_far:.3e}, "
          f"z in [{z_far.min():.4f}, {z_far.max():.4f}]")
    assert e_far < 5e-7

    # ---- float64 path (injected-uniform mode): same structure in w = -ln(t*(2-t)), natural log ----
    # samples are (v, t) pairs with t = 1 - v exact in float64
    def tgt_vt(v, t):
        return np.where(v < 0.5, ndtri(0.5 + 0.5 * v), -ndtri(0.5 * t))

    k = np.unique(np.concatenate([np.arange(1, 1 << 16), rng.randint(1, 1 << 30, size=300000),
                                  (1 << 30) - np.unique(np.logspace(0, 9.03, 200000).astype(np.int64))]))
    v64 = k.astype(np.float64) * 2.0 ** -30
    t64 = 1.0 - v64
    w64 = -np.log(t64 * (2.0 - t64))
    g64 = tgt_vt(v64, t64)
    W_SPLIT = args.w_split
    cm = w64 < W_SPLIT
    o = np.argsort(w64[cm])
    c64, e64c = lawson_fit(w64[cm][o] - 0.5 * W_SPLIT, (g64[cm] / v64[cm])[o], args.deg64, -0.5 * W_SPLIT - 1e-9,
                           0.5 * W_SPLIT + 1e-9)
    # tail A: w in [W_SPLIT, WA]; tail B: w in [WA, 38] (t down to 2^-54, the smallest p the reference can see)
    WA = 17.0
    tb = np.unique(np.concatenate([2.0 ** -np.linspace(0.0, 54.5, 400000), t64[~cm]]))
    tb = tb[tb < 0.5]
    wb = -np.log(tb * (2.0 - tb))
    gb = -ndtri(0.5 * tb)
    sb = np.sqrt(wb)
    ma = (wb >= W_SPLIT - 1e-6) & (wb <= WA + 1e-6)
    mb = (wb >= WA - 1e-6) & (wb <= 38.2)
    SA0 = 0.5 * (np.sqrt(W_SPLIT) + np.sqrt(WA))
    SB0 = 0.5 * (np.sqrt(WA) + np.sqrt(38.2))
    o = np.argsort(sb[ma])
    ca64, e64a = lawson_fit(sb[ma][o] - SA0, gb[ma][o], args.deg64, sb[ma].min() - SA0 - 1e-9, sb[ma].max() - SA0 + 1e-9)
    o = np.argsort(sb[mb])
    cb64, e64b = lawson_fit(sb[mb][o] - SB0, gb[mb][o], args.deg64, sb[mb].min() - SB0 - 1e-9, sb[mb].max() - SB0 + 1e-9)
    print(f"float64 path (deg {args.deg64}): fit error central {e64c:.3e}  tailA {e64a:.3e}  tailB {e64b:.3e}")

    if not args.no_write:
        with open(args.out, "w") as fo:
            fo.write("// Generated by tools/fit_halfnormal_quantile.py -- do not edit by hand.\n")
            fo.write(f"// deg_central={args.deg_central} deg_tail={args.deg_tail} w_split={args.w_split}\n")
            fo.write(f"#define GSWM_HNQ_KSCALE {2.0 ** c_shift!r}f\n")
            fo.write(f"#define GSWM_HNQ_CSHIFT {c_shift!r}f\n")
            fo.write(f"#define GSWM_HNQ_XSPLIT {float(X_SPLIT)!r}f\n")
            fo.write(f"#define GSWM_HNQ_S0 {S0!r}f\n")
            fo.write(f"#define GSWM_HNQ_FGUARD_BITS 0x{0x3F800000 | m_guard:08X}u   // X < XSPLIT  =>  bits(f) > this  (m > {m_guard})\n")
            fo.write("// highest power first\n")
            fo.write("#define GSWM_HNQ_CENTRAL_COEFFS " + ", ".join(f"{float(np.float32(c))!r}f" for c in cc) + "\n")
            fo.write("#define GSWM_HNQ_TAIL_COEFFS " + ", ".join(f"{float(np.float32(c))!r}f" for c in ct) + "\n")
            fo.write(f"#define GSWM_HNQ_S1 {S1!r}f\n")
            fo.write("#define GSWM_HNQ_FARTAIL_COEFFS " + ", ".join(f"{float(np.float32(c))!r}f" for c in cf) + "\n")
            fo.write("// float64 path: g = v*Pc(w - WSPLIT/2) | Qa(sqrt(w) - SA0) | Qb(sqrt(w) - SB0),  w = -ln(t*(2-t))\n")
            fo.write(f"#define GSWM_HNQ64_WSPLIT {W_SPLIT!r}\n#define GSWM_HNQ64_WA {WA!r}\n")
            fo.write(f"#define GSWM_HNQ64_SA0 {float(SA0)!r}\n#define GSWM_HNQ64_SB0 {float(SB0)!r}\n")
            fo.write("#define GSWM_HNQ64_CENTRAL_COEFFS " + ", ".join(f"{float(c)!r}" for c in c64) + "\n")
            fo.write("#define GSWM_HNQ64_TAILA_COEFFS " + ", ".join(f"{float(c)!r}" for c in ca64) + "\n")
            fo.write("#define GSWM_HNQ64_TAILB_COEFFS " + ", ".join(f"{float(c)!r}" for c in cb64) + "\n")
        print("wrote", os.path.normpath(args.out))


if __name__ == "__main__":
    main()
