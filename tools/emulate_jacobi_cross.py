"""numpy emulation of k_jacobi's sweep at config 4 (L_A from the committed golden's D + the oracle's K_A): Gram screen, active panel
pairs by distance class, one inner 16 x 16 Jacobi sweep per visit -- the full cyclic sweep (15 steps) against a variant that rotates
only the 64 CROSS pairs (8 steps) on visits other than a panel's home pair (2k, 2k+1).  Prints the relative off-diagonal after every
sweep and the eigenvalue error: whether the cheaper inner sweep still converges in the same number of outer sweeps.

    python tools/emulate_jacobi_cross.py [full|cross|hybrid] [fixture.npz | large]   (hybrid = what the kernel does: cross-only while the sweep
    starts above 4 tol)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_c as oc  # noqa: E402

JB, JP = 8, 16
TOL = 5e-5


def tournament(step, pair, nb):
    mth = nb - 1
    if pair == 0:
        a, b = mth, step % mth
    else:
        a, b = (step + pair) % mth, (step - pair + mth) % mth
    return (a, b) if a < b else (b, a)


def rotation(app, aqq, apq):
    f = np.float32
    c, s, rel = 1.0, 0.0, f(0)
    if app > 0 and aqq > 0:
        rel = abs(f(apq)) / np.sqrt(f(app) * f(aqq))
        if rel > 1e-12:
            tau = f(aqq - app) / (f(2) * f(apq))
            t = float(np.copysign(f(1), tau) / (abs(tau) + np.sqrt(f(1) + tau * tau)))
            c = 1.0 / np.sqrt(1.0 + t * t)
            s = t * c
    return c, s, rel


def inner(B, cross_only):
    B = B.astype(np.float64).copy()
    Q = np.eye(JP)
    steps = []
    if cross_only:
        for r in range(JB):
            steps.append([(k, JB + ((k + r) & 7)) for k in range(JB)])
    else:
        for st in range(JP - 1):
            steps.append([tournament(st, k, JP) for k in range(JB)])
    for prs in steps:
        J = np.eye(JP)
        for a, b in prs:
            c, s, _ = rotation(B[a, a], B[b, b], B[a, b])
            J[a, a] = c; J[a, b] = s; J[b, a] = -s; J[b, b] = c
        B = J.T @ B @ J
        Q = Q @ J
    return Q


def screen(G, nb):
    C = (G.T.astype(np.float64) @ G.astype(np.float64))
    d = np.sqrt(np.diag(C))
    with np.errstate(divide="ignore", invalid="ignore"):
        R = np.abs(C) / np.outer(d, d)
    R[~np.isfinite(R)] = 0
    np.fill_diagonal(R, 0)
    rel = R.reshape(nb, JB, nb, JB).max(axis=(1, 3))
    return rel


def run(L, mode):
    p = L.shape[0]
    cols = (p + JP - 1) // JP * JP
    nb = cols // JB
    G = np.zeros((p, cols), dtype=np.float32)
    G[:, :p] = L.astype(np.float32)
    for sweep in range(12):
        rel = screen(G, nb)
        # a panel's within-block goes with its home pair (2k, 2k+1)
        pair_rel = np.triu(rel, 1)
        for k in range(nb // 2):
            pair_rel[2 * k, 2 * k + 1] = max(rel[2 * k, 2 * k + 1], rel[2 * k, 2 * k], rel[2 * k + 1, 2 * k + 1])
        off = pair_rel.max()
        active = np.argwhere(pair_rel > 0.2 * TOL)
        print(f"sweep {sweep}: off {off:.3e}, active pairs {len(active)}", flush=True)
        if off <= TOL:
            break
        classes = {}
        if len(active) > nb * (nb - 1) // 4:      # more than half of the pairs: round-robin tournament
            act = {(int(I), int(J)) for I, J in active}
            for st in range(nb - 1):
                for k in range(nb // 2):
                    pr = tournament(st, k, nb)
                    if pr in act:
                        classes.setdefault(st, []).append(pr)
        else:
            for I, J in active:
                d = J - I
                classes.setdefault(2 * d + ((I // d) & 1), []).append((I, J))
        visits = 0
        for key in sorted(classes):
            for I, J in classes[key]:
                idx = np.r_[I * JB:(I + 1) * JB, J * JB:(J + 1) * JB]
                P = G[:, idx]
                B = P.T.astype(np.float64) @ P.astype(np.float64)
                dd = np.sqrt(np.diag(B))
                with np.errstate(divide="ignore", invalid="ignore"):
                    R = np.abs(B) / np.outer(dd, dd)
                R[~np.isfinite(R)] = 0
                np.fill_diagonal(R, 0)
                if R.max() <= 0.25 * TOL:
                    continue
                home = (J == I + 1) and (I % 2 == 0)
                Q = inner(B, cross_only=((mode == "cross" or (mode == "hybrid" and off > 4 * TOL)) and not home))
                G[:, idx] = (P.astype(np.float64) @ Q).astype(np.float32)
                visits += 1
        print(f"   visits {visits}, classes {len(classes)}")
    lam = np.sort(np.linalg.norm(G[:, :p].astype(np.float64), axis=0))
    mu = np.linalg.eigvalsh(L)
    err = float(np.max(np.abs(lam[1:] - mu[1:]) / mu[1:]))
    print("max rel eigenvalue error (column norms, before the Rayleigh refinement):", err)
    return sweep, float(off), err


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "full"
    if len(sys.argv) > 2 and sys.argv[2] == "large":      # tests/test_gpu_parity.py::test_large_sample_count (350 x 300, p = 4200 uniform)
        from oracle import oracle_np as o
        img = o.synthetic_image(350, 300, 1, seed=17)
        smp = oc.uniform_sampling(350, 300, 4000)
        r = o.run_pipeline(img, smp)
        print("mode", mode, "large p", r["L_A"].shape[0])
        run(r["L_A"], mode)
        sys.exit(0)
    if len(sys.argv) > 2:      # a small committed fixture instead of config 4
        from oracle import oracle_np as o
        g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", sys.argv[2]))
        img = g["image"]
        if int(g["rgb"]):
            img = np.repeat(img[:, :, None], 3, axis=2) if img.ndim == 2 else img
        r = o.run_pipeline(img, g["sample_indices"], kind=str(g["kind"]), h_loc=float(g["h_loc"]), h_val=float(g["h_val"]))
        print("mode", mode, sys.argv[2], "p", r["L_A"].shape[0])
        run(r["L_A"], mode)
        sys.exit(0)
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "c4_full.npz"))
    W, H = int(g["width"]), int(g["height"])
    img = oc.synthetic_image(W, H, 1, int(g["seed_img"]))
    s = g["sample_indices"].astype(np.int64)
    K_A = oc.affinity_rows(img, s, s)
    D = g["D"]
    L = (np.diag(D) - K_A) / D.mean()
    print("mode", mode)
    run(L, mode)
