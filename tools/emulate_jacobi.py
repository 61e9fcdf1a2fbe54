"""numpy (float32) emulation of csrc/eigen_jacobi.cu's algorithm -- used in the build container (no GPU)
to validate the tournament, the rotation formulas and the convergence behaviour before spending GPU time."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

JB, JP = 8, 16


def tournament(step, pair, nb):
    mth = nb - 1
    if pair == 0:
        a, b = mth, step % mth
    else:
        a, b = (step + pair) % mth, (step - pair + mth) % mth
    return (a, b) if a < b else (b, a)


def inner_jacobi(B, tol, max_inner=12):
    B = B.astype(np.float32).copy()
    Q = np.eye(JP, dtype=np.float32)
    f = np.float32
    for isw in range(max_inner):
        off = f(0)
        for st in range(JP - 1):
            J = np.eye(JP, dtype=np.float32)
            for k in range(JB):
                a, b = tournament(st, k, JP)
                app, aqq, apq = B[a, a], B[b, b], B[a, b]
                c, s = f(1), f(0)
                if app > 0 and aqq > 0:
                    rel = abs(apq) / np.sqrt(app * aqq)
                    if rel > 1e-9:
                        tau = (aqq - app) / (f(2) * apq)
                        t = np.copysign(f(1), tau) / (abs(tau) + np.sqrt(f(1) + tau * tau))
                        c = f(1) / np.sqrt(f(1) + t * t)
                        s = t * c
                    off = max(off, rel)
                J[a, a] = c; J[a, b] = s; J[b, a] = -s; J[b, b] = c
            B = (J.T @ B @ J).astype(np.float32)
            Q = (Q @ J).astype(np.float32)
        if off <= 0.1 * tol:
            break
    return Q, isw + 1


def block_jacobi(A, tol=5e-6, max_sweeps=40, verbose=True):
    p = A.shape[0]
    cols = (p + JP - 1) // JP * JP
    nb = cols // JB
    G = np.zeros((p, cols), dtype=np.float32)
    G[:, :p] = A.astype(np.float32)
    inner_total = 0
    for sweep in range(max_sweeps):
        off = 0.0
        for step in range(nb - 1):
            for pair in range(nb // 2):
                I, J = tournament(step, pair, nb)
                idx = np.r_[I * JB:(I + 1) * JB, J * JB:(J + 1) * JB]
                P = G[:, idx]
                B = (P.T @ P).astype(np.float32)
                d = np.diag(B)
                with np.errstate(divide="ignore", invalid="ignore"):
                    rel = np.abs(B) / np.sqrt(np.outer(d, d))
                rel[~np.isfinite(rel)] = 0
                np.fill_diagonal(rel, 0)
                po = rel.max()
                off = max(off, po)
                if po > 0.25 * tol:
                    Q, ni = inner_jacobi(B, tol)
                    inner_total += ni
                    G[:, idx] = (P @ Q).astype(np.float32)
        if verbose:
            print("sweep", sweep, "off", off, "inner sweeps so far", inner_total)
        if off <= tol:
            break
    lam = np.linalg.norm(G[:, :p].astype(np.float64), axis=0)
    order = np.argsort(lam)
    U = G[:, :p][:, order] / lam[order]
    return lam[order], U, sweep + 1


if __name__ == "__main__":
    from oracle import oracle_np as o
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", sys.argv[1] if len(sys.argv) > 1 else "cat_small_random50.npz"))
    img = g["image"]
    if int(g["rgb"]):
        img = np.repeat(img[:, :, None], 3, axis=2)
    r = o.run_pipeline(img, g["sample_indices"], kind=str(g["kind"]), h_loc=float(g["h_loc"]), h_val=float(g["h_val"]))
    L = r["L_A"]
    lam, U, sweeps = block_jacobi(L)
    mu = np.linalg.eigvalsh(L)
    print("sweeps", sweeps, "max rel eig err", np.max(np.abs(lam - mu) / mu))
    print("orth err", np.abs(U.T @ U - np.eye(U.shape[1])).max(), "resid", np.abs(L @ U - U * lam).max())
