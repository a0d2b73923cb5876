"""Research probe (CPU, numpy): the column signs LAPACK's gesdd gives the top right singular vectors of a tall
matrix X (M >= 11N/6: QR -> gebrd -> bdsdc path) as a closed-form function of G = X^T X and the top-left R x R block
of X.  Finding (validated against torch.linalg.svd = MKL sgesdd and scipy = netlib dgesdd):

  * V = PB * V_B with PB from gebrd(R): PB equals the orthogonal factor of the e_1-preserving Householder
    tridiagonalisation of G (same reflectors).
  * bdsdc leaves the signs of the dominant vectors to the implicit-QR leaf solver, whose limit obeys the classical
    "leading principal minor" rule: det(VB[0:i+1, 0:i+1]) / det(VB[0:i, 0:i]) has the sign of d_i, the i-th diagonal
    entry of the bidiagonal B (before bdsqr makes the singular values positive by flipping rows of V^T).
  * sign(d_i) = D_i * sign(d_i(C)): D_i = sign of the i-th diagonal entry of the Householder-QR factor of X
    (depends on the first i+1 rows/columns of X and on G only), d_i(C) from the left Householder history of gebrd
    run on the Cholesky factor C of G.

Run:  python tools/research/lapack_sign_rule.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import qmf_port as port  # noqa: E402


def sgn(x):  # LAPACK SIGN(1, x): +1 for +0
    return -1.0 if (x < 0 or (x == 0 and np.signbit(x))) else 1.0


def larfg(alpha, x):
    xnorm = np.linalg.norm(x)
    if x.size == 0 or xnorm == 0.0:
        return alpha, x, 0.0
    beta = -sgn(alpha) * np.hypot(alpha, xnorm)
    tau = (beta - alpha) / beta
    return beta, x / (alpha - beta), tau


def tridiag_forward(G, steps):
    """e_1-preserving Householder tridiagonalisation (dsytd2 'L' conventions), first `steps` reflectors.
    Returns the first steps+1 columns of Q (= PB of gebrd)."""
    A = G.copy()
    n = A.shape[0]
    Q = np.eye(n)
    for j in range(steps):
        beta, v, tau = larfg(A[j + 1, j], A[j + 2:, j])
        w = np.concatenate(([1.0], v))
        if tau != 0:
            H = np.eye(n - j - 1) - tau * np.outer(w, w)
            A[j + 1:, :] = H @ A[j + 1:, :]
            A[:, j + 1:] = A[:, j + 1:] @ H
            Q[:, j + 1:] = Q[:, j + 1:] @ H
    return Q[:, :steps + 1]


def qr_diag_signs(Xtop, Gtop):
    """signs of the first R diagonal entries of the Householder-QR factor of X from its top R x R block and the
    R x R Gram block: QR of [Xtop; Y] with Y^T Y = Gtop - Xtop^T Xtop."""
    R = Xtop.shape[0]
    rem = Gtop - Xtop.T @ Xtop
    Y = np.linalg.cholesky(rem + 1e-9 * np.trace(rem) / R * np.eye(R)).T
    Z = np.vstack([Xtop, Y])
    D = np.zeros(R)
    for k in range(R):
        beta, v, tau = larfg(Z[k, k], Z[k + 1:, k])
        w = np.concatenate(([1.0], v))
        D[k] = sgn(beta)
        Z[k, k] = beta
        Z[k + 1:, k] = 0
        if tau != 0 and k + 1 < R:
            Z[k:, k + 1:] -= tau * np.outer(w, w @ Z[k:, k + 1:])
    return D


def predict_flips(G, Xtop, V):
    """V: (N, R) orthonormal top-R eigenvectors of G in any sign.  Returns +-1 per column so that V*flip has
    LAPACK's signs."""
    N, R = V.shape
    P = tridiag_forward(G, R - 1) if R > 1 else np.eye(N)[:, :1]  # p_0..p_{R-1}
    C = np.linalg.cholesky(G).T
    D = qr_diag_signs(Xtop[:R, :R], G[:R, :R])
    ws = []
    dsign = np.zeros(R)
    for l in range(R):
        z = C @ P[:, l]
        for j, (w, tau) in enumerate(ws):
            z[j:] -= tau * w * (w @ z[j:])
        beta, v, tau = larfg(z[l], z[l + 1:])
        ws.append((np.concatenate(([1.0], v)), tau))
        dsign[l] = sgn(beta) * D[l]
    VB = P.T @ V  # (R, R)
    flips = np.ones(R)
    prev = 1.0
    for i in range(R):
        m = np.linalg.det(VB[:i + 1, :i + 1] * flips[None, :i + 1])
        if sgn(m / prev) != dsign[i]:
            flips[i] = -1.0
            m = -m
        prev = m
    return flips


def check(name, imgs, ranks=(4, 2, 2)):
    ok = np.zeros(max(ranks)); tot = np.zeros(max(ranks)); img_ok = 0
    for img in imgs:
        good = True
        for pl, (x, _, _) in enumerate(port.qmf_planes(img)):
            R = ranks[pl]
            _, _, vh = torch.linalg.svd(x, full_matrices=False)
            X = x.numpy().astype(np.float64)
            G = X.T @ X
            w, E = np.linalg.eigh(G)
            V = E[:, ::-1][:, :R]
            fl = predict_flips(G, X[:R, :R], V)
            ref = np.sign(np.sum(V * vh[:R].numpy().T, axis=0))
            ok[:R] += fl == ref
            tot[:R] += 1
            good &= bool(np.all(fl == ref))
        img_ok += good
    print(f"{name}: per-component agreement {ok / np.maximum(tot, 1)}; images with every sign right {img_ok}/{len(imgs)}")


if __name__ == "__main__":
    torch.set_num_threads(8)
    from PIL import Image
    check("S-nat 512x768", [port.s_nat(5000 + i, 512, 768) for i in range(24)])
    check("S-nat 256x384", [port.s_nat(400 + i, 256, 384) for i in range(16)])
    check("S-nat 128x192", [port.s_nat(400 + i, 128, 192) for i in range(16)])
    check("S-iid 512x768", [port.s_iid(7000 + i, 512, 768) for i in range(12)])
    k = np.array(Image.open(os.path.join(os.path.dirname(__file__), "../../tests/golden/kodim01.png")).convert("RGB"))
    check("kodim01", [torch.from_numpy(k.transpose(2, 0, 1).copy())])
    check("S-nat 1365x2048", [port.s_nat(1000, 1365, 2048)])
