"""Desk check (numpy) of the two-columns-per-exchange algebra of panel_columns_smem (kernels_panel.cuh):
the formulas for H_{k+1} from the Gram entries a, b and the rows g, g+1 must reproduce plain
column-by-column Householder QR (dlarfg/dlarf) including the V'V entries used by dlarft."""
import numpy as np

def house_ref(F, stair):
    F = F.copy(); m, n = F.shape; g = 0; taus = []; 
    for k in range(n):
        if g >= m: break
        t = max(g + 1, stair[k])
        x = F[g + 1:t, k]; alpha = F[g, k]; ss = x @ x
        if t - g > 1 and ss != 0:
            beta = -np.copysign(np.sqrt(alpha * alpha + ss), alpha); tau = (beta - alpha) / beta; scale = 1 / (alpha - beta)
        else:
            beta, tau, scale = alpha, 0.0, 0.0
        v = np.zeros(m); v[g] = 1; v[g + 1:t] = x * scale
        F[:, k + 1:] -= np.outer(v, tau * (v @ F[:, k + 1:]))
        F[g, k] = beta; F[g + 1:t, k] = v[g + 1:t]
        taus.append(tau); g += 1
    return F, np.array(taus)

def house_two(F, stair, theta=1 / 32):
    F = F.copy(); m, n = F.shape; g = 0; taus = np.zeros(n); k = 0; ntwo = 0
    G = np.zeros((n, n))     # V'V strictly upper entries G[j, k] = v_j' v_k
    while k < n:
        if g >= m: break
        t = max(g + 1, stair[k]); la = k + 1 < n and g + 1 < m
        t1 = max(g + 2, stair[k + 1]) if la else t
        c, c1 = k, k + 1
        A = np.array([F[g + 2:t, c] @ F[g + 2:t, j] for j in range(n)])
        B = np.array([F[g + 2:t1, c1] @ F[g + 2:t1, j] for j in range(n)]) if la else np.zeros(n)
        r0 = F[g, :].copy(); r1 = F[g + 1, :].copy() if g + 1 < m else np.zeros(n)
        x0g1 = r1[c] if g + 1 < t else 0.0
        s = A + x0g1 * r1
        ss = s[c]; alpha = r0[c]
        if t - g > 1 and ss != 0:
            beta = -np.copysign(np.sqrt(alpha * alpha + ss), alpha); tau = (beta - alpha) / beta; scale = 1 / (alpha - beta)
        else:
            beta, tau, scale = alpha, 0.0, 0.0
        wv = tau * (r0 + scale * s); vg1 = scale * x0g1
        two = False
        if la:
            A_c, A_c1, B_c1, om = A[c], A[c1], B[c1], wv[c1]
            alpha1 = r1[c1] - vg1 * om; so = scale * om
            if t1 - (g + 1) <= 1:
                beta1, tau1, scale1, two = alpha1, 0.0, 0.0, True
            else:
                m1, m2 = 2 * so * A_c1, so * so * A_c
                ss1 = (B_c1 - m1) + m2; mag = B_c1 + abs(m1) + m2
                if ss1 >= theta * mag and ss1 > 1e-280:
                    beta1 = -np.copysign(np.sqrt(alpha1 * alpha1 + ss1), alpha1); tau1 = (beta1 - alpha1) / beta1
                    scale1 = 1 / (alpha1 - beta1); two = True
        lanes = np.arange(n)
        r1p = np.where(lanes > c, r1 - vg1 * wv, np.where(lanes == c, vg1, r1))
        w1 = np.zeros(n); d1 = np.zeros(n)
        if two:
            swv = np.where(lanes > c, scale * wv, 0.0)
            d1 = np.where(lanes == c, scale * (A_c1 - so * A_c), (B - swv * A_c1) - so * (A - swv * A_c))
            w1 = tau1 * (r1p + scale1 * d1)
        ee = t1 if two else t
        x0 = F[g + 2:ee, c].copy(); x1 = F[g + 2:ee, c1].copy() if la else None
        for j in range(c, n):
            ya, fa, fb = 1.0, 0.0, 0.0
            if j == c:
                if tau != 0: ya = scale
            elif two and j == c1:
                if tau1 != 0: ya = scale1
                fa = ya * so
            else:
                fb = scale1 * w1[j] if two else 0.0
                fa = scale * wv[j] - fb * (so if two else 0.0)
            if tau != 0 or two:
                y = F[g + 2:ee, j]
                F[g + 2:ee, j] = y * ya - x0 * fa - (x1 * fb if two else 0.0)
        # pivot rows
        for j in range(c, n):
            if j == c: F[g, j] = beta
            elif tau != 0: F[g, j] -= wv[j]
        if g + 1 < m:
            for j in range(c, n):
                if two and j == c1: F[g + 1, j] = beta1
                elif two and j > c1: F[g + 1, j] = r1p[j] - w1[j]
                elif tau != 0 and g + 1 < t: F[g + 1, j] = r1p[j]
        for j in range(c): G[j, c] = scale * s[j] + r0[j]
        taus[c] = tau; g += 1; k += 1
        if two:
            for j in range(c1): G[j, c1] = r1p[j] + scale1 * d1[j]
            taus[c1] = tau1; g += 1; k += 1; ntwo += 1
    return F, taus, G, ntwo

rng = np.random.default_rng(0)
for trial in range(200):
    m = int(rng.integers(3, 60)); n = int(rng.integers(1, min(m, 32) + 1))
    F = rng.standard_normal((m, n))
    # random nondecreasing staircase, zeros below it
    stair = np.sort(rng.integers(1, m + 1, size=n))
    if trial % 3 == 0: stair[:] = m
    for j in range(n): F[stair[j]:, j] = 0
    if trial % 5 == 0 and n > 2: F[:, 2] = F[:, 1] * 1.0000001 + 1e-9 * F[:, 2]   # near-dependent: guard must reject
    for j in range(n): F[stair[j]:, j] = 0
    Fr, tr = house_ref(F, stair)
    Ft, tt, G, ntwo = house_two(F, stair)
    nr = min(len(tr), n)
    scale = np.abs(Fr).max()
    err = np.abs(Fr - Ft).max() / scale
    neardep = (trial % 5 == 0 and n > 2)
    if not neardep:
        assert err < 1e-11, (trial, m, n, err)
        assert np.abs(tr - tt[:len(tr)]).max() < 1e-11, trial
    # backward error: Q R = F with Q = H_0 H_1 ... from the two-column variant's own V, tau
    nrf = len(tr)
    Vt = np.tril(Ft, -1)[:, :nrf].copy()
    for j in range(nrf): Vt[j, j] = 1
    Rt = np.triu(Ft)[:, :]
    QR = Rt.copy()
    for j in reversed(range(nrf)):
        v = Vt[:, j]; QR -= np.outer(v, tt[j] * (v @ QR))
    assert np.abs(QR - F).max() < 1e-13 * max(1, np.abs(F).max()) * m, (trial, np.abs(QR - F).max())
    if neardep: continue
    # G vs V'V
    V = np.tril(Ft, -1)[:, :n].copy()
    for j in range(min(m, n)): V[j, j] = 1
    VtV = V.T @ V
    for j in range(n):
        for kk in range(j + 1, min(n, len(tr))):
            assert abs(G[j, kk] - VtV[j, kk]) < 1e-9 * (1 + abs(VtV[j, kk])), (trial, j, kk, G[j, kk], VtV[j, kk])
print("ok")
