"""numpy MODEL of the register-resident partitioned-Thomas kernel (Layout B) -- test infrastructure.

It mirrors, operation for operation, what kwinto-cuda_b200/csrc/fd1d_reg.cuh does on the device
(constant-dt hoisted LU in pivot-scaled variables, v' = max(D*u~ - v, p), per-thread chunks of M
nodes, Kogge-Stone carries with precomputed multipliers, Moebius-composed pivots), so the
numerical design can be checked against the oracle on the CPU before GPU time is spent
(tests/test_scheme_model.py).  The product never imports this file.
"""
import numpy as np


def coefficients(x, a0, ax, axx, dt):
    """bl, b, bu of B = 1 - theta*dt*A exactly as the reference assembles them
    (src/Math/kwFd1d.cpp:73-114) but with the constant dt = t/(tDim-1)."""
    n = x.shape[0]
    th = 0.5
    bl = np.zeros(n)
    b = np.zeros(n)
    bu = np.zeros(n)
    inv = 1.0 / (x[1] - x[0])
    b[0] = 1 - th * dt * (a0 - inv * ax)
    bu[0] = -th * dt * (inv * ax)
    iu = 1.0 / (x[2:] - x[1:-1])
    im = 1.0 / (x[2:] - x[:-2])
    idn = 1.0 / (x[1:-1] - x[:-2])
    i2u = 2.0 * iu * im
    i2m = 2.0 * idn * iu
    i2l = 2.0 * idn * im
    bl[1:-1] = -th * dt * (-im * ax + i2l * axx)
    b[1:-1] = 1 - th * dt * (a0 - i2m * axx)
    bu[1:-1] = -th * dt * (im * ax + i2u * axx)
    inv = 1.0 / (x[-1] - x[-2])
    bl[-1] = -th * dt * (-inv * ax)
    b[-1] = 1 - th * dt * (a0 + inv * ax)
    return bl, b, bu


def pivots_serial(bl, b, bu):
    n = b.shape[0]
    beta = np.empty(n)
    beta[0] = b[0]
    for j in range(1, n):
        beta[j] = b[j] - bl[j] * (bu[j - 1] / beta[j - 1])
    return beta


def pivots_moebius(bl, b, bu, M):
    """beta_j = b_j - c_j / beta_{j-1}, c_j = bl_j*bu_{j-1}: each thread composes the M Moebius maps of
    its chunk into one normalised 2x2 matrix, the chunk-boundary pivots come from a scan over those
    matrices, then every thread re-runs the recurrence inside its chunk from the true incoming pivot."""
    n = b.shape[0]
    P = n // M
    c = np.zeros(n)
    c[1:] = bl[1:] * bu[:-1]
    bb = b.reshape(P, M)
    cc = c.reshape(P, M)
    # chunk matrix [[m00, m01], [m10, m11]] acting on (num; den)
    m00 = np.ones(P); m01 = np.zeros(P); m10 = np.zeros(P); m11 = np.ones(P)
    for i in range(M):
        n00 = bb[:, i] * m00 - cc[:, i] * m10
        n01 = bb[:, i] * m01 - cc[:, i] * m11
        m10, m11 = m00, m01
        m00, m01 = n00, n01
        s = 1.0 / np.maximum(np.maximum(np.abs(m00), np.abs(m01)), np.maximum(np.abs(m10), np.abs(m11)))
        m00, m01, m10, m11 = m00 * s, m01 * s, m10 * s, m11 * s
    # serial walk over chunk matrices (device: warp scan); homogeneous (num, den), start = infinity
    num, den = 1.0, 0.0
    beta_in = np.empty(P)  # pivot of the node just before each chunk (inf for chunk 0)
    for k in range(P):
        beta_in[k] = np.inf if den == 0.0 else num / den
        nn = m00[k] * num + m01[k] * den
        dd = m10[k] * num + m11[k] * den
        s = 1.0 / max(abs(nn), abs(dd))
        num, den = nn * s, dd * s
    beta = np.empty((P, M))
    prev = beta_in.copy()
    for i in range(M):
        cur = bb[:, i] - cc[:, i] / prev
        beta[:, i] = cur
        prev = cur
    return beta.reshape(n)


def _shift_up(a, d, W):
    """__shfl_up_sync by d inside W-lane warps; lanes < d keep their own value."""
    r = a.reshape(-1, W).copy()
    src = a.reshape(-1, W)
    r[:, d:] = src[:, :-d]
    return r.reshape(a.shape)


def _shift_down(a, d, W):
    r = a.reshape(-1, W).copy()
    src = a.reshape(-1, W)
    r[:, :-d] = src[:, d:]
    return r.reshape(a.shape)


class PartitionedModel:
    def __init__(self, x, payoff, a0, ax, axx, dt, american, M=8, W=32, moebius=True):
        n0 = x.shape[0]
        self.n0 = n0
        P = -(-n0 // M)
        P = -(-P // W) * W if P > W else P  # whole warps (small grids: one partial warp is fine in the model)
        W = min(W, P)
        n = P * M
        self.P, self.M, self.W, self.n = P, M, W, n
        bl, b, bu = coefficients(x, a0, ax, axx, dt)
        # pad with identity rows: v stays 0 there and nothing couples back (bu[n0-1] = 0 already)
        pad = n - n0
        bl = np.concatenate([bl, np.zeros(pad)])
        b = np.concatenate([b, np.ones(pad)])
        bu = np.concatenate([bu, np.zeros(pad)])
        beta = pivots_moebius(bl, b, bu, M) if moebius else pivots_serial(bl, b, bu)
        self.beta = beta
        ib = 1.0 / beta
        at = np.zeros(n); gt = np.zeros(n)
        at[1:] = -bl[1:] * ib[:-1]
        gt[:-1] = -bu[:-1] * ib[1:]
        D = 2.0 * ib
        D[n0:] = 0.0
        proj = np.full(n, -np.inf)
        if american:
            proj[: n0 - 1] = payoff[: n0 - 1]
        self.v = np.concatenate([payoff, np.zeros(pad)]).reshape(P, M).copy()
        self.proj = proj.reshape(P, M)
        a = at.reshape(P, M); g = gt.reshape(P, M); D = D.reshape(P, M)
        self.a, self.g, self.D = a, g, D
        Pp = np.empty((P, M)); Q = np.empty((P, M)); R = np.empty((P, M))
        Pp[:, 0] = a[:, 0]
        for i in range(1, M):
            Pp[:, i] = a[:, i] * Pp[:, i - 1]
        Q[:, M - 1] = g[:, M - 1]
        for i in range(M - 2, -1, -1):
            Q[:, i] = g[:, i] * Q[:, i + 1]
        R[:, M - 1] = Pp[:, M - 1]
        for i in range(M - 2, -1, -1):
            R[:, i] = Pp[:, i] + g[:, i] * R[:, i + 1]
        self.R0 = R[:, 0].copy()
        self.DR = D * R
        self.DQ = D * Q
        lane = np.arange(P) % W
        self.lane = lane
        self.nlev = int(np.log2(W))
        # forward multipliers
        A = Pp[:, M - 1].copy()
        self.Af = []
        for d in range(self.nlev):
            s = 1 << d
            self.Af.append(np.where(lane >= s, A, 0.0))
            A = np.where(lane >= s, A * _shift_up(A, s, W), A)
        self.PWf = A  # product warp-start..k
        PWex = _shift_up(A, 1, W)
        self.PWexf = np.where(lane == 0, 1.0, PWex)
        self.AWf = A.reshape(-1, W)[:, W - 1].copy()
        # backward multipliers
        G = Q[:, 0].copy()
        self.Gb = []
        for d in range(self.nlev):
            s = 1 << d
            self.Gb.append(np.where(lane < W - s, G, 0.0))
            G = np.where(lane < W - s, G * _shift_down(G, s, W), G)
        PWex = _shift_down(G, 1, W)
        self.PWexb = np.where(lane == W - 1, 1.0, PWex)
        self.AWb = G.reshape(-1, W)[:, 0].copy()

    def step(self):
        P, M, W = self.P, self.M, self.W
        v, a, g = self.v, self.a, self.g
        nw = P // W
        # forward local
        y = np.empty((P, M))
        y[:, 0] = v[:, 0]
        for i in range(1, M):
            y[:, i] = a[:, i] * y[:, i - 1] + v[:, i]
        S = y[:, M - 1].copy()
        for d in range(self.nlev):
            S = self.Af[d] * _shift_up(S, 1 << d, W) + S
        Z = S.reshape(nw, W)[:, W - 1]
        X = np.zeros(nw)
        for w in range(1, nw):
            X[w] = self.AWf[w - 1] * X[w - 1] + Z[w - 1]
        Sm1 = np.where(self.lane == 0, 0.0, _shift_up(S, 1, W))
        Yin = self.PWexf * np.repeat(X, W) + Sm1
        # backward local (on the local forward result)
        u = np.empty((P, M))
        u[:, M - 1] = y[:, M - 1]
        for i in range(M - 2, -1, -1):
            u[:, i] = g[:, i] * u[:, i + 1] + y[:, i]
        T = self.R0 * Yin + u[:, 0]
        for d in range(self.nlev):
            T = self.Gb[d] * _shift_down(T, 1 << d, W) + T
        Z = T.reshape(nw, W)[:, 0]
        X = np.zeros(nw)
        for w in range(nw - 2, -1, -1):
            X[w] = self.AWb[w + 1] * X[w + 1] + Z[w + 1]
        Tp1 = np.where(self.lane == W - 1, 0.0, _shift_down(T, 1, W))
        Uin = self.PWexb * np.repeat(X, W) + Tp1
        r = self.D * u - v
        r = self.DR * Yin[:, None] + r
        r = self.DQ * Uin[:, None] + r
        self.v = np.maximum(r, self.proj)

    def solution(self):
        return self.v.reshape(-1)[: self.n0]


def price_with_model(oracle, option, tdim, xdim, density=0.25, scale=50.0, M=8, moebius=True):
    """One option through the model; x grid and payoff taken from the oracle's libm so that only the
    ALGORITHMIC differences (hoisted LU, scaled variables, partition carries) are measured."""
    t, k, z, r, q, s, e, w = (option[f] for f in ("t", "k", "z", "r", "q", "s", "e", "w"))
    x = oracle.x_grid(float(z), float(t), xdim, density, scale)
    pay = np.maximum(0.0, 1.0 - np.exp(x)) if w < 0 else np.maximum(0.0, np.exp(x) - 1.0)
    a0 = -r
    ax = r - q - z * z / 2
    axx = z * z / 2
    dt = t / (tdim - 1)
    mdl = PartitionedModel(x, pay, a0, ax, axx, dt, bool(e), M=M, moebius=moebius)
    for _ in range(tdim - 1):
        mdl.step()
    v = mdl.solution()
    xq = np.log(s / k)
    xi = int(np.searchsorted(x, xq, side="left"))
    val = ((x[xi] - xq) * v[xi - 1] + (xq - x[xi - 1]) * v[xi]) / (x[xi] - x[xi - 1])
    return k * val, x, v, mdl
