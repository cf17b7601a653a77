"""Independent restatement of the physics spec mps-planar-v3 in numpy float32 scalars.

TEST INFRASTRUCTURE.  This file is written from the spec text, not from oracle/mps_oracle.c, and shares no code
with it or with the CUDA kernels: Python objects instead of arrays of structs, exact fused multiply-add built
from double arithmetic (fma32 below, checked against libm's fmaf in tests/test_oracle_physics.py), colour-order
sweep written as a sort key instead of nested loops.  tests/golden/make_physics_kat.py runs the scenarios the
spec is pinned on through this file and commits their outputs; the oracle and the GPU must reproduce them bit
for bit.  Slow on purpose (a few thousand steps per second) — only small cases.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
BOX, DISC = 0, 1
PI_F = f32(3.14159265358979323846)
TWO_PI_F = f32(6.28318530717958647692)

F_RAN, F_OK, F_NOT_AT_REST, F_REST_CONTACT = 1 << 0, 1 << 1, 1 << 4, 1 << 5
F_BOUNDS, F_INFEASIBLE, F_BAD_CONTACT, F_OVERFLOW = 1 << 6, 1 << 7, 1 << 8, 1 << 9
F_ACTIVE_CONTACT, F_JAMMED = 1 << 10, 1 << 11


def fma32(a, b, c):
    """round_to_nearest_even_binary32(a * b + c) for binary32 inputs, one rounding."""
    a, b, c = float(a), float(b), float(c)
    p = a * b                      # exact: 24 + 24 significant bits fit in 53
    s = p + c                      # rounded to binary64 ...
    bb = s - p                     # ... with the exact remainder by TwoSum
    err = (p - (s - bb)) + (c - bb)
    r = f32(s)
    if err != 0.0 and math.isfinite(s):
        d = s - float(r)
        if d != 0.0:
            # s is not a binary32 number: the second rounding went wrong only if s sat exactly on the midpoint
            # of two neighbouring binary32 numbers, in which case the sign of the remainder decides
            other = np.nextafter(r, f32(np.inf) if d > 0 else f32(-np.inf))
            if d == (float(other) - float(r)) * 0.5:
                if (err > 0) == (d > 0):
                    r = other
        # s exactly representable: the remainder is below half an ulp of binary64, r stands
    return r


def fmin(a, b):
    return a if a < b else b


def fmax(a, b):
    return a if a > b else b


def sincos(a):
    """spec §3: Cody-Waite reduction by pi/2 in three parts, degree-7 / degree-8 minimax polynomials."""
    a = f32(a)
    two_over_pi = f32(0.63661977236758134)
    p_hi, p_mid, p_lo = f32(1.5703125), f32(4.837512969970703125e-4), f32(7.54978995489188216e-8)
    kf = f32(np.rint(a * two_over_pi))
    k = int(kf)
    r = fma32(-kf, p_hi, a)
    r = fma32(-kf, p_mid, r)
    r = fma32(-kf, p_lo, r)
    z = r * r
    sp = fma32(f32(-1.9515295891e-4), z, f32(8.3321608736e-3))
    sp = fma32(sp, z, f32(-1.6666654611e-1))
    sp = sp * z
    sp = fma32(sp, r, r)
    cp = fma32(f32(2.443315711809948e-5), z, f32(-1.388731625493765e-3))
    cp = fma32(cp, z, f32(4.166664568298827e-2))
    cp = cp * z
    cp = fma32(cp, z, fma32(f32(-0.5), z, f32(1.0)))
    q = k & 3
    return [(sp, cp), (cp, -sp), (-sp, -cp), (-cp, sp)][q]


def wrap(th):
    if th >= PI_F:
        return th - TWO_PI_F
    if th < -PI_F:
        return th + TWO_PI_F
    return th


class Shape:
    __slots__ = ("kind", "x", "y", "c", "s", "hx", "hy")

    def __init__(self, kind, x, y, c, s, hx, hy):
        self.kind, self.x, self.y, self.c, self.s, self.hx, self.hy = kind, f32(x), f32(y), f32(c), f32(s), f32(hx), f32(hy)


class Contact:
    """a touching pair: normal from a to b, 1 or 2 points with lever arms from both centres and separation"""

    def __init__(self, n, pts):
        self.n = n
        self.pts = pts  # [(rax, ray, rbx, rby, sep)]


def disc_disc(A, B):
    dx, dy = B.x - A.x, B.y - A.y
    d2 = fma32(dx, dx, dy * dy)
    R = A.hx + B.hx
    if d2 > R * R:
        return None
    d = np.sqrt(d2)
    n = (dx / d, dy / d) if d > f32(1e-9) else (f32(1), f32(0))
    sep = d - R
    h = fma32(f32(0.5), sep, A.hx)
    rax, ray = n[0] * h, n[1] * h
    return Contact(n, [(rax, ray, rax - dx, ray - dy, sep)])


def box_disc(X, D, box_is_a):
    ddx, ddy = D.x - X.x, D.y - X.y
    lx = fma32(ddx, X.c, ddy * X.s)
    ly = fma32(ddy, X.c, -(ddx * X.s))
    qx = fmin(fmax(lx, -X.hx), X.hx)
    qy = fmin(fmax(ly, -X.hy), X.hy)
    r = D.hx
    if qx == lx and qy == ly:  # centre inside the box: deepest axis
        px, py = X.hx - abs(lx), X.hy - abs(ly)
        if px < py:
            nl = (f32(1) if lx >= 0 else f32(-1), f32(0))
            sep = -(px + r)
            sp = (nl[0] * X.hx, ly)
        else:
            nl = (f32(0), f32(1) if ly >= 0 else f32(-1))
            sep = -(py + r)
            sp = (lx, nl[1] * X.hy)
    else:
        vx, vy = lx - qx, ly - qy
        d2 = fma32(vx, vx, vy * vy)
        if d2 > r * r:
            return None
        d = np.sqrt(d2)
        nl = (vx / d, vy / d)
        sep = d - r
        sp = (qx, qy)
    hs = f32(0.5) * sep
    cpx, cpy = fma32(nl[0], hs, sp[0]), fma32(nl[1], hs, sp[1])
    rx = (fma32(X.c, cpx, -(X.s * cpy)), fma32(X.s, cpx, X.c * cpy))       # from the box centre, world frame
    n = (fma32(X.c, nl[0], -(X.s * nl[1])), fma32(X.s, nl[0], X.c * nl[1]))
    rd = (rx[0] - ddx, rx[1] - ddy)                                          # from the disc centre
    if box_is_a:
        return Contact(n, [(rx[0], rx[1], rd[0], rd[1], sep)])
    return Contact((-n[0], -n[1]), [(rd[0], rd[1], rx[0], rx[1], sep)])


def box_box(A, B, ref_tol):
    dx, dy = B.x - A.x, B.y - A.y
    d00 = fma32(A.c, B.c, A.s * B.s)
    d01 = fma32(A.s, B.c, -(A.c * B.s))
    a00, a01 = abs(d00), abs(d01)
    pax, pay = fma32(dx, A.c, dy * A.s), fma32(dy, A.c, -(dx * A.s))
    pbx, pby = fma32(dx, B.c, dy * B.s), fma32(dy, B.c, -(dx * B.s))
    s_ax = abs(pax) - A.hx - fma32(B.hx, a00, B.hy * a01)
    s_ay = abs(pay) - A.hy - fma32(B.hx, a01, B.hy * a00)
    s_bx = abs(pbx) - B.hx - fma32(A.hx, a00, A.hy * a01)
    s_by = abs(pby) - B.hy - fma32(A.hx, a01, A.hy * a00)
    if s_ax > 0 or s_ay > 0 or s_bx > 0 or s_by > 0:
        return None
    a_y, b_y = s_ay > s_ax, s_by > s_bx
    sa, sb = (s_ay if a_y else s_ax), (s_by if b_y else s_bx)
    ref_b = sb > sa + ref_tol
    R, I = (B, A) if ref_b else (A, B)
    ddx, ddy = (-dx, -dy) if ref_b else (dx, dy)
    fy = b_y if ref_b else a_y
    u = (-R.s, R.c) if fy else (R.c, R.s)
    hu, ht = (R.hy, R.hx) if fy else (R.hx, R.hy)
    sg = f32(1) if fma32(ddx, u[0], ddy * u[1]) >= 0 else f32(-1)
    nr = (sg * u[0], sg * u[1])
    tr = (-nr[1], nr[0])
    e0 = fma32(nr[0], I.c, nr[1] * I.s)
    e1 = fma32(nr[1], I.c, -(nr[0] * I.s))
    if abs(e0) >= abs(e1):
        sm = f32(-1) if e0 > 0 else f32(1)
        m, hm, he = (sm * I.c, sm * I.s), I.hx, I.hy
    else:
        sm = f32(-1) if e1 > 0 else f32(1)
        m, hm, he = (sm * -I.s, sm * I.c), I.hy, I.hx
    cx, cy = fma32(m[0], hm, ddx), fma32(m[1], hm, ddy)
    ex, ey = -m[1], m[0]
    v1 = (fma32(ex, he, cx), fma32(ey, he, cy))
    v2 = (fma32(-ex, he, cx), fma32(-ey, he, cy))
    t = [fma32(v1[0], tr[0], v1[1] * tr[1]), fma32(v2[0], tr[0], v2[1] * tr[1])]
    s = [fma32(v1[0], nr[0], v1[1] * nr[1]) - hu, fma32(v2[0], nr[0], v2[1] * nr[1]) - hu]
    if (t[0] > ht and t[1] > ht) or (t[0] < -ht and t[1] < -ht):
        return None
    ct, cs = list(t), list(s)
    for k in (0, 1):  # clip end k of the incident edge to the side planes |t| <= ht
        o = 1 - k
        lim = ht if t[k] > ht else (-ht if t[k] < -ht else None)
        if lim is not None:
            al = (lim - t[k]) / (t[o] - t[k])
            cs[k] = fma32(al, s[o] - s[k], s[k])
            ct[k] = lim
    pts = []
    for k in (0, 1):
        if cs[k] <= 0:
            h = fma32(f32(0.5), cs[k], hu)
            rr = (fma32(tr[0], ct[k], nr[0] * h), fma32(tr[1], ct[k], nr[1] * h))
            ri = (rr[0] - ddx, rr[1] - ddy)
            pts.append((ri[0], ri[1], rr[0], rr[1], cs[k]) if ref_b else (rr[0], rr[1], ri[0], ri[1], cs[k]))
    if not pts:
        return None
    return Contact((-nr[0], -nr[1]) if ref_b else nr, pts)


def collide(A, B, ref_tol):
    if A.kind == BOX and B.kind == BOX:
        return box_box(A, B, ref_tol)
    if A.kind == BOX:
        return box_disc(A, B, True)
    if B.kind == BOX:
        return box_disc(B, A, False)
    return disc_disc(A, B)


class Rows:
    """§3.4: the solver's view of a touching pair"""


class World:
    def __init__(self, sc):
        """sc: anything with the fields of struct mps_scene (include/mps_b200.h)"""
        self.sc = sc
        self.B, self.S, self.robot = int(sc.n_bodies), int(sc.n_statics), int(sc.robot_id)
        B = self.B
        g = lambda arr, i: f32(arr[i])  # noqa: E731
        self.dt = f32(sc.dt)
        self.mass = [g(sc.mass, i) for i in range(B)]
        self.inertia = [g(sc.inertia, i) for i in range(B)]
        self.inv_m, self.inv_i, self.lin_cap, self.ang_cap = [], [], [], []
        for i in range(B):
            if i == self.robot:
                vals = (f32(0),) * 4
            else:
                f = g(sc.mu_ground, i) * (self.mass[i] * f32(sc.gravity))
                vals = (f32(1) / self.mass[i], f32(1) / self.inertia[i], self.dt * f, self.dt * (f * g(sc.fr_radius, i)))
            self.inv_m.append(vals[0]); self.inv_i.append(vals[1]); self.lin_cap.append(vals[2]); self.ang_cap.append(vals[3])
        self.beta_dt = f32(sc.baumgarte) * (f32(1) / self.dt)
        self.ref_tol = f32(0.1) * f32(sc.slop)
        self.statics = []
        for k in range(self.S):
            s, c = sincos(sc.s_theta[k])
            self.statics.append(Shape(int(sc.s_shape[k]), sc.s_x[k], sc.s_y[k], c, s, sc.s_hx[k], sc.s_hy[k]))
        self.x = [f32(0)] * B; self.y = [f32(0)] * B; self.th = [f32(0)] * B
        self.c = [f32(1)] * B; self.s = [f32(0)] * B
        self.vx = [f32(0)] * B; self.vy = [f32(0)] * B; self.w = [f32(0)] * B
        self.overflow = False
        self.max_pen = f32(0)

    # -- state ----------------------------------------------------------------------------------------------
    def set_state(self, state, vel=None):
        for i in range(self.B):
            self.x[i], self.y[i] = f32(state[3 * i]), f32(state[3 * i + 1])
            self.th[i] = wrap(f32(state[3 * i + 2]))
            self.s[i], self.c[i] = sincos(self.th[i])
            self.vx[i] = f32(vel[3 * i]) if vel is not None else f32(0)
            self.vy[i] = f32(vel[3 * i + 1]) if vel is not None else f32(0)
            self.w[i] = f32(vel[3 * i + 2]) if vel is not None else f32(0)

    def pose(self):
        return np.array([v for i in range(self.B) for v in (self.x[i], self.y[i], self.th[i])], np.float32)

    def twist(self):
        return np.array([v for i in range(self.B) for v in (self.vx[i], self.vy[i], self.w[i])], np.float32)

    def shape(self, j):
        if j >= self.B:
            return self.statics[j - self.B]
        sc = self.sc
        return Shape(int(sc.shape[j]), self.x[j], self.y[j], self.c[j], self.s[j], sc.hx[j], sc.hy[j])

    def allowed(self, a, b):
        if b >= self.B:
            return not ((int(self.sc.static_forbidden) >> a) & 1)
        return not ((int(self.sc.pair_forbidden[a]) >> b) & 1)

    # -- §3.2 step 2 ----------------------------------------------------------------------------------------
    def detect(self):
        """all touching pairs (exact geometry on every pair: no broad phase is needed for the result)"""
        self.max_pen = f32(0)
        bad = False
        found = []
        for i in range(self.B):
            A = self.shape(i)
            for j in range(i + 1, self.B + self.S):
                ct = collide(A, self.shape(j), self.ref_tol)
                if ct is None:
                    continue
                for p in ct.pts:
                    if -p[4] > self.max_pen:
                        self.max_pen = -p[4]
                if not self.allowed(i, j):
                    bad = True
                if len(found) >= int(self.sc.max_manifolds):
                    self.overflow = True
                    continue
                found.append(self.rows(i, j, ct))
        return found, bad

    # -- §3.4 -----------------------------------------------------------------------------------------------
    def rows(self, a, b, ct):
        sc = self.sc
        r = Rows()
        r.a, r.b = a, b
        r.n = ct.n
        r.t = (ct.n[1], -ct.n[0])
        r.ma, r.ia = self.inv_m[a], self.inv_i[a]
        if b < self.B:
            r.mb, r.ib, mub = self.inv_m[b], self.inv_i[b], f32(sc.mu_contact[b])
        else:
            r.mb, r.ib, mub = f32(0), f32(0), f32(sc.s_mu_contact[b - self.B])
        r.mu = np.sqrt(f32(sc.mu_contact[a]) * mub)
        z = f32(0)
        r.rna, r.rnb, r.rta, r.rtb = [z, z], [z, z], [z, z], [z, z]
        r.kn, r.kt, r.kb = [z, z], [z, z], [z, z]
        r.acc_n, r.acc_t = [z, z], [z, z]
        ia_rn, ib_rn, ia_rt, ib_rt = [z, z], [z, z], [z, z], [z, z]
        msum = r.ma + r.mb
        for k, (rax, ray, rbx, rby, sep) in enumerate(ct.pts):
            r.rna[k] = fma32(rax, r.n[1], -(ray * r.n[0]))
            r.rnb[k] = fma32(rbx, r.n[1], -(rby * r.n[0]))
            r.rta[k] = fma32(rax, r.t[1], -(ray * r.t[0]))
            r.rtb[k] = fma32(rbx, r.t[1], -(rby * r.t[0]))
            ia_rn[k], ib_rn[k] = r.ia * r.rna[k], r.ib * r.rnb[k]
            ia_rt[k], ib_rt[k] = r.ia * r.rta[k], r.ib * r.rtb[k]
            kn = fma32(ib_rn[k], r.rnb[k], fma32(ia_rn[k], r.rna[k], msum))
            kt = fma32(ib_rt[k], r.rtb[k], fma32(ia_rt[k], r.rta[k], msum))
            r.kn[k] = f32(1) / kn if kn > 0 else z
            r.kt[k] = f32(1) / kt if kt > 0 else z
            pen = -sep - f32(sc.slop)
            bias = fmin(self.beta_dt * pen, f32(sc.max_bias_vel)) if pen > 0 else z
            r.kb[k] = r.kn[k] * bias
        # couplings K_sr between the rows of the manifold
        r.k_t0n0 = fma32(ia_rt[0], r.rna[0], ib_rt[0] * r.rnb[0])
        r.k_t0t1 = r.k_n0n1 = r.k_t0n1 = r.k_t1n0 = r.k_t1n1 = z
        if len(ct.pts) > 1:
            r.k_t0t1 = fma32(ib_rt[0], r.rtb[1], fma32(ia_rt[0], r.rta[1], msum))
            r.k_n0n1 = fma32(ib_rn[0], r.rnb[1], fma32(ia_rn[0], r.rna[1], msum))
            r.k_t0n1 = fma32(ia_rt[0], r.rna[1], ib_rt[0] * r.rnb[1])
            r.k_t1n0 = fma32(ia_rt[1], r.rna[0], ib_rt[1] * r.rnb[0])
            r.k_t1n1 = fma32(ia_rt[1], r.rna[1], ib_rt[1] * r.rnb[1])
        return r

    def visit(self, r):
        a, b = r.a, r.b
        va, wa = (self.vx[a], self.vy[a]), self.w[a]
        vb, wb = ((self.vx[b], self.vy[b]), self.w[b]) if b < self.B else ((f32(0), f32(0)), f32(0))
        dv = (vb[0] - va[0], vb[1] - va[1])

        def along(d, cb, ca):
            return fma32(dv[0], d[0], fma32(dv[1], d[1], fma32(wb, cb, -(wa * ca))))

        ut = [along(r.t, r.rtb[k], r.rta[k]) for k in (0, 1)]
        un = [along(r.n, r.rnb[k], r.rna[k]) for k in (0, 1)]

        def friction(k, u):
            lim = r.mu * r.acc_n[k]
            d = fmin(fmax(r.kt[k] * -u, -lim - r.acc_t[k]), lim - r.acc_t[k])
            r.acc_t[k] = r.acc_t[k] + d
            return d

        def normal(k, u):
            d = fmax(fma32(-r.kn[k], u, r.kb[k]), -r.acc_n[k])
            r.acc_n[k] = r.acc_n[k] + d
            return d

        dt0 = friction(0, ut[0])
        dt1 = friction(1, fma32(r.k_t0t1, dt0, ut[1]))
        dn0 = normal(0, fma32(r.k_t1n0, dt1, fma32(r.k_t0n0, dt0, un[0])))
        dn1 = normal(1, fma32(r.k_n0n1, dn0, fma32(r.k_t1n1, dt1, fma32(r.k_t0n1, dt0, un[1]))))
        st = dt0 + dt1
        ra = fma32(r.rna[1], dn1, fma32(r.rta[1], dt1, fma32(r.rna[0], dn0, r.rta[0] * dt0)))
        rb = fma32(r.rnb[1], dn1, fma32(r.rtb[1], dt1, fma32(r.rnb[0], dn0, r.rtb[0] * dt0)))
        px = fma32(r.n[0], dn1, fma32(r.n[0], dn0, r.t[0] * st))
        py = fma32(r.n[1], dn1, fma32(r.n[1], dn0, r.t[1] * st))
        self.vx[a], self.vy[a], self.w[a] = fma32(-r.ma, px, va[0]), fma32(-r.ma, py, va[1]), fma32(-r.ia, ra, wa)
        if b < self.B:
            self.vx[b], self.vy[b], self.w[b] = fma32(r.mb, px, vb[0]), fma32(r.mb, py, vb[1]), fma32(r.ib, rb, wb)

    # -- §3.2 -----------------------------------------------------------------------------------------------
    def step(self, u):
        """one step with robot twist u; returns True when a forbidden pair touched"""
        sc, B = self.sc, self.B
        self.vx[self.robot], self.vy[self.robot], self.w[self.robot] = f32(u[0]), f32(u[1]), f32(u[2])
        found, bad = self.detect()
        self.n_manifolds = len(found)
        if self.overflow:
            return bad
        for i in range(B):
            if i == self.robot:
                continue
            imp = -self.inertia[i] * self.w[i]
            impc = fmin(fmax(imp, -self.ang_cap[i]), self.ang_cap[i])
            self.w[i] = f32(0) if impc == imp else fma32(self.inv_i[i], impc, self.w[i])
            ix, iy = -self.mass[i] * self.vx[i], -self.mass[i] * self.vy[i]
            n2 = fma32(ix, ix, iy * iy)
            cap = self.lin_cap[i]
            if n2 > cap * cap:
                k = cap / np.sqrt(n2)
                self.vx[i] = fma32(self.inv_m[i], ix * k, self.vx[i])
                self.vy[i] = fma32(self.inv_m[i], iy * k, self.vy[i])
            else:
                self.vx[i] = self.vy[i] = f32(0)
        # colour order: body pairs by ((i + j) mod B, i), then the pairs with static s by (B + s, i)
        order = sorted(found, key=lambda r: (((r.a + r.b) % B) if r.b < B else r.b, r.a))
        for _ in range(int(sc.vel_iters)):
            for r in order:
                self.visit(r)
        for i in range(B):
            self.x[i] = fma32(self.dt, self.vx[i], self.x[i])
            self.y[i] = fma32(self.dt, self.vy[i], self.y[i])
            self.th[i] = wrap(fma32(self.dt, self.w[i], self.th[i]))
            self.s[i], self.c[i] = sincos(self.th[i])
        return bad

    def at_rest(self):
        v2 = f32(self.sc.v_rest) * f32(self.sc.v_rest)
        return all(not (fma32(self.vx[i], self.vx[i], self.vy[i] * self.vy[i]) > v2) and not (abs(self.w[i]) > f32(self.sc.w_rest))
                   for i in range(self.B))


def normalize_d(v):
    """util/Math.cpp:8-16 in double"""
    v = math.fmod(v, 2.0 * math.pi)
    if v < -math.pi:
        v += 2.0 * math.pi
    elif v >= math.pi:
        v -= 2.0 * math.pi
    return v


def ramp_of(params, alim):
    """RampVelocityControl.cpp:98-158 (same restatement as make_kat.py)"""
    v = [f32(params[i]) for i in range(3)]
    t_pl = f32(params[3])
    t_acc = f32(abs(v[0] / f32(alim[0])))
    for i in range(3):
        t = f32(abs(v[i] / f32(alim[i])))
        t_acc = t if t_acc < t else t_acc
    a = [f32((f32(1.0) / t_acc) * v[i]) for i in range(3)] if t_acc != 0 else [f32(0)] * 3
    return v, a, t_acc, t_pl


def ramp_velocity(v, a, t_acc, t_pl, t):
    t = f32(t)
    if t < t_acc:
        return [f32(t * a[i]) for i in range(3)]
    if t < f32(t_acc + t_pl):
        return list(v)
    if float(t) < 2.0 * float(t_acc) + float(t_pl):
        s = f32(f32(t - t_acc) - t_pl)
        return [f32(v[i] - f32(s * a[i])) for i in range(3)]
    return [f32(0)] * 3


def in_bounds(sc, state):
    eps = np.finfo(np.float64).eps
    for i in range(int(sc.n_bodies)):
        x, y, th = float(state[3 * i]), float(state[3 * i + 1]), float(state[3 * i + 2])
        if x - eps > float(sc.x_max) or x + eps < float(sc.x_min) or y - eps > float(sc.y_max) or y + eps < float(sc.y_min):
            return False
        if not (th < math.pi + eps and th > -math.pi - eps):
            return False
    return True


def is_valid(sc, state):
    """§3.7 -> flags (0 = valid)"""
    if not in_bounds(sc, state):
        return F_BOUNDS
    wd = World(sc)
    wd.set_state(state)
    _, bad = wd.detect()
    f = 0
    if wd.max_pen > f32(sc.pen_max):
        f |= F_INFEASIBLE
    if bad:
        f |= F_BAD_CONTACT
    if wd.overflow:
        f |= F_OVERFLOW
    return f


def propagate(sc, state, control):
    """§3.6 -> (ok, state_out[3B], rest_time, flags, n_steps)"""
    wd = World(sc)
    wd.set_state(state)
    v, a, t_acc, t_pl = ramp_of(control, sc.accel_limits)
    T = f32(f32(f32(2.0) * t_acc + t_pl) + f32(0.0))
    flags, ok, steps = F_RAN, True, 0
    time = f32(0)
    while time < T and ok:
        bad = wd.step(ramp_velocity(v, a, t_acc, t_pl, time))
        time = time + wd.dt
        steps += 1
        if wd.overflow:
            flags |= F_OVERFLOW
            ok = False
        if bad and int(sc.strict_active_contacts):
            flags |= F_ACTIVE_CONTACT
            ok = False
    rest = f32(0)
    if ok:
        jam = 0
        while rest <= f32(sc.t_max) and not wd.at_rest() and ok:
            bad = wd.step((0, 0, 0))
            ok = not bad
            rest = rest + wd.dt
            steps += 1
            if bad:
                flags |= F_REST_CONTACT
            if wd.overflow:
                flags |= F_OVERFLOW
                ok = False
            jam = jam + 1 if wd.max_pen > f32(1.25) * f32(sc.slop) else 0
            if int(sc.jam_steps) > 0 and jam >= int(sc.jam_steps):
                flags |= F_JAMMED
                ok = False
        rested = bool(rest <= f32(sc.t_max)) and wd.at_rest()
        if not rested:
            flags |= F_NOT_AT_REST
        ok = ok and rested
    out = wd.pose()
    for i in range(wd.B):
        out[3 * i + 2] = f32(normalize_d(float(out[3 * i + 2])))
    vf = is_valid(sc, out)
    flags |= vf
    ok = ok and vf == 0
    if ok:
        flags |= F_OK
    return ok, out, rest, flags, steps
