"""Integration strategies used by the CPU oracle.

TEST INFRASTRUCTURE ONLY -- never imported by chomp_b200/.

Two strategies share one call signature ``integ(f, a, b, rtol, breaks=(),
singular=())``:

``Romberg``
    The reference's own rule: every integral in CHOMP is
    ``scipy.integrate.romberg(f, a, b, vec_func=True, tol=1.48e-32, rtol=<module
    precision>, divmax=20)`` (66 call sites, e.g. cosmology.py:633-638,
    halo.py:910-916, kernel.py:698-704, correlation.py:253-259).  ``breaks`` and
    ``singular`` are ignored, exactly as the reference ignores kinks.

``Tight``
    The converged value of the *same* integrand: composite Gauss-Legendre
    between every point where the integrand is not smooth (``breaks``: spline
    knots, HOD kinks, exponent switches) with geometric grading towards
    ``singular`` points (algebraic end-point behaviour such as
    ((M-M0)/M1')**alpha at hod.py:226-230).  This is the comparison target for
    the 1e-5 parity bar: tightening the reference's
    Romberg rtol instead stalls at divmax=20 on the kinks.
"""
import numpy as np

from . import _py2compat

_GL_CACHE = {}


def gl_nodes(n):
    """Gauss-Legendre nodes/weights on [-1, 1] (cached)."""
    if n not in _GL_CACHE:
        _GL_CACHE[n] = np.polynomial.legendre.leggauss(n)
    return _GL_CACHE[n]


def gl_panels(edges, n):
    """Nodes and weights of an n-point GL rule on each [edges[i], edges[i+1]]."""
    x, w = gl_nodes(n)
    edges = np.asarray(edges, dtype=float)
    mid = 0.5*(edges[1:] + edges[:-1])
    half = 0.5*(edges[1:] - edges[:-1])
    return ((mid[:, None] + half[:, None]*x[None, :]).ravel(),
            (half[:, None]*w[None, :]).ravel())


class Romberg(object):
    name = "romberg"

    def __init__(self, divmax=20, tol=1.48e-32):
        self.divmax = divmax
        self.tol = tol

    def __call__(self, f, a, b, rtol, breaks=(), singular=(), args=()):
        return _py2compat.romberg(f, a, b, args=args, tol=self.tol, rtol=rtol,
                                  divmax=self.divmax, vec_func=True)


class Tight(object):
    name = "tight"

    def __init__(self, order=40, grade_levels=24, grade_ratio=0.2):
        self.order = order
        self.grade_levels = grade_levels
        self.grade_ratio = grade_ratio

    def edges(self, a, b, breaks=(), singular=()):
        pts = [a, b]
        for p in breaks:
            if a < p < b:
                pts.append(float(p))
        sing = [float(s) for s in singular if a <= s <= b]
        for s in sing:
            if a < s < b:
                pts.append(s)
        pts = np.unique(np.asarray(pts, dtype=float))
        # geometric grading on both sides of each singular point
        extra = []
        for s in sing:
            i = int(np.argmin(np.abs(pts - s)))
            for j in (i - 1, i + 1):
                if 0 <= j < pts.size:
                    d = pts[j] - s
                    for lev in range(1, self.grade_levels + 1):
                        extra.append(s + d*self.grade_ratio**lev)
        if extra:
            pts = np.unique(np.concatenate([pts, np.asarray(extra)]))
        return pts

    def __call__(self, f, a, b, rtol, breaks=(), singular=(), args=()):
        if b == a:
            return 0.0
        sign = 1.0
        if b < a:
            a, b, sign = b, a, -1.0
        x, w = gl_panels(self.edges(a, b, breaks, singular), self.order)
        _py2compat.N_EVAL[0] += x.size
        return sign*float(np.sum(w*f(x, *args)))


class Fixed(Tight):
    """The fixed-order rule the CUDA kernels use, for tuning node counts on the
    CPU: ``order``-point Gauss-Legendre on every panel between ``breaks``; on
    the panels touching a ``singular`` point the substitution
    x = s + (e - s) t**sing_power removes the algebraic end-point behaviour."""
    name = "tight"   # asks chomp_oracle for the same break lists as Tight

    def __init__(self, order=4, sing_power=3, sing_order=None):
        Tight.__init__(self, order, 0, 0.5)
        self.sing_power = sing_power
        self.sing_order = sing_order or order

    def __call__(self, f, a, b, rtol, breaks=(), singular=(), args=()):
        if b == a:
            return 0.0
        edges = self.edges(a, b, breaks, singular)
        sing = set(float(s) for s in singular if a <= s <= b)
        x, w = gl_nodes(self.order)
        xs, ws = gl_nodes(self.sing_order)
        t, tw = 0.5*(xs + 1.0), 0.5*ws
        total = 0.0
        for lo, hi in zip(edges[:-1], edges[1:]):
            if lo in sing or hi in sing:
                s, e = (lo, hi) if lo in sing else (hi, lo)
                p = self.sing_power
                xx = s + (e - s)*t**p
                ww = (e - s)*p*t**(p - 1)*tw
                total += float(np.sum(ww*f(xx, *args)))*(1.0 if e > s else -1.0)
            else:
                xx = 0.5*(lo + hi) + 0.5*(hi - lo)*x
                total += 0.5*(hi - lo)*float(np.sum(w*f(xx, *args)))
        return total
