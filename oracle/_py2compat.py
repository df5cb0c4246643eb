"""Compatibility shims injected into the generated reference copy (oracle/_ref).

TEST INFRASTRUCTURE ONLY.  Nothing under chomp_b200/ may import this.

Two things the reference (Python 2.7 + old scipy) relies on no longer exist:

* ``scipy.integrate.romberg`` (removed in scipy 1.15).  ``romberg`` below is a
  restatement of the published algorithm of scipy <= 1.14
  (``scipy/integrate/_quadrature.py``; third-party, BSD, *not* part of
  /root/reference): a trapezoid rule on 1, 2, 4, ... panels that reuses the
  previous ordinates, a Richardson table
  ``R[i][k] = (4**k R[i][k-1] - R[i-1][k-1]) / (4**k - 1)`` and the stopping rule
  ``|R[i][i] - R[i-1][i-1]| < tol  or  < rtol*|R[i][i]|`` with at most
  ``divmax`` halvings (after which the last estimate is returned).
  The reference always calls it with tol=1.48e-32, divmax=20, vec_func=True
  (e.g. cosmology.py:106-110, halo.py:910-916, correlation.py:253-259).
* Python-2 integer division (``1/b`` in kernel.py:164-168 when ``b`` is an int).
"""
import numpy as _np

#: counters used by bench.py / tests to report integrand evaluations
N_EVAL = [0]


def _midpoint_sum(func, args, lo, hi, panels):
    """Sum of ordinates that the trapezoid rule with ``panels`` panels adds to
    the rule with ``panels/2`` panels (the end-point average for panels==1)."""
    if panels == 1:
        N_EVAL[0] += 2
        return 0.5*(func(lo, *args) + func(hi, *args))
    n_new = panels//2
    h = float(hi - lo)/n_new
    pts = lo + 0.5*h + h*_np.arange(n_new)
    N_EVAL[0] += n_new
    return _np.sum(func(pts, *args), axis=0)


def romberg(function, a, b, args=(), tol=1.48e-8, rtol=1.48e-8, show=False,
            divmax=10, vec_func=False):
    if _np.isinf(a) or _np.isinf(b):
        raise ValueError("Romberg integration only available for finite limits.")
    if vec_func:
        func = function
    else:
        def func(x, *fargs):
            if _np.isscalar(x):
                return function(x, *fargs)
            return _np.array([function(xi, *fargs) for xi in x])
    if not isinstance(args, tuple):
        args = (args,)
    span = b - a
    panels = 1
    ordsum = _midpoint_sum(func, args, a, b, panels)
    prev_row = [span*ordsum]
    result = prev_row[0]
    for i in range(1, divmax + 1):
        panels *= 2
        ordsum = ordsum + _midpoint_sum(func, args, a, b, panels)
        row = [span*ordsum/panels]
        for k in range(i):
            p4 = 4.0**(k + 1)
            row.append((p4*row[k] - prev_row[k])/(p4 - 1.0))
        result = row[i]
        err = abs(result - prev_row[i - 1])
        if err < tol or err < rtol*abs(result):
            break
        prev_row = row
    return result


def py2_div(a, b):
    """``a / b`` with Python-2 semantics (floor division for two ints)."""
    if isinstance(a, (int, _np.integer)) and isinstance(b, (int, _np.integer)):
        return a//b
    return a/b


def install():
    """Give scipy.integrate a ``romberg`` attribute if it lost it."""
    from scipy import integrate
    if not hasattr(integrate, "romberg") or getattr(
            integrate.romberg, "__module__", "") != __name__:
        integrate.romberg = romberg


install()


def py2_lt(a, b):
    """``a < b`` with Python 2's ordering of None below every number (correlation.py:106)."""
    if a is None:
        return b is not None
    if b is None:
        return False
    return a < b


def py2_gt(a, b):
    """``a > b`` with Python 2's ordering of None below every number (correlation.py:106)."""
    return py2_lt(b, a)


_PY2_CORRELATION_KEY_ORDER = ("D_z", "kernel", "_ln_k_max", "power_spec", "log_theta_min", "log_theta_max", "theta_array",
                              "wtheta_array", "_ln_k_min", "halo")


def py2_corr_eq(a, b):
    """``a == b`` for two Correlation objects as CPython 2.7 evaluates correlation.py:125-133: the __dict__s are
    compared value by value in the hash order of a 64-bit CPython 2.7 dictionary holding Correlation's attribute
    names (the order above, from the string hash and the open-addressing probe sequence), stopping at the first
    unequal value -- so different correlations compare unequal at 'D_z' or 'kernel', before any array is reached."""
    if a is b:
        return True
    if not isinstance(b, a.__class__):
        return False
    da, db = a.__dict__, b.__dict__
    if len(da) != len(db):
        return False
    keys = [k for k in _PY2_CORRELATION_KEY_ORDER if k in da] + [k for k in da if k not in _PY2_CORRELATION_KEY_ORDER]
    for k in keys:
        if k not in db:
            return False
        if not bool(da[k] == db[k]):
            return False
    return True
