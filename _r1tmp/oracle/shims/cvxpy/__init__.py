"""Stand-in for the six `cvxpy` names the reference's hot path uses.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  `cvxpy` is not installable in
this sandbox, so the reference package (`/root/reference/gym_anm`) is imported
UNMODIFIED with this module first on `sys.path`.  The only problems the
reference builds on the hot path are

    x = cp.Variable(2)
    cp.Problem(cp.Minimize(cp.sum_squares(x - point)), [G @ x <= h]).solve()

(reference gym_anm/simulator/components/devices.py:299-304 and :517-522), i.e. the
Euclidean projection of a 2-D point onto the convex polygon {x : G x <= h}.
This module solves exactly that problem, exactly (the unique minimiser of the
strictly convex QP), instead of OSQP's ~1e-5 first-order answer.

The agent-only names (`Parameter`, `abs`, `maximum`, ...) exist so that
`import gym_anm` succeeds; constructing an MPC agent raises.
"""
import numpy as np

from ._projection import project_onto_polygon

__version__ = "0.0-standin"


class _Affine:
    """`point`-shifted identity expression: value(x) = x - shift."""

    def __init__(self, var, shift):
        self.var, self.shift = var, np.asarray(shift, dtype=np.float64)


class _LinearMap:
    """G @ x."""

    __array_ufunc__ = None

    def __init__(self, G, var):
        self.G, self.var = np.asarray(G, dtype=np.float64), var

    def __le__(self, h):
        return _HalfPlanes(self.G, np.asarray(h, dtype=np.float64), self.var)


class _HalfPlanes:
    def __init__(self, G, h, var):
        self.G, self.h, self.var = G, h, var


class _SumSquares:
    def __init__(self, expr):
        if not isinstance(expr, _Affine):
            raise NotImplementedError("stand-in cvxpy: sum_squares(x - point) only")
        self.expr = expr


class Variable:
    __array_ufunc__ = None  # make `ndarray @ Variable` defer to __rmatmul__

    def __init__(self, shape=None, **kwargs):
        self.shape = shape
        self.value = None

    def __sub__(self, other):
        return _Affine(self, other)

    def __rmatmul__(self, G):
        return _LinearMap(G, self)


class Minimize:
    def __init__(self, objective):
        self.objective = objective


class Problem:
    def __init__(self, objective, constraints=()):
        self.objective = objective
        self.constraints = list(constraints)
        self.status = None

    def solve(self, *args, **kwargs):
        obj = self.objective.objective
        if not isinstance(obj, _SumSquares) or len(self.constraints) != 1:
            raise NotImplementedError("stand-in cvxpy only solves 2-D polygon projections")
        con = self.constraints[0]
        point = obj.expr.shift
        x = project_onto_polygon(con.G, con.h, point)
        obj.expr.var.value = x
        self.status = "optimal"
        return float(np.sum((x - point) ** 2))


def sum_squares(expr):
    return _SumSquares(expr)


def _agent_only(*args, **kwargs):
    raise NotImplementedError("stand-in cvxpy: MPC-agent LP features are not provided")


Parameter = abs = maximum = _agent_only
