"""Drop-in for the reference's ``run_math_model.py``: the FULL-tree controller of
``math_model.py`` driven over random scenarios (run_math_model.py:231-280).

Module-level names and function signatures follow run_math_model.py:13-228; the 1000-scenario
loop that the reference runs at import time is ``run(n_scenarios, seed)`` here, and no figure
is saved (plotting is out of scope).
"""
import math

import numpy as np

from .config import phi_0, L, y_t, y_0, x_t, x_0, beta_max, v_max, eps, delta_t, delta_beta, delta_v
from . import _full_impl

beta = 0
v = 0
phi = phi_0
x = x_0
y = y_0
coord_actual = [x, y]

prediction_horizon = 3

vector_v, vector_beta = _full_impl.default_grids(globals())

result_vector_x = [x]
result_vector_y = [y]
result_vector_phi = [phi]

size_max_1 = vector_beta.size * vector_v.size
size_max_2 = pow(size_max_1, 2)
size_max_3 = pow(size_max_1, 3)

t = 0

_full_impl.install(globals(), placeholder_trajectory=[0])   # run_math_model.py:128-130


def draw_scenario(rng=np.random):
    """run_math_model.py:235-239."""
    sx = rng.uniform(-10, 10)
    sy = rng.uniform(-10, 10)
    sphi = rng.uniform(-math.pi, math.pi)
    return sx, sy, sphi, rng.uniform(sx - 10, sx + 10), rng.uniform(sy - 10, sy + 10)


def run(n_scenarios=1000, seed=None, max_ticks=None, verbose=False):
    """The experiment of run_math_model.py:231-280; returns [(scenario, path)]."""
    rng = np.random if seed is None else np.random.RandomState(seed)
    out = []
    for _ in range(n_scenarios):
        sc = draw_scenario(rng)
        reset_scenario(*sc)                                  # noqa: F821 (installed above)
        if verbose:
            print([sc[0], sc[1], sc[2]], [sc[3], sc[4]])
        out.append((sc, run_scenario(max_ticks=max_ticks, verbose=verbose)))   # noqa: F821
    return out
