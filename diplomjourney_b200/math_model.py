"""Drop-in for the reference's ``math_model.py``: same module-level names, same function
signatures (math_model.py:40-231), FULL control tree evaluated by the CUDA library.

Differences a user can observe, all deliberate:
  * importing the module does not run the experiment (the reference's module-level loop,
    math_model.py:234-258, is ``run_scenario()`` here) and opens no matplotlib figure;
  * ``predictive_control`` finishes: the reference allocates ``np.empty([S**3, 3])`` = 345 TB
    at the default 201 x 121 grid and cannot run it at all.
"""
from .config import phi_0, L, y_t, y_0, x_t, x_0, beta_max, v_max, eps, delta_t, delta_beta, delta_v
from . import _full_impl

# state of the vehicle (math_model.py:11-17)
beta = 0
v = 0
phi = phi_0
x = x_0
y = y_0
coord_actual = [x, y]

prediction_horizon = 3                                   # math_model.py:20

vector_v, vector_beta = _full_impl.default_grids(globals())   # math_model.py:23-30: 201 x 121

result_vector_x = [x]
result_vector_y = [y]
result_vector_phi = [phi]

size_max_1 = vector_beta.size * vector_v.size            # math_model.py:117-119
size_max_2 = pow(size_max_1, 2)
size_max_3 = pow(size_max_1, 3)

t = 0

# defines is_on_target, get_distance_from_line, get_distance_from_target, saturation, v_x, v_y,
# v_phi, control_criterion, integrate_velocity, integrate_angle, coordinate_x, coordinate_y,
# angle_phi, iteration_of_predict, predictive_control, run_scenario, reset_scenario and the
# module state optimal_trajectory / optimal_criterion (math_model.py:131-133)
_full_impl.install(globals(), placeholder_trajectory=[[[0]]])
