"""Controller constants under the names the reference's scripts import (``config.py`` there: a flat
list of assignments, lines 3-28).  The names and values are part of the drop-in contract; here they
are declared by physical group and unit and bound into the module namespace once, so a typo in a
name or unit shows up in one place (and in ``tests/test_host_api.py::test_config_names_and_values``).
"""
import math

_ANGLES_DEG = {            # stored in radians, like the reference
    "eps_beta": 5,         # slack on the steering limit (note: the window filter applies radians() to it again)
    "beta_max": 60,        # steering angle limit
    "delta_beta": 1,       # steering grid step
    "beta_acc_max": 400,   # steering rate limit, per second
}
_VEHICLE_AND_GRID = {
    "L": 0.5,              # wheelbase, m
    "delta_t": 0.05,       # control tick, s
    "v_max": 1,            # m/s
    "v_min": 0.4,          # floor applied while slowing down
    "delta_v": 0.005,      # velocity grid step
    "v_acc_max": 0.5,      # m/s^2
    "eps": 0.001,          # is_on_target tolerance on the SQUARED distance
}
_SCENARIO = {
    "x_0": 0, "y_0": 0, "phi_0": 0,    # initial pose = origin of the tracked line
    "x_t": 1, "y_t": 5,                # operator target
}

globals().update({name: math.radians(deg) for name, deg in _ANGLES_DEG.items()})
globals().update(_VEHICLE_AND_GRID)
globals().update(_SCENARIO)

__all__ = sorted(list(_ANGLES_DEG) + list(_VEHICLE_AND_GRID) + list(_SCENARIO))
