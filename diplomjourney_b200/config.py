"""Constants of the controller -- same names and values as the reference's config.py:3-28
(the names are part of the drop-in contract: the scripts import them by name)."""
import math

eps = 0.001
eps_beta = math.radians(5)

# vehicle
L = 0.5

# discretisation / limits
delta_t = 0.05

beta_max = math.radians(60)
delta_beta = math.radians(1)
beta_acc_max = math.radians(400)

v_max = 1
v_min = 0.4
delta_v = 0.005
v_acc_max = 0.5

# initial pose
x_0 = 0
y_0 = 0
phi_0 = 0

# operator target
x_t = 1
y_t = 5
