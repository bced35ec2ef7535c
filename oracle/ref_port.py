"""TEST/BENCH INFRASTRUCTURE ONLY -- scalar Python port of the reference's CPU path, kept in the
reference's own implementation style so that its speed is representative of what a user of
the reference runs today:

* every node expansion is three ``scipy.integrate.quad`` calls on Python integrands
  (math_model.py:69-78,90-114), exactly like ``iteration_of_predict``;
* the tree is walked layer by layer with Python loops (math_model.py:159-200) and every
  leaf is scored with the scalar ``control_criterion`` (math_model.py:82-86) and compared
  with a strict ``<`` against the running optimum.

Unlike the reference it can be restricted to a sub-range of the tree (first control i0 and a
slice of second controls), because the reference's dense ``np.empty([S**3, 3])`` cannot be
allocated for the benchmark grids; the per-leaf work is unchanged.  ``bench.py --impl
reference`` and the ``cpu_baseline`` leg time this file (the unmodified reference itself is
only available in the build container, via oracle/ref_exec.py).
"""
from __future__ import annotations

import math
import time

import numpy as np
import scipy.integrate as sp

from .closed_form import CONFIG


class RefPort:
    def __init__(self, vector_v, vector_beta, target, origin, cost="mm", cfg=CONFIG):
        self.vector_v = list(vector_v)
        self.vector_beta = list(vector_beta)
        self.x_t, self.y_t = target
        self.x_0, self.y_0 = origin
        self.L, self.delta_t = cfg["L"], cfg["delta_t"]
        self.cost = cost
        self.t = 0.0

    # --- math_model.py:69-78
    def v_x(self, _time, v, phi):
        return v * np.cos(phi)

    def v_y(self, _time, v, phi):
        return v * np.sin(phi)

    def v_phi(self, _time, v, beta):
        return (v / self.L) * math.tan(beta)

    # --- math_model.py:110-114
    def iteration_of_predict(self, state, v, beta):
        t0, t1 = self.t, self.t + self.delta_t
        dphi = sp.quad(self.v_phi, t0, t1, args=(v, beta))[0]
        dx = sp.quad(self.v_x, t0, t1, args=(v, state[2] + dphi))[0]
        dy = sp.quad(self.v_y, t0, t1, args=(v, state[2] + dphi))[0]
        return [state[0] + dx, state[1] + dy, state[2] + dphi]

    # --- math_model.py:48-58,82-86 / math_model_tree.py:56-66,82-87
    def distance_from_line(self, x, y):
        if x == self.x_0 and y == self.y_0:
            return 1000
        return abs((self.y_t - self.y_0) * x - (self.x_t - self.x_0) * y + self.x_t * self.y_0 - self.y_t * self.x_0) \
            / math.sqrt((self.y_t - self.y_0) ** 2 + (self.x_t - self.x_0) ** 2)

    def control_criterion(self, s):
        d = math.sqrt((self.x_t - s[0]) ** 2 + (self.y_t - s[1]) ** 2)
        dl = self.distance_from_line(s[0], s[1])
        if self.cost == "mm":
            ang = np.arctan(self.x_t / self.y_t) - s[2]
            return 10000 * d + 10 * ang ** 2 + 100 * dl ** 2
        return 10000 * d + 10000 * dl ** 2

    # --- math_model.py:159-200 restricted to first control i0 and second controls [i1_lo, i1_hi)
    def full_h3_subtree(self, state, i0, i1_lo, i1_hi, threshold=math.inf):
        ctr = [(v, b) for v in self.vector_v for b in self.vector_beta]
        S = len(ctr)
        best, best_j, leaves = threshold, -1, 0
        n0 = self.iteration_of_predict(state, *ctr[i0])
        for i1 in range(i1_lo, i1_hi):
            n1 = self.iteration_of_predict(n0, *ctr[i1])
            row = np.array(n1)                      # the reference stores np.array rows (math_model.py:183)
            for i2 in range(S):
                n2 = self.iteration_of_predict(row, *ctr[i2])
                leaf = np.array(n2)                 # math_model.py:194
                if self.control_criterion(n2) < best:
                    best = self.control_criterion(leaf)   # evaluated twice on improvement (math_model.py:195,198)
                    best_j = (i0 * S + i1) * S + i2
                leaves += 1
        return best, best_j, leaves

    # --- math_model_tree.py:308-361
    def held(self, state, H=3):
        ctr = [(v, b) for v in self.vector_v for b in self.vector_beta]
        best, best_k = math.inf, -1
        nodes = [list(state)] * len(ctr)
        for _ in range(H):
            nodes = [self.iteration_of_predict(nodes[k], *ctr[k]) + list(ctr[k]) for k in range(len(ctr))]
        for k, n in enumerate(nodes):
            J = self.control_criterion(n)
            if J < best:
                best, best_k = J, k
        return best, best_k, len(ctr)


def _worker(args):
    """One process = one core: time a slice of one robot's FULL tree."""
    (V, B, scen, cost, i0, i1_lo, i1_hi) = args
    rp = RefPort(V, B, (scen[3], scen[4]), (scen[0], scen[1]), cost)
    t = time.perf_counter()
    _, _, leaves = rp.full_h3_subtree(list(scen[:3]), i0, i1_lo, i1_hi)
    return leaves, time.perf_counter() - t


def timed_sample_full(V, B, scenarios, cost, n_i1, procs):
    """Each of ``procs`` worker processes walks n_i1 second-level subtrees (n_i1 * S leaves) of a
    different robot.  Returns (total leaves, wall seconds, per-process seconds)."""
    import multiprocessing as mp

    S = len(V) * len(B)
    jobs = [(list(V), list(B), list(map(float, scenarios[p % len(scenarios)])), cost, (7 * p) % S, 0, min(n_i1, S))
            for p in range(procs)]
    t = time.perf_counter()
    if procs == 1:
        res = [_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t
    return sum(r[0] for r in res), wall, [r[1] for r in res]
