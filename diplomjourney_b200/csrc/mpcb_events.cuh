// mpcb_events.cuh -- operator events of the online controller, as the device-resident closed loop (mpcb_loop.cu)
// applies them between two ticks.  Plain host/device code: the CPU test-suite compiles exactly this with g++ and
// compares it with the reference's own functions (tests/test_shipped_code_on_host.py, tests/golden/operator_events.json).
#pragma once
#include <cmath>

#include "mpcb_types.cuh"

namespace mpcb {

// ---- one operator event of the online controller (math_model_tree.py:118-129 new_target, :142-215 turn_left /
// turn_right, :219-226 slow_down), applied by the device-resident closed loop between two ticks:
// line[0..3] = x_t, y_t, x_0, y_0.  The turn target is the reference's quadrant formula
//   x (+/-) distance * f(tau) (+/-) radius_u_turn * g(tau),   tau = phi - {pi/2, pi, 3pi/2, 0}
// evaluated left to right like the reference's expression; the tracked line restarts at the robot's pose;
// new_target ends in slow_down(30 deg) = 10 slowed ticks, a turn then calls slow_down(90 deg) = 20.
MPCB_HD void apply_event(const mpcb_loop_event &e, double radius, double x, double y, double phi, double *line,
                         int &slow_steps) {
    const double pi = 3.141592653589793;
    double tx, ty;
    if (e.kind == MPCB_EVENT_NEW_TARGET) {
        tx = e.a; ty = e.b;
        slow_steps = 10;
    } else {
        const double d = e.a;
        const double sg = e.kind == MPCB_EVENT_TURN_LEFT ? -1.0 : 1.0;     // right turn = left turn with the distance terms negated
        int q;
        if (pi / 2 <= phi && phi <= 3 * pi / 2) q = phi <= pi ? 0 : 1;
        else q = phi <= 2 * pi ? 2 : 3;
        const double tau = q == 0 ? phi - pi / 2 : q == 1 ? phi - pi : q == 2 ? phi - 3 * pi / 2 : phi;
        const double sn = sin(tau), cs = cos(tau);
        const double dc = dmul(d, cs), ds = dmul(d, sn), rc = dmul(radius, cs), rs = dmul(radius, sn);
        // left turn:  q0  x - d cos - R sin, y - d sin + R cos     q1  x + d sin - R cos, y - d cos - R sin
        //             q2  x + d cos + R sin, y + d sin - R cos     q3  x - d sin + R cos, y + d cos + R sin
        switch (q) {
        case 0:  tx = dadd(dadd(x, sg * dc), -rs); ty = dadd(dadd(y, sg * ds), rc); break;
        case 1:  tx = dadd(dadd(x, -sg * ds), -rc); ty = dadd(dadd(y, sg * dc), -rs); break;
        case 2:  tx = dadd(dadd(x, -sg * dc), rs); ty = dadd(dadd(y, -sg * ds), -rc); break;
        default: tx = dadd(dadd(x, sg * ds), rc); ty = dadd(dadd(y, -sg * dc), rs); break;
        }
        slow_steps = 20;
    }
    line[0] = tx; line[1] = ty; line[2] = x; line[3] = y;
}

}  // namespace mpcb
