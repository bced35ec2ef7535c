// mpcb_bounds.cuh -- the float64 prefix walk and the lower bounds behind the exact branch-and-bound.
//
// Kept in a header of plain host/device functions so that the CPU test-suite can compile exactly this code with g++
// and check the inequality it relies on -- bound <= cost of every leaf below the node -- against the float64 oracle
// for every node of small trees (tests/test_pruning_bound_math.py); the kernels in mpcb_kernels.cu include it as is.
#pragma once
#include <cmath>

#include "mpcb_types.cuh"

namespace mpcb {

// cos / sin of the heading range reachable in i+1 steps, (i+1) dphi_max; cos = -2 marks "the whole circle"
inline void bounds_set_heading_ranges(LaunchArgs &a, double dphimax) {
    for (int i = 0; i < kMaxH; ++i) {
        const double ang = (i + 1) * dphimax;
        a.cosk[i] = ang < 3.141592653589793 ? std::cos(ang) : -2.0;
        a.sink[i] = ang < 3.141592653589793 ? std::sin(ang) : 0.0;
    }
}

// min over |q| <= Q of q^2 + e q  (the line / heading offset terms of leaf_val are of this form)
MPCB_HD double quad_min(double e, double Q) {
    const double ae = fabs(e);
    return ae <= 2.0 * Q ? -0.25 * e * e : Q * (Q - ae);
}

// one step of the float64 prefix walk in the start frame: heading by the angle-addition recurrence
MPCB_HD void walk_step(const double4 t, double &xi, double &eta, double &psi, double &cp, double &sp) {
    double cn = cp * t.x - sp * t.y;
    double sn = sp * t.x + cp * t.y;
    cp = cn; sp = sn;
    xi = fma(t.z, cp, xi);
    eta = fma(t.z, sp, eta);
    psi += t.w;
}

// Range [lo, hi] of the displacement of `steps` further control steps along a FIXED direction that makes the angle
// gamma, given by (cg, sg) = (cos gamma, |sin gamma|), with the node's present heading.  Step i moves by s_i in
// [s_min, s_max] along a heading that has turned by at most i dphi_max, so its projection is s_i cos(angle_i) with
// max(0, gamma - i dphi_max) <= angle_i <= min(pi, gamma + i dphi_max).
//  * along the bearing to the target, hi is the closest approach: the distance of a leaf is at least its projection
//    on that bearing, d >= D - hi.  A robot that faces away from its target, or cannot stop, is thereby known to
//    move AWAY from it;
//  * along the gradient of the signed line distance, wl [lo, hi] brackets the line offset q of every leaf.
// Grids with negative speeds fall back to the isotropic range +-steps max|s|.
MPCB_HD void projection_range(const LaunchArgs &a, double cg, double sg, int steps, double &lo,
                                                 double &hi) {
    if (a.g.smin < 0.0 || !(cg * cg + sg * sg < 2.0)) { hi = steps * a.g.smax; lo = -hi; return; }   // (NaN: isotropic)
    lo = 0.0; hi = 0.0;
    for (int i = 0; i < steps; ++i) {
        const double ci = a.cosk[i], si = a.sink[i];                         // cos, sin of (i+1) dphi_max; -2, 0: whole circle
        const double cmax = cg >= ci ? 1.0 : cg * ci + sg * si;              // cos(max(0, gamma - (i+1) dphi_max))
        const double cmin = -cg >= ci ? -1.0 : cg * ci - sg * si;            // cos(min(pi, gamma + (i+1) dphi_max))
        hi += cmax >= 0.0 ? a.g.smax * cmax : a.g.smin * cmax;
        lo += cmin <= 0.0 ? a.g.smax * cmin : a.g.smin * cmin;
    }
}

// min over q in [qlo, qhi] of q^2 + e q
MPCB_HD double quad_min_range(double e, double qlo, double qhi) {
    const double q = fmin(fmax(-0.5 * e, qlo), qhi);
    return q * (q + e);
}

// Lower bound of J_rel over every leaf `steps` control steps below a node, from its quantities in ANY frame in which
// (ch, sh) is its heading: (rx, ry) target relative to the node and D its length, (nx, ny) gradient of the scaled line
// distance (length wl), ep scaled line distance, hp scaled heading error -- the closest approach the steering limits
// allow, the most favourable line offset inside the bracket they allow, the most favourable heading offset.
MPCB_HD double lower_bound_from(const LaunchArgs &a, const SolveParams &P, double rx, double ry,
                                                   double D, double nx, double ny, double ch, double sh, double ep,
                                                   double hp, double base0, int steps) {
    double lo, reach, qlo, qhi;
    if (D > 0.0) {
        const double inv = 1.0 / D;
        projection_range(a, (rx * ch + ry * sh) * inv, fabs(rx * sh - ry * ch) * inv, steps, lo, reach);
    } else {
        reach = steps * a.g.smax;
    }
    const double invl = 1.0 / P.wl;
    projection_range(a, (nx * ch + ny * sh) * invl, fabs(nx * sh - ny * ch) * invl, steps, qlo, qhi);
    // a distance cannot become negative: a leaf is no closer to the target than max(0, D - reach)
    return base0 - kWd * fmin(reach, D) + quad_min_range(2.0 * ep, P.wl * qlo, P.wl * qhi) +
           quad_min(-2.0 * hp, P.wh * steps * a.g.dphimax);
}

// ... for a node at (xi, eta, psi) with heading (cp, sp) in the start frame
MPCB_HD double subtree_lower_bound(const LaunchArgs &a, const SolveParams &P, double xi, double eta,
                                                      double psi, double cp, double sp, int steps) {
    const double relx = P.u0 - xi, rely = P.w0 - eta;
    const double D = sqrt(relx * relx + rely * rely);
    const double ep = P.e0 + P.nx0 * xi + P.ny0 * eta;
    const double hp = P.hp0 - P.wh * psi;
    const double base0 = kWd * (D - P.d0) + (ep - P.e0) * (ep + P.e0) + (hp - P.hp0) * (hp + P.hp0);
    return lower_bound_from(a, P, relx, rely, D, P.nx0, P.ny0, cp, sp, ep, hp, base0, steps);
}

}  // namespace mpcb
