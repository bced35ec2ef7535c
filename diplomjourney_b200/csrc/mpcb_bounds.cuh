// mpcb_bounds.cuh -- the float64 prefix walk and the lower bounds behind the exact branch-and-bound.
//
// Kept in a header of plain host/device functions so that the CPU test-suite can compile exactly this code with g++
// and check the inequality it relies on -- bound <= cost of every leaf below the node -- against the float64 oracle
// for every node of small trees (tests/test_pruning_bound_math.py); the kernels in mpcb_kernels.cu include it as is.
#pragma once
#include <cmath>

#include "mpcb_types.cuh"

namespace mpcb {

using std::fabs; using std::fmin; using std::fmax; using std::sqrt; using std::fma;   // the float overloads too (host build)

// cos / sin of the heading range reachable in i+1 steps, (i+1) dphi_max; cos = -2 marks "the whole circle"
// (a.g.smax / smin / dphimax must be set: their float roundings for the fp32 pre-filter are taken here too)
inline void bounds_set_heading_ranges(LaunchArgs &a, double dphimax) {
    for (int i = 0; i < kMaxH; ++i) {
        const double ang = (i + 1) * dphimax;
        a.cosk[i] = ang < 3.141592653589793 ? std::cos(ang) : -2.0;
        a.sink[i] = ang < 3.141592653589793 ? std::sin(ang) : 0.0;
    }
    a.bc32[0] = (float)a.g.smax; a.bc32[1] = (float)a.g.smin; a.bc32[2] = (float)a.g.dphimax;
    a.bc32[3] = (float)a.cosk[0]; a.bc32[4] = (float)a.sink[0]; a.bc32[5] = (float)a.cosk[1]; a.bc32[6] = (float)a.sink[1];
}

// The scalars the bounds read, in the arithmetic type T they are evaluated in: double for the bounds that decide
// (every cut is a float64 statement), float for the pre-filter of the pruned pass 1 (node_prefilter32 below).
template <typename T>
struct BoundConsts {
    T smax, smin, dphimax, wl, inv_wl, wh;
    T cosk[kMaxH], sink[kMaxH];
};

// (only the heading ranges of the first `steps` steps are filled in: a bound with `steps` steps of slack reads no more)
template <typename T>
MPCB_HD BoundConsts<T> bound_consts(const LaunchArgs &a, const SolveParams &P, int steps = kMaxH) {
    BoundConsts<T> c;
    c.smax = (T)a.g.smax; c.smin = (T)a.g.smin; c.dphimax = (T)a.g.dphimax;
    c.wl = (T)P.wl; c.inv_wl = (T)P.inv_wl; c.wh = (T)P.wh;
    for (int i = 0; i < kMaxH && i < steps; ++i) { c.cosk[i] = (T)a.cosk[i]; c.sink[i] = (T)a.sink[i]; }
    return c;
}

// min over |q| <= Q of q^2 + e q  (the line / heading offset terms of leaf_val are of this form)
template <typename T>
MPCB_HD T quad_min(T e, T Q) {
    const T ae = fabs(e);
    return ae <= T(2) * Q ? T(-0.25) * e * e : Q * (Q - ae);
}

// one step of the prefix walk in the start frame: heading by the angle-addition recurrence
// (t = {cos dphi, sin dphi, s, dphi} of the step's control)
template <typename T>
MPCB_HD void walk_step_t(T tc, T ts, T tstep, T tdphi, T &xi, T &eta, T &psi, T &cp, T &sp) {
    const T cn = cp * tc - sp * ts;
    const T sn = sp * tc + cp * ts;
    cp = cn; sp = sn;
    xi = fma(tstep, cp, xi);
    eta = fma(tstep, sp, eta);
    psi += tdphi;
}
MPCB_HD void walk_step(const double4 t, double &xi, double &eta, double &psi, double &cp, double &sp) {
    walk_step_t<double>(t.x, t.y, t.z, t.w, xi, eta, psi, cp, sp);
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
template <typename T>
MPCB_HD void projection_range(const BoundConsts<T> &k, T cg, T sg, int steps, T &lo, T &hi) {
    if (k.smin < T(0) || !(cg * cg + sg * sg < T(2))) { hi = steps * k.smax; lo = -hi; return; }   // (NaN: isotropic)
    lo = T(0); hi = T(0);
    for (int i = 0; i < steps; ++i) {
        const T ci = k.cosk[i], si = k.sink[i];                          // cos, sin of (i+1) dphi_max; -2, 0: whole circle
        const T cmax = cg >= ci ? T(1) : cg * ci + sg * si;              // cos(max(0, gamma - (i+1) dphi_max))
        const T cmin = -cg >= ci ? T(-1) : cg * ci - sg * si;            // cos(min(pi, gamma + (i+1) dphi_max))
        hi += cmax >= T(0) ? k.smax * cmax : k.smin * cmax;
        lo += cmin <= T(0) ? k.smax * cmin : k.smin * cmin;
    }
}

// min over q in [qlo, qhi] of q^2 + e q
template <typename T>
MPCB_HD T quad_min_range(T e, T qlo, T qhi) {
    const T q = fmin(fmax(T(-0.5) * e, qlo), qhi);
    return q * (q + e);
}

// Lower bound of J_rel over every leaf `steps` control steps below a node, from its quantities in ANY frame in which
// (ch, sh) is its heading: (rx, ry) target relative to the node and D its length, (nx, ny) gradient of the scaled line
// distance (length wl), ep scaled line distance, hp scaled heading error -- the closest approach the steering limits
// allow, the most favourable line offset inside the bracket they allow, the most favourable heading offset.
template <typename T>
MPCB_HD T lower_bound_from_t(const BoundConsts<T> &k, T rx, T ry, T D, T nx, T ny, T ch, T sh, T ep, T hp, T base0,
                             int steps) {
    T lo, reach, qlo, qhi;
    if (D > T(0)) {
        const T inv = T(1) / D;
        projection_range<T>(k, (rx * ch + ry * sh) * inv, fabs(rx * sh - ry * ch) * inv, steps, lo, reach);
    } else {
        reach = steps * k.smax;
    }
    projection_range<T>(k, (nx * ch + ny * sh) * k.inv_wl, fabs(nx * sh - ny * ch) * k.inv_wl, steps, qlo, qhi);
    // a distance cannot become negative: a leaf is no closer to the target than max(0, D - reach)
    return base0 - T(kWd) * fmin(reach, D) + quad_min_range<T>(T(2) * ep, k.wl * qlo, k.wl * qhi) +
           quad_min<T>(T(-2) * hp, k.wh * steps * k.dphimax);
}

MPCB_HD double lower_bound_from(const LaunchArgs &a, const SolveParams &P, double rx, double ry,
                                                   double D, double nx, double ny, double ch, double sh, double ep,
                                                   double hp, double base0, int steps) {
    return lower_bound_from_t<double>(bound_consts<double>(a, P, steps), rx, ry, D, nx, ny, ch, sh, ep, hp, base0, steps);
}

// ... for a node at (xi, eta, psi) with heading (cp, sp) in the start frame, from the solve's start-frame quantities
// (u0, w0 target, d0 its distance, e0 scaled line distance, (nx0, ny0) its gradient, hp0 scaled heading error)
template <typename T>
MPCB_HD T subtree_lower_bound_t(const BoundConsts<T> &k, T u0, T w0, T d0, T e0, T nx0, T ny0, T hp0, T xi, T eta, T psi,
                                T cp, T sp, int steps) {
    const T relx = u0 - xi, rely = w0 - eta;
    const T D = sqrt(relx * relx + rely * rely);
    const T dl = nx0 * xi + ny0 * eta;                    // ep - e0
    const T ep = e0 + dl;
    const T dh = -k.wh * psi;                             // hp - hp0
    const T hp = hp0 + dh;
    const T base0 = T(kWd) * (D - d0) + dl * (ep + e0) + dh * (hp + hp0);
    return lower_bound_from_t<T>(k, relx, rely, D, nx0, ny0, cp, sp, ep, hp, base0, steps);
}

MPCB_HD double subtree_lower_bound(const LaunchArgs &a, const SolveParams &P, double xi, double eta,
                                                      double psi, double cp, double sp, int steps) {
    return subtree_lower_bound_t<double>(bound_consts<double>(a, P, steps), P.u0, P.w0, P.d0, P.e0, P.nx0, P.ny0, P.hp0, xi, eta,
                                         psi, cp, sp, steps);
}

// fp32 PRE-FILTER of the pruned pass 1: the same bound over the children of a depth-(H-1) node evaluated in float from
// a float walk.  It decides nothing on its own: a node is dropped early only if  lb32 > bound + 8 tol1 , and
//   |lb32 - lb64| <= 2^-24 [kWd (8 Dmax + 6 smax) + 16 Q(2E + Q) + 10 G(2H + G)] <= tol1
// (the roundings of u0, w0, the walk, D, the three differences of squares and the two projected ranges, with
// Dmax = d0 + H smax, E, H the anchor bounds and Q, G the one-step offsets of prep_kernel's error model; tol1 =
// 2^-22 [4 kWd Dmax + 2 kWd smax + 5 (E+Q)^2 + 4 (H+G)^2]), so every node it drops the float64 test drops too and the set of
// nodes that are scored -- and every number downstream -- is unchanged; it only spares the float64 set-up of the 99.9 %
// of the nodes that are nowhere near the bound.
struct Prefilter32 {
    BoundConsts<float> k;
    float u0, w0, d0, e0, nx0, ny0, hp0;
};

// once per solve (prep_kernel): the start-frame quantities rounded to float
MPCB_HD void prefilter_solve_consts(SolveParams &P) {
    P.pf[0] = (float)P.u0; P.pf[1] = (float)P.w0; P.pf[2] = (float)P.d0; P.pf[3] = (float)P.e0;
    P.pf[4] = (float)P.nx0; P.pf[5] = (float)P.ny0; P.pf[6] = (float)P.hp0;
    P.pf[7] = (float)P.wl; P.pf[8] = (float)P.inv_wl; P.pf[9] = (float)P.wh;
    P.pf[10] = 0.f; P.pf[11] = 0.f;
}

// per node: nothing but loads (the pre-filter runs on every node of a listed tile; conversions per node were a tenth of it)
MPCB_HD Prefilter32 prefilter32(const LaunchArgs &a, const SolveParams &P) {
    Prefilter32 f;
    f.k.smax = a.bc32[0]; f.k.smin = a.bc32[1]; f.k.dphimax = a.bc32[2];
    f.k.cosk[0] = a.bc32[3]; f.k.sink[0] = a.bc32[4]; f.k.cosk[1] = a.bc32[5]; f.k.sink[1] = a.bc32[6];
    f.u0 = P.pf[0]; f.w0 = P.pf[1]; f.d0 = P.pf[2]; f.e0 = P.pf[3];
    f.nx0 = P.pf[4]; f.ny0 = P.pf[5]; f.hp0 = P.pf[6];
    f.k.wl = P.pf[7]; f.k.inv_wl = P.pf[8]; f.k.wh = P.pf[9];
    return f;
}

// steps = 1: the children of a depth-(H-1) node; steps = 2: the two levels below a depth-(H-2) node (the tile test).
// With two steps the reach, the line bracket Q and the heading bracket G double, so the error model above reads
//   2^-24 [kWd (8 Dmax + 12 smax) + 64 Q (E + Q) + 40 G (H + G)]
// which is below 8 tol1 = 2^-24 [128 kWd Dmax + 64 kWd smax + 160 (E+Q)^2 + 128 (H+G)^2] term by term: the same margin
// of 8 tol1 serves both (tests/test_pruning_bound_math.py measures either against the float64 bound).
MPCB_HD float node_prefilter32(const Prefilter32 &f, float xi, float eta, float psi, float cp, float sp, int steps = 1) {
    return subtree_lower_bound_t<float>(f.k, f.u0, f.w0, f.d0, f.e0, f.nx0, f.ny0, f.hp0, xi, eta, psi, cp, sp, steps);
}

}  // namespace mpcb
