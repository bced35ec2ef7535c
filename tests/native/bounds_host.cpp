// TEST HARNESS (not part of the product): compiles the SHIPPED bound code of the CUDA library,
// diplomjourney_b200/csrc/mpcb_bounds.cuh, for the host with g++ so that tests/test_pruning_bound_math.py can
// evaluate exactly what the kernels evaluate.
#include "mpcb_bounds.cuh"

extern "C" void mpcb_test_subtree_lower_bounds(double smax, double smin, double dphimax,
                                               const double *solve,   // u0, w0, d0, e0, nx0, ny0, hp0, wl, wh
                                               long long n, const double *xi, const double *eta, const double *psi,
                                               int steps, double *out) {
    mpcb::LaunchArgs a = {};
    a.g.smax = smax; a.g.smin = smin; a.g.dphimax = dphimax;
    mpcb::bounds_set_heading_ranges(a, dphimax);
    mpcb::SolveParams P = {};
    P.u0 = solve[0]; P.w0 = solve[1]; P.d0 = solve[2]; P.e0 = solve[3]; P.nx0 = solve[4]; P.ny0 = solve[5];
    P.hp0 = solve[6]; P.wl = solve[7]; P.wh = solve[8]; P.inv_wl = 1.0 / P.wl;
    for (long long i = 0; i < n; ++i)
        out[i] = mpcb::subtree_lower_bound(a, P, xi[i], eta[i], psi[i], std::cos(psi[i]), std::sin(psi[i]), steps);
}

// the fp32 pre-filter of the pruned pass 1 on the same nodes (poses given in float64, rounded to float as the float walk
// would hold them)
extern "C" void mpcb_test_prefilter32(double smax, double smin, double dphimax, const double *solve, long long n,
                                      const double *xi, const double *eta, const double *psi, float *out, int steps) {
    mpcb::LaunchArgs a = {};
    a.g.smax = smax; a.g.smin = smin; a.g.dphimax = dphimax;
    mpcb::bounds_set_heading_ranges(a, dphimax);
    mpcb::SolveParams P = {};
    P.u0 = solve[0]; P.w0 = solve[1]; P.d0 = solve[2]; P.e0 = solve[3]; P.nx0 = solve[4]; P.ny0 = solve[5];
    P.hp0 = solve[6]; P.wl = solve[7]; P.wh = solve[8]; P.inv_wl = 1.0 / P.wl;
    mpcb::prefilter_solve_consts(P);                      // what prep_kernel does once per solve
    const mpcb::Prefilter32 f = mpcb::prefilter32(a, P);
    for (long long i = 0; i < n; ++i)
        out[i] = mpcb::node_prefilter32(f, (float)xi[i], (float)eta[i], (float)psi[i], std::cos((float)psi[i]),
                                        std::sin((float)psi[i]), steps);
}
