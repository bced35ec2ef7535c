// TEST HARNESS (not part of the product): compiles SHIPPED pieces of the CUDA library for the host with g++ --
// csrc/mpcb_exact.cuh (the float64 leaf evaluation whose results the library returns) and the invariant-divisor
// index decoding of csrc/mpcb_types.cuh, and the operator events of csrc/mpcb_events.cuh -- so that tests/test_shipped_code_on_host.py can check them against the
// reference's golden outputs and the oracle without a GPU.
#include <cmath>
#include <vector>

#include "mpcb_events.cuh"
#include "mpcb_exact.cuh"

// one operator event (kind 1 new_target(a, b), 2 turn_left(a), 3 turn_right(a)) on a pose: line[4] = x_t, y_t, x_0, y_0
extern "C" int mpcb_test_apply_event(int kind, double a, double b, double radius, double x, double y, double phi,
                                     int slow_before, double *line) {
    mpcb_loop_event e = {0, kind, a, b};
    int slow = slow_before;
    mpcb::apply_event(e, radius, x, y, phi, line, slow);
    return slow;
}

extern "C" void mpcb_test_fastdiv64(unsigned long long d, long long count, const unsigned long long *n,
                                    unsigned long long *q) {
    mpcb::FastDiv64 f;
    mpcb::fastdiv_init(f, d);
    for (long long i = 0; i < count; ++i) q[i] = f.div(n[i]);
}

extern "C" void mpcb_test_fastdiv32(unsigned d, long long count, const unsigned *n, unsigned *q) {
    mpcb::FastDiv32 f;
    mpcb::fastdiv32_init(f, d);
    for (long long i = 0; i < count; ++i) q[i] = f.div(n[i]);
}

// cost of leaf j of one solve (mode 0 FULL / 1 HELD), poses after each step and the first control, computed by
// mpcb::exact_cost from tables built as mpcb_set_grid builds them and solve parameters as prep_kernel builds them
extern "C" double mpcb_test_exact_cost(const double *v, int nv, const double *beta, int nb, double L, double delta_t,
                                       int mode, int cost_kind, int H, const double *state, const double *target,
                                       const double *origin, long long j, double *traj, int *first_c, int slow,
                                       double v_min) {
    const int S = nv * nb;
    std::vector<double4> tab(S), tab_slow(S);
    std::vector<double> vt(S), vt_slow(S);
    double vmin_grid = v[0];
    for (int i = 1; i < nv; ++i) vmin_grid = std::fmin(vmin_grid, v[i]);
    const double v_slow = vmin_grid > v_min ? vmin_grid : v_min;             // math_model_tree.py:312-316
    for (int iv = 0; iv < nv; ++iv)
        for (int ib = 0; ib < nb; ++ib) {
            const int c = iv * nb + ib;
            for (int variant = 0; variant < 2; ++variant) {
                const double vc = variant ? v_slow : v[iv];
                const double dphi = (vc / L) * std::tan(beta[ib]) * delta_t;   // math_model.py:77-78 x delta_t
                (variant ? tab_slow : tab)[c] = make_double4(std::cos(dphi), std::sin(dphi), vc * delta_t, dphi);
                (variant ? vt_slow : vt)[c] = vc;
            }
        }
    mpcb::LaunchArgs a = {};
    a.g.tab64 = tab.data(); a.g.vtab = vt.data(); a.g.tab64_slow = tab_slow.data(); a.g.vtab_slow = vt_slow.data();
    a.g.S = S; a.g.nb = nb; a.g.dt = delta_t;
    a.mode = mode; a.H = H; a.cost_kind = cost_kind;
    unsigned long long pw = 1;
    for (int k = H - 1; k >= 0; --k) {   // fd[k].d = S^(H-1-k), as fill_args (mpcb_api.cu)
        mpcb::fastdiv_init(a.fd[k], pw);
        if (mode == 0) pw *= (unsigned long long)S;
    }
    mpcb::SolveParams P = {};
    P.xs = state[0]; P.ys = state[1]; P.phi0 = state[2];
    P.xt = target[0]; P.yt = target[1];
    P.ox = origin[0]; P.oy = origin[1];
    P.theta = std::atan(P.xt / P.yt);
    P.lineA = P.yt - P.oy; P.lineB = P.xt - P.ox;
    P.lineC = P.xt * P.oy - P.yt * P.ox;
    P.line_norm = std::sqrt(P.lineA * P.lineA + P.lineB * P.lineB);
    P.flags = slow ? mpcb::kFlagSlow : 0;
    return mpcb::exact_cost(a, P, j, traj, first_c);
}
