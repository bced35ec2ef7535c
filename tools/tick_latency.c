/* Latency of one online tick through the C ABI alone (no Python): mpcb_held_tick_host in a loop.
 *   gcc -O2 -Iinclude tools/tick_latency.c -o build/tick_latency -Ldiplomjourney_b200/lib -lmpcb200 -Wl,-rpath,$PWD/diplomjourney_b200/lib -lm
 *   ./build/tick_latency            (on a GPU box) */
#include <math.h>
#include <stdio.h>
#include <time.h>

#include "mpcb200.h"

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(void) {
    mpcb_handle *h;
    if (mpcb_create(0, &h) != 0) { fprintf(stderr, "no CUDA device\n"); return 1; }
    double v[11], b[41];
    for (int i = 0; i < 11; ++i) v[i] = 0.5 + 0.005 * (i - 5);                 /* vector_of_velocities(0.5) */
    for (int i = 0; i < 41; ++i) b[i] = (M_PI / 180.0) * (i - 20);             /* vector_of_beta_angles(0) */
    double state[3] = {0.3, -0.2, 0.1}, target[2] = {2, 3}, origin[2] = {0, 0};
    double cost, traj[9], ctl[2];
    int64_t idx;
    for (int zc = 1; zc >= 0; --zc) {
        mpcb_set_option(h, "zero_copy", zc);
        const int reps = 2000;
        double t0 = 0;
        for (int i = -100; i < reps; ++i) {
            if (i == 0) t0 = now();
            state[0] += 1e-6;
            if (mpcb_held_tick_host(h, v, 11, b, 41, 0.5, 0.05, 0.4, MPCB_COST_TREE, 3, state, target, origin, INFINITY, 0,
                                    &cost, &idx, traj, ctl) != 0) { fprintf(stderr, "%s\n", mpcb_last_error(h)); return 1; }
        }
        printf("mpcb_held_tick_host (S=451, H=3, %s): %.2f us per tick, leaf %lld cost %.6f\n",
               zc ? "mapped pinned memory, no copies" : "one staged copy each way", 1e6 * (now() - t0) / reps, (long long)idx, cost);
    }
    mpcb_destroy(h);
    return 0;
}
