/* Plain-C consumer of libsdde_b200.so: the drop-in boundary needs nothing but <stdint.h> types and pointers.
 * Builds the C1 and C2 schedules through the ABI and prints the facts SURVEY.md K4 pins (1001 / 2001 steps, clipped
 * steps at the lag multiples); with a GPU it also runs one tiny solve.  Compiled and run by tests/test_abi_host.py. */
#include <stdio.h>
#include <string.h>

#include "sdde_b200.h"

static void fill(sdde_desc *d, int model, int n_param, const double *p, const double *u0, double tf, double dt, double lag,
                 const double *sv, int n_sv) {
    memset(d, 0, sizeof(*d));
    d->struct_size = (int32_t)sizeof(*d);
    d->model_id = model; d->alg = SDDE_ALG_EM_SPLIT;
    d->n_state = 1; d->n_noise = 1; d->n_lag = 1; d->n_param = n_param;
    d->constant_lags_declared = 1; d->save_start = 1; d->save_end = 1;
    d->t0 = 0.0; d->tf = tf; d->dt = dt; d->lags[0] = lag;
    d->hist_c0[0] = u0[0];
    d->params = p; d->u0 = u0; d->saveat = sv; d->n_saveat = n_sv;
    d->want_paths = 1;
}

int main(void) {
    double p1[3] = {-1.0, 0.5, 0.2}, u1[1] = {1.0}, sv1[10];
    double p2[4] = {0.2, 0.1, 10.0, 0.1}, u2[1] = {1.2}, sv2[100];
    sdde_desc d;
    sdde_schedule s;
    int i, rc;
    for (i = 0; i < 10; ++i) sv1[i] = 1.0 * (i + 1);
    for (i = 0; i < 100; ++i) sv2[i] = 1.0 * (i + 1);
    printf("version %d\n", sdde_b200_version());
    fill(&d, SDDE_MODEL_LINEAR1, 3, p1, u1, 10.0, 0.01, 1.0, sv1, 10);
    rc = sdde_b200_schedule(&d, &s);
    if (rc) { printf("schedule failed: %s\n", sdde_b200_last_error()); return 1; }
    printf("C1 steps %lld saves %lld ring %d\n", (long long)s.n_steps, (long long)s.n_save, s.ring_slots);
    sdde_b200_schedule_free(&s);
    fill(&d, SDDE_MODEL_MACKEY_GLASS, 4, p2, u2, 100.0, 0.05, 17.0, sv2, 100);
    rc = sdde_b200_schedule(&d, &s);
    if (rc) { printf("schedule failed: %s\n", sdde_b200_last_error()); return 1; }
    printf("C2 steps %lld saves %lld ring %d tstops", (long long)s.n_steps, (long long)s.n_save, s.ring_slots);
    for (i = 0; i < s.n_tstops_hit; ++i) printf(" %.17g", s.tstops_hit[i]);
    printf("\n");
    sdde_b200_schedule_free(&s);
    d.struct_size = 8;
    rc = sdde_b200_schedule(&d, &s);
    printf("bad struct_size -> %d (%s)\n", rc, rc ? "refused" : "accepted");
    if (sdde_b200_device_count() > 0) {
        double out[101 * 8];
        uint8_t st[8];
        sdde_ctx *c = sdde_b200_create(0);
        if (!c) { printf("create failed: %s\n", sdde_b200_last_error()); return 1; }
        fill(&d, SDDE_MODEL_MACKEY_GLASS, 4, p2, u2, 100.0, 0.05, 17.0, sv2, 100);
        d.n_traj = 8; d.seed = 42;
        rc = sdde_b200_solve(c, &d, out, NULL, st, NULL);
        printf("solve -> %d, u(0) = %g, u(100) of path 0 = %g\n", rc, out[0], out[100 * 8]);
        sdde_b200_destroy(c);
        if (rc) return 1;
    } else {
        printf("no CUDA device: compute entry points not called\n");
    }
    return 0;
}
