/* The C example of INTEGRATION.md 1b: libdsp_lp.so driven from plain C through dsp_lp_template_create_csr -- no Python-side
 * symbolic setup.  tests/test_cabi.py compiles and links it (CPU); tests/test_gpu_parity.py runs it on the GPU.
 *   min -x0 - 2 x1   s.t.  x0 + x1 + s = p,   0 <= x0 <= 3,  x1, s >= 0     ->   obj = -2 p,  x = (0, p, 0)              */
#include <math.h>
#include <stdio.h>

#include "dsp_lp.h"

int main(void) {
    int32_t A_ptr[] = {0, 3}, A_idx[] = {0, 1, 2};
    double A_val[] = {1, 1, 1};
    double c0[] = {-1, -2, 0}, b0[] = {0}, u0[] = {3, 1e300, 1e300};
    int32_t bptr[] = {0, 1}, bidx[] = {0};
    double bval[] = {1};
    int32_t zptr[] = {0, 0, 0, 0};
    dsp_lp_desc d = {0};
    d.m = 1; d.n = 3; d.Pc = 0; d.Pr = 1;
    d.A_ptr = A_ptr; d.A_idx = A_idx; d.A_val = A_val;
    d.c0 = c0; d.cmap.ptr = zptr;
    d.b0 = b0; d.bmap.ptr = bptr; d.bmap.idx = bidx; d.bmap.val = bval;
    d.u0 = u0; d.umap.ptr = zptr;
    int32_t nb, w;
    if (dsp_lp_analyze_csr(&d, &nb, &w, NULL, NULL, NULL, NULL) != 0 || nb != 1) { printf("analyze failed\n"); return 2; }
    dsp_template *t = NULL;
    int rc = dsp_lp_template_create_csr(&d, &t);
    if (rc != 0) { printf("create failed: %s\n", dsp_lp_last_error()); return 3; }
    double rp[4] = {1, 2, 5, 8}, obj[4], x[12];
    int32_t st[4], it[4];
    rc = dsp_lp_solve_batch_host(t, 4, NULL, rp, 1, NULL, obj, st, it, x, NULL);
    if (rc != 0) { printf("solve failed: %s\n", dsp_lp_last_error()); return 4; }
    int bad = 0;
    for (int k = 0; k < 4; ++k) {
        printf("p = %g: status %d, %d iterations, obj %.9f, x = (%.6f, %.6f, %.6f)\n", rp[k], st[k], it[k], obj[k], x[3 * k], x[3 * k + 1], x[3 * k + 2]);
        if (st[k] != DSP_OPTIMAL || fabs(obj[k] + 2 * rp[k]) > 1e-6 * (1 + 2 * rp[k]) || fabs(x[3 * k + 1] - rp[k]) > 1e-4 * rp[k]) bad = 1;
    }
    dsp_lp_template_destroy(t);
    printf(bad ? "MISMATCH\n" : "C ABI OK\n");
    return bad;
}
