/* Plain-C float64 pieces of the ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/plb_oracle.py for the rules).
 *
 * build_target_sdf: Loss.update_target / update_target_sdf (plb/engine/losses/loss.py:81-106), literally:
 * 2*n Jacobi sweeps; per node, neighbours with offsets in [-3,3)^3 in lexicographic order, strict '<' update,
 * norm with eps 1e-8 (loss.py:77-79).  A sweep is a pure function of the previous sweep's (sdf, nearest) pair, so
 * iteration stops at the first fixed point.  Validated against the numpy statement in plb_oracle.build_target_sdf.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

int oracle_build_target_sdf(int n, double dx, const double* density, double* sdf_out) {
    const double inf = 1000.0;
    long long total = (long long)n * n * n;
    double* sdf_c = (double*)malloc(sizeof(double) * total);
    double* sdf_o = (double*)malloc(sizeof(double) * total);
    double* near_c = (double*)calloc(3 * total, sizeof(double));
    double* near_o = (double*)calloc(3 * total, sizeof(double));
    if (!sdf_c || !sdf_o || !near_c || !near_o) return -1;
    for (long long i = 0; i < total; i++) sdf_c[i] = inf;
    int sweeps = 0;
    for (int it = 0; it < 2 * n; it++) {
        int changed = 0;
#pragma omp parallel for collapse(2) reduction(| : changed)
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++)
                for (int k = 0; k < n; k++) {
                    long long node = ((long long)i * n + j) * n + k;
                    double gx = i * dx, gy = j * dx, gz = k * dx;
                    double best = inf, bx = near_c[node * 3], by = near_c[node * 3 + 1], bz = near_c[node * 3 + 2];
                    if (density[node] > 1e-4) {
                        best = 0.0; bx = gx; by = gy; bz = gz;
                    } else {
                        for (int a = -3; a < 3; a++)
                            for (int b = -3; b < 3; b++)
                                for (int c = -3; c < 3; c++) {
                                    int vi = i + a, vj = j + b, vk = k + c;
                                    if (vi < 0 || vj < 0 || vk < 0 || vi >= n || vj >= n || vk >= n) continue;
                                    if (a == 0 && b == 0 && c == 0) continue;
                                    long long v = ((long long)vi * n + vj) * n + vk;
                                    if (sdf_c[v] < inf) {
                                        double px = near_c[v * 3], py = near_c[v * 3 + 1], pz = near_c[v * 3 + 2];
                                        double ddx = gx - px, ddy = gy - py, ddz = gz - pz;
                                        double dist = sqrt(ddx * ddx + ddy * ddy + ddz * ddz + 1e-8);
                                        if (dist < best) { best = dist; bx = px; by = py; bz = pz; }
                                    }
                                }
                    }
                    if (best != sdf_c[node] || bx != near_c[node * 3] || by != near_c[node * 3 + 1] || bz != near_c[node * 3 + 2]) changed |= 1;
                    sdf_o[node] = best;
                    near_o[node * 3] = bx; near_o[node * 3 + 1] = by; near_o[node * 3 + 2] = bz;
                }
        double* t = sdf_c; sdf_c = sdf_o; sdf_o = t;
        t = near_c; near_c = near_o; near_o = t;
        sweeps++;
        if (!changed) break;
    }
    memcpy(sdf_out, sdf_c, sizeof(double) * total);
    free(sdf_c); free(sdf_o); free(near_c); free(near_o);
    return sweeps;
}
