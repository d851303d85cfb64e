/* minimal_c_abi.c — the C ABI from plain C (no openMVG, no C++, no Python).
 *   gcc -std=c99 examples/minimal_c_abi.c -Iinclude -Lopenmvg_b200 -lomvg_b200 -Wl,-rpath,$PWD/openmvg_b200 -lm -o /tmp/minimal_c_abi
 * Needs a B200 at run time; without one every call returns OMVG_E_CUDA and the program says so.
 * MATCH: two images of random descriptors sharing 50 rows -> those 50 matches.
 * BA: a 4-camera / 60-point scene with perturbed landmarks -> the cost drops to the noise floor. */
#include "omvg_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double urand(void) { return rand() / (double)RAND_MAX; }

int main(void)
{
  srand(7);
  /* ---------------------------------------------------------------- MATCH */
  enum { NA = 400, NB = 300, SHARED = 50 };
  static uint8_t a[NA][OMVG_DESC_LEN], b[NB][OMVG_DESC_LEN];
  for (int i = 0; i < NA; ++i) for (int k = 0; k < OMVG_DESC_LEN; ++k) a[i][k] = (uint8_t)(rand() & 255);
  for (int i = 0; i < NB; ++i) for (int k = 0; k < OMVG_DESC_LEN; ++k) b[i][k] = (uint8_t)(rand() & 255);
  for (int i = 0; i < SHARED; ++i) memcpy(b[i], a[2 * i], OMVG_DESC_LEN);
  omvg_match_ctx *m = NULL;
  if (omvg_match_create(&m, 0) != OMVG_OK) { printf("no B200: %s\n", omvg_last_error()); return 0; }
  const uint32_t counts[2] = {NA, NB}, pi[1] = {0}, pj[1] = {1};
  omvg_match_set_images(m, 2, counts);
  omvg_match_upload_host(m, 0, &a[0][0]);
  omvg_match_upload_host(m, 1, &b[0][0]);
  omvg_match_prepare(m);
  const uint64_t *off; const uint32_t *ij; uint64_t n = 0;
  if (omvg_match_run(m, pi, pj, 1, 0.8f) != OMVG_OK || omvg_match_fetch(m, &off, &ij, &n) != OMVG_OK) { printf("match failed: %s\n", omvg_last_error()); return 1; }
  printf("MATCH: %llu matches (expected %d), first (i=%u, j=%u)\n", (unsigned long long)n, SHARED, n ? ij[0] : 0u, n ? ij[1] : 0u);
  omvg_match_destroy(m);

  /* ---------------------------------------------------------------- BA */
  enum { NC = 4, NP = 60 };
  double poses[NC][6], intr[1][OMVG_BA_INTR_STRIDE] = {{1000, 500, 500, 0, 0, 0, 0, 0}}, pts[NP][3], xy[NC * NP][2];
  int32_t model[1] = {OMVG_PINHOLE_CAMERA}, vp[NC], vi[NC], ov[NC * NP], op[NC * NP];
  for (int c = 0; c < NC; ++c) { vp[c] = c; vi[c] = 0; poses[c][0] = 0; poses[c][1] = 0.1 * (c - 1.5); poses[c][2] = 0; poses[c][3] = 0.2 * (c - 1.5); poses[c][4] = 0; poses[c][5] = 4.0; }
  for (int j = 0; j < NP; ++j) { pts[j][0] = urand() - 0.5; pts[j][1] = urand() - 0.5; pts[j][2] = urand() - 0.5; }
  for (int c = 0; c < NC; ++c) for (int j = 0; j < NP; ++j) {           /* project with R = Ry(angle), t */
    const double an = poses[c][1], cs = cos(an), sn = sin(an), X = pts[j][0], Y = pts[j][1], Z = pts[j][2];
    const double x = cs * X + sn * Z + poses[c][3], y = Y + poses[c][4], z = -sn * X + cs * Z + poses[c][5];
    const int o = c * NP + j; ov[o] = c; op[o] = j; xy[o][0] = 500 + 1000 * x / z + (urand() - 0.5); xy[o][1] = 500 + 1000 * y / z + (urand() - 0.5);
  }
  for (int j = 0; j < NP; ++j) for (int k = 0; k < 3; ++k) pts[j][k] += 0.02 * (urand() - 0.5);
  omvg_ba_problem P; memset(&P, 0, sizeof P);
  P.n_poses = NC; P.n_intrinsics = 1; P.n_points = NP; P.n_views = NC; P.n_obs = NC * NP;
  P.poses = &poses[0][0]; P.intrinsics = &intr[0][0]; P.intr_model = model; P.points = &pts[0][0];
  P.view_pose = vp; P.view_intr = vi; P.obs_view = ov; P.obs_point = op; P.obs_xy = &xy[0][0];
  omvg_ba_options O; omvg_ba_default_options(&O);
  omvg_ba_summary S;
  const int rc = omvg_ba_solve(&P, &O, &S);
  printf("BA: rc %d, cost %.3f -> %.3f in %d iterations (%.2f ms on the device)\n", rc, S.initial_cost, S.final_cost, S.iterations, S.device_ms);
  return rc == OMVG_OK && S.final_cost < S.initial_cost ? 0 : 1;
}
