/* Minimal C client of liblio_b200.so: one sweep through stage A, then the host-only wire-format helpers.
 * Shows that include/lio_b200.h is plain C (C99), with no C++ or torch types at the boundary.
 *   gcc -std=c99 -Iinclude examples/minimal.c -Llio_mapping_b200 -llio_b200 -Wl,-rpath,$PWD/lio_mapping_b200 -lm -o minimal */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "lio_b200.h"

int main(void) {
  printf("liblio_b200 version %d, %d CUDA device(s)\n", lio_version(), lio_device_count());

  /* host-only: /compact_data round trip */
  float tf7[7] = {0.f, 0.f, 0.f, 1.f, 1.f, 2.f, 3.f};
  float corner[2 * 4] = {1, 2, 3, 0.1f, 4, 5, 6, 1.2f}, surf[1 * 4] = {7, 8, 9, 2.3f}, full[1 * 4] = {0, 0, 1, 3.4f};
  float msg[(3 + 4) * 4], tf_out[7], c2[8], s2[4], f2[4];
  int n_points = 0, sizes[3];
  if (lio_compact_encode(tf7, corner, 2, surf, 1, full, 1, msg, 7, &n_points) != LIO_OK) return 1;
  if (lio_compact_sizes(msg, n_points, sizes) != LIO_OK || sizes[0] != 2 || sizes[1] != 1 || sizes[2] != 1) return 2;
  if (lio_compact_decode(msg, n_points, tf_out, c2, s2, f2) != LIO_OK || tf_out[6] != 3.f || c2[7] != 1.2f) return 3;

  if (lio_device_count() <= 0) {
    printf("no device: compute entry points return LIO_ERR_NO_DEVICE (%d)\n", LIO_ERR_NO_DEVICE);
    return 0;
  }

  /* stage A on a synthetic 16-ring sweep */
  const int rings = 16, cols = 900, n = rings * cols;
  float *sweep = (float *)malloc(sizeof(float) * 4 * n);
  for (int c = 0; c < cols; ++c)
    for (int r = 0; r < rings; ++r) {
      const float az = -6.2831853f * (float)c / cols, el = (-15.f + 2.f * r) * 0.01745329f, range = 5.f + 2.f * sinf(3.f * az);
      float *p = sweep + 4 * (c * rings + r);
      p[0] = range * cosf(el) * cosf(az); p[1] = range * cosf(el) * sinf(az); p[2] = range * sinf(el); p[3] = 0.f;
    }
  lio_pp_config cfg;
  lio_pp_default_config(&cfg);
  cfg.lower_bound = -15.f; cfg.upper_bound = 15.f; cfg.num_rings = rings;
  lio_pp *pp = NULL;
  if (lio_pp_create(&cfg, n, 0, NULL, &pp) != LIO_OK) { printf("create failed: %s\n", lio_last_error()); return 4; }
  if (lio_pp_process_host(pp, sweep, n) != LIO_OK) { printf("process failed: %s\n", lio_last_error()); return 5; }
  int cloud_sizes[LIO_PP_NUM_CLOUDS];
  lio_pp_cloud_sizes(pp, cloud_sizes);
  printf("sharp %d  less-sharp %d  flat %d  less-flat %d\n", cloud_sizes[LIO_PP_CORNER_SHARP], cloud_sizes[LIO_PP_CORNER_LESS_SHARP], cloud_sizes[LIO_PP_SURF_FLAT],
         cloud_sizes[LIO_PP_SURF_LESS_FLAT]);
  lio_pp_destroy(pp);
  free(sweep);
  return 0;
}
