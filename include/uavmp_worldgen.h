/* uavmp_worldgen.h — synthetic worlds for tests and bench.py (HOST utility, its own library: libuavmp_worldgen.so).
 *
 * It defines the INPUTS both the CPU oracle and the CUDA path consume and is deliberately not part of
 * libuavmp.so: the CPU reference arm of bench.py generates its inputs without mapping the product library.
 */
#ifndef UAVMP_WORLDGEN_H
#define UAVMP_WORLDGEN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* synthetic world generator parameters (host utility; random_forest.cpp:509-535, simulator.xml:16-41) */
typedef struct {
  int map_type; /* 0 random pillars + rings, 2 two-slab wall */
  uint32_t seed;
  double x_size, y_size, resolution;
  double init_x, init_y, init_radius;
  int polar_num, circle_num;
  double w_l, w_h, h_l, h_h;
  double radius_l, radius_h, z_l, z_h, theta;
  double wall_x, wall_y, wall_w;
} uavmp_mapgen_params;

void uavmp_mapgen_params_default(uavmp_mapgen_params* p, double x_size, double y_size, uint32_t seed);
int uavmp_mapgen_cloud(const uavmp_mapgen_params* p, float* cloud_xyz, int cap); /* returns point count */
/* GridMap::cloudCallback's inflation on the host (grid_map.cpp:733-785): the definition the device version is tested against */
int uavmp_grid_inflate_host(const float* cloud_xyz, int n, const double origin[3], const double map_size[3],
                            double resolution, double obstacles_inflation, int8_t* occ_inflate, int nx, int ny, int nz);

#ifdef __cplusplus
}
#endif
#endif /* UAVMP_WORLDGEN_H */
