/* oracle/oracle.h — C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY; see oracle/README.md).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this. */
#ifndef UAVMP_ORACLE_H
#define UAVMP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int allocated_node_num;   /* kino_astar/allocated_node_num   (kino_astar.cpp:8)  */
  int collision_check_type; /* kino_astar/collision_check_type (kino_astar.cpp:9)  */
  double rou_time, lambda_heu, goal_tolerance, time_step_size, max_velocity, max_acceleration, acc_resolution,
      sample_tau;           /* kino_astar.cpp:10-17 */
  double robot_r, robot_h;  /* kino_astar.cpp:18-19 */
  int libm_mode;            /* 0: csrc/fpmath.h (bit-identical to the device), 1: glibc like the reference */
} oracle_kino_params;

typedef struct {
  long long n_pop, n_occ_lookup, n_cloud_pts_tested, n_hash_probe, n_insert, n_update, n_heuristic, n_shot,
      heap_len_sum;
} oracle_kino_counters;

typedef struct {
  int status;       /* 1 REACH_END, 2 NO_PATH_FOUND (kino_astar.h:155-159) */
  int use_node_num; /* kino_astar.h:127 */
  int n_pop;
  int n_path;       /* points appended to `path` */
  int n_path_nodes; /* nodes on the parent chain */
  double shot_duration;
  uint64_t pop_hash; /* digest of the ordered pop sequence (voxel index + state bits) */
  oracle_kino_counters counters;
  uint64_t lookup_digest;   /* digest of the positions the search passes to GridMap::isInMap, in call order */
  long long n_in_map_calls;
} oracle_kino_result;

typedef struct oracle_kino oracle_kino;

oracle_kino* oracle_kino_create(const oracle_kino_params* p, const int8_t* occ_inflate, int nx, int ny, int nz,
                                const double origin[3], const double map_size[3], double resolution,
                                const float* cloud_xyz, int n_cloud);
void oracle_kino_destroy(oracle_kino* k);
int oracle_kino_search(oracle_kino* k, const double start_pt[3], const double start_vel[3], const double end_pt[3],
                       const double end_vel[3], oracle_kino_result* res, double* path_xyz, int path_cap,
                       int32_t* pop_trace, int pop_cap);

double oracle_fp_eval(int op, double x, int n);
void oracle_fp_eval_n(int op, const double* x, int n_pow, double* y, long long n);

#ifdef __cplusplus
}
#endif
#endif
