/*
 * oracle.h -- TEST INFRASTRUCTURE.  Interface of the CPU restatement in oracle.c.
 * See the header of oracle.c for what it follows in the reference and how it is pinned.
 * Scene structs come from the repo's include/raytracer.h (ABI-identical to the
 * reference's raytracer.h:60-131).
 */
#ifndef RTB200_ORACLE_H
#define RTB200_ORACLE_H

#include <stdint.h>
#include <stddef.h>

enum { ORACLE_RNG_LIBC = 0, ORACLE_RNG_STREAM = 1, ORACLE_RNG_PHILOX = 2 };
enum { ORACLE_DIELECTRIC_STOCHASTIC = 0, ORACLE_DIELECTRIC_SPLIT = 1 };
enum { ORACLE_INTEGRATOR_PATH = 0, ORACLE_INTEGRATOR_WHITTED = 1 };

typedef struct
{
  int max_depth;       /* run-time MAX_DEPTH (raytracer.h:25) */
  int rng_mode;        /* ORACLE_RNG_* */
  int dielectric_mode; /* ORACLE_DIELECTRIC_* */
  int sample_offset;   /* keyed RNG: global index of the first sample */
  uint64_t seed;       /* srand() seed (LIBC) or Philox key (PHILOX) */
  int threads;         /* OpenMP threads (keyed RNG only; LIBC is sequential) */
  int integrator;      /* ORACLE_INTEGRATOR_*: trace_path (raytracer.c:482) or cast_ray (raytracer.c:556) */
} OracleParams;

void oracle_params_default(OracleParams *p);

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void oracle_keyed_jitter(uint64_t seed, uint32_t pixel, uint32_t sample, double *j2);

void oracle_init_camera(Camera *camera, const double *pos, const double *target, int width, int height);
void oracle_camera_ray(const Camera *camera, double u, double v, double *ray6);

int oracle_intersect_sphere(const double *ray6, const double *center, double radius, double *t);
int oracle_intersect_triangle(const double *ray6, const double *verts15, double *tuv);
void oracle_surface_normal(const double *v9, double *n3);
void oracle_reflect(const double *in3, const double *n3, double *out3);
void oracle_refract(const double *in3, const double *n3, double iot, double *out3);
void oracle_checkered(const double *color3, double u, double v, double M, double *out3);

void oracle_render_sum(double *sum_rgb, const SceneObject *objects, size_t n, const Camera *camera,
                       int width, int height, int samples, const OracleParams *p, long long *counters);
void oracle_tonemap(uint8_t *fb, const double *sum_rgb, int width, int height, int total_samples);
void oracle_render(uint8_t *fb, const SceneObject *objects, size_t n, const Camera *camera, int width,
                   int height, int samples, const OracleParams *p, long long *counters);

void oracle_intersect_rays(const SceneObject *objects, size_t n_obj, const double *rays, long long n_rays,
                           int32_t *ids, int64_t *prims, double *ts, double *points, double *normals,
                           double *uvs, int threads);
long long oracle_trace_path_stream(const SceneObject *objects, size_t n_obj, const double *ray6, int depth,
                                   const OracleParams *p, const int32_t *stream, long long stream_len,
                                   double *radiance3);
void oracle_path_records(const SceneObject *objects, size_t n, const Camera *camera, int width, int height,
                         int sample, int n_vertices, const OracleParams *p, int32_t *ids, double *points,
                         double *normals, double *dists, double *radiance);

void oracle_cast_rays(const SceneObject *objects, size_t n_obj, const double *rays, long long n_rays,
                      int max_depth, double *rgb, long long *ray_counts, int threads);

#endif
