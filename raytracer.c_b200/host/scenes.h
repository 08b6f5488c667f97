/*
 * scenes.h -- host-side scene builders for the benchmark configurations
 * (BASELINE.json `configs`).  C99, no GPU code.
 */
#ifndef RTB200_SCENES_H
#define RTB200_SCENES_H

#include "raytracer.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SCENE_DEFAULT_COUNT 38
#define SCENE_SEED 1666943821u /* the reference's fixed seed, main.c:182 */

/* C1: the reference main.c default scene (main.c:244-397): 6 wall spheres of r=10000,
 * 30 packed spheres, 2 lights.  Wall positions depend on the aspect ratio (quirk Q15).
 * `out` must hold SCENE_DEFAULT_COUNT records. */
size_t scene_default(Object *out, int width, int height);

/* Deterministic host RNG for the generators (PCG32); independent of libc rand(). */
typedef struct { uint64_t state, inc; } ScenePcg;
void scene_pcg_seed(ScenePcg *g, uint64_t seed, uint64_t stream);
double scene_pcg_double(ScenePcg *g); /* [0,1) */

/* Material mix of a generated scene, as cumulative-free fractions (sum <= 1; the rest is
 * plain diffuse).  The reference's own mix (main.c:104-119) is {.5, .2, .2}. */
typedef struct
{
  double emissive;   /* white diffuse with emission in U[0,1)^3 */
  double refraction; /* M_REFRACTION */
  double reflection; /* M_REFLECTION */
} SceneMix;

/* The six wall spheres (main.c:258-299) for a room of the given half extents. */
size_t scene_room_walls(Object *out, double half_w, double half_h, double depth);

/* generate_random_spheres semantics (main.c:65-138): rejection-pack `count`
 * non-overlapping spheres with radius U[rmin,rmax) into the box, material drawn from
 * `mix`.  Returns the number placed (== count unless the box is too full). */
size_t scene_random_spheres(Object *out, size_t count, vec3 box_min, vec3 box_max, double rmin,
                            double rmax, SceneMix mix, uint64_t seed);

/* C2 / C4 / C5 style scene: walls + `count` packed spheres (+ the two reference lights).
 * Allocates *out (free with free()). */
size_t scene_sphere_field(Object **out, size_t count, int width, int height, SceneMix mix, uint64_t seed);

/* C3: procedural height-field mesh, (grid x grid) cells -> 2*grid*grid triangles, wound so
 * that calculate_surface_normal (cross(v2-v0, v1-v0)) points up.  Allocates mesh->vertices. */
void scene_heightfield_mesh(TriangleMesh *mesh, int grid, double half_w, double half_d, double y0,
                            double amplitude);

/* writes the same mesh as an OBJ file (v / vt / f), to exercise load_obj */
bool scene_write_obj(const char *filename, const TriangleMesh *mesh);

/* C3 scene: walls + mesh object + light spheres.  Allocates *out; `mesh` and the Sphere
 * records it points to must outlive the scene (spheres are allocated in one block that is
 * returned through *sphere_block; free both with free()). */
size_t scene_mesh_room(SceneObject **out, Sphere **sphere_block, TriangleMesh *mesh, int width, int height);

/* wrap flat Objects as SceneObjects (allocates *out and *sphere_block) */
size_t scene_from_objects(SceneObject **out, Sphere **sphere_block, const Object *objects, size_t n);

#ifdef __cplusplus
}
#endif
#endif
