/*
 * scenes.c -- scene data for the benchmark configurations.  Host-side C99.
 *
 * scene_default() is the reference's hard-coded scene (main.c:244-397) as DATA: the
 * harness must reproduce it verbatim for config C1.  The
 * generators follow generate_random_spheres / collision (main.c:50-138) but draw from a
 * private PCG32 stream instead of libc rand(), so a scene is a pure function of its seed.
 */
#include "scenes.h"

/* ---- C1: the reference default scene --------------------------------------- */

/* the 30 packed spheres of main.c:350-379: flags, centre, radius, emission; colour is
 * white for all of them */
typedef struct { uint flags; double cx, cy, cz, r, ex, ey, ez; } PackedSphere;

#define D M_DEFAULT
#define R M_REFLECTION
static const PackedSphere k_packed[30] = {
  { D, 11.8823, 12.8165, -3.43022, 3.47138, 0, 0, 0 },
  { R, -4.78617, -10.565, -11.8307, 7.8185, 0, 0, 0 },
  { R, 16.3283, 15.7456, 8.02745, 3.38449, 0, 0, 0 },
  { D, -7.74563, 7.10781, -1.14851, 5.68239, 0.129721, 1.08691, 0.15077 },
  { D, 0.604958, 13.8198, -10.0857, 3.63955, 0, 0, 0 },
  { D, 2.72773, -3.47742, 7.21287, 5.756, 0.419482, 0.406897, 0.301653 },
  { R, -11.6808, -15.0112, 10.6413, 3.40004, 0, 0, 0 },
  { R, 5.28438, -2.58167, -3.87996, 2.20867, 0, 0, 0 },
  { D, -15.1722, -0.318264, -14.8739, 3.31716, 0, 0, 0 },
  { D, 7.05345, -11.9375, -4.08415, 5.01176, 0, 0, 0 },
  { D, -6.64606, 12.5952, -11.8074, 3.57727, 2.02456, 1.14375, 0.22395 },
  { R, 15.3284, 7.63569, -7.88126, 2.26494, 0, 0, 0 },
  { R, 5.15508, -13.4632, 12.9555, 4.41505, 0, 0, 0 },
  { D, 6.61409, 15.9581, 13.6585, 2.76828, 0, 0, 0 },
  { R, 0.00113487, 8.35296, -14.4917, 2.58514, 0, 0, 0 },
  { R, 9.63578, 9.63074, -16.0336, 2.22603, 0, 0, 0 },
  { R, 13.105, 1.55555, 2.67293, 4.00109, 0, 0, 0 },
  { R, -0.0637789, 6.39925, 11.777, 4.99425, 0, 0, 0 },
  { D, 7.11587, 6.96992, 7.24724, 3.28273, 0.403171, 1.90743, 1.59559 },
  { D, -17.0139, 4.27765, 11.924, 2.14903, 0, 0, 0 },
  { D, 15.3924, -4.96949, 12.4327, 3.48512, 0.647167, 1.99216, 1.4463 },
  { R, -16.0135, 15.9701, 12.4844, 3.00053, 0, 0, 0 },
  { D, -2.87246, -15.5185, 7.78116, 3.4779, 3.16375, 4.44267, 3.49332 },
  { R, -8.89639, -10.9745, -1.80553, 2.39033, 0, 0, 0 },
  { D, -0.653194, 9.99867, 4.17957, 3.28669, 0.662701, 2.82942, 1.50879 },
  { D, -14.6767, -6.47449, 4.48493, 4.77854, 1.6413, 2.60242, 0.421142 },
  { D, -9.76604, 16.8809, -0.605894, 2.89667, 0.479186, 0.149559, 0.3761 },
  { D, 4.07601, 5.6942, -3.07305, 4.91388, 0, 0, 0 },
  { R, 15.1469, -13.988, 9.5646, 4.6719, 0, 0, 0 },
  { D, -7.2047, -5.0758, 7.74727, 2.86742, 0, 0, 0 },
};
#undef D
#undef R

static Object make_sphere(uint flags, vec3 center, double radius, vec3 color, vec3 emission)
{
  Object o;
  memset(&o, 0, sizeof(o));
  o.flags = flags;
  o.radius = radius;
  o.center = center;
  o.color = color;
  o.emission = emission;
  return o;
}

/* main.c:258-299 -- floor, back, left (green), right (red), ceiling, front */
size_t scene_room_walls(Object *out, double half_w, double half_h, double depth)
{
  const double radius = 10000;
  const vec3 grey = VECTOR(0.75, 0.75, 0.75);
  const vec3 black = BLACK;
  out[0] = make_sphere(M_DEFAULT, VECTOR(0, -radius - half_h, 0), radius, grey, black);
  out[1] = make_sphere(M_DEFAULT, VECTOR(0, 0, -radius - depth), radius, grey, black);
  out[2] = make_sphere(M_DEFAULT, VECTOR(-radius - half_w, 0, 0), radius, VECTOR(0.25, 0.75, 0.25), black);
  out[3] = make_sphere(M_DEFAULT, VECTOR(radius + half_w, 0, 0), radius, VECTOR(0.75, 0.25, 0.25), black);
  out[4] = make_sphere(M_DEFAULT, VECTOR(0, radius + half_h, 0), radius, grey, black);
  out[5] = make_sphere(M_DEFAULT, VECTOR(0, 0, radius + depth * 2), radius, grey, black);
  return 6;
}

/* the two lights of main.c:383-395 */
static size_t room_lights(Object *out, double room_height)
{
  const double light_radius = 15;
  const double y = -room_height;
  out[0] = make_sphere(M_DEFAULT, VECTOR(0, room_height + light_radius * 0.9, 0), light_radius, WHITE,
                       RGB(0x00 * 15, 0x32 * 15, 0xA0 * 15));
  /* SPHERE(2, y, 12, 3): centre is lifted by the radius (main.c:18-20) */
  out[1] = make_sphere(M_DEFAULT, VECTOR(2, (y) + (3), 12), 3, WHITE, RGB(0xD0, 0x00, 0x70));
  return 2;
}

size_t scene_default(Object *out, int width, int height)
{
  const double aspect_ratio = (double)width / (double)height;
  const double room_depth = 30;
  const double room_height = 20;
  const double room_width = room_height * aspect_ratio;
  size_t n = scene_room_walls(out, room_width, room_height, room_depth);
  for (int i = 0; i < 30; i++)
  {
    const PackedSphere *p = &k_packed[i];
    out[n++] = make_sphere(p->flags, VECTOR(p->cx, p->cy, p->cz), p->r, VECTOR(1, 1, 1),
                           VECTOR(p->ex, p->ey, p->ez));
  }
  n += room_lights(out + n, room_height);
  return n;
}

/* ---- PCG32 (O'Neill 2014, pcg32_random_r of the minimal C implementation) --- */

static uint32_t pcg_next(ScenePcg *g)
{
  uint64_t old = g->state;
  g->state = old * 6364136223846793005ULL + (g->inc | 1);
  uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  uint32_t rot = (uint32_t)(old >> 59u);
  return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
}

void scene_pcg_seed(ScenePcg *g, uint64_t seed, uint64_t stream)
{
  g->state = 0;
  g->inc = (stream << 1u) | 1u;
  pcg_next(g);
  g->state += seed;
  pcg_next(g);
}

double scene_pcg_double(ScenePcg *g)
{
  /* 31-bit resolution like random_double() (raytracer.c:227) */
  return (double)(pcg_next(g) >> 1) / 2147483648.0;
}

static double pcg_range(ScenePcg *g, double lo, double hi) { return scene_pcg_double(g) * (hi - lo) + lo; }

/* ---- generate_random_spheres (main.c:65-138) -------------------------------- */

static bool spheres_collide(vec3 c0, double r0, vec3 c1, double r1) /* main.c:50-53 */
{
  return vec3_length(vec3_sub(c0, c1)) < (r0 + r1);
}

size_t scene_random_spheres(Object *out, size_t count, vec3 box_min, vec3 box_max, double rmin,
                            double rmax, SceneMix mix, uint64_t seed)
{
  ScenePcg g;
  scene_pcg_seed(&g, seed, 54u);
  const long long max_iterations = 100000000LL;
  long long iterations = 0;
  size_t found = 0;

  /* uniform grid over the box so the overlap test stays O(1) per candidate at 10k+ spheres
   * (the reference scans all placed spheres; same accept/reject decisions) */
  const double cell = 2.0 * rmax;
  int gx = (int)fmax(1.0, floor((box_max.x - box_min.x) / cell));
  int gy = (int)fmax(1.0, floor((box_max.y - box_min.y) / cell));
  int gz = (int)fmax(1.0, floor((box_max.z - box_min.z) / cell));
  size_t n_cells = (size_t)gx * gy * gz;
  int *head = (int *)malloc(sizeof(int) * n_cells);
  int *next = (int *)malloc(sizeof(int) * (count ? count : 1));
  if (!head || !next)
  {
    fprintf(stderr, "scene_random_spheres: out of memory\n");
    exit(EXIT_FAILURE);
  }
  for (size_t i = 0; i < n_cells; i++)
    head[i] = -1;

  while (found < count && iterations++ < max_iterations)
  {
    double radius = pcg_range(&g, rmin, rmax);
    vec3 vr = { radius, radius, radius };
    vec3 lo = vec3_add(box_min, vr);
    vec3 hi = vec3_sub(box_max, vr);
    vec3 center = { pcg_range(&g, lo.x, hi.x), pcg_range(&g, lo.y, hi.y), pcg_range(&g, lo.z, hi.z) };

    int cx = (int)fmin(gx - 1, fmax(0, floor((center.x - box_min.x) / (box_max.x - box_min.x) * gx)));
    int cy = (int)fmin(gy - 1, fmax(0, floor((center.y - box_min.y) / (box_max.y - box_min.y) * gy)));
    int cz = (int)fmin(gz - 1, fmax(0, floor((center.z - box_min.z) / (box_max.z - box_min.z) * gz)));

    bool hit = false;
    for (int dz = -1; dz <= 1 && !hit; dz++)
      for (int dy = -1; dy <= 1 && !hit; dy++)
        for (int dx = -1; dx <= 1 && !hit; dx++)
        {
          int x = cx + dx, y = cy + dy, z = cz + dz;
          if (x < 0 || y < 0 || z < 0 || x >= gx || y >= gy || z >= gz)
            continue;
          for (int k = head[((size_t)z * gy + y) * gx + x]; k >= 0; k = next[k])
            if (spheres_collide(out[k].center, out[k].radius, center, radius))
            {
              hit = true;
              break;
            }
        }
    if (hit)
      continue;

    uint flags = M_DEFAULT;
    vec3 color = WHITE;
    vec3 emission = BLACK;
    double r = scene_pcg_double(&g);
    if (r < mix.emissive)
      emission = (vec3){ scene_pcg_double(&g), scene_pcg_double(&g), scene_pcg_double(&g) };
    else if (r < mix.emissive + mix.refraction)
      flags = M_REFRACTION;
    else if (r < mix.emissive + mix.refraction + mix.reflection)
      flags = M_REFLECTION;

    out[found] = make_sphere(flags, center, radius, color, emission);
    size_t c = ((size_t)cz * gy + cy) * gx + cx;
    next[found] = head[c];
    head[c] = (int)found;
    found++;
  }
  free(head);
  free(next);
  return found;
}

size_t scene_sphere_field(Object **out, size_t count, int width, int height, SceneMix mix, uint64_t seed)
{
  const double aspect_ratio = (double)width / (double)height;
  const double room_depth = 30, room_height = 20;
  const double room_width = room_height * aspect_ratio;
  Object *scene = (Object *)malloc(sizeof(Object) * (count + 8));
  if (!scene)
  {
    fprintf(stderr, "scene_sphere_field: out of memory\n");
    exit(EXIT_FAILURE);
  }
  size_t n = scene_room_walls(scene, room_width, room_height, room_depth);
  /* Radii scaled from the reference's [2,8) so that `count` spheres fill the room about as
   * densely as its 30 do: fill ~ count * r^3, so r scales with (30/count)^(1/3). */
  double scale = cbrt(30.0 / (double)(count > 30 ? count : 30));
  double rmin = 2.0 * scale, rmax = 8.0 * scale;
  n += scene_random_spheres(scene + n, count, VECTOR(-room_width, -room_height, -room_depth),
                            VECTOR(room_width, room_height, room_depth), rmin, rmax, mix, seed);
  n += room_lights(scene + n, room_height);
  *out = scene;
  return n;
}

/* ---- C3: height-field mesh --------------------------------------------------- */

static double terrain_height(double x, double z, double amplitude)
{
  return amplitude * (0.55 * sin(0.21 * x + 0.3) * cos(0.17 * z - 0.4) + 0.30 * sin(0.53 * x - 0.37 * z) +
                      0.15 * cos(1.31 * x + 0.9) * sin(1.13 * z));
}

static Vertex terrain_vertex(int i, int j, int grid, double half_w, double half_d, double y0, double amplitude)
{
  double fx = (double)i / grid, fz = (double)j / grid;
  double x = -half_w + 2.0 * half_w * fx;
  double z = -half_d + 2.0 * half_d * fz;
  Vertex v;
  /* float-representable, like positions that went through an OBJ file (tinyobj parses
   * `float`, tinyobj_loader.h:1152-1158) */
  v.pos.x = (double)(float)x;
  v.pos.y = (double)(float)(y0 + terrain_height(x, z, amplitude));
  v.pos.z = (double)(float)z;
  v.tex.x = (double)(float)fx;
  v.tex.y = (double)(float)fz;
  return v;
}

void scene_heightfield_mesh(TriangleMesh *mesh, int grid, double half_w, double half_d, double y0,
                            double amplitude)
{
  size_t n_tri = (size_t)2 * grid * grid;
  Vertex *v = (Vertex *)malloc(sizeof(Vertex) * 3 * n_tri);
  if (!v)
  {
    fprintf(stderr, "scene_heightfield_mesh: out of memory\n");
    exit(EXIT_FAILURE);
  }
  size_t k = 0;
  for (int j = 0; j < grid; j++)
    for (int i = 0; i < grid; i++)
    {
      Vertex p00 = terrain_vertex(i, j, grid, half_w, half_d, y0, amplitude);
      Vertex p10 = terrain_vertex(i + 1, j, grid, half_w, half_d, y0, amplitude);
      Vertex p01 = terrain_vertex(i, j + 1, grid, half_w, half_d, y0, amplitude);
      Vertex p11 = terrain_vertex(i + 1, j + 1, grid, half_w, half_d, y0, amplitude);
      /* winding chosen so cross(v2-v0, v1-v0) (raytracer.c:44) points to +y */
      v[k++] = p00; v[k++] = p10; v[k++] = p01;
      v[k++] = p11; v[k++] = p01; v[k++] = p10;
    }
  mesh->num_triangles = n_tri;
  mesh->vertices = v;
}

bool scene_write_obj(const char *filename, const TriangleMesh *mesh)
{
  FILE *f = fopen(filename, "w");
  if (!f)
    return false;
  fprintf(f, "# generated by scene_write_obj\no mesh\n");
  size_t nv = mesh->num_triangles * 3;
  for (size_t i = 0; i < nv; i++)
    fprintf(f, "v %.9g %.9g %.9g\n", mesh->vertices[i].pos.x, mesh->vertices[i].pos.y, mesh->vertices[i].pos.z);
  for (size_t i = 0; i < nv; i++)
    fprintf(f, "vt %.9g %.9g\n", mesh->vertices[i].tex.x, mesh->vertices[i].tex.y);
  for (size_t t = 0; t < mesh->num_triangles; t++)
    fprintf(f, "f %zu/%zu %zu/%zu %zu/%zu\n", 3 * t + 1, 3 * t + 1, 3 * t + 2, 3 * t + 2, 3 * t + 3, 3 * t + 3);
  fprintf(f, "# end\n");
  return fclose(f) == 0;
}

static SceneObject wrap_sphere(const Object *o, Sphere *slot)
{
  SceneObject s;
  memset(&s, 0, sizeof(s));
  slot->center = o->center;
  slot->radius = o->radius;
  s.type = GEOMETRY_SPHERE;
  s.material.flags = o->flags;
  s.material.color = o->color;
  s.material.emission = o->emission;
  s.geometry.sphere = slot;
  return s;
}

size_t scene_from_objects(SceneObject **out, Sphere **sphere_block, const Object *objects, size_t n)
{
  SceneObject *so = (SceneObject *)malloc(sizeof(SceneObject) * (n ? n : 1));
  Sphere *sp = (Sphere *)malloc(sizeof(Sphere) * (n ? n : 1));
  if (!so || !sp)
  {
    fprintf(stderr, "scene_from_objects: out of memory\n");
    exit(EXIT_FAILURE);
  }
  for (size_t i = 0; i < n; i++)
    so[i] = wrap_sphere(&objects[i], &sp[i]);
  *out = so;
  *sphere_block = sp;
  return n;
}

size_t scene_mesh_room(SceneObject **out, Sphere **sphere_block, TriangleMesh *mesh, int width, int height)
{
  const double aspect_ratio = (double)width / (double)height;
  const double room_depth = 30, room_height = 20;
  const double room_width = room_height * aspect_ratio;
  Object flat[16];
  size_t n = scene_room_walls(flat, room_width, room_height, room_depth);
  /* emissive spheres above the terrain (the lights), a mirror and a dielectric ball */
  flat[n++] = make_sphere(M_DEFAULT, VECTOR(0, room_height + 15 * 0.9, 0), 15, WHITE,
                          RGB(0x00 * 15, 0x32 * 15, 0xA0 * 15));
  flat[n++] = make_sphere(M_DEFAULT, VECTOR(-14, 6, -8), 3.5, WHITE, VECTOR(3.16375, 4.44267, 3.49332));
  flat[n++] = make_sphere(M_DEFAULT, VECTOR(15, 2, 6), 3, WHITE, VECTOR(2.02456, 1.14375, 0.22395));
  flat[n++] = make_sphere(M_DEFAULT, VECTOR(2, -3, 14), 2, WHITE, RGB(0xD0 * 3, 0x00, 0x70 * 3));
  flat[n++] = make_sphere(M_REFLECTION, VECTOR(-6, -1, 4), 5, WHITE, BLACK);
  flat[n++] = make_sphere(M_REFRACTION, VECTOR(9, -2, -4), 4.5, WHITE, BLACK);

  SceneObject *so = (SceneObject *)malloc(sizeof(SceneObject) * (n + 1));
  Sphere *sp = (Sphere *)malloc(sizeof(Sphere) * n);
  if (!so || !sp)
  {
    fprintf(stderr, "scene_mesh_room: out of memory\n");
    exit(EXIT_FAILURE);
  }
  size_t k = 0;
  for (size_t i = 0; i < 6; i++, k++) /* walls first, like the reference scene */
    so[k] = wrap_sphere(&flat[i], &sp[i]);
  memset(&so[k], 0, sizeof(SceneObject));
  so[k].type = GEOMETRY_MESH;
  so[k].material.flags = M_DEFAULT;
  so[k].material.color = VECTOR(0.75, 0.75, 0.75);
  so[k].material.emission = BLACK;
  so[k].geometry.mesh = mesh;
  k++;
  for (size_t i = 6; i < n; i++, k++)
    so[k] = wrap_sphere(&flat[i], &sp[i]);
  *out = so;
  *sphere_block = sp;
  return k;
}
