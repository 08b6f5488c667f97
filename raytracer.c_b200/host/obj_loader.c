/*
 * obj_loader.c -- load_obj(): Wavefront OBJ -> TriangleMesh.
 *
 * The reference declares `bool load_obj(const char*, TriangleMesh*)` (raytracer.h:158) but
 * never defines it; it vendors tinyobjloader-c (lib/tinyobj_loader.h) without calling it.
 * That third-party file is not copied here.  This is a small purpose-built reader with
 * the observable behaviour the tinyobj path would have for this use
 * (tinyobj_parse_obj(..., TINYOBJ_FLAG_TRIANGULATE), tinyobj_loader.h:118-144):
 *   - `v` positions and `vt` texcoords are parsed as double and narrowed to float
 *     (tinyobj_loader.h:470-481), so every coordinate is float-representable;
 *   - faces with more than 3 corners are fan-triangulated (i0, i[k-1], i[k])
 *     (tinyobj_loader.h:1203-1224); corners may be `v`, `v/vt`, `v//vn`, `v/vt/vn`;
 *   - indices are 1-based, negative indices are relative to the current count;
 *   - a corner without a texcoord gets (0,0);
 *   - `vn`, `o`, `g`, `s`, `usemtl`, `mtllib` and comments are skipped: the path tracer
 *     uses flat normals from calculate_surface_normal and per-object materials.
 * Unlike tinyobj there is no line-length limit, no per-face corner limit and the last
 * line needs no trailing newline (cube.obj:30-31 carries a workaround for that).
 */
#include "raytracer.h"

#include <ctype.h>
#include <errno.h>

typedef struct { float x, y, z; } F3;
typedef struct { float u, v; } F2;

typedef struct
{
  void *data;
  size_t count, capacity, elem;
} Vec;

static bool vec_push(Vec *v, const void *item)
{
  if (v->count == v->capacity)
  {
    size_t cap = v->capacity ? v->capacity * 2 : 1024;
    void *p = realloc(v->data, cap * v->elem);
    if (!p)
      return false;
    v->data = p;
    v->capacity = cap;
  }
  memcpy((char *)v->data + v->count * v->elem, item, v->elem);
  v->count++;
  return true;
}

static const char *skip_blank(const char *p, const char *end)
{
  while (p < end && (*p == ' ' || *p == '\t' || *p == '\r'))
    p++;
  return p;
}

static bool parse_real(const char **pp, const char *end, float *out)
{
  const char *p = skip_blank(*pp, end);
  if (p >= end)
    return false;
  char *stop = NULL;
  double val = strtod(p, &stop); /* the buffer is NUL-terminated */
  if (stop == p)
    return false;
  *out = (float)val;
  *pp = stop;
  return true;
}

/* one face corner: v[/[vt][/vn]] */
static bool parse_corner(const char **pp, const char *end, long *vi, long *ti)
{
  const char *p = skip_blank(*pp, end);
  if (p >= end || *p == '\n' || *p == '#')
    return false;
  char *stop = NULL;
  *vi = strtol(p, &stop, 10);
  if (stop == p)
    return false;
  p = stop;
  *ti = 0;
  if (p < end && *p == '/')
  {
    p++;
    if (p < end && *p != '/' && !isspace((unsigned char)*p))
    {
      *ti = strtol(p, &stop, 10);
      p = stop;
    }
    if (p < end && *p == '/')
    {
      p++;
      (void)strtol(p, &stop, 10); /* normal index, unused */
      p = stop;
    }
  }
  *pp = p;
  return true;
}

static bool resolve(long idx, size_t count, size_t *out)
{
  if (idx > 0 && (size_t)idx <= count)
  {
    *out = (size_t)idx - 1;
    return true;
  }
  if (idx < 0 && (size_t)(-idx) <= count)
  {
    *out = count - (size_t)(-idx);
    return true;
  }
  return false;
}

void free_mesh(TriangleMesh *mesh)
{
  if (!mesh)
    return;
  free(mesh->vertices);
  mesh->vertices = NULL;
  mesh->num_triangles = 0;
}

bool load_obj(const char *filename, TriangleMesh *mesh)
{
  if (!filename || !mesh)
    return false;
  mesh->num_triangles = 0;
  mesh->vertices = NULL;

  FILE *f = fopen(filename, "rb");
  if (!f)
  {
    fprintf(stderr, "load_obj: cannot open '%s': %s\n", filename, strerror(errno));
    return false;
  }
  fseek(f, 0, SEEK_END);
  long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (size < 0)
  {
    fclose(f);
    return false;
  }
  char *buf = (char *)malloc((size_t)size + 1);
  if (!buf)
  {
    fclose(f);
    return false;
  }
  size_t got = fread(buf, 1, (size_t)size, f);
  fclose(f);
  buf[got] = '\0';
  const char *end = buf + got;

  Vec pos = { NULL, 0, 0, sizeof(F3) };
  Vec tex = { NULL, 0, 0, sizeof(F2) };
  Vec out = { NULL, 0, 0, sizeof(Vertex) };
  Vec corners = { NULL, 0, 0, sizeof(Vertex) };
  bool ok = true;

  for (const char *line = buf; ok && line < end;)
  {
    const char *eol = memchr(line, '\n', (size_t)(end - line));
    if (!eol)
      eol = end;
    const char *p = skip_blank(line, eol);
    if (p + 1 < eol && p[0] == 'v' && (p[1] == ' ' || p[1] == '\t'))
    {
      p += 2;
      F3 v = { 0, 0, 0 };
      if (parse_real(&p, eol, &v.x) && parse_real(&p, eol, &v.y) && parse_real(&p, eol, &v.z))
        ok = vec_push(&pos, &v);
      else
        ok = false;
    }
    else if (p + 2 < eol && p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t'))
    {
      p += 3;
      F2 t = { 0, 0 };
      if (parse_real(&p, eol, &t.u))
      {
        parse_real(&p, eol, &t.v); /* v is optional */
        ok = vec_push(&tex, &t);
      }
      else
        ok = false;
    }
    else if (p + 1 < eol && p[0] == 'f' && (p[1] == ' ' || p[1] == '\t'))
    {
      p += 2;
      corners.count = 0;
      long vi, ti;
      while (ok && parse_corner(&p, eol, &vi, &ti))
      {
        size_t v_at, t_at;
        Vertex vert;
        memset(&vert, 0, sizeof(vert));
        if (!resolve(vi, pos.count, &v_at))
        {
          fprintf(stderr, "load_obj: vertex index %ld out of range\n", vi);
          ok = false;
          break;
        }
        F3 pv = ((F3 *)pos.data)[v_at];
        vert.pos = (vec3){ pv.x, pv.y, pv.z };
        if (ti != 0 && resolve(ti, tex.count, &t_at))
        {
          F2 tv = ((F2 *)tex.data)[t_at];
          vert.tex = (vec2){ tv.u, tv.v };
        }
        ok = vec_push(&corners, &vert);
      }
      const Vertex *c = (const Vertex *)corners.data;
      for (size_t k = 2; ok && k < corners.count; k++)
        ok = vec_push(&out, &c[0]) && vec_push(&out, &c[k - 1]) && vec_push(&out, &c[k]);
    }
    line = eol + 1;
  }

  free(buf);
  free(pos.data);
  free(tex.data);
  free(corners.data);
  if (!ok)
  {
    free(out.data);
    fprintf(stderr, "load_obj: failed to parse '%s'\n", filename);
    return false;
  }
  mesh->num_triangles = out.count / 3;
  mesh->vertices = (Vertex *)out.data;
  return true;
}
