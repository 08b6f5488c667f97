/*
 * obj_loader.c -- load_obj(): Wavefront OBJ -> TriangleMesh.
 *
 * The reference declares `bool load_obj(const char*, TriangleMesh*)` (raytracer.h:158) but
 * never defines it; it vendors tinyobjloader-c (lib/tinyobj_loader.h) without calling it.
 * That third-party file is not copied here.  This is a small purpose-built reader with
 * the observable behaviour the tinyobj path would have for this use
 * (tinyobj_parse_obj(..., TINYOBJ_FLAG_TRIANGULATE), tinyobj_loader.h:118-144):
 *   - `v` positions and `vt` texcoords are parsed as double with tinyobj's own (not correctly
 *     rounded) decimal arithmetic and narrowed to float (tinyobj_loader.h:338-481), so every
 *     coordinate is float-representable and equal to what the tinyobj path gives
 *     (tests/test_obj_loader.py compares the two parsers bit for bit);
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

/* One real number, with the arithmetic of the reference's vendored parser so that the floats are
 * the ones the tinyobj path produces (tinyobj_loader.h:338-468, restated, not copied): the token runs
 * to the next blank; grammar [sign] digits [. digits] [(e|E) [sign] digits], parsed greedily;
 *   - integer digits:    m = m * 10 + digit                       (double)
 *   - k-th decimal digit: m += digit * (0.1 * 0.1 * ... k times)   (double, the power by repeated product)
 *   - exponent e:        value = m * 5^e * 2^e, both powers by repeated product, inverted for e < 0
 * and the double is narrowed to float (tinyobj_loader.h:470-481).  A token that does not start with a
 * sign or a digit, or has no digit after the sign (".5", "-.5"), or an empty exponent, gives 0 -- as
 * upstream, where the failure is silent.  Not strtod(): that one is correctly rounded, this is not,
 * and the two differ in the last bit of the double for a few percent of inputs. */
static bool parse_real(const char **pp, const char *end, float *out)
{
  const char *p = skip_blank(*pp, end);
  if (p >= end)
    return false;
  const char *stop = p;
  while (stop < end && *stop != ' ' && *stop != '\t' && *stop != '\r' && *stop != '\n')
    stop++;
  *pp = stop;
  *out = 0.0f;

  const char *c = p;
  bool negative = false, exp_negative = false, ok = true;
  double m = 0.0;
  int e10 = 0, n_digits = 0;
  if (*c == '+' || *c == '-')
  {
    negative = *c == '-';
    c++;
  }
  else if (!isdigit((unsigned char)*c))
    return true; /* value 0 */
  for (; c < stop && isdigit((unsigned char)*c); c++, n_digits++)
    m = m * 10 + (double)(*c - '0');
  if (n_digits == 0)
    return true;
  if (c < stop && *c == '.')
  {
    c++;
    for (int k = 1; c < stop && isdigit((unsigned char)*c); c++, k++)
    {
      double scale = 1.0;
      for (int j = 0; j < k; j++)
        scale *= 0.1;
      m += (double)(*c - '0') * scale;
    }
  }
  if (c < stop && (*c == 'e' || *c == 'E'))
  {
    c++;
    if (c < stop && (*c == '+' || *c == '-'))
    {
      exp_negative = *c == '-';
      c++;
    }
    else if (!(c < stop && isdigit((unsigned char)*c)))
      ok = false;
    int n_exp = 0;
    for (; ok && c < stop && isdigit((unsigned char)*c); c++, n_exp++)
      e10 = e10 * 10 + (*c - '0');
    if (n_exp == 0)
      ok = false;
  }
  if (!ok)
    return true; /* value 0 */
  double p5 = 1.0, p2 = 1.0;
  for (int k = 0; k < e10; k++)
    p5 = p5 * 5.0;
  for (int k = 0; k < e10; k++)
    p2 = p2 * 2.0;
  if (exp_negative)
  {
    p5 = 1.0 / p5;
    p2 = 1.0 / p2;
  }
  *out = (float)((negative ? -1 : 1) * (m * p5 * p2));
  return true;
}

/* strtol() skips white space, newlines included: an index is only parsed where a digit or a
 * sign actually starts before `end` (the end of the line), so `f 1// 2// 3//` keeps its three
 * corners and a line ending in `/` never reads into the next line */
static bool parse_index(const char **pp, const char *end, long *out)
{
  const char *p = *pp;
  if (p >= end)
    return false;
  const char *q = p;
  if (*q == '-' || *q == '+')
    q++;
  if (q >= end || !isdigit((unsigned char)*q))
    return false;
  char *stop = NULL;
  *out = strtol(p, &stop, 10); /* the digits end at a non-digit no later than `end`'s newline or NUL */
  *pp = stop <= end ? stop : end;
  return true;
}

/* one face corner: v[/[vt][/vn]] */
static bool parse_corner(const char **pp, const char *end, long *vi, long *ti)
{
  const char *p = skip_blank(*pp, end);
  if (p >= end || *p == '\n' || *p == '#')
    return false;
  if (!parse_index(&p, end, vi))
    return false;
  *ti = 0;
  if (p < end && *p == '/')
  {
    p++;
    long t = 0, n = 0;
    if (parse_index(&p, end, &t))
      *ti = t;
    if (p < end && *p == '/')
    {
      p++;
      (void)parse_index(&p, end, &n); /* normal index, unused */
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
  if (idx < 0)
  {
    const unsigned long back = 0ul - (unsigned long)idx; /* defined for LONG_MIN too */
    if (back <= count)
    {
      *out = count - (size_t)back;
      return true;
    }
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
