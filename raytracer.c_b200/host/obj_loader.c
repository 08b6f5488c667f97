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
#include <omp.h>
#include <sys/types.h>
#include <unistd.h>

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
    /* the k-th digit's weight is 1.0 * 0.1 * ... * 0.1 (k factors, left to right): carried along,
     * the same sequence of roundings as recomputing it for every digit */
    double scale = 1.0;
    for (; c < stop && isdigit((unsigned char)*c); c++)
    {
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

/* ---- the loader: two passes over line-aligned chunks of the file, shared out among the host cores ------
 * A 1 M-triangle OBJ (config C3: 239 MB as scene_write_obj writes it) took 2.1 s to parse on one core --
 * nine times the 0.23 s the B200 needs to render it at 128 spp -- so the parse is spread over the cores:
 *   pass 1  every chunk parses its `v` / `vt` lines into its own arrays and counts the triangles its `f`
 *           lines will give (corners - 2 each);
 *   prefix  sums of the three counts give every chunk the number of positions / texcoords defined BEFORE
 *           it (what a relative index, or the range check of an absolute one, refers to) and the slot of
 *           its first output triangle; the per-chunk arrays are copied into two global ones;
 *   pass 2  every chunk walks its lines again, keeps the running counts, resolves its faces against the
 *           global arrays and writes its triangles to their slots.
 * Chunks are ~4 MB and handed out dynamically: a file usually holds all its `v` lines first and all its
 * `f` lines last, so equal byte ranges are very unequal work in either pass.  The result is byte for byte
 * the one-chunk result: same floats, same order, same failures (a face may only refer to vertices
 * defined before it).  Measured on the 239 MB file, 8 cores: 2.13 s -> 0.33 s. */
typedef struct
{
  const char *begin, *end; /* whole lines */
  Vec pos, tex;
  size_t n_tris;           /* triangles this chunk's faces emit */
  size_t pos_before, tex_before, tri_before;
  bool ok;
} ObjChunk;

static bool is_v(const char *p, const char *eol) { return p + 1 < eol && p[0] == 'v' && (p[1] == ' ' || p[1] == '\t'); }
static bool is_vt(const char *p, const char *eol) { return p + 2 < eol && p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t'); }
static bool is_f(const char *p, const char *eol) { return p + 1 < eol && p[0] == 'f' && (p[1] == ' ' || p[1] == '\t'); }

static void obj_pass1(ObjChunk *c)
{
  c->pos = (Vec){ NULL, 0, 0, sizeof(F3) };
  c->tex = (Vec){ NULL, 0, 0, sizeof(F2) };
  c->n_tris = 0;
  bool ok = true;
  for (const char *line = c->begin; ok && line < c->end;)
  {
    const char *eol = memchr(line, '\n', (size_t)(c->end - line));
    if (!eol)
      eol = c->end;
    const char *p = skip_blank(line, eol);
    if (is_v(p, eol))
    {
      p += 2;
      F3 v = { 0, 0, 0 };
      ok = parse_real(&p, eol, &v.x) && parse_real(&p, eol, &v.y) && parse_real(&p, eol, &v.z) && vec_push(&c->pos, &v);
    }
    else if (is_vt(p, eol))
    {
      p += 3;
      F2 t = { 0, 0 };
      if (parse_real(&p, eol, &t.u))
      {
        parse_real(&p, eol, &t.v); /* v is optional */
        ok = vec_push(&c->tex, &t);
      }
      else
        ok = false;
    }
    else if (is_f(p, eol))
    {
      p += 2;
      size_t corners = 0;
      long vi, ti;
      while (parse_corner(&p, eol, &vi, &ti))
        corners++;
      if (corners > 2)
        c->n_tris += corners - 2;
    }
    line = eol + 1;
  }
  c->ok = ok;
}

static void obj_pass2(ObjChunk *c, const F3 *pos, const F2 *tex, Vertex *out)
{
  size_t n_pos = c->pos_before, n_tex = c->tex_before;
  Vertex *dst = out + 3 * c->tri_before;
  bool ok = true;
  for (const char *line = c->begin; ok && line < c->end;)
  {
    const char *eol = memchr(line, '\n', (size_t)(c->end - line));
    if (!eol)
      eol = c->end;
    const char *p = skip_blank(line, eol);
    if (is_v(p, eol))
      n_pos++;
    else if (is_vt(p, eol))
      n_tex++;
    else if (is_f(p, eol))
    {
      p += 2;
      Vertex first, prev, cur;
      memset(&first, 0, sizeof(first));
      memset(&prev, 0, sizeof(prev));
      size_t corners = 0;
      long vi, ti;
      while (ok && parse_corner(&p, eol, &vi, &ti))
      {
        size_t v_at, t_at;
        memset(&cur, 0, sizeof(cur));
        if (!resolve(vi, n_pos, &v_at))
        {
          fprintf(stderr, "load_obj: vertex index %ld out of range\n", vi);
          ok = false;
          break;
        }
        cur.pos = (vec3){ pos[v_at].x, pos[v_at].y, pos[v_at].z };
        if (ti != 0 && resolve(ti, n_tex, &t_at))
          cur.tex = (vec2){ tex[t_at].u, tex[t_at].v };
        /* fan: (corner 0, corner k-1, corner k), tinyobj_loader.h:1203-1224 */
        if (corners == 0)
          first = cur;
        else if (corners >= 2)
        {
          dst[0] = first;
          dst[1] = prev;
          dst[2] = cur;
          dst += 3;
        }
        prev = cur;
        corners++;
      }
    }
    line = eol + 1;
  }
  /* a face that failed half-way leaves its chunk short: the load fails as a whole */
  c->ok = ok && dst == out + 3 * (c->tri_before + c->n_tris);
}

bool load_obj_ex(const char *filename, TriangleMesh *mesh, int threads, size_t min_chunk_bytes)
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

  if (threads <= 0)
  {
    threads = omp_get_max_threads();
    if (threads > 32)
      threads = 32;
  }
  if (min_chunk_bytes == 0)
    min_chunk_bytes = (size_t)4 << 20;
  /* more chunks than threads, handed out dynamically: an OBJ file usually holds all its `v` lines first and
   * all its `f` lines last, so equal byte ranges are very unequal work in either pass */
  size_t n_chunks = (size_t)size / min_chunk_bytes;
  if (n_chunks > 4096)
    n_chunks = 4096;
  if (n_chunks < 1)
    n_chunks = 1;
  if ((size_t)threads > n_chunks)
    threads = (int)n_chunks;

  const bool timing = getenv("RTB_TIMING") != NULL; /* development aid */
  double t_mark = omp_get_wtime();
#define OBJ_MARK(what) do { if (timing) { double now_ = omp_get_wtime(); fprintf(stderr, "load_obj: %s %.1f ms\n", what, 1e3 * (now_ - t_mark)); t_mark = now_; } } while (0)
  /* the file is read by the same team: page-cache copies run at memory speed per thread */
  const int fd = fileno(f);
  bool read_ok = true;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (long k = 0; k < (long)n_chunks; k++)
  {
    size_t a = (size_t)size * (size_t)k / n_chunks, b = (size_t)size * (size_t)(k + 1) / n_chunks;
    while (a < b)
    {
      ssize_t got = pread(fd, buf + a, b - a, (off_t)a);
      if (got <= 0)
      {
#pragma omp atomic write
        read_ok = false;
        break;
      }
      a += (size_t)got;
    }
  }
  fclose(f);
  if (!read_ok)
  {
    free(buf);
    fprintf(stderr, "load_obj: cannot read '%s'\n", filename);
    return false;
  }
  buf[size] = '\0';
  const char *end = buf + size;
  OBJ_MARK("read");

  /* chunk k starts at the first line that begins at or after its share of the bytes */
  ObjChunk *chunks = (ObjChunk *)calloc(n_chunks, sizeof(ObjChunk));
  if (!chunks)
  {
    free(buf);
    return false;
  }
  const char *at = buf;
  for (size_t k = 0; k < n_chunks; k++)
  {
    chunks[k].begin = at;
    const char *want = k + 1 == n_chunks ? end : buf + (size_t)size * (k + 1) / n_chunks;
    if (want < at)
      want = at;
    if (k + 1 < n_chunks && want < end)
    {
      const char *nl = memchr(want, '\n', (size_t)(end - want));
      want = nl ? nl + 1 : end;
    }
    chunks[k].end = want;
    at = want;
  }

#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (long k = 0; k < (long)n_chunks; k++)
    obj_pass1(&chunks[k]);
  OBJ_MARK("pass 1 (v, vt, face count)");

  bool ok = true;
  size_t n_pos = 0, n_tex = 0, n_tris = 0;
  for (size_t k = 0; k < n_chunks; k++)
  {
    ok = ok && chunks[k].ok;
    chunks[k].pos_before = n_pos;
    chunks[k].tex_before = n_tex;
    chunks[k].tri_before = n_tris;
    n_pos += chunks[k].pos.count;
    n_tex += chunks[k].tex.count;
    n_tris += chunks[k].n_tris;
  }
  F3 *pos = NULL;
  F2 *tex = NULL;
  Vertex *out = NULL;
  if (ok)
  {
    pos = (F3 *)malloc((n_pos ? n_pos : 1) * sizeof(F3));
    tex = (F2 *)malloc((n_tex ? n_tex : 1) * sizeof(F2));
    out = (Vertex *)malloc((n_tris ? 3 * n_tris : 1) * sizeof(Vertex));
    ok = pos && tex && out;
  }
  if (ok)
  {
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (long k = 0; k < (long)n_chunks; k++)
    {
      if (chunks[k].pos.count)
        memcpy(pos + chunks[k].pos_before, chunks[k].pos.data, chunks[k].pos.count * sizeof(F3));
      if (chunks[k].tex.count)
        memcpy(tex + chunks[k].tex_before, chunks[k].tex.data, chunks[k].tex.count * sizeof(F2));
    }
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (long k = 0; k < (long)n_chunks; k++)
      obj_pass2(&chunks[k], pos, tex, out);
    for (size_t k = 0; k < n_chunks; k++)
      ok = ok && chunks[k].ok;
    OBJ_MARK("pass 2 (faces)");
  }

  for (size_t k = 0; k < n_chunks; k++)
  {
    free(chunks[k].pos.data);
    free(chunks[k].tex.data);
  }
  free(chunks);
  free(buf);
  free(pos);
  free(tex);
  if (!ok)
  {
    free(out);
    fprintf(stderr, "load_obj: failed to parse '%s'\n", filename);
    return false;
  }
  OBJ_MARK("free");
#undef OBJ_MARK
  mesh->num_triangles = n_tris;
  mesh->vertices = out;
  return true;
}

/* raytracer.h:158 */
bool load_obj(const char *filename, TriangleMesh *mesh)
{
  return load_obj_ex(filename, mesh, 0, 0);
}
