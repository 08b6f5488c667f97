/*
 * vector.h -- double-precision small-vector math for the b200 path tracer host side.
 *
 * API-compatible with the reference's math header (type and function names, argument
 * order, evaluation order of every expression) so that unchanged reference-side callers
 * (`main.c`, `test.c`) compile against it.  Reference: /root/reference/vector.h:7-84.
 *
 * Differences on purpose:
 *   - every function is `static inline`, so a translation unit never depends on an
 *     out-of-line definition existing somewhere else (the reference relies on -O3
 *     inlining bare C99 `inline`s);
 *   - `mat4_mult` computes a real 4x4 product (the reference indexes with the wrong
 *     stride and is unused; quirk Q14).
 * All arithmetic is IEEE double, left-to-right, matching vector.h:16-61 so host-side
 * values (camera basis, scene setup) are bit-identical to the reference's.
 */
#ifndef RTB200_VECTOR_H
#define RTB200_VECTOR_H
#ifndef VECTOR_M
#define VECTOR_M /* keeps a later #include of the reference header from redefining */
#endif

#include <math.h>
#include <assert.h>

typedef double REAL;

typedef struct { REAL x, y; } vec2;
typedef struct { REAL x, y, z; } vec3;
typedef struct { REAL x, y, z, w; } vec4;
typedef REAL mat2[4];
typedef REAL mat3[9];
typedef REAL mat4[16];

#define MAT4_D (4)
#define MAT4_P (1)

/* ---- vec2 ---------------------------------------------------------------- */

static inline vec2 vec2_add(vec2 a, vec2 b)
{
  vec2 r = { a.x + b.x, a.y + b.y };
  return r;
}

static inline vec2 vec2_scalar_mult(vec2 v, REAL s)
{
  vec2 r = { v.x * s, v.y * s };
  return r;
}

/* ---- vec3: componentwise -------------------------------------------------- */

static inline vec3 vec3_add(vec3 a, vec3 b)
{
  vec3 r = { a.x + b.x, a.y + b.y, a.z + b.z };
  return r;
}

static inline vec3 vec3_sub(vec3 a, vec3 b)
{
  vec3 r = { a.x - b.x, a.y - b.y, a.z - b.z };
  return r;
}

static inline vec3 vec3_mult(vec3 a, vec3 b)
{
  vec3 r = { a.x * b.x, a.y * b.y, a.z * b.z };
  return r;
}

static inline vec3 vec3_scalar_mult(vec3 v, REAL s)
{
  vec3 r = { v.x * s, v.y * s, v.z * s };
  return r;
}

/* multiplies by the reciprocal, like vector.h:34-35 (not three divisions) */
static inline vec3 vec3_scalar_div(vec3 v, REAL s)
{
  return vec3_scalar_mult(v, 1.0 / s);
}

static inline int vec3_equal(vec3 a, vec3 b)
{
  return (a.x == b.x) && (a.y == b.y) && (a.z == b.z);
}

/* ---- vec3: products, norms ------------------------------------------------ */

static inline REAL vec3_dot(vec3 a, vec3 b)
{
  return a.x * b.x + a.y * b.y + a.z * b.z;
}

static inline vec3 vec3_cross(vec3 a, vec3 b)
{
  vec3 r;
  r.x = a.y * b.z - a.z * b.y;
  r.y = a.z * b.x - a.x * b.z;
  r.z = a.x * b.y - a.y * b.x;
  return r;
}

static inline REAL vec3_length(vec3 v)
{
  return sqrt(vec3_dot(v, v));
}

static inline vec3 vec3_normalize(vec3 v)
{
  REAL len = vec3_length(v);
  assert(len > 0);
  return vec3_scalar_mult(v, 1.0 / len);
}

/* ---- mat4 (row-major) ------------------------------------------------------ */

/* affine transform of a point: (A * [v,1]).xyz, vector.h:63-74 */
static inline vec3 mat4_vector_mult(mat4 A, vec3 v)
{
  REAL in[4] = { v.x, v.y, v.z, 1.0 };
  REAL out[4];
  for (unsigned row = 0; row < 4; row++)
  {
    REAL acc = 0;
    for (unsigned k = 0; k < 4; k++)
      acc += A[row * 4 + k] * in[k];
    out[row] = acc;
  }
  vec3 r = { out[0], out[1], out[2] };
  return r;
}

/* C = A * B */
static inline void mat4_mult(mat4 A, mat4 B, mat4 C)
{
  for (unsigned row = 0; row < 4; row++)
    for (unsigned col = 0; col < 4; col++)
    {
      REAL acc = 0;
      for (unsigned k = 0; k < 4; k++)
        acc += A[row * 4 + k] * B[k * 4 + col];
      C[row * 4 + col] = acc;
    }
}

#endif /* RTB200_VECTOR_H */
