/* abi_client.c -- a plain C client of libddf_b200.so: what a `ccall` from Julia does, without Python in between.
 * Built by tests/test_gpu_parity.py::test_plain_c_client with gcc against include/ddf_b200.h and run on the GPU.
 *
 * It uploads the two-triangle square through the reference's storage (1-based Int64 CSC of a `One` matrix,
 * src/Manifolds.jl:319-330), assembles the faces on the device and checks, against values worked out by hand
 * from the reference's rules (Manifolds.jl:717-768: face ids = lexicographic ranks, sign of the face omitting
 * the n-th smallest vertex = (-1)^(n-1)):
 *   counts 4 / 5 / 2; edges (1,2) (1,3) (2,3) (2,4) (3,4); boundary_2 columns [+,-,+] on edges {1,2,3} and {3,4,5};
 *   d0 x = edge differences; d1 d0 x = 0; hodge_2 = 1/area; error status + message for a shape mismatch. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "ddf_b200.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    if ((call) != 0) {                                                     \
      fprintf(stderr, "FAIL %s: %s\n", #call, ddf_last_error());           \
      return 1;                                                            \
    }                                                                      \
  } while (0)
#define EXPECT(cond)                                                       \
  do {                                                                     \
    if (!(cond)) {                                                         \
      fprintf(stderr, "EXPECT failed at line %d: %s\n", __LINE__, #cond);  \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(void) {
  ddf_ctx* ctx = NULL;
  CHECK(ddf_ctx_create(0, &ctx));
  /* unit square, vertices 1..4 = (0,0) (1,0) (0,1) (1,1); triangles {1,2,3}, {2,3,4} */
  const int64_t colptr[3] = {1, 4, 7};
  const int64_t rowval[6] = {1, 2, 3, 2, 3, 4};
  const double coords[8] = {0, 0, 1, 0, 0, 1, 1, 1};
  const double weights[4] = {0, 0, 0, 0};
  ddf_mesh* m = NULL;
  CHECK(ddf_mesh_create(ctx, 2, 2, 4, 2, colptr, rowval, coords, weights, &m));
  CHECK(ddf_mesh_assemble(m));
  int64_t n0, n1, n2;
  CHECK(ddf_mesh_nsimplices(m, 0, &n0));
  CHECK(ddf_mesh_nsimplices(m, 1, &n1));
  CHECK(ddf_mesh_nsimplices(m, 2, &n2));
  EXPECT(n0 == 4 && n1 == 5 && n2 == 2);
  int64_t ecp[6], erv[10];
  CHECK(ddf_mesh_get_simplices_csc(m, 1, ecp, erv));
  const int64_t want_e[10] = {1, 2, 1, 3, 2, 3, 2, 4, 3, 4};
  EXPECT(memcmp(erv, want_e, sizeof(want_e)) == 0);
  int64_t bcp[3], brv[6];
  int8_t bnz[6];
  CHECK(ddf_mesh_get_boundary_csc(m, 2, bcp, brv, bnz));
  const int64_t want_b[6] = {1, 2, 3, 3, 4, 5};
  const int8_t want_s[6] = {1, -1, 1, 1, -1, 1};
  EXPECT(bcp[0] == 1 && bcp[1] == 4 && bcp[2] == 7);
  EXPECT(memcmp(brv, want_b, sizeof(want_b)) == 0 && memcmp(bnz, want_s, sizeof(want_s)) == 0);

  ddf_op *d0 = NULL, *d1 = NULL, *h2 = NULL;
  CHECK(ddf_op_deriv(m, DDF_PR, 0, &d0));
  CHECK(ddf_op_deriv(m, DDF_PR, 1, &d1));
  ddf_vec *x = NULL, *y = NULL, *z = NULL;
  CHECK(ddf_vec_create(ctx, n0, &x));
  CHECK(ddf_vec_create(ctx, n1, &y));
  CHECK(ddf_vec_create(ctx, n2, &z));
  const double xv[4] = {1.0, 4.0, 9.0, 16.0};
  CHECK(ddf_vec_upload(x, xv));
  CHECK(ddf_op_apply(d0, x, y));
  double yv[5];
  CHECK(ddf_vec_download(y, yv));
  const double want_y[5] = {3.0, 8.0, 5.0, 12.0, 7.0};       /* x_b - x_a per edge (a,b) */
  EXPECT(memcmp(yv, want_y, sizeof(want_y)) == 0);
  CHECK(ddf_op_apply(d1, y, z));
  double zv[2];
  CHECK(ddf_vec_download(z, zv));
  EXPECT(zv[0] == 0.0 && zv[1] == 0.0);                       /* dd = 0, exact */

  CHECK(ddf_mesh_geometry(m, DDF_DUAL_CIRCUMCENTRIC, 1));
  double vol2[2];
  CHECK(ddf_mesh_get_volumes(m, 2, vol2));
  EXPECT(fabs(vol2[0] - 0.5) < 1e-15 && fabs(vol2[1] - 0.5) < 1e-15);
  CHECK(ddf_op_hodge(m, DDF_PR, 2, &h2));
  const double ones[2] = {1.0, 1.0};
  ddf_vec* w = NULL;
  CHECK(ddf_vec_create(ctx, n2, &w));
  CHECK(ddf_vec_upload(z, ones));
  CHECK(ddf_op_apply(h2, z, w));
  CHECK(ddf_vec_download(w, zv));
  EXPECT(fabs(zv[0] - 2.0) < 1e-14 && fabs(zv[1] - 2.0) < 1e-14);   /* dualvol_2 / vol_2 = 1 / 0.5 */

  /* errors come back as a status + message, never as an exception or exit */
  EXPECT(ddf_op_apply(d0, y, x) != 0);
  EXPECT(strstr(ddf_last_error(), "columns") != NULL || strstr(ddf_last_error(), "length") != NULL);
  ddf_op* bad = NULL;
  EXPECT(ddf_op_deriv(m, DDF_PR, 2, &bad) != 0 && bad == NULL);

  ddf_vec_destroy(w); ddf_vec_destroy(z); ddf_vec_destroy(y); ddf_vec_destroy(x);
  ddf_op_destroy(h2); ddf_op_destroy(d1); ddf_op_destroy(d0);
  ddf_mesh_destroy(m);
  ddf_ctx_destroy(ctx);
  printf("abi_client ok\n");
  return 0;
}
