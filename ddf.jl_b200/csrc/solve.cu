// solve.cu -- N2: Krylov solve on device vectors around any ddf_op (examples/poisson.jl:203-227).
//
// The reference solves A' x = b' with IterativeSolvers.bicgstabl_iterator! (IterativeSolvers
// 0.9.0, Manifest.toml:471-475; l = 2, Pl = Identity) when D >= 3 and with UMFPACK `\` when
// D <= 2.  This is BiCGStab(l) (Sleijpen & Fokkema 1993) in the arrangement IterativeSolvers
// uses -- l BiCG sub-steps that build the residual / search-direction Krylov blocks, then one
// minimal-residual step of degree l through the (l+1)x(l+1) Gram matrix -- on device vectors:
// every matrix-vector product is ddf_op_apply, every inner product the deterministic
// reduction of core.cu.  IterativeSolvers draws its shadow residual with rand(), so the
// reference's iterates are not reproducible; the shadow residual here is r0.  Parity is a
// statement about the converged solution (tests compare with a direct solve of the oracle's
// restatement of A', b').
#include <cmath>
#include <memory>
#include <vector>

#include "common.cuh"

int ddf_vec_gram(ddf_ctx* c, int nv, ddf_vec* const* vs, double* out);   // core.cu

namespace {
struct VecSet {
  std::vector<ddf_vec*> v;
  ~VecSet() {
    for (ddf_vec* p : v) ddf_vec_destroy(p);
  }
  int add(ddf_ctx* ctx, int64_t n, ddf_vec** out) {
    ddf_vec* p = nullptr;
    DDF_CHECK(ddf_vec_create(ctx, n, &p));
    v.push_back(p);
    *out = p;
    return 0;
  }
};

// solve the small dense system G y = rhs (n <= 8) by LU with partial pivoting, in place
int small_solve(int n, double* G, double* rhs) {
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++)
      if (std::fabs(G[r * n + c]) > std::fabs(G[piv * n + c])) piv = r;
    if (G[piv * n + c] == 0.0) return 1;
    if (piv != c) {
      for (int k = 0; k < n; k++) std::swap(G[c * n + k], G[piv * n + k]);
      std::swap(rhs[c], rhs[piv]);
    }
    for (int r = c + 1; r < n; r++) {
      const double f = G[r * n + c] / G[c * n + c];
      for (int k = c; k < n; k++) G[r * n + k] -= f * G[c * n + k];
      rhs[r] -= f * rhs[c];
    }
  }
  for (int r = n - 1; r >= 0; r--) {
    double s = rhs[r];
    for (int k = r + 1; k < n; k++) s -= G[r * n + k] * rhs[k];
    rhs[r] = s / G[r * n + r];
  }
  return 0;
}
}  // namespace

extern "C" int ddf_solve_bicgstabl(ddf_op* A, ddf_vec* b, ddf_vec* x, int l, double reltol, double abstol,
                                   int max_mv_products, int* mv_products, double* residual, int* converged) {
  DDF_TRY_BEGIN
  if (l < 1 || l > 7) DDF_FAIL("bicgstabl: 1 <= l <= 7, got %d", l);
  if (b->n != x->n) DDF_FAIL("bicgstabl: b has length %lld, x has %lld", (long long)b->n, (long long)x->n);
  if (b == x) DDF_FAIL("bicgstabl: b and x must not alias");
  ddf_ctx* ctx = b->ctx;
  const int64_t n = b->n;
  VecSet ws;
  std::vector<ddf_vec*> rs(l + 1), us(l + 2);
  ddf_vec* shadow = nullptr;
  for (auto& p : rs) DDF_CHECK(ws.add(ctx, n, &p));
  for (auto& p : us) DDF_CHECK(ws.add(ctx, n, &p));
  DDF_CHECK(ws.add(ctx, n, &shadow));
  int mv = 0;
  // r0 = b - A x
  DDF_CHECK(ddf_op_apply(A, x, rs[0]));
  mv++;
  DDF_CHECK(ddf_vec_axpby(1.0, b, -1.0, rs[0], rs[0]));
  DDF_CHECK(ddf_vec_copy(shadow, rs[0]));
  for (auto& p : us) DDF_CHECK(ddf_vec_fill(p, 0.0));
  double nrm = 0;
  DDF_CHECK(ddf_vec_norm(rs[0], DDF_NORM_2, &nrm));
  const double tol = std::fmax(reltol * nrm, abstol);
  double omega = 1.0, sigma = 1.0;
  bool ok = nrm <= tol;
  while (!ok && mv + 2 * l <= max_mv_products) {
    sigma = -omega * sigma;
    // ---- BiCG part
    for (int j = 0; j < l; j++) {
      double rho = 0;
      DDF_CHECK(ddf_vec_dot(shadow, rs[j], &rho));
      if (sigma == 0.0) DDF_FAIL("bicgstabl: breakdown (sigma = 0) after %d products", mv);
      const double beta = rho / sigma;
      for (int i = 0; i <= j; i++) DDF_CHECK(ddf_vec_axpby(1.0, rs[i], -beta, us[i], us[i]));
      DDF_CHECK(ddf_op_apply(A, us[j], us[j + 1]));
      mv++;
      DDF_CHECK(ddf_vec_dot(shadow, us[j + 1], &sigma));
      if (sigma == 0.0) DDF_FAIL("bicgstabl: breakdown (sigma = 0) after %d products", mv);
      const double alpha = rho / sigma;
      for (int i = 0; i <= j; i++) DDF_CHECK(ddf_vec_axpby(1.0, rs[i], -alpha, us[i + 1], rs[i]));
      DDF_CHECK(ddf_op_apply(A, rs[j], rs[j + 1]));
      mv++;
      DDF_CHECK(ddf_vec_axpby(1.0, x, alpha, us[0], x));
    }
    // ---- minimal-residual part: gamma = argmin | rs[0] - sum_k gamma_k rs[k] |,  k = 1..l
    double G[49], g[7], W[64];
    DDF_CHECK(ddf_vec_gram(ctx, l + 1, rs.data(), W));      // one pass, one synchronisation for all (l+1)(l+2)/2 products
    for (int r = 1; r <= l; r++) {
      for (int c = 1; c <= l; c++) G[(r - 1) * l + (c - 1)] = W[r * (l + 1) + c];
      g[r - 1] = W[r * (l + 1)];
    }
    if (small_solve(l, G, g)) DDF_FAIL("bicgstabl: singular minimal-residual system after %d products", mv);
    // us[0] -= sum g_k us[k];  x += sum g_k rs[k-1];  rs[0] -= sum g_k rs[k]: one pass each (x before rs[0] changes)
    double mg[7];
    for (int k = 0; k < l; k++) mg[k] = -g[k];
    DDF_CHECK(ddf_vec_lincomb(us[0], us[0], 1.0, l, mg, us.data() + 1));
    DDF_CHECK(ddf_vec_lincomb(x, x, 1.0, l, g, rs.data()));
    DDF_CHECK(ddf_vec_lincomb(rs[0], rs[0], 1.0, l, mg, rs.data() + 1));
    omega = g[l - 1];
    DDF_CHECK(ddf_vec_norm(rs[0], DDF_NORM_2, &nrm));
    ok = nrm <= tol;
    if (!std::isfinite(nrm)) DDF_FAIL("bicgstabl: residual is not finite after %d products", mv);
  }
  if (mv_products) *mv_products = mv;
  if (residual) *residual = nrm;
  if (converged) *converged = ok ? 1 : 0;
  return 0;
  DDF_TRY_END
}
