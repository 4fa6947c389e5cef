/* ddf_b200.h -- C ABI of libddf_b200.so: a B200 (sm_100a) CUDA implementation of
 * DDF.jl's discrete-exterior-calculus operator hot path.
 *
 * The reference (eschnett/DDF.jl, pure Julia) has no FFI seam of its own; the seam
 * is Julia multiple dispatch on the container inside `Op.values` (src/Ops.jl:18-34)
 * and `Fun.values` (src/Funs.jl:16-31), with examples/wave-gpu.jl:22-34,121-126 as
 * the precedent (swap the array type, overload `mul!`).  Every entry point below
 * names the reference interface it replaces (file:line under /root/reference).
 * The Julia binding a maintainer would add is shown in INTEGRATION.md; the Python
 * ctypes binding used by the tests is ddf.jl_b200/_lib.py.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; the message is
 *    available from ddf_last_error() (thread-local).  Nothing throws or exits.
 *  - handles are opaque and library-owned; *_destroy releases device memory.
 *  - host pointers are borrowed for the duration of the call only.
 *  - indices cross the ABI as the reference stores them: int64, 1-based, CSC
 *    (colptr[n+1], rowval[nnz], nzval[nnz]); on the device they are 0-based int32.
 *  - one ctx == one GPU == one process rank (torch.distributed-style launch);
 *    multi-GPU meshes are slabs joined by NCCL halos (ddf_comm_*, ddf_halo_*).
 *  - there is NO CPU fallback: without an sm_100 device ddf_ctx_create fails.
 */
#ifndef DDF_B200_H
#define DDF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

typedef struct ddf_ctx ddf_ctx;
typedef struct ddf_mesh ddf_mesh;
typedef struct ddf_vec ddf_vec;
typedef struct ddf_op ddf_op;

/* PrimalDual (src/Manifolds.jl: `@enum PrimalDual Pr Dl`, `!Pr == Dl`). */
enum { DDF_PR = 1, DDF_DL = 0 };
/* DualKind (src/Manifolds.jl: BarycentricDuals / CircumcentricDuals). Only the
 * circumcentric kind is implemented (the barycentric code path of the reference
 * is dead: called with the wrong arity at Manifolds.jl:649-652). */
enum { DDF_DUAL_BARYCENTRIC = 0, DDF_DUAL_CIRCUMCENTRIC = 1 };
/* norms for ddf_vec_norm (LinearAlgebra.norm(x, p); examples/wave.jl:127,162) */
enum { DDF_NORM_1 = 1, DDF_NORM_2 = 2, DDF_NORM_INF = 0 };

const char* ddf_last_error(void);
int ddf_version(void);

/* ------------------------------------------------------------------ context */
int ddf_ctx_create(int device, ddf_ctx** out);
int ddf_ctx_destroy(ddf_ctx* ctx);
int ddf_sync(ddf_ctx* ctx);
/* bytes currently held by the library on this ctx's device */
int ddf_ctx_bytes_in_use(ddf_ctx* ctx, int64_t* out);
/* Device memory is drawn from the device's stream-ordered pool and stays cached there when tables are dropped;
 * ddf_ctx_trim synchronises and returns the cached blocks to the driver. */
int ddf_ctx_trim(ddf_ctx* ctx);
/* CUDA-event timing on the library stream (bench.py's roofline timer):
 * tic records an event, toc records another, synchronises and returns ms. */
int ddf_timer_tic(ddf_ctx* ctx);
int ddf_timer_toc(ddf_ctx* ctx, double* ms);
/* numbered CUDA events on the library stream (0 <= slot < 4096): record without
 * synchronising, read the elapsed time between two recorded slots later */
int ddf_event_record(ddf_ctx* ctx, int slot);
int ddf_event_elapsed(ddf_ctx* ctx, int slot_a, int slot_b, double* ms);
/* number of kernels the library has launched on this ctx since creation */
int ddf_ctx_launch_count(ddf_ctx* ctx, int64_t* out);
/* kernel-variant selector for profiling / A-B timing ("fixed", "wave", "asm"; "packbits" narrows the offset field of
 * the prefix-packed d_k rows -- and switches the bit-packed rows off -- so that tests can reach the explicit-table
 * fallback; "prefetch" = how many blocks ahead a block asks L2 for the index words of a later block, default 592,
 * 0 = off); every variant returns bit-identical results; for "fixed", "wave" and "asm" 0 selects the default */
int ddf_set_tuning(const char* key, int value);
/* the same selector for ONE context (ddf_set_tuning sets the defaults of later contexts and every live one);
 * tuning state lives in the context, so distinct contexts may be driven from distinct threads */
int ddf_ctx_set_tuning(ddf_ctx* ctx, const char* key, int value);
/* overwrite a >L2-sized scratch buffer (bench hygiene between timed launches) */
int ddf_flush_l2(ddf_ctx* ctx);

/* ------------------------------------------------------------------- meshes */
/* Manifold outer constructor input (src/Manifolds.jl:319-330): the top-simplex
 * table `simplicesD::SparseOp{0,D,One}` as CSC (nvertices x nsimplices, D+1 rows
 * per column, index-only), vertex coordinates (C x V column-major == V rows of C
 * doubles) and vertex weights. */
int ddf_mesh_create(ddf_ctx* ctx, int D, int C, int64_t nvertices,
                    int64_t nsimplices, const int64_t* colptr,
                    const int64_t* rowval, const double* coords,
                    const double* weights, ddf_mesh** out);
/* large_hypercube_manifold (src/ManifoldConstructors.jl:196-284) generated on the
 * device, same simplex order / vertex numbering / coords / level weights.
 * `hdiv` > 0 generalises the reference's equal-nelts assert (:262): coordinates
 * are i/hdiv and weights levelweight/hdiv^2 (slab meshes of BASELINE config 5);
 * hdiv == 0 reproduces the reference (requires all nelts equal).
 * `z0`: offset (in cells) of this slab along the last axis -- added to the last
 * coordinate before the division, so slab meshes share global coordinates. */
int ddf_mesh_create_hypercube(ddf_ctx* ctx, int D, const int64_t* nelts,
                              int64_t hdiv, int64_t z0, ddf_mesh** out);
int ddf_mesh_destroy(ddf_mesh* mesh);
/* calc_faces_boundaries for R = D..1 (src/Manifolds.jl:361-368, :717-768):
 * lexicographic face ranking + signed boundary matrices, bit-exact. */
int ddf_mesh_assemble(ddf_mesh* mesh);
/* nsimplices(mfd, R) */
int ddf_mesh_nsimplices(ddf_mesh* mesh, int R, int64_t* out);
int ddf_mesh_dims(ddf_mesh* mesh, int* D, int* C);
/* get_simplices(mfd, R).op (src/Manifolds.jl:434-441): colptr[n_R+1], rowval[(R+1) n_R] */
int ddf_mesh_get_simplices_csc(ddf_mesh* mesh, int R, int64_t* colptr, int64_t* rowval);
/* get_boundaries(mfd, R).op (src/Manifolds.jl:443-450): n_{R-1} x n_R, Int8 values */
int ddf_mesh_get_boundary_csc(ddf_mesh* mesh, int R, int64_t* colptr,
                              int64_t* rowval, int8_t* nzval);
/* get_lookup(mfd, Ri, Rj).op (src/Manifolds.jl:455-498): the n_Ri x n_Rj index-only (`One`) incidence pattern --
 * Ri == 0: the vertex tables, Ri == Rj: identity, Ri < Rj: sub-simplices (|boundaries| and its chains), Ri > Rj: the
 * transpose -- as 1-based Int64 CSC with rows ascending.  Call with colptr == NULL to query nnz. */
int ddf_mesh_get_lookup_csc(ddf_mesh* mesh, int Ri, int Rj, int64_t* nnz, int64_t* colptr, int64_t* rowval);
/* refine_coords(get_lookup(mfd,0,1), get_coords(mfd)) (src/Meshing.jl:127-146; ManifoldConstructors.jl:523): the old
 * vertices followed by the edge midpoints in edge-id order, (n_0 + n_1) rows of C doubles */
int ddf_mesh_refine_coords(ddf_mesh* mesh, double* out);
/* calc_volumes / weighted calc_dualcoords / circumcentric calc_dualvolumes
 * (src/Manifolds.jl:834-855, :896-913, :1160-1223) for all R. */
int ddf_mesh_geometry(ddf_mesh* mesh, int dualkind, int use_weighted_duals);
int ddf_mesh_get_coords(ddf_mesh* mesh, double* out /* V*C */);
int ddf_mesh_get_weights(ddf_mesh* mesh, double* out /* V */);
int ddf_mesh_get_volumes(ddf_mesh* mesh, int R, double* out);
int ddf_mesh_get_dualvolumes(ddf_mesh* mesh, int R, double* out);
int ddf_mesh_get_dualcoords(ddf_mesh* mesh, int R, double* out /* n_R*C */);
/* get_isboundary(mfd, R) as a 0/1 byte mask (src/Manifolds.jl:510-550) */
int ddf_mesh_get_isboundary(ddf_mesh* mesh, int R, uint8_t* out);

/* ------------------------------------------------------------------ cochains */
/* Fun{D,P,R}.values (src/Funs.jl:16-31): a device vector of n doubles. */
int ddf_vec_create(ddf_ctx* ctx, int64_t n, ddf_vec** out);
int ddf_vec_destroy(ddf_vec* v);
int ddf_vec_length(ddf_vec* v, int64_t* out);
int ddf_vec_upload(ddf_vec* v, const double* host);
int ddf_vec_download(ddf_vec* v, double* host);
/* Asynchronous variants on dedicated copy streams (one per PCIe direction, overlapped with
 * compute and with each other).  `host` must be page-locked and must stay valid until
 * ddf_sync(ctx); ordering against kernels that touch `v` is tracked by the library. */
int ddf_vec_upload_async(ddf_vec* v, const double* host);
int ddf_vec_download_async(ddf_vec* v, double* host);
/* out = base + dt*(coef[0] ks[0] + ... + coef[nk-1] ks[nk-1]), 1 <= nk <= 8, products and sums separately
 * rounded left to right: the stage update of an explicit Runge-Kutta step, i.e. the broadcast
 * `@.. tmp = uprev + dt*(a31*k1 + a32*k2)` OrdinaryDiffEq issues on the state vector (examples/wave.jl:134).
 * `out` may alias `base`, not a term. */
int ddf_vec_lincomb(ddf_vec* out, ddf_vec* base, double dt, int nk, const double* coef, ddf_vec** ks);
int ddf_vec_fill(ddf_vec* v, double a);
int ddf_vec_copy(ddf_vec* dst, ddf_vec* src);
/* sample(Fun{D,Pr,0}, f, mfd) for the fields of the examples, evaluated at the
 * vertex coordinates on the device (src/Continuum.jl:124-158, R = 0):
 * kind 0: prod_d sinpi(x_d)  (examples/wave.jl:45 at t=0)
 * kind 1: sin(x1) cos(x2)    (README.md:63-65)
 * kind 2: poisson.jl Gaussian rho (examples/poisson.jl:85-92) */
int ddf_vec_sample(ddf_mesh* mesh, int kind, ddf_vec* out);
/* uniform [0,1) pseudo-random fill (Funs.jl:80-84 `rand`), counter-based, seeded */
int ddf_vec_rand(ddf_vec* v, uint64_t seed);
/* z = a*x + b*y  (Funs.jl:156-188: +, -, scalar *, /); z may alias x or y; y may be NULL when b == 0 */
int ddf_vec_axpby(double a, ddf_vec* x, double b, ddf_vec* y, ddf_vec* z);
/* z = x / a (Funs.jl:186-188) */
int ddf_vec_div(ddf_vec* x, double a, ddf_vec* z);
/* z = x .* y */
int ddf_vec_mul(ddf_vec* x, ddf_vec* y, ddf_vec* z);
int ddf_vec_norm(ddf_vec* v, int kind, double* out);
int ddf_vec_dot(ddf_vec* x, ddf_vec* y, double* out);
int ddf_vec_sum(ddf_vec* x, double* out);
/* dot(f, g) of src/ManifoldOps.jl:154-164: sum f_i (1/vol_D[i]) g_i for Fun{D,Pr,D}, sum f_i (1/dualvol_0[i]) g_i
 * for Fun{D,Dl,0}; any other (P, R) is an error like the reference's missing method.  One fused pass. */
int ddf_vec_wdot(ddf_mesh* mesh, int P, int R, ddf_vec* f, ddf_vec* g, double* out);

/* ----------------------------------------------------------------- operators */
/* Each returns a new op handle (src/ManifoldOps.jl).  P is DDF_PR / DDF_DL. */
int ddf_op_boundary(ddf_mesh* mesh, int P, int R, ddf_op** out);   /* ManifoldOps.jl:17-26  */
int ddf_op_deriv(ddf_mesh* mesh, int P, int R, ddf_op** out);      /* ManifoldOps.jl:40-47  */
int ddf_op_hodge(ddf_mesh* mesh, int P, int R, ddf_op** out);      /* ManifoldOps.jl:58-68,84-91 */
int ddf_op_invhodge(ddf_mesh* mesh, int P, int R, ddf_op** out);   /* ManifoldOps.jl:71-82,93-100 */
int ddf_op_coderiv(ddf_mesh* mesh, int P, int R, ddf_op** out);    /* ManifoldOps.jl:110-121 */
int ddf_op_laplace(ddf_mesh* mesh, int P, int R, ddf_op** out);    /* ManifoldOps.jl:128-144 */
int ddf_op_isboundary(ddf_mesh* mesh, int P, int R, ddf_op** out); /* ManifoldOps.jl:31-35  */
int ddf_op_identity(ddf_mesh* mesh, int R, double a, ddf_op** out);/* Ops.jl:216-220 `one`, scaled */
int ddf_op_zero(ddf_mesh* mesh, int R1, int R2, ddf_op** out);     /* Ops.jl:152-157 `zero` */
/* arbitrary Op from the reference's storage: SparseMatrixCSC{Float64,Int} */
int ddf_op_from_csc(ddf_ctx* ctx, int64_t nrows, int64_t ncols, const int64_t* colptr,
                    const int64_t* rowval, const double* nzval, ddf_op** out);
/* Diagonal{Float64} */
int ddf_op_from_diag(ddf_ctx* ctx, int64_t n, const double* diag, ddf_op** out);
int ddf_op_destroy(ddf_op* op);
int ddf_op_shape(ddf_op* op, int64_t* nrows, int64_t* ncols);
/* Op*Op (Ops.jl:228-232): lazy product C = A*B (applied right to left, diagonal
 * factors fused into the neighbouring sparse kernel) */
int ddf_op_compose(ddf_op* A, ddf_op* B, ddf_op** out);
/* A + b*B (Ops.jl:186-196) */
int ddf_op_add(ddf_op* A, double b, ddf_op* B, ddf_op** out);
/* a*A (Ops.jl:198-212) */
int ddf_op_scale(double a, ddf_op* A, ddf_op** out);
/* A / a (Ops.jl:210-212): the assembled values divided by a (one rounding each) + the constructor's chop */
int ddf_op_div(ddf_op* A, double a, ddf_op** out);
/* adjoint(A) (Ops.jl:258-260) */
int ddf_op_adjoint(ddf_op* A, ddf_op** out);
/* Op*Fun (Ops.jl:273-276): y = A x */
int ddf_op_apply(ddf_op* A, ddf_vec* x, ddf_vec* y);
/* Storage of the rows of d_k = deriv(Pr,k) the apply kernels use on this mesh (chosen on first use):
 * 2 / 3 = bit-packed rows: ONE 4- / 8-byte word of index differences per row (index 0 against the smallest index 0
 * of its 32-row group, every later index against its predecessor; field widths chosen per table from the largest
 * differences it holds) + a base per 32 rows; 1 = prefix-packed (the first face of every row is its prefix, whose
 * ids run non-decreasing down the rows: W-1 words per row + a base per 32 rows); 0 = explicit index table (tables
 * whose differences do not fit, e.g. top simplices given in a scattered order).  All give identical bits (same
 * gathers, same order). */
int ddf_mesh_deriv_format(ddf_mesh* m, int k, int* format);
/* The operator in the reference's own representation: every Op holds its assembled sparse matrix (Ops.jl:25),
 * rounded and chopped after each product / sum (Ops.jl:36-58,186-232), and `A * f` is a sparse mat-vec on it
 * (Ops.jl:273-276).  Returns a new operator whose apply is one row kernel over those values; the iterates of a
 * Krylov solve on it then see the matrix the reference's solver sees (examples/poisson.jl:203-227). */
int ddf_op_assemble(ddf_op* A, ddf_op** out);
/* Assembled export in the reference's storage, with the Op constructor's
 * dnz/chop applied (Ops.jl:36-58).  Call with NULL arrays to query nnz. */
int ddf_op_to_csc(ddf_op* A, int64_t* nnz, int64_t* colptr, int64_t* rowval, double* nzval);

/* ---------------------------------------------------------------------- wave */
/* examples/wave.jl:87-100: dU = F U, i.e. du = (E-B) udot, dudot = (E-B)(L u). */
int ddf_wave_rhs(ddf_mesh* mesh, ddf_vec* u, ddf_vec* udot, ddf_vec* du, ddf_vec* dudot);
/* examples/wave.jl:106,118: dt = minimum(vol_1)/2 */
int ddf_wave_dt(ddf_mesh* mesh, double* dt);
/* north_star's fused leapfrog on the same operators (not in the reference):
 *   v += dt (1-b) .* (L u);  u += dt (1-b) .* v     repeated nsteps times,
 * one kernel per step, L applied matrix-free from the vertex adjacency. */
int ddf_wave_leapfrog(ddf_mesh* mesh, ddf_vec* u, ddf_vec* v, double dt, int nsteps);
/* `nsteps` fixed steps of Tsit5 on dU = F U, U = [u; udot] -- what examples/wave.jl:134 runs
 * (`solve(prob, Tsit5(); adaptive=false, dt=dt)`); in place.  Same roundings as the unfused perform_step! (stage
 * broadcasts with separately rounded products and sums, left to right); on meshes whose rows are staged a step is three
 * passes over L, two stages each (a stage input of u does not depend on the previous stage's derivative of udot). */
int ddf_wave_tsit5(ddf_mesh* mesh, ddf_vec* u, ddf_vec* udot, double dt, int nsteps);
/* Fused Poisson relaxation sweep (weighted Jacobi on L u = rho with the boundary values of u held fixed --
 * the Dirichlet problem of examples/poisson.jl:75-81): u <- u + theta * ((1-b) .* ((rho - L u) ./ diag(L))),
 * `nsweeps` times, one pass over the rows per sweep (same kernels as the wave step). */
int ddf_poisson_jacobi(ddf_mesh* mesh, ddf_vec* u, ddf_vec* rho, double theta, int nsweeps);
/* one leapfrog step timed alone is exposed through ddf_timer_*; the algorithmic
 * byte count of a step (E*16 + V*41) is returned here for reporting */
int ddf_wave_step_bytes(ddf_mesh* mesh, int64_t* algorithmic, int64_t* format_bytes);
/* storage the row kernels use for L on this mesh: 0 = row-CSR, 1 = explicit-index SELL-32-256,
 * 2 = staged rows (TMA-staged gather of u, 16-bit slots, jagged slices; csrc/stage.cuh) */
int ddf_wave_format(ddf_mesh* mesh, int* format);

/* ------------------------------------------------------------ Krylov solve (N2) */
/* BiCGStab(l) on device vectors around any operator: the solver examples/poisson.jl:203-227 runs for D >= 3
 * (IterativeSolvers.bicgstabl_iterator!(x, A', b'; Pl = Identity): l = 2, reltol = sqrt(eps), abstol = 0,
 * max_mv_products = size(A,2)).  x holds the initial guess on entry and the solution on return; the loop stops
 * when |r| <= max(reltol*|r0|, abstol) or another sweep would exceed max_mv_products.  Outputs may be NULL. */
int ddf_solve_bicgstabl(ddf_op* A, ddf_vec* b, ddf_vec* x, int l, double reltol, double abstol, int max_mv_products,
                        int* mv_products, double* residual, int* converged);

/* --------------------------------------------------------- multi-GPU (NCCL) */
/* ncclGetUniqueId: 128 bytes, created on rank 0 and broadcast by the host side */
int ddf_comm_unique_id(void* id128);
int ddf_comm_init(ddf_ctx* ctx, int nranks, int rank, const void* id128);
int ddf_comm_destroy(ddf_ctx* ctx);
/* Slab decomposition of the distributed wave step: this rank's
 * mesh is a slab along the slowest axis with `plane` vertices per plane; its first
 * `ghost_lo` and last `ghost_hi` vertex planes are ghosts owned by rank-1 / rank+1
 * (0 at the domain ends).  Owned rows are the planes in between. */
int ddf_mesh_set_halo(ddf_mesh* mesh, int64_t plane, int64_t ghost_lo, int64_t ghost_hi);
/* Partition of a mesh of ANY origin (uploaded, refined, Delaunay) by contiguous ranges of the global vertex numbering
 * (north_star: "large refined meshes are partitioned by simplex ranges"; the reference has no parallel code).  The
 * local mesh is the closed star of the owned vertices, renumbered monotonically (host side: ddf.dist.partition_mesh),
 * so owned local vertices are the range [own_lo, own_hi) and every owned row of the geometry / star / L tables equals
 * the undivided mesh's bit for bit.  peers[k] (k < npeers) receives the owned values send_idx[send_ptr[k] ..
 * send_ptr[k+1]) and fills the ghosts recv_idx[recv_ptr[k] .. recv_ptr[k+1]); 0-based local vertex ids, both sides of
 * a pair in ascending global id.  After this call ddf_wave_leapfrog_dist, ddf_poisson_jacobi_dist, ddf_wave_tsit5,
 * ddf_halo_exchange and ddf_mesh_owned_rows work as on a slab (one grouped ncclSend/ncclRecv per exchange). */
int ddf_mesh_set_halo_lists(ddf_mesh* mesh, int64_t own_lo, int64_t own_hi, int npeers, const int32_t* peers,
                            const int64_t* send_ptr, const int64_t* send_idx, const int64_t* recv_ptr,
                            const int64_t* recv_idx);
/* The rank-R simplices this rank owns (smallest vertex in an owned plane) as ONE row range [lo, hi) of its local
 * tables (lexicographic numbering makes them contiguous); a mesh without a halo owns everything. */
int ddf_mesh_owned_rows(ddf_mesh* mesh, int R, int64_t* lo, int64_t* hi);
/* A copy of a leaf operator (deriv(Pr,k), hodge / invhodge / isboundary) whose apply writes rows [row0, row1) only:
 * a slab rank applies d_k / star_k to the rows it owns and leaves the ghost rows to their owner. */
int ddf_op_restrict_rows(ddf_op* A, int64_t row0, int64_t row1, ddf_op** out);
/* fused leapfrog on the owned rows + one-plane halo exchange of u per step with
 * ncclSend/ncclRecv to rank-1 / rank+1, overlapped with the interior rows. */
int ddf_wave_leapfrog_dist(ddf_mesh* mesh, ddf_vec* u, ddf_vec* v, double dt, int nsteps);
/* ddf_poisson_jacobi on a slab mesh: the same halo protocols as the distributed wave step */
int ddf_poisson_jacobi_dist(ddf_mesh* mesh, ddf_vec* u, ddf_vec* rho, double theta, int nsweeps);
/* how the last ddf_wave_leapfrog_dist moved its ghost planes: 0 = not run yet, 1 = fused peer-memory stores
 * (the row kernel writes the interface planes straight into the neighbours' ghost planes over NVLink; neighbours
 * are kept within one step by release/acquire step flags in peer memory), 2 = ncclSend/ncclRecv (fallback when
 * CUDA IPC mappings are unavailable on any rank, or forced by ddf_set_tuning("wave", 1024)) */
int ddf_halo_mode(ddf_mesh* mesh, int* mode);
/* halo exchange of a vertex cochain alone (fills the ghost planes next to the owned rows) */
int ddf_halo_exchange(ddf_mesh* mesh, ddf_vec* u);
/* ghost exchange of a cochain of ANY rank R on a slab mesh: every local R-simplex whose smallest vertex lies in a
 * ghost plane receives the value held by the rank that owns it (ownership = owner of the smallest vertex).  The
 * pack / unpack lists are built on the first call per (mesh, R); neighbours find matching entries by position
 * because local simplex ids are lexicographic ranks and neighbouring slabs differ by a constant vertex shift. */
int ddf_halo_exchange_k(ddf_mesh* mesh, ddf_vec* x, int R);
/* sum / max all-reduce of a host scalar across ranks (norms, CFL minimum) */
int ddf_comm_allreduce(ddf_ctx* ctx, double* value, int op /* 0 sum, 1 max, 2 min */);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DDF_B200_H */
