// mesh.cuh -- device-resident manifold (src/Manifolds.jl:37-101 restated as flat
// int32 tables).  Layout in HBM (all 0-based):
//   simp[R]  : n_R x (R+1) int32, row-major, vertices ascending   (_simplices[R]; index-only, ZeroOrOne.jl:43-68)
//   bnd[R]   : n_R x (R+1) int32, row-major, face ids ascending   (_boundaries[R] columns; the +-1 values are
//              implicit: entry p (0-based) carries sign (-1)^(R+p), constant for every column)
//              bnd[1] aliases simp[1] (the faces of an edge are its vertices)
//   brow_ptr[R], bcol[R] : CSR of d_R by rows (face -> parents ascending), parent id | sign bit 31
//   vol[R], dualvol[R], cc[R] : geometry (Float64)
#pragma once
#include "common.cuh"

#define DDF_SELL_PAD 0xFFFFFFFFu
#define DDF_SELL_WIN 256
// SELL-32-256 storage (built by sell.cuh)
struct SellBufs {
  DevBuf<uint32_t> ptr;     // entry offset per slice of 32 sorted rows (nwindows*8 + 1)
  DevBuf<uint8_t> perm8;    // sorted position -> row offset inside its 256-row window
  DevBuf<uint32_t> col;     // explicit indices (released when compressed or stored as 16-bit offsets)
  DevBuf<uint16_t> col16;   // column - row as int16 (0x8000 = padding) when every offset fits: 10 B per entry
  bool idx16 = false;
  DevBuf<double> val;
  // compressed index stream: per slice a bit mask of uniform-offset columns and a word offset
  bool compressed = false;
  DevBuf<uint32_t> umask, iptr, cidx;
  int64_t padded = 0;       // stored entries incl. padding
  int64_t index_words = 0;  // 32-bit words of index data actually stored
  void release() {
    ptr.release(); perm8.release(); col.release(); col16.release(); val.release(); umask.release(); iptr.release();
    cidx.release();
    padded = 0; index_words = 0; compressed = false; idx16 = false;
  }
};

// staged rows (stage.cuh): jagged slices, 16-bit slots into a per-window staging buffer of u
struct StageBufs {
  DevBuf<uint32_t> ptr;     // entry offset per slice of 32 sorted rows (nwindows*8 + 1), no padding
  DevBuf<uint8_t> perm8;    // sorted position -> row offset inside its 256-row window
  DevBuf<uint8_t> len8;     // row length by sorted position
  DevBuf<uint32_t> tab;     // per window: nruns, nslots, (start_col, slot_off) per run
  // per window ONE contiguous record: [values f64 x nent][slots u16 x nent], nent = the window's stored entries
  // (a multiple of 8), at byte offset 10 * (first entry): one TMA bulk copy per window
  DevBuf<unsigned char> packed;
  // per window 816 B: {u32 nent, pad to 16, u32 slice offset x 8} | perm8[256] | len8[256] | boundary mask by row offset [256]
  DevBuf<unsigned char> meta;
  uint32_t max_slots = 0;   // largest staging buffer of any window (Float64 elements)
  uint32_t max_entries = 0; // most stored entries of any group of K windows (incl. their <= 7 padding entries each)
  int K = 1;                // consecutive windows that share one run table / one pipeline stage
  int64_t ngroups = 0;
  int64_t entries = 0;      // stored entries incl. the per-window padding
  int64_t nnz = 0;
  bool ok = false;
  void release() {
    ptr.release(); perm8.release(); len8.release(); tab.release(); packed.release(); meta.release();
    max_slots = 0; max_entries = 0; K = 1; ngroups = 0; entries = 0; nnz = 0; ok = false;
  }
};

// prefix-packed fixed-width rows of d_k (ops.cu, k_pack_fixed): W-1 words per row + one base per 32 rows
struct PackedBufs {
  DevBuf<uint32_t> words;   // row r: word 0 = (idx0 - base[r/32]) << 24 | (idx1 - idx0), words 1.. = idx2.. explicit
  DevBuf<uint32_t> base;    // idx0 of the first row of every group of 32 rows
  int state = 0;            // 0 = not built, 1 = in use, 2 = does not pack (explicit table)
  void release() { words.release(); base.release(); state = 0; }
};

// bit-packed fixed-width rows of d_k (ops.cu, k_bits_*): ONE 4- or 8-byte word per row + one base per 32 rows.
// Field widths are chosen per table from the largest differences it holds (index 0 against the smallest index 0 of
// its 32-row group, every later index against its predecessor in the row: face ids ascend inside a row).
struct BitRows {
  DevBuf<uint32_t> words;   // nrows * wbytes / 4 words
  DevBuf<uint32_t> base;    // smallest idx0 of every group of 32 rows
  int wbytes = 0;           // 4 or 8
  int b[4] = {0, 0, 0, 0};  // bits of idx0 - base, idx1 - idx0, idx2 - idx1, idx3 - idx2
  int state = 0;            // 0 = not built, 1 = in use, 2 = does not fit (prefix-packed rows / explicit table)
  void release() { words.release(); base.release(); wbytes = 0; state = 0; }
};

struct ddf_mesh {
  ddf_ctx* ctx = nullptr;
  int D = 0, C = 0;
  int64_t n[DDF_MAXD + 1] = {0};         // nsimplices per rank R
  bool assembled = false;
  bool has_geometry = false;
  bool weighted = true;

  DevBuf<int32_t> simp[DDF_MAXD + 1];    // R >= 1 (R = 0 is the identity table)
  DevBuf<int32_t> bnd[DDF_MAXD + 1];     // R >= 2 owns; R == 1 aliases simp[1]
  DevBuf<uint32_t> brow_ptr[DDF_MAXD + 1];  // n_{R-1}+1 offsets into bcol[R]
  DevBuf<uint32_t> bcol[DDF_MAXD + 1];      // (R+1) n_R entries: parent | (negative ? 1u<<31 : 0)

  DevBuf<double> coords;                 // V x C row-major
  DevBuf<double> weights;                // V
  DevBuf<double> vol[DDF_MAXD + 1];
  DevBuf<double> dualvol[DDF_MAXD + 1];
  DevBuf<double> cc[DDF_MAXD + 1];       // dual coords n_R x C (R >= 1; R = 0 is coords)
  DevBuf<uint8_t> isbnd[DDF_MAXD + 1];   // R < D
  bool has_isbnd = false;
  // diagonal operator payloads, built lazily by ops.cu
  DevBuf<double> star[DDF_MAXD + 1];     // dualvol ./ vol   (hodge(Pr,R),   ManifoldOps.jl:67)
  DevBuf<double> istar[DDF_MAXD + 1];    // vol ./ dualvol   (invhodge(Dl,R), ManifoldOps.jl:81)
  DevBuf<double> isbnd_d[DDF_MAXD + 1];  // 0.0 / 1.0 mask   (isboundary,    ManifoldOps.jl:34)

  // wave / Laplacian working set (built lazily by wave.cu): the assembled
  // L = laplace(Pr,0) = delta_1 d_0 in row-CSR with the diagonal split off.
  // Row i's neighbours are the other endpoints of its incident edges in edge-id
  // order == ascending vertex order, so adj_ptr aliases brow_ptr[1].
  bool has_wave = false;
  DevBuf<uint32_t> adj_ptr;              // V+1 row offsets (diagonal inline, chopped entries removed)
  DevBuf<int32_t> adj_nbr;               // nnz column j, ascending within a row
  DevBuf<double> adj_l;                  // nnz L[i,j] = fl(ih0[i] * star1[e]); L[i,i] = sequential sum of -L[i,j]
  int64_t adj_nnz = 0;
  // the same rows in SELL-32-256 (sell.cuh), used by the row kernels when the padding is small
  bool wave_sell = false;
  SellBufs lsell;
  // ... or as staged rows (stage.cuh) when every window's columns fit a staging buffer (preferred)
  StageBufs lstage;
  // SELL copies of the row-CSR of d_R (boundary / dual derivative apply); state 0 = not
  // built, 1 = in use, 2 = rejected (padding too large -> CSR kernel)
  int bsell_state[DDF_MAXD + 1] = {0};
  SellBufs bsell[DDF_MAXD + 1];
  // rows of <= 2 entries (faces of the top simplices): (entry0, entry1 | 0xFFFFFFFF) pairs, one 8-byte load per row
  DevBuf<uint32_t> bpair[DDF_MAXD + 1];
  uint32_t bup_cap[DDF_MAXD + 1] = {0};  // longest upper run of a 256-row window (doubles, rounded): shared memory of the staged kernel
  DevBuf<uint32_t> bup[DDF_MAXD + 1];    // upper-run offsets of the rows of the boundary matrix (ops.cu, upper_build): bsell then holds the lower entries only
  PackedBufs bpack[DDF_MAXD + 1];
  BitRows bbits[DDF_MAXD + 1];
  int64_t wave2_reach = -1;              // largest |column - row| of the L rows (dependency distance of the two-step wavefront)
  DevBuf<uint32_t> wave2_done;           // per-iteration completion counters of the two-step wavefront (zeroed per launch)
  DevBuf<double> utmp;                   // V   ping-pong buffer of the fused step
  DevBuf<double> ldinv;                  // V   1 ./ diag(L) for the fused Poisson relaxation step
  DevBuf<double> tsit_work;              // 16 V  stage derivatives and scratch of ddf_wave_tsit5 (allocated on first use)

  // slab halo of the distributed wave step (comm.cu): vertex planes along the
  // slowest axis; the first halo_lo / last halo_hi planes are ghosts owned by rank-1 / rank+1
  int64_t halo_plane = 0, halo_lo = 0, halo_hi = 0;
  // general halo of a mesh partitioned by vertex ranges (comm.cu, ddf_mesh_set_halo_lists): owned local vertices are
  // the contiguous range [gen_own_lo, gen_own_hi); ghost vertices are filled from explicit per-peer index lists
  void* halo_gen = nullptr;
  int64_t gen_own_lo = 0, gen_own_hi = -1;
  // fused peer-memory halo (comm.cu): arena + peer mappings, the step counter the flags carry;
  // state 0 = untried, 1 = in use, 2 = unavailable (NCCL send/recv is used instead)
  void* halo_p2p = nullptr;
  void* halo_lists = nullptr;            // per-rank-R pack / unpack index lists of ddf_halo_exchange_k (comm.cu)
  unsigned long long halo_step = 0;
  int halo_p2p_state = 0;
  int wave2_dist_tune = 0;               // the "wave" tuning value that agreement was made under
  int wave2_dist_state = 0;              // two leapfrog steps per launch on this slab: 0 = not asked yet, 1 = every rank can, 2 = no
  int halo_last = 0;                     // what the last distributed step used: 1 = peer memory, 2 = NCCL

  // hypercube generator parameters (slab meshes of the distributed wave step)
  bool is_hypercube = false;
  int64_t nelts[DDF_MAXD] = {0};
  int64_t hdiv = 0, z0 = 0;

  const int32_t* bnd_table(int R) const { return R == 1 ? simp[1].p : bnd[R].p; }
};

// implemented in mesh.cu
int ddf_mesh_require_assembled(ddf_mesh* m);
// implemented in geometry.cu
int ddf_mesh_require_geometry(ddf_mesh* m);
int ddf_mesh_require_isboundary(ddf_mesh* m);
// implemented in wave.cu
int ddf_mesh_require_wave(ddf_mesh* m);
// implemented in comm.cu
void ddf_halo_release(ddf_mesh* m);
