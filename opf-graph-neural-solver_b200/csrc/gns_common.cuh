// gns_common.cuh — layouts and device helpers shared by the forward and backward kernels.
//
// Data layout in shared memory ("grid-interleaved SoA"): every per-bus / per-line /
// per-generator quantity of the G grids a CTA works on is stored as  [quantity][item][G]
// with the grid index fastest.  A thread owns one item (bus or line) for VG consecutive
// grids and moves them with one 4*VG-byte access, so consecutive lanes touch consecutive
// words (conflict-free) and every warp-uniform weight fetched from shared memory is
// amortised over VG items.  Buses are stored in the plan's internal order (in-degree
// descending) so that the lanes of a warp loop over the same number of incoming lines.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gns {

constexpr int kMaxK = 64;
constexpr float kSlope = 0.01f;  // nn.LeakyReLU default, ref GNS/main.py:23

__host__ __device__ constexpr int pad4(int x) { return (x + 3) & ~3; }
// Row stride (floats) of a [rows][items] shared-memory array: multiple of 4 (16-byte rows) with
// an odd number of 16-byte groups, so that 128-bit reads of 8 consecutive ROWS at the same item
// offset fall into 8 different bank groups (used by the backward weight-gradient tiles).
__host__ __device__ constexpr int row_stride(int items) {
  int p = pad4(items);
  return ((p / 4) % 2 == 0) ? p + 4 : p;
}

// Backward kernel: per-bus rows double as MMA B operands ([feature][item], 128-bit loads by lanes
// (row g, quad t)); a stride of 16 mod 32 floats puts the 8 lanes of every 128-bit phase on 32 distinct banks.
__host__ __device__ constexpr int bwd_bus_stride(int items) {
  int p = pad4(items);
  while (p % 32 != 16) p += 4;
  return p;
}

// ---------------------------------------------------------------------------------
// Packed per-step weight layout.  Every matrix is stored [wide][HP]: the fast index is
// always the hidden-side index (padded to a multiple of 4 for 128-bit broadcast loads),
// the slow index the wide side (input of a first layer, output of a last layer).  The
// same rows serve forward (stream the wide index) and backward (dX and dW^T).
// Pairing: phi_v feeds L_v, phi_theta feeds L_theta, phi_m feeds L_m (ref GNS/main.py:157-176).
// ---------------------------------------------------------------------------------
struct WLayout {
  int L, H, multi;
  int HP, PO, POP, LPAD, DIN_L, NPHI;
  // inside a phi block
  int phi_w1m, phi_w1f, phi_b1, phi_w2, phi_b2, phi_w4, phi_b4, phi_size;
  // inside an L-net block
  int ln_w1, ln_b1, ln_w2, ln_b2, ln_wo, ln_bo_s, ln_bo_m, ln_size_s, ln_size_m;
  // step block: phi nets in pair order (v, theta, m | single), then L_v, L_theta, L_m, then the
  // fused aggregate->hidden blocks of the three pairs:  M[j][o] = sum_i W4[i][j] * W1_L[4+L+i][o]
  // ([H][HP]) followed by c[o] = sum_i b4[i] * W1_L[4+L+i][o] ([HP]).  They are derived from the
  // parameters after packing (forward) and carry dM / dc in the gradient buffer (backward).
  int off_phi[3], off_ln[3], off_mf[3], mf_size, wparams, wstep;
};

__host__ __device__ constexpr WLayout make_wlayout(int L, int H, bool multi) {
  WLayout w{};
  w.L = L; w.H = H; w.multi = multi ? 1 : 0;
  w.HP = pad4(H);
  w.PO = multi ? L : 1;
  w.POP = pad4(w.PO);
  w.LPAD = pad4(L);
  w.DIN_L = 4 + 2 * L;
  w.NPHI = multi ? 3 : 1;
  w.phi_w1m = 0;
  w.phi_w1f = w.phi_w1m + L * w.HP;
  w.phi_b1 = w.phi_w1f + 5 * w.HP;
  w.phi_w2 = w.phi_b1 + w.HP;
  w.phi_b2 = w.phi_w2 + H * w.HP;
  w.phi_w4 = w.phi_b2 + w.HP;
  w.phi_b4 = w.phi_w4 + w.PO * w.HP;
  w.phi_size = w.phi_b4 + w.POP;
  w.ln_w1 = 0;
  w.ln_b1 = w.ln_w1 + w.DIN_L * w.HP;
  w.ln_w2 = w.ln_b1 + w.HP;
  w.ln_b2 = w.ln_w2 + H * w.HP;
  w.ln_wo = w.ln_b2 + w.HP;
  w.ln_bo_s = w.ln_wo + w.HP;
  w.ln_bo_m = w.ln_wo + L * w.HP;
  w.ln_size_s = w.ln_bo_s + 4;
  w.ln_size_m = w.ln_bo_m + w.LPAD;
  int o = 0;
  for (int p = 0; p < 3; ++p) { w.off_phi[p] = (p < w.NPHI) ? o : 0; if (p < w.NPHI) o += w.phi_size; }
  // pair order q: 0 = v, 1 = theta, 2 = m
  w.off_ln[0] = o; o += w.ln_size_s;
  w.off_ln[1] = o; o += w.ln_size_s;
  w.off_ln[2] = o; o += w.ln_size_m;
  w.wparams = o;                       // floats that map 1:1 to parameters (+ padding)
  w.mf_size = H * w.HP + w.HP;
  for (int q = 0; q < 3; ++q) { w.off_mf[q] = o; o += w.mf_size; }
  w.wstep = o;
  return w;
}

// ---------------------------------------------------------------------------------
// Gradient accumulators of the backward kernel, in MMA-fragment order.  Every weight-gradient GEMM
// call owns a block of NT x kFragTile floats inside its warp's private accumulator, so a lane adds its
// cells with one or two 64-bit reductions and no index arithmetic; build_frag_map() (host) inverts the order once per model shape.
// Per pair q: phi per-line W2/b2 (11 wide rows), phi per-line W1f (5), phi W1m/b1 (L+1),
// L-net output layer (m-net: L rows x (H+1) columns; scalar nets: H+1 values, stored directly),
// L-net second layer (H+1), L-net first layer + fused block + bias (4+L+H+2).
// ---------------------------------------------------------------------------------
struct FragLayout { int w2l, w1f, w1m, out, w2, w1, net, step; };
// One 8-row tile owns kFragTile floats: [lane][2] cells of hidden columns 0..7 (64 floats), then [lane][2] cells of
// hidden columns 8..10 for the 12 lanes that hold them (24 floats).  Columns 11..15 of the m16 tile are padding
// and are never stored, which keeps the accumulators of 1,480 warps at 100 MB (< L2) instead of 145 MB.
constexpr int kFragTile = 64 + 24;
// The warp-specialised kernel (gns_backward2.cuh) maps MMA row g to hidden column 2g and row g+8 to column 2g+1, so a
// lane holds 4 cells of one 8-row tile and adds them with ONE 128-bit reduction: [lane = (c/2)*4 + (r%8)/2][4], lanes
// of hidden columns 0..11 only (24 lanes).
constexpr int kFragTile2 = 24 * 4;
__host__ __device__ constexpr int frag_tiles(int rows) { return (rows + 7) / 8; }
__host__ __device__ constexpr FragLayout make_frag_layout(int L, int H, int tile = kFragTile) {
  FragLayout f{};
  int o = 0;
  f.w2l = o; o += tile * frag_tiles(H + 1);
  f.w1f = o; o += tile * frag_tiles(5);
  f.w1m = o; o += tile * frag_tiles(L + 1);
  f.out = o; o += tile * frag_tiles(L);
  f.w2 = o; o += tile * frag_tiles(H + 1);
  f.w1 = o; o += tile * frag_tiles(4 + L + H + 2);
  f.net = o;
  f.step = (3 * o + 31) & ~31;
  return f;
}
__host__ __device__ constexpr int frag2_index(int r, int c) {
  return (r / 8) * kFragTile2 + ((c / 2) * 4 + (r % 8) / 2) * 4 + (c % 2) * 2 + (r % 2);
}
// position of cell (wide row r, hidden column c <= 10) inside a call's block
__host__ __device__ constexpr int frag_index(int r, int c) {
  return (r / 8) * kFragTile + (c < 8 ? 0 : 64) + ((c % 8) * 4 + (r % 8) / 2) * 2 + (r % 2);
}

// ---------------------------------------------------------------------------------
// Activation checkpoints (training only).  Besides the state entering every step, the forward
// kernel keeps the post-LeakyReLU hidden activations of every MLP evaluation, so that the backward
// kernel neither recomputes the forward nor needs the pre-activations (h > 0 <=> z > 0):
//   per bus and pair q:   A (aggregate, H), h1 and h2 of the L-net (H each)    rows [q][3H][NGs]
//   per line and phi net: h1 and h2 of the phi net (H each)                    rows [p][2H][EGs]
// a line is addressed through the plan's in_pos table (iteration-major over the slots that walk it).  172 floats per bus and step on
// case300: 867 KB per grid for K=4, streamed once out and once in (~1.3 TB/s at 0.75 M grids/s).
// ---------------------------------------------------------------------------------
// Two row layouts, chosen by the backward geometry (the reader):
//   grid-major  [row][grid][item]  when the backward kernel works on ONE grid per CTA (large grids): its lanes are
//               consecutive items of one grid and read consecutive floats; every (row, grid) segment is padded to
//               a 128-byte multiple.  The forward kernel pays one 32-bit store per grid instead of a vector store.
//   interleaved [row][item][grid]  when several grids share a CTA (small grids): the lanes of both kernels are
//               (item, grid) pairs with the grid index fastest, exactly this order.
// Element of (item, grid): item * is + grid * gs (bus block; ls / gl for the line block); an array of H rows takes H * rb
// floats (rl for the line block).  Grid-major: ordered as quads of rows over the elements (stg_rows4 / ldg_rows4);
// interleaved: plain rows, (row, element) at row * rb + element (a thread's VG grids are adjacent: one vector store per
// row).  The line block holds, per phi net, h1 (H rows) and ONE row of words with the LeakyReLU slope bits of h2 (bit o:
// h2[o] > 0) - the backward pass needs h2 only for its slope.
struct ActLayout { int rb, is, gs, rl, ls, gl, line_off, total; };
__host__ __device__ inline ActLayout make_act_layout(int H, int nphi, int Ns, int E, int G, bool grid_major) {
  ActLayout a{};
  if (grid_major) {
    const int nsp = (Ns + 31) & ~31, esp = (E + 31) & ~31;
    a.rb = G * nsp; a.is = 1; a.gs = nsp;
    a.rl = G * esp; a.ls = 1; a.gl = esp;
  } else {
    a.rb = pad4(Ns * G); a.is = G; a.gs = 1;
    a.rl = pad4(E * G); a.ls = G; a.gl = 1;
  }
  a.line_off = 3 * 3 * H * a.rb;
  a.total = a.line_off + nphi * (H + 1) * a.rl;      // per phi net: h1 (H rows) and one row of h2 slope bits
  return a;
}

// ---------------------------------------------------------------------------------
// Checkpoints for the warp-specialised backward kernel (gns_backward2.cuh, one grid per CTA): everything a
// (grid, step) needs is contiguous, every block is [row][item] with a row stride = 16 mod 32 floats (the MMA
// fragment loads of 8 rows x 4 quads are conflict-free once a block sits in shared memory), so a block moves
// with ONE bulk copy (cp.async.bulk).  Bus blocks are indexed by BUS RANK (no twin columns), line blocks by
// the plan's in_pos column.  Per pair q:  h2L, h1L, A (H rows x NbP each), h1 of the phi net per line
// (H rows x EP), and the LeakyReLU slope bits: row 0 = bits of (h1L, h2L) per slot, row 1+it = bits of
// (h1, h2) of the it-th line the slot walks (bit o: h1[o] > 0, bit H+o: h2[o] > 0), MR = 1 + max walk rows x NsM.
// h2L is stored ITEM-major ([item][H]): it is the hidden-side operand of the output-layer gradient (see cons_call).
// The forward zero-fills the padding columns of the float blocks (0 x garbage must not be NaN); there is always at
// least one padding column / item, which the backward kernel uses as the dump target of its branch-free stores.
// ---------------------------------------------------------------------------------
__host__ __device__ constexpr int mma_stride(int items) {
  int p = pad4(items);
  while (p % 32 != 16) p += 4;
  return p;
}
struct Act2Layout {
  int NbP, EP, NsM, MR;
  int h2L[3], h1L[3], A[3], h1line[3], mask[3];
  int step;               // floats per (grid, step)
  int state;              // floats per state checkpoint: (4+L) * NbP
};
__host__ __device__ inline Act2Layout make_act2_layout(int L, int H, int N, int Ns, int E, int maxwalk) {
  Act2Layout a{};
  a.NbP = mma_stride(N + 1); a.EP = mma_stride(E + 1); a.NsM = pad4(Ns); a.MR = 1 + maxwalk;
  int o = 0;
  for (int q = 0; q < 3; ++q) {
    a.h2L[q] = o; o += H * a.NbP;
    a.h1L[q] = o; o += H * a.NbP;
    a.A[q] = o; o += H * a.NbP;
    a.h1line[q] = o; o += H * a.EP;
    a.mask[q] = o; o += a.MR * a.NsM;
  }
  a.step = o;
  a.state = (4 + L) * a.NbP;
  return a;
}

// ---------------------------------------------------------------------------------
// Topology index block (shared by all grids; copied to shared memory once per CTA).
// All entries are uint16 (n_bus, n_line < 65536).
// ---------------------------------------------------------------------------------
struct TopoOffsets {      // offsets in uint16 units inside the index block
  // Internal index space = SLOTS.  Every bus owns one primary slot; a bus with more than
  // `cap` incoming lines owns 2 or 4 adjacent slots ("twins") that split its in-lines, so
  // no lane walks more than ~cap lines.  Ns = number of slots >= N.
  int fi, ti;             // [E] primary slot of from / to bus
  int fa, ta;             // [E] alias line ids: external from / to bus number re-read as a line id
  int in_b, in_e, in_fe;  // [Ns] in-line range walked by this slot; end of the bus's full range (primary)
  int in_ids;             // [E] line ids grouped by receiving bus (ascending line id inside a bus)
  int in_pos;             // [E] activation-checkpoint column of in-list position e: positions walked in the same iteration by
                          //     consecutive slots are consecutive, so the line records of a warp are written / read coalesced
  int out_b, out_e;       // [Ns] out-line range (primary slots only, else empty)
  int out_ids;            // [E]
  int gen_b, gen_e;       // [Ns] generator range (primary slots only)
  int gen_ids;            // [Gn]
  int ext_of;             // [Ns] external bus of the slot
  int prim_of;            // [Ns] primary slot of the slot's bus
  int gsz;                // [Ns] slots in this bus's twin group (1, 2 or 4), aligned to gsz
  int rank_of;            // [N]  primary slot of external bus
  int brank;              // [Ns] bus rank (position in the degree-descending bus order, 0..N-1) of the slot's bus
  int total;              // padded to a multiple of 8: what the forward / first backward kernel copy to shared memory
  // extension read by the warp-specialised backward kernel only (gns_backward2.cuh)
  int fr, tr;             // [E] bus rank of the from / to bus
  int ext_rank;           // [N] bus rank -> external bus
  // lines as items (gns_backward3.cuh), indexed by the line's activation column c = in_pos
  int col_slot, col_it;   // [E] slot that walks the line and its position in that slot's walk (address of the slope word)
  int col_brank;          // [E] bus rank of the receiving bus
  int col_line;           // [E] line id
  int rin_b;              // [N+1] in-line range of the bus with this rank inside rin_cols
  int rin_cols;           // [E] activation columns of the lines entering a bus, grouped by bus rank (ascending line id)
  int total_ext;          // padded to a multiple of 8
};

__host__ __device__ inline TopoOffsets make_topo_offsets(int N, int Ns, int E, int Gn) {
  TopoOffsets t{};
  int o = 0;
  t.fi = o; o += E;
  t.ti = o; o += E;
  t.fa = o; o += E;
  t.ta = o; o += E;
  t.in_b = o; o += Ns;
  t.in_e = o; o += Ns;
  t.in_fe = o; o += Ns;
  t.in_ids = o; o += E;
  t.in_pos = o; o += E;
  t.out_b = o; o += Ns;
  t.out_e = o; o += Ns;
  t.out_ids = o; o += E;
  t.gen_b = o; o += Ns;
  t.gen_e = o; o += Ns;
  t.gen_ids = o; o += Gn;
  t.ext_of = o; o += Ns;
  t.prim_of = o; o += Ns;
  t.gsz = o; o += Ns;
  t.rank_of = o; o += N;
  t.brank = o; o += Ns;
  t.total = (o + 7) & ~7;
  o = t.total;
  t.fr = o; o += E;
  t.tr = o; o += E;
  t.ext_rank = o; o += N;
  t.col_slot = o; o += E;
  t.col_it = o; o += E;
  t.col_brank = o; o += E;
  t.col_line = o; o += E;
  t.rin_b = o; o += N + 1;
  t.rin_cols = o; o += E;
  t.total_ext = (o + 7) & ~7;
  return t;
}

// ---------------------------------------------------------------------------------
// Vector-over-grids helpers.
// ---------------------------------------------------------------------------------
template <int VG> struct VecIO;
template <> struct VecIO<1> {
  __device__ static __forceinline__ void ld(float (&x)[1], const float* p) { x[0] = *p; }
  __device__ static __forceinline__ void st(float* p, const float (&x)[1]) { *p = x[0]; }
};
template <> struct VecIO<2> {
  __device__ static __forceinline__ void ld(float (&x)[2], const float* p) {
    float2 t = *reinterpret_cast<const float2*>(p); x[0] = t.x; x[1] = t.y;
  }
  __device__ static __forceinline__ void st(float* p, const float (&x)[2]) {
    *reinterpret_cast<float2*>(p) = make_float2(x[0], x[1]);
  }
};
template <> struct VecIO<4> {
  __device__ static __forceinline__ void ld(float (&x)[4], const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p); x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
  }
  __device__ static __forceinline__ void st(float* p, const float (&x)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  }
};

// streaming global stores / loads of VG consecutive grids (activation checkpoints)
template <int VG> __device__ __forceinline__ void stg_stream(float* p, const float (&x)[VG]);
template <> __device__ __forceinline__ void stg_stream<1>(float* p, const float (&x)[1]) { __stcs(p, x[0]); }
template <> __device__ __forceinline__ void stg_stream<2>(float* p, const float (&x)[2]) {
  __stcs(reinterpret_cast<float2*>(p), make_float2(x[0], x[1]));
}
template <> __device__ __forceinline__ void stg_stream<4>(float* p, const float (&x)[4]) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(x[0], x[1], x[2], x[3]));
}
// one streaming 32-bit store per grid, `gstride` floats apart (grid-major activation rows)
template <int VG> __device__ __forceinline__ void stg_grids(float* p, int gstride, const float (&x)[VG]) {
#pragma unroll
  for (int g = 0; g < VG; ++g) __stcs(p + g * gstride, x[g]);
}
// N rows of VG grids each, `rstride` floats apart.  INTER: the thread's grids are adjacent in memory (interleaved
// activation layout): one vector store per row; else one 32-bit store per grid, `gstride` floats apart.
template <int N, int VG, bool INTER> __device__ __forceinline__ void stg_rows(float* p, int rstride, int gstride, const float (&x)[N][VG]) {
#pragma unroll
  for (int o = 0; o < N; ++o) {     // (pointer bumped row by row: N independent row offsets would be hoisted into spills)
    if constexpr (INTER) stg_stream<VG>(p, x[o]);
    else stg_grids<VG>(p, gstride, x[o]);
    p += rstride;
  }
}

// Activation arrays of the first backward kernel (ActLayout): the H rows of one hidden vector are kept as QUADS of rows,
// [element][4] for rows 4c..4c+3 at arr + 4 c rb and [element][2] for the last two rows (H = 10) at arr + (H-2) rb, with
// element = item * is + grid * gs.  A thread writes / reads a vector with H/4 128-bit and one 64-bit access instead of H
// 32-bit ones, and the lanes of a warp (consecutive elements) still cover consecutive addresses.
// checkpoint store policy (A/B knob): 0 = st.global.cs (evict first), 1 = default policy, 3 = diagnostic: no store at all
#ifndef GNS_STORE_MODE
#define GNS_STORE_MODE 0
#endif
template <class V> __device__ __forceinline__ void ckpt_store(V* p, V v) {
#if GNS_STORE_MODE == 0
  __stcs(p, v);
#elif GNS_STORE_MODE == 1
  *p = v;
#else
  if (reinterpret_cast<size_t>(p) == 1) *p = v;      // keeps the value alive, never stores
#endif
}
template <int H, int VG>
__device__ __forceinline__ void stg_rows4(float* arr, int rb, int e0, int gs, const float (&x)[H][VG]) {
  static_assert(H % 4 == 2, "quads of rows plus one pair");
#pragma unroll
  for (int g = 0; g < VG; ++g) {
    const int e = e0 + g * gs;
    float* p = arr + 4 * e;
#pragma unroll
    for (int c = 0; c < H / 4; ++c) {
      ckpt_store(reinterpret_cast<float4*>(p), make_float4(x[4 * c][g], x[4 * c + 1][g], x[4 * c + 2][g], x[4 * c + 3][g]));
      p += 4 * rb;
    }
    ckpt_store(reinterpret_cast<float2*>(arr + (size_t)(H - 2) * rb + 2 * e), make_float2(x[H - 2][g], x[H - 1][g]));
  }
}
template <int H>
__device__ __forceinline__ void ldg_rows4(float* x /* [H] */, const float* arr, size_t rb, int e) {
  static_assert(H % 4 == 2, "quads of rows plus one pair");
#pragma unroll
  for (int c = 0; c < H / 4; ++c) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(arr + 4 * c * rb + 4 * (size_t)e));
    x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
  }
  const float2 v = __ldg(reinterpret_cast<const float2*>(arr + (H - 2) * rb + 2 * (size_t)e));
  x[H - 2] = v.x; x[H - 1] = v.y;
}

__device__ __forceinline__ float lrelu(float x) { return fmaxf(x, kSlope * x); }
// LeakyReLU / accumulate over the VG grids of a thread; two grids use the packed f32x2 multiply / add of sm_100
// (FMUL2 / FADD2: one issue slot for both grids), the max stays scalar
template <int VG>
__device__ __forceinline__ void lrelu_vec(float (&x)[VG]) {
  if constexpr (VG == 2) {
    const float2 t = __fmul2_rn(make_float2(x[0], x[1]), make_float2(kSlope, kSlope));
    x[0] = fmaxf(x[0], t.x); x[1] = fmaxf(x[1], t.y);
  } else {
#pragma unroll
    for (int g = 0; g < VG; ++g) x[g] = lrelu(x[g]);
  }
}
template <int VG>
__device__ __forceinline__ void add_vec(float (&a)[VG], const float (&b)[VG]) {
  if constexpr (VG == 2) {
    const float2 t = __fadd2_rn(make_float2(a[0], a[1]), make_float2(b[0], b[1]));
    a[0] = t.x; a[1] = t.y;
  } else {
#pragma unroll
    for (int g = 0; g < VG; ++g) a[g] += b[g];
  }
}
// LeakyReLU slope bits of the post-activation values h (h > 0 <=> z > 0): bit (shift + o) of w[g] is set where the slope is 1
template <int H, int VG>
__device__ __forceinline__ void slope_bits(uint32_t (&w)[VG], const float (&h)[H][VG], int shift) {
#pragma unroll
  for (int o = 0; o < H; ++o)
#pragma unroll
    for (int g = 0; g < VG; ++g) w[g] |= (h[o][g] > 0.f) ? (1u << (shift + o)) : 0u;
}
__device__ __forceinline__ float lrelu_grad(float z) { return z > 0.f ? 1.f : kSlope; }   // also valid on h = lrelu(z)

// sin and cos of one float, branch-free: three-constant Cody-Waite reduction by pi/2 with FMAs,
// then minimax polynomials on [-pi/4, pi/4].  Max abs error 9.3e-8 for |x| <= 1e5, 1.2e-7 for
// |x| <= 1e6 (measured against double); beyond that it degrades smoothly (1e-4 at 1e7) but stays
// far below the effect of the argument's own rounding (ulp(1e6) = 0.06 rad).  Arguments here are
// angle differences of a few radians.  Inf / NaN give NaN like sinf / cosf.  Keeping the
// library's Payne-Hanek slow path (and any call or branch) out of the persistent kernels saves
// registers, ~60 KB of code, and lets the compiler interleave the independent sincos chains.
__device__ __forceinline__ void fast_sincos(float x, float& s, float& c) {
  const float q = rintf(x * 0.63661977236758134f);
  float r = fmaf(q, -1.57079637050628662e+00f, x);
  r = fmaf(q, 4.37113900018624283e-08f, r);
  r = fmaf(q, 1.71512449632429490e-15f, r);
  const int qi = (int)q;
  const float r2 = r * r;
  float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, r2, -1.6666654611e-1f);
  const float sr = fmaf(ps * r2, r, r);
  float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, r2, 4.166664568298827e-2f);
  const float cr = fmaf(pc * r2, r2, fmaf(-0.5f, r2, 1.0f));
  const float s0 = (qi & 1) ? cr : sr;
  const float c0 = (qi & 1) ? sr : cr;
  s = (qi & 2) ? -s0 : s0;
  c = ((qi + 1) & 2) ? -c0 : c0;
}

// ---- packed pairs: the two grids of a VG=2 thread as one f32x2 value (FADD2 / FMUL2 / FFMA2 of sm_100) ----
// Same IEEE roundings per component as the scalar operators; a - b is fma(b, -1, a) (exact product).
struct Pk2 { float2 v; };
__device__ __forceinline__ Pk2 pk2(float a, float b) { return Pk2{make_float2(a, b)}; }
__device__ __forceinline__ Pk2 pk2(float a) { return Pk2{make_float2(a, a)}; }
__device__ __forceinline__ Pk2 operator+(Pk2 a, Pk2 b) { return Pk2{__fadd2_rn(a.v, b.v)}; }
__device__ __forceinline__ Pk2 operator-(Pk2 a, Pk2 b) { return Pk2{__ffma2_rn(b.v, make_float2(-1.f, -1.f), a.v)}; }
__device__ __forceinline__ Pk2 operator*(Pk2 a, Pk2 b) { return Pk2{__fmul2_rn(a.v, b.v)}; }
__device__ __forceinline__ Pk2 operator*(Pk2 a, float b) { return Pk2{__fmul2_rn(a.v, make_float2(b, b))}; }
__device__ __forceinline__ Pk2 operator-(Pk2 a) { return Pk2{make_float2(-a.v.x, -a.v.y)}; }
__device__ __forceinline__ Pk2 vfma(Pk2 a, Pk2 b, Pk2 c) { return Pk2{__ffma2_rn(a.v, b.v, c.v)}; }
__device__ __forceinline__ Pk2 vfma(Pk2 a, float b, Pk2 c) { return Pk2{__ffma2_rn(a.v, make_float2(b, b), c.v)}; }
__device__ __forceinline__ Pk2 vfma(Pk2 a, float b, float c) { return Pk2{__ffma2_rn(a.v, make_float2(b, b), make_float2(c, c))}; }
__device__ __forceinline__ Pk2 vabs(Pk2 a) { return Pk2{make_float2(fabsf(a.v.x), fabsf(a.v.y))}; }
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vabs(float a) { return fabsf(a); }

// fast_sincos for a packed pair: the Cody-Waite reduction and both polynomials run as f32x2 operations,
// only the rounding to the quadrant and the final selects are per component.  Bit-identical to fast_sincos.
__device__ __forceinline__ void fast_sincos(Pk2 x, Pk2& s, Pk2& c) {
  const Pk2 t = x * 0.63661977236758134f;
  const Pk2 q = pk2(rintf(t.v.x), rintf(t.v.y));
  Pk2 r = vfma(q, -1.57079637050628662e+00f, x);
  r = vfma(q, 4.37113900018624283e-08f, r);
  r = vfma(q, 1.71512449632429490e-15f, r);
  const int qx = (int)q.v.x, qy = (int)q.v.y;
  const Pk2 r2 = r * r;
  Pk2 ps = vfma(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = vfma(ps, r2, pk2(-1.6666654611e-1f));
  const Pk2 sr = vfma(ps * r2, r, r);
  Pk2 pc = vfma(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = vfma(pc, r2, pk2(4.166664568298827e-2f));
  const Pk2 cr = vfma(pc * r2, r2, vfma(r2, -0.5f, 1.0f));
  const float sx0 = (qx & 1) ? cr.v.x : sr.v.x, cx0 = (qx & 1) ? sr.v.x : cr.v.x;
  const float sy0 = (qy & 1) ? cr.v.y : sr.v.y, cy0 = (qy & 1) ? sr.v.y : cr.v.y;
  s = pk2((qx & 2) ? -sx0 : sx0, (qy & 2) ? -sy0 : sy0);
  c = pk2(((qx + 1) & 2) ? -cx0 : cx0, ((qy + 1) & 2) ? -cy0 : cy0);
}

// Per-line Kirchhoff terms of one step (ref GNS/main.py:38-41, 66-72, 87-99), for one
// grid (T = float) or the two grids of a thread at once (T = Pk2).  Inputs: v / theta of the line's ends,
// Y, b, 1/tau, shift of the alias lines A = f_e and B = t_e, and D, sin D, cos D of those alias lines.
template <class T>
struct LineTerms { T msg, pf, qf, pt, qt; };
template <class T>
__device__ __forceinline__ LineTerms<T> line_terms(T vf, T vt, T thf, T tht, T Yf, T bf, T itf, T shf, T Df, T sDf, T cDf,
                                                  T Yt, T bt, T itt, T sht, T Dt, T sDt) {
  const T a1 = ((thf - tht) - Df) - shf;
  const T a2 = ((tht - thf) - Df) + shf;
  const T a3 = ((tht - thf) + Dt) - sht;              // delta_ji = -delta_ij, re-read through the alias
  T s1, c1, s2, c2, s3, c3;
  fast_sincos(a1, s1, c1);
  fast_sincos(a2, s2, c2);
  fast_sincos(a3, s3, c3);
  (void)c2;
  const T t1 = ((vf * vt) * Yf) * itf;
  const T vft = vf * itf;                               // v_f / tau_f
  const T vft2 = vft * vft;
  const T yfs = Yf * sDf;
  LineTerms<T> o;
  // | v_f v_t Y/tau (sin a1 + sin a2) + (v_f / tau^2) Y sin D + v_t^2 Y sin D |   (note v_f / tau^2, quirk Q5)
  o.msg = vabs(vfma(t1, s1 + s2, vfma(vf * (itf * itf), yfs, (vt * vt) * yfs)));
  o.pf = vfma(t1, s1, vft2 * yfs);
  o.qf = vfma(t1 * -1.0f, c1, vft2 * vfma(Yf, cDf, bf * -0.5f));
  const T u1 = ((vt * vf) * Yt) * itt;
  const T sdt = sDt * -1.0f;
  const T vt2 = vt * vt;
  o.pt = vfma(u1, s3, (vt2 * Yt) * sdt);
  o.qt = vfma(u1 * -1.0f, c3, vt2 * vfma(Yt, sdt, bt * -0.5f));   // yes sin, the reference's own line (quirk Q5)
  return o;
}

// An integer zero the optimiser cannot see through.  Added to a shared-memory weight
// pointer inside a data-dependent loop it stops loop-invariant code motion from hoisting
// every weight row of the loop body into registers (and from there into local-memory
// spills): the rows must be re-read from shared memory as warp-uniform broadcasts.
__device__ __forceinline__ int opaque_zero() {
  int z;
  asm volatile("mov.u32 %0, 0;" : "=r"(z));
  return z;
}

// Load one padded weight row (HP floats, 16-byte aligned, warp-uniform address).
template <int HP>
__device__ __forceinline__ void load_row(float (&w)[HP], const float* row) {
#pragma unroll
  for (int c = 0; c < HP / 4; ++c) {
    float4 t = *reinterpret_cast<const float4*>(row + 4 * c);
    w[4 * c + 0] = t.x; w[4 * c + 1] = t.y; w[4 * c + 2] = t.z; w[4 * c + 3] = t.w;
  }
}

// Packed FP32 pairs: sm_100 issues two FMAs per lane in one FFMA2 (fma.rn.f32x2) and accepts a
// scalar register as a broadcast operand, so a weight-times-two-grids (VG=2) or a
// two-weights-times-one-activation (VG=1) update costs one issue slot instead of two.  Measured
// FFMA2 peak on B200 = FFMA peak (74.0 vs 72.3 TFLOP/s): the freed issue slots go to LDS / ALU work.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// acc[o][g] += x[g] * row[o]
template <int H, int HP, int VG>
__device__ __forceinline__ void row_axpy(float (&acc)[H][VG], const float (&x)[VG], const float* row) {
  float w[HP];
  load_row<HP>(w, row);
  if constexpr (VG == 2) {
    const float2 xv = make_float2(x[0], x[1]);
#pragma unroll
    for (int o = 0; o < H; ++o) {
      const float2 r = fma2(xv, make_float2(w[o], w[o]), make_float2(acc[o][0], acc[o][1]));
      acc[o][0] = r.x; acc[o][1] = r.y;
    }
  } else if constexpr (VG == 1) {
    const float2 xv = make_float2(x[0], x[0]);
#pragma unroll
    for (int o = 0; o + 1 < H; o += 2) {
      const float2 r = fma2(make_float2(w[o], w[o + 1]), xv, make_float2(acc[o][0], acc[o + 1][0]));
      acc[o][0] = r.x; acc[o + 1][0] = r.y;
    }
    if (H & 1) acc[H - 1][0] = fmaf(x[0], w[H - 1], acc[H - 1][0]);
  } else {
#pragma unroll
    for (int o = 0; o < H; ++o)
#pragma unroll
      for (int g = 0; g < VG; ++g) acc[o][g] = fmaf(x[g], w[o], acc[o][g]);
  }
}

// out[g] += sum_o h[o][g] * row[o]
template <int H, int HP, int VG>
__device__ __forceinline__ void row_dot(float (&out)[VG], const float (&h)[H][VG], const float* row) {
  float w[HP];
  load_row<HP>(w, row);
  if constexpr (VG == 2) {
    float2 t0 = make_float2(out[0], out[1]), t1 = make_float2(0.f, 0.f);   // two chains: shorter dependency
#pragma unroll
    for (int o = 0; o + 1 < H; o += 2) {
      t0 = fma2(make_float2(h[o][0], h[o][1]), make_float2(w[o], w[o]), t0);
      t1 = fma2(make_float2(h[o + 1][0], h[o + 1][1]), make_float2(w[o + 1], w[o + 1]), t1);
    }
    if (H & 1) t0 = fma2(make_float2(h[H - 1][0], h[H - 1][1]), make_float2(w[H - 1], w[H - 1]), t0);
    const float2 tt = __fadd2_rn(t0, t1);
    out[0] = tt.x; out[1] = tt.y;
  } else if constexpr (VG == 1) {
    float2 t = make_float2(0.f, 0.f);
#pragma unroll
    for (int o = 0; o + 1 < H; o += 2) t = fma2(make_float2(h[o][0], h[o + 1][0]), make_float2(w[o], w[o + 1]), t);
    float r = out[0] + (t.x + t.y);
    if (H & 1) r = fmaf(h[H - 1][0], w[H - 1], r);
    out[0] = r;
  } else {
#pragma unroll
    for (int o = 0; o < H; ++o)
#pragma unroll
      for (int g = 0; g < VG; ++g) out[g] = fmaf(h[o][g], w[o], out[g]);
  }
}

// ---- TMA (cp.async.bulk) staging of the next batch's raw input rows, tracked by an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier over a subset of the CTA's warps (id 1..15; `count` = participating threads, a multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// global -> shared bulk copy (16-byte aligned addresses and size), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 4-byte asynchronous global -> shared copies (LDGSTS): no register staging, no scoreboard wait at the issue
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 16-byte aligned window [begin, begin+bytes) covering elements [first, first+count) of a float array
struct BulkWindow { long long begin; uint32_t bytes; int shift; };   // shift = floats between window start and `first`
__device__ __forceinline__ BulkWindow bulk_window(long long first, int count) {
  const long long o = first * 4, oa = o & ~15LL;
  BulkWindow w;
  w.begin = oa;
  w.bytes = (uint32_t)(((o + (long long)count * 4 + 15) & ~15LL) - oa);
  w.shift = (int)((o - oa) >> 2);
  return w;
}

// same de-interleave as load_block, but from the staged raw rows in shared memory; one thread per
// (grid, row): one integer division per row instead of two per element
__device__ __forceinline__ void unpack_block(const float* __restrict__ raw, float* __restrict__ dst, int G, int rows,
                                             int cols, int c0, int dst_stride, const uint16_t* __restrict__ slot_of_row) {
  const int total = rows * G;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int gl = idx / rows;
    const int row = idx - gl * rows;
    const int slot = slot_of_row ? (int)slot_of_row[row] : row;
    const float* src = raw + idx * cols;
    float* d = dst + slot * G + gl;
    for (int c = c0; c < cols; ++c) d[(c - c0) * dst_stride] = src[c];
  }
}

// ---- staged input load: reference AoS rows -> grid-interleaved SoA in shared memory ----
// src: [S][rows][cols]; keeps columns c0..cols-1 as dst[(c-c0)][slot(row)][G] (+gl).
__device__ __forceinline__ void load_block(const float* __restrict__ src, float* __restrict__ dst,
                                           long long g0, long long S, int G, int rows, int cols, int c0,
                                           int dst_stride, const uint16_t* __restrict__ slot_of_row) {
  const int total = rows * G;                 // one thread per (grid, row): its columns are independent loads
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int gl = idx / rows;
    const int row = idx - gl * rows;
    long long g = g0 + gl;
    if (g >= S) g = S - 1;  // tail batch: replicate the last grid (results are not stored)
    const float* s = src + (g * rows + row) * cols;
    const int slot = slot_of_row ? (int)slot_of_row[row] : row;
    float* d = dst + slot * G + gl;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = (c >= c0 && c < cols) ? __ldg(s + c) : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c >= c0 && c < cols) d[(c - c0) * dst_stride] = v[c];
  }
}

// Deterministic per-grid block reduction of NV values at once.  Thread `tid` contributes
// x[NV][VG] for grids (tid % NGQ)*VG .. +VG.  Requires 32 % NGQ == 0 and blockDim.x % 32 == 0.
// `red` holds 2 * nwarps * kRedNV * G floats: the two halves alternate (`parity`), so ONE barrier
// per call is enough - a thread can only re-enter the half it is still being read from after
// every thread has passed the barrier of the call in between.  Every thread receives the totals
// of its own grids; the warp partials are added in a fixed order.
constexpr int kRedNV = 4;
template <int VG, int NV>
__device__ __forceinline__ void block_sum_multi(float (&x)[NV][VG], float* red, int NGQ, int& parity) {
  static_assert(NV <= kRedNV, "too many values");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int G = NGQ * VG;
  float* buf = red + parity * (nwarps * kRedNV * G);
  parity ^= 1;
  for (int off = 16; off >= NGQ; off >>= 1) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int g = 0; g < VG; ++g) x[v][g] += __shfl_xor_sync(0xffffffffu, x[v][g], off);
  }
  if (lane < NGQ) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int g = 0; g < VG; ++g) buf[(warp * kRedNV + v) * G + lane * VG + g] = x[v][g];
  }
  __syncthreads();
  const int gq = lane % NGQ;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int g = 0; g < VG; ++g) x[v][g] = 0.f;
#pragma unroll 1
  for (int w = 0; w < nwarps; ++w) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int g = 0; g < VG; ++g) x[v][g] += buf[(w * kRedNV + v) * G + gq * VG + g];
  }
}

template <int VG>
__device__ __forceinline__ void block_sum_per_grid(float (&x)[VG], float* red, int NGQ, int& parity) {
  float (&xv)[1][VG] = reinterpret_cast<float (&)[1][VG]>(x);
  block_sum_multi<VG, 1>(xv, red, NGQ, parity);
}

// ---------------------------------------------------------------------------------
// Kernel argument blocks
// ---------------------------------------------------------------------------------
struct SmemPlan {          // offsets in floats from the start of dynamic shared memory
  int state;               // [(4+L)][N][G]  v, theta, dP, dQ, m[0..L)
  int busc;                // [4][N][G]      Pd, Qd, Gs, Bs
  int genc;                // [6][Gn][G]     Pmax, Pmin, Pset, vg, qg0, Pg0
  int linef;               // [5][E][G]      r, x, b, tau, shift
  int yline;               // [2][N][G]      1/sqrt(r^2+x^2) and 1/tau of lines 0..N-1 (alias lines)
  int trig;                // [3][N][G]      D, sin D, cos D of lines 0..N-1
  int flows;               // [4][E][G]      p_from, q_from, p_to, q_to
  int gsum;                // [4][G]         sum Pd, sum Pset, sum Pmin, sum Pmax
  int red;                 // [nwarps][G]
  int weights;             // [wstep]
  int topo;                // uint16 block (offset in floats)
  int stage_b, stage_l, stage_g;   // raw AoS staging of the next batch (forward; TMA bulk copies), 0 if unused
  int mbar;                // 8-byte mbarrier of the input staging
  int weights2;            // second weight buffer (TMA double buffering of the per-step weights), 0 if unused
  int mbar_w;              // its mbarrier
  int extra;               // backward-only regions start here
  int total_floats;
};

struct FwdArgs {
  const float* params;     // packed [K][wstep]
  const float* buses; const float* lines; const float* gens;
  float* v; float* theta; float* total; float* last;
  float* ckpt;             // [nbatch][K][(4+L)][N][G] or null
  float* pglob;            // [nbatch][K][G] or null
  float* act;              // [nbatch][K][ActLayout.total] hidden activations for the backward pass, or null
  const uint16_t* topo;    // index block in global memory
  long long S;
  int N, Ns, E, Gn, K, NGQ, G, nbatch;   // Ns = bus slots (>= N)
  int NGs, EGs;            // padded row strides of the [.][Ns][G] and [.][E][G] arrays
  int need_grad;
  ActLayout al;
  Act2Layout a2;           // GRADV = 3: per-grid checkpoints for the warp-specialised backward kernel
  float* ck2;              // [S][K+1][a2.state] state entering step k (k = 0..K-1) and the final state (K)
  int compact;             // inference on the compact input format: buses / lines / gens point to [S][N][2] / [S][E][5] / [S][Gn][2]
  const float* cbus;       // [N][4]  bus_i, type, Gs, Bs   (compact format: constants of the case)
  const float* cgen;       // [Gn][4] bus_i, Pmax, Pmin, qg
  int use_tma;             // inputs 16-byte aligned and staging present: prefetch the next batch with cp.async.bulk
  unsigned char grp_of_warp[32];   // warp -> group of 32/NGQ consecutive bus slots
  SmemPlan sm;
  TopoOffsets to;
  float wk[kMaxK];         // gamma^(K-k) rounded to float like the reference's python scalar
};

}  // namespace gns
