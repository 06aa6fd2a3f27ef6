// gns_backward3.cuh — backward kernel with the MLP adjoint in MMA-fragment space (one grid per CTA, all warps on
// the same code).
//
// What the two earlier kernels showed: the cost of the backward pass is instruction issue, a warp
// of this code runs at ~0.13 instructions per cycle whatever it does, so every warp must stay busy on the same
// instruction stream, and the savings have to come from instructions per MAC.  Here the per-item matrix-vector
// products of the adjoint (dX: one instruction per MAC as FFMA2 chains in gns_backward.cuh) run on the tensor cores
// as well:
//   * items are tiled by 16 (MMA M); a hidden vector of a tile is an m16n8 accumulator fragment per 8 columns.  Under
//     the k permutation  k-slot t <-> column 8ks+2t,  k-slot t+4 <-> column 8ks+2t+1  the C fragment of one product
//     IS the A fragment of the next (rows g, g+8 / two adjacent columns per lane), so the layers
//     g_out -> dh2 -> d2 -> d1 -> (adj state, adj aggregate) chain in registers without shuffles; LeakyReLU slopes
//     come from the bit masks the training forward stores; weights are B fragments straight from the packed rows
//     ([row][hidden] = two adjacent k values per 64-bit load), pre-split once per step into TF32 big / small parts.
//   * buses (by rank, no twin slots) and lines (by activation column) are two item spaces: the aggregate's adjoint
//     goes bus -> line through a shared-memory block (gather by the line's receiving bus), the pre-activation
//     adjoint of the phi net's first layer goes line -> bus through another (CSR gather in ascending line order).
//   * weight gradients: every 16-item tile is one k16 chunk of the long-K scheme of gns_backward2.cuh (hid operand
//     item-major, wide operand = bulk-copied row blocks of the forward's per-grid checkpoints), flushed per tile
//     with one red.global.add.v4 per lane and 8-row tile (the phi net's two small matrices accumulate over the warp's
//     line tiles first).
// Checkpoint layout: Act2Layout (training forward GRADV = 3).  Physics adjoint: one thread per bus, as before.
// Measured: it executes as many instructions as the first kernel (hidden_dim 10 pads to 16 on every MMA
// axis, a FP32-accurate product is three HMMA plus two operand splits) and is 11-25 % slower; opt-in, GNS_BWD3=1.
#pragma once
#include "gns_backward2.cuh"

namespace gns {

constexpr int kB3Tile = 128;      // accumulator cells of the m-net's output-layer gradient: [mt][nt][lane][4]

struct Bwd3Smem {          // offsets in floats
  int state;               // [(4+L)][NbP]  state entering the step (wide rows), bulk copy
  int nxt;                 // [2][3][NbP]   v', theta', dP' leaving the step, bulk copy, double buffered
  int adj4;                // [4][NbP]      adjoint of (v, theta, dP, dQ) between steps
  int a4;                  // [4][NbP]      this step's additions to it (from the fragments)
  int gout;                // [2][NbP]      output adjoints of the scalar nets (q = 0: v, 1: theta)
  int gdP;                 // [NbP]
  int cst;                 // [6][NbP]
  int trig;                // [3][NbP]
  int adjD;                // [NbP]
  int deg;                 // [NbP]
  int ones_b;              // [NbP]
  int ones_l;              // [EP]
  int zrow;                // [EP]
  int lineg;               // [5][EP]
  int featp;               // [5][EP]
  int w;                   // [wstep]  weights (the raw value is the TF32 "big" part)
  int ws;                  // [wstep]  their TF32 "small" parts
  int act;                 // [3][H][NbP] h2L (item-major), h1L, A of the pair; then [H][EP] h1 of the lines
  int act_l;
  int adjA;                // [NbP][H]  item-major: adjoint of the aggregate, bus -> line
  int d1l;                 // [EP][H]   item-major: adjoint of the phi first-layer pre-activation, line -> bus (and hid of dW1f)
  int scratch;             // [nwarps][kScr] per-warp item-major tiles
  int red;
  int topo;                // uint16 arrays
  int mbar;
  int total;
};
constexpr int kB3Scr = 16 * 24;     // adjoint-state tile [16][24]; later in the same tile: two hidden tiles [16][10]

struct Bwd3Topo { int fa, ta, fr, tr, in_ids, in_pos, out_ids, col_slot, col_it, col_brank, rin_cols, rin_b; int total; };
__host__ __device__ inline Bwd3Topo make_bwd3_topo(int N, int E) {
  Bwd3Topo t{};
  const int Ep = pad4(E);
  int o = 0;
  t.fa = o; o += Ep; t.ta = o; o += Ep; t.fr = o; o += Ep; t.tr = o; o += Ep;
  t.in_ids = o; o += Ep; t.in_pos = o; o += Ep; t.out_ids = o; o += Ep;
  t.col_slot = o; o += Ep; t.col_it = o; o += Ep; t.col_brank = o; o += Ep; t.rin_cols = o; o += Ep;
  t.rin_b = o; o += pad4(N + 1);
  t.total = o;
  return t;
}

__host__ __device__ inline Bwd3Smem make_bwd3_smem(int L, int H, int N, int E, int wstep, int nwarps, const Act2Layout& a2) {
  Bwd3Smem s{};
  int o = 0;
  auto take = [&](int n) { int r = o; o += pad4(n); return r; };
  const int NbP = a2.NbP, EP = a2.EP;
  s.state = take((4 + L) * NbP);
  s.nxt = take(2 * 3 * NbP);
  s.adj4 = take(4 * NbP);
  s.a4 = take(4 * NbP);
  s.gout = take(2 * NbP);
  s.cst = take(6 * NbP);
  s.deg = take(NbP);
  s.ones_b = take(NbP);
  s.ones_l = take(EP);
  s.zrow = take(EP);
  s.featp = take(5 * EP);
  s.w = take(wstep);
  s.ws = take(wstep);
  s.act = take(3 * H * NbP);
  s.act_l = take(H * EP);
  // one region, two lives: the physics adjoint's per-step arrays (gdP, trig, adjD, lineg) and, after the barrier that
  // ends the physics, the two item-major blocks that carry adjoints between the bus and the line tiles
  {
    const int phys = 5 * NbP + 5 * EP, mlp = NbP * H + EP * H;
    const int u = take(phys > mlp ? phys : mlp);
    s.gdP = u; s.trig = u + NbP; s.adjD = u + 4 * NbP; s.lineg = u + 5 * NbP;
    s.adjA = u; s.d1l = u + NbP * H;
  }
  s.scratch = take(nwarps * kB3Scr);
  s.red = take(2 * 4 * 32);
  s.topo = take((make_bwd3_topo(N, E).total + 1) / 2);
  s.mbar = take(2 * 8);
  s.total = o;
  return s;
}

// accumulator layout of this kernel: like gns_backward2 (tile of 96) except the output layer of the m net, which is
// computed with the adjoint state as the hidden-side operand: [2 mt][2 nt][32 lanes][4]
__host__ __device__ constexpr FragLayout make_frag_layout3(int L, int H) {
  FragLayout f{};
  int o = 0;
  f.w2l = o; o += kFragTile2 * frag_tiles(H + 1);
  f.w1f = o; o += kFragTile2 * frag_tiles(5);
  f.w1m = o; o += kFragTile2 * frag_tiles(L + 1);
  f.out = o; o += 4 * kB3Tile;
  f.w2 = o; o += kFragTile2 * frag_tiles(H + 1);
  f.w1 = o; o += kFragTile2 * frag_tiles(4 + L + H + 2);
  f.net = o;
  f.step = (3 * o + 31) & ~31;
  return f;
}
// cell (adjoint-state column r' = 4 + i, hidden column j or H = bias) of the m net's output-layer gradient
__host__ __device__ constexpr int frag3_out_index(int rp, int j) {
  return (((rp / 16) * 2 + j / 8) * 32 + ((rp % 16) / 2) * 4 + (j % 8) / 2) * 4 + (rp % 2) * 2 + (j % 2);
}

struct Bwd3Args {
  const float* params;
  const float* buses; const float* lines; const float* gens;
  const float* ck2; const float* pglob; const float* act;
  const float* grad_total; const float* grad_last; const float* grad_v; const float* grad_theta;
  float* gacc;              // [ctas * nwarps][K][FragLayout3.step]  (acc_shared: [ctas][K][step], the warps' reductions race)
  int acc_shared;
  const uint16_t* topo;
  long long S;
  int N, Ns, E, Gn, K;
  Act2Layout a2;
  Bwd3Smem sm;
  TopoOffsets to;
  float wk[kMaxK];
};

// ---- fragment helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_small(float x) {
  return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
}
// mma.sync without `volatile`: the products of a tile are pure data flow, so ptxas may interleave the independent
// accumulator chains of neighbouring calls (a chain of dependent HMMAs waits ~20 cycles per link)
__device__ __forceinline__ void mma_nv(float (&d)[4], const uint32_t (&a)[4], float b0, float b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
// c += A (big / small quads) x B (big / small pairs), 3-term TF32
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ab)[4], const uint32_t (&as)[4], float2 bb, float2 bs) {
  mma_nv(c, as, bb.x, bb.y);
  mma_nv(c, ab, bb.x, bb.y);
  mma_nv(c, ab, bs.x, bs.y);
}
// accumulator fragment (rows g, g+8; columns col, col+1) -> A quad of the k step that covers these columns,
// each element multiplied by the LeakyReLU slope of its (item, column) bit
__device__ __forceinline__ void frag_to_quad(const float (&c)[4], uint32_t wlo, uint32_t whi, int col, int shift, bool masked,
                                             uint32_t (&ab)[4], uint32_t (&as)[4], float (&q)[4]) {
  if (masked) {
    q[0] = c[0] * slope_of(wlo, shift + col); q[2] = c[1] * slope_of(wlo, shift + col + 1);
    q[1] = c[2] * slope_of(whi, shift + col); q[3] = c[3] * slope_of(whi, shift + col + 1);
  } else {
    q[0] = c[0]; q[2] = c[1]; q[1] = c[2]; q[3] = c[3];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { ab[i] = __float_as_uint(q[i]); as[i] = tf32_small(q[i]); }
}
// A-order quad -> item-major tile [16][stride] (columns col, col+1 of items g and g+8)
__device__ __forceinline__ void store_quad(float* tile, int stride, int g, int col, const float (&q)[4]) {
  *reinterpret_cast<float2*>(tile + g * stride + col) = make_float2(q[0], q[2]);
  *reinterpret_cast<float2*>(tile + (g + 8) * stride + col) = make_float2(q[1], q[3]);
}

// c[nt] += A x W^T for weight rows [row0, row0 + nrows) mapped to output columns colshift + r:
// B(k = hidden o, n = output column) = w[(row0 + n - colshift) * HP + o], a 64-bit load per k step (o = 8ks + 2t, +1)
template <int KS, int NT, int HP>
__device__ __forceinline__ void mma_rows(float (&c)[NT][4], const uint32_t (&ab)[KS][4], const uint32_t (&as)[KS][4],
                                         const float* wbig, const float* wsmall, int row0, int nrows, int colshift, int g, int t) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    float2 bb[NT], bs[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int r = 8 * nt + g - colshift;
      const bool rv = r >= 0 && r < nrows;
      const int off = (row0 + (rv ? r : 0)) * HP + 2 * t;
      bb[nt] = make_float2(0.f, 0.f); bs[nt] = bb[nt];
      if (rv && 8 * ks + 2 * t < HP) {
        bb[nt] = *reinterpret_cast<const float2*>(wbig + off + 8 * ks);
        bs[nt] = *reinterpret_cast<const float2*>(wsmall + off + 8 * ks);
      }
    }
    // term-major: consecutive HMMAs go to different accumulators
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_nv(c[nt], as[ks], bb[nt].x, bb[nt].y);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_nv(c[nt], ab[ks], bb[nt].x, bb[nt].y);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_nv(c[nt], ab[ks], bs[nt].x, bs[nt].y);
  }
}

// One k16 chunk of a weight-gradient call on the 16 items of a tile (see cons_call of gns_backward2.cuh):
// hid = item-major [16][hs] tile (columns 2g, 2g+1 per lane; C = 11 adds the constant-1 column), row(r) = wide rows
// at the tile's first item.  Accumulates into acc.
template <int C, int R, class RowFn>
__device__ __forceinline__ void dw_chunk(float (&acc)[(R + 7) / 8][4], const float* hid, int hs, RowFn row, const float* zrow,
                                         const float* ones16 /* [16] 1 where the item exists (C == 11) */) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  constexpr int NT = (R + 7) / 8;
  uint32_t ab[2][4], as[2][4];
#pragma unroll
  for (int st = 0; st < 2; ++st) {
    float2 lo = make_float2(0.f, 0.f), hi = lo;
    const int i0 = 4 * t + 2 * st;
    if (g < 5) {
      lo = *reinterpret_cast<const float2*>(hid + i0 * hs + 2 * g);
      hi = *reinterpret_cast<const float2*>(hid + (i0 + 1) * hs + 2 * g);
    } else if (C == 11 && g == 5) {
      lo.x = ones16[i0]; hi.x = ones16[i0 + 1];
    }
    ab[st][0] = __float_as_uint(lo.x); ab[st][1] = __float_as_uint(lo.y);
    ab[st][2] = __float_as_uint(hi.x); ab[st][3] = __float_as_uint(hi.y);
#pragma unroll
    for (int i = 0; i < 4; ++i) as[st][i] = tf32_small(__uint_as_float(ab[st][i]));
  }
  float4 b[NT], bs[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int r = nt * 8 + g;
    const float* rp = (((nt * 8 + 8 <= R) || (r < R)) ? row(r) : zrow) + 4 * t;
    b[nt] = *reinterpret_cast<const float4*>(rp);
    bs[nt] = make_float4(__uint_as_float(tf32_small(b[nt].x)), __uint_as_float(tf32_small(b[nt].y)),
                         __uint_as_float(tf32_small(b[nt].z)), __uint_as_float(tf32_small(b[nt].w)));
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_nv(acc[nt], as[0], b[nt].x, b[nt].y);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_nv(acc[nt], ab[0], b[nt].x, b[nt].y);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_nv(acc[nt], ab[0], bs[nt].x, bs[nt].y);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_nv(acc[nt], as[1], b[nt].z, b[nt].w);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_nv(acc[nt], ab[1], b[nt].z, b[nt].w);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) mma_nv(acc[nt], ab[1], bs[nt].z, bs[nt].w);
}
template <int NT>
__device__ __forceinline__ void dw_flush(float (&acc)[NT][4], float* __restrict__ gfrag) {
  const int lane = threadIdx.x & 31;
  if (lane < 24) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) red_add_v4(gfrag + nt * kFragTile2 + lane * 4, acc[nt][0], acc[nt][1], acc[nt][2], acc[nt][3]);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
}

// BT / LT: bus / line tiles per warp (the host picks the warp count so that they suffice); TB, MINB: launch bounds
template <int L, int H, int BT, int LT, int TB, int MINB>
__global__ void __launch_bounds__(TB, MINB) gns_backward3_kernel(const Bwd3Args a) {
  static_assert(H == 10 && (L == 10 || L == 20), "instantiated for multiple phi, hidden 10, latent 10 / 20");
  constexpr bool MULTI = true;
  constexpr WLayout W = make_wlayout(L, H, MULTI);
  constexpr FragLayout FL = make_frag_layout3(L, H);
  constexpr int HP = pad4(H);
  constexpr int SC = 4 + L;                 // adjoint-state columns: v, theta, dP, dQ (this step's additions) | m
  constexpr int SNT = (SC + 7) / 8;         // their n8 tiles
  constexpr int SW = 24;                    // item-major stride of the adjoint-state tile
  static_assert(SC <= SW, "adjoint-state tile");

  extern __shared__ __align__(16) float smem[];
  const int N = a.N, E = a.E, Gn = a.Gn, K = a.K;
  const int NbP = a.a2.NbP, EP = a.a2.EP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, nwarps = T >> 5;
  const int g = lane >> 2, t = lane & 3;
  float* const s_state = smem + a.sm.state;
  float* const s_nxt = smem + a.sm.nxt;
  float* const s_adj4 = smem + a.sm.adj4;
  float* const s_a4 = smem + a.sm.a4;
  float* const s_gout = smem + a.sm.gout;
  float* const s_gdP = smem + a.sm.gdP;
  float* const s_cst = smem + a.sm.cst;
  float* const s_trig = smem + a.sm.trig;
  float* const s_adjD = smem + a.sm.adjD;
  float* const s_deg = smem + a.sm.deg;
  float* const s_ones_b = smem + a.sm.ones_b;
  float* const s_ones_l = smem + a.sm.ones_l;
  const float* const s_zrow = smem + a.sm.zrow;
  float* const s_lineg = smem + a.sm.lineg;
  float* const s_featp = smem + a.sm.featp;
  float* const s_w = smem + a.sm.w;
  float* const s_ws = smem + a.sm.ws;
  float* const s_h2L = smem + a.sm.act;                 // item-major [NbP][H]
  float* const s_h1L = smem + a.sm.act + H * NbP;       // rows
  float* const s_A = smem + a.sm.act + 2 * H * NbP;     // rows
  float* const s_h1n = smem + a.sm.act_l;               // rows [H][EP]
  float* const s_adjA = smem + a.sm.adjA;
  float* const s_d1l = smem + a.sm.d1l;
  float* const scr = smem + a.sm.scratch + warp * kB3Scr;
  float* const scr_st = scr;                            // [16][SW]   (dead once the output-layer gradient is done)
  float* const scr_h0 = scr;                            // [16][H]
  float* const scr_h1 = scr_h0 + 16 * H;                // [16][H]
  float* const s_red = smem + a.sm.red;
  uint16_t* const s_topo = reinterpret_cast<uint16_t*>(smem + a.sm.topo);
  uint64_t* const s_bar = reinterpret_cast<uint64_t*>(smem + a.sm.mbar);
  enum { BAR_STATE = 0, BAR_NXT = 1 /* 2 */, BAR_W = 3, BAR_ACT = 4 /* h2L, h1L, A, h1n */ };
  const Bwd3Topo TP = make_bwd3_topo(N, E);
  const uint16_t* const t_fa = s_topo + TP.fa;
  const uint16_t* const t_ta = s_topo + TP.ta;
  const uint16_t* const t_fr = s_topo + TP.fr;
  const uint16_t* const t_tr = s_topo + TP.tr;
  const uint16_t* const t_ini = s_topo + TP.in_ids;
  const uint16_t* const t_inp = s_topo + TP.in_pos;
  const uint16_t* const t_outi = s_topo + TP.out_ids;
  const uint16_t* const t_cslot = s_topo + TP.col_slot;
  const uint16_t* const t_cit = s_topo + TP.col_it;
  const uint16_t* const t_cbr = s_topo + TP.col_brank;
  const uint16_t* const t_rcols = s_topo + TP.rin_cols;
  const uint16_t* const t_rinb = s_topo + TP.rin_b;

  // ---- one-time setup ----
  for (int i = tid; i < a.sm.total; i += T) smem[i] = 0.f;
  __syncthreads();
  {
    const int src[11] = {a.to.fa, a.to.ta, a.to.fr, a.to.tr, a.to.in_ids, a.to.in_pos, a.to.out_ids, a.to.col_slot, a.to.col_it,
                         a.to.col_brank, a.to.rin_cols};
    const int Ep = pad4(E);
    for (int i = tid; i < 11 * E; i += T) {
      const int w = i / E, e = i - w * E;
      s_topo[w * Ep + e] = a.topo[src[w] + e];
    }
    for (int i = tid; i <= N; i += T) s_topo[TP.rin_b + i] = a.topo[a.to.rin_b + i];
    for (int i = tid; i < N; i += T) s_ones_b[i] = 1.f;
    for (int i = tid; i < E; i += T) s_ones_l[i] = 1.f;
    for (int s = tid; s < a.Ns; s += T)
      if ((int)a.topo[a.to.prim_of + s] == s)
        s_deg[a.topo[a.to.brank + s]] = (float)((int)a.topo[a.to.in_fe + s] - (int)a.topo[a.to.in_b + s]);
    if (tid == 0) {
      for (int i = 0; i < 8; ++i) mbar_init(s_bar + i, 1);
      fence_proxy_async();
    }
  }
  __syncthreads();

  // ---- per-thread bus (physics adjoint: one thread per bus rank) ----
  const int r_me = tid;                                  // bus rank
  const bool bus_on = r_me < N;
  int ext = 0, pslot = 0, e_in0 = 0, e_in1 = 0, e_out0 = 0, e_out1 = 0;
  bool is_gen = false;
  if (bus_on) {
    ext = a.topo[a.to.ext_rank + r_me];
    pslot = a.topo[a.to.rank_of + ext];
    e_in0 = a.topo[a.to.in_b + pslot]; e_in1 = a.topo[a.to.in_fe + pslot];
    e_out0 = a.topo[a.to.out_b + pslot]; e_out1 = a.topo[a.to.out_e + pslot];
    is_gen = a.topo[a.to.gen_e + pslot] > a.topo[a.to.gen_b + pslot];
  }
  // ---- tiles of this warp: bus tiles w, w + nwarps; line tiles w, w + nwarps, ... ----
  const int nbt = (N + 15) / 16, nlt = (E + 15) / 16;
  int ps_lo[BT], ps_hi[BT];               // primary slots of this lane's items (g, g+8) in its bus tiles: slope-word columns
#pragma unroll
  for (int j = 0; j < BT; ++j) {
    const int rlo = (warp + j * nwarps) * 16 + g, rhi = rlo + 8;
    ps_lo[j] = rlo < N ? (int)a.topo[a.to.rank_of + (int)a.topo[a.to.ext_rank + rlo]] : -1;
    ps_hi[j] = rhi < N ? (int)a.topo[a.to.rank_of + (int)a.topo[a.to.ext_rank + rhi]] : -1;
  }
  int red_parity = 0;
  uint32_t ph_state = 0, ph_nxt = 0, ph_w = 0, ph_act = 0;
  int nxt_buf = 0;
  float* const gacc_w = a.gacc + (a.acc_shared ? (size_t)blockIdx.x : (size_t)blockIdx.x * nwarps + warp) * ((size_t)K * FL.step);

  auto block_sum = [&](float (&x)[3], int nv) {          // deterministic sum over the CTA (barrier inside)
    float* buf = s_red + red_parity * 128;
    red_parity ^= 1;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
      for (int v = 0; v < 3; ++v) x[v] += __shfl_xor_sync(0xffffffffu, x[v], off);
    if (lane == 0) { for (int v = 0; v < nv; ++v) buf[warp * 4 + v] = x[v]; }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < 3; ++v) x[v] = 0.f;
    for (int w = 0; w < nwarps; ++w)
      for (int v = 0; v < nv; ++v) x[v] += buf[w * 4 + v];
  };
  auto issue_state = [&](long long grid, int k) {
    fence_proxy_async();
    mbar_expect_tx(s_bar + BAR_STATE, (uint32_t)((4 + L) * NbP * 4));
    bulk_g2s(s_state, a.ck2 + ((size_t)grid * (K + 1) + k) * (size_t)a.a2.state, (uint32_t)((4 + L) * NbP * 4), s_bar + BAR_STATE);
  };
  auto issue_nxt = [&](long long grid, int kidx, int buf) {
    mbar_expect_tx(s_bar + BAR_NXT + buf, 3 * NbP * 4);
    bulk_g2s(s_nxt + buf * 3 * NbP, a.ck2 + ((size_t)grid * (K + 1) + kidx) * (size_t)a.a2.state, 3 * NbP * 4, s_bar + BAR_NXT + buf);
  };
  auto issue_w = [&](int k) {
    fence_proxy_async();
    mbar_expect_tx(s_bar + BAR_W, W.wstep * 4);
    bulk_g2s(s_w, a.params + (size_t)k * W.wstep, W.wstep * 4, s_bar + BAR_W);
  };
  auto issue_act_bus = [&](long long grid, int k, int q) {       // h2L, h1L, A of pair q (one barrier, three copies)
    fence_proxy_async();
    const float* base = a.act + ((size_t)grid * K + k) * (size_t)a.a2.step;
    mbar_expect_tx(s_bar + BAR_ACT, (uint32_t)(3 * H * NbP * 4));
    bulk_g2s(s_h2L, base + a.a2.h2L[q], H * NbP * 4, s_bar + BAR_ACT);
    bulk_g2s(s_h1L, base + a.a2.h1L[q], H * NbP * 4, s_bar + BAR_ACT);
    bulk_g2s(s_A, base + a.a2.A[q], H * NbP * 4, s_bar + BAR_ACT);
  };
  auto issue_act_line = [&](long long grid, int k, int q) {
    fence_proxy_async();
    const float* base = a.act + ((size_t)grid * K + k) * (size_t)a.a2.step;
    mbar_expect_tx(s_bar + BAR_ACT + 1, (uint32_t)(H * EP * 4));
    bulk_g2s(s_h1n, base + a.a2.h1line[q], H * EP * 4, s_bar + BAR_ACT + 1);
  };
  const int first_grid = blockIdx.x, grid_step = gridDim.x;
  if (tid == 0 && first_grid < a.S) {
    issue_w(K - 1);
    issue_nxt(first_grid, K, 0);
    issue_state(first_grid, K - 1);
    issue_act_bus(first_grid, K - 1, 2);
    issue_act_line(first_grid, K - 1, 2);
  }

  // adjoint of the latent (and this step's additions to adj v, theta, dP, dQ in columns 0..3) of this warp's two bus
  // tiles, as accumulator fragments: columns 8 nt + 2t, +1 of items g, g+8
  float ast[BT][SNT][4];

  for (long long grid = first_grid; grid < a.S; grid += grid_step) {
    // ---------------- per-grid constants ----------------
    const float gtot = a.grad_total[grid];
    const float glast = a.grad_last ? a.grad_last[grid] : 0.f;
    float s3[3] = {0.f, 0.f, 0.f};
    for (int j = tid; j < Gn; j += T) {
      const float* gr = a.gens + ((size_t)grid * Gn + j) * 7;
      s3[0] += __ldg(gr + 3); s3[1] += __ldg(gr + 2); s3[2] += __ldg(gr + 1);
    }
    if (bus_on) {
      s_cst[0 * NbP + r_me] = __ldg(a.buses + ((size_t)grid * N + ext) * 6 + 4);
      float lo = 0.f, hi = 0.f;
      const int j0 = a.topo[a.to.gen_b + pslot], j1 = a.topo[a.to.gen_e + pslot];
      for (int j = j0; j < j1; ++j) {
        const float* gr = a.gens + ((size_t)grid * Gn + (int)a.topo[a.to.gen_ids + j]) * 7;
        const float Pmax = __ldg(gr + 1), Pmin = __ldg(gr + 2), Pset = __ldg(gr + 3);
        lo += 2.f * (Pset - Pmin); hi += 2.f * (Pmax - Pset);
      }
      s_cst[1 * NbP + r_me] = lo; s_cst[2 * NbP + r_me] = hi;
      s_adj4[0 * NbP + r_me] = a.grad_v ? a.grad_v[(size_t)grid * N + ext] : 0.f;
      s_adj4[1 * NbP + r_me] = a.grad_theta ? a.grad_theta[(size_t)grid * N + ext] : 0.f;
      s_adj4[2 * NbP + r_me] = 0.f; s_adj4[3 * NbP + r_me] = 0.f;
      const float* lr = a.lines + ((size_t)grid * E + r_me) * 7;       // alias line id = this thread's index (< N <= E)
      const float r = __ldg(lr + 2), x = __ldg(lr + 3);
      s_cst[3 * NbP + r_me] = 1.0f / sqrtf(r * r + x * x);
      s_cst[4 * NbP + r_me] = 1.0f / __ldg(lr + 5);
      s_cst[5 * NbP + r_me] = __ldg(lr + 6);
    }
    for (int i = tid; i < E; i += T) {
      const float* lr = a.lines + ((size_t)grid * E + (int)t_ini[i]) * 7 + 2;
      const int col = t_inp[i];
#pragma unroll
      for (int c = 0; c < 5; ++c) s_featp[c * EP + col] = __ldg(lr + c);
    }
#pragma unroll
    for (int j = 0; j < BT; ++j)
#pragma unroll
      for (int nt = 0; nt < SNT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) ast[j][nt][i] = 0.f;
    block_sum(s3, 3);
    const float sPset = s3[0], sPmin = s3[1], sPmax = s3[2];
    float pglob = __ldg(a.pglob + (size_t)grid * K + (K - 1));

    for (int k = K - 1; k >= 0; --k) {
      if (tid == 0) {
        if (k >= 1) issue_nxt(grid, k, nxt_buf ^ 1);
        else if (grid + grid_step < a.S) issue_nxt(grid + grid_step, K, nxt_buf ^ 1);
      }
      const float pg_now = pglob;
      if (k >= 1) pglob = __ldg(a.pglob + (size_t)grid * K + (k - 1));
      else if (grid + grid_step < a.S) pglob = __ldg(a.pglob + (size_t)(grid + grid_step) * K + (K - 1));

      // ---------------- physics adjoint (one thread per bus / line), as in gns_backward.cuh ----------------
      const bool lo_branch = pg_now < sPset;
      const float lam = lo_branch ? (pg_now - sPmin) / (2.f * (sPset - sPmin))
                                  : (pg_now - 2.f * sPset + sPmax) / (2.f * (sPmax - sPset));
      const bool lo_arm = lam < 0.5f;
      const float coef = (gtot * a.wk[k] + ((k == K - 1) ? glast : 0.f)) * (2.0f / (float)N);
      mbar_wait(s_bar + BAR_NXT + nxt_buf, (ph_nxt >> nxt_buf) & 1u);
      ph_nxt ^= 1u << nxt_buf;
      const float* const nx = s_nxt + nxt_buf * 3 * NbP;
      nxt_buf ^= 1;
      float gdP = 0.f, vpr = 0.f, Gsv = 0.f;
      float part[3] = {0.f, 0.f, 0.f};
      if (bus_on) {
        gdP = s_adj4[2 * NbP + r_me] + coef * nx[2 * NbP + r_me];
        s_gdP[r_me] = gdP;
        part[0] = gdP * (lo_arm ? s_cst[1 * NbP + r_me] : s_cst[2 * NbP + r_me]);
        vpr = nx[r_me];
        Gsv = s_cst[r_me];
        const float d = nx[NbP + (int)t_fr[r_me]] - nx[NbP + (int)t_tr[r_me]];     // alias line r_me
        float sd, cd;
        fast_sincos(d, sd, cd);
        s_trig[r_me] = d; s_trig[NbP + r_me] = sd; s_trig[2 * NbP + r_me] = cd;
        s_a4[0 * NbP + r_me] = 0.f; s_a4[1 * NbP + r_me] = 0.f; s_a4[2 * NbP + r_me] = 0.f; s_a4[3 * NbP + r_me] = 0.f;
      }
      block_sum(part, 1);
      const float adj_pg = part[0] / (lo_branch ? 2.f * (sPset - sPmin) : 2.f * (sPmax - sPset));
#pragma unroll 1
      for (int e = tid; e < E; e += T) {
        const int fi = t_fr[e], ti = t_tr[e], fa = t_fa[e], ta = t_ta[e];
        const float vf = nx[fi], vt = nx[ti];
        const float thf = nx[NbP + fi], tht = nx[NbP + ti];
        const float g_pf = s_gdP[ti];
        const float g_pt = s_gdP[fi];
        const float Yf = s_cst[3 * NbP + fa], itf = s_cst[4 * NbP + fa], shf = s_cst[5 * NbP + fa];
        const float Df = s_trig[fa], sDf = s_trig[NbP + fa], cDf = s_trig[2 * NbP + fa];
        const float Yt = s_cst[3 * NbP + ta], itt = s_cst[4 * NbP + ta], sht = s_cst[5 * NbP + ta];
        const float DB = s_trig[ta], sDB = s_trig[NbP + ta], cDB = s_trig[2 * NbP + ta];
        const float a1 = thf - tht - Df - shf;
        const float a2 = tht - thf - Df + shf;
        const float a3 = tht - thf + DB - sht;
        float s1, c1, s2, c2, s3v, c3;
        fast_sincos(a1, s1, c1);
        fast_sincos(a2, s2, c2);
        fast_sincos(a3, s3v, c3);
        const float yft = Yf * itf, yftt = Yf * (itf * itf), ytt = Yt * itt;
        const float t1 = vf * vt * yft, u1 = vt * vf * ytt;
        const float sDt = -sDB;
        const float ss = s1 + s2;
        const float inner = t1 * ss + vf * yftt * sDf + vt * vt * Yf * sDf;
        const float g_in = adj_pg * ((inner > 0.f) ? 1.f : ((inner < 0.f) ? -1.f : 0.f));
        const float gvf = g_in * (vt * yft * ss + yftt * sDf) + g_pf * (vt * yft * s1 + 2.f * vf * yftt * sDf) +
                          g_pt * (vt * ytt * s3v);
        const float gvt = g_in * (vf * yft * ss + 2.f * vt * Yf * sDf) + g_pf * (vf * yft * s1) +
                          g_pt * (vf * ytt * s3v + 2.f * vt * Yt * sDt);
        const float G1 = (g_in + g_pf) * t1 * c1, G2 = g_in * t1 * c2, G3 = g_pt * u1 * c3;
        s_lineg[0 * EP + e] = gvf;
        s_lineg[1 * EP + e] = gvt;
        s_lineg[2 * EP + e] = G1 - G2 - G3;
        s_lineg[3 * EP + e] = -G1 - G2 + (g_in * (vf * yftt + vt * vt * Yf) + g_pf * vf * vf * yftt) * cDf;
        s_lineg[4 * EP + e] = G3 - g_pt * vt * vt * Yt * cDB;
      }
      __syncthreads();
      float adjv = 0.f, adjth = 0.f;
      if (bus_on) {
        float sv = 0.f, sth = 0.f, sD = 0.f;
        for (int e = e_out0; e < e_out1; ++e) {
          const int eo = t_outi[e];
          sv += s_lineg[0 * EP + eo]; sth += s_lineg[2 * EP + eo]; sD += s_lineg[3 * EP + eo];
        }
        for (int e = e_in0; e < e_in1; ++e) {
          const int eo = t_ini[e];
          sv += s_lineg[1 * EP + eo]; sth -= s_lineg[2 * EP + eo]; sD += s_lineg[4 * EP + eo];
        }
        adjv = s_adj4[0 * NbP + r_me] + sv + gdP * (-2.f * Gsv * vpr) + adj_pg * (2.f * vpr * Gsv);
        adjth = s_adj4[1 * NbP + r_me] + sth;
        s_adjD[ext] = sD;
      }
      __syncthreads();
      if (bus_on) {
        for (int e = e_out0; e < e_out1; ++e) { const int l = t_outi[e]; if (l < N) adjth += s_adjD[l]; }
        for (int e = e_in0; e < e_in1; ++e) { const int l = t_ini[e]; if (l < N) adjth -= s_adjD[l]; }
        s_gout[0 * NbP + r_me] = is_gen ? 0.f : adjv;
        s_gout[1 * NbP + r_me] = adjth;
        s_adj4[0 * NbP + r_me] = adjv;      // the step's MLP additions arrive at the end of the step
        s_adj4[1 * NbP + r_me] = adjth;
      }
      // ---------------- weights of this step: wait, split once ----------------
      mbar_wait(s_bar + BAR_W, ph_w);
      ph_w ^= 1u;
      for (int i = tid; i < W.wstep; i += T) s_ws[i] = __uint_as_float(tf32_small(s_w[i]));
      mbar_wait(s_bar + BAR_STATE, ph_state);
      ph_state ^= 1u;
      __syncthreads();           // s_gout, s_ws, s_a4 zeros visible

      float* const gk = gacc_w + (size_t)k * FL.step;
#pragma unroll 1
      for (int qq = 0; qq < 3; ++qq) {
        const int q = (qq == 0) ? 2 : qq - 1;      // m net first: its output layer reads adj m' before anyone adds to it
        const float* const wphi = s_w + q * W.phi_size;
        const float* const wphis = s_ws + q * W.phi_size;
        const int ln0 = W.off_ln[0] + q * W.ln_size_s;
        const float* const wln = s_w + ln0;
        const float* const wlns = s_ws + ln0;
        const float* const wmf = s_w + W.off_mf[0] + q * W.mf_size;
        const float* const wmfs = s_ws + W.off_mf[0] + q * W.mf_size;
        float* const gln = gk + q * FL.net;
        float* const gphi = gk + q * FL.net;
        // slope words of this lane's items (bus tiles: bits 0..9 h1L, 10..19 h2L; line tiles: h1, h2 of the phi net):
        // fetched here, first used after the activation blocks have landed
        uint32_t mw[BT][2], lw[LT][2];
        {
          const uint32_t* const mrow =
              reinterpret_cast<const uint32_t*>(a.act + ((size_t)grid * K + k) * (size_t)a.a2.step + a.a2.mask[q]);
#pragma unroll
          for (int j = 0; j < BT; ++j) {
            mw[j][0] = ps_lo[j] >= 0 ? __ldg(mrow + ps_lo[j]) : 0u;
            mw[j][1] = ps_hi[j] >= 0 ? __ldg(mrow + ps_hi[j]) : 0u;
          }
          const uint32_t* const lmask = mrow + a.a2.NsM;          // row 1 + it
#pragma unroll
          for (int j = 0; j < LT; ++j) {
            const int clo = (warp + j * nwarps) * 16 + g, chi = clo + 8;
            lw[j][0] = clo < E ? __ldg(lmask + (int)t_cit[clo] * a.a2.NsM + (int)t_cslot[clo]) : 0u;
            lw[j][1] = chi < E ? __ldg(lmask + (int)t_cit[chi] * a.a2.NsM + (int)t_cslot[chi]) : 0u;
          }
        }
        mbar_wait(s_bar + BAR_ACT, ph_act & 1u);
        ph_act ^= 1u;

        // ======== phase A: L net of the pair on this warp's bus tiles ========
#pragma unroll
        for (int j = 0; j < BT; ++j) {
          const int tile = warp + j * nwarps;
          if (tile < nbt) {
            const int i0 = tile * 16;                        // first bus rank of the tile
            const int rlo = i0 + g, rhi = i0 + g + 8;
            const uint32_t wlo = mw[j][0], whi = mw[j][1];
            // ---- output layer: dh2 = g_out x Wout, and its weight gradient ----
            float c2[2][4];
            float qd[4];
            if (q == 2) {
#pragma unroll
              for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) c2[nt][i] = 0.f;
              {
                // adjoint state tile -> item-major scratch (hidden-side operand of the output-layer gradient) and A quads
                uint32_t ab[SNT][4], as[SNT][4];
                __syncwarp();
#pragma unroll
                for (int nt = 0; nt < SNT; ++nt) {
                  frag_to_quad(ast[j][nt], 0u, 0u, 8 * nt + 2 * t, 0, false, ab[nt], as[nt], qd);
                  store_quad(scr_st, SW, g, 8 * nt + 2 * t, qd);
                }
                // dh2[item][o] = sum_r' ast[item][r'] Wout[r' - 4][o]: B(k = r', n = o) = wln[ln_wo + (r' - 4) HP + o]
#pragma unroll
                for (int ks = 0; ks < SNT; ++ks) {
                  float2 bb[2], bs[2];
#pragma unroll
                  for (int nt = 0; nt < 2; ++nt) {
                    const int o = 8 * nt + g;
                    const int r0 = 8 * ks + 2 * t - 4;      // weight rows of k slots t, t+4
                    bb[nt] = make_float2(0.f, 0.f); bs[nt] = bb[nt];
                    if (o < HP) {
                      if (r0 >= 0 && r0 < L) { bb[nt].x = wln[W.ln_wo + r0 * HP + o]; bs[nt].x = wlns[W.ln_wo + r0 * HP + o]; }
                      if (r0 + 1 >= 0 && r0 + 1 < L) { bb[nt].y = wln[W.ln_wo + (r0 + 1) * HP + o]; bs[nt].y = wlns[W.ln_wo + (r0 + 1) * HP + o]; }
                    }
                  }
#pragma unroll
                  for (int nt = 0; nt < 2; ++nt) mma_nv(c2[nt], as[ks], bb[nt].x, bb[nt].y);
#pragma unroll
                  for (int nt = 0; nt < 2; ++nt) mma_nv(c2[nt], ab[ks], bb[nt].x, bb[nt].y);
#pragma unroll
                  for (int nt = 0; nt < 2; ++nt) mma_nv(c2[nt], ab[ks], bs[nt].x, bs[nt].y);
                }
              }
              __syncwarp();
              // dWout[i][j] += adjm'[i] h2[j], dbout[i] += adjm'[i]: A = adjoint state tile (columns 2g, 2g+1 of 16 per
              // M tile), B = [h2L | 1] item-major (two scalar loads per k step)
              float accOutM[2][2][4];
#pragma unroll
              for (int i = 0; i < 4; ++i) { accOutM[0][0][i] = 0.f; accOutM[0][1][i] = 0.f; accOutM[1][0][i] = 0.f; accOutM[1][1][i] = 0.f; }
#pragma unroll
              for (int st = 0; st < 2; ++st) {
                const int it0 = 4 * t + 2 * st;
                uint32_t ab2[2][4], as2[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                  float2 lo = make_float2(0.f, 0.f), hi = lo;
                  if (16 * mt + 2 * g < SW) {
                    lo = *reinterpret_cast<const float2*>(scr_st + it0 * SW + 16 * mt + 2 * g);
                    hi = *reinterpret_cast<const float2*>(scr_st + (it0 + 1) * SW + 16 * mt + 2 * g);
                  }
                  ab2[mt][0] = __float_as_uint(lo.x); ab2[mt][1] = __float_as_uint(lo.y);
                  ab2[mt][2] = __float_as_uint(hi.x); ab2[mt][3] = __float_as_uint(hi.y);
#pragma unroll
                  for (int i = 0; i < 4; ++i) as2[mt][i] = tf32_small(__uint_as_float(ab2[mt][i]));
                }
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                  const int n = 8 * nt + g;
                  float b0 = 0.f, b1 = 0.f;
                  if (n < H) { b0 = s_h2L[(i0 + it0) * H + n]; b1 = s_h2L[(i0 + it0 + 1) * H + n]; }
                  else if (n == H) { b0 = s_ones_b[i0 + it0]; b1 = s_ones_b[i0 + it0 + 1]; }
                  const float2 bb = make_float2(b0, b1);
                  const float2 bs = make_float2(__uint_as_float(tf32_small(b0)), __uint_as_float(tf32_small(b1)));
#pragma unroll
                  for (int mt = 0; mt < 2; ++mt) mma3(accOutM[mt][nt], ab2[mt], as2[mt], bb, bs);
                }
              }
#pragma unroll
              for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
                  red_add_v4(gln + FL.out + ((mt * 2 + nt) * 32 + lane) * 4, accOutM[mt][nt][0], accOutM[mt][nt][1], accOutM[mt][nt][2],
                             accOutM[mt][nt][3]);
            } else {
              // scalar net: outer product of the output adjoint with the output row
              const float glo = s_gout[q * NbP + rlo], ghi = s_gout[q * NbP + rhi];
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) {
                const int o = 8 * nt + 2 * t;
                const float w0 = o < HP ? wln[W.ln_wo + o] : 0.f, w1 = o + 1 < HP ? wln[W.ln_wo + o + 1] : 0.f;
                c2[nt][0] = glo * w0; c2[nt][1] = glo * w1; c2[nt][2] = ghi * w0; c2[nt][3] = ghi * w1;
              }
              float accOutS[1][4] = {{0.f, 0.f, 0.f, 0.f}};
              dw_chunk<H + 1, 1>(accOutS, s_h2L + i0 * H, H, [&](int) { return s_gout + q * NbP + i0; }, s_zrow + i0, s_ones_b + i0);
              dw_flush<1>(accOutS, gln + FL.out);
            }
            // ---- d2 = dh2 * slope(h2L); its item-major copy is the hidden-side operand of dW2 ----
            uint32_t db[2][4], ds[2][4];
            __syncwarp();              // the adjoint-state tile in the scratch is dead: the hidden tiles take its place
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              frag_to_quad(c2[nt], wlo, whi, 8 * nt + 2 * t, H, 8 * nt + 2 * t < H, db[nt], ds[nt], qd);
              if (8 * nt + 2 * t < H) store_quad(scr_h0, H, g, 8 * nt + 2 * t, qd);
            }
            // ---- second layer: d1 = (d2 x W2) * slope(h1L) ----
            float c1[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int i = 0; i < 4; ++i) c1[nt][i] = 0.f;
            mma_rows<2, 2, HP>(c1, db, ds, wln + W.ln_w2, wlns + W.ln_w2, 0, H, 0, g, t);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              frag_to_quad(c1[nt], wlo, whi, 8 * nt + 2 * t, 0, 8 * nt + 2 * t < H, db[nt], ds[nt], qd);
              if (8 * nt + 2 * t < H) store_quad(scr_h1, H, g, 8 * nt + 2 * t, qd);
            }
            // ---- first layer dX: adjoint state (+= d1 x W1[:4+L]) and adjoint of the aggregate (d1 x M) ----
            mma_rows<2, SNT, HP>(ast[j], db, ds, wln + W.ln_w1, wlns + W.ln_w1, 0, SC, 0, g, t);
            {
              float cA[2][4];
#pragma unroll
              for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) cA[nt][i] = 0.f;
              mma_rows<2, 2, HP>(cA, db, ds, wmf, wmfs, 0, H, 0, g, t);
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) {
                const int col = 8 * nt + 2 * t;
                if (col < H) {
                  *reinterpret_cast<float2*>(s_adjA + (i0 + g) * H + col) = make_float2(cA[nt][0], cA[nt][1]);
                  *reinterpret_cast<float2*>(s_adjA + (i0 + g + 8) * H + col) = make_float2(cA[nt][2], cA[nt][3]);
                }
              }
            }
            __syncwarp();
            // ---- weight gradients of the hidden layers on this tile (one k16 chunk each), flushed per tile ----
            {
              float accW2[2][4];
#pragma unroll
              for (int i = 0; i < 4; ++i) { accW2[0][i] = 0.f; accW2[1][i] = 0.f; }
              dw_chunk<H, H + 1>(accW2, scr_h0, H, [&](int r) { return (r < H ? s_h1L + r * NbP : s_ones_b) + i0; }, s_zrow + i0, nullptr);
              dw_flush<2>(accW2, gln + FL.w2);
            }
            {
              // W1 rows: 4 + L state columns | H aggregate columns | degree | 1, in two calls to bound the live registers
              constexpr int RW1 = 4 + L + H + 2;
              constexpr int R0 = RW1 > 24 ? 24 : 16;       // first call: whole n tiles
              auto w1row = [&](int r) {
                return (r < 4 + L ? s_state + r * NbP
                                  : (r < 4 + L + H ? s_A + (r - 4 - L) * NbP : (r == 4 + L + H ? s_deg : s_ones_b))) + i0;
              };
              float acc0[R0 / 8][4];
#pragma unroll
              for (int x = 0; x < R0 / 8; ++x)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc0[x][i] = 0.f;
              dw_chunk<H, R0>(acc0, scr_h1, H, w1row, s_zrow + i0, nullptr);
              dw_flush<R0 / 8>(acc0, gln + FL.w1);
              constexpr int R1 = RW1 - R0;
              float acc1[(R1 + 7) / 8][4];
#pragma unroll
              for (int x = 0; x < (R1 + 7) / 8; ++x)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc1[x][i] = 0.f;
              dw_chunk<H, R1>(acc1, scr_h1, H, [&](int r) { return w1row(r + R0); }, s_zrow + i0, nullptr);
              dw_flush<(R1 + 7) / 8>(acc1, gln + FL.w1 + (R0 / 8) * kFragTile2);
            }
          }
        }
        mbar_wait(s_bar + BAR_ACT + 1, (ph_act >> 1) & 1u);     // (ph_act bit 0 was flipped above; bit 1 tracks the line block)
        __syncthreads();                       // adjA complete; everyone is done with h2L / h1L / A of this pair
        ph_act ^= 2u;
        if (tid == 0) {                         // next pair's (or step's, or grid's) L-net activation blocks
          if (qq < 2) issue_act_bus(grid, k, qq);                       // qq -> next q: (qq+1 == 1) ? 0 : 1  == qq
          else if (k >= 1) issue_act_bus(grid, k - 1, 2);
          else if (grid + grid_step < a.S) issue_act_bus(grid + grid_step, K - 1, 2);
        }

        // ======== phase C: phi net of the pair, lines as items ========
        float accW2l[2][4], accW1f[1][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { accW2l[0][i] = 0.f; accW2l[1][i] = 0.f; accW1f[0][i] = 0.f; }
#pragma unroll
        for (int j = 0; j < LT; ++j) {
          const int tile = warp + j * nwarps;
          if (tile >= nlt) break;
          const int c0 = tile * 16;
          const int clo = c0 + g, chi = c0 + g + 8;
          const bool vlo = clo < E, vhi = chi < E;
          const uint32_t wlo = lw[j][0], whi = lw[j][1];
          const float* alo = s_adjA + (vlo ? (int)t_cbr[clo] : 0) * H;
          const float* ahi = s_adjA + (vhi ? (int)t_cbr[chi] : 0) * H;
          uint32_t db[2][4], ds[2][4];
          float qd[4];
          __syncwarp();
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const int col = 8 * nt + 2 * t;
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            if (col < H) {
              const float2 x = *reinterpret_cast<const float2*>(alo + col), y = *reinterpret_cast<const float2*>(ahi + col);
              if (vlo) { c[0] = x.x; c[1] = x.y; }
              if (vhi) { c[2] = y.x; c[3] = y.y; }
            }
            frag_to_quad(c, wlo, whi, col, H, col < H, db[nt], ds[nt], qd);          // d2 = adjA[bus] * slope(h2)
            if (col < H) store_quad(scr_h0, H, g, col, qd);
          }
          float c1[2][4];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) c1[nt][i] = 0.f;
          mma_rows<2, 2, HP>(c1, db, ds, wphi + W.phi_w2, wphis + W.phi_w2, 0, H, 0, g, t);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const int col = 8 * nt + 2 * t;
            frag_to_quad(c1[nt], wlo, whi, col, 0, col < H, db[nt], ds[nt], qd);    // d1 = (d2 x W2) * slope(h1)
            if (col < H) store_quad(s_d1l + c0 * H, H, g, col, qd);
          }
          __syncwarp();
          dw_chunk<H, H + 1>(accW2l, scr_h0, H, [&](int r) { return (r < H ? s_h1n + r * EP : s_ones_l) + c0; }, s_zrow + c0, nullptr);
          dw_chunk<H, 5>(accW1f, s_d1l + c0 * H, H, [&](int r) { return s_featp + r * EP + c0; }, s_zrow + c0, nullptr);
        }
        dw_flush<2>(accW2l, gphi + FL.w2l);
        dw_flush<1>(accW1f, gphi + FL.w1f);
        __syncthreads();                       // d1l complete; everyone is done with the line block
        if (tid == 0) {
          if (qq < 2) issue_act_line(grid, k, qq);
          else if (k >= 1) issue_act_line(grid, k - 1, 2);
          else if (grid + grid_step < a.S) issue_act_line(grid + grid_step, K - 1, 2);
        }

        // ======== phase D: adjP per bus (gather over its in-lines), adj m += adjP x W1m, dW1m ========
#pragma unroll
        for (int j = 0; j < BT; ++j) {
          const int tile = warp + j * nwarps;
          if (tile < nbt) {
            const int i0 = tile * 16;
            const int rlo = i0 + g, rhi = i0 + g + 8;
            const int lb = rlo < N ? (int)t_rinb[rlo] : 0, le = rlo < N ? (int)t_rinb[rlo + 1] : 0;
            const int hb = rhi < N ? (int)t_rinb[rhi] : 0, he = rhi < N ? (int)t_rinb[rhi + 1] : 0;
            uint32_t db[2][4], ds[2][4];
            float qd[4];
            __syncwarp();
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              const int col = 8 * nt + 2 * t;
              float c[4] = {0.f, 0.f, 0.f, 0.f};
              if (col < H) {
                for (int e = lb; e < le; ++e) {
                  const float2 x = *reinterpret_cast<const float2*>(s_d1l + (int)t_rcols[e] * H + col);
                  c[0] += x.x; c[1] += x.y;
                }
                for (int e = hb; e < he; ++e) {
                  const float2 x = *reinterpret_cast<const float2*>(s_d1l + (int)t_rcols[e] * H + col);
                  c[2] += x.x; c[3] += x.y;
                }
              }
              frag_to_quad(c, 0u, 0u, col, 0, false, db[nt], ds[nt], qd);
              if (col < H) store_quad(scr_h0, H, g, col, qd);
            }
            // adjoint state columns 4.. += adjP x W1m
            mma_rows<2, SNT, HP>(ast[j], db, ds, wphi + W.phi_w1m, wphis + W.phi_w1m, 0, L, 4, g, t);
            __syncwarp();
            float accW1m[(L + 1 + 7) / 8][4];
#pragma unroll
            for (int x = 0; x < (L + 1 + 7) / 8; ++x)
#pragma unroll
              for (int i = 0; i < 4; ++i) accW1m[x][i] = 0.f;
            dw_chunk<H, L + 1>(accW1m, scr_h0, H, [&](int r) { return (r < L ? s_state + (4 + r) * NbP : s_ones_b) + i0; }, s_zrow + i0,
                               nullptr);
            dw_flush<(L + 1 + 7) / 8>(accW1m, gphi + FL.w1m);
          }
        }
      }  // pairs

      // ---------------- this step's additions to adj (v, theta, dP, dQ): fragments -> rows; then per bus ----------------
#pragma unroll
      for (int j = 0; j < BT; ++j) {
        const int tile = warp + j * nwarps;
        if (tile < nbt && t < 2) {
          const int i0 = tile * 16;
          s_a4[(2 * t) * NbP + i0 + g] = ast[j][0][0]; s_a4[(2 * t + 1) * NbP + i0 + g] = ast[j][0][1];
          s_a4[(2 * t) * NbP + i0 + g + 8] = ast[j][0][2]; s_a4[(2 * t + 1) * NbP + i0 + g + 8] = ast[j][0][3];
          ast[j][0][0] = 0.f; ast[j][0][1] = 0.f; ast[j][0][2] = 0.f; ast[j][0][3] = 0.f;
        }
      }
      __syncthreads();
      if (bus_on) {
        s_adj4[0 * NbP + r_me] += s_a4[0 * NbP + r_me];
        s_adj4[1 * NbP + r_me] += s_a4[1 * NbP + r_me];
        s_adj4[2 * NbP + r_me] = s_a4[2 * NbP + r_me];
        s_adj4[3 * NbP + r_me] = s_a4[3 * NbP + r_me];
      }
      if (tid == 0) {        // every warp is past its last use of the weights and the state rows
        if (k >= 1) { issue_w(k - 1); issue_state(grid, k - 1); }
        else if (grid + grid_step < a.S) { issue_w(K - 1); issue_state(grid + grid_step, K - 1); }
      }
      __syncthreads();
    }  // k
  }  // grid
}

}  // namespace gns
