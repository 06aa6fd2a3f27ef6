// gns_backward2.cuh — warp-specialised backward kernel for large grids (one grid per CTA at a time).
//
// Same mathematics as gns_backward.cuh (the implicit autograd backward of ref GNS/main.py:288, BPTT over the K
// steps), reorganised around where the instructions went in the first kernel (ncu, profiles/r01h_*):
//
//   PRODUCER warps (threads 0 .. 32 PW): one thread owns TWO adjacent bus slots of the grid, so every dX product of
//     the MLP adjoint is one packed FFMA2 against a warp-uniform weight row (the forward kernel's trick, with the
//     second slot in place of the second grid).  LeakyReLU slopes come from the bit masks the training forward
//     stored (no activation values are read here), the physics adjoint runs on these warps as well.  Every hidden-
//     side vector a weight gradient needs (d2L, d1L, adjP per bus; d2, d1 per line; the output adjoints) is written
//     once to a shared-memory block [item][column] (branch-free: slots without a bus write to a padding item).
//   CONSUMER warps (the other 32 CW threads): the weight gradients  dW^T[w][o] = sum_items wide[w][item] hid[o][item]
//     as mma.sync m16n8k8 TF32 products with the 3-term split of gns_backward.cuh, but over ALL items of the grid per
//     call: a warp accumulates its share of the items in registers and flushes every call with ONE red.global per
//     lane and tile (the first kernel flushed every 32 items).  The wide rows (state, A, h1L, h2L, per-line h1) are
//     bulk-copied (cp.async.bulk + mbarrier) from the forward's per-grid checkpoint blocks straight into shared
//     memory: no register staging, no exposed load latency, no transposition.
//   The two sides meet through mbarrier full / empty pairs per block (static roles, see Bwd2Bar); producers run up
//   to one pair of nets ahead of the consumers.
//
// Block layouts are Act2Layout (gns_common.cuh); gradients go to the fragment-order accumulators of FragLayout, one
// block per CONSUMER warp (deterministic: one owner, program order), summed by reduce_partials_kernel.
#pragma once
#include "gns_backward.cuh"

namespace gns {

constexpr int kB2MaxWalk = 4;       // in-lines walked by one slot (register array of slope words)
constexpr int kB2ActSlots = 3;      // activation staging ring (consumer side)

// mbarrier indices
struct Bwd2Bar {
  enum : int {
    HID_FULL = 0,      // [5] producers -> consumers: D2L, D1L, D2LN, D1LN, ADJP written
    HID_EMPTY = 5,     // [5] consumers -> producers: block consumed
    ADJM_FULL = 10,    // adj m' rows written (they live in the D2LN + D1LN blocks)
    GROW_FULL = 11,    // [2] output adjoint row of the scalar nets (q = 0, 1)
    GROW_EMPTY = 13,   // [2]
    ACT_FULL = 15,     // [3] bulk copy of an activation block landed
    ACT_EMPTY = 18,    // [3] all consumer warps are done with the block
    STATE_FULL = 21, STATE_EMPTY = 22,
    NXT_FULL = 23,     // [2] v', theta', dP' rows of the state leaving the step (producers)
    W_FULL = 25,       // the step's weights (producers)
    COUNT = 26
  };
};
enum : int { B2_D2L = 0, B2_D1L = 1, B2_D2LN = 2, B2_D1LN = 3, B2_ADJP = 4 };

struct Bwd2Smem {          // offsets in floats from the start of dynamic shared memory
  int state;               // [(4+L)][NbP]  state entering the step (consumers), bulk copy
  int nxt;                 // [2][3][NbP]   v', theta', dP' of the state leaving the step (producers), bulk copy
  int gdP;                 // [NbP]         adjoint of dP' (published for the line phase)
  int cst;                 // [6][NbP]      Gs, sum 2(Pset-Pmin), sum 2(Pmax-Pset) by bus rank; Y, 1/tau, shift by alias line id
  int trig;                // [3][NbP]      D, sin D, cos D of the alias lines
  int adjD;                // [NbP]
  int deg;                 // [NbP]         in-degree of the bus (row of dc)
  int ones_b;              // [NbP]         1 for columns < N
  int ones_l;              // [EP]          1 for columns < E
  int grow;                // [2][NbP]
  int lineg;               // [5][EP]
  int featp;               // [5][EP]       line features in in_pos order (wide rows of dW1f)
  int weights;             // [wstep]
  int hid_b;               // [3][H][NbP]   D2L, D1L, ADJP
  int hid_l;               // [2][H][EP]    D2LN, D1LN (contiguous: also hosts the adj m' rows [L][NbP])
  int act;                 // [kB2ActSlots][H][EP]
  int zrow;                // [EP] zeros: wide row of the padding rows of a tile (loads stay unconditional)
  int zc;                  // [2][40] hidden-side constants for MMA rows past the real columns: zeros, and (1, 0) every 10 floats
  int red;                 // [2][8][4]
  int topo;                // uint16 [7][Epad]: fa, ta, fr, tr, in_ids, in_pos, out_ids
  int mbar;                // [Bwd2Bar::COUNT] x 8 bytes
  int total;
};

__host__ __device__ inline Bwd2Smem make_bwd2_smem(int L, int H, int E, int wstep, const Act2Layout& a2) {
  Bwd2Smem s{};
  int o = 0;
  auto take = [&](int n) { int r = o; o += pad4(n); return r; };
  const int NbP = a2.NbP, EP = a2.EP;
  s.state = take((4 + L) * NbP);
  s.nxt = take(2 * 3 * NbP);
  s.gdP = take(NbP);
  s.cst = take(6 * NbP);
  s.trig = take(3 * NbP);
  s.adjD = take(NbP);
  s.deg = take(NbP);
  s.ones_b = take(NbP);
  s.ones_l = take(EP);
  s.grow = take(2 * NbP);
  s.lineg = take(5 * EP);
  s.featp = take(5 * EP);
  s.weights = take(wstep);
  s.hid_b = take(3 * H * NbP);
  s.hid_l = take(2 * H * EP);
  s.act = take(kB2ActSlots * H * EP);
  s.zrow = take(EP);
  s.zc = take(80);
  s.red = take(2 * 8 * 4);
  s.topo = take((7 * pad4(E) + 1) / 2);
  s.mbar = take(2 * Bwd2Bar::COUNT);
  s.total = o;
  return s;
}

struct Bwd2Args {
  const float* params;      // packed [K][wstep]
  const float* buses; const float* lines; const float* gens;
  const float* ck2;         // [Sg][K+1][a2.state]
  const float* pglob;       // [Sg][K]
  const float* act;         // [Sg][K][a2.step]
  const float* grad_total; const float* grad_last; const float* grad_v; const float* grad_theta;
  float* gacc;              // [ctas * CW][K][FragLayout.step], zeroed by the host
  const uint16_t* topo;
  long long S;
  int N, Ns, E, Gn, K;
  int PW, CW;
  unsigned char role[32];   // warp -> producer rank (0 .. PW-1) or 0x80 | consumer rank
  int maxwalk;
  Act2Layout a2;
  Bwd2Smem sm;
  TopoOffsets to;
  float wk[kMaxK];
};

__device__ __forceinline__ float slope_of(uint32_t w, int bit) { return ((w >> bit) & 1u) ? 1.f : kSlope; }
// x[o][h] *= slope(bit shift+o of w[h])
template <int H>
__device__ __forceinline__ void apply_slope(float (&x)[H][2], const uint32_t (&w)[2], int shift) {
#pragma unroll
  for (int o = 0; o < H; ++o) {
    const float2 r = __fmul2_rn(make_float2(x[o][0], x[o][1]), make_float2(slope_of(w[0], shift + o), slope_of(w[1], shift + o)));
    x[o][0] = r.x; x[o][1] = r.y;
  }
}

// deterministic sum over the producer warps of NV values (barrier 1); every producer thread gets the totals
template <int NV>
__device__ __forceinline__ void prod_sum(float (&x)[NV], float* red, int PW, int warp, int& parity) {
  const int lane = threadIdx.x & 31;
  float* buf = red + parity * 32;
  parity ^= 1;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
    for (int v = 0; v < NV; ++v) x[v] += __shfl_xor_sync(0xffffffffu, x[v], off);
  if (lane == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) buf[warp * 4 + v] = x[v];
  }
  named_bar_sync(1, PW * 32);
#pragma unroll
  for (int v = 0; v < NV; ++v) x[v] = 0.f;
  for (int w = 0; w < PW; ++w)
#pragma unroll
    for (int v = 0; v < NV; ++v) x[v] += buf[w * 4 + v];
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {   // fire-and-forget, 16-byte aligned
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One weight-gradient call of a consumer warp: D[hid c][wide r] += sum over this warp's share of the items, C hidden
// columns (10, or 11 = the H columns and a constant 1), R wide rows (N of the MMA, 8 per tile).
//   hid:  ITEM-major block [item][10]; MMA row g is hidden column 2g and row g+8 column 2g+1, so (a0, a1) and (a2, a3)
//         of a k8 step are two 64-bit loads (items 4t+2s and 4t+2s+1 of the chunk, s = k8 step) - no register
//         shuffling in front of the HMMA; lanes of rows past the real columns read constants (zc) with stride 0.
//   row(r): ROW-major [item] rows with a stride = 16 mod 32 floats; one 128-bit load (items 4t..4t+3) feeds both k8
//         steps.  Rows past R read the zero row, so every load is unconditional.
// The accumulators stay in registers over all chunks; one 128-bit reduction per lane and tile at the end.
template <int C, int R, class RowFn>
__device__ __forceinline__ void cons_call(const float* hid, RowFn row, const float* zrow, const float* zc, int nch, int cw,
                                          int CW, float* __restrict__ gfrag) {
  static_assert(C == 10 || C == 11, "hidden columns");
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  constexpr int NT = (R + 7) / 8;
  const int c0 = (nch * cw) / CW, c1 = (nch * (cw + 1)) / CW;
  const bool a_ok = g < 5;
  const int hs = a_ok ? 10 : 0;                                   // floats between consecutive items (0: constants)
  const float* ap = a_ok ? hid + 2 * g + (c0 * 16 + 4 * t) * 10 : zc + ((C == 11 && g == 5) ? 40 : 0);
  const float* rp[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int r = nt * 8 + g;
    rp[nt] = (((nt * 8 + 8 <= R) || (r < R)) ? row(r) : zrow) + c0 * 16 + 4 * t;
  }
  float accA[NT][4], accB[NT][4];      // big x big, and the two cross terms: separate chains (accuracy and latency)
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) { accA[nt][j] = 0.f; accB[nt][j] = 0.f; }
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    uint32_t ab[2][4], as[2][4];
#pragma unroll
    for (int st = 0; st < 2; ++st) {
      const float2 lo = *reinterpret_cast<const float2*>(ap + (2 * st) * 10);
      const float2 hi = *reinterpret_cast<const float2*>(ap + (2 * st + 1) * 10);
      // the tensor core reads the top 19 bits of a TF32 operand: the raw value IS its "big" part
      ab[st][0] = __float_as_uint(lo.x); ab[st][1] = __float_as_uint(lo.y);
      ab[st][2] = __float_as_uint(hi.x); ab[st][3] = __float_as_uint(hi.y);
      as[st][0] = __float_as_uint(lo.x - __uint_as_float(ab[st][0] & 0xffffe000u));
      as[st][1] = __float_as_uint(lo.y - __uint_as_float(ab[st][1] & 0xffffe000u));
      as[st][2] = __float_as_uint(hi.x - __uint_as_float(ab[st][2] & 0xffffe000u));
      as[st][3] = __float_as_uint(hi.y - __uint_as_float(ab[st][3] & 0xffffe000u));
    }
    ap += 16 * hs;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float4 b = *reinterpret_cast<const float4*>(rp[nt]);
      rp[nt] += 16;
      const uint32_t bb[2][2] = {{__float_as_uint(b.x), __float_as_uint(b.y)}, {__float_as_uint(b.z), __float_as_uint(b.w)}};
      uint32_t bs[2][2];
      bs[0][0] = __float_as_uint(b.x - __uint_as_float(bb[0][0] & 0xffffe000u));
      bs[0][1] = __float_as_uint(b.y - __uint_as_float(bb[0][1] & 0xffffe000u));
      bs[1][0] = __float_as_uint(b.z - __uint_as_float(bb[1][0] & 0xffffe000u));
      bs[1][1] = __float_as_uint(b.w - __uint_as_float(bb[1][1] & 0xffffe000u));
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        mma_tf32(accB[nt], as[st], bb[st][0], bb[st][1]);
        mma_tf32(accA[nt], ab[st], bb[st][0], bb[st][1]);
        mma_tf32(accB[nt], ab[st], bs[st][0], bs[st][1]);
      }
    }
  }
  if (lane < 24) {     // hidden columns 0..11
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
      red_add_v4(gfrag + nt * kFragTile2 + lane * 4, accA[nt][0] + accB[nt][0], accA[nt][1] + accB[nt][1],
                 accA[nt][2] + accB[nt][2], accA[nt][3] + accB[nt][3]);
  }
}

// hidden-side vectors of the thread's two items -> item-major block [item][H]: H/2 64-bit stores per item
template <int H>
__device__ __forceinline__ void store_hid(float* blk, const int (&item)[2], const float (&x)[H][2]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float* p = blk + item[h] * H;
#pragma unroll
    for (int o = 0; o < H; o += 2) *reinterpret_cast<float2*>(p + o) = make_float2(x[o][h], x[o + 1][h]);
  }
}

template <int L, int H, bool MULTI>
__global__ void __launch_bounds__(384, 1) gns_backward2_kernel(const Bwd2Args a) {
  constexpr WLayout W = make_wlayout(L, H, MULTI);
  constexpr FragLayout FL = make_frag_layout(L, H, kFragTile2);
  constexpr int HP = pad4(H);
  static_assert(H == 10, "item-major hid blocks: 10 floats per item");
  static_assert(H <= 16, "slope words hold 2H bits");
  using B = Bwd2Bar;

  extern __shared__ __align__(16) float smem[];
  const int N = a.N, Ns = a.Ns, E = a.E, Gn = a.Gn, K = a.K, PW = a.PW, CW = a.CW;
  const int NbP = a.a2.NbP, EP = a.a2.EP;
  const int PT = PW * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x;
  float* const s_state = smem + a.sm.state;
  float* const s_nxt = smem + a.sm.nxt;
  float* const s_gdP = smem + a.sm.gdP;
  float* const s_cst = smem + a.sm.cst;
  float* const s_trig = smem + a.sm.trig;
  float* const s_adjD = smem + a.sm.adjD;
  float* const s_deg = smem + a.sm.deg;
  float* const s_ones_b = smem + a.sm.ones_b;
  float* const s_ones_l = smem + a.sm.ones_l;
  float* const s_grow = smem + a.sm.grow;
  float* const s_lineg = smem + a.sm.lineg;
  float* const s_featp = smem + a.sm.featp;
  float* const s_w = smem + a.sm.weights;
  float* const s_hid_b = smem + a.sm.hid_b;
  float* const s_hid_l = smem + a.sm.hid_l;
  float* const s_act = smem + a.sm.act;
  float* const s_red = smem + a.sm.red;
  const float* const s_zrow = smem + a.sm.zrow;
  float* const s_zc = smem + a.sm.zc;
  uint16_t* const s_topo = reinterpret_cast<uint16_t*>(smem + a.sm.topo);
  uint64_t* const s_bar = reinterpret_cast<uint64_t*>(smem + a.sm.mbar);
  const int Epad = pad4(E);
  const uint16_t* const t_fa = s_topo + 0 * Epad;
  const uint16_t* const t_ta = s_topo + 1 * Epad;
  const uint16_t* const t_fr = s_topo + 2 * Epad;
  const uint16_t* const t_tr = s_topo + 3 * Epad;
  const uint16_t* const t_ini = s_topo + 4 * Epad;
  const uint16_t* const t_inp = s_topo + 5 * Epad;
  const uint16_t* const t_outi = s_topo + 6 * Epad;
  float* const s_hid[5] = {s_hid_b, s_hid_b + H * NbP, s_hid_l, s_hid_l + H * EP, s_hid_b + 2 * H * NbP};
  float* const s_adjm = s_hid_l;     // adj m' rows [L][NbP] of the m-net's output layer call

  // ---- one-time setup by the whole CTA ----
  for (int i = tid; i < a.sm.total; i += T) smem[i] = 0.f;   // padding columns and tails must hold finite values
  __syncthreads();
  {
    uint16_t* st = s_topo;
    const int src[7] = {a.to.fa, a.to.ta, a.to.fr, a.to.tr, a.to.in_ids, a.to.in_pos, a.to.out_ids};
    for (int i = tid; i < 7 * E; i += T) {
      const int w = i / E, e = i - w * E;
      st[w * Epad + e] = a.topo[src[w] + e];
    }
    for (int i = tid; i < N; i += T) s_ones_b[i] = 1.f;
    if (tid < 4) s_zc[40 + 10 * tid] = 1.f;          // (1, 0) at every item offset of a chunk quarter: the constant-1 column
    for (int i = tid; i < E; i += T) s_ones_l[i] = 1.f;
    for (int s = tid; s < Ns; s += T)
      if ((int)a.topo[a.to.prim_of + s] == s)
        s_deg[a.topo[a.to.brank + s]] = (float)((int)a.topo[a.to.in_fe + s] - (int)a.topo[a.to.in_b + s]);
    if (tid == 0) {
      for (int i = 0; i < 5; ++i) { mbar_init(s_bar + B::HID_FULL + i, PW); mbar_init(s_bar + B::HID_EMPTY + i, CW); }
      mbar_init(s_bar + B::ADJM_FULL, PW);
      for (int i = 0; i < 2; ++i) { mbar_init(s_bar + B::GROW_FULL + i, PW); mbar_init(s_bar + B::GROW_EMPTY + i, CW); }
      for (int i = 0; i < kB2ActSlots; ++i) { mbar_init(s_bar + B::ACT_FULL + i, 1); mbar_init(s_bar + B::ACT_EMPTY + i, CW); }
      mbar_init(s_bar + B::STATE_FULL, 1); mbar_init(s_bar + B::STATE_EMPTY, CW);
      mbar_init(s_bar + B::NXT_FULL, 1); mbar_init(s_bar + B::NXT_FULL + 1, 1);
      mbar_init(s_bar + B::W_FULL, 1);
      fence_proxy_async();
    }
  }
  __syncthreads();
  const int first_grid = blockIdx.x, grid_step = gridDim.x;

  const int wrole = a.role[warp];
  if (wrole < 0x80) {
    // =====================================================================================================
    // PRODUCERS
    // =====================================================================================================
    const int p = wrole * 32 + lane;
    int slot[2], br[2], ext[2], e_in0[2], deg[2], e_full1[2], e_out0[2], e_out1[2];
    bool on[2], prim[2], is_gen[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      slot[h] = 2 * p + h;
      on[h] = slot[h] < Ns;
      const int sl = on[h] ? slot[h] : 0;
      prim[h] = on[h] && (int)a.topo[a.to.prim_of + sl] == sl;
      br[h] = a.topo[a.to.brank + sl];
      ext[h] = a.topo[a.to.ext_of + sl];
      e_in0[h] = on[h] ? (int)a.topo[a.to.in_b + sl] : 0;
      deg[h] = on[h] ? (int)a.topo[a.to.in_e + sl] - e_in0[h] : 0;
      e_full1[h] = prim[h] ? (int)a.topo[a.to.in_fe + sl] : e_in0[h];
      e_out0[h] = prim[h] ? (int)a.topo[a.to.out_b + sl] : 0;
      e_out1[h] = prim[h] ? (int)a.topo[a.to.out_e + sl] : 0;
      is_gen[h] = prim[h] && a.topo[a.to.gen_e + sl] > a.topo[a.to.gen_b + sl];
    }
    // branch-free stores: a slot that does not own a bus writes to the first padding column / item (never read as
    // a non-zero product: every wide row is zero there)
    const int colb[2] = {prim[0] ? br[0] : N, prim[1] ? br[1] : N};
    // twin groups (2 or 4 adjacent, aligned slots of one bus): both slots of a thread belong to the same group
    const int gsz_t = on[0] ? (int)a.topo[a.to.gsz + slot[0]] : 1;
    const int lead_lane = gsz_t >= 2 ? lane - (p - (int)a.topo[a.to.prim_of + slot[0]] / 2) : lane;
    const bool warp_has_twins = __any_sync(0xffffffffu, gsz_t > 1);
    const int warp_max_deg = __reduce_max_sync(0xffffffffu, max(deg[0], deg[1]));
    const int maxwalk = a.maxwalk;
    int red_parity = 0;
    uint32_t ph_empty = 0x7fu;            // parity to wait for on HID_EMPTY[0..4], GROW_EMPTY[0..1] (bits 5, 6): first wait passes
    uint32_t ph_nxt = 0, ph_w = 0;        // NXT_FULL[2] (bits 0, 1), W_FULL
    int nxt_buf = 0;

    auto wait_empty = [&](int i) {        // i = 0..4 hid blocks, 5..6 output-adjoint rows
      mbar_wait(s_bar + (i < 5 ? B::HID_EMPTY + i : B::GROW_EMPTY + (i - 5)), (ph_empty >> i) & 1u);
      ph_empty ^= 1u << i;
    };
    auto signal = [&](int bar) {          // all lanes of the warp wrote their part
      __syncwarp();
      if (lane == 0) mbar_arrive(s_bar + bar);
    };
    auto issue_weights = [&](int k) {
      fence_proxy_async();
      mbar_expect_tx(s_bar + B::W_FULL, W.wstep * 4);
      bulk_g2s(s_w, a.params + (size_t)k * W.wstep, W.wstep * 4, s_bar + B::W_FULL);
    };
    auto issue_nxt = [&](long long grid, int kidx, int buf) {   // rows v, theta, dP of state checkpoint kidx
      mbar_expect_tx(s_bar + B::NXT_FULL + buf, 3 * NbP * 4);
      bulk_g2s(s_nxt + buf * 3 * NbP, a.ck2 + ((size_t)grid * (K + 1) + kidx) * (size_t)a.a2.state, 3 * NbP * 4,
               s_bar + B::NXT_FULL + buf);
    };
    if (p == 0 && first_grid < a.S) {
      issue_weights(K - 1);
      issue_nxt(first_grid, K, 0);
    }

    float amr[L][2];       // adjoint of the bus latent (lives in registers across the steps of a grid)
    float adj4[4][2];      // adjoint of v, theta, dP, dQ of the state between two steps

    for (long long grid = first_grid; grid < a.S; grid += grid_step) {
      // ---------------- per-grid constants from the raw input rows ----------------
      const float gtot = a.grad_total[grid];
      const float glast = a.grad_last ? a.grad_last[grid] : 0.f;
      float s3[3] = {0.f, 0.f, 0.f};
      for (int j = p; j < Gn; j += PT) {
        const float* gr = a.gens + ((size_t)grid * Gn + j) * 7;
        s3[0] += __ldg(gr + 3); s3[1] += __ldg(gr + 2); s3[2] += __ldg(gr + 1);     // Pset, Pmin, Pmax
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (prim[h]) {
          s_cst[0 * NbP + br[h]] = __ldg(a.buses + ((size_t)grid * N + ext[h]) * 6 + 4);   // Gs
          float lo = 0.f, hi = 0.f;
          const int j0 = a.topo[a.to.gen_b + slot[h]], j1 = a.topo[a.to.gen_e + slot[h]];
          for (int j = j0; j < j1; ++j) {
            const float* gr = a.gens + ((size_t)grid * Gn + (int)a.topo[a.to.gen_ids + j]) * 7;
            const float Pmax = __ldg(gr + 1), Pmin = __ldg(gr + 2), Pset = __ldg(gr + 3);
            lo += 2.f * (Pset - Pmin); hi += 2.f * (Pmax - Pset);
          }
          s_cst[1 * NbP + br[h]] = lo; s_cst[2 * NbP + br[h]] = hi;
          adj4[0][h] = a.grad_v ? a.grad_v[(size_t)grid * N + ext[h]] : 0.f;
          adj4[1][h] = a.grad_theta ? a.grad_theta[(size_t)grid * N + ext[h]] : 0.f;
        } else {
          adj4[0][h] = 0.f; adj4[1][h] = 0.f;
        }
        adj4[2][h] = 0.f; adj4[3][h] = 0.f;
        const int j = 2 * p + h;               // alias line id (ref GNS/main.py:41: bus numbers re-read as line numbers)
        if (j < N) {
          const float* lr = a.lines + ((size_t)grid * E + j) * 7;
          const float r = __ldg(lr + 2), x = __ldg(lr + 3);
          s_cst[3 * NbP + j] = 1.0f / sqrtf(r * r + x * x);
          s_cst[4 * NbP + j] = 1.0f / __ldg(lr + 5);          // 1 / tau, like the forward kernel
          s_cst[5 * NbP + j] = __ldg(lr + 6);
        }
      }
#pragma unroll
      for (int i = 0; i < L; ++i) { amr[i][0] = 0.f; amr[i][1] = 0.f; }
      prod_sum<3>(s3, s_red, PW, wrole, red_parity);          // its barrier also publishes s_cst
      const float sPset = s3[0], sPmin = s3[1], sPmax = s3[2];
      float pglob = __ldg(a.pglob + (size_t)grid * K + (K - 1));

      for (int k = K - 1; k >= 0; --k) {
        const float* const act_k = a.act + ((size_t)grid * K + k) * (size_t)a.a2.step;
        // slope words of the first pair (m) travel under the physics phase
        uint2 mbw = make_uint2(0u, 0u);
        if (on[0]) mbw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(act_k + a.a2.mask[2]) + 2 * p));
        // next state rows for the producers (other buffer): needed at the start of step k-1 / the next grid
        if (p == 0) {
          if (k >= 1) issue_nxt(grid, k, nxt_buf ^ 1);
          else if (grid + grid_step < a.S) issue_nxt(grid + grid_step, K, nxt_buf ^ 1);
        }
        const float pg_now = pglob;
        if (k >= 1) pglob = __ldg(a.pglob + (size_t)grid * K + (k - 1));
        else if (grid + grid_step < a.S) pglob = __ldg(a.pglob + (size_t)(grid + grid_step) * K + (K - 1));

        // ---------------- physics adjoint (a): dP' adjoint, lambda coupling, alias-line trig ----------------
        const bool lo_branch = pg_now < sPset;
        const float lam = lo_branch ? (pg_now - sPmin) / (2.f * (sPset - sPmin))
                                    : (pg_now - 2.f * sPset + sPmax) / (2.f * (sPmax - sPset));
        const bool lo_arm = lam < 0.5f;
        const float coef = (gtot * a.wk[k] + ((k == K - 1) ? glast : 0.f)) * (2.0f / (float)N);
        mbar_wait(s_bar + B::NXT_FULL + nxt_buf, (ph_nxt >> nxt_buf) & 1u);
        ph_nxt ^= 1u << nxt_buf;
        const float* const nx = s_nxt + nxt_buf * 3 * NbP;
        nxt_buf ^= 1;
        float gdP[2], vpr[2], Gsv[2];
        float part[1] = {0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          gdP[h] = 0.f; vpr[h] = 0.f; Gsv[h] = 0.f;
          if (prim[h]) {
            gdP[h] = adj4[2][h] + coef * nx[2 * NbP + br[h]];
            s_gdP[br[h]] = gdP[h];
            part[0] += gdP[h] * (lo_arm ? s_cst[1 * NbP + br[h]] : s_cst[2 * NbP + br[h]]);
            vpr[h] = nx[br[h]];
            Gsv[h] = s_cst[br[h]];
          }
          const int j = 2 * p + h;
          if (j < N) {   // alias line j: D_j = theta'[f_j] - theta'[t_j]
            const float d = nx[NbP + (int)t_fr[j]] - nx[NbP + (int)t_tr[j]];
            float sd, cd;
            fast_sincos(d, sd, cd);
            s_trig[j] = d; s_trig[NbP + j] = sd; s_trig[2 * NbP + j] = cd;
          }
        }
        prod_sum<1>(part, s_red, PW, wrole, red_parity);       // its barrier also publishes s_gdP and s_trig
        const float adj_pg = part[0] / (lo_branch ? 2.f * (sPset - sPmin) : 2.f * (sPmax - sPset));

        // ---------------- physics adjoint (b): per-line partials ----------------
#pragma unroll 1
        for (int e = p; e < E; e += PT) {
          const int fi = t_fr[e], ti = t_tr[e], fa = t_fa[e], ta = t_ta[e];
          const float vf = nx[fi], vt = nx[ti];
          const float thf = nx[NbP + fi], tht = nx[NbP + ti];
          const float g_pf = s_gdP[ti];      // p_from lands on the receiving bus
          const float g_pt = s_gdP[fi];      // p_to lands on the sending bus
          const float Yf = s_cst[3 * NbP + fa], itf = s_cst[4 * NbP + fa], shf = s_cst[5 * NbP + fa];
          const float Df = s_trig[fa], sDf = s_trig[NbP + fa], cDf = s_trig[2 * NbP + fa];
          const float Yt = s_cst[3 * NbP + ta], itt = s_cst[4 * NbP + ta], sht = s_cst[5 * NbP + ta];
          const float DB = s_trig[ta], sDB = s_trig[NbP + ta], cDB = s_trig[2 * NbP + ta];
          const float a1 = thf - tht - Df - shf;
          const float a2 = tht - thf - Df + shf;
          const float a3 = tht - thf + DB - sht;               // delta_ji[dst] = -D_B
          float s1, c1, s2, c2, s3v, c3;
          fast_sincos(a1, s1, c1);
          fast_sincos(a2, s2, c2);
          fast_sincos(a3, s3v, c3);
          const float yft = Yf * itf, yftt = Yf * (itf * itf), ytt = Yt * itt;
          const float t1 = vf * vt * yft, u1 = vt * vf * ytt;
          const float sDt = -sDB;
          const float ss = s1 + s2;
          const float inner = t1 * ss + vf * yftt * sDf + vt * vt * Yf * sDf;       // |.| of ref GNS/main.py:41
          const float g_in = adj_pg * ((inner > 0.f) ? 1.f : ((inner < 0.f) ? -1.f : 0.f));
          const float gvf = g_in * (vt * yft * ss + yftt * sDf) + g_pf * (vt * yft * s1 + 2.f * vf * yftt * sDf) +
                            g_pt * (vt * ytt * s3v);
          const float gvt = g_in * (vf * yft * ss + 2.f * vt * Yf * sDf) + g_pf * (vf * yft * s1) +
                            g_pt * (vf * ytt * s3v + 2.f * vt * Yt * sDt);
          const float G1 = (g_in + g_pf) * t1 * c1, G2 = g_in * t1 * c2, G3 = g_pt * u1 * c3;
          const float gth = G1 - G2 - G3;
          const float gDA = -G1 - G2 + (g_in * (vf * yftt + vt * vt * Yf) + g_pf * vf * vf * yftt) * cDf;
          const float gDB = G3 - g_pt * vt * vt * Yt * cDB;
          s_lineg[0 * EP + e] = gvf;
          s_lineg[1 * EP + e] = gvt;
          s_lineg[2 * EP + e] = gth;
          s_lineg[3 * EP + e] = gDA;
          s_lineg[4 * EP + e] = gDB;
        }
        named_bar_sync(1, PT);

        // ---------------- physics adjoint (c): CSR gathers replace the forward scatter-adds ----------------
        float adjv[2], adjth[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          adjv[h] = 0.f; adjth[h] = 0.f;
          if (prim[h]) {
            float sv = 0.f, sth = 0.f, sD = 0.f;
            for (int e = e_out0[h]; e < e_out1[h]; ++e) {
              const int eo = t_outi[e];
              sv += s_lineg[0 * EP + eo]; sth += s_lineg[2 * EP + eo]; sD += s_lineg[3 * EP + eo];
            }
            for (int e = e_in0[h]; e < e_full1[h]; ++e) {
              const int eo = t_ini[e];
              sv += s_lineg[1 * EP + eo]; sth -= s_lineg[2 * EP + eo]; sD += s_lineg[4 * EP + eo];
            }
            adjv[h] = adj4[0][h] + sv + gdP[h] * (-2.f * Gsv[h] * vpr[h]) + adj_pg * (2.f * vpr[h] * Gsv[h]);
            adjth[h] = adj4[1][h] + sth;
            s_adjD[ext[h]] = sD;                               // alias line id == external bus number
          }
        }
        named_bar_sync(1, PT);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (prim[h]) {                                       // D_e = theta[f_e] - theta[t_e] for alias lines e < N
            for (int e = e_out0[h]; e < e_out1[h]; ++e) { const int l = t_outi[e]; if (l < N) adjth[h] += s_adjD[l]; }
            for (int e = e_in0[h]; e < e_full1[h]; ++e) { const int l = t_ini[e]; if (l < N) adjth[h] -= s_adjD[l]; }
          }
        }

        // ---------------- MLP adjoint: dX chains (two slots per FFMA2) and the hid blocks of the weight gradients ----------------
        mbar_wait(s_bar + B::W_FULL, ph_w);
        ph_w ^= 1u;
        float a4[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a4[i][0] = 0.f; a4[i][1] = 0.f; }
        float adjA[H][2];
#pragma unroll
        for (int o = 0; o < H; ++o) { adjA[o][0] = 0.f; adjA[o][1] = 0.f; }

#pragma unroll 1
        for (int qq = 0; qq < 3; ++qq) {
          const int q = (qq == 0) ? 2 : qq - 1;      // m-net first: its output layer reads adj m' before anyone adds to it
          const int ql = MULTI ? q : 0;
          const float* wphi = s_w + (MULTI ? q * W.phi_size : 0);
          const float* wln = s_w + W.off_ln[0] + q * W.ln_size_s;
          const float* wmf = s_w + W.off_mf[0] + q * W.mf_size;
          const uint32_t mb[2] = {mbw.x, mbw.y};
          const bool do_phi = MULTI || qq == 2;
          // slope words: the next pair's bus word and this pair's line words travel under the L-net adjoint
          if (qq < 2 && on[0])
            mbw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(act_k + a.a2.mask[qq]) + 2 * p));
          const uint2* const mlp = reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(act_k + a.a2.mask[ql]) + a.a2.NsM + 2 * p);
          uint2 mlw = make_uint2(0u, 0u);     // slope word of the line walked in iteration 0; the next ones are fetched one iteration ahead
          if (do_phi && maxwalk > 0 && on[0]) mlw = __ldg(mlp);

          // ---- output layer ----
          float d2[H][2];
#pragma unroll
          for (int o = 0; o < H; ++o) { d2[o][0] = 0.f; d2[o][1] = 0.f; }
          if (q < 2) {
            float gv[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) gv[h] = prim[h] ? ((q == 0) ? (is_gen[h] ? 0.f : adjv[h]) : adjth[h]) : 0.f;
            row_axpy<H, HP, 2>(d2, gv, wln + W.ln_wo);
            wait_empty(5 + q);
#pragma unroll
            for (int h = 0; h < 2; ++h) s_grow[q * NbP + colb[h]] = gv[h];
            signal(B::GROW_FULL + q);
          } else {
            // adj m' rows for the consumers (they live in the line blocks, free at this point of the step)
            wait_empty(B2_D2LN);
            wait_empty(B2_D1LN);
            if (k == K - 1) {      // the previous grid's last line call is over: its line features may go
              for (int i = p; i < E; i += PT) {
                const float* lr = a.lines + ((size_t)grid * E + (int)t_ini[i]) * 7 + 2;
                const int col = t_inp[i];
#pragma unroll
                for (int c = 0; c < 5; ++c) s_featp[c * EP + col] = __ldg(lr + c);
              }
            }
            // (their padding columns hold stale line records; the constant-1 hidden column of the bias gradient
            // is not zero there, so they are cleared: dump column N included, the slots without a bus write zeros too)
            for (int i = p; i < L * (NbP - N); i += PT) {
              const int r = i / (NbP - N);
              s_adjm[r * NbP + N + (i - r * (NbP - N))] = 0.f;
            }
            {
              float* p0 = s_adjm + colb[0];
              float* p1 = s_adjm + colb[1];
#pragma unroll
              for (int i = 0; i < L; ++i) { *p0 = amr[i][0]; *p1 = amr[i][1]; p0 += NbP; p1 += NbP; }
            }
            signal(B::ADJM_FULL);
#pragma unroll
            for (int i = 0; i < L; ++i) row_axpy<H, HP, 2>(d2, amr[i], wln + W.ln_wo + i * HP);
          }
          apply_slope<H>(d2, mb, H);
          wait_empty(B2_D2L);
          store_hid(s_hid[B2_D2L], colb, d2);
          signal(B::HID_FULL + B2_D2L);
          // ---- second layer ----
          float d1[H][2];
#pragma unroll
          for (int j = 0; j < H; ++j) {
            d1[j][0] = 0.f; d1[j][1] = 0.f;
            row_dot<H, HP, 2>(d1[j], d2, wln + W.ln_w2 + j * HP);
          }
          apply_slope<H>(d1, mb, 0);
          wait_empty(B2_D1L);
          store_hid(s_hid[B2_D1L], colb, d1);
          signal(B::HID_FULL + B2_D1L);
          // ---- dX of the first layer: state adjoints, latent adjoint, aggregate adjoint (fused block) ----
#pragma unroll
          for (int i = 0; i < 4; ++i) row_dot<H, HP, 2>(a4[i], d1, wln + W.ln_w1 + i * HP);
#pragma unroll
          for (int i = 0; i < L; ++i) row_dot<H, HP, 2>(amr[i], d1, wln + W.ln_w1 + (4 + i) * HP);
          if (MULTI) {
#pragma unroll
            for (int o = 0; o < H; ++o) { adjA[o][0] = 0.f; adjA[o][1] = 0.f; }
          }
#pragma unroll
          for (int j = 0; j < H; ++j) row_dot<H, HP, 2>(adjA[j], d1, wmf + j * HP);

          if (do_phi) {
            // ---- phi net: every slot walks its in-lines; twins take the aggregate's adjoint from the bus owner ----
            if (warp_has_twins) {
#pragma unroll
              for (int o = 0; o < H; ++o) {
                const float t = __shfl_sync(0xffffffffu, adjA[o][0], lead_lane);
                if (gsz_t >= 2) { adjA[o][0] = t; adjA[o][1] = t; }
              }
            }
            float adjP[H][2];
#pragma unroll
            for (int o = 0; o < H; ++o) { adjP[o][0] = 0.f; adjP[o][1] = 0.f; }
            wait_empty(B2_D2LN);
            wait_empty(B2_D1LN);
#pragma unroll 1
            for (int it = 0; it < warp_max_deg; ++it) {
              {
                const float* wp = wphi + opaque_zero();     // keep the weight rows in shared memory (no hoisting into spills)
                const bool live[2] = {it < deg[0], it < deg[1]};
                const uint32_t ml[2] = {mlw.x, mlw.y};
                if (it + 1 < maxwalk && on[0]) mlw = __ldg(mlp + (size_t)(it + 1) * (a.a2.NsM / 2));
                float e2[H][2], e1[H][2];
#pragma unroll
                for (int o = 0; o < H; ++o) { e2[o][0] = live[0] ? adjA[o][0] : 0.f; e2[o][1] = live[1] ? adjA[o][1] : 0.f; }
                apply_slope<H>(e2, ml, H);
#pragma unroll
                for (int j = 0; j < H; ++j) {
                  e1[j][0] = 0.f; e1[j][1] = 0.f;
                  row_dot<H, HP, 2>(e1[j], e2, wp + W.phi_w2 + j * HP);
                }
                apply_slope<H>(e1, ml, 0);
#pragma unroll
                for (int j = 0; j < H; ++j) { adjP[j][0] += e1[j][0]; adjP[j][1] += e1[j][1]; }
                const int coll[2] = {live[0] ? (int)t_inp[e_in0[0] + it] : E, live[1] ? (int)t_inp[e_in0[1] + it] : E};
                store_hid(s_hid[B2_D2LN], coll, e2);
                store_hid(s_hid[B2_D1LN], coll, e1);
              }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(s_bar + B::HID_FULL + B2_D2LN); mbar_arrive(s_bar + B::HID_FULL + B2_D1LN); }
            if (warp_has_twins) {     // the bus owner needs the sum over its twins; the others keep zeros
#pragma unroll
              for (int o = 0; o < H; ++o) {
                float s = adjP[o][0] + adjP[o][1];
                const float u = __shfl_down_sync(0xffffffffu, s, 1);
                if (gsz_t == 4) s += u;
                if (gsz_t >= 2) { adjP[o][0] = (lead_lane == lane) ? s : 0.f; adjP[o][1] = 0.f; }
              }
            }
            wait_empty(B2_ADJP);
            store_hid(s_hid[B2_ADJP], colb, adjP);
            signal(B::HID_FULL + B2_ADJP);
#pragma unroll
            for (int i = 0; i < L; ++i) row_dot<H, HP, 2>(amr[i], adjP, wphi + W.phi_w1m + i * HP);
          }
        }  // pairs
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          adj4[0][h] = adjv[h] + a4[0][h];
          adj4[1][h] = adjth[h] + a4[1][h];
          adj4[2][h] = a4[2][h];
          adj4[3][h] = a4[3][h];
        }
        named_bar_sync(1, PT);      // every producer is done with this step's weights
        if (p == 0) {
          if (k >= 1) issue_weights(k - 1);
          else if (grid + grid_step < a.S) issue_weights(K - 1);
        }
      }  // k
    }  // grid
  } else {
    // =====================================================================================================
    // CONSUMERS
    // =====================================================================================================
    const int cw = wrole & 0x7f;
    const bool issuer = cw == 0 && lane == 0;
    float* const gacc_w = a.gacc + ((size_t)blockIdx.x * CW + cw) * ((size_t)K * FL.step);
    const int nch_b = NbP / 16, nch_l = EP / 16;
    constexpr int NB = MULTI ? 12 : 10;          // activation blocks per step
    const int ACTSZ = H * EP;
    uint32_t ph_full = 0;        // parity to wait for on HID_FULL[0..4], ADJM_FULL (bit 5), GROW_FULL[0..1] (bits 6, 7), STATE_FULL (bit 8)
    uint32_t ph_sempty = 0;      // STATE_EMPTY (issuer)

    auto wait_full = [&](int i) {
      const int bar = i < 5 ? B::HID_FULL + i : (i == 5 ? B::ADJM_FULL : (i < 8 ? B::GROW_FULL + (i - 6) : B::STATE_FULL));
      mbar_wait(s_bar + bar, (ph_full >> i) & 1u);
      ph_full ^= 1u << i;
    };
    auto release = [&](int bar) {       // this warp is done reading the block
      __syncwarp();
      if (lane == 0) mbar_arrive(s_bar + bar);
    };
    // The issuing lane walks the activation blocks in consumption order: grid, step k = K-1..0, block r of the step
    // (per pair in the order m, v, theta: h2L, h1L, A, and the per-line h1 of the pair's phi net).
    long long ib_grid = first_grid;
    int ib_k = K - 1, ib_r = 0;
    auto issue_act = [&](int slot) {      // next block of the walk -> ring slot
      if (ib_grid >= a.S) return;
      const int r = ib_r;
      int qq, j;
      if (MULTI) { qq = r >> 2; j = r & 3; }
      else { qq = r < 9 ? r / 3 : 2; j = r - 3 * qq; }
      const int q = (qq == 0) ? 2 : qq - 1;
      const int off = j == 0 ? a.a2.h2L[q] : (j == 1 ? a.a2.h1L[q] : (j == 2 ? a.a2.A[q] : a.a2.h1line[MULTI ? q : 0]));
      const uint32_t bytes = (uint32_t)(H * (j == 3 ? EP : NbP) * 4);
      mbar_expect_tx(s_bar + B::ACT_FULL + slot, bytes);
      bulk_g2s(s_act + slot * ACTSZ, a.act + ((size_t)ib_grid * K + ib_k) * (size_t)a.a2.step + off, bytes, s_bar + B::ACT_FULL + slot);
      if (++ib_r == NB) {
        ib_r = 0;
        if (--ib_k < 0) { ib_k = K - 1; ib_grid += grid_step; }
      }
    };
    long long is_grid = first_grid;
    int is_k = K - 1;
    auto issue_state = [&]() {            // state entering the next step of the walk
      if (is_grid >= a.S) return;
      mbar_expect_tx(s_bar + B::STATE_FULL, (uint32_t)((4 + L) * NbP * 4));
      bulk_g2s(s_state, a.ck2 + ((size_t)is_grid * (K + 1) + is_k) * (size_t)a.a2.state, (uint32_t)((4 + L) * NbP * 4), s_bar + B::STATE_FULL);
      if (--is_k < 0) { is_k = K - 1; is_grid += grid_step; }
    };
    if (issuer) {
      for (int i = 0; i < kB2ActSlots; ++i) issue_act(i);
      issue_state();
    }
    int aslot = 0;               // ring slot of the next activation block
    uint32_t ph_act = 0;         // parity of the slot's next completion (ACT_FULL and ACT_EMPTY advance together)
    // wait for the next activation block; returns its shared-memory address
    auto act_wait = [&]() {
      mbar_wait(s_bar + B::ACT_FULL + aslot, (ph_act >> aslot) & 1u);
      return s_act + aslot * ACTSZ;
    };
    // all reads of the block by this warp are done; the issuing lane refills the slot once every warp is done
    auto act_done = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(s_bar + B::ACT_EMPTY + aslot);
      if (issuer) {
        mbar_wait(s_bar + B::ACT_EMPTY + aslot, (ph_act >> aslot) & 1u);
        issue_act(aslot);
      }
      ph_act ^= 1u << aslot;
      aslot = aslot == kB2ActSlots - 1 ? 0 : aslot + 1;
    };

    for (long long grid = first_grid; grid < a.S; grid += grid_step) {
      for (int k = K - 1; k >= 0; --k) {
        float* const gk = gacc_w + (size_t)k * FL.step;
        bool have_state = false;
#pragma unroll 1
        for (int qq = 0; qq < 3; ++qq) {
          const int q = (qq == 0) ? 2 : qq - 1;
          float* const gln = gk + q * FL.net;
          float* const gphi = gk + (MULTI ? q * FL.net : 0);
          const bool do_phi = MULTI || qq == 2;
          // ---- output layer: dWout, dbout ----
          {
            const float* h2L = act_wait();       // item-major; the 11th column is the constant 1 (bias gradient)
            if (q == 2) {
              wait_full(5);
              cons_call<H + 1, L>(h2L, [&](int r) { return s_adjm + r * NbP; }, s_zrow, s_zc, nch_b, cw, CW, gln + FL.out);
              release(B::HID_EMPTY + B2_D2LN);
              release(B::HID_EMPTY + B2_D1LN);
            } else {
              wait_full(6 + q);
              cons_call<H + 1, 1>(h2L, [&](int) { return s_grow + q * NbP; }, s_zrow, s_zc, nch_b, cw, CW, gln + FL.out);
              release(B::GROW_EMPTY + q);
            }
            act_done();
          }
          // ---- second layer: dW2, db2 ----
          {
            const float* h1L = act_wait();
            wait_full(B2_D2L);
            cons_call<H, H + 1>(s_hid[B2_D2L], [&](int r) { return r < H ? h1L + r * NbP : s_ones_b; }, s_zrow, s_zc, nch_b, cw,
                                CW, gln + FL.w2);
            release(B::HID_EMPTY + B2_D2L);
            act_done();
          }
          // ---- first layer: dW1[:4+L], dM, dc, db1 ----
          {
            const float* Ab = act_wait();
            wait_full(B2_D1L);
            if (!have_state) { wait_full(8); have_state = true; }
            cons_call<H, 4 + L + H + 2>(
                s_hid[B2_D1L],
                [&](int r) {
                  return r < 4 + L ? s_state + r * NbP
                                   : (r < 4 + L + H ? Ab + (r - 4 - L) * NbP : (r == 4 + L + H ? s_deg : s_ones_b));
                },
                s_zrow, s_zc, nch_b, cw, CW, gln + FL.w1);
            release(B::HID_EMPTY + B2_D1L);
            act_done();
          }
          if (do_phi) {
            // ---- phi net: per-line dW2 / db2 and dW1f over all lines at once, then dW1m / db1 per bus ----
            const float* h1n = act_wait();
            wait_full(B2_D2LN);
            cons_call<H, H + 1>(s_hid[B2_D2LN], [&](int r) { return r < H ? h1n + r * EP : s_ones_l; }, s_zrow, s_zc, nch_l, cw,
                                CW, gphi + FL.w2l);
            wait_full(B2_D1LN);
            cons_call<H, 5>(s_hid[B2_D1LN], [&](int r) { return s_featp + r * EP; }, s_zrow, s_zc, nch_l, cw, CW, gphi + FL.w1f);
            release(B::HID_EMPTY + B2_D2LN);
            release(B::HID_EMPTY + B2_D1LN);
            act_done();
            wait_full(B2_ADJP);
            cons_call<H, L + 1>(s_hid[B2_ADJP], [&](int r) { return r < L ? s_state + (4 + r) * NbP : s_ones_b; }, s_zrow, s_zc,
                                nch_b, cw, CW, gphi + FL.w1m);
            release(B::HID_EMPTY + B2_ADJP);
          }
        }  // pairs
        // the state rows of this step are dead: fetch those of the next one
        release(B::STATE_EMPTY);
        if (issuer) {
          mbar_wait(s_bar + B::STATE_EMPTY, ph_sempty);
          ph_sempty ^= 1u;
          issue_state();
        }
      }  // k
    }  // grid
  }
}

}  // namespace gns
