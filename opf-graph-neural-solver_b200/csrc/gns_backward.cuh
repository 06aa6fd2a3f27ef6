// gns_backward.cuh — persistent backward kernel: BPTT through the K steps of one CTA-batch
// of grids, the implicit autograd backward of ref GNS/main.py:288 restated by hand.
//
// Per step k = K-1 .. 0 (state_k = checkpoint written by the forward kernel):
//   1. physics adjoint: loss -> dP' -> (lambda / p_global coupling, line flows) -> v', theta'.
//      All scatter-adds of the forward become CSR gathers here, so there are no scatter atomics.
//   2. MLP adjoint per bus, thread-local: the bus thread reads the hidden activations the
//      forward kernel kept (ActLayout: post-LeakyReLU values, whose sign gives the slope) and
//      back-propagates them (dX), mirroring the bus-centric forward.  Nothing is recomputed.
//   3. weight gradients (the only GEMM-shaped work): dW^T[w][o] = sum_items wide[w] * hid[o].
//      Each warp transposes the per-item vectors of its 32 items through a private shared
//      memory tile and runs the 32-item product on the tensor cores (mma.sync m16n8k8 TF32 with
//      a 3-term split that keeps FP32 accuracy).  State rows (v, theta, dP, dQ, m, adj m) are
//      already stored [feature][item] and are read in place as B fragments.  The accumulator
//      cells go to a per-warp private block in global memory in fragment order (one or two 64-bit
//      reductions per lane and 8-row tile, L2 resident), summed over warps and re-ordered by a
//      second small kernel: deterministic, no contended atomics.
#pragma once
#include "gns_common.cuh"

namespace gns {

#ifndef GNS_KTS
#define GNS_KTS 48
#endif
constexpr int kTS = GNS_KTS;   // tile row stride in floats: 48 = 16 mod 32, so the 128-bit fragment loads of 8 rows x 4 quads are conflict-free

struct BwdSmem {          // offsets in floats from SmemPlan.extra
  int adj;                // [(4+L)][NGs]  adjoint of (v, theta, dP, dQ, m)
  int nxt;                // [4][NGs]      v', theta', dP', dQ' of the state leaving the step
  int lineg;              // [5][EGs]      per-line partials: d/dv_f, d/dv_t, d/dtheta_f, d/dD_A, d/dD_B
  int adjD;               // [NGs]         adjoint of the alias-line angle differences
  int tiles;              // [nwarps][tfloats]
  int tfloats;
  int mbar;               // 8-byte mbarrier of the per-step weight copy
  int total;
};

// Per-warp tile (floats): [ones row][hid block A][16 wide rows][S rows | hid block B].
// Wide rows are [row][item] with stride kTS.  A hid block holds the hidden-side vectors of the 32 items
// interleaved for the MMA A fragment: [pair p = c % 8][item][c / 8] with row stride kSA, so one LDS.128 at
// item a yields {c[a], c+8[a], c[a+1], c+8[a+1]} = (a0, a1, a2, a3) of an m16n8k8 with k slots (t, t+4) =
// items (a, a+1): no register shuffling in front of the HMMA.
#ifndef GNS_AMREG_MAX_L
#define GNS_AMREG_MAX_L 32   // latents up to this size keep a step's adj m additions in registers
#endif
#ifndef GNS_MMA_CHAINS
#define GNS_MMA_CHAINS 2   // measured on case300 training: 1 chain 23.8 ms, 2 chains 22.9 ms, 4 chains 24.9 ms (registers)
#endif
constexpr int kSA = 68;   // 64 + 4: the 8 lanes of a 128-bit phase (2 pair rows x 4 quads 8 floats apart) hit 32 distinct banks
__host__ __device__ constexpr int bwd_hid_floats() { return 8 * kSA; }
__host__ __device__ constexpr int bwd_srows_floats(int H) { return (H + 1) * kTS > bwd_hid_floats() ? (H + 1) * kTS : bwd_hid_floats(); }
__host__ __device__ constexpr int bwd_tile_floats(int H) { return kTS + bwd_hid_floats() + 16 * kTS + bwd_srows_floats(H); }

__host__ __device__ inline BwdSmem make_bwd_smem(int N, int E, int G, int L, int H, int PO, int nwarps,
                                                bool mglobal = false) {
  BwdSmem b{};
  const int NGs = bwd_bus_stride(N * G), EGs = row_stride(E * G);
  int o = 0;
  b.adj = o; o += (mglobal ? 4 : 4 + L) * NGs;   // large latent: m / adj m rows live in a global scratch instead
  b.nxt = o; o += 4 * NGs;
  b.lineg = o; o += 5 * EGs;
  b.adjD = o; o += NGs;
  (void)PO;
  b.tfloats = bwd_tile_floats(H);
  b.tiles = o; o += nwarps * b.tfloats;
  b.mbar = o; o += 4;
  b.total = o;
  return b;
}

struct BwdArgs {
  const float* params;      // packed [K][wstep]
  const float* buses; const float* lines; const float* gens;
  const float* ckpt;        // forward checkpoints [nbatch_f][K][(4+L)*NGs_f], grid-interleaved with G_f
  const float* pglob;       // [nbatch_f][K][G_f]
  const float* act;         // [nbatch_f][K][ActLayout.total] hidden activations kept by the forward kernel
  const float* grad_total; const float* grad_last; const float* grad_v; const float* grad_theta;
  float* gacc;              // [ctas*nwarps][K][FragLayout.step] per-warp gradient accumulators (zeroed by the host)
  float* mscratch;          // [ctas][2][L][NGs] m and adj m rows when they do not fit shared memory (L > 32)
  const uint16_t* topo;
  long long S;
  int N, Ns, E, Gn, K, NGQ, G, nbatch;
  int NGs, EGs;
  int acc_per_warp;         // 1: one accumulator block per warp (bit-reproducible); 0: one per CTA, shared by its warps
  int Gf, NGs_f, EGs_f;     // forward geometry of the checkpoints
  ActLayout al;
  unsigned char grp_of_warp[32];
  SmemPlan sm;
  BwdSmem bs;
  TopoOffsets to;
  float wk[kMaxK];
};

// ---- weight-gradient tiles on the tensor cores: mma.sync m16n8k8 TF32 with a 3-term split ----
// D[hid c][wide r] += sum_item hid_c[item] * row_r[item]: M = hidden columns (C <= 16), N = 8 wide rows per
// tile, K = the warp's 32 items.  Every FP32 operand x is split into big = x with the low 13 mantissa
// bits cleared (exactly a TF32 number) and small = x - big (exact in FP32; the tensor core keeps its
// top 10 mantissa bits), and D += big*big + big*small + small*big: the dropped terms are below 2^-20
// relative per product, against 2^-11 for plain TF32, so the FP32 parity tolerances hold.
// The k index of an MMA is a free permutation as long as A and B agree: k slot t (t+4) of MMA i is item
// 4t+i (16+4t+i), so both fragments come straight from two LDS.128 per row and need no shuffles.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& big, uint32_t& small) {
  big = __float_as_uint(x) & 0xffffe000u;
  small = __float_as_uint(x - __uint_as_float(big));
}

__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {   // fire-and-forget: no load, no scoreboard wait
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// gfrag: this call's block of the warp-private accumulator (FragLayout); only this warp ever adds to it,
// in program order, so the reductions are deterministic.
// CG: the wide rows live in global memory that this kernel updates with reductions: read them from L2 (ld.global.cg)
// Tiles [skip0, skip1) are left out (their wide rows are known to be all zeros: the latent entering step 0).
template <int C, int R, bool CG, class RowFn>
__device__ __forceinline__ void tile_gemm_mma(RowFn rowfn, const float* __restrict__ hidblk, float* __restrict__ gfrag,
                                              int skip0 = 0, int skip1 = 0) {
  static_assert(C <= 11, "hidden columns 11..15 of the m16 tile are not stored");
  const int lane = threadIdx.x & 31, gi = lane >> 2, t = lane & 3;
  // A fragments of the four MMAs: MMA m covers items a(t, m) = 16 (m / 2) + 4 t + 2 (m % 2) and a + 1
  uint32_t ab[4][4], as[4][4];            // [mma][a0..a3], big and small parts
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const float4 q = *reinterpret_cast<const float4*>(hidblk + gi * kSA + 2 * ((m >> 1) * 16 + 4 * t + (m & 1) * 2));
    split_tf32(q.x, ab[m][0], as[m][0]);
    split_tf32(q.z, ab[m][2], as[m][2]);
    if (C > 8) {
      split_tf32(q.y, ab[m][1], as[m][1]);
      split_tf32(q.w, ab[m][3], as[m][3]);
    } else {
      ab[m][1] = as[m][1] = ab[m][3] = as[m][3] = 0u;
    }
  }
  constexpr int NT = (R + 7) / 8;
  constexpr int UNR = NT <= 5 ? NT : 1;   // long tiles (latent_dim 64) stay rolled: unrolling them only buys spills
  // B fragments of tile nt: wide row nt*8+gi, items 4t..4t+3 and 16+4t..16+4t+3 (zeros past the last row)
  auto load_b = [&](int nt, float4& lo, float4& hi) {
    const int row = nt * 8 + gi;
    const bool rv = (nt * 8 + 8 <= R) || (row < R);
    const float* rp = rowfn(rv ? row : 0) + 4 * t;
    if constexpr (CG) {
      lo = __ldcg(reinterpret_cast<const float4*>(rp));
      hi = __ldcg(reinterpret_cast<const float4*>(rp + 16));
    } else {
      lo = *reinterpret_cast<const float4*>(rp);
      hi = *reinterpret_cast<const float4*>(rp + 16);
    }
    if (!rv) { lo = make_float4(0.f, 0.f, 0.f, 0.f); hi = lo; }
  };
  float4 nlo, nhi;                        // rolled loops (long tiles, rows possibly in global memory): one tile ahead
  if constexpr (UNR == 1) load_b(0, nlo, nhi);
#pragma unroll UNR
  for (int nt = 0; nt < NT; ++nt) {
    float4 blo, bhi;
    if constexpr (UNR != 1) {
      if (nt >= skip0 && nt < skip1) continue;
    }
    if constexpr (UNR == 1) {
      blo = nlo; bhi = nhi;
      if (nt + 1 < NT) load_b(nt + 1, nlo, nhi);
    } else {
      load_b(nt, blo, bhi);
    }
    const float bv[4][2] = {{blo.x, blo.y}, {blo.z, blo.w}, {bhi.x, bhi.y}, {bhi.z, bhi.w}};
    uint32_t bb[4][2], bs[4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      split_tf32(bv[m][0], bb[m][0], bs[m][0]);
      split_tf32(bv[m][1], bb[m][1], bs[m][1]);
    }
    float di[GNS_MMA_CHAINS][4];          // independent accumulator chains (HMMA latency)
#pragma unroll
    for (int c = 0; c < GNS_MMA_CHAINS; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) di[c][j] = 0.f;
#pragma unroll
    for (int m = 0; m < 4; ++m) mma_tf32(di[m % GNS_MMA_CHAINS], as[m], bb[m][0], bb[m][1]);
#pragma unroll
    for (int m = 0; m < 4; ++m) mma_tf32(di[m % GNS_MMA_CHAINS], ab[m], bs[m][0], bs[m][1]);
#pragma unroll
    for (int m = 0; m < 4; ++m) mma_tf32(di[m % GNS_MMA_CHAINS], ab[m], bb[m][0], bb[m][1]);
    float d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      d[j] = di[0][j];
#pragma unroll
      for (int c = 1; c < GNS_MMA_CHAINS; ++c) d[j] += di[c][j];
    }
    red_add_v2(gfrag + nt * kFragTile + lane * 2, d[0], d[1]);                               // hidden columns 0..7
    if (C > 8 && lane < 4 * (C - 8)) red_add_v2(gfrag + nt * kFragTile + 64 + lane * 2, d[2], d[3]);   // columns 8..C-1
  }
}

// one weight-gradient GEMM call: rows r < R (wide side) x C hidden columns over the warp's 32 items, added
// into the call's accumulator block
template <int C, int R, bool CG = false, class RowFn>
__device__ __forceinline__ void tile_gemm_r(RowFn rowfn, const float* __restrict__ hidblk, float* __restrict__ gfrag,
                                            int skip0 = 0, int skip1 = 0) {
  tile_gemm_mma<C, R, CG>(rowfn, hidblk, gfrag, skip0, skip1);
}

// Large latent dimensions (L > 32): state m and its adjoint (2 x L x Ns floats, 183 KB for L=64 on
// case300) do not fit next to the tiles, so both live in an L2-resident per-CTA global scratch with
// the same [feature][item] layout; they are only ever touched by the owning thread (the message
// uses the receiver's own latent) and by the in-place rows of the weight-gradient tiles.
template <int L, int H, bool MULTI, int TMAX>
__global__ void __launch_bounds__(TMAX, 1) gns_backward_kernel(const BwdArgs a) {
  constexpr bool MG = L > 32;
  constexpr WLayout W = make_wlayout(L, H, MULTI);
  constexpr FragLayout FL = make_frag_layout(L, H);
  constexpr int HP = pad4(H);
  // tile row map
  constexpr int T_ONES = 0;                              // wide row of ones (bias gradients)
  constexpr int T_HIDA = kTS;                            // hid block A
  constexpr int T_WIDE = T_HIDA + bwd_hid_floats();      // 16 wide rows
  constexpr int T_S = T_WIDE + 16 * kTS;                 // H+1 wide rows: the aggregate A and the in-degree (rows of dM / dc) ...
  constexpr int T_HIDB = T_S;                            // ... and, once those are dead (phi adjoint), hid block B
  static_assert(H + 1 + 5 <= 16 && H + 1 <= 16, "wide staging rows / one m16 tile");

  extern __shared__ __align__(16) float smem[];
  const int N = a.N, E = a.E, Gn = a.Gn, G = a.G, NGQ = a.NGQ, K = a.K;
  const int NG = a.NGs, EG = a.EGs, GnG = Gn * G;
  float* const s_state = smem + a.sm.state;
  float* const s_busc = smem + a.sm.busc;
  float* const s_genc = smem + a.sm.genc;
  float* const s_linef = smem + a.sm.linef;
  float* const s_y = smem + a.sm.yline;
  float* const s_trig = smem + a.sm.trig;
  float* const s_red = smem + a.sm.red;
  float* const s_w = smem + a.sm.weights;
  uint16_t* const s_topo = reinterpret_cast<uint16_t*>(smem + a.sm.topo);
  float* const s_adj = smem + a.sm.extra + a.bs.adj;
  float* const s_nxt = smem + a.sm.extra + a.bs.nxt;
  float* const s_lineg = smem + a.sm.extra + a.bs.lineg;
  float* const s_adjD = smem + a.sm.extra + a.bs.adjD;
  // rows of m / adj m: [i][NGs]
  float* const m_rows = MG ? a.mscratch + (size_t)blockIdx.x * 2 * L * a.NGs : s_state + 4 * a.NGs;
  float* const am_rows = MG ? a.mscratch + ((size_t)blockIdx.x * 2 + 1) * L * a.NGs : s_adj + 4 * a.NGs;

  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  int red_parity = 0;
  const int grp = a.grp_of_warp[warp];                 // bus group of this warp (sub-partition balancing)
  const int slot = grp * (32 / NGQ) + lane / NGQ;
  const int gq = lane % NGQ;
  const bool slot_on = slot < a.Ns;           // owns a bus slot (primary or twin)
  float* const tile = smem + a.sm.extra + a.bs.tiles + warp * a.bs.tfloats;
  float* const gacc_w = a.gacc + ((size_t)blockIdx.x * (a.acc_per_warp ? nwarps : 1) + (a.acc_per_warp ? warp : 0)) * ((size_t)K * FL.step);

  // zero all dynamic shared memory once: padding lanes and tail rows must hold finite values
  for (int i = tid; i < a.sm.total_floats; i += T) smem[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < a.to.total / 2; i += T)
    reinterpret_cast<uint32_t*>(s_topo)[i] = reinterpret_cast<const uint32_t*>(a.topo)[i];
  tile[T_ONES + lane] = 1.f;
  // per-step weights arrive by one bulk copy (cp.async.bulk + mbarrier) issued as soon as the previous
  // step's MLP phase is over, so the copy runs under the physics adjoint instead of in front of it
  uint64_t* const s_mbar = reinterpret_cast<uint64_t*>(smem + a.sm.extra + a.bs.mbar);
  uint32_t w_phase = 0;
  auto weights_issue = [&](int k) {   // one thread, after a block barrier that follows the last read of s_w
    fence_proxy_async();
    mbar_expect_tx(s_mbar, W.wstep * 4);
    bulk_g2s(s_w, a.params + (size_t)k * W.wstep, W.wstep * 4, s_mbar);
  };
  const uint16_t* const t_fi = s_topo + a.to.fi;
  const uint16_t* const t_ti = s_topo + a.to.ti;
  const uint16_t* const t_fa = s_topo + a.to.fa;
  const uint16_t* const t_ta = s_topo + a.to.ta;
  const uint16_t* const t_inb = s_topo + a.to.in_b;
  const uint16_t* const t_ine = s_topo + a.to.in_e;
  const uint16_t* const t_infe = s_topo + a.to.in_fe;
  const uint16_t* const t_ini = s_topo + a.to.in_ids;
  const uint16_t* const t_inp = s_topo + a.to.in_pos;
  const uint16_t* const t_outb = s_topo + a.to.out_b;
  const uint16_t* const t_oute = s_topo + a.to.out_e;
  const uint16_t* const t_outi = s_topo + a.to.out_ids;
  const uint16_t* const t_genb = s_topo + a.to.gen_b;
  const uint16_t* const t_gene = s_topo + a.to.gen_e;
  const uint16_t* const t_geni = s_topo + a.to.gen_ids;
  const uint16_t* const t_ext = s_topo + a.to.ext_of;
  const uint16_t* const t_rank = s_topo + a.to.rank_of;
  const uint16_t* const t_prim = s_topo + a.to.prim_of;
  const uint16_t* const t_gsz = s_topo + a.to.gsz;
  if (tid == 0) { mbar_init(s_mbar, 1); fence_proxy_async(); }
  __syncthreads();
  if (tid == 0 && (int)blockIdx.x < a.nbatch) weights_issue(K - 1);

  // slot bookkeeping: a bus's state / adjoint live in its primary slot; twins only help with its lines
  const int sl = slot_on ? slot : 0;
  const int n = t_prim[sl];                             // primary slot of this slot's bus
  const bool bus_on = slot_on && n == sl;               // this thread owns the bus
  const int nb = n * G + gq;                            // offset of (bus, grid gq) inside a [Ns][G] row
  const int jb = slot * G + gq;                         // offset when `slot` is used as an alias LINE id (< N)
  const int gsz = slot_on ? (int)t_gsz[sl] : 1;
  const bool warp_has_twins = __any_sync(0xffffffffu, gsz > 1);
  const int leader_lane = lane - (sl - n) * NGQ;        // lane of the primary slot of this twin group
  const int e_in0 = slot_on ? (int)t_inb[sl] : 0, e_in1 = slot_on ? (int)t_ine[sl] : 0;   // lines this slot walks
  const int e_full1 = bus_on ? (int)t_infe[sl] : e_in0;                                     // end of the bus's in-list
  const int e_out0 = bus_on ? (int)t_outb[sl] : 0, e_out1 = bus_on ? (int)t_oute[sl] : 0;
  const int j0 = bus_on ? (int)t_genb[sl] : 0, j1 = bus_on ? (int)t_gene[sl] : 0;
  const int deg = e_in1 - e_in0;                        // lines walked by this slot
  const float degf = (float)(e_full1 - e_in0);          // in-degree of the bus (primary)
  const bool is_gen = j1 > j0;
  const int warp_max_deg = __reduce_max_sync(0xffffffffu, deg);

  auto stage = [&](int off, int row, float val) { tile[off + row * kTS + lane] = bus_on ? val : 0.f; };          // wide row
  auto stage_hid = [&](int blk, int c, float val) { tile[blk + (c & 7) * kSA + lane * 2 + (c >> 3)] = bus_on ? val : 0.f; };

  for (int batch = blockIdx.x; batch < a.nbatch; batch += gridDim.x) {
    const long long g0 = (long long)batch * G;
    const long long gme = g0 + gq;                       // this thread's grid
    const bool grid_ok = gme < a.S;
    const long long gld = grid_ok ? gme : a.S - 1;       // replicated tail grid (zero upstream gradient)
    const long long bf = gld / a.Gf;                     // forward CTA-batch and column of that grid
    const int cf = (int)(gld - bf * a.Gf);
    const size_t ck_stride = (size_t)(4 + L) * a.NGs_f;
    const float* const ck_base = a.ckpt + (size_t)bf * K * ck_stride + cf;
    const float gtot = grid_ok ? a.grad_total[gld] : 0.f;
    const float glast = (grid_ok && a.grad_last) ? a.grad_last[gld] : 0.f;

    // ---------------- inputs (same staging as the forward kernel) ----------------
    load_block(a.buses, s_busc, g0, a.S, G, N, 6, 2, NG, t_rank);
    load_block(a.lines, s_linef, g0, a.S, G, E, 7, 2, EG, nullptr);
    load_block(a.gens, s_genc, g0, a.S, G, Gn, 7, 1, GnG, nullptr);
    __syncthreads();
    float part4[4] = {0.f, 0.f, 0.f, 0.f};
    if (bus_on) part4[0] = s_busc[0 * NG + nb];
    if (slot < N) {                                                        // `slot` is an alias LINE id here
      const float r = s_linef[0 * EG + jb], x = s_linef[1 * EG + jb];
      s_y[jb] = 1.0f / sqrtf(r * r + x * x);
      s_y[NG + jb] = 1.0f / s_linef[3 * EG + jb];           // 1 / tau, like the forward kernel
    }
    for (int it = tid; it < Gn * NGQ; it += T) {
      part4[1] += s_genc[2 * GnG + it];
      part4[2] += s_genc[1 * GnG + it];
      part4[3] += s_genc[0 * GnG + it];
    }
    float sums[3][1] = {{part4[1]}, {part4[2]}, {part4[3]}};
    block_sum_multi<1, 3>(sums, s_red, NGQ, red_parity);
    const float sPset = sums[0][0], sPmin = sums[1][0], sPmax = sums[2][0];

    // ---------------- adjoint of the outputs; v', theta', dP', dQ' of the final state ----------------
    if (bus_on) {
      const int ext = t_ext[n];
      s_adj[0 * NG + nb] = (grid_ok && a.grad_v) ? a.grad_v[gld * N + ext] : 0.f;
      s_adj[1 * NG + nb] = (grid_ok && a.grad_theta) ? a.grad_theta[gld * N + ext] : 0.f;
      s_adj[2 * NG + nb] = 0.f; s_adj[3 * NG + nb] = 0.f;
      for (int i = 0; i < L; ++i) am_rows[i * NG + nb] = 0.f;
      const float* ck = ck_base + (size_t)(K - 1) * ck_stride + (size_t)n * a.Gf;
      for (int i = 0; i < 4; ++i) s_nxt[i * NG + nb] = ck[(size_t)i * a.NGs_f];
    }
    __syncthreads();

    for (int k = K - 1; k >= 0; --k) {
      // ---------------- stage weights and state_k ----------------
      if (bus_on) {
        if (k >= 1) {
          const float* ck = ck_base + (size_t)(k - 1) * ck_stride + (size_t)n * a.Gf;
          // asynchronous copies straight into shared memory; they are waited for in front of the barrier of the
          // lambda-coupling reduction below, under the trig of the alias lines
          for (int i = 0; i < 4; ++i) cp_async4(s_state + i * NG + nb, ck + (size_t)i * a.NGs_f);
          if constexpr (!MG) {
            for (int i = 0; i < L; ++i) cp_async4(m_rows + i * NG + nb, ck + (size_t)(4 + i) * a.NGs_f);
          } else if (a.Gf != 1) {     // (large latents with one grid per forward column read m in place, see rows_m)
            for (int i = 0; i < L; ++i) m_rows[i * NG + nb] = ck[(size_t)(4 + i) * a.NGs_f];
          }
          cp_async_commit();
        } else {   // state before step 0 (ref GNS/main.py:141-152)
          float vv = 0.f, pg = 0.f, qg = 0.f;
          for (int j = j0; j < j1; ++j) {
            const int gid = t_geni[j];
            vv += s_genc[3 * GnG + gid * G + gq];
            pg += s_genc[5 * GnG + gid * G + gq];
            qg += s_genc[4 * GnG + gid * G + gq];
          }
          vv = (vv == 0.f) ? 1.f : vv;
          s_state[0 * NG + nb] = vv;
          s_state[1 * NG + nb] = 0.f;
          s_state[2 * NG + nb] = pg - s_busc[0 * NG + nb] - s_busc[2 * NG + nb] * (vv * vv);
          s_state[3 * NG + nb] = qg - s_busc[1 * NG + nb] + s_busc[3 * NG + nb] * (vv * vv);
          for (int i = 0; i < L; ++i) m_rows[i * NG + nb] = 0.f;
        }
      }

      // ---------------- physics adjoint (a): dP' adjoint, lambda coupling, alias-line trig ----------------
      const float pglob = a.pglob[((size_t)bf * K + k) * a.Gf + cf];
      const bool lo_branch = pglob < sPset;
      const float lam = lo_branch ? (pglob - sPmin) / (2.f * (sPset - sPmin))
                                  : (pglob - 2.f * sPset + sPmax) / (2.f * (sPmax - sPset));
      const bool lo_arm = lam < 0.5f;
      float gdP = 0.f, vprime = 0.f, Gs = 0.f;
      float part[1] = {0.f};
      if (bus_on) {
        const float coef = (gtot * a.wk[k] + ((k == K - 1) ? glast : 0.f)) * (2.0f / (float)N);
        gdP = s_adj[2 * NG + nb] + coef * s_nxt[2 * NG + nb];
        s_adj[2 * NG + nb] = gdP;                       // published for the line phase (upstream of p_from / p_to)
        float cs = 0.f;
        for (int j = j0; j < j1; ++j) {
          const int gid = t_geni[j];
          const float Pmax = s_genc[0 * GnG + gid * G + gq], Pmin = s_genc[1 * GnG + gid * G + gq];
          const float Pset = s_genc[2 * GnG + gid * G + gq];
          cs += lo_arm ? 2.f * (Pset - Pmin) : 2.f * (Pmax - Pset);
        }
        part[0] = gdP * cs;
        vprime = s_nxt[0 * NG + nb];
        Gs = s_busc[2 * NG + nb];
      }
      if (slot < N) {   // alias line j == slot: D_j = theta'[f_j] - theta'[t_j]
        const float d = s_nxt[1 * NG + (int)t_fi[slot] * G + gq] - s_nxt[1 * NG + (int)t_ti[slot] * G + gq];
        float sd, cd;
        fast_sincos(d, sd, cd);
        s_trig[0 * NG + jb] = d; s_trig[1 * NG + jb] = sd; s_trig[2 * NG + jb] = cd;
      }
      cp_async_wait_all();                                      // this thread's state_k rows have landed
      block_sum_per_grid<1>(part, s_red, NGQ, red_parity);      // its barrier also publishes s_state, s_trig, gdP
      const float adj_pg = part[0] / (lo_branch ? 2.f * (sPset - sPmin) : 2.f * (sPmax - sPset));

      // ---------------- physics adjoint (b): per-line partials ----------------
      for (int it = tid; it < E * NGQ; it += T) {
        const int e = it / NGQ;
        const int fi = t_fi[e], ti = t_ti[e], fa = t_fa[e], ta = t_ta[e];
        const float vf = s_nxt[0 * NG + fi * G + gq], vt = s_nxt[0 * NG + ti * G + gq];
        const float thf = s_nxt[1 * NG + fi * G + gq], tht = s_nxt[1 * NG + ti * G + gq];
        const float g_pf = s_adj[2 * NG + ti * G + gq];      // p_from lands on the receiving bus
        const float g_pt = s_adj[2 * NG + fi * G + gq];      // p_to lands on the sending bus
        const float Yf = s_y[fa * G + gq], itf = s_y[NG + fa * G + gq], shf = s_linef[4 * EG + fa * G + gq];
        const float Df = s_trig[0 * NG + fa * G + gq], sDf = s_trig[1 * NG + fa * G + gq], cDf = s_trig[2 * NG + fa * G + gq];
        const float Yt = s_y[ta * G + gq], itt = s_y[NG + ta * G + gq], sht = s_linef[4 * EG + ta * G + gq];
        const float DB = s_trig[0 * NG + ta * G + gq], sDB = s_trig[1 * NG + ta * G + gq], cDB = s_trig[2 * NG + ta * G + gq];
        const float a1 = thf - tht - Df - shf;
        const float a2 = tht - thf - Df + shf;
        const float a3 = tht - thf + DB - sht;               // delta_ji[dst] = -D_B
        float s1, c1, s2, c2, s3, c3;
        fast_sincos(a1, s1, c1);
        fast_sincos(a2, s2, c2);
        fast_sincos(a3, s3, c3);
        const float yft = Yf * itf, yftt = Yf * (itf * itf), ytt = Yt * itt;
        const float t1 = vf * vt * yft, u1 = vt * vf * ytt;
        const float sDt = -sDB;
        const float ss = s1 + s2;
        const float inner = t1 * ss + vf * yftt * sDf + vt * vt * Yf * sDf;       // |.| of ref GNS/main.py:41
        const float g_in = adj_pg * ((inner > 0.f) ? 1.f : ((inner < 0.f) ? -1.f : 0.f));
        const float gvf = g_in * (vt * yft * ss + yftt * sDf) + g_pf * (vt * yft * s1 + 2.f * vf * yftt * sDf) +
                          g_pt * (vt * ytt * s3);
        const float gvt = g_in * (vf * yft * ss + 2.f * vt * Yf * sDf) + g_pf * (vf * yft * s1) +
                          g_pt * (vf * ytt * s3 + 2.f * vt * Yt * sDt);
        const float G1 = (g_in + g_pf) * t1 * c1, G2 = g_in * t1 * c2, G3 = g_pt * u1 * c3;
        const float gth = G1 - G2 - G3;
        const float gDA = -G1 - G2 + (g_in * (vf * yftt + vt * vt * Yf) + g_pf * vf * vf * yftt) * cDf;
        const float gDB = G3 - g_pt * vt * vt * Yt * cDB;
        const int eo = e * G + gq;
        s_lineg[0 * EG + eo] = gvf;
        s_lineg[1 * EG + eo] = gvt;
        s_lineg[2 * EG + eo] = gth;
        s_lineg[3 * EG + eo] = gDA;
        s_lineg[4 * EG + eo] = gDB;
      }
      __syncthreads();

      // ---------------- physics adjoint (c): CSR gathers replace the forward scatter-adds ----------------
      float adjv = 0.f, adjth = 0.f;
      if (bus_on) {
        float sv = 0.f, sth = 0.f, sD = 0.f;
        for (int e = e_out0; e < e_out1; ++e) {
          const int eo = (int)t_outi[e] * G + gq;
          sv += s_lineg[0 * EG + eo]; sth += s_lineg[2 * EG + eo]; sD += s_lineg[3 * EG + eo];
        }
        for (int e = e_in0; e < e_full1; ++e) {
          const int eo = (int)t_ini[e] * G + gq;
          sv += s_lineg[1 * EG + eo]; sth -= s_lineg[2 * EG + eo]; sD += s_lineg[4 * EG + eo];
        }
        adjv = s_adj[0 * NG + nb] + sv + gdP * (-2.f * Gs * vprime) + adj_pg * (2.f * vprime * Gs);
        adjth = s_adj[1 * NG + nb] + sth;
        s_adjD[(int)t_ext[n] * G + gq] = sD;               // alias line id == external bus number
      }
      __syncthreads();
      if (bus_on) {                                        // D_e = theta[f_e] - theta[t_e] for alias lines e < N
        for (int e = e_out0; e < e_out1; ++e) { const int l = t_outi[e]; if (l < N) adjth += s_adjD[l * G + gq]; }
        for (int e = e_in0; e < e_full1; ++e) { const int l = t_ini[e]; if (l < N) adjth -= s_adjD[l * G + gq]; }
      }
      // adjv / adjth now hold d loss / d v', d theta' of this thread's bus (registers).

      // ---------------- MLP adjoint + weight gradients (bus-centric, warp tiles) ----------------
      mbar_wait(s_mbar, w_phase);                      // this step's weights have landed
      w_phase ^= 1;
      float* const gk = gacc_w + (size_t)k * FL.step;
      {
        float* adjrow = s_adj + nb;
        float* adjm = am_rows + nb;
        float adj4[4] = {0.f, 0.f, 0.f, 0.f};
        const float* rows_state = s_state + 32 * grp;          // [f][item] rows of this warp's 32 items
        // Large latents: with one grid per forward column the checkpoint's m rows already are [feature][item], so the
        // tiles read them in place (no copy into the scratch); step 0 reads the zeroed scratch rows.
        const bool m_in_ckpt = MG && a.Gf == 1 && k >= 1;
        const float* rows_m = m_in_ckpt ? ck_base + (size_t)(k - 1) * ck_stride + (size_t)4 * a.NGs_f + 32 * grp : m_rows + 32 * grp;
        const int MS = m_in_ckpt ? a.NGs_f : NG;               // row stride of rows_m
        const float* rows_am = am_rows + 32 * grp;
        // this grid's column of the activations the forward kernel kept for step k (no recompute here)
        const float* const act_k = a.act + ((size_t)bf * K + k) * (size_t)a.al.total;
        const size_t RB = (size_t)a.al.rb, RL = (size_t)a.al.rl;
        const int AIS = a.al.is, ALS = a.al.ls;     // item strides of the two activation layouts (locals: read once per step)
        float adjA[H];
#pragma unroll
        for (int o = 0; o < H; ++o) adjA[o] = 0.f;
        // small latents: this step's additions to adj m stay in registers and are folded in once at the end
        // (the only readers of adj m' inside the step, the m-net output layer and its tile, come first)
        constexpr bool AMREG = (L <= GNS_AMREG_MAX_L);
        float amr[AMREG ? L : 1];
#pragma unroll
        for (int i = 0; i < (AMREG ? L : 1); ++i) amr[i] = 0.f;

        // line activations (h1, h2 of one phi evaluation) travel one iteration ahead of their use: the loads of
        // iteration 0 are issued in front of the L-net's first-layer tile, those of it+1 at the top of iteration it
        // (the iteration-major in_pos columns are read once, so there is no L1 reuse to wait for)
        float nh[H];
        uint32_t nbits = 0u;          // slope bits of h2
        const bool quads = a.al.gs != 1;       // grid-major layout: quads of rows (else plain rows)
        auto load_line = [&](const float* actl, int it) {
          const int el = (int)t_inp[(slot_on && it < deg) ? e_in0 + it : 0] * ALS + cf * a.al.gl;
          if (quads) {
            ldg_rows4<H>(nh, actl, RL, el);
          } else {
#pragma unroll
            for (int o = 0; o < H; ++o) nh[o] = __ldg(actl + o * RL + el);
          }
          nbits = __ldg(reinterpret_cast<const uint32_t*>(actl + H * RL) + el);
        };
        // adjoint of one phi net given adjA: line loop, dW2 / db2 / dW1f, adjP, dW1m / db1, adj m
        auto phi_backward = [&](const float* wphi, float* gphi, const float* actl) {
          float adjP[H];
#pragma unroll
          for (int o = 0; o < H; ++o) adjP[o] = 0.f;
          if (warp_has_twins) {   // twins take the aggregate's adjoint from the bus owner
#pragma unroll
            for (int o = 0; o < H; ++o) adjA[o] = __shfl_sync(0xffffffffu, adjA[o], leader_lane);
          }
          for (int it = 0; it < warp_max_deg; ++it) {
            const bool live = slot_on && (it < deg);
            const float* wp = wphi + opaque_zero();
            const int e = live ? e_in0 + it : 0;                 // position in the in-list
            float h1[H], d2[H][1], d1[H], feat[5];
#pragma unroll
            for (int o = 0; o < H; ++o) h1[o] = nh[o];                          // fetched one iteration ahead
            const uint32_t h2bits = nbits;
            if (it + 1 < warp_max_deg) load_line(actl, it + 1);
            const float* lf = s_linef + (int)t_ini[e] * G + gq;
#pragma unroll
            for (int c = 0; c < 5; ++c) feat[c] = lf[c * EG];
#pragma unroll
            for (int o = 0; o < H; ++o) d2[o][0] = live ? adjA[o] * (((h2bits >> o) & 1u) ? 1.f : kSlope) : 0.f;
#pragma unroll
            for (int j = 0; j < H; ++j) {
              float t[1] = {0.f};
              row_dot<H, HP, 1>(t, d2, wp + W.phi_w2 + j * HP);
              d1[j] = t[0] * lrelu_grad(h1[j]);
              adjP[j] += d1[j];
            }
            __syncwarp();
#pragma unroll
            for (int o = 0; o < H; ++o) {
              tile[T_HIDA + (o & 7) * kSA + lane * 2 + (o >> 3)] = d2[o][0];
              tile[T_HIDB + (o & 7) * kSA + lane * 2 + (o >> 3)] = d1[o];
              tile[T_WIDE + o * kTS + lane] = live ? h1[o] : 0.f;
            }
#pragma unroll
            for (int c = 0; c < 5; ++c) tile[T_WIDE + (H + 1 + c) * kTS + lane] = live ? feat[c] : 0.f;
            __syncwarp();
            // dW2^T[j][o] += h1[j] d2[o];  db2[o] += d2[o]
            tile_gemm_r<H, H + 1>([&](int r) { return tile + (r < H ? T_WIDE + r * kTS : T_ONES); }, tile + T_HIDA, gphi + FL.w2l);
            // dW1f^T[c5][o] += feat[c5] d1[o]
            tile_gemm_r<H, 5>([&](int r) { return tile + T_WIDE + (H + 1 + r) * kTS; }, tile + T_HIDB, gphi + FL.w1f);
          }
          if (warp_has_twins) {   // the bus owner needs the sum over its twins
#pragma unroll
            for (int d = 1; d < 4; d *= 2) {
#pragma unroll
              for (int o = 0; o < H; ++o) {
                const float t = __shfl_xor_sync(0xffffffffu, adjP[o], d * NGQ);
                if (gsz > d) adjP[o] += t;
              }
            }
          }
          __syncwarp();
#pragma unroll
          for (int o = 0; o < H; ++o) stage_hid(T_HIDA, o, adjP[o]);
          __syncwarp();
          // dW1m^T[i][o] += m[i] adjP[o];  db1[o] += adjP[o]
          // (step 0: the latent rows are zeros, ref GNS/main.py:152 - the tiles made of them only add zeros)
          tile_gemm_r<H, L + 1>([&](int r) { return r < L ? rows_m + r * MS : tile + T_ONES; }, tile + T_HIDA, gphi + FL.w1m, 0,
                                k == 0 ? L / 8 : 0);
          {
            float (&pv)[H][1] = reinterpret_cast<float (&)[H][1]>(adjP);
            if constexpr (AMREG) {
#pragma unroll
              for (int i = 0; i < L; ++i) {
                float t[1] = {amr[i]};
                row_dot<H, HP, 1>(t, pv, wphi + W.phi_w1m + i * HP);
                amr[i] = t[0];
              }
            } else {
#pragma unroll 4
              for (int i = 0; i < L; ++i) {
                float t[1] = {0.f};
                row_dot<H, HP, 1>(t, pv, wphi + W.phi_w1m + i * HP);
                if (bus_on) atomicAdd(adjm + i * NG, t[0]);   // global scratch (L > 32): fire-and-forget reduction, no load
              }
            }
          }
        };

#pragma unroll 1
        for (int qq = 0; qq < 3; ++qq) {
          const int q = (qq == 0) ? 2 : qq - 1;      // m-net first: it reads adj m' before anyone adds to it
          if (q == 2 && k == K - 1) continue;        // the last step's m net is dead (adj m' = 0; the forward skips it too)
          const float* wphi = s_w + (MULTI ? q * W.phi_size : 0);
          const float* wln = s_w + W.off_ln[0] + q * W.ln_size_s;
          float* gphi = gk + (MULTI ? q * FL.net : 0);   // accumulator blocks of this pair (fragment order)
          float* gln = gk + q * FL.net;
          const float* wmf = s_w + W.off_mf[0] + q * W.mf_size;     // fused block M (H rows) and c (1 row)

          // ---- activations of this bus and pair kept by the forward kernel: A, h1, h2 of the L-net ----
          float h1L[H], h2L[H];
          {
            const float* ab = act_k + (size_t)(q * 3 * H) * RB;
            const int eb = cf * a.al.gs + n * AIS;
            float Aq[H];
            if (quads) {
              ldg_rows4<H>(Aq, ab, RB, eb);
              ldg_rows4<H>(h1L, ab + H * RB, RB, eb);
              ldg_rows4<H>(h2L, ab + 2 * H * RB, RB, eb);
            } else {
#pragma unroll
              for (int o = 0; o < H; ++o) {
                Aq[o] = __ldg(ab + o * RB + eb); h1L[o] = __ldg(ab + (H + o) * RB + eb); h2L[o] = __ldg(ab + (2 * H + o) * RB + eb);
              }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < H; ++j) stage(T_S, j, Aq[j]);        // wide rows of dM
            stage(T_S, H, degf);                                     // wide row of dc
          }

          // ---- output layer adjoint and its weight gradient ----
          float dh2[H][1];
#pragma unroll
          for (int o = 0; o < H; ++o) dh2[o][0] = 0.f;
          if (q < 2) {
            const float g = (q == 0) ? (is_gen ? 0.f : adjv) : adjth;
            float gv[1] = {g};
            row_axpy<H, HP, 1>(dh2, gv, wln + W.ln_wo);
#pragma unroll
            for (int o = 0; o < H; ++o) stage(T_WIDE, o, h2L[o]);
            stage_hid(T_HIDA, 0, g);
            __syncwarp();
            // dWout[j] += g h2[j];  dbout += g      (rows = [h2 (H), ones], one column g)
            tile_gemm_r<1, H + 1>([&](int r) { return tile + (r < H ? T_WIDE + r * kTS : T_ONES); }, tile + T_HIDA, gln + FL.out);
          } else {
            constexpr int UG = MG ? 8 : 2;      // adj m' in the global scratch: more L2 loads in flight
#pragma unroll UG
            for (int i = 0; i < L; ++i) {
              float gm[1] = {bus_on ? (MG ? __ldcg(adjm + i * NG) : adjm[i * NG]) : 0.f};   // (the scratch is updated by reductions in L2)
              row_axpy<H, HP, 1>(dh2, gm, wln + W.ln_wo + i * HP);
            }
#pragma unroll
            for (int o = 0; o < H; ++o) stage_hid(T_HIDA, o, h2L[o]);
            stage_hid(T_HIDA, H, 1.f);
            __syncwarp();
            // dWout[i][j] += adjm'[i] h2[j];  dbout[i] += adjm'[i]    (rows = adj m' rows in place)
            tile_gemm_r<H + 1, L, MG>([&](int r) { return rows_am + r * NG; }, tile + T_HIDA, gln + FL.out);
          }
          __syncwarp();
          // ---- second layer ----
          float d2[H][1], d1[H][1];
#pragma unroll
          for (int o = 0; o < H; ++o) {
            d2[o][0] = dh2[o][0] * lrelu_grad(h2L[o]);
            stage_hid(T_HIDA, o, d2[o][0]);
            stage(T_WIDE, o, h1L[o]);
          }
          __syncwarp();
          tile_gemm_r<H, H + 1>([&](int r) { return tile + (r < H ? T_WIDE + r * kTS : T_ONES); }, tile + T_HIDA, gln + FL.w2);
#pragma unroll
          for (int j = 0; j < H; ++j) {
            float t[1] = {0.f};
            row_dot<H, HP, 1>(t, d2, wln + W.ln_w2 + j * HP);
            d1[j][0] = t[0] * lrelu_grad(h1L[j]);
          }
          __syncwarp();
          // ---- first layer: dW1^T[i][o] += x[i] d1[o] with x = [v,theta,dP,dQ, m, S], db1 += d1 ----
#pragma unroll
          for (int o = 0; o < H; ++o) stage_hid(T_HIDA, o, d1[o][0]);
          __syncwarp();
          const float* const actl = act_k + a.al.line_off + (size_t)((MULTI ? q : 0) * (H + 1)) * RL;
          if ((MULTI || qq == 2) && warp_max_deg > 0) load_line(actl, 0);
          // rows: state (4+L, in place), A (H) and deg (1) from the tile, ones -> dW1[:4+L], dM, dc, db1
          tile_gemm_r<H, 4 + L + H + 2>(
              [&](int r) {
                return r < 4 ? rows_state + r * NG
                       : r < 4 + L ? rows_m + (r - 4) * MS
                                 : (r < 4 + L + H + 1 ? tile + T_S + (r - 4 - L) * kTS : tile + T_ONES);
              },
              tile + T_HIDA, gln + FL.w1, 1, k == 0 ? (4 + L) / 8 : 0);      // (step 0: tiles of zero latent rows left out)
          __syncwarp();
          // ---- dX of the first layer ----
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float t[1] = {0.f};
            row_dot<H, HP, 1>(t, d1, wln + W.ln_w1 + i * HP);
            adj4[i] += t[0];
          }
          if constexpr (AMREG) {
#pragma unroll
            for (int i = 0; i < L; ++i) {
              float t[1] = {amr[i]};
              row_dot<H, HP, 1>(t, d1, wln + W.ln_w1 + (4 + i) * HP);
              amr[i] = t[0];
            }
          } else {
#pragma unroll 4
            for (int i = 0; i < L; ++i) {
              float t[1] = {0.f};
              row_dot<H, HP, 1>(t, d1, wln + W.ln_w1 + (4 + i) * HP);
              if (bus_on) atomicAdd(adjm + i * NG, t[0]);
            }
          }
          if (MULTI) {
#pragma unroll
            for (int o = 0; o < H; ++o) adjA[o] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < H; ++j) {       // adjoint of the aggregate through the fused block: adjA[j] += M[j] . d1
            float t[1] = {0.f};
            row_dot<H, HP, 1>(t, d1, wmf + j * HP);
            adjA[j] += t[0];
          }
          __syncwarp();
          if (MULTI || qq == 2)
            phi_backward(wphi, gphi, actl);
        }
        if constexpr (AMREG) {
          if (bus_on) {
#pragma unroll
            for (int i = 0; i < L; ++i) adjm[i * NG] += amr[i];
          }
        }
        if (bus_on) {
          adjrow[0 * NG] = adjv + adj4[0];
          adjrow[1 * NG] = adjth + adj4[1];
          adjrow[2 * NG] = adj4[2];
          adjrow[3 * NG] = adj4[3];
        }
      }
      __syncthreads();
      if (tid == 0) {                                  // s_w is free: fetch the next step's weights (or the next batch's first)
        if (k >= 1) weights_issue(k - 1);
        else if (batch + (int)gridDim.x < a.nbatch) weights_issue(K - 1);
      }
      // state_k's (v, theta, dP, dQ) are the primes of step k-1
      if (bus_on && k >= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s_nxt[i * NG + nb] = s_state[i * NG + nb];
      }
      __syncthreads();
    }  // k
  }  // batch
}

}  // namespace gns
