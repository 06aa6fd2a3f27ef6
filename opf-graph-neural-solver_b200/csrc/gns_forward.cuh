// gns_forward.cuh — persistent K-step forward kernel (one CTA = G grids, all K steps).
//
// Restates ref GNS/main.py:140-202 (+ :34-104 for the physics) for G grids at a time.
// Per step and per bus the whole message-passing pipeline runs inside ONE thread:
//   P   = W1[:, :L] m[n] + b1                       (receiver-only message, quirk Q2)
//   A   = sum_{lines e into n} lrelu(W2 lrelu(P + W1[:, L:] feat_e) + b2)
//   S   = W4 A + deg(n) b4                           (= scatter_add of phi outputs, by linearity)
//   out = L_net([v, theta, dP, dQ, m, S])
// so the bus aggregation needs neither atomics nor a barrier; the lines entering a bus are
// walked through the plan's CSR in ascending line order (the reference's scatter order).
// Then, after one barrier, the Kirchhoff physics runs per line and per bus.
#pragma once
#include "gns_common.cuh"

namespace gns {

#ifndef GNS_GRAD_MREG
#define GNS_GRAD_MREG 1   // training variant: keep the latent in registers too
#endif

// GRADV: 0 = inference, 1 = training with grid-major activation rows, 2 = training with interleaved rows (ActLayout),
//        3 = training with the per-grid block checkpoints of the warp-specialised backward kernel (Act2Layout)
template <int L, int H, bool MULTI, int VG, int TMAX, int GRADV>
__global__ void __launch_bounds__(TMAX, 1) gns_forward_kernel(const FwdArgs a) {
  constexpr bool GRAD = GRADV != 0;
  constexpr bool INTER = GRADV == 2;
  constexpr bool V3 = GRADV == 3;
  constexpr WLayout W = make_wlayout(L, H, MULTI);
  constexpr int HP = pad4(H);
  constexpr int PO = MULTI ? L : 1;
  using IO = VecIO<VG>;

  extern __shared__ __align__(16) float smem[];
  const int N = a.N, E = a.E, Gn = a.Gn, G = a.G, NGQ = a.NGQ, K = a.K;
  const int NG = a.NGs, EG = a.EGs, GnG = Gn * G;   // padded row strides
  float* const s_state = smem + a.sm.state;
  float* const s_busc = smem + a.sm.busc;
  float* const s_genc = smem + a.sm.genc;
  float* const s_linef = smem + a.sm.linef;
  float* const s_y = smem + a.sm.yline;
  float* const s_trig = smem + a.sm.trig;
  float* const s_flow = smem + a.sm.flows;
  float* const s_gsum = smem + a.sm.gsum;
  float* const s_red = smem + a.sm.red;
  // current step's weights: offset select on the shared-memory base (keeps the address space provable;
  // a pointer picked from an array of pointers degrades every weight load to a generic LD)
  const float* s_w = smem + a.sm.weights;
  uint16_t* const s_topo = reinterpret_cast<uint16_t*>(smem + a.sm.topo);

  const int tid = threadIdx.x, T = blockDim.x;
  int red_parity = 0;
  // bus slot (internal order) in bus phases; the warp -> bus-group map balances the 4 sub-partitions
  const int slot = (int)a.grp_of_warp[tid >> 5] * (32 / NGQ) + (tid & 31) / NGQ;
  const int gq = (tid & 31) % NGQ;
  const int gcol = gq * VG;
  const bool slot_on = slot < a.Ns;     // this thread owns a bus slot (primary or twin)

  // topology indices: once per CTA
  for (int i = tid; i < a.to.total / 2; i += T)
    reinterpret_cast<uint32_t*>(s_topo)[i] = reinterpret_cast<const uint32_t*>(a.topo)[i];
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
  const uint16_t* const t_brank = s_topo + a.to.brank;
  __syncthreads();
  // slot bookkeeping (constant over batches)
  const int sl = slot_on ? slot : 0;
  const int pslot = t_prim[sl];                       // primary slot of this slot's bus: owns the state
  const bool prim = slot_on && pslot == sl;           // this thread owns the bus (twins only help with lines)
  const bool bus_on = prim;
  const int gsz = slot_on ? (int)t_gsz[sl] : 1;       // twin group size (1, 2, 4), group aligned to gsz
  const bool warp_has_twins = __any_sync(0xffffffffu, gsz > 1);
  const int e_in0 = slot_on ? (int)t_inb[sl] : 0, e_in1 = slot_on ? (int)t_ine[sl] : 0;   // lines this slot walks
  const int e_full1 = prim ? (int)t_infe[sl] : e_in0;                                       // end of the bus's in-list
  const int e_out0 = prim ? (int)t_outb[sl] : 0, e_out1 = prim ? (int)t_oute[sl] : 0;
  const int j0 = prim ? (int)t_genb[sl] : 0, j1 = prim ? (int)t_gene[sl] : 0;

  if (!GRAD && a.compact) {   // constant columns of the case, once per CTA (the per-batch unpack never touches these rows)
    for (int i = tid; i < N * G; i += T) {
      const int ext = i / G, gl = i - ext * G, sl2 = t_rank[ext];
      smem[a.sm.busc + 2 * NG + sl2 * G + gl] = a.cbus[ext * 4 + 2];      // Gs
      smem[a.sm.busc + 3 * NG + sl2 * G + gl] = a.cbus[ext * 4 + 3];      // Bs
    }
    for (int i = tid; i < Gn * G; i += T) {
      const int j = i / G;
      smem[a.sm.genc + 0 * GnG + i] = a.cgen[j * 4 + 1];                  // Pmax
      smem[a.sm.genc + 1 * GnG + i] = a.cgen[j * 4 + 2];                  // Pmin
      smem[a.sm.genc + 4 * GnG + i] = a.cgen[j * 4 + 3];                  // qg
    }
  }
  // ---- TMA staging: the raw rows of batch b+gridDim are bulk-copied while batch b computes ----
  float* const s_raw_b = smem + a.sm.stage_b;
  float* const s_raw_l = smem + a.sm.stage_l;
  float* const s_raw_g = smem + a.sm.stage_g;
  uint64_t* const s_mbar = reinterpret_cast<uint64_t*>(smem + a.sm.mbar);
  const bool tma = a.use_tma != 0;
  uint32_t tma_phase = 0;
  // a batch takes the bulk path when it is full and its 16-byte windows stay inside the tensors
  // Input rows: the reference's packed rows (6 / 7 / 7 floats), or - inference only - the compact format that carries
  // just the columns that vary between the samples of a case (Pd,Qd | r,x,b,tau,shift | vg,Pg: 2 / 5 / 2 floats, see
  // gns_forward_compact); the constant columns then come once per CTA from the per-case blocks a.cbus / a.cgen.
  const bool compact = !GRAD && a.compact != 0;
  const int cb = compact ? 2 : 6, cl = compact ? 5 : 7, cg = compact ? 2 : 7;     // floats per raw row
  const int kb = compact ? 0 : 2, kl = compact ? 0 : 2;                             // first column kept
  auto bulk_ok = [&](int b) {
    const long long gb = (long long)b * G;
    if (!tma || gb + G > a.S) return false;
    const BulkWindow wb = bulk_window(gb * N * cb, G * N * cb), wl = bulk_window(gb * E * cl, G * E * cl),
                     wg = bulk_window(gb * Gn * cg, G * Gn * cg);
    return wb.begin + wb.bytes <= a.S * N * cb * 4 && wl.begin + wl.bytes <= a.S * E * cl * 4 &&
           wg.begin + wg.bytes <= a.S * Gn * cg * 4;
  };
  // compact generator rows (vg, Pg) -> vg, and Pg twice: Pg_set is a copy of Pg (ref GNS/utils.py:38)
  auto unpack_gens_compact = [&](const float* raw, long long g0, bool from_global) {
    for (int idx = tid; idx < Gn * G; idx += T) {
      const int gl = idx / Gn, j = idx - gl * Gn;
      const float* src = raw + (size_t)idx * 2;
      if (from_global) {
        long long g = g0 + gl;
        if (g >= a.S) g = a.S - 1;
        src = raw + ((size_t)g * Gn + j) * 2;
      }
      const float vg = src[0], pg = src[1];
      float* d = s_genc + j * G + gl;
      d[3 * GnG] = vg; d[2 * GnG] = pg; d[5 * GnG] = pg;
    }
  };
  auto bulk_issue = [&](int b) {     // one thread
    const long long gb = (long long)b * G;
    const BulkWindow wb = bulk_window(gb * N * cb, G * N * cb), wl = bulk_window(gb * E * cl, G * E * cl),
                     wg = bulk_window(gb * Gn * cg, G * Gn * cg);
    mbar_expect_tx(s_mbar, wb.bytes + wl.bytes + wg.bytes);
    bulk_g2s(s_raw_b, reinterpret_cast<const char*>(a.buses) + wb.begin, wb.bytes, s_mbar);
    bulk_g2s(s_raw_l, reinterpret_cast<const char*>(a.lines) + wl.begin, wl.bytes, s_mbar);
    if (wg.bytes) bulk_g2s(s_raw_g, reinterpret_cast<const char*>(a.gens) + wg.begin, wg.bytes, s_mbar);
  };
  // per-step weights: double-buffered bulk copies, step t+1 is in flight while step t computes
  uint64_t* const s_mbar_w = reinterpret_cast<uint64_t*>(smem + a.sm.mbar_w);
  const bool tma_w = tma && a.sm.weights2 != 0;
  uint32_t w_phase = 0, w_buf = 0;
  auto weights_issue = [&](int k, int buf) {   // one thread
    mbar_expect_tx(s_mbar_w, W.wstep * 4);
    bulk_g2s(smem + (buf ? a.sm.weights2 : a.sm.weights), a.params + (size_t)k * W.wstep, W.wstep * 4, s_mbar_w);
  };
  if (tma) {
    if (tid == 0) {
      mbar_init(s_mbar, 1);
      if (tma_w) mbar_init(s_mbar_w, 1);
      fence_proxy_async();
      if ((int)blockIdx.x < a.nbatch && bulk_ok(blockIdx.x)) bulk_issue(blockIdx.x);
      if (tma_w && (int)blockIdx.x < a.nbatch) weights_issue(0, 0);
    }
    __syncthreads();
  }

  // GRADV = 3: state checkpoint `kidx` of the thread's own bus (rows indexed by bus rank, one block per grid) and
  // zeros in the padding columns; the values were written by this thread, so no barrier is needed in front
  const int my_brank = (V3 && slot_on) ? (int)t_brank[sl] : 0;
  auto ckpt2_store = [&](long long g0, int kidx) {
    const int NbP = a.a2.NbP;
    if (bus_on) {
#pragma unroll 4
      for (int r = 0; r < 4 + L; ++r) {
        float x[VG];
        IO::ld(x, s_state + r * NG + slot * G + gcol);
#pragma unroll
        for (int g = 0; g < VG; ++g)
          __stcs(a.ck2 + ((size_t)(g0 + gcol + g) * (K + 1) + kidx) * (size_t)a.a2.state + r * NbP + my_brank, x[g]);
      }
    }
    const int npad = NbP - N;
    for (int i = tid; i < npad * (4 + L) * G; i += T) {
      const int gl = i / (npad * (4 + L)), rem = i - gl * (npad * (4 + L));
      const int r = rem / npad, c = N + rem - r * npad;
      __stcs(a.ck2 + ((size_t)(g0 + gl) * (K + 1) + kidx) * (size_t)a.a2.state + r * NbP + c, 0.f);
    }
  };

  for (int batch = blockIdx.x; batch < a.nbatch; batch += gridDim.x) {
    const long long g0 = (long long)batch * G;

    // ---------------- load + de-interleave the G grids of this batch ----------------
    if (bulk_ok(batch)) {
      mbar_wait(s_mbar, tma_phase);
      tma_phase ^= 1;
      unpack_block(s_raw_b + bulk_window(g0 * N * cb, 0).shift, s_busc, G, N, cb, kb, NG, t_rank);
      unpack_block(s_raw_l + bulk_window(g0 * E * cl, 0).shift, s_linef, G, E, cl, kl, EG, nullptr);
      if (compact) unpack_gens_compact(s_raw_g + bulk_window(g0 * Gn * cg, 0).shift, g0, false);
      else unpack_block(s_raw_g + bulk_window(g0 * Gn * 7, 0).shift, s_genc, G, Gn, 7, 1, GnG, nullptr);
    } else {
      load_block(a.buses, s_busc, g0, a.S, G, N, cb, kb, NG, t_rank);
      load_block(a.lines, s_linef, g0, a.S, G, E, cl, kl, EG, nullptr);
      if (compact) unpack_gens_compact(a.gens, g0, true);
      else load_block(a.gens, s_genc, g0, a.S, G, Gn, 7, 1, GnG, nullptr);
    }
    __syncthreads();
    {   // staging buffers are free again: start the copy of this CTA's next batch
      const int nb_next = batch + gridDim.x;
      if (tid == 0 && nb_next < a.nbatch && bulk_ok(nb_next)) {
        fence_proxy_async();
        bulk_issue(nb_next);
      }
    }

    // ---------------- state init (ref GNS/main.py:141-152) ----------------
    float part4[4][VG];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int g = 0; g < VG; ++g) part4[q][g] = 0.f;
    if (bus_on) {
      const int n = slot;
      float vv[VG], pg[VG], qg[VG];
#pragma unroll
      for (int g = 0; g < VG; ++g) { vv[g] = 0.f; pg[g] = 0.f; qg[g] = 0.f; }
      for (int j = j0; j < j1; ++j) {
        const int gid = t_geni[j];
        float x[VG];
        IO::ld(x, s_genc + 3 * GnG + gid * G + gcol);
#pragma unroll
        for (int g = 0; g < VG; ++g) vv[g] += x[g];
        IO::ld(x, s_genc + 5 * GnG + gid * G + gcol);
#pragma unroll
        for (int g = 0; g < VG; ++g) pg[g] += x[g];
        IO::ld(x, s_genc + 4 * GnG + gid * G + gcol);
#pragma unroll
        for (int g = 0; g < VG; ++g) qg[g] += x[g];
      }
      float Pd[VG], Qd[VG], Gs[VG], Bs[VG], o[VG];
      IO::ld(Pd, s_busc + 0 * NG + n * G + gcol);
      IO::ld(Qd, s_busc + 1 * NG + n * G + gcol);
      IO::ld(Gs, s_busc + 2 * NG + n * G + gcol);
      IO::ld(Bs, s_busc + 3 * NG + n * G + gcol);
      float* st = s_state + n * G + gcol;
#pragma unroll
      for (int g = 0; g < VG; ++g) vv[g] = (vv[g] == 0.f) ? 1.f : vv[g];
      IO::st(st + 0 * NG, vv);
#pragma unroll
      for (int g = 0; g < VG; ++g) o[g] = 0.f;
      IO::st(st + 1 * NG, o);
#pragma unroll
      for (int g = 0; g < VG; ++g) o[g] = pg[g] - Pd[g] - Gs[g] * (vv[g] * vv[g]);
      IO::st(st + 2 * NG, o);
#pragma unroll
      for (int g = 0; g < VG; ++g) o[g] = qg[g] - Qd[g] + Bs[g] * (vv[g] * vv[g]);
      IO::st(st + 3 * NG, o);
#pragma unroll
      for (int g = 0; g < VG; ++g) o[g] = 0.f;
      for (int i = 0; i < L; ++i) IO::st(st + (4 + i) * NG, o);
#pragma unroll
      for (int g = 0; g < VG; ++g) part4[0][g] = Pd[g];
    }
    if (slot < N) {   // alias-line admittance magnitude, lines 0..N-1 (ref GNS/main.py:38,87); slot = LINE id here
      float r[VG], x[VG], y[VG];
      IO::ld(r, s_linef + 0 * EG + slot * G + gcol);
      IO::ld(x, s_linef + 1 * EG + slot * G + gcol);
#pragma unroll
      for (int g = 0; g < VG; ++g) y[g] = 1.0f / sqrtf(r[g] * r[g] + x[g] * x[g]);
      IO::st(s_y + slot * G + gcol, y);
      // reciprocal tap ratio once per batch: the per-step flows multiply instead of dividing, which
      // removes the IEEE-division slow-path branches from the line phase (<= 1 ulp vs x / tau)
      IO::ld(x, s_linef + 3 * EG + slot * G + gcol);
#pragma unroll
      for (int g = 0; g < VG; ++g) y[g] = 1.0f / x[g];
      IO::st(s_y + NG + slot * G + gcol, y);
    }
    for (int it = tid; it < Gn * NGQ; it += T) {
      const int j = it / NGQ;
      float x[VG];
      IO::ld(x, s_genc + 2 * GnG + j * G + gcol);   // Pset
#pragma unroll
      for (int g = 0; g < VG; ++g) part4[1][g] += x[g];
      IO::ld(x, s_genc + 1 * GnG + j * G + gcol);   // Pmin
#pragma unroll
      for (int g = 0; g < VG; ++g) part4[2][g] += x[g];
      IO::ld(x, s_genc + 0 * GnG + j * G + gcol);   // Pmax
#pragma unroll
      for (int g = 0; g < VG; ++g) part4[3][g] += x[g];
    }
    float sPd[VG], sPset[VG], sPmin[VG], sPmax[VG];
    block_sum_multi<VG, 4>(part4, s_red, NGQ, red_parity);
#pragma unroll
    for (int g = 0; g < VG; ++g) { sPd[g] = part4[0][g]; sPset[g] = part4[1][g]; sPmin[g] = part4[2][g]; sPmax[g] = part4[3][g]; }
    if (tid < NGQ) {  // keep per-grid sums for the backward pass as well
      IO::st(s_gsum + 0 * G + gcol, sPd);
    }

    float loss_tot[VG], loss_last[VG];
#pragma unroll
    for (int g = 0; g < VG; ++g) { loss_tot[g] = 0.f; loss_last[g] = 0.f; }

    for (int k = 0; k < K; ++k) {
      // ---------------- stage this step's weights ----------------
      if (tma_w) {
        mbar_wait(s_mbar_w, w_phase);            // this step's weights have landed
        w_phase ^= 1;
        s_w = smem + (w_buf ? a.sm.weights2 : a.sm.weights);
        w_buf ^= 1;
      } else {
        const float4* src = reinterpret_cast<const float4*>(a.params + (size_t)k * W.wstep);
        float4* dst = reinterpret_cast<float4*>(smem + a.sm.weights);
        for (int i = tid; i < W.wstep / 4; i += T) dst[i] = __ldg(src + i);
      }
      // ---------------- checkpoint: state entering step k (k >= 1) ----------------
      if constexpr (V3) {
        ckpt2_store(g0, k);
        // zeros in the padding columns of this step's activation blocks
        const int pb = a.a2.NbP - N, pl = a.a2.EP - E;
        const int nb = 9 * H * pb, nl = (MULTI ? 3 : 1) * H * pl;
        for (int i = tid; i < (nb + nl) * G; i += T) {
          const int gl = i / (nb + nl);
          int rem = i - gl * (nb + nl);
          float* base = a.act + ((size_t)(g0 + gl) * K + k) * (size_t)a.a2.step;
          if (rem < nb) {
            const int blk = rem / (H * pb); rem -= blk * (H * pb);       // blk = 3 q + {h2L, h1L, A}
            const int r = rem / pb, c = N + rem - r * pb;
            const int q = blk / 3, w = blk - 3 * q;
            if (w == 0) __stcs(base + a.a2.h2L[q] + c * H + r, 0.f);       // item-major block
            else __stcs(base + (w == 1 ? a.a2.h1L[q] : a.a2.A[q]) + r * a.a2.NbP + c, 0.f);
          } else {
            rem -= nb;
            const int q = rem / (H * pl); rem -= q * (H * pl);
            const int r = rem / pl, c = E + rem - r * pl;
            __stcs(base + a.a2.h1line[q] + r * a.a2.EP + c, 0.f);
          }
        }
      } else if (GRAD && k >= 1) {
        const int nst4 = (4 + L) * NG / 4;
        float4* dstg = reinterpret_cast<float4*>(a.ckpt + ((size_t)batch * K + (k - 1)) * (size_t)((4 + L) * NG));
        const float4* srcs = reinterpret_cast<const float4*>(s_state);
        for (int i = tid; i < nst4; i += T) ckpt_store(dstg + i, srcs[i]);
      }
      __syncthreads();
      // Every thread is past the weight wait, and the other buffer was last read before the barrier
      // that ended the previous step: refill it with the next step's weights.  (Issuing before this
      // barrier could complete two phases ahead of a slow waiter, which would then spin forever.)
      if (tma_w && tid == 0 && ((k + 1 < K) || (batch + (int)gridDim.x < a.nbatch))) {
        fence_proxy_async();
        weights_issue(k + 1 < K ? k + 1 : 0, w_buf);
      }

      // ---------------- bus phase: phi nets, aggregation, L nets ----------------
      {
        const int n = pslot;                         // the bus's state lives in its primary slot
        // training: hidden activations of this step for the backward kernel (see ActLayout)
        float* const act_k = !GRAD ? nullptr
                             : (V3 ? a.act + ((size_t)(g0 + gcol) * K + k) * (size_t)a.a2.step
                                   : a.act + ((size_t)batch * K + k) * (size_t)a.al.total);
        const int gstr2 = V3 ? K * a.a2.step : 0;          // floats between the blocks of consecutive grids (GRADV = 3)
        float* st = s_state + n * G + gcol;
        const float* sm_m = st + 4 * NG;
        const float degf = (float)(e_full1 - e_in0); // in-degree of the bus (primary)
        // the bus's latent is read by six matrix-vector products per step (3 phi + 3 L nets): keep it
        // in registers when it is small enough (L*VG <= 40), else re-read it from shared memory
        constexpr bool MREG = (L * VG <= 40) && (GNS_GRAD_MREG || !GRAD);
        float mreg[MREG ? L : 1][VG];
        if constexpr (MREG) {
#pragma unroll
          for (int i = 0; i < L; ++i) IO::ld(mreg[i], sm_m + i * NG);
        }
        float st4[4][VG];
#pragma unroll
        for (int q = 0; q < 4; ++q) IO::ld(st4[q], st + q * NG);
        float dv[VG], dth[VG];
        float A[H][VG];
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
          // The m net of the LAST step only updates the latent, which nothing reads afterwards (the outputs are v, theta
          // and the physics losses; ref GNS/main.py:176-202 evaluates it and drops the result, its parameters get no
          // gradient): dead work, skipped.  (GRADV = 3 keeps it: the two experimental backward kernels walk every pair.)
          if (!V3 && q == 2 && k == K - 1) break;
          if (MULTI || q == 0) {
#pragma unroll
            for (int o = 0; o < H; ++o)
#pragma unroll
              for (int g = 0; g < VG; ++g) A[o][g] = 0.f;
           if (slot_on) {
            const float* wphi = s_w + (MULTI ? q * W.phi_size : 0);
            float P[H][VG];
            {
              float b[HP];
              load_row<HP>(b, wphi + W.phi_b1);
#pragma unroll
              for (int o = 0; o < H; ++o)
#pragma unroll
                for (int g = 0; g < VG; ++g) P[o][g] = b[o];
            }
            if constexpr (MREG) {
#pragma unroll
              for (int i = 0; i < L; ++i) row_axpy<H, HP, VG>(P, mreg[i], wphi + W.phi_w1m + i * HP);
            } else {
#pragma unroll 4
              for (int i = 0; i < L; ++i) {
                float x[VG];
                IO::ld(x, sm_m + i * NG);
                row_axpy<H, HP, VG>(P, x, wphi + W.phi_w1m + i * HP);
              }
            }
            for (int e = e_in0; e < e_in1; ++e) {
              const float* lf = s_linef + (int)t_ini[e] * G + gcol;
              wphi += opaque_zero();   // keep the weight rows in shared memory (no LICM into spills)
              float z[H][VG];
#pragma unroll
              for (int o = 0; o < H; ++o)
#pragma unroll
                for (int g = 0; g < VG; ++g) z[o][g] = P[o][g];
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                float x[VG];
                IO::ld(x, lf + c * EG);
                row_axpy<H, HP, VG>(z, x, wphi + W.phi_w1f + c * HP);
              }
              float z2[H][VG];
              {
                float b[HP];
                load_row<HP>(b, wphi + W.phi_b2);
#pragma unroll
                for (int o = 0; o < H; ++o) {
#pragma unroll
                  for (int g = 0; g < VG; ++g) z2[o][g] = b[o];
                  lrelu_vec<VG>(z[o]);
                }
              }
#pragma unroll
              for (int j = 0; j < H; ++j) row_axpy<H, HP, VG>(z2, z[j], wphi + W.phi_w2 + j * HP);
#pragma unroll
              for (int o = 0; o < H; ++o) { lrelu_vec<VG>(z2[o]); add_vec<VG>(A[o], z2[o]); }
              if constexpr (V3) {
                // h1 rows of the phi net (column in_pos) and the slope bits of (h1, h2) of this slot's (e - e_in0)-th line
                const int ep = a.a2.EP + opaque_zero();
                stg_rows<H, VG, false>(act_k + a.a2.h1line[MULTI ? q : 0] + (int)t_inp[e], ep, gstr2, z);
                uint32_t wb[VG];
#pragma unroll
                for (int g = 0; g < VG; ++g) wb[g] = 0u;
                slope_bits<H, VG>(wb, z, 0);
                slope_bits<H, VG>(wb, z2, H);
                uint32_t* mp = reinterpret_cast<uint32_t*>(act_k + a.a2.mask[MULTI ? q : 0]) + (1 + e - e_in0) * a.a2.NsM + sl;
#pragma unroll
                for (int g = 0; g < VG; ++g) __stcs(mp + (size_t)g * gstr2, wb[g]);
              } else if constexpr (GRAD) {
                // (the opaque zero keeps the 2H row offsets from being hoisted out of the line loop into spills)
                const int rl = a.al.rl + opaque_zero();
                float* ap = act_k + a.al.line_off + (size_t)((MULTI ? q : 0) * (H + 1)) * rl;
                const int el = gcol * a.al.gl + (int)t_inp[e] * a.al.ls;
                if constexpr (INTER) stg_rows<H, VG, true>(ap + el, rl, a.al.gl, z);
                else stg_rows4<H, VG>(ap, rl, el, a.al.gl, z);
                uint32_t wb[VG];
#pragma unroll
                for (int g = 0; g < VG; ++g) wb[g] = 0u;
                slope_bits<H, VG>(wb, z2, 0);
                uint32_t* mp = reinterpret_cast<uint32_t*>(ap + H * rl) + el;
#pragma unroll
                for (int g = 0; g < VG; ++g) ckpt_store(mp + g * a.al.gl, wb[g]);
              }
            }
           }
            if (warp_has_twins) {   // twins: combine the partial aggregates of a bus (lanes NGQ apart)
#pragma unroll
              for (int d = 1; d < 4; d *= 2) {
#pragma unroll
                for (int o = 0; o < H; ++o)
#pragma unroll
                  for (int g = 0; g < VG; ++g) {
                    const float t = __shfl_xor_sync(0xffffffffu, A[o][g], d * NGQ);
                    if (gsz > d) A[o][g] += t;
                  }
              }
            }
          }
          if (!prim) continue;
          float* const ab = (GRAD && !V3) ? act_k + (size_t)(q * 3 * H) * a.al.rb : nullptr;   // arrays A, h1L, h2L of the pair
          const int eb = (GRAD && !V3) ? gcol * a.al.gs + n * a.al.is : 0;                      // this thread's first element
          if constexpr (V3) {
            stg_rows<H, VG, false>(act_k + a.a2.A[q] + my_brank, a.a2.NbP, gstr2, A);
          } else if constexpr (GRAD) {
            if constexpr (INTER) stg_rows<H, VG, true>(ab + eb, a.al.rb, a.al.gs, A);
            else stg_rows4<H, VG>(ab, a.al.rb, eb, a.al.gs, A);
          }
          const float* wln = s_w + W.off_ln[0] + q * W.ln_size_s;   // L_v, L_theta, L_m are consecutive
          float zL[H][VG];
          {
            float b[HP];
            load_row<HP>(b, wln + W.ln_b1);
#pragma unroll
            for (int o = 0; o < H; ++o)
#pragma unroll
              for (int g = 0; g < VG; ++g) zL[o][g] = b[o];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) row_axpy<H, HP, VG>(zL, st4[i], wln + W.ln_w1 + i * HP);
          if constexpr (MREG) {
#pragma unroll
            for (int i = 0; i < L; ++i) row_axpy<H, HP, VG>(zL, mreg[i], wln + W.ln_w1 + (4 + i) * HP);
          } else {
#pragma unroll 4
            for (int i = 0; i < L; ++i) {
              float x[VG];
              IO::ld(x, sm_m + i * NG);
              row_axpy<H, HP, VG>(zL, x, wln + W.ln_w1 + (4 + i) * HP);
            }
          }
          {   // S = W4 A + deg b4 feeds the first layer linearly: use the pre-multiplied block
              // M = W4^T W1[:, 4+L:]^T and c = b4 W1[:, 4+L:]^T (fuse_params_kernel): H*H instead of 2*L*H MACs
            const float* wmf = s_w + W.off_mf[0] + q * W.mf_size;
#pragma unroll
            for (int j = 0; j < H; ++j) row_axpy<H, HP, VG>(zL, A[j], wmf + j * HP);
            float dg[VG];
#pragma unroll
            for (int g = 0; g < VG; ++g) dg[g] = degf;
            row_axpy<H, HP, VG>(zL, dg, wmf + H * HP);
          }
          float z2[H][VG];
          {
            float b[HP];
            load_row<HP>(b, wln + W.ln_b2);
#pragma unroll
            for (int o = 0; o < H; ++o) {
#pragma unroll
              for (int g = 0; g < VG; ++g) z2[o][g] = b[o];
              lrelu_vec<VG>(zL[o]);
            }
          }
#pragma unroll
          for (int j = 0; j < H; ++j) row_axpy<H, HP, VG>(z2, zL[j], wln + W.ln_w2 + j * HP);
#pragma unroll
          for (int o = 0; o < H; ++o) lrelu_vec<VG>(z2[o]);
          if constexpr (V3) {
            stg_rows<H, VG, false>(act_k + a.a2.h1L[q] + my_brank, a.a2.NbP, gstr2, zL);
            {   // h2 of the L-net, item-major: H consecutive floats per bus (hidden-side operand of dWout)
              static_assert(H % 2 == 0, "64-bit stores");
              float* hp = act_k + a.a2.h2L[q] + my_brank * H;
#pragma unroll
              for (int g = 0; g < VG; ++g)
#pragma unroll
                for (int o = 0; o < H; o += 2)
                  __stcs(reinterpret_cast<float2*>(hp + (size_t)g * gstr2 + o), make_float2(z2[o][g], z2[o + 1][g]));
            }
            uint32_t wb[VG];
#pragma unroll
            for (int g = 0; g < VG; ++g) wb[g] = 0u;
            slope_bits<H, VG>(wb, zL, 0);
            slope_bits<H, VG>(wb, z2, H);
            uint32_t* mp = reinterpret_cast<uint32_t*>(act_k + a.a2.mask[q]) + sl;
#pragma unroll
            for (int g = 0; g < VG; ++g) __stcs(mp + (size_t)g * gstr2, wb[g]);
          } else if constexpr (GRAD) {
            if constexpr (INTER) {
              stg_rows<H, VG, true>(ab + H * a.al.rb + eb, a.al.rb, a.al.gs, zL);
              stg_rows<H, VG, true>(ab + 2 * H * a.al.rb + eb, a.al.rb, a.al.gs, z2);
            } else {
              stg_rows4<H, VG>(ab + H * a.al.rb, a.al.rb, eb, a.al.gs, zL);
              stg_rows4<H, VG>(ab + 2 * H * a.al.rb, a.al.rb, eb, a.al.gs, z2);
            }
          }
          if (q < 2) {
            float out[VG];
            const float bo = wln[W.ln_bo_s];
#pragma unroll
            for (int g = 0; g < VG; ++g) out[g] = bo;
            row_dot<H, HP, VG>(out, z2, wln + W.ln_wo);
            if (q == 0) {
#pragma unroll
              for (int g = 0; g < VG; ++g) dv[g] = out[g];
            } else {
#pragma unroll
              for (int g = 0; g < VG; ++g) dth[g] = out[g];
            }
          } else {
#pragma unroll 2
            for (int i = 0; i < L; ++i) {
              float dm[VG], mi[VG];
              const float bo = wln[W.ln_bo_m + i];
#pragma unroll
              for (int g = 0; g < VG; ++g) dm[g] = bo;
              row_dot<H, HP, VG>(dm, z2, wln + W.ln_wo + i * HP);
              IO::ld(mi, st + (4 + i) * NG);
              add_vec<VG>(mi, dm);
              IO::st(st + (4 + i) * NG, mi);
            }
          }
        }
        // state update (ref GNS/main.py:182-188); v only moves on non-generator buses
        if (prim) {
          const bool is_gen = j1 > j0;
#pragma unroll
          for (int g = 0; g < VG; ++g) {
            st4[1][g] = st4[1][g] + dth[g];
            if (!is_gen) st4[0][g] = st4[0][g] + dv[g];
          }
          IO::st(st + 0 * NG, st4[0]);
          IO::st(st + 1 * NG, st4[1]);
        }
      }
      __syncthreads();

      // ---------------- physics 1: angle difference of the alias lines 0..N-1 ----------------
      if (slot < N) {
        const int j = slot;  // used as a LINE id here
        float tf[VG], tt[VG], d[VG], sd[VG], cd[VG];
        IO::ld(tf, s_state + 1 * NG + (int)t_fi[j] * G + gcol);
        IO::ld(tt, s_state + 1 * NG + (int)t_ti[j] * G + gcol);
        if constexpr (VG == 2) {
          const Pk2 dd = pk2(tf[0], tf[1]) - pk2(tt[0], tt[1]);
          Pk2 s2v, c2v;
          fast_sincos(dd, s2v, c2v);
          d[0] = dd.v.x; d[1] = dd.v.y; sd[0] = s2v.v.x; sd[1] = s2v.v.y; cd[0] = c2v.v.x; cd[1] = c2v.v.y;
        } else {
#pragma unroll
          for (int g = 0; g < VG; ++g) { d[g] = tf[g] - tt[g]; fast_sincos(d[g], sd[g], cd[g]); }
        }
        IO::st(s_trig + 0 * NG + j * G + gcol, d);
        IO::st(s_trig + 1 * NG + j * G + gcol, sd);
        IO::st(s_trig + 2 * NG + j * G + gcol, cd);
      }
      __syncthreads();

      // ---------------- physics 2: per-line flows (ref GNS/main.py:38-41,66-72,87-99) ----------------
      float pj[VG];
#pragma unroll
      for (int g = 0; g < VG; ++g) pj[g] = 0.f;
      // two line items of a thread are interleaved (E is just above the thread count on case300: the
      // second, nearly empty round would otherwise cost a full sincos latency chain)
#pragma unroll 2
      for (int it = tid; it < E * NGQ; it += T) {
        const int e = it / NGQ;
        const int fi = t_fi[e], ti = t_ti[e], fa = t_fa[e], ta = t_ta[e];
        float vf[VG], vt[VG], thf[VG], tht[VG];
        IO::ld(vf, s_state + 0 * NG + fi * G + gcol);
        IO::ld(vt, s_state + 0 * NG + ti * G + gcol);
        IO::ld(thf, s_state + 1 * NG + fi * G + gcol);
        IO::ld(tht, s_state + 1 * NG + ti * G + gcol);
        float Yf[VG], tauf[VG], shf[VG], bf[VG], Df[VG], sDf[VG], cDf[VG];
        IO::ld(Yf, s_y + fa * G + gcol);
        IO::ld(bf, s_linef + 2 * EG + fa * G + gcol);
        IO::ld(tauf, s_y + NG + fa * G + gcol);      // 1 / tau
        IO::ld(shf, s_linef + 4 * EG + fa * G + gcol);
        IO::ld(Df, s_trig + 0 * NG + fa * G + gcol);
        IO::ld(sDf, s_trig + 1 * NG + fa * G + gcol);
        IO::ld(cDf, s_trig + 2 * NG + fa * G + gcol);
        float Yt[VG], taut[VG], sht[VG], bt[VG], Dt[VG], sDt[VG];
        IO::ld(Yt, s_y + ta * G + gcol);
        IO::ld(bt, s_linef + 2 * EG + ta * G + gcol);
        IO::ld(taut, s_y + NG + ta * G + gcol);      // 1 / tau
        IO::ld(sht, s_linef + 4 * EG + ta * G + gcol);
        IO::ld(Dt, s_trig + 0 * NG + ta * G + gcol);
        IO::ld(sDt, s_trig + 1 * NG + ta * G + gcol);
        float pf[VG], qf[VG], pt[VG], qt[VG];
        if constexpr (VG == 2) {   // both grids in packed f32x2 arithmetic
          auto P = [](const float (&x)[VG]) { return pk2(x[0], x[1]); };
          const LineTerms<Pk2> o = line_terms<Pk2>(P(vf), P(vt), P(thf), P(tht), P(Yf), P(bf), P(tauf), P(shf), P(Df), P(sDf),
                                                   P(cDf), P(Yt), P(bt), P(taut), P(sht), P(Dt), P(sDt));
          pj[0] += o.msg.v.x; pj[1] += o.msg.v.y;
          pf[0] = o.pf.v.x; pf[1] = o.pf.v.y; qf[0] = o.qf.v.x; qf[1] = o.qf.v.y;
          pt[0] = o.pt.v.x; pt[1] = o.pt.v.y; qt[0] = o.qt.v.x; qt[1] = o.qt.v.y;
        } else {
#pragma unroll
          for (int g = 0; g < VG; ++g) {
            const LineTerms<float> o = line_terms<float>(vf[g], vt[g], thf[g], tht[g], Yf[g], bf[g], tauf[g], shf[g], Df[g],
                                                         sDf[g], cDf[g], Yt[g], bt[g], taut[g], sht[g], Dt[g], sDt[g]);
            pj[g] += o.msg;
            pf[g] = o.pf; qf[g] = o.qf; pt[g] = o.pt; qt[g] = o.qt;
          }
        }
        IO::st(s_flow + 0 * EG + e * G + gcol, pf);
        IO::st(s_flow + 1 * EG + e * G + gcol, qf);
        IO::st(s_flow + 2 * EG + e * G + gcol, pt);
        IO::st(s_flow + 3 * EG + e * G + gcol, qt);
      }
      float v_own[VG], Gs[VG], Bs[VG];
      if (bus_on) {
        IO::ld(v_own, s_state + 0 * NG + slot * G + gcol);
        IO::ld(Gs, s_busc + 2 * NG + slot * G + gcol);
        IO::ld(Bs, s_busc + 3 * NG + slot * G + gcol);
#pragma unroll
        for (int g = 0; g < VG; ++g) pj[g] += (v_own[g] * v_own[g]) * Gs[g];
      }
      block_sum_per_grid<VG>(pj, s_red, NGQ, red_parity);   // its barrier also orders the s_flow writes before the gathers

      // ---------------- physics 3: slack redistribution + per-bus mismatch ----------------
      float lam[VG];
      bool lo_arm[VG];
#pragma unroll
      for (int g = 0; g < VG; ++g) {
        const float pglob = sPd[g] + pj[g];
        pj[g] = pglob;
        const float l1 = (pglob - sPmin[g]) / (2.f * (sPset[g] - sPmin[g]));
        const float l2 = (pglob - 2.f * sPset[g] + sPmax[g]) / (2.f * (sPmax[g] - sPset[g]));
        lam[g] = (pglob < sPset[g]) ? l1 : l2;
        lo_arm[g] = lam[g] < 0.5f;
      }
      if constexpr (V3) {
        if (tid < NGQ) {
#pragma unroll
          for (int g = 0; g < VG; ++g) a.pglob[(size_t)(g0 + gcol + g) * K + k] = pj[g];
        }
      } else if (GRAD && tid < NGQ) IO::st(a.pglob + ((size_t)batch * K + k) * G + gcol, pj);
      if (bus_on) {
        const int n = slot;
        float pgs[VG];
#pragma unroll
        for (int g = 0; g < VG; ++g) pgs[g] = 0.f;
        for (int j = j0; j < j1; ++j) {
          const int gid = t_geni[j];
          float Pmax[VG], Pmin[VG], Pset[VG];
          IO::ld(Pmax, s_genc + 0 * GnG + gid * G + gcol);
          IO::ld(Pmin, s_genc + 1 * GnG + gid * G + gcol);
          IO::ld(Pset, s_genc + 2 * GnG + gid * G + gcol);
#pragma unroll
          for (int g = 0; g < VG; ++g)
            pgs[g] += lo_arm[g] ? (Pmin[g] + 2.f * (Pset[g] - Pmin[g]) * lam[g])
                                : (2.f * Pset[g] - Pmax[g] + 2.f * (Pmax[g] - Pset[g]) * lam[g]);
        }
        float spf[VG], sqf[VG], spt[VG], sqt[VG];
#pragma unroll
        for (int g = 0; g < VG; ++g) { spf[g] = 0.f; sqf[g] = 0.f; spt[g] = 0.f; sqt[g] = 0.f; }
        for (int e = e_in0; e < e_full1; ++e) {
          const int line = t_ini[e];
          float x[VG];
          IO::ld(x, s_flow + 0 * EG + line * G + gcol);
#pragma unroll
          for (int g = 0; g < VG; ++g) spf[g] += x[g];
          IO::ld(x, s_flow + 1 * EG + line * G + gcol);
#pragma unroll
          for (int g = 0; g < VG; ++g) sqf[g] += x[g];
        }
        for (int e = e_out0; e < e_out1; ++e) {
          const int line = t_outi[e];
          float x[VG];
          IO::ld(x, s_flow + 2 * EG + line * G + gcol);
#pragma unroll
          for (int g = 0; g < VG; ++g) spt[g] += x[g];
          IO::ld(x, s_flow + 3 * EG + line * G + gcol);
#pragma unroll
          for (int g = 0; g < VG; ++g) sqt[g] += x[g];
        }
        float Pd[VG], Qd[VG], dP[VG], dQ[VG];
        IO::ld(Pd, s_busc + 0 * NG + n * G + gcol);
        IO::ld(Qd, s_busc + 1 * NG + n * G + gcol);
        const float wk = a.wk[k];
#pragma unroll
        for (int g = 0; g < VG; ++g) {
          const float v2 = v_own[g] * v_own[g];
          const float qg = (Qd[g] - Bs[g] * v2) - sqf[g] - sqt[g];
          dP[g] = pgs[g] - Pd[g] - Gs[g] * v2 + spf[g] + spt[g];
          dQ[g] = qg - Qd[g] + Bs[g] * v2 + sqf[g] + sqt[g];
          const float sq = dP[g] * dP[g] + dQ[g] * dQ[g];
          loss_tot[g] = fmaf(wk, sq, loss_tot[g]);
          loss_last[g] = sq;
        }
        IO::st(s_state + 2 * NG + n * G + gcol, dP);
        IO::st(s_state + 3 * NG + n * G + gcol, dQ);
      }
      __syncthreads();
    }  // k

    // ---------------- outputs ----------------
    if constexpr (V3) {
      ckpt2_store(g0, K);
    } else if (GRAD) {   // final state: v, theta, dP, dQ are what the backward pass reads; the latent rows are not stored
      const int nst4 = 4 * NG / 4;
      float4* dstg = reinterpret_cast<float4*>(a.ckpt + ((size_t)batch * K + (K - 1)) * (size_t)((4 + L) * NG));
      const float4* srcs = reinterpret_cast<const float4*>(s_state);
      for (int i = tid; i < nst4; i += T) ckpt_store(dstg + i, srcs[i]);
    }
    {
      float l2[2][VG];
#pragma unroll
      for (int g = 0; g < VG; ++g) { l2[0][g] = loss_tot[g]; l2[1][g] = loss_last[g]; }
      block_sum_multi<VG, 2>(l2, s_red, NGQ, red_parity);
#pragma unroll
      for (int g = 0; g < VG; ++g) { loss_tot[g] = l2[0][g]; loss_last[g] = l2[1][g]; }
    }
    if (tid < NGQ) {
#pragma unroll
      for (int g = 0; g < VG; ++g) {
        const long long gg = g0 + gcol + g;
        if (gg < a.S) {
          a.total[gg] = loss_tot[g] / (float)N;
          a.last[gg] = loss_last[g] / (float)N;
        }
      }
    }
    if (bus_on) {
      const int ext = t_ext[slot];
      float vv[VG], th[VG];
      IO::ld(vv, s_state + 0 * NG + slot * G + gcol);
      IO::ld(th, s_state + 1 * NG + slot * G + gcol);
#pragma unroll
      for (int g = 0; g < VG; ++g) {
        const long long gg = g0 + gcol + g;
        if (gg < a.S) {
          a.v[gg * N + ext] = (vv[g] < 0.f) ? 0.f : vv[g];   // ref GNS/main.py:201
          a.theta[gg * N + ext] = th[g];
        }
      }
    }
    __syncthreads();
  }  // batch
}

}  // namespace gns
