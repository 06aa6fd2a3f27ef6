// gns_plan.cu — topology plan: CSR construction, internal bus order, index block upload,
// launch geometry, workspace carving, canonical<->packed parameter maps.
//
// Replaces the index tensors the reference rebuilds inside every forward call
// (ref GNS/main.py:35-36, 85-86, 144, 153, 184-185) by a one-time plan.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "gns_host.h"
#include "../../include/gns_b200.h"

namespace gns {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error_cstr() { return g_err.c_str(); }

#define GNS_CUDA_OK(expr)                                                            \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
      return -2;                                                                     \
    }                                                                                \
  } while (0)

// stable counting sort of line ids by key
static void csr_by(const std::vector<int32_t>& key, int nb, std::vector<int32_t>& rowptr,
                   std::vector<int32_t>& ids) {
  rowptr.assign(nb + 1, 0);
  for (int32_t k : key) rowptr[k + 1]++;
  for (int i = 0; i < nb; ++i) rowptr[i + 1] += rowptr[i];
  ids.assign(key.size(), 0);
  std::vector<int32_t> cur(rowptr.begin(), rowptr.end() - 1);
  for (int32_t e = 0; e < (int32_t)key.size(); ++e) ids[cur[key[e]]++] = e;
}

static int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return (s && *s) ? std::atoi(s) : dflt;
}

static SmemPlan plan_smem(int N /* bus slots */, int E, int Gn, int G, int L, int wstep, int nwarps, int topo_u16,
                          int extra_floats, bool backward, int stage_rows_n /* real buses, 0 = no staging */,
                          int state_rows, bool double_weights) {
  SmemPlan s{};
  int o = 0;
  auto take = [&](int n) { int r = o; o += pad4(n); return r; };
  const int NGs = backward ? bwd_bus_stride(N * G) : row_stride(N * G), EGs = row_stride(E * G);
  (void)L;
  s.state = take(state_rows * NGs);
  s.busc = take(4 * NGs);
  s.genc = take(6 * Gn * G);
  s.linef = take(5 * EGs);
  s.yline = take(2 * NGs);       // Y and 1/tau of the alias lines
  s.trig = take(3 * NGs);
  s.flows = take(backward ? 0 : 4 * EGs);   // the backward kernel keeps its own per-line block
  s.gsum = take(4 * G);
  s.red = take(2 * nwarps * kRedNV * G);   // two alternating halves, kRedNV values per warp and grid
  s.weights = take(wstep);
  s.topo = take((topo_u16 + 1) / 2);
  if (!backward && stage_rows_n > 0) {   // raw staging of the next batch's rows (+ slack for the 16-byte window)
    s.stage_b = take(G * stage_rows_n * 6 + 8);
    s.stage_l = take(G * E * 7 + 8);
    s.stage_g = take(G * Gn * 7 + 8);
    s.mbar = take(4);
    if (double_weights) {
      s.weights2 = take(wstep);
      s.mbar_w = take(4);
    }
  }
  s.extra = o;
  o += pad4(extra_floats);
  s.total_floats = o;
  return s;
}

// extra shared memory of the backward kernel, in floats (see gns_backward.cuh)
int backward_extra_floats(int N, int E, int G, int L, int H, int T);
// shared memory of the warp-specialised backward kernel, in floats (see gns_backward2.cuh)
int make_bwd2_smem_floats(int L, int H, int E, int wstep, const Act2Layout& a2);
int make_bwd3_smem_floats(int L, int H, int N, int E, int wstep, int nwarps, const Act2Layout& a2);
int frag3_step_floats(int L, int H);

// Warp w runs on SM sub-partition w % 4.  A warp's bus phase costs ~ (1 + c * max lines walked by
// a lane of its bus group), c ~ 0.2 forward / 0.3 backward (instruction counts).  Groups are placed
// longest-first on the sub-partition with the least load that still has a free warp, then pairs of
// groups are swapped between sub-partitions while that lowers the maximum load (with 10 warps the
// sub-partitions hold 3,3,2,2 warps, so the heaviest groups must end up on the 2-warp ones).
static void balance_warps(const gns_plan* plan, Geometry* g, bool backward) {
  const int nw = g->T / 32, spw = 32 / g->NGQ;
  float per_line = backward ? 0.3f : 0.2f;
  if (const char* e = std::getenv(backward ? "GNS_BWD_LINE_COST" : "GNS_FWD_LINE_COST")) per_line = (float)std::atof(e);
  std::vector<float> cost(nw, 0.f);
  for (int grp = 0; grp < nw; ++grp) {
    int mx = -1;
    for (int s = grp * spw; s < std::min((grp + 1) * spw, plan->Ns); ++s)
      mx = std::max(mx, plan->slot_in_end[s] - plan->slot_in_begin[s]);
    cost[grp] = mx < 0 ? 0.f : 1.f + per_line * mx;
  }
  std::vector<int> order(nw);
  for (int i = 0; i < nw; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  float load[4] = {0, 0, 0, 0};
  std::vector<int> bin[4];
  for (int grp : order) {
    int best = -1;
    for (int sp = 0; sp < 4; ++sp) {
      const int cap = (nw - sp + 3) / 4;           // warps sp, sp+4, ...
      if ((int)bin[sp].size() >= cap) continue;
      if (best < 0 || load[sp] < load[best]) best = sp;
    }
    bin[best].push_back(grp);
    load[best] += cost[grp];
  }
  for (bool improved = true; improved;) {            // pairwise swaps that lower the heavier of the two loads
    improved = false;
    for (int a = 0; a < 4; ++a)
      for (int b = a + 1; b < 4; ++b)
        for (size_t i = 0; i < bin[a].size(); ++i)
          for (size_t j = 0; j < bin[b].size(); ++j) {
            const float d = cost[bin[a][i]] - cost[bin[b][j]];
            const float na = load[a] - d, nb = load[b] + d;
            if (std::max(na, nb) + 1e-6f < std::max(load[a], load[b])) {
              std::swap(bin[a][i], bin[b][j]);
              load[a] = na; load[b] = nb;
              improved = true;
            }
          }
  }
  for (int sp = 0; sp < 4; ++sp)
    for (size_t i = 0; i < bin[sp].size(); ++i) g->grp_of_warp[sp + 4 * i] = (unsigned char)bin[sp][i];
  if (std::getenv("GNS_NO_BALANCE")) for (int w = 0; w < nw; ++w) g->grp_of_warp[w] = (unsigned char)w;
}

bool choose_geometry(const gns_plan* plan, const ModelDims& md, long long S, bool backward, Geometry* out) {
  const int N = plan->Ns, E = plan->E, Gn = plan->Gn;   // N = bus slots here
  const WLayout W = make_wlayout(md.L, md.H, md.multi != 0);
  const int limit = plan->smem_optin;
  Geometry best{};
  bool found = false;
  const int force_vg = env_int(backward ? "GNS_BWD_VG" : "GNS_FWD_VG", 0);
  const int force_ngq = env_int(backward ? "GNS_BWD_NGQ" : "GNS_FWD_NGQ", 0);
  const int target_threads = env_int("GNS_TARGET_THREADS", 320);
  const bool allow_tma = !backward && !env_int("GNS_NO_TMA", 0);
  for (int VG : {2, 1}) {
    if (force_vg && VG != force_vg) continue;
    if (backward && VG != 1) continue;          // the backward kernel handles one grid per thread
    // small batches: one grid per thread so that the batch spreads over more SMs
    if (!force_vg && !backward && VG == 2 && (S + 1) / 2 < plan->num_sms) continue;
    for (int NGQ = 32; NGQ >= 1; NGQ >>= 1) {
      if (force_ngq && NGQ != force_ngq) continue;
      const int T = ((N * NGQ + 31) / 32) * 32;
      if (T > 1024) continue;
      if (32 / NGQ < plan->max_gsz) continue;     // a twin group must stay inside one warp
      const int G = VG * NGQ;
      const long long nb = (S + G - 1) / G;
      if (!force_ngq && NGQ > 1 && (T > std::max(target_threads, 32) || nb < 2LL * plan->num_sms)) continue;
      Geometry g{};
      g.VG = VG; g.NGQ = NGQ; g.G = G; g.T = T;
      g.tmax = (T <= 320) ? 320 : ((T <= 384) ? 384 : 1024);   // launch-bounds variant (registers per thread)
      if (g.tmax == 1024 && VG != 1) continue;   // the wide-CTA variant exists for VG=1 only
      const int extra = backward ? backward_extra_floats(N, E, G, md.L, md.H, T) : 0;
      // TMA extras (input staging, second weight buffer) are taken only while they do not lower the
      // number of CTAs that fit on one SM (228 KB per SM, 1 KB reserved per CTA)
      const int state_rows = (backward && md.L > 32) ? 4 : 4 + md.L;
      const int reg_ctas = std::max(1, 65536 / (T * (g.tmax == 320 ? 200 : (g.tmax == 384 ? 168 : 64))));      // register-file limit
      auto ctas_per_sm = [&](size_t b) { return std::min(reg_ctas, (int)((size_t)233472 / (b + 1024))); };
      SmemPlan sm = plan_smem(N, E, Gn, G, md.L, W.wstep, T / 32, plan->to.total, extra, backward, 0, state_rows, false);
      size_t bytes = (size_t)sm.total_floats * 4;
      if ((int)bytes > limit) continue;
      if (allow_tma) {
        for (int level = 2; level >= 1; --level) {
          const SmemPlan t = plan_smem(N, E, Gn, G, md.L, W.wstep, T / 32, plan->to.total, extra, backward, plan->N,
                                       state_rows, level == 2);
          const size_t tb = (size_t)t.total_floats * 4;
          if ((int)tb <= limit && ctas_per_sm(tb) >= ctas_per_sm(bytes)) { sm = t; bytes = tb; break; }
        }
      }
      g.smem_bytes = bytes; g.sm = sm;
      balance_warps(plan, &g, backward);
      g.nbatch = (int)nb; g.num_sms = plan->num_sms;
      best = g; found = true;
      break;
    }
    if (found) break;
  }
  if (!found) {
    set_error("no launch geometry fits: n_bus=" + std::to_string(N) + " n_line=" + std::to_string(E) +
              " latent_dim=" + std::to_string(md.L) + " (shared memory limit " + std::to_string(limit) + " B)");
    return false;
  }
  *out = best;
  return true;
}

int bwd2_threads_limit() { return 384; }

Bwd2Geom choose_bwd2(const gns_plan* plan, const ModelDims& md, long long S) {
  Bwd2Geom g{};
  // Opt-in (GNS_BWD2=1).  Measured on B200, case300 K=4, 16,384 grids (profiles/r02_bwd2_*): 17.8 ms against 16.6 ms of
  // the first kernel - the producer side runs on 5 warps at the same ~0.13 instructions / cycle / warp as every warp of
  // the other kernels, so the instruction savings (330 k instead of 428 k warp instructions per grid) do not pay for the
  // lost warp-level parallelism.  Kept as the evidence of that experiment and as a second, independent implementation
  // of the adjoint (parity tests run both).
  const int force = env_int("GNS_BWD2", 0);
  if (S <= 0) return g;
  if (env_int("GNS_BWD3", 0) == 1 && md.multi && md.H == 10 && (md.L == 10 || md.L == 20) && plan->max_walk >= 1) {
    // fragment-space kernel (gns_backward3.cuh): one thread per bus for the physics, 16-item tiles per warp.
    // Two geometries.  Wide: one bus tile and at most two line tiles per warp (up to 20 warps of 96 registers, or two
    // CTAs of up to 8 warps and 128 registers per SM); narrow: two bus / three line tiles per warp (up to 10 warps of 168
    // registers).  Measured on B200, K=4, 16,384 grids, backward pass only: case300 wide 20.0 ms (18.6 ms with shared
    // accumulator blocks) / narrow 18.5 ms / first kernel 16.7 ms; case118 wide 10.2 ms / narrow 13.7 ms / first kernel
    // 8.2 ms.  With 19 warps the legacy mma.sync pipe throttles and the CTA barriers take 22 % of the warp time, so the
    // wide geometry is the default only where it brings a second CTA onto the SM.  GNS_BWD3_TILES=1 / 2 forces one.
    // GNS_DETERMINISTIC=0: the warps of a CTA share one accumulator block (their reductions race: gradients reproducible
    // to rounding only); otherwise one block per warp.
    const int nbt = (plan->N + 15) / 16, nlt = (plan->E + 15) / 16;
    const int tiles_env = env_int("GNS_BWD3_TILES", 0);
    const int nw_wide = std::max(std::max(nbt, (nlt + 1) / 2), std::max(2, (plan->N + 31) / 32));
    const bool wide = tiles_env == 1 || (tiles_env == 0 && nw_wide * 32 <= 256);
    const int bt = wide ? 1 : 2, lt = wide ? 2 : 3;
    const int nw = std::max(std::max((nbt + bt - 1) / bt, (nlt + lt - 1) / lt), std::max(2, (plan->N + 31) / 32));
    const int T = nw * 32;
    const char* det_env = std::getenv("GNS_DETERMINISTIC");
    const bool shared_acc = det_env && det_env[0] == '0';
    if (T <= (wide ? 640 : 320) && plan->N >= env_int("GNS_BWD2_MIN_SLOTS", 96)) {
      g.a2 = make_act2_layout(md.L, md.H, plan->N, plan->Ns, plan->E, plan->max_walk);
      const WLayout W = make_wlayout(md.L, md.H, true);
      const size_t bytes = (size_t)make_bwd3_smem_floats(md.L, md.H, plan->N, plan->E, W.wstep, nw, g.a2) * 4;
      if ((int)bytes <= plan->smem_optin) {
        g.variant = 3; g.PW = nw; g.CW = shared_acc ? 1 : nw; g.T = T; g.smem_bytes = bytes; g.bt = bt; g.lt = lt;
        g.parts_per_cta = shared_acc ? 1 : nw;
        const int regs = !wide ? 168 : (T <= 256 ? 128 : 96);
        const int per_sm = std::max(1, std::min((int)((size_t)233472 / (bytes + 1024)), 65536 / (T * regs)));
        g.ctas = (int)std::min<long long>(S, (long long)plan->num_sms * per_sm);
        g.ok = true;
        return g;
      }
    }
  }
  if (force != 1) return g;
  if (md.H != 10 || (md.L != 10 && md.L != 20)) return g;           // instantiated dims (latent in registers: L <= 20)
  if (plan->max_walk > 4 || plan->max_walk < 1) return g;            // kB2MaxWalk
  const int min_slots = env_int("GNS_BWD2_MIN_SLOTS", 96);
  if (force != 1 && plan->Ns < min_slots) return g;                   // small grids: several grids per CTA (first kernel)
  const int PW = ((plan->Ns + 1) / 2 + 31) / 32;
  const int CW = std::max(1, env_int("GNS_BWD2_CW", std::max(1, (3 * PW + 4) / 5)));   // measured best: 3 consumer warps for 5 producers
  if (PW > 8 || CW > 8 || (PW + CW) * 32 > bwd2_threads_limit()) return g;
  g.a2 = make_act2_layout(md.L, md.H, plan->N, plan->Ns, plan->E, plan->max_walk);
  if (md.L * g.a2.NbP > 2 * md.H * g.a2.EP) return g;                // the adj m' rows live in the two line blocks
  const WLayout W = make_wlayout(md.L, md.H, md.multi != 0);
  const size_t bytes = (size_t)make_bwd2_smem_floats(md.L, md.H, plan->E, W.wstep, g.a2) * 4;
  if ((int)bytes > plan->smem_optin) return g;
  g.PW = PW; g.CW = CW; g.T = (PW + CW) * 32; g.smem_bytes = bytes;
  const int per_sm = std::max(1, std::min((int)((size_t)233472 / (bytes + 1024)), 65536 / (g.T * 168)));
  g.ctas = (int)std::min<long long>(S, (long long)plan->num_sms * per_sm);
  g.ok = true;
  return g;
}

bool act_grid_major(const Geometry& bwd) {
  if (const char* e = std::getenv("GNS_ACT_LAYOUT")) {
    if (e[0] == 'g') return true;
    if (e[0] == 'i') return false;
  }
  return bwd.G == 1;
}

int64_t canonical_param_count(const ModelDims& md) {
  const int64_t L = md.L, H = md.H;
  const int64_t phi_out = md.multi ? L : 1;
  const int64_t phi = (5 + L) * H + H + H * H + H + phi_out * H + phi_out;
  const int64_t ls = (4 + 2 * L) * H + H + H * H + H + H + 1;
  const int64_t lm = (4 + 2 * L) * H + H + H * H + H + L * H + L;
  return (int64_t)md.K * ((md.multi ? 3 : 1) * phi + 2 * ls + lm);
}

// canonical (state_dict order) index -> packed [K][wstep] index
std::vector<int32_t> build_pack_map(const ModelDims& md) {
  const int L = md.L, H = md.H, K = md.K;
  const bool multi = md.multi != 0;
  const WLayout W = make_wlayout(L, H, multi);
  std::vector<int32_t> map;
  map.reserve((size_t)canonical_param_count(md));
  // canonical net order: phi_v, phi_theta, phi_m | phi ; L_theta, L_v, L_m
  struct Net { bool is_phi; int pair; };   // pair: 0 = v, 1 = theta, 2 = m
  std::vector<Net> nets;
  if (multi) { nets.push_back({true, 0}); nets.push_back({true, 1}); nets.push_back({true, 2}); }
  else nets.push_back({true, 0});
  nets.push_back({false, 1}); nets.push_back({false, 0}); nets.push_back({false, 2});
  for (const Net& net : nets) {
    for (int k = 0; k < K; ++k) {
      const int base = k * W.wstep + (net.is_phi ? W.off_phi[net.pair] : W.off_ln[net.pair]);
      if (net.is_phi) {
        const int din = 5 + L, dout = multi ? L : 1;
        for (int o = 0; o < H; ++o)            // linear1.weight [H][din], inputs = [m (L), features (5)]
          for (int i = 0; i < din; ++i)
            map.push_back(base + (i < L ? W.phi_w1m + i * W.HP + o : W.phi_w1f + (i - L) * W.HP + o));
        for (int o = 0; o < H; ++o) map.push_back(base + W.phi_b1 + o);
        for (int o = 0; o < H; ++o)            // linear2.weight [H][H]
          for (int j = 0; j < H; ++j) map.push_back(base + W.phi_w2 + j * W.HP + o);
        for (int o = 0; o < H; ++o) map.push_back(base + W.phi_b2 + o);
        for (int i = 0; i < dout; ++i)         // linear4.weight [dout][H]
          for (int j = 0; j < H; ++j) map.push_back(base + W.phi_w4 + i * W.HP + j);
        for (int i = 0; i < dout; ++i) map.push_back(base + W.phi_b4 + i);
      } else {
        const int din = 4 + 2 * L, dout = (net.pair == 2) ? L : 1;
        for (int o = 0; o < H; ++o)
          for (int i = 0; i < din; ++i) map.push_back(base + W.ln_w1 + i * W.HP + o);
        for (int o = 0; o < H; ++o) map.push_back(base + W.ln_b1 + o);
        for (int o = 0; o < H; ++o)
          for (int j = 0; j < H; ++j) map.push_back(base + W.ln_w2 + j * W.HP + o);
        for (int o = 0; o < H; ++o) map.push_back(base + W.ln_b2 + o);
        for (int i = 0; i < dout; ++i)
          for (int j = 0; j < H; ++j) map.push_back(base + W.ln_wo + i * W.HP + j);
        const int bo = (net.pair == 2) ? W.ln_bo_m : W.ln_bo_s;
        for (int i = 0; i < dout; ++i) map.push_back(base + bo + i);
      }
    }
  }
  return map;
}

const gns_plan::PackMap* get_pack_map(gns_plan* plan, const ModelDims& md) {
  auto key = std::make_tuple(md.K, md.L, md.H, md.multi);
  auto it = plan->pack_maps.find(key);
  if (it != plan->pack_maps.end()) return &it->second;
  std::vector<int32_t> map = build_pack_map(md);
  gns_plan::PackMap pm;
  pm.n_canon = (int64_t)map.size();
  pm.n_packed = (int64_t)md.K * make_wlayout(md.L, md.H, md.multi != 0).wstep;
  if (cudaMalloc(&pm.d_map, map.size() * sizeof(int32_t)) != cudaSuccess) { set_error("cudaMalloc(pack map) failed"); return nullptr; }
  if (cudaMemcpy(pm.d_map, map.data(), map.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("cudaMemcpy(pack map) failed"); return nullptr;
  }
  auto res = plan->pack_maps.emplace(key, pm);
  return &res.first->second;
}

Workspace plan_workspace(const gns_plan* plan, const ModelDims& md, long long S, bool need_grad,
                         const Geometry& fwd, const Geometry& bwd, const Bwd2Geom& b2) {
  const WLayout W = make_wlayout(md.L, md.H, md.multi != 0);
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  Workspace w{};
  size_t o = 0;
  w.packed_params = o; o = align(o + (size_t)md.K * W.wstep * 4);
  if (need_grad && b2.ok) {
    // per-grid block checkpoints of the warp-specialised backward kernel (Act2Layout); the forward writes whole
    // CTA-batches, so the grid count is rounded up to its batch size
    const size_t Sg = (size_t)fwd.nbatch * fwd.G;
    const FragLayout FL = make_frag_layout(md.L, md.H, kFragTile2);
    const size_t fstep = b2.variant == 3 ? (size_t)frag3_step_floats(md.L, md.H) : (size_t)FL.step;
    w.ckpt = o; o = align(o + Sg * (md.K + 1) * (size_t)b2.a2.state * 4);
    w.pglob = o; o = align(o + Sg * md.K * 4);
    w.act = o; o = align(o + Sg * md.K * (size_t)b2.a2.step * 4);
    w.gpartial = o; o = align(o + (size_t)b2.ctas * b2.CW * md.K * fstep * 4);   // one block per (consumer) warp
    w.fragsum = o; o = align(o + (size_t)md.K * fstep * 4);
    w.packed_grad = o; o = align(o + (size_t)md.K * W.wstep * 4);
  } else if (need_grad) {
    const size_t nst = (size_t)(4 + md.L) * row_stride(plan->Ns * fwd.G);
    w.ckpt = o; o = align(o + (size_t)fwd.nbatch * md.K * nst * 4);
    w.pglob = o; o = align(o + (size_t)fwd.nbatch * md.K * fwd.G * 4);
    const ActLayout al = make_act_layout(md.H, md.multi ? 3 : 1, plan->Ns, plan->E, fwd.G, act_grid_major(bwd));
    w.act = o; o = align(o + (size_t)fwd.nbatch * md.K * (size_t)al.total * 4);
    w.gpartial = o; o = align(o + (size_t)bwd.ctas * (bwd.T / 32) * md.K * make_frag_layout(md.L, md.H).step * 4);   // one block per warp
    w.fragsum = o; o = align(o + (size_t)md.K * make_frag_layout(md.L, md.H).step * 4);
    w.packed_grad = o; o = align(o + (size_t)md.K * W.wstep * 4);
    if (md.L > 32) { w.mscratch = o; o = align(o + (size_t)bwd.ctas * 2 * md.L * bwd_bus_stride(plan->Ns * bwd.G) * 4); }
  }
  w.total = o;
  return w;
}

}  // namespace gns

using namespace gns;

extern "C" int gns_plan_create(int n_bus, int n_line, int n_gen, const int32_t* f_bus, const int32_t* t_bus,
                               const int32_t* gen_bus, int device, gns_plan** out_plan) {
  if (!out_plan) { set_error("out_plan is null"); return -1; }
  *out_plan = nullptr;
  if (n_bus <= 0 || n_line <= 0 || n_gen < 0 || !f_bus || !t_bus || (n_gen > 0 && !gen_bus)) {
    set_error("gns_plan_create: empty topology"); return -1;
  }
  if (n_bus > 65535 || n_line > 65535 || n_gen > 65535) { set_error("gns_plan_create: sizes must be < 65536"); return -1; }
  if (n_bus > n_line) {
    set_error("gns_plan_create: n_bus > n_line; the reference indexes per-line vectors with bus numbers "
              "(ref GNS/main.py:41) and would raise IndexError");
    return -1;
  }
  for (int e = 0; e < n_line; ++e)
    if (f_bus[e] < 0 || f_bus[e] >= n_bus || t_bus[e] < 0 || t_bus[e] >= n_bus) {
      set_error("gns_plan_create: line " + std::to_string(e) + " has a bus index outside [0, n_bus) "
                "(bus ids must be contiguous 1..n_bus like the reference requires, ref GNS/main.py:153)");
      return -1;
    }
  for (int j = 0; j < n_gen; ++j)
    if (gen_bus[j] < 0 || gen_bus[j] >= n_bus) { set_error("gns_plan_create: generator bus out of range"); return -1; }

  gns_plan* p = new gns_plan();
  p->device = device; p->N = n_bus; p->E = n_line; p->Gn = n_gen;
  p->f_bus.assign(f_bus, f_bus + n_line);
  p->t_bus.assign(t_bus, t_bus + n_line);
  p->gen_bus.assign(gen_bus, gen_bus + n_gen);

  // CSR in EXTERNAL bus numbering (exported, bit-exact against numpy argsort(stable)+bincount)
  csr_by(p->t_bus, n_bus, p->in_rowptr, p->in_lines);
  csr_by(p->f_bus, n_bus, p->out_rowptr, p->out_lines);
  csr_by(p->gen_bus, n_bus, p->gen_rowptr, p->gen_ids);
  // internal order: in-degree descending, stable
  p->bus_order.resize(n_bus);
  std::iota(p->bus_order.begin(), p->bus_order.end(), 0);
  std::stable_sort(p->bus_order.begin(), p->bus_order.end(), [&](int32_t a, int32_t b) {
    return (p->in_rowptr[a + 1] - p->in_rowptr[a]) > (p->in_rowptr[b + 1] - p->in_rowptr[b]);
  });
  p->bus_rank.resize(n_bus);
  for (int s = 0; s < n_bus; ++s) p->bus_rank[p->bus_order[s]] = s;

  // ---- slots: a bus with more than deg_cap incoming lines is split over 2 or 4 adjacent slots ----
  p->deg_cap = std::max(1, env_int("GNS_DEG_CAP", 3));   // measured best on case300 (2: 3.62, 3: 3.68, 4: 3.51 M grids/s)
  const int max_group = std::max(1, std::min(4, env_int("GNS_MAX_TWINS", 4)));
  auto group_size = [&](int deg) {
    int need = (deg + p->deg_cap - 1) / p->deg_cap, g = 1;
    while (g < need && g < max_group) g *= 2;
    return g;
  };
  // bus_order is degree-descending, so group sizes are non-increasing along it and every group
  // starts at a multiple of its own size: twin groups never straddle a warp's slot range.
  p->slot_bus.clear(); p->slot_primary.clear(); p->slot_in_begin.clear(); p->slot_in_end.clear(); p->slot_gsz.clear();
  std::vector<int32_t> prim_slot_of_bus(n_bus, 0);
  {
    int in_off = 0;
    for (int s0 = 0; s0 < n_bus; ++s0) {
      const int b = p->bus_order[s0];
      const int deg = p->in_rowptr[b + 1] - p->in_rowptr[b];
      const int g = group_size(deg);
      p->max_gsz = std::max(p->max_gsz, g);
      const int chunk = (deg + g - 1) / g;
      const int first = (int)p->slot_bus.size();
      prim_slot_of_bus[b] = first;
      for (int i = 0; i < g; ++i) {
        const int lo = std::min(deg, i * chunk), hi = std::min(deg, (i + 1) * chunk);
        p->slot_bus.push_back(b);
        p->slot_primary.push_back(first);
        p->slot_in_begin.push_back(in_off + lo);
        p->slot_in_end.push_back(in_off + hi);
        p->slot_gsz.push_back(g);
      }
      in_off += deg;
    }
  }
  p->Ns = (int)p->slot_bus.size();
  const int Ns = p->Ns;
  if (Ns > 65535) { set_error("gns_plan_create: too many bus slots"); delete p; return -1; }

  // device index block in slot numbering
  p->to = make_topo_offsets(n_bus, Ns, n_line, n_gen);
  std::vector<uint16_t> blk(p->to.total_ext, 0);
  for (int e = 0; e < n_line; ++e) {
    blk[p->to.fi + e] = (uint16_t)prim_slot_of_bus[f_bus[e]];
    blk[p->to.ti + e] = (uint16_t)prim_slot_of_bus[t_bus[e]];
    blk[p->to.fa + e] = (uint16_t)f_bus[e];
    blk[p->to.ta + e] = (uint16_t)t_bus[e];
  }
  {
    int oi = 0, oo = 0, og = 0;
    for (int s0 = 0; s0 < n_bus; ++s0) {       // id lists in primary-slot order
      const int b = p->bus_order[s0];
      const int ps = prim_slot_of_bus[b];
      const int in0 = oi;
      for (int q = p->in_rowptr[b]; q < p->in_rowptr[b + 1]; ++q) blk[p->to.in_ids + oi++] = (uint16_t)p->in_lines[q];
      for (int s = ps; s < ps + p->slot_gsz[ps]; ++s) blk[p->to.in_fe + s] = (uint16_t)oi;
      (void)in0;
      blk[p->to.out_b + ps] = (uint16_t)oo;
      for (int q = p->out_rowptr[b]; q < p->out_rowptr[b + 1]; ++q) blk[p->to.out_ids + oo++] = (uint16_t)p->out_lines[q];
      blk[p->to.out_e + ps] = (uint16_t)oo;
      blk[p->to.gen_b + ps] = (uint16_t)og;
      for (int q = p->gen_rowptr[b]; q < p->gen_rowptr[b + 1]; ++q) blk[p->to.gen_ids + og++] = (uint16_t)p->gen_ids[q];
      blk[p->to.gen_e + ps] = (uint16_t)og;
      blk[p->to.rank_of + b] = (uint16_t)ps;
    }
    {   // in_pos: iteration-major numbering of the in-list positions
      int maxw = 0, next = 0;
      for (int s = 0; s < Ns; ++s) maxw = std::max(maxw, p->slot_in_end[s] - p->slot_in_begin[s]);
      for (int it = 0; it < maxw; ++it)
        for (int s = 0; s < Ns; ++s)
          if (p->slot_in_end[s] - p->slot_in_begin[s] > it) blk[p->to.in_pos + p->slot_in_begin[s] + it] = (uint16_t)next++;
    }
    for (int s = 0; s < Ns; ++s) {
      blk[p->to.in_b + s] = (uint16_t)p->slot_in_begin[s];
      blk[p->to.in_e + s] = (uint16_t)p->slot_in_end[s];
      blk[p->to.ext_of + s] = (uint16_t)p->slot_bus[s];
      blk[p->to.prim_of + s] = (uint16_t)p->slot_primary[s];
      blk[p->to.gsz + s] = (uint16_t)p->slot_gsz[s];
      blk[p->to.brank + s] = (uint16_t)p->bus_rank[p->slot_bus[s]];
      p->max_walk = std::max(p->max_walk, p->slot_in_end[s] - p->slot_in_begin[s]);
      // non-primary slots keep empty out / generator ranges (zero-initialised begin == end)
    }
    for (int e = 0; e < n_line; ++e) {
      blk[p->to.fr + e] = (uint16_t)p->bus_rank[f_bus[e]];
      blk[p->to.tr + e] = (uint16_t)p->bus_rank[t_bus[e]];
    }
    for (int r = 0; r < n_bus; ++r) blk[p->to.ext_rank + r] = (uint16_t)p->bus_order[r];
    // lines as items: per activation column (in_pos) the walking slot, walk position, receiving bus rank, line id;
    // and per bus rank the columns of its in-lines (in_ids is already grouped by bus in rank order, ascending line id)
    for (int s = 0; s < Ns; ++s)
      for (int e = p->slot_in_begin[s]; e < p->slot_in_end[s]; ++e) {
        const int c = blk[p->to.in_pos + e];
        blk[p->to.col_slot + c] = (uint16_t)s;
        blk[p->to.col_it + c] = (uint16_t)(e - p->slot_in_begin[s]);
        blk[p->to.col_brank + c] = (uint16_t)p->bus_rank[p->slot_bus[s]];
        blk[p->to.col_line + c] = blk[p->to.in_ids + e];
      }
    {
      int off = 0;
      for (int r = 0; r < n_bus; ++r) {
        const int b = p->bus_order[r];
        blk[p->to.rin_b + r] = (uint16_t)off;
        const int ps = prim_slot_of_bus[b];
        const int e0 = p->slot_in_begin[ps], e1 = e0 + (p->in_rowptr[b + 1] - p->in_rowptr[b]);
        for (int e = e0; e < e1; ++e) blk[p->to.rin_cols + off++] = blk[p->to.in_pos + e];
      }
      blk[p->to.rin_b + n_bus] = (uint16_t)off;
    }
  }
  std::vector<float> expect(2 * n_line + n_gen);
  for (int e = 0; e < n_line; ++e) { expect[e] = (float)(f_bus[e] + 1); expect[n_line + e] = (float)(t_bus[e] + 1); }
  for (int j = 0; j < n_gen; ++j) expect[2 * n_line + j] = (float)(gen_bus[j] + 1);

  auto fail = [&](const char* what) { set_error(what); gns_plan_destroy(p); return -2; };
  if (device < 0) {   // host-only plan: index arrays for export, nothing uploaded (used by CPU tests)
    *out_plan = p;
    return 0;
  }
  if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail("cudaGetDeviceProperties failed");
  p->num_sms = prop.multiProcessorCount;
  p->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (cudaMalloc(&p->d_topo, blk.size() * sizeof(uint16_t)) != cudaSuccess) return fail("cudaMalloc(topo) failed");
  if (cudaMemcpy(p->d_topo, blk.data(), blk.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail("cudaMemcpy(topo) failed");
  if (cudaMalloc(&p->d_expect, expect.size() * sizeof(float)) != cudaSuccess) return fail("cudaMalloc(expect) failed");
  if (cudaMemcpy(p->d_expect, expect.data(), expect.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail("cudaMemcpy(expect) failed");
  if (cudaMalloc(&p->d_flag, sizeof(int)) != cudaSuccess) return fail("cudaMalloc(flag) failed");
  *out_plan = p;
  return 0;
}

extern "C" void gns_plan_destroy(gns_plan* p) {
  if (!p) return;
  if (p->device < 0) { delete p; return; }
  cudaSetDevice(p->device);
  if (p->d_topo) cudaFree(p->d_topo);
  if (p->d_expect) cudaFree(p->d_expect);
  if (p->d_flag) cudaFree(p->d_flag);
  for (auto& kv : p->pack_maps) if (kv.second.d_map) cudaFree(kv.second.d_map);
  for (auto& kv : p->frag_maps) if (kv.second) cudaFree(kv.second);
  delete p;
}

extern "C" int gns_plan_export(const gns_plan* p, const char* name, int32_t* out, int capacity) {
  if (!p || !name) return -1;
  const std::vector<int32_t>* v = nullptr;
  const std::string s(name);
  if (s == "in_rowptr") v = &p->in_rowptr;
  else if (s == "in_lines") v = &p->in_lines;
  else if (s == "out_rowptr") v = &p->out_rowptr;
  else if (s == "out_lines") v = &p->out_lines;
  else if (s == "gen_rowptr") v = &p->gen_rowptr;
  else if (s == "gen_ids") v = &p->gen_ids;
  else if (s == "bus_order") v = &p->bus_order;
  else if (s == "bus_rank") v = &p->bus_rank;
  else if (s == "slot_bus") v = &p->slot_bus;
  else if (s == "slot_primary") v = &p->slot_primary;
  else if (s == "slot_in_begin") v = &p->slot_in_begin;
  else if (s == "slot_in_end") v = &p->slot_in_end;
  else if (s == "slot_gsz") v = &p->slot_gsz;
  else { set_error("gns_plan_export: unknown array '" + s + "'"); return -1; }
  if (out) {
    if (capacity < (int)v->size()) { set_error("gns_plan_export: capacity too small"); return -1; }
    std::memcpy(out, v->data(), v->size() * sizeof(int32_t));
  }
  return (int)v->size();
}

extern "C" int64_t gns_param_count(int K, int L, int H, int multi) {
  return canonical_param_count(ModelDims{K, L, H, multi});
}
