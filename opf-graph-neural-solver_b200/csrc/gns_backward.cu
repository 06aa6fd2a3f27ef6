// gns_backward.cu — host side of the backward pass: geometry, launch, and the two small
// kernels that fold the per-warp gradient accumulators into the state_dict-order gradient.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "gns_backward.cuh"
#include "gns_backward2.cuh"
#include "gns_backward3.cuh"
#include "gns_host.h"
#include "../../include/gns_b200.h"

namespace gns {

BwdLauncher find_backward(int L, int H, int multi, int tmax);

int backward_extra_floats(int N, int E, int G, int L, int H, int T) {
  // PO <= L: size the tiles for the multiple-phi case
  return make_bwd_smem(N, E, G, L, H, L, T / 32, L > 32).total;
}

int make_bwd2_smem_floats(int L, int H, int E, int wstep, const Act2Layout& a2) {
  return make_bwd2_smem(L, H, E, wstep, a2).total;
}

int make_bwd3_smem_floats(int L, int H, int N, int E, int wstep, int nwarps, const Act2Layout& a2) {
  return make_bwd3_smem(L, H, N, E, wstep, nwarps, a2).total;
}
int frag3_step_floats(int L, int H) { return make_frag_layout3(L, H).step; }

int backward_ctas(const gns_plan* plan, const ModelDims&, const Geometry& g) {
  // one persistent CTA per SM slot; the exact occupancy is clamped again at launch
  return std::min(g.nbatch, plan->num_sms * std::max(1, (int)(plan->smem_optin / std::max<size_t>(g.smem_bytes, 1))));
}

// packed index (inside one step's [wstep] block) -> index inside the step's fragment-order accumulator
// block (FragLayout), or -1 for padding and for the entries the backward kernel never writes (W4 / b4 and
// the W1 slice behind the fused block, which unfuse_grads_kernel derives).  Mirrors the GEMM calls of
// gns_backward_kernel one to one.
// `v2`: the warp-specialised kernel computes the scalar nets' output layer with the roles swapped (hid = [h2, 1],
// one wide row = the output adjoint), so its cells are (row 0, column c) instead of (row r, column 0).
std::vector<int32_t> build_frag_map(const ModelDims& md, int variant) {
  const bool v2 = variant >= 2;
  const int L = md.L, H = md.H;
  const bool multi = md.multi != 0;
  const WLayout W = make_wlayout(L, H, multi);
  const FragLayout F = variant == 3 ? make_frag_layout3(L, H) : make_frag_layout(L, H, v2 ? kFragTile2 : kFragTile);
  auto frag_index = [&](int r, int c) { return v2 ? gns::frag2_index(r, c) : gns::frag_index(r, c); };
  std::vector<int32_t> inv(W.wstep, -1);
  auto put = [&](int packed, int frag) { inv[packed] = frag; };
  for (int q = 0; q < 3; ++q) {
    const int fb = q * F.net;
    if (multi || q == 0) {                       // phi net of this pair (single phi: pair 0 carries it)
      const int pb = multi ? q * W.phi_size : 0;
      for (int r = 0; r < H + 1; ++r)
        for (int c = 0; c < H; ++c) put(pb + (r < H ? W.phi_w2 + r * W.HP : W.phi_b2) + c, fb + F.w2l + frag_index(r, c));
      for (int r = 0; r < 5; ++r)
        for (int c = 0; c < H; ++c) put(pb + W.phi_w1f + r * W.HP + c, fb + F.w1f + frag_index(r, c));
      for (int r = 0; r < L + 1; ++r)
        for (int c = 0; c < H; ++c) put(pb + (r < L ? W.phi_w1m + r * W.HP : W.phi_b1) + c, fb + F.w1m + frag_index(r, c));
    }
    const int lb = W.off_ln[0] + q * W.ln_size_s;
    if (q < 2) {
      for (int r = 0; r < H + 1; ++r)
        put(lb + (r < H ? W.ln_wo + r : W.ln_bo_s), fb + F.out + (v2 ? frag_index(0, r) : frag_index(r, 0)));
    } else {
      for (int r = 0; r < L; ++r)
        for (int c = 0; c < H + 1; ++c)
          put(lb + (c < H ? W.ln_wo + r * W.HP + c : W.ln_bo_m + r),
              fb + F.out + (variant == 3 ? frag3_out_index(4 + r, c) : frag_index(r, c)));
    }
    for (int r = 0; r < H + 1; ++r)
      for (int c = 0; c < H; ++c) put(lb + (r < H ? W.ln_w2 + r * W.HP : W.ln_b2) + c, fb + F.w2 + frag_index(r, c));
    const int mfb = W.off_mf[0] + q * W.mf_size;
    for (int r = 0; r < 4 + L + H + 2; ++r)
      for (int c = 0; c < H; ++c) {
        const int packed = r < 4 + L ? lb + W.ln_w1 + r * W.HP + c
                                     : (r < 4 + L + H + 1 ? mfb + (r - 4 - L) * W.HP + c : lb + W.ln_b1 + c);
        put(packed, fb + F.w1 + frag_index(r, c));
      }
  }
  return inv;
}

static const int32_t* get_frag_map(gns_plan* plan, const ModelDims& md, int v2) {
  auto key = std::make_tuple(md.L, md.H, md.multi + 2 * v2);
  auto it = plan->frag_maps.find(key);
  if (it != plan->frag_maps.end()) return it->second;
  const std::vector<int32_t> inv = build_frag_map(md, v2);
  int32_t* d = nullptr;
  if (cudaMalloc(&d, inv.size() * sizeof(int32_t)) != cudaSuccess ||
      cudaMemcpy(d, inv.data(), inv.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("frag map upload failed");
    return nullptr;
  }
  plan->frag_maps.emplace(key, d);
  return d;
}

// fragsum[f] = sum_w gacc[w][f] over the per-warp accumulator blocks, in fragment order: consecutive threads read
// consecutive floats of one block (coalesced), 16 part lanes per block split the blocks between them and are folded
// in a fixed order (deterministic).  A block covers 64 cells: 264 CTAs for K=4 on case-sized models.
constexpr int kRedF = 64, kRedP = 16;
__global__ void __launch_bounds__(kRedF * kRedP) reduce_partials_kernel(const float* __restrict__ gacc, float* __restrict__ fragsum,
                                                                        long long n, int nparts) {
  __shared__ float sm[kRedP][kRedF];
  const int fx = threadIdx.x, py = threadIdx.y;
  const long long f = blockIdx.x * (long long)kRedF + fx;
  float s0 = 0.f, s1 = 0.f;
  if (f < n) {
    int w = py;
    for (; w + kRedP < nparts; w += 2 * kRedP) {
      s0 += gacc[(size_t)w * n + f];
      s1 += gacc[(size_t)(w + kRedP) * n + f];
    }
    if (w < nparts) s0 += gacc[(size_t)w * n + f];
  }
  sm[py][fx] = s0 + s1;
  __syncthreads();
  if (py == 0 && f < n) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < kRedP; ++j) t += sm[j][fx];
    fragsum[f] = t;
  }
}

// packed[k][p] = fragsum[k][inv[p]]   (0 where inv[p] < 0): fragment order -> packed parameter layout
__global__ void gather_frag_kernel(const float* __restrict__ fragsum, const int32_t* __restrict__ inv,
                                   float* __restrict__ packed, int K, int wstep, int fstep) {
  const long long n = (long long)K * wstep;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(p / wstep);
    const int f = inv[p - (long long)k * wstep];
    packed[p] = f >= 0 ? fragsum[(size_t)k * fstep + f] : 0.f;
  }
}

// Chain rule of the fused aggregate->hidden block back to the parameters it was built from:
//   dW4[i][j]  = sum_o W1S[i][o] dM[j][o]
//   dW1S[i][o] = sum_j W4[i][j] dM[j][o] + b4[i] dc[o]
//   db4[i]     = sum_o W1S[i][o] dc[o]
// (single phi: W4 / b4 are shared by the three pairs and receive the sum).  One thread per
// destination element; these rows get no other contribution.
__global__ void unfuse_grads_kernel(float* __restrict__ grad, const float* __restrict__ packed, WLayout W, int K) {
  const int nphi = W.multi ? 3 : 1;
  const int n_w4 = nphi * W.PO * W.H, n_b4 = nphi * W.PO, n_w1s = 3 * W.PO * W.H;
  const int per_step = n_w4 + n_b4 + n_w1s;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < K * per_step; t += gridDim.x * blockDim.x) {
    const int k = t / per_step;
    int r = t - k * per_step;
    const float* pb = packed + (size_t)k * W.wstep;
    float* gb = grad + (size_t)k * W.wstep;
    if (r < n_w4 + n_b4) {
      const bool is_b = r >= n_w4;
      if (is_b) r -= n_w4;
      const int p = is_b ? r / W.PO : r / (W.PO * W.H);
      const int rem = is_b ? r - p * W.PO : r - p * W.PO * W.H;
      const int i = is_b ? rem : rem / W.H, j = is_b ? 0 : rem - (rem / W.H) * W.H;
      float acc = 0.f;
      for (int q = (W.multi ? p : 0); q < (W.multi ? p + 1 : 3); ++q) {
        const float* w1s = pb + W.off_ln[q] + W.ln_w1 + (4 + W.L) * W.HP + i * W.HP;
        const float* dm = gb + W.off_mf[q] + (is_b ? W.H * W.HP : j * W.HP);
        for (int o = 0; o < W.H; ++o) acc = fmaf(w1s[o], dm[o], acc);
      }
      gb[W.off_phi[p] + (is_b ? W.phi_b4 + i : W.phi_w4 + i * W.HP + j)] = acc;
    } else {
      r -= n_w4 + n_b4;
      const int q = r / (W.PO * W.H);
      const int rem = r - q * W.PO * W.H;
      const int i = rem / W.H, o = rem - i * W.H;
      const float* phi = pb + W.off_phi[W.multi ? q : 0];
      const float* dm = gb + W.off_mf[q];
      float acc = phi[W.phi_b4 + i] * dm[W.H * W.HP + o];
      for (int j = 0; j < W.H; ++j) acc = fmaf(phi[W.phi_w4 + i * W.HP + j], dm[j * W.HP + o], acc);
      gb[W.off_ln[q] + W.ln_w1 + (4 + W.L + i) * W.HP + o] = acc;
    }
  }
}

// canon[c] = packed[map[c]]
__global__ void unpack_grads_kernel(const float* __restrict__ packed, const int32_t* __restrict__ map,
                                    float* __restrict__ canon, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    canon[i] = packed[map[i]];
}

// fragment-order accumulators -> state_dict-order gradient (shared by both backward kernels)
static int fold_gradients(gns_plan* plan, const ModelDims& md, const float* gacc, int nparts, const int32_t* d_inv,
                          const float* packed_params, char* wsb, const Workspace& ws, float* grad_params, cudaStream_t st,
                          int tile = kFragTile) {
  const gns_plan::PackMap* pm = get_pack_map(plan, md);
  if (!pm) return -2;
  const WLayout W = make_wlayout(md.L, md.H, md.multi != 0);
  const FragLayout FL = tile == 0 ? make_frag_layout3(md.L, md.H) : make_frag_layout(md.L, md.H, tile);
  const size_t per_part = (size_t)md.K * FL.step;
  float* packed_grad = reinterpret_cast<float*>(wsb + ws.packed_grad);
  const int th = 256;
  float* fragsum = reinterpret_cast<float*>(wsb + ws.fragsum);
  reduce_partials_kernel<<<(unsigned)((per_part + kRedF - 1) / kRedF), dim3(kRedF, kRedP), 0, st>>>(gacc, fragsum, (long long)per_part, nparts);
  const int bl = (int)std::min<long long>(((long long)md.K * W.wstep + th - 1) / th, 2048);
  gather_frag_kernel<<<bl, th, 0, st>>>(fragsum, d_inv, packed_grad, md.K, W.wstep, FL.step);
  const int nphi = md.multi ? 3 : 1, PO = md.multi ? md.L : 1;
  const int tot = md.K * (nphi * PO * md.H + nphi * PO + 3 * PO * md.H);
  unfuse_grads_kernel<<<(tot + th - 1) / th, th, 0, st>>>(packed_grad, packed_params, W, md.K);
  const int bl2 = (int)std::min<long long>((pm->n_canon + th - 1) / th, 1024);
  unpack_grads_kernel<<<bl2, th, 0, st>>>(packed_grad, pm->d_map, grad_params, pm->n_canon);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error(std::string("gradient reduce: ") + cudaGetErrorString(e)); return -2; }
  return 0;
}

// warp-specialised kernel (large grids, see gns_backward2.cuh)
static int run_backward2(gns_plan* plan, const ModelDims& md, const Bwd2Geom& b2, const Workspace& ws, const float* buses,
                         const float* lines, const float* gens, long long S, float gamma, const float* grad_total,
                         const float* grad_last, const float* grad_v, const float* grad_theta, float* grad_params,
                         char* wsb, cudaStream_t st) {
  const WLayout W = make_wlayout(md.L, md.H, md.multi != 0);
  const FragLayout FL = make_frag_layout(md.L, md.H, kFragTile2);
  const size_t per_part = (size_t)md.K * FL.step;
  const int32_t* d_inv = get_frag_map(plan, md, 2);
  if (!d_inv) return -2;
  const int nparts = b2.ctas * b2.CW;
  float* gacc = reinterpret_cast<float*>(wsb + ws.gpartial);
  cudaError_t e = cudaMemsetAsync(gacc, 0, (size_t)nparts * per_part * 4, st);
  if (e != cudaSuccess) { set_error(std::string("memset gacc: ") + cudaGetErrorString(e)); return -2; }
  Bwd2Launcher launch = find_backward2(md.L, md.H, md.multi);
  if (!launch) { set_error("gns_backward: warp-specialised kernel not built for these dims"); return -1; }
  Bwd2Args a{};
  a.params = reinterpret_cast<const float*>(wsb + ws.packed_params);   // packed by gns_forward
  a.buses = buses; a.lines = lines; a.gens = gens;
  a.ck2 = reinterpret_cast<const float*>(wsb + ws.ckpt);
  a.pglob = reinterpret_cast<const float*>(wsb + ws.pglob);
  a.act = reinterpret_cast<const float*>(wsb + ws.act);
  a.grad_total = grad_total; a.grad_last = grad_last; a.grad_v = grad_v; a.grad_theta = grad_theta;
  a.gacc = gacc;
  a.topo = plan->d_topo;
  a.S = S; a.N = plan->N; a.Ns = plan->Ns; a.E = plan->E; a.Gn = plan->Gn; a.K = md.K;
  a.PW = b2.PW; a.CW = b2.CW; a.maxwalk = plan->max_walk;
  {
    // warp -> role.  Mode 0: the first PW warps produce.  Mode 1 (GNS_BWD2_ROLEMAP=1): warps of SM sub-partitions 0, 1
    // (warp % 4 < 2) produce and those of 2, 3 consume, so the warps that share an L0 instruction cache run the same code.
    const char* rm = std::getenv("GNS_BWD2_ROLEMAP");
    const int mode = rm ? std::atoi(rm) : 0;
    int np = 0, nc = 0;
    const int nw = b2.PW + b2.CW;
    for (int w = 0; w < nw; ++w) {
      bool prod = w < b2.PW;
      if (mode == 1) prod = (w % 4 < 2) ? (np < b2.PW) : (nc >= b2.CW);
      a.role[w] = prod ? (unsigned char)(np++) : (unsigned char)(0x80 | nc++);
    }
  }
  a.a2 = b2.a2;
  a.sm = make_bwd2_smem(md.L, md.H, plan->E, W.wstep, b2.a2);
  a.to = plan->to;
  for (int k = 0; k < md.K; ++k) a.wk[k] = (float)std::pow((double)gamma, (double)(md.K - k));
  e = launch(a, b2, plan->num_sms, st);
  if (e != cudaSuccess) { set_error(std::string("backward2 launch: ") + cudaGetErrorString(e)); return -2; }
  return fold_gradients(plan, md, gacc, nparts, d_inv, a.params, wsb, ws, grad_params, st, kFragTile2);
}

// fragment-space kernel (gns_backward3.cuh)
static int run_backward3(gns_plan* plan, const ModelDims& md, const Bwd2Geom& b2, const Workspace& ws, const float* buses,
                         const float* lines, const float* gens, long long S, float gamma, const float* grad_total,
                         const float* grad_last, const float* grad_v, const float* grad_theta, float* grad_params,
                         char* wsb, cudaStream_t st) {
  const WLayout W = make_wlayout(md.L, md.H, md.multi != 0);
  const FragLayout FL = make_frag_layout3(md.L, md.H);
  const size_t per_part = (size_t)md.K * FL.step;
  const int32_t* d_inv = get_frag_map(plan, md, 3);
  if (!d_inv) return -2;
  const int nw = b2.T / 32;
  const int nparts = b2.ctas * b2.parts_per_cta;
  float* gacc = reinterpret_cast<float*>(wsb + ws.gpartial);
  cudaError_t e = cudaMemsetAsync(gacc, 0, (size_t)nparts * per_part * 4, st);
  if (e != cudaSuccess) { set_error(std::string("memset gacc: ") + cudaGetErrorString(e)); return -2; }
  Bwd3Launcher launch = find_backward3(md.L, md.H, md.multi);
  if (!launch) { set_error("gns_backward: fragment-space kernel not built for these dims"); return -1; }
  Bwd3Args a{};
  a.params = reinterpret_cast<const float*>(wsb + ws.packed_params);
  a.buses = buses; a.lines = lines; a.gens = gens;
  a.ck2 = reinterpret_cast<const float*>(wsb + ws.ckpt);
  a.pglob = reinterpret_cast<const float*>(wsb + ws.pglob);
  a.act = reinterpret_cast<const float*>(wsb + ws.act);
  a.grad_total = grad_total; a.grad_last = grad_last; a.grad_v = grad_v; a.grad_theta = grad_theta;
  a.gacc = gacc;
  a.acc_shared = b2.parts_per_cta == 1 ? 1 : 0;
  a.topo = plan->d_topo;
  a.S = S; a.N = plan->N; a.Ns = plan->Ns; a.E = plan->E; a.Gn = plan->Gn; a.K = md.K;
  a.a2 = b2.a2;
  a.sm = make_bwd3_smem(md.L, md.H, plan->N, plan->E, W.wstep, nw, b2.a2);
  a.to = plan->to;
  for (int k = 0; k < md.K; ++k) a.wk[k] = (float)std::pow((double)gamma, (double)(md.K - k));
  e = launch(a, b2, plan->num_sms, st);
  if (e != cudaSuccess) { set_error(std::string("backward3 launch: ") + cudaGetErrorString(e)); return -2; }
  return fold_gradients(plan, md, gacc, nparts, d_inv, a.params, wsb, ws, grad_params, st, 0);
}

int run_backward(gns_plan* plan, const ModelDims& md, const float* params, const float* buses, const float* lines,
                 const float* gens, long long S, float gamma, const float* grad_total, const float* grad_last,
                 const float* grad_v, const float* grad_theta, float* grad_params, void* workspace,
                 long long workspace_bytes, cudaStream_t st) {
  (void)params;
  Geometry gf, gb;
  if (!choose_geometry(plan, md, S, false, &gf)) return -1;
  if (!choose_geometry(plan, md, S, true, &gb)) return -1;
  gb.ctas = backward_ctas(plan, md, gb);
  const Bwd2Geom b2 = choose_bwd2(plan, md, S);
  const Workspace ws = plan_workspace(plan, md, S, true, gf, gb, b2);
  if (b2.ok) {
    if ((long long)ws.total > workspace_bytes) { set_error("gns_backward: workspace too small"); return -1; }
    if (b2.variant == 3)
      return run_backward3(plan, md, b2, ws, buses, lines, gens, S, gamma, grad_total, grad_last, grad_v, grad_theta,
                           grad_params, static_cast<char*>(workspace), st);
    return run_backward2(plan, md, b2, ws, buses, lines, gens, S, gamma, grad_total, grad_last, grad_v, grad_theta,
                         grad_params, static_cast<char*>(workspace), st);
  }
  if ((long long)ws.total > workspace_bytes) { set_error("gns_backward: workspace too small"); return -1; }
  char* wsb = static_cast<char*>(workspace);
  const int nwarps = gb.T / 32;
  const FragLayout FL = make_frag_layout(md.L, md.H);
  const size_t per_part = (size_t)md.K * FL.step;
  const int32_t* d_inv = get_frag_map(plan, md, 0);
  if (!d_inv) return -2;
  // GNS_DETERMINISTIC=0: the warps of a CTA share one accumulator block (10x smaller, L2 resident); the order of
  // their floating-point reductions is then not fixed, so gradients are reproducible to rounding only
  const char* det_env = std::getenv("GNS_DETERMINISTIC");
  const bool per_warp = !(det_env && det_env[0] == '0');
  const int nparts = gb.ctas * (per_warp ? nwarps : 1);
  float* gacc = reinterpret_cast<float*>(wsb + ws.gpartial);
  cudaError_t e = cudaMemsetAsync(gacc, 0, (size_t)nparts * per_part * 4, st);
  if (e != cudaSuccess) { set_error(std::string("memset gacc: ") + cudaGetErrorString(e)); return -2; }

  BwdLauncher launch = find_backward(md.L, md.H, md.multi, gb.tmax);
  if (!launch) {
    set_error("gns_backward: no backward kernel built for latent_dim=" + std::to_string(md.L) + " hidden_dim=" +
              std::to_string(md.H));
    return -1;
  }
  BwdArgs a{};
  a.params = reinterpret_cast<const float*>(wsb + ws.packed_params);   // packed by gns_forward
  a.buses = buses; a.lines = lines; a.gens = gens;
  a.ckpt = reinterpret_cast<const float*>(wsb + ws.ckpt);
  a.pglob = reinterpret_cast<const float*>(wsb + ws.pglob);
  a.act = reinterpret_cast<const float*>(wsb + ws.act);
  a.grad_total = grad_total; a.grad_last = grad_last; a.grad_v = grad_v; a.grad_theta = grad_theta;
  a.gacc = gacc;
  a.acc_per_warp = per_warp ? 1 : 0;
  a.topo = plan->d_topo;
  a.S = S; a.N = plan->N; a.Ns = plan->Ns; a.E = plan->E; a.Gn = plan->Gn; a.K = md.K; a.NGQ = gb.NGQ; a.G = gb.G; a.nbatch = gb.nbatch;
  a.NGs = bwd_bus_stride(plan->Ns * gb.G); a.EGs = row_stride(plan->E * gb.G);
  a.Gf = gf.G; a.NGs_f = row_stride(plan->Ns * gf.G); a.EGs_f = row_stride(plan->E * gf.G);
  a.al = make_act_layout(md.H, md.multi ? 3 : 1, plan->Ns, plan->E, gf.G, act_grid_major(gb));
  std::memcpy(a.grp_of_warp, gb.grp_of_warp, 32);
  a.sm = gb.sm;
  a.bs = make_bwd_smem(plan->Ns, plan->E, gb.G, md.L, md.H, md.L, nwarps, md.L > 32);
  a.mscratch = md.L > 32 ? reinterpret_cast<float*>(wsb + ws.mscratch) : nullptr;
  a.to = plan->to;
  for (int k = 0; k < md.K; ++k) a.wk[k] = (float)std::pow((double)gamma, (double)(md.K - k));
  e = launch(a, gb, st);
  if (e != cudaSuccess) { set_error(std::string("backward launch: ") + cudaGetErrorString(e)); return -2; }
  return fold_gradients(plan, md, gacc, nparts, d_inv, a.params, wsb, ws, grad_params, st);
}

}  // namespace gns

extern "C" int gns_layout_export(const char* name, int K, int latent_dim, int hidden_dim, int multiple_phi,
                                 int32_t* out, int capacity) {
  using namespace gns;
  if (!name || K < 1 || K > kMaxK || !gns_dims_supported(latent_dim, hidden_dim)) {
    set_error("gns_layout_export: bad arguments"); return -1;
  }
  const ModelDims md{K, latent_dim, hidden_dim, multiple_phi};
  std::vector<int32_t> v;
  const std::string n(name);
  if (n == "pack") v = build_pack_map(md);
  else if (n == "frag") v = build_frag_map(md, 0);
  else if (n == "frag2") v = build_frag_map(md, 2);
  else if (n == "frag3") v = build_frag_map(md, 3);
  else { set_error("gns_layout_export: unknown map '" + n + "'"); return -1; }
  if (out) {
    if (capacity < (int)v.size()) { set_error("gns_layout_export: capacity too small"); return -1; }
    std::memcpy(out, v.data(), v.size() * sizeof(int32_t));
  }
  return (int)v.size();
}

