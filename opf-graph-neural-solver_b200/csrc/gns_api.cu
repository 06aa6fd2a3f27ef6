// gns_api.cu — the C ABI entry points (include/gns_b200.h) and the small utility kernels
// around the two persistent kernels: parameter packing, gradient un-packing, topology
// check, Adam, and the FP32-FMA throughput probe used as the roofline denominator.
#include <cmath>
#include <cstring>

#include "gns_host.h"
#include "../../include/gns_b200.h"

namespace gns {
const char* last_error_cstr();
int run_backward(gns_plan* plan, const ModelDims& md, const float* params, const float* buses, const float* lines,
                 const float* gens, long long S, float gamma, const float* grad_total, const float* grad_last,
                 const float* grad_v, const float* grad_theta, float* grad_params, void* workspace,
                 long long workspace_bytes, cudaStream_t st);
int backward_ctas(const gns_plan* plan, const ModelDims& md, const Geometry& g);

// packed[map[c]] = canon[c]   (packed is zero-filled first so the padding lanes stay 0)
__global__ void pack_params_kernel(const float* __restrict__ canon, const int32_t* __restrict__ map,
                                   float* __restrict__ packed, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    packed[map[i]] = canon[i];
}

// M[j][o] = sum_i W4[i][j] W1S[i][o],  c[o] = sum_i b4[i] W1S[i][o]  for every step and pair
// (W1S = rows 4+L.. of the L-net's first layer).  One thread per output element.
__global__ void fuse_params_kernel(float* __restrict__ packed, WLayout W, int K) {
  const int per_pair = W.H * W.H + W.H;
  const int total = K * 3 * per_pair;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int k = t / (3 * per_pair);
    int r = t - k * 3 * per_pair;
    const int q = r / per_pair;
    r -= q * per_pair;
    float* base = packed + (size_t)k * W.wstep;
    const float* phi = base + W.off_phi[W.multi ? q : 0];
    const float* w1s = base + W.off_ln[q] + W.ln_w1 + (4 + W.L) * W.HP;
    float acc = 0.f;
    if (r < W.H * W.H) {
      const int j = r / W.H, o = r - j * W.H;
      for (int i = 0; i < W.PO; ++i) acc = fmaf(phi[W.phi_w4 + i * W.HP + j], w1s[i * W.HP + o], acc);
      base[W.off_mf[q] + j * W.HP + o] = acc;
    } else {
      const int o = r - W.H * W.H;
      for (int i = 0; i < W.PO; ++i) acc = fmaf(phi[W.phi_b4 + i], w1s[i * W.HP + o], acc);
      base[W.off_mf[q] + W.H * W.HP + o] = acc;
    }
  }
}

__global__ void check_topology_kernel(const float* __restrict__ lines, const float* __restrict__ gens,
                                      const float* __restrict__ expect, long long S, int E, int Gn, int* flag) {
  const long long per = 2LL * E + Gn;
  const long long total = S * per;
  int bad = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / per;
    const int r = (int)(i - s * per);
    float got;
    if (r < E) got = lines[(s * E + r) * 7 + 0];
    else if (r < 2 * E) got = lines[(s * E + (r - E)) * 7 + 1];
    else got = gens[(s * Gn + (r - 2 * E)) * 7 + 0];
    bad |= (got != expect[r]);
  }
  if (bad) atomicOr(flag, 1);
}

// Compact input format -> the reference's packed rows (ref GNS/utils.py:17-41).  Between the samples of one case only
// Pd, Qd | r, x, b, tau, shift | vg, Pg vary (ref GNS/augment_grids.py:35-53); bus_i, type, Gs, Bs | f_bus, t_bus |
// bus_i, Pmax, Pmin, qg are constants of the case and Pg_set is a copy of Pg (ref GNS/utils.py:38), so a host batch
// ships 2N + 5E + 2Gn floats per grid instead of 6N + 7E + 7Gn.  One thread per output row.
__global__ void expand_inputs_kernel(const float* __restrict__ bv, const float* __restrict__ lv, const float* __restrict__ gv,
                                     const float* __restrict__ cb, const float* __restrict__ cl, const float* __restrict__ cg,
                                     long long S, int N, int E, int Gn, float* __restrict__ buses, float* __restrict__ lines,
                                     float* __restrict__ gens) {
  const long long per = (long long)N + E + Gn, total = S * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / per;
    int r = (int)(i - s * per);
    if (r < N) {
      const float* v = bv + (s * N + r) * 2;
      const float* c = cb + r * 4;
      float* o = buses + (s * N + r) * 6;
      o[0] = c[0]; o[1] = c[1]; o[2] = v[0]; o[3] = v[1]; o[4] = c[2]; o[5] = c[3];
    } else if (r < N + E) {
      r -= N;
      const float* v = lv + (s * E + r) * 5;
      const float* c = cl + r * 2;
      float* o = lines + (s * E + r) * 7;
      o[0] = c[0]; o[1] = c[1]; o[2] = v[0]; o[3] = v[1]; o[4] = v[2]; o[5] = v[3]; o[6] = v[4];
    } else {
      r -= N + E;
      const float* v = gv + (s * Gn + r) * 2;
      const float* c = cg + r * 4;
      float* o = gens + (s * Gn + r) * 7;
      o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = v[1]; o[4] = v[0]; o[5] = c[3]; o[6] = v[1];
    }
  }
}

// torch.optim.Adam defaults (ref GNS/main.py:243): no weight decay, no amsgrad
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

// register-only FFMA loop: 16 independent chains per thread, 2 flops per FFMA
__global__ void ffma_probe_kernel(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-6f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456f) out[0] = s;   // keep the loop alive
}

// same probe with the packed FFMA2 (fma.rn.f32x2) instruction of sm_100: 2 FMAs per lane per issue slot
__global__ void ffma2_probe_kernel(float* out, int iters, float a, float b) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-6f + i, threadIdx.x * 2e-6f + i);
  const float2 bb = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], make_float2(a, a), bb);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  if (s == 123.456f) out[0] = s;
}

static bool check_plan(const gns_plan* plan) {
  if (plan && plan->device < 0) { set_error("host-only plan (device < 0) cannot launch kernels"); return false; }
  return plan != nullptr;
}

static bool check_model(const ModelDims& md) {
  if (md.K < 1 || md.K > kMaxK) { set_error("K must be in [1, " + std::to_string(kMaxK) + "]"); return false; }
  if (!gns_dims_supported(md.L, md.H)) {
    set_error("no sm_100a kernel instantiated for latent_dim=" + std::to_string(md.L) + ", hidden_dim=" +
              std::to_string(md.H) + " (built: latent_dim in {10,20,64}, hidden_dim=10); there is no fallback path");
    return false;
  }
  return true;
}

}  // namespace gns

using namespace gns;

extern "C" const char* gns_last_error(void) { return last_error_cstr(); }
extern "C" const char* gns_version(void) { return "gns_b200 0.1 (sm_100a)"; }

extern "C" int gns_dims_supported(int L, int H) {
  return (H == 10 && (L == 10 || L == 20 || L == 64)) ? 1 : 0;
}

extern "C" int64_t gns_workspace_bytes(const gns_plan* plan, int64_t S, int K, int L, int H, int multi, int need_grad) {
  if (!plan || S <= 0) { set_error("gns_workspace_bytes: bad arguments"); return -1; }
  ModelDims md{K, L, H, multi};
  if (!check_model(md)) return -1;
  Geometry gf, gb;
  if (!choose_geometry(plan, md, S, false, &gf)) return -1;
  if (need_grad) {
    if (!choose_geometry(plan, md, S, true, &gb)) return -1;
    gb.ctas = backward_ctas(plan, md, gb);
  }
  const Bwd2Geom b2 = need_grad ? choose_bwd2(plan, md, S) : Bwd2Geom{};
  return (int64_t)plan_workspace(plan, md, S, need_grad != 0, gf, gb, b2).total;
}

extern "C" int gns_launch_info(const gns_plan* plan, int64_t S, int K, int L, int H, int multi, int backward, int32_t out[8]) {
  if (!plan || !out) return -1;
  ModelDims md{K, L, H, multi};
  if (!check_model(md)) return -1;
  Geometry g;
  if (!choose_geometry(plan, md, S, backward != 0, &g)) return -1;
  // persistent kernels: CTAs = min(batches, SMs x CTAs resident per SM); residency from shared memory and the
  // register budget of the launch-bounds variant (the launchers clamp with the occupancy API, same numbers)
  const int regs = g.tmax == 320 ? 200 : (g.tmax == 384 ? 168 : 64);
  const int per_sm = std::max(1, std::min((int)((size_t)233472 / (g.smem_bytes + 1024)), 65536 / (g.T * regs)));
  int ctas = backward ? std::min(backward_ctas(plan, md, g), plan->num_sms * per_sm) : std::min(g.nbatch, plan->num_sms * per_sm);
  if (backward) {
    const Bwd2Geom b2 = choose_bwd2(plan, md, S);
    if (b2.ok) {   // warp-specialised kernel: one grid per CTA at a time
      out[0] = 1; out[1] = b2.T; out[2] = (int32_t)b2.smem_bytes; out[3] = b2.ctas; out[4] = b2.variant == 3 ? 16 * b2.bt : 2;
      out[5] = (int32_t)std::min<int64_t>(S, 2147483647); out[6] = plan->num_sms;
      out[7] = b2.variant == 3 ? (b2.bt == 2 ? 320 : (b2.T <= 256 ? 256 : 640)) : 384;
      return 0;
    }
  }
  out[0] = g.G; out[1] = g.T; out[2] = (int32_t)g.smem_bytes; out[3] = ctas; out[4] = g.VG;
  out[5] = g.nbatch; out[6] = plan->num_sms; out[7] = g.tmax;
  return 0;
}

static int forward_impl(const gns_plan* cplan, const float* params, const float* buses, const float* lines,
                        const float* gens, const float* cbus, const float* cgen, int64_t S, int K, int L, int H, int multi,
                        float gamma, float* v, float* theta, float* total_loss, float* last_loss, void* workspace,
                        int64_t workspace_bytes, int need_grad, void* stream) {
  gns_plan* plan = const_cast<gns_plan*>(cplan);
  const bool compact = cbus != nullptr;
  if (!plan || !params || !buses || !lines || !gens || !v || !theta || !total_loss || !last_loss || !workspace) {
    set_error("gns_forward: null argument"); return -1;
  }
  if (S <= 0) { set_error("gns_forward: empty batch"); return -1; }
  if (!check_plan(plan)) return -1;
  ModelDims md{K, L, H, multi};
  if (!check_model(md)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaSetDevice(plan->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return -2; }
  Geometry gf, gb;
  if (!choose_geometry(plan, md, S, false, &gf)) return -1;
  if (need_grad) {
    if (!choose_geometry(plan, md, S, true, &gb)) return -1;
    gb.ctas = backward_ctas(plan, md, gb);
  }
  const Bwd2Geom b2 = need_grad ? choose_bwd2(plan, md, S) : Bwd2Geom{};
  const Workspace ws = plan_workspace(plan, md, S, need_grad != 0, gf, gb, b2);
  if ((int64_t)ws.total > workspace_bytes) {
    set_error("gns_forward: workspace too small (" + std::to_string(workspace_bytes) + " < " + std::to_string(ws.total) + ")");
    return -1;
  }
  const gns_plan::PackMap* pm = get_pack_map(plan, md);
  if (!pm) return -2;
  char* wsb = static_cast<char*>(workspace);
  float* packed = reinterpret_cast<float*>(wsb + ws.packed_params);
  cudaError_t e = cudaMemsetAsync(packed, 0, (size_t)pm->n_packed * 4, st);
  if (e != cudaSuccess) { set_error(std::string("memset packed: ") + cudaGetErrorString(e)); return -2; }
  {
    const int th = 256;
    const int bl = (int)std::min<int64_t>((pm->n_canon + th - 1) / th, 1024);
    pack_params_kernel<<<bl, th, 0, st>>>(params, pm->d_map, packed, pm->n_canon);
    const WLayout Wl = make_wlayout(L, H, multi != 0);
    const int tot = K * 3 * (H * H + H);
    fuse_params_kernel<<<(tot + th - 1) / th, th, 0, st>>>(packed, Wl, K);
  }
  FwdLauncher launch = find_forward(L, H, multi, gf.VG, gf.tmax);
  if (!launch) { set_error("gns_forward: kernel variant not built"); return -1; }
  FwdArgs a{};
  a.params = packed;
  a.buses = buses; a.lines = lines; a.gens = gens;
  a.v = v; a.theta = theta; a.total = total_loss; a.last = last_loss;
  a.ckpt = need_grad ? reinterpret_cast<float*>(wsb + ws.ckpt) : nullptr;
  a.pglob = need_grad ? reinterpret_cast<float*>(wsb + ws.pglob) : nullptr;
  a.act = need_grad ? reinterpret_cast<float*>(wsb + ws.act) : nullptr;
  a.topo = plan->d_topo;
  a.S = S; a.N = plan->N; a.Ns = plan->Ns; a.E = plan->E; a.Gn = plan->Gn; a.K = K; a.NGQ = gf.NGQ; a.G = gf.G; a.nbatch = gf.nbatch;
  a.NGs = row_stride(plan->Ns * gf.G); a.EGs = row_stride(plan->E * gf.G);
  a.need_grad = need_grad ? (b2.ok ? 2 : 1) : 0;     // 2: per-grid block checkpoints (Act2Layout) for gns_backward2
  a.a2 = b2.a2;
  a.ck2 = a.ckpt;
  a.al = make_act_layout(H, multi ? 3 : 1, plan->Ns, plan->E, gf.G, (need_grad && !b2.ok) ? act_grid_major(gb) : true);
  a.compact = compact ? 1 : 0;
  a.cbus = cbus; a.cgen = cgen;
  a.use_tma = (gf.sm.stage_l != 0 && ((uintptr_t)buses % 16 == 0) && ((uintptr_t)lines % 16 == 0) &&
               ((uintptr_t)gens % 16 == 0)) ? 1 : 0;
  std::memcpy(a.grp_of_warp, gf.grp_of_warp, 32);
  a.sm = gf.sm; a.to = plan->to;
  for (int k = 0; k < K; ++k) a.wk[k] = (float)std::pow((double)gamma, (double)(K - k));   // ref GNS/main.py:198
  e = launch(a, gf, st);
  if (e != cudaSuccess) { set_error(std::string("forward launch: ") + cudaGetErrorString(e)); return -2; }
  return 0;
}

extern "C" int gns_forward(const gns_plan* plan, const float* params, const float* buses, const float* lines,
                           const float* gens, int64_t S, int K, int L, int H, int multi, float gamma, float* v,
                           float* theta, float* total_loss, float* last_loss, void* workspace,
                           int64_t workspace_bytes, int need_grad, void* stream) {
  return forward_impl(plan, params, buses, lines, gens, nullptr, nullptr, S, K, L, H, multi, gamma, v, theta, total_loss,
                      last_loss, workspace, workspace_bytes, need_grad, stream);
}

extern "C" int gns_forward_compact(const gns_plan* plan, const float* params, const float* bus_var, const float* line_var,
                                   const float* gen_var, const float* bus_const, const float* gen_const, int64_t S, int K,
                                   int L, int H, int multi, float gamma, float* v, float* theta, float* total_loss,
                                   float* last_loss, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!bus_const || (plan && plan->Gn > 0 && !gen_const)) { set_error("gns_forward_compact: null constant block"); return -1; }
  return forward_impl(plan, params, bus_var, line_var, gen_var, bus_const, gen_const ? gen_const : bus_const, S, K, L, H, multi,
                      gamma, v, theta, total_loss, last_loss, workspace, workspace_bytes, 0, stream);
}

extern "C" int gns_backward(const gns_plan* cplan, const float* params, const float* buses, const float* lines,
                            const float* gens, int64_t S, int K, int L, int H, int multi, float gamma,
                            const float* grad_total, const float* grad_last, const float* grad_v,
                            const float* grad_theta, float* grad_params, void* workspace, int64_t workspace_bytes,
                            void* stream) {
  gns_plan* plan = const_cast<gns_plan*>(cplan);
  if (!plan || !params || !buses || !lines || !gens || !grad_total || !grad_params || !workspace) {
    set_error("gns_backward: null argument"); return -1;
  }
  ModelDims md{K, L, H, multi};
  if (!check_model(md) || !check_plan(plan)) return -1;
  if (cudaSetDevice(plan->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return -2; }
  return run_backward(plan, md, params, buses, lines, gens, S, gamma, grad_total, grad_last, grad_v, grad_theta,
                      grad_params, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int gns_check_topology(const gns_plan* plan, const float* lines, const float* gens, int64_t S, void* stream) {
  if (!plan || !lines || (plan->Gn > 0 && !gens) || S <= 0) { set_error("gns_check_topology: bad arguments"); return -1; }
  if (!check_plan(plan)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaSetDevice(plan->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return -2; }
  if (cudaMemsetAsync(plan->d_flag, 0, sizeof(int), st) != cudaSuccess) { set_error("memset flag failed"); return -2; }
  const long long total = S * (2LL * plan->E + plan->Gn);
  const int th = 256;
  const int bl = (int)std::min<long long>((total + th - 1) / th, 4096);
  check_topology_kernel<<<bl, th, 0, st>>>(lines, gens, plan->d_expect, S, plan->E, plan->Gn, plan->d_flag);
  int flag = 0;
  if (cudaMemcpyAsync(&flag, plan->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    set_error(std::string("gns_check_topology: ") + cudaGetErrorString(cudaGetLastError())); return -2;
  }
  return flag ? 1 : 0;
}

extern "C" int gns_expand_inputs(const float* bus_var, const float* line_var, const float* gen_var, const float* bus_const,
                                 const float* line_const, const float* gen_const, int64_t S, int n_bus, int n_line, int n_gen,
                                 float* buses, float* lines, float* gens, void* stream) {
  if (!bus_var || !line_var || !bus_const || !line_const || !buses || !lines || S <= 0 || n_bus <= 0 || n_line <= 0 ||
      n_gen < 0 || (n_gen > 0 && (!gen_var || !gen_const || !gens))) {
    set_error("gns_expand_inputs: bad arguments"); return -1;
  }
  const long long total = S * ((long long)n_bus + n_line + n_gen);
  const int th = 256;
  const int bl = (int)std::min<long long>((total + th - 1) / th, 148 * 16);
  expand_inputs_kernel<<<bl, th, 0, (cudaStream_t)stream>>>(bus_var, line_var, gen_var, bus_const, line_const, gen_const, S, n_bus,
                                                          n_line, n_gen, buses, lines, gens);
  if (cudaGetLastError() != cudaSuccess) { set_error("gns_expand_inputs: launch failed"); return -2; }
  return 0;
}

extern "C" int gns_check_topology_async(const gns_plan* plan, const float* lines, const float* gens, int64_t S, int* flag,
                                        void* stream) {
  if (!plan || !lines || (plan->Gn > 0 && !gens) || S <= 0 || !flag) { set_error("gns_check_topology_async: bad arguments"); return -1; }
  if (!check_plan(plan)) return -1;
  if (cudaSetDevice(plan->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return -2; }
  const long long total = S * (2LL * plan->E + plan->Gn);
  const int th = 256;
  const int bl = (int)std::min<long long>((total + th - 1) / th, 4096);
  check_topology_kernel<<<bl, th, 0, (cudaStream_t)stream>>>(lines, gens, plan->d_expect, S, plan->E, plan->Gn, flag);
  if (cudaGetLastError() != cudaSuccess) { set_error("gns_check_topology_async: launch failed"); return -2; }
  return 0;
}

extern "C" int gns_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, int64_t step, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) { set_error("gns_adam_step: bad arguments"); return -1; }
  const float bc1 = 1.f - (float)std::pow((double)beta1, (double)step);
  const float bc2 = 1.f - (float)std::pow((double)beta2, (double)step);
  const int th = 256;
  const int bl = (int)std::min<int64_t>((n + th - 1) / th, 2048);
  adam_kernel<<<bl, th, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1,
                                                   std::sqrt(bc2));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error(std::string("adam launch: ") + cudaGetErrorString(e)); return -2; }
  return 0;
}

extern "C" double gns_measure_ffma2_flops(int device, int iters) {
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return -1.0; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  float* d = nullptr;
  cudaMalloc(&d, 4);
  const int threads = 256, blocks = prop.multiProcessorCount * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  ffma2_probe_kernel<<<blocks, threads>>>(d, iters / 4 + 1, 1.0000001f, 1e-9f);
  cudaDeviceSynchronize();
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    ffma2_probe_kernel<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 2 * 8 * 8 * (double)iters * threads * (double)blocks;
    best = std::max(best, flops / (ms * 1e-3));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  if (cudaGetLastError() != cudaSuccess) { set_error("ffma2 probe failed"); return -1.0; }
  return best;
}

extern "C" double gns_measure_ffma_flops(int device, int iters) {
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return -1.0; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  float* d = nullptr;
  cudaMalloc(&d, 4);
  const int threads = 256, blocks = prop.multiProcessorCount * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  ffma_probe_kernel<<<blocks, threads>>>(d, iters / 4 + 1, 1.0000001f, 1e-9f);   // warm-up
  cudaDeviceSynchronize();
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    ffma_probe_kernel<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * 8 * (double)iters * threads * (double)blocks;
    best = std::max(best, flops / (ms * 1e-3));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  if (cudaGetLastError() != cudaSuccess) { set_error("ffma probe failed"); return -1.0; }
  return best;
}
