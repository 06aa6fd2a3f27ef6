// gns_backward.cu — host side of the backward pass: geometry, launch, and the two small
// kernels that fold the per-warp gradient accumulators into the state_dict-order gradient.
#include <algorithm>
#include <cstring>

#include "gns_backward.cuh"
#include "gns_host.h"

namespace gns {

BwdLauncher find_backward(int L, int H, int multi, int tmax);

int backward_extra_floats(int N, int E, int G, int L, int H, int T) {
  // PO <= L: size the tiles for the multiple-phi case
  return make_bwd_smem(N, E, G, L, H, L, T / 32, L > 32).total;
}

int backward_ctas(const gns_plan* plan, const ModelDims&, const Geometry& g) {
  // one persistent CTA per SM slot; the exact occupancy is clamped again at launch
  return std::min(g.nbatch, plan->num_sms * std::max(1, (int)(plan->smem_optin / std::max<size_t>(g.smem_bytes, 1))));
}

// packed_grad[p] = sum_w gacc[w][p]
__global__ void reduce_partials_kernel(const float* __restrict__ gacc, float* __restrict__ packed, long long n,
                                       int nparts) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int w = 0;
    for (; w + 3 < nparts; w += 4) {
      s0 += gacc[(size_t)w * n + p];
      s1 += gacc[(size_t)(w + 1) * n + p];
      s2 += gacc[(size_t)(w + 2) * n + p];
      s3 += gacc[(size_t)(w + 3) * n + p];
    }
    for (; w < nparts; ++w) s0 += gacc[(size_t)w * n + p];
    packed[p] = (s0 + s1) + (s2 + s3);
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

int run_backward(gns_plan* plan, const ModelDims& md, const float* params, const float* buses, const float* lines,
                 const float* gens, long long S, float gamma, const float* grad_total, const float* grad_last,
                 const float* grad_v, const float* grad_theta, float* grad_params, void* workspace,
                 long long workspace_bytes, cudaStream_t st) {
  (void)params;
  Geometry gf, gb;
  if (!choose_geometry(plan, md, S, false, &gf)) return -1;
  if (!choose_geometry(plan, md, S, true, &gb)) return -1;
  gb.ctas = backward_ctas(plan, md, gb);
  const Workspace ws = plan_workspace(plan, md, S, true, gf, gb);
  if ((long long)ws.total > workspace_bytes) { set_error("gns_backward: workspace too small"); return -1; }
  const gns_plan::PackMap* pm = get_pack_map(plan, md);
  if (!pm) return -2;
  const WLayout W = make_wlayout(md.L, md.H, md.multi != 0);
  char* wsb = static_cast<char*>(workspace);
  const int nwarps = gb.T / 32;
  const size_t per_part = (size_t)md.K * W.wstep;
  const int nparts = gb.ctas * nwarps;
  float* gacc = reinterpret_cast<float*>(wsb + ws.gpartial);
  float* packed_grad = reinterpret_cast<float*>(wsb + ws.packed_grad);
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
  a.grad_total = grad_total; a.grad_last = grad_last; a.grad_v = grad_v; a.grad_theta = grad_theta;
  a.gacc = gacc;
  a.topo = plan->d_topo;
  a.S = S; a.N = plan->N; a.Ns = plan->Ns; a.E = plan->E; a.Gn = plan->Gn; a.K = md.K; a.NGQ = gb.NGQ; a.G = gb.G; a.nbatch = gb.nbatch;
  a.NGs = row_stride(plan->Ns * gb.G); a.EGs = row_stride(plan->E * gb.G);
  a.Gf = gf.G; a.NGs_f = row_stride(plan->Ns * gf.G);
  std::memcpy(a.grp_of_warp, gb.grp_of_warp, 32);
  a.sm = gb.sm;
  a.bs = make_bwd_smem(plan->Ns, plan->E, gb.G, md.L, md.H, md.L, nwarps, md.L > 32);
  a.mscratch = md.L > 32 ? reinterpret_cast<float*>(wsb + ws.mscratch) : nullptr;
  a.to = plan->to;
  for (int k = 0; k < md.K; ++k) a.wk[k] = (float)std::pow((double)gamma, (double)(md.K - k));
  e = launch(a, gb, st);
  if (e != cudaSuccess) { set_error(std::string("backward launch: ") + cudaGetErrorString(e)); return -2; }
  {
    const int th = 256;
    const int bl = (int)std::min<long long>(((long long)per_part + th - 1) / th, 2048);
    reduce_partials_kernel<<<bl, th, 0, st>>>(gacc, packed_grad, (long long)per_part, nparts);
    const int nphi = md.multi ? 3 : 1, PO = md.multi ? md.L : 1;
    const int tot = md.K * (nphi * PO * md.H + nphi * PO + 3 * PO * md.H);
    unfuse_grads_kernel<<<(tot + th - 1) / th, th, 0, st>>>(packed_grad, a.params, W, md.K);
    const int bl2 = (int)std::min<long long>((pm->n_canon + th - 1) / th, 1024);
    unpack_grads_kernel<<<bl2, th, 0, st>>>(packed_grad, pm->d_map, grad_params, pm->n_canon);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error(std::string("gradient reduce: ") + cudaGetErrorString(e)); return -2; }
  return 0;
}

}  // namespace gns
