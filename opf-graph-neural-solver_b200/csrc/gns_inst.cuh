// gns_inst.cuh — explicit instantiation of the kernels for one latent_dim (GNS_INST_L),
// hidden_dim 10.  Included by gns_inst_l*.cu so the variants compile in parallel.
#pragma once
#include "gns_forward.cuh"
#include "gns_backward.cuh"
#include "gns_backward2.cuh"
#include "gns_backward3.cuh"
#include "gns_host.h"

namespace gns {

template <int L, int H, bool MULTI, int VG, int TMAX, int GRADV>
static cudaError_t launch_forward_g(const FwdArgs& a, const Geometry& g, cudaStream_t st) {
  auto kern = gns_forward_kernel<L, H, MULTI, VG, TMAX, GRADV>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, g.T, g.smem_bytes);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  const int ctas = std::min(g.nbatch, occ * g.num_sms);
  kern<<<ctas, g.T, g.smem_bytes, st>>>(a);
  return cudaGetLastError();
}

// the training variants also write the state and activation checkpoints (grid-major or interleaved rows)
template <int L, int H, bool MULTI, int VG, int TMAX>
static cudaError_t launch_forward(const FwdArgs& a, const Geometry& g, cudaStream_t st) {
  if (!a.need_grad) return launch_forward_g<L, H, MULTI, VG, TMAX, 0>(a, g, st);
  if (a.need_grad == 2) {
    if constexpr (L <= 20 && TMAX != 1024) return launch_forward_g<L, H, MULTI, VG, TMAX, 3>(a, g, st);
    else return cudaErrorInvalidValue;
  }
  return a.al.gs == 1 ? launch_forward_g<L, H, MULTI, VG, TMAX, 2>(a, g, st)
                      : launch_forward_g<L, H, MULTI, VG, TMAX, 1>(a, g, st);
}

template <int L, int H>
static FwdLauncher pick_forward(int multi, int VG, int tmax) {
  if (tmax == 320) {
    if (VG == 2) return multi ? launch_forward<L, H, true, 2, 320> : launch_forward<L, H, false, 2, 320>;
    if (VG == 1) return multi ? launch_forward<L, H, true, 1, 320> : launch_forward<L, H, false, 1, 320>;
  } else if (tmax == 384) {
    if (VG == 2) return multi ? launch_forward<L, H, true, 2, 384> : launch_forward<L, H, false, 2, 384>;
    if (VG == 1) return multi ? launch_forward<L, H, true, 1, 384> : launch_forward<L, H, false, 1, 384>;
  } else if (tmax == 1024 && VG == 1) {
    return multi ? launch_forward<L, H, true, 1, 1024> : launch_forward<L, H, false, 1, 1024>;
  }
  return nullptr;
}


template <int L, int H, bool MULTI, int TMAX>
static cudaError_t launch_backward(const BwdArgs& a, const Geometry& g, cudaStream_t st) {
  auto kern = gns_backward_kernel<L, H, MULTI, TMAX>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, g.T, g.smem_bytes);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  // g.ctas accumulator blocks were provisioned by the host; never launch more than that
  const int ctas = std::min(std::min(g.nbatch, occ * g.num_sms), g.ctas);
  kern<<<ctas, g.T, g.smem_bytes, st>>>(a);
  return cudaGetLastError();
}

template <int L, int H, bool MULTI>
static cudaError_t launch_backward2(const Bwd2Args& a, const Bwd2Geom& g, int num_sms, cudaStream_t st) {
  auto kern = gns_backward2_kernel<L, H, MULTI>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, g.T, g.smem_bytes);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  // g.ctas x CW accumulator blocks were provisioned by the host; never launch more CTAs than that
  const int ctas = (int)std::min<long long>(std::min<long long>(a.S, (long long)occ * num_sms), g.ctas);
  kern<<<ctas, g.T, g.smem_bytes, st>>>(a);
  return cudaGetLastError();
}
template <int L, int H>
static Bwd2Launcher pick_backward2(int multi) {
  if constexpr (L <= 20) return multi ? launch_backward2<L, H, true> : launch_backward2<L, H, false>;
  else return nullptr;
}

template <int L, int H, int BT, int LT, int TB, int MINB>
static cudaError_t launch_backward3_g(const Bwd3Args& a, const Bwd2Geom& g, int num_sms, cudaStream_t st) {
  auto kern = gns_backward3_kernel<L, H, BT, LT, TB, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, g.T, g.smem_bytes);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  const int ctas = (int)std::min<long long>(std::min<long long>(a.S, (long long)occ * num_sms), g.ctas);
  kern<<<ctas, g.T, g.smem_bytes, st>>>(a);
  return cudaGetLastError();
}
// geometry variants: (1, 2) tiles per warp with up to 640 threads of 96 registers, or up to 256 threads of 128 registers
// and two CTAs per SM; (2, 3) tiles per warp with up to 320 threads of 168 registers
template <int L, int H>
static cudaError_t launch_backward3(const Bwd3Args& a, const Bwd2Geom& g, int num_sms, cudaStream_t st) {
  if (g.bt == 1 && g.lt == 2) {
    if (g.T <= 256) return launch_backward3_g<L, H, 1, 2, 256, 2>(a, g, num_sms, st);
    if (g.T <= 640) return launch_backward3_g<L, H, 1, 2, 640, 1>(a, g, num_sms, st);
  } else if (g.bt == 2 && g.lt == 3 && g.T <= 320) {
    return launch_backward3_g<L, H, 2, 3, 320, 1>(a, g, num_sms, st);
  }
  return cudaErrorInvalidValue;
}
template <int L, int H>
static Bwd3Launcher pick_backward3(int multi) {
  if constexpr (L <= 20) return multi ? launch_backward3<L, H> : nullptr;
  else return nullptr;
}

template <int L, int H>
static BwdLauncher pick_backward(int multi, int tmax) {
  if (tmax == 320) return multi ? launch_backward<L, H, true, 320> : launch_backward<L, H, false, 320>;
  if (tmax == 384) return multi ? launch_backward<L, H, true, 384> : launch_backward<L, H, false, 384>;
  if (tmax == 1024) return multi ? launch_backward<L, H, true, 1024> : launch_backward<L, H, false, 1024>;
  return nullptr;
}

}  // namespace gns
