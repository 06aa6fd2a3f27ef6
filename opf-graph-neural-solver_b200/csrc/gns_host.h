// gns_host.h — host-side plan object and launch geometry (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "gns_common.cuh"

struct gns_plan {
  int device = 0;
  int N = 0, E = 0, Gn = 0;
  int Ns = 0;                       // bus slots (high-degree buses own 2 or 4)
  int deg_cap = 2, max_gsz = 1;
  int max_walk = 0;                 // most in-lines walked by one slot
  int num_sms = 0;
  int smem_optin = 0;
  // host copies (int32) for export and tests
  std::vector<int32_t> f_bus, t_bus, gen_bus;
  std::vector<int32_t> in_rowptr, in_lines, out_rowptr, out_lines, gen_rowptr, gen_ids;
  std::vector<int32_t> bus_order, bus_rank;
  std::vector<int32_t> slot_bus, slot_primary, slot_in_begin, slot_in_end, slot_gsz;   // slot tables
  // device index block (uint16) + float copies of the expected index columns
  gns::TopoOffsets to{};
  uint16_t* d_topo = nullptr;
  float* d_expect = nullptr;        // [2E + Gn]: f_bus+1, t_bus+1, gen_bus+1 as float (topology check)
  int* d_flag = nullptr;
  // canonical(state_dict order) -> packed index maps, cached per (K,L,H,multi)
  struct PackMap { int32_t* d_map = nullptr; int64_t n_canon = 0; int64_t n_packed = 0; };
  std::map<std::tuple<int, int, int, int>, PackMap> pack_maps;
  // packed index -> fragment-order accumulator index (device), cached per (L,H,multi)
  std::map<std::tuple<int, int, int>, int32_t*> frag_maps;
};

namespace gns {

struct Geometry {
  int VG = 1, NGQ = 1, G = 1, T = 32;
  int tmax = 384;           // launch-bounds variant
  int nbatch = 0, ctas = 0, num_sms = 0;
  size_t smem_bytes = 0;
  SmemPlan sm{};
  unsigned char grp_of_warp[32] = {0};   // bus group handled by each warp (balances the 4 SM sub-partitions)
};

struct ModelDims { int K, L, H, multi; };

// launch geometry of the warp-specialised backward kernel (gns_backward2.cuh): one grid per CTA at a time
struct Bwd2Geom {
  bool ok = false;
  int variant = 2;          // 2: warp-specialised kernel (gns_backward2.cuh), 3: fragment-space kernel (gns_backward3.cuh)
  int PW = 0, CW = 0, T = 0, ctas = 0;
  int bt = 0, lt = 0;       // variant 3: bus / line tiles per warp
  int parts_per_cta = 0;    // accumulator blocks per CTA (the warps that own one; 1 when the CTA's warps share a block)
  size_t smem_bytes = 0;
  Act2Layout a2{};
};
// eligible iff the grid is large enough to fill the producer warps, the blocks fit shared memory and kernels are
// built for the dims; GNS_BWD2=0 / 1 forces the first / this kernel (A/B measurements, tests)
Bwd2Geom choose_bwd2(const gns_plan* plan, const ModelDims& md, long long S);

void set_error(const std::string& msg);

// workspace carving (all offsets in bytes, 256-byte aligned)
struct Workspace {
  size_t packed_params = 0;   // [K][wstep] floats
  size_t ckpt = 0;            // [nbatch][K][pad4((4+L) N G)] floats   (need_grad)
  size_t pglob = 0;           // [nbatch][K][G] floats                 (need_grad)
  size_t act = 0;             // [nbatch][K][ActLayout.total] floats   (need_grad): hidden activations
  size_t gpartial = 0;        // [ctas*nwarps][K][FragLayout.step] floats (need_grad): per-warp gradient partial sums
  size_t fragsum = 0;         // [K][FragLayout.step] floats           (need_grad): accumulators summed over warps
  size_t packed_grad = 0;     // [K][wstep] floats                     (need_grad)
  size_t mscratch = 0;        // [ctas][2][L][NGs] floats              (need_grad, L > 32)
  size_t total = 0;
};

bool choose_geometry(const gns_plan* plan, const ModelDims& md, long long S, bool backward, Geometry* out);
// activation checkpoints: grid-major rows iff the backward kernel works on one grid per CTA (see ActLayout);
// GNS_ACT_LAYOUT=grid / interleaved overrides (A/B measurements)
bool act_grid_major(const Geometry& bwd);
Workspace plan_workspace(const gns_plan* plan, const ModelDims& md, long long S, bool need_grad,
                         const Geometry& fwd, const Geometry& bwd, const Bwd2Geom& b2);

// canonical <-> packed parameter maps
int64_t canonical_param_count(const ModelDims& md);
std::vector<int32_t> build_pack_map(const ModelDims& md);   // canonical index -> packed index
const gns_plan::PackMap* get_pack_map(gns_plan* plan, const ModelDims& md);

// kernel launchers (defined in the per-dimension translation units)
typedef cudaError_t (*FwdLauncher)(const FwdArgs& a, const Geometry& g, cudaStream_t st);
FwdLauncher find_forward(int L, int H, int multi, int VG, int tmax);
struct BwdArgs;
typedef cudaError_t (*BwdLauncher)(const BwdArgs& a, const Geometry& g, cudaStream_t st);
BwdLauncher find_backward(int L, int H, int multi, int tmax);
struct Bwd2Args;
typedef cudaError_t (*Bwd2Launcher)(const Bwd2Args& a, const Bwd2Geom& g, int num_sms, cudaStream_t st);
Bwd2Launcher find_backward2(int L, int H, int multi);
struct Bwd3Args;
typedef cudaError_t (*Bwd3Launcher)(const Bwd3Args& a, const Bwd2Geom& g, int num_sms, cudaStream_t st);
Bwd3Launcher find_backward3(int L, int H, int multi);

}  // namespace gns
