/*
 * gns_b200.h — C ABI of the B200-native Graph Neural Solver hot path.
 *
 * The reference (LeonOrou/OPF-Graph-Neural-Solver) has no FFI or plugin boundary; its
 * boundary for this path is the Python call
 *
 *     GNS.forward(buses, lines, generators, B, L, G) -> (v, theta, total_loss, last_loss)
 *                                                        (ref GNS/main.py:140-202)
 *
 * and the implicit autograd backward of `mean(total_loss)` (ref GNS/main.py:284-288).
 * This header is what a ctypes / cffi binding for that call binds (INTEGRATION.md
 * shows the stub).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every `const float*` / `float*` below that is not marked (host) is a DEVICE
 *     pointer to contiguous float32 on the plan's device;
 *   - the library never allocates or frees caller memory; scratch comes from the
 *     caller-provided workspace (size from gns_workspace_bytes);
 *   - all launches go to the `stream` argument (a cudaStream_t passed as void*);
 *     nothing synchronises unless stated;
 *   - return value 0 = ok, non-zero = error (text via gns_last_error(), thread local);
 *   - batched tensors use the reference's packed row layout (ref GNS/utils.py:5-9):
 *       buses [S][n_bus][6]   bus_i,type,Pd,Qd,Gs,Bs
 *       lines [S][n_line][7]  f_bus,t_bus,r,x,b,tau,shift
 *       gens  [S][n_gen][7]   bus_i,Pmax,Pmin,Pg_set,vg,qg,Pg
 *     with ONE topology (f_bus,t_bus,gen bus_i) shared by the whole batch;
 *   - parameters are one flat float32 buffer in the reference's state_dict order
 *     (ref GNS/main.py:113-134): for net in [phi_v,phi_theta,phi_m |
 *     phi], L_theta, L_v, L_m: for k in 0..K-1: linear1.weight[H][in], linear1.bias[H],
 *     linear2.weight[H][H], linear2.bias[H], linear4.weight[out][H], linear4.bias[out].
 */
#ifndef GNS_B200_H
#define GNS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gns_plan gns_plan;

/* Topology plan: CSR of lines by receiving bus and by sending bus, generators by bus,
 * the degree-sorted internal bus order, alias-gather indices.  Replaces the per-call
 * index tensors the reference rebuilds on every forward (ref GNS/main.py:35-36,85-86,
 * 144,153,184-185).  f_bus/t_bus/gen_bus are HOST arrays of 0-based bus numbers.
 * Preconditions (checked): 0 <= bus < n_bus; n_bus <= n_line (the reference indexes
 * per-line vectors with bus numbers, ref GNS/main.py:41,68-72,91-92). */
int gns_plan_create(int n_bus, int n_line, int n_gen,
                    const int32_t* f_bus, const int32_t* t_bus, const int32_t* gen_bus,
                    int device, gns_plan** out_plan);
void gns_plan_destroy(gns_plan* plan);

/* Copy one of the plan's host-side index arrays into `out` (capacity in elements).
 * Returns the element count, or -1.  Names: "in_rowptr" [n_bus+1], "in_lines" [n_line]
 * (lines grouped by t_bus, ascending line id inside a bus), "out_rowptr", "out_lines"
 * (by f_bus), "gen_rowptr" [n_bus+1], "gen_ids" [n_gen], "bus_order" [n_bus] (internal
 * slot -> bus, in-degree descending, stable), "bus_rank" [n_bus] (bus -> slot). */
int gns_plan_export(const gns_plan* plan, const char* name, int32_t* out, int capacity);

/* 1 if kernels are instantiated for this (latent_dim, hidden_dim). */
int gns_dims_supported(int latent_dim, int hidden_dim);

/* Number of float32 parameters in the state_dict-order flat buffer. */
int64_t gns_param_count(int K, int latent_dim, int hidden_dim, int multiple_phi);

/* Bytes of device scratch for a batch of S grids.  need_grad != 0 adds the per-step
 * checkpoints (state and hidden activations of every step, about 1 MB per case300
 * grid for K=4) and the gradient partial sums that gns_backward consumes. */
int64_t gns_workspace_bytes(const gns_plan* plan, int64_t S, int K, int latent_dim,
                            int hidden_dim, int multiple_phi, int need_grad);

/* K-step forward for S grids (ref GNS/main.py:140-202).
 * Outputs: v [S][n_bus] (clamped at 0 like ref :201), theta [S][n_bus],
 * total_loss [S], last_loss [S].  If need_grad != 0 the workspace keeps what
 * gns_backward needs and must stay untouched until it ran. */
int gns_forward(const gns_plan* plan, const float* params,
                const float* buses, const float* lines, const float* gens,
                int64_t S, int K, int latent_dim, int hidden_dim, int multiple_phi,
                float gamma,
                float* v, float* theta, float* total_loss, float* last_loss,
                void* workspace, int64_t workspace_bytes, int need_grad, void* stream);

/* Inference forward on the COMPACT input format (see gns_expand_inputs for the columns): the kernel reads
 * bus_var [S][n_bus][2], line_var [S][n_line][5], gen_var [S][n_gen][2] directly and takes the constant columns from
 * bus_const [n_bus][4] / gen_const [n_gen][4]; the topology is the plan's (line_const is not needed).  Same outputs,
 * bit for bit, as gns_forward on the expanded rows; no checkpoints (need_grad = 0).  Replaces ref GNS/utils.py:17-41
 * + GNS/main.py:140-202 for host batches of one case: 11.2 KB instead of 20.6 KB per case300 grid cross PCIe. */
int gns_forward_compact(const gns_plan* plan, const float* params,
                        const float* bus_var, const float* line_var, const float* gen_var,
                        const float* bus_const, const float* gen_const,
                        int64_t S, int K, int latent_dim, int hidden_dim, int multiple_phi, float gamma,
                        float* v, float* theta, float* total_loss, float* last_loss,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* Backward of  sum_s ( grad_total[s]*total_loss[s] + grad_last[s]*last_loss[s]
 *                      + <grad_v[s], v[s]> + <grad_theta[s], theta[s]> )
 * w.r.t. the parameters (the implicit autograd backward of the reference,
 * ref GNS/main.py:288).  grad_last / grad_v / grad_theta may be NULL (= zeros).
 * grad_params (state_dict order, gns_param_count floats) is overwritten. */
int gns_backward(const gns_plan* plan, const float* params,
                 const float* buses, const float* lines, const float* gens,
                 int64_t S, int K, int latent_dim, int hidden_dim, int multiple_phi,
                 float gamma,
                 const float* grad_total, const float* grad_last,
                 const float* grad_v, const float* grad_theta,
                 float* grad_params,
                 void* workspace, int64_t workspace_bytes, void* stream);

/* Verify on the device that every grid of the batch carries the plan's topology
 * (the reference would silently use per-sample indices, ref GNS/main.py:153).
 * Synchronises `stream`.  Returns 0 if all match, 1 if a mismatch was found, <0 on error. */
int gns_check_topology(const gns_plan* plan, const float* lines, const float* gens,
                       int64_t S, void* stream);

/* Same check without synchronisation: ORs 1 into the caller's device int `flag` (zeroed by the caller) when a grid
 * of the batch carries another topology; the caller reads the flag when it next synchronises (pipelined host
 * batches check every chunk this way).  Replaces the per-sample index rebuild of ref GNS/main.py:153. */
int gns_check_topology_async(const gns_plan* plan, const float* lines, const float* gens,
                             int64_t S, int* flag, void* stream);

/* Compact input format -> the reference's packed rows, on the device (replaces the host-side row build of
 * ref GNS/utils.py:17-41 for batches of ONE case).  Between samples only Pd,Qd | r,x,b,tau,shift | vg,Pg vary
 * (ref GNS/augment_grids.py:35-53); the rest is constant per case and Pg_set is a copy of Pg (ref GNS/utils.py:38).
 *   bus_var [S][n_bus][2] = Pd,Qd      line_var [S][n_line][5] = r,x,b,tau,shift    gen_var [S][n_gen][2] = vg,Pg
 *   bus_const [n_bus][4] = bus_i,type,Gs,Bs   line_const [n_line][2] = f_bus,t_bus  gen_const [n_gen][4] = bus_i,Pmax,Pmin,qg
 * writes buses [S][n_bus][6], lines [S][n_line][7], gens [S][n_gen][7] (all device pointers). */
int gns_expand_inputs(const float* bus_var, const float* line_var, const float* gen_var,
                      const float* bus_const, const float* line_const, const float* gen_const,
                      int64_t S, int n_bus, int n_line, int n_gen,
                      float* buses, float* lines, float* gens, void* stream);

/* Layout maps of the gradient path, for tests and tools (host only, no GPU needed).
 *   "pack": canonical (state_dict order, all K steps) index -> packed index        [gns_param_count]
 *   "frag": packed index inside ONE step's block -> index inside that step's fragment-order
 *           accumulator block of gns_backward, or -1 (padding, and the W4 / b4 / W1-slice entries the
 *           fused block's chain rule fills in)                                      [packed step size]
 *   "frag2": the same map for the warp-specialised backward kernel (large grids)    [packed step size]
 * Returns the number of entries (writes them when out != NULL and capacity suffices), -1 on error.
 * Replaces nothing in the reference: autograd keeps its own bookkeeping (ref GNS/main.py:288). */
int gns_layout_export(const char* name, int K, int latent_dim, int hidden_dim, int multiple_phi,
                      int32_t* out, int capacity);

/* Launch geometry chosen for this plan/model (for bench/roofline reporting; nothing in the reference corresponds):
 * out[0]=grids per CTA, out[1]=threads per CTA, out[2]=dynamic smem bytes, out[3]=CTAs launched for S grids
 * (persistent kernels: min(batches, SMs x resident CTAs)), out[4]=items per thread (grids, or bus slots in the
 * warp-specialised backward kernel) or bus items per warp (fragment-space backward kernel), out[5]=CTA batches,
 * out[6]=SMs, out[7]=launch-bounds variant. */
int gns_launch_info(const gns_plan* plan, int64_t S, int K, int latent_dim, int hidden_dim,
                    int multiple_phi, int backward, int32_t out[8]);

/* Adam step on flat buffers (ref GNS/main.py:243,289: torch.optim.Adam defaults,
 * no weight decay, no amsgrad).  step is the 1-based step count. */
int gns_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  int64_t n, float lr, float beta1, float beta2, float eps, int64_t step,
                  void* stream);

/* FP32 FMA throughput probe used for the roofline denominator: runs a register-only
 * FFMA kernel for about `iters` dependent-chain rounds and returns achieved FLOP/s
 * (synchronises the device). */
double gns_measure_ffma_flops(int device, int iters);
/* Same probe issued as packed FFMA2 (fma.rn.f32x2, two FMAs per lane per instruction). */
double gns_measure_ffma2_flops(int device, int iters);

const char* gns_last_error(void);
const char* gns_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GNS_B200_H */
