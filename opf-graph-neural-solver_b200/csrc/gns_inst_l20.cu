// kernels for latent_dim = 20, hidden_dim = 10
#include <algorithm>
#include "gns_inst.cuh"
namespace gns {
FwdLauncher find_forward_l20(int multi, int VG, int tmax) { return pick_forward<20, 10>(multi, VG, tmax); }
BwdLauncher find_backward_l20(int multi, int tmax) { return pick_backward<20, 10>(multi, tmax); }
Bwd2Launcher find_backward2_l20(int multi) { return pick_backward2<20, 10>(multi); }
Bwd3Launcher find_backward3_l20(int multi) { return pick_backward3<20, 10>(multi); }
}  // namespace gns
