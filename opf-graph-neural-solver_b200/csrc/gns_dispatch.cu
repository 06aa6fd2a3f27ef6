// gns_dispatch.cu — (latent_dim, hidden_dim) -> instantiated kernel launcher.
#include "gns_host.h"
namespace gns {
FwdLauncher find_forward_l10(int multi, int VG, int tmax);
FwdLauncher find_forward_l20(int multi, int VG, int tmax);
FwdLauncher find_forward_l64(int multi, int VG, int tmax);

FwdLauncher find_forward(int L, int H, int multi, int VG, int tmax) {
  if (H != 10) return nullptr;
  switch (L) {
    case 10: return find_forward_l10(multi, VG, tmax);
    case 20: return find_forward_l20(multi, VG, tmax);
    case 64: return find_forward_l64(multi, VG, tmax);
  }
  return nullptr;
}
BwdLauncher find_backward_l10(int multi, int tmax);
BwdLauncher find_backward_l20(int multi, int tmax);
BwdLauncher find_backward_l64(int multi, int tmax);

BwdLauncher find_backward(int L, int H, int multi, int tmax) {
  if (H != 10) return nullptr;
  switch (L) {
    case 10: return find_backward_l10(multi, tmax);
    case 20: return find_backward_l20(multi, tmax);
    case 64: return find_backward_l64(multi, tmax);
  }
  return nullptr;
}
Bwd2Launcher find_backward2_l10(int multi);
Bwd2Launcher find_backward2_l20(int multi);
Bwd2Launcher find_backward2_l64(int multi);
Bwd2Launcher find_backward2(int L, int H, int multi) {
  if (H != 10) return nullptr;
  switch (L) {
    case 10: return find_backward2_l10(multi);
    case 20: return find_backward2_l20(multi);
    case 64: return find_backward2_l64(multi);
  }
  return nullptr;
}
Bwd3Launcher find_backward3_l10(int multi);
Bwd3Launcher find_backward3_l20(int multi);
Bwd3Launcher find_backward3_l64(int multi);
Bwd3Launcher find_backward3(int L, int H, int multi) {
  if (H != 10) return nullptr;
  switch (L) {
    case 10: return find_backward3_l10(multi);
    case 20: return find_backward3_l20(multi);
    case 64: return find_backward3_l64(multi);
  }
  return nullptr;
}
}  // namespace gns
