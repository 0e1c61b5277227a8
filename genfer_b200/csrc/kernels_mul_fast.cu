// Register-tiled DFMA product kernel for dense cube-like operands (placeholder: not yet enabled).
#include "kernels.cuh"

namespace gtp {

bool fast_mul_applicable(const Ctx&, const MulArgs&) { return false; }
void launch_mul_fast(Ctx&, const MulArgs&) { throw Error(GTP_ERR_ARG, "fast product kernel not built"); }

}  // namespace gtp
