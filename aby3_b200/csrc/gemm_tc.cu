// gemm_tc.cu -- tcgen05 / TMEM u8-limb GEMM (placeholder until the kernel lands).
#include "common.cuh"
namespace aby3cu {
bool gemm_tc_profitable(u64, u64, u64) { return false; }
int gemm_cross_tc(aby3cu_ctx*, const i64*, const i64*, const i64*, const i64*, u64, u64, u64, i64*, int) {
    set_error("tcgen05 GEMM not built yet");
    return 4;
}
}  // namespace aby3cu
