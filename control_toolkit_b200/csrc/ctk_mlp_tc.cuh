// ctk_mlp_tc.cuh -- tcgen05 (UMMA + TMEM) engine for the MLP predictor.  Placeholder interface; the kernel lands in
// a later commit.  Until then selecting CTK_MLP_TCGEN05 fails loudly (no silent fallback to the SIMT engine).
#pragma once
#include <string>

#include "../../include/ctk_b200.h"
#include "ctk_args.cuh"

namespace ctk {

struct MlpTcDev {
  void* blob = nullptr;
};

inline void mlp_tc_free(MlpTcDev&) {}
inline bool mlp_tc_upload(MlpTcDev&, const ctk_mlp_weights*, std::string& err) {
  err = "tcgen05 MLP engine not built";
  return false;
}
inline bool mlp_tc_launch_mppi(MlpTcDev&, const MppiArgs&, int, bool, cudaStream_t, int*, int64_t*, std::string& err) {
  err = "tcgen05 MLP engine not built";
  return false;
}

}  // namespace ctk
