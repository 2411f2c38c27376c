// net_umma.cu — tcgen05 engine (placeholder until the kernels land; SIMT mode is complete).
#include "net.cuh"
#include "net_umma.cuh"

namespace b200 {
int umma_net_init(b200dqn_net* n) {
  B2_REQUIRE(n->cfg.math_mode != B200DQN_MATH_TCGEN05, B200DQN_ENOTIMPL, "math_mode TCGEN05 is not built yet");
  return B200DQN_OK;
}
void umma_net_destroy(b200dqn_net*) {}
int umma_weights_changed(b200dqn_net*, cudaStream_t) { return B200DQN_OK; }
int umma_target_synced(b200dqn_net*, cudaStream_t) { return B200DQN_OK; }
int umma_forward(b200dqn_net*, const uint8_t* const*, const int32_t* const*, const int*, int, int, cudaStream_t) {
  set_error("tcgen05 forward not built");
  return B200DQN_ENOTIMPL;
}
bool umma_has_backward() { return false; }
int umma_backward(b200dqn_net*, const uint8_t*, const int32_t*, int, int, cudaStream_t) { return B200DQN_ENOTIMPL; }
int umma_forward_launches() { return 4; }
int umma_backward_launches() { return 7; }
}  // namespace b200
