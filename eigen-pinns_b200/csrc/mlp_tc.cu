// bf16 tcgen05 corrector MLP (perf mode) — placeholder entry points until the kernel lands.
#include "ep_common.cuh"
extern "C" {
size_t ep_mlp_tc_packed_weight_bytes(int, const int*) { return 0; }
size_t ep_mlp_tc_packed_input_bytes(int, int) { return 0; }
size_t ep_mlp_tc_act_bytes(int, int, const int*) { return 0; }
size_t ep_mlp_tc_bwd_workspace_bytes(int, int, const int*) { return 0; }
int ep_mlp_tc_pack_weights(int, const int*, const float* const*, void*, void*, ep_stream_t) {
  ep::set_error("ep_mlp_tc_*: not built"); return EP_ERR_UNSUPPORTED; }
int ep_mlp_tc_pack_input(int, int, const float*, int, void*, ep_stream_t) {
  ep::set_error("ep_mlp_tc_*: not built"); return EP_ERR_UNSUPPORTED; }
int ep_mlp_tc_fwd(int, int, const int*, const void*, const void*, const float* const*, void*, const float*, float,
                  const float*, float*, float*, int, ep_stream_t) {
  ep::set_error("ep_mlp_tc_*: not built"); return EP_ERR_UNSUPPORTED; }
int ep_mlp_tc_bwd(int, int, const int*, const void*, const void*, const void*, const void*, const float*, int, float,
                  const float*, float* const*, float* const*, void*, size_t, ep_stream_t) {
  ep::set_error("ep_mlp_tc_*: not built"); return EP_ERR_UNSUPPORTED; }
}
