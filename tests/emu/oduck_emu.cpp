// liboduck_emu: csrc/oduck_cuda.cu -- the env side of liboduck_cuda (oduck_create / randomize / reset / step / physics /
// buffers) -- compiled for the HOST against tests/emu/cuda_runtime.h (TEST INFRASTRUCTURE).  Every kernel launch runs block by
// block as 32 threads, one per lane, so the device code's logic is exercised against the oracle without a GPU.  Built with
// -DWPB=1 (one warp per block).  The tensor-core entry points (policy, learner) are not part of it.
#include <cuda_runtime.h>

#include "../../open_duck_playground_b200/csrc/oduck_cuda.cu"

extern "C" {
int oduck_policy_forward(OduckHandle*, const OduckPolicyWeights*, const float*, const uint32_t*, int, float*, float*, float*, void*) {
  return oduck_fail(ODUCK_ERR_UNSUPPORTED, "oduck_policy_forward: not part of the CPU emulation");
}
int oduck_policy_invalidate(OduckHandle*) { return ODUCK_OK; }
// the env half of oduck_rollout_step (oduck_step_into_sink: k_step writing the Transition into the attached sink) IS part of the
// emulation and is what tests/test_env_emu.py drives; the actor half is a tensor-core kernel
int oduck_rollout_step(OduckHandle*, const OduckPolicyWeights*, const uint32_t*, int, void*) {
  return oduck_fail(ODUCK_ERR_UNSUPPORTED, "oduck_rollout_step: the actor is not part of the CPU emulation (use oduck_step_into_sink)");
}
}
