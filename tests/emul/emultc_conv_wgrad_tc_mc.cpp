// TEST INFRASTRUCTURE: csrc/conv_wgrad_tc_mc.cu compiled for the host against the functional tensor-core model
// (tests/emul/tc_emul.h).  Fiber execution model only.
#define SG2IM_EMUL 1
#include "../../sg2im_b200/csrc/conv_wgrad_tc_mc.cu"
