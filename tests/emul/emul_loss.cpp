// TEST INFRASTRUCTURE: csrc/loss.cu compiled for the host (SIMT emulation, -DSG2IM_EMUL)
#include "../../sg2im_b200/csrc/loss.cu"
