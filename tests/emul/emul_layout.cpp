// TEST INFRASTRUCTURE: the layout kernels' own source, compiled for the host.
#define SG2IM_EMUL 1
#include "../../sg2im_b200/csrc/layout.cu"
