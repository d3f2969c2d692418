// TEST INFRASTRUCTURE: csrc/layout_boxes.cu compiled for the host.
#define SG2IM_EMUL 1
#include "../../sg2im_b200/csrc/layout_boxes.cu"
