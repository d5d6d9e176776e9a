#include "bp_inst.cuh"
QLDPC_DEFINE_SHAPE(4, 2)
