#include "bp_inst.cuh"
QLDPC_DEFINE_SHAPE(6, 2)
