#include "bp_inst.cuh"
QLDPC_DEFINE_SHAPE(8, 2)
