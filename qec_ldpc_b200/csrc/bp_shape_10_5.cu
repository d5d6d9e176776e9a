#include "bp_inst.cuh"
QLDPC_DEFINE_SHAPE_M(10, 5, 305)
