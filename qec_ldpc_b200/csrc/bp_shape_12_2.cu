#include "bp_inst.cuh"
QLDPC_DEFINE_SHAPE(12, 2)
