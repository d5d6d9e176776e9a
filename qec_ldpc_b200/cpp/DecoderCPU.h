// Source compatibility only: the reference's driver constructs `DecoderCPU decoder(code)` (QEC_LDPC/main.cu:79).  This
// framework has NO CPU decode path (north star: "no CPU fallback"), so the name is an alias of DecoderGPU and code
// written against DecoderCPU runs on the sm_100a kernels -- and fails loudly (std::string exception carrying
// QLDPC_ERR_NO_DEVICE's text) on a machine without a CUDA device.  The CPU restatement of the reference used for
// parity checking lives under oracle/ and is test infrastructure, not part of this package.
#pragma once
#include "DecoderGPU.h"
typedef DecoderGPU DecoderCPU;
