/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_dyn_impl.h for the full header).
 * Instantiates the CPU restatement of the dynamics in float64 (`_f64`, the checker) and float32
 * (`_f32`, same arithmetic width as the CUDA kernels, used to separate rounding from logic errors
 * and as the timed CPU baseline).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include "b200gym.h"

#define ORC_REAL double
#define ORC_SUF _f64
#include "oracle_dyn_impl.h"
#undef ORC_REAL
#undef ORC_SUF

#define ORC_REAL float
#define ORC_SUF _f32
#include "oracle_dyn_impl.h"
#undef ORC_REAL
#undef ORC_SUF
