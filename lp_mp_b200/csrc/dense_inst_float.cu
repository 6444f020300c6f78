#define LPMP_INST_REAL float
#include "dense_inst.cuh"
