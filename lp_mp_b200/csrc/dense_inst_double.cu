#define LPMP_INST_REAL double
#include "dense_inst.cuh"
