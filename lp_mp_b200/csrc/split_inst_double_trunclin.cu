#define LPMP_INST_REAL double
#define LPMP_INST_PW LPMP_PAIRWISE_TRUNCLIN
#include "split_inst.cuh"
