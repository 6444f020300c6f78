#define LPMP_INST_REAL double
#define LPMP_INST_PW LPMP_PAIRWISE_POTTS
#include "split_inst.cuh"
