#define LPMP_INST_REAL float
#define LPMP_INST_PW LPMP_PAIRWISE_POTTS
#include "split_inst.cuh"
