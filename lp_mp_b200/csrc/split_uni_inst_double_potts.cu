#define LPMP_INST_REAL double
#define LPMP_INST_PW LPMP_PAIRWISE_POTTS
#define LPMP_INST_UNI 1
#include "split_inst.cuh"
