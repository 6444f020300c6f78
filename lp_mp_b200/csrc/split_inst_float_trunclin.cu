#define LPMP_INST_REAL float
#define LPMP_INST_PW LPMP_PAIRWISE_TRUNCLIN
#include "split_inst.cuh"
