#define LPMP_INST_REAL float
#define LPMP_INST_PW LPMP_PAIRWISE_POTTS
#define LPMP_INST_ANISO false
#include "wavefront_inst.cuh"
