#define LPMP_INST_REAL double
#define LPMP_INST_PW LPMP_PAIRWISE_TRUNCLIN
#define LPMP_INST_ANISO false
#include "wavefront_inst.cuh"
