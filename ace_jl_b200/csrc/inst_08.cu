#define ACE_INST_NMAX 8
#include "inst_template.cuh"
