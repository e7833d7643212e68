#define ACE_INST_NMAX 12
#include "inst_template.cuh"
