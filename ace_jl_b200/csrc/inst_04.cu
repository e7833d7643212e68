#define ACE_INST_NMAX 4
#include "inst_template.cuh"
