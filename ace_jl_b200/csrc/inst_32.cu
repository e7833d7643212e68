#define ACE_INST_NMAX 32
#include "inst_template.cuh"
