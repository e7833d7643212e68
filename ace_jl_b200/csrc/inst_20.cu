#define ACE_INST_NMAX 20
#include "inst_template.cuh"
