#define ACE_INST_NMAX 24
#include "inst_template.cuh"
