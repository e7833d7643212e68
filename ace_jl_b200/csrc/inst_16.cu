#define ACE_INST_NMAX 16
#include "inst_template.cuh"
