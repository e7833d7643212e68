#define ACE_STREAM_NF 3
#include "stream_template.cuh"
