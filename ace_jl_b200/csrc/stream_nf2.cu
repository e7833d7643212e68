#define ACE_STREAM_NF 2
#include "stream_template.cuh"
