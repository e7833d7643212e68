#define ACE_STREAM_NF 4
#include "stream_template.cuh"
