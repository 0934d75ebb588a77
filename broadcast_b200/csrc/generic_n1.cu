#define BCAST_N 1
#include "generic_impl.cuh"
