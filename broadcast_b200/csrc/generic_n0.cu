#define BCAST_N 0
#include "generic_impl.cuh"
