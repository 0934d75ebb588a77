#define BCAST_N 5
#include "generic_impl.cuh"
