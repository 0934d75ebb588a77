// order-3 member of the scheme family (flux_num_dnc3.F90), tangent width 1: see generic_impl.cuh
#define BCAST_N 1
#define BCAST_ORD 3
#include "generic_impl.cuh"
