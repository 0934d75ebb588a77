// order-7 member of the scheme family (flux_num_dnc7.F90), tangent width 0: see generic_impl.cuh
#define BCAST_N 0
#define BCAST_ORD 7
#include "generic_impl.cuh"
