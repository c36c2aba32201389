/** @file flat_partitioned.cuh  algorithms::spmv::flat_partitioned is declared in loops/algorithms/spmv/spmv.cuh
 *  (reference include/loops/algorithms/spmv/flat_partitioned.cuh). */
#pragma once
#include <loops/algorithms/spmv/spmv.cuh>
