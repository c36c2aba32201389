/** @file group_mapped.cuh  algorithms::spmv::group_mapped is declared in loops/algorithms/spmv/spmv.cuh
 *  (reference include/loops/algorithms/spmv/group_mapped.cuh). */
#pragma once
#include <loops/algorithms/spmv/spmv.cuh>
