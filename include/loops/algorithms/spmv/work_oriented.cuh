/** @file work_oriented.cuh  algorithms::spmv::work_oriented is declared in loops/algorithms/spmv/spmv.cuh
 *  (reference include/loops/algorithms/spmv/work_oriented.cuh). */
#pragma once
#include <loops/algorithms/spmv/spmv.cuh>
