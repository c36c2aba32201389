/** @file merge_path_flat.cuh  algorithms::spmv::merge_path_flat is declared in loops/algorithms/spmv/spmv.cuh
 *  (reference include/loops/algorithms/spmv/merge_path_flat.cuh). */
#pragma once
#include <loops/algorithms/spmv/spmv.cuh>
