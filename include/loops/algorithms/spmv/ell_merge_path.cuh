/** @file ell_merge_path.cuh  algorithms::spmv::ell_merge_path is declared in loops/algorithms/spmv/spmv.cuh
 *  (reference include/loops/algorithms/spmv/ell_merge_path.cuh). */
#pragma once
#include <loops/algorithms/spmv/spmv.cuh>
