/** @file original.cuh  algorithms::spmv::original is declared in loops/algorithms/spmv/spmv.cuh
 *  (reference include/loops/algorithms/spmv/original.cuh). */
#pragma once
#include <loops/algorithms/spmv/spmv.cuh>
