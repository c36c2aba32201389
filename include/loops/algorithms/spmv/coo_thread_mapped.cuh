/** @file coo_thread_mapped.cuh  algorithms::spmv::coo_thread_mapped is declared in loops/algorithms/spmv/spmv.cuh
 *  (reference include/loops/algorithms/spmv/coo_thread_mapped.cuh). */
#pragma once
#include <loops/algorithms/spmv/spmv.cuh>
