/** @file csc_thread_mapped.cuh  algorithms::spmv::csc_thread_mapped is declared in loops/algorithms/spmv/spmv.cuh
 *  (reference include/loops/algorithms/spmv/csc_thread_mapped.cuh). */
#pragma once
#include <loops/algorithms/spmv/spmv.cuh>
