/** @file partitioning.hxx  layout::flat_uniform_occupancy<K, base> lives in loops/container/layout.hxx
 *  (reference include/loops/container/partitioning.hxx:71-141). */
#pragma once
#include <loops/container/layout.hxx>
