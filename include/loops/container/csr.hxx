/** @file csr.hxx  csr_t lives in loops/container/formats.hxx (reference include/loops/container/csr.hxx). */
#pragma once
#include <loops/container/formats.hxx>
