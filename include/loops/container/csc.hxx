/** @file csc.hxx  csc_t lives in loops/container/formats.hxx (reference include/loops/container/csc.hxx). */
#pragma once
#include <loops/container/formats.hxx>
