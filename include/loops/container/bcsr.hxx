/** @file bcsr.hxx  bcsr_t lives in loops/container/formats.hxx (reference include/loops/container/bcsr.hxx). */
#pragma once
#include <loops/container/formats.hxx>
