/** @file coo.hxx  coo_t lives in loops/container/formats.hxx (reference include/loops/container/coo.hxx). */
#pragma once
#include <loops/container/formats.hxx>
