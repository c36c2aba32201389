/** @file ell.hxx  ell_t lives in loops/container/formats.hxx (reference include/loops/container/ell.hxx). */
#pragma once
#include <loops/container/formats.hxx>
