/** @file dia.hxx  dia_t lives in loops/container/formats.hxx (reference include/loops/container/dia.hxx). */
#pragma once
#include <loops/container/formats.hxx>
