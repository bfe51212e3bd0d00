// Build shim (test infrastructure): stands in for the reference's pod5_format/types.h so that the
// reference's Cx codec headers compile without Boost/flatbuffers. Declares only what those headers use.
#pragma once
#include <arrow/buffer.h>
#include <arrow/memory_pool.h>
#include <arrow/result.h>
#include <arrow/status.h>
#include <gsl/gsl-lite.hpp>
#include <cstdint>
#include <cstdio>
#include <memory>
namespace pod5 {
using Status = arrow::Status;
template <typename T> using Result = arrow::Result<T>;
using SampleType = std::int16_t;
}  // namespace pod5
