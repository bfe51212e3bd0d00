// Stand-in (standalone build of integration/ only): the few names of the reference's pod5_format/types.h and result.h
// that the per-chunk codec interface uses. Inside the reference tree the real header is found instead.
#pragma once
#include <arrow/buffer.h>
#include <arrow/memory_pool.h>
#include <arrow/result.h>
#include <arrow/status.h>
#include <gsl/gsl-lite.hpp>
#include <cstdint>
#include <memory>
namespace pod5 {
using Status = arrow::Status;
template <typename T> using Result = arrow::Result<T>;
using SampleType = std::int16_t;
}  // namespace pod5
