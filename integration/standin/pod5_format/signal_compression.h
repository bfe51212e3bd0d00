// Stand-in (standalone build of integration/ only): the declarations of pod5_format/signal_compression.h:17-36.
#pragma once
#include "pod5_format/types.h"
namespace pod5 {
std::size_t compressed_signal_max_size(std::size_t sample_count);
arrow::Result<std::size_t> compress_signal(gsl::span<SampleType const> const& samples, arrow::MemoryPool* pool, gsl::span<std::uint8_t> const& destination);
arrow::Result<std::shared_ptr<arrow::Buffer>> compress_signal(gsl::span<SampleType const> const& samples, arrow::MemoryPool* pool);
arrow::Result<std::shared_ptr<arrow::Buffer>> decompress_signal(gsl::span<std::uint8_t const> const& compressed_bytes, std::uint32_t samples_count, arrow::MemoryPool* pool);
arrow::Status decompress_signal(gsl::span<std::uint8_t const> const& compressed_bytes, arrow::MemoryPool* pool, gsl::span<std::int16_t> const& destination);
}  // namespace pod5
