// Stand-in (standalone build of integration/ only): the declarations of pod5_format/pgnano/pgnano.h:13-23.
#pragma once
#include "pod5_format/types.h"
#include "pod5_format/read_table_utils.h"
namespace pgnano {
pod5::Status decompress_signal(gsl::span<std::uint8_t const> const& compressed_bytes, arrow::MemoryPool* pool, gsl::span<std::int16_t> const& destination, pgnano::PGNanoReaderState& state);
arrow::Result<std::shared_ptr<arrow::Buffer>> compress_signal(gsl::span<std::int16_t const> const& samples, arrow::MemoryPool* pool, pod5::ReadData const& read_data, bool is_last_batch);
}  // namespace pgnano
