// rnsc_signal_compression.cpp -- the reference's C++ per-chunk codec functions, same signatures, on librnsc.
//
// This is the file a maintainer of tomas-gr/RawNanoporeSignalCompression drops into pod5/c++/pod5_format/ IN PLACE OF
//   signal_compression.cpp   (pod5::compressed_signal_max_size :30-35, compress_signal :37-94, decompress_signal :96-141)
//   pgnano/pgnano.cpp        (pgnano::compress_signal :59-96, pgnano::decompress_signal :98-126)
// so that SignalTableWriter::add_signal (signal_table_writer.cpp:98,107), SignalTableRecordBatch::extract_signal_row
// (signal_table_reader.cpp:133,138), the C API (c_api.cpp:1199,1236,1270) and the pybind module (api.h:518,528) run the
// B200 codec with no other change. Same conventions as the originals: inputs are borrowed, outputs go to the caller's
// span or to a pool-allocated arrow::Buffer, errors are arrow::Status::Invalid with the reference's messages, nothing
// throws. Every call runs on the calling thread's own librnsc context (rnsc_thread_context), so AsyncSignalLoader's
// workers (async_signal_loader.cpp:176-220) decode concurrently -- the reference's own Cx code is not thread-safe
// (unsynchronised globals, C5.hpp:27-48). The pgnano variant, compile-time in the reference (pgnano.cpp:1), is
// rnsc_set_pgnano_variant()'s.
#include "pod5_format/signal_compression.h"
#include "pod5_format/pgnano/pgnano.h"

#include "rnsc.h"

#include <arrow/buffer.h>

namespace {

arrow::Status to_status(int r)
{
    if (r == RNSC_OK) return arrow::Status::OK();
    const char* last = rnsc_last_error();
    return arrow::Status::Invalid((last && *last) ? last : rnsc_status_string(r));
}

arrow::Result<rnsc_ctx*> context()
{
    rnsc_ctx* c = rnsc_thread_context();
    if (!c) return arrow::Status::Invalid(rnsc_last_error());
    return c;
}

arrow::Result<std::shared_ptr<arrow::Buffer>> compress_to_buffer(int codec, gsl::span<std::int16_t const> const& samples, arrow::MemoryPool* pool)
{
    ARROW_ASSIGN_OR_RAISE(rnsc_ctx * c, context());
    std::size_t size = rnsc_max_compressed_size(codec, samples.size());      // VBZ: ZSTD_compressBound(ceil(n/8)+2n); Cx: max(2n+26, 1024) (compressor.h:39-45)
    ARROW_ASSIGN_OR_RAISE(std::shared_ptr<arrow::ResizableBuffer> out, arrow::AllocateResizableBuffer(size, pool));
    ARROW_RETURN_NOT_OK(to_status(rnsc_compress_chunk(c, codec, samples.data(), samples.size(), out->mutable_data(), &size)));
    ARROW_RETURN_NOT_OK(out->Resize(size));
    return out;
}

}  // namespace

namespace pod5 {

std::size_t compressed_signal_max_size(std::size_t sample_count) { return rnsc_max_compressed_size(RNSC_VBZ, sample_count); }

arrow::Result<std::size_t> compress_signal(gsl::span<SampleType const> const& samples, arrow::MemoryPool*, gsl::span<std::uint8_t> const& destination)
{
    ARROW_ASSIGN_OR_RAISE(rnsc_ctx * c, context());
    std::size_t size = destination.size();
    ARROW_RETURN_NOT_OK(to_status(rnsc_compress_chunk(c, RNSC_VBZ, samples.data(), samples.size(), destination.data(), &size)));
    return size;
}

arrow::Result<std::shared_ptr<arrow::Buffer>> compress_signal(gsl::span<SampleType const> const& samples, arrow::MemoryPool* pool)
{
    return compress_to_buffer(RNSC_VBZ, samples, pool);
}

arrow::Status decompress_signal(gsl::span<std::uint8_t const> const& compressed_bytes, arrow::MemoryPool*, gsl::span<std::int16_t> const& destination)
{
    ARROW_ASSIGN_OR_RAISE(rnsc_ctx * c, context());
    return to_status(rnsc_decompress_chunk(c, RNSC_VBZ, compressed_bytes.data(), compressed_bytes.size(), destination.size(), destination.data()));
}

arrow::Result<std::shared_ptr<arrow::Buffer>> decompress_signal(gsl::span<std::uint8_t const> const& compressed_bytes, std::uint32_t samples_count, arrow::MemoryPool* pool)
{
    ARROW_ASSIGN_OR_RAISE(std::shared_ptr<arrow::ResizableBuffer> out, arrow::AllocateResizableBuffer(samples_count * sizeof(SampleType), pool));
    auto dst = gsl::make_span(reinterpret_cast<std::int16_t*>(out->mutable_data()), samples_count);
    ARROW_RETURN_NOT_OK(decompress_signal(compressed_bytes, pool, dst));
    return out;
}

}  // namespace pod5

namespace pgnano {

arrow::Result<std::shared_ptr<arrow::Buffer>> compress_signal(gsl::span<std::int16_t const> const& samples, arrow::MemoryPool* pool,
                                                              pod5::ReadData const&, bool)      // both ignored by every Cx codec (C5.hpp:282-287)
{
    return compress_to_buffer(rnsc_get_pgnano_variant(), samples, pool);
}

pod5::Status decompress_signal(gsl::span<std::uint8_t const> const& compressed_bytes, arrow::MemoryPool*, gsl::span<std::int16_t> const& destination,
                               pgnano::PGNanoReaderState&)                                     // ignored (C5.hpp:477-481)
{
    ARROW_ASSIGN_OR_RAISE(rnsc_ctx * c, context());
    return to_status(rnsc_decompress_chunk(c, rnsc_get_pgnano_variant(), compressed_bytes.data(), compressed_bytes.size(), destination.size(), destination.data()));
}

}  // namespace pgnano
