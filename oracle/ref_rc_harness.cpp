// oracle/ref_rc_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Wraps the reference's legacy pgnano range coder (pod5_format/pgnano/{compressor,decompressor}.cpp,
// commonLib/{model,codecs}.cpp, third_party/fqzcomp/{clr,simple_model}.h), compiled where it lies under
// /root/reference by oracle/Makefile. A fresh Compressor/Decompressor per chunk, as benchmark.cpp:203
// and tests/compression_decompression.cpp:76-113 do.
#include "compressor.h"
#include "decompressor.h"
#include "pore_type_server.h"
#include "pgnano_writer_state.h"
#include "pgnano_reader_state.h"
#include <vector>

extern "C" {

size_t ref_rc_max_size(size_t n) { return pgnano::Compressor::compressed_signal_max_size(n, pod5::ReadData{}); }

// pore_desc: string registered for pore index 0 ("R10.4.1" in benchmark.cpp:207 -> header byte 0).
size_t ref_rc_encode(const int16_t* in, size_t n, uint8_t* out, const char* pore_desc)
{
    pgnano::Compressor c;
    pgnano::PGNanoWriterState ws;
    ws.m_pore_type_server->put_pore_type(0, pore_desc);
    pod5::ReadData rd; rd.pore_type = 0;
    // The reference's raw-fallback branch (compressor.cpp:60-65) memcpy's INTO the caller's samples;
    // work on a private copy and report that branch as 0 ("would fall back"), like oracle/rnsc_oracle.c.
    std::vector<int16_t> tmp(in, in + n);
    size_t r = c.compress(rd, n, tmp.data(), out, ws);
    return out[0] ? 0 : r;
}

void ref_rc_decode(const uint8_t* in, size_t clen, int16_t* out)
{
    pgnano::Decompressor d;
    pgnano::PGNanoReaderState rs;
    d.decompress(in, out, clen, rs);
}

}  // extern "C"
