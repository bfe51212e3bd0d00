// Build shim (test infrastructure): minimal ReadData / reader+writer state types. Every Cx codec
// ignores these parameters.
#pragma once
#include <cstdint>
namespace pod5 { struct ReadData { std::int16_t pore_type = 0; }; }
namespace pgnano { struct PGNanoReaderState {}; struct PGNanoWriterState {}; }
