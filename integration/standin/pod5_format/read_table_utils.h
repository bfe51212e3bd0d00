// Stand-in (standalone build of integration/ only) for the parameter types every Cx codec ignores
// (pod5_format/read_table_utils.h ReadData; pgnano/pgnano_reader_state.h PGNanoReaderState).
#pragma once
#include <cstdint>
namespace pod5 { struct ReadData { std::int16_t pore_type = 0; }; }
namespace pgnano { class PGNanoReaderState {}; }
