// Build shim (test infrastructure): ReadData with the one field the range coder's header stamps.
#pragma once
#include <cstdint>
namespace pod5 { using PoreDictionaryIndex = std::int16_t; struct ReadData { PoreDictionaryIndex pore_type = 0; }; }
