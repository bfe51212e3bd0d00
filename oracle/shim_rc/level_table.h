// Build shim (test infrastructure).
#pragma once
namespace pgnano { class LevelTable {}; }
