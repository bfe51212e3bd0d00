// Build shim (test infrastructure) for the legacy range-coder sources: std headers only.
#pragma once
#include <array>
#include <cassert>
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <memory>
#include <string>
