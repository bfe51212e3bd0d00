// Build shim (test infrastructure): intentionally empty.
#pragma once
