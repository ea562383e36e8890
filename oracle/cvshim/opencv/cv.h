// ORACLE (test infrastructure): forwards to the single shim header.
#pragma once
#include "opencv2/core/core.hpp"
